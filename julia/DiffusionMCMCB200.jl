# DiffusionMCMCB200.jl -- the reference-side binding of libdmcmc_b200.so (include/dmcmc.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image (SURVEY.md §0).
# It is the stub a DiffusionMCMC.jl maintainer would add: the plugin's own methods for the imputation
# path (src/run!_alterations.jl, src/workspaces.jl) forward to `ccall`s instead of calling
# GuidedProposals.jl / DiffusionDefinition.jl.  Everything else (ExtensibleMCMC's run!, schedule,
# callbacks, parameter chain, adaptation) stays as it is.  The Python package
# `diffusionmcmc.jl_b200/` mirrors exactly this control flow and is what the tests drive.
module DiffusionMCMCB200

using ExtensibleMCMC, ObservationSchemes, StaticArrays
const eMCMC = ExtensibleMCMC
const OBS = ObservationSchemes

const LIB = get(ENV, "DMCMC_B200_LIB", "libdmcmc_b200.so")
const ABI_VERSION = Int32(4)

struct Config
    abi_version::Int32; model::Int32; d::Int32; n_chains::Int32
    chain_offset::Int32; interval_offset::Int32; device::Int32; seed::UInt64; pin_eps::Float64; store_noise::Int32
end

mutable struct Ctx
    h::Ptr{Cvoid}
end

check(ctx, rc) = rc == 0 || error(unsafe_string(ccall((:dmcmc_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h)))

function Ctx(model, d, n_chains; device=0, seed=0x1, chain_offset=0, pin_eps=1e-4)
    cfg = Ref(Config(ABI_VERSION, model, d, n_chains, chain_offset, 0, device, seed, pin_eps, 1))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:dmcmc_create, LIB), Cint, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, h)
    rc == 0 || error(unsafe_string(ccall((:dmcmc_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    ctx = Ctx(h[])
    finalizer(c -> ccall((:dmcmc_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.h), ctx)
end

# ---- replaces DiffusionGlobalSubworkspace's constructor (src/workspaces.jl:55-71, :164-170) ----
function upload_problem!(ctx, data::AllObservations, tts, θ, x0s)
    rec_nobs = Int32[length(r.obs) for r in data.recordings]
    n_steps  = Int32[length(tt) - 1 for rec_tts in tts for tt in rec_tts]
    t        = Float64[s for rec_tts in tts for tt in rec_tts for s in tt]
    check(ctx, ccall((:dmcmc_set_grid, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
                     ctx.h, length(rec_nobs), rec_nobs, n_steps, t))
    obs = [o for r in data.recordings for o in r.obs]
    p = length(obs[1].obs)
    L = Float64[obs[j].L[a, b] for j in eachindex(obs) for a in 1:p for b in 1:size(obs[j].L, 2)]
    Σ = Float64[obs[j].Σ[a, b] for j in eachindex(obs) for a in 1:p for b in 1:p]
    v = Float64[obs[j].obs[a] - obs[j].μ[a] for j in eachindex(obs) for a in 1:p]
    check(ctx, ccall((:dmcmc_set_obs, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                     ctx.h, p, L, Σ, v))
    check(ctx, ccall((:dmcmc_set_start, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, x0s))
    check(ctx, ccall((:dmcmc_set_theta, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32), ctx.h, θ, length(θ)))
end

# ---- block layouts (src/workspaces.jl:242-261; src/schedule.jl:97-100; src/blocking.jl:7-9) ----
function set_layout!(ctx, layout)           # layout = ((rec_idx, obs_range), ...), 1-based like the reference
    offs = cumsum([0; ctx_rec_nobs(ctx)[1:end-1]])
    i1 = Int32[offs[r] + first(rg) - 1 for (r, rg) in layout]
    i2 = Int32[offs[r] + last(rg) - 1 for (r, rg) in layout]
    pinned = Int32[last(rg) < ctx_rec_nobs(ctx)[r] for (r, rg) in layout]
    check(ctx, ccall((:dmcmc_relayout, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                     ctx.h, length(i1), i1, i2, pinned))
end

# ---- eMCMC.update!(::PathImputation, ...)  (src/run!_alterations.jl:5-26, :36-75) ----
function eMCMC.update!(update::PathImputation, gws::DiffusionGlobalWorkspace, lws::DiffusionLocalWorkspace, step)
    ctx = gws.ctx
    ρ = Float64[r for ρs in update.ρs for r in ρs]                        # src/updates.jl:53-74
    check(ctx, ccall((:dmcmc_set_rho, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, ρ))
    nblk = length(lws.sub_ws_diff.ll)
    ll°, ll, acc = Vector{Float64}(undef, nblk), Vector{Float64}(undef, nblk), Vector{UInt8}(undef, nblk)
    check(ctx, ccall((:dmcmc_impute, LIB), Cint,
                     (Ptr{Cvoid}, UInt32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
                     ctx.h, step.mcmciter, C_NULL, C_NULL, ll°, ll, acc))
    for i in 1:nblk                                                       # register (src/run!_alterations.jl:57-75)
        accepted(lws, step.mcmciter)[i] = acc[i] != 0
        DiffusionMCMC.ll°(lws, step.mcmciter)[i] = ll°[i]
        DiffusionMCMC.ll(lws)[i] = ll[i]
        DiffusionMCMC.ll(lws, step.mcmciter)[i] = ll[i]
    end
    save_path!(gws, step, 1)                                              # src/run!_alterations.jl:51-54
end

# ---- eMCMC.set_proposal!(::MCMCParamUpdate, ...)  (src/run!_alterations.jl:177-211) ----
function eMCMC.set_proposal!(updt::eMCMC.MCMCParamUpdate, gws::DiffusionGlobalWorkspace, lws::DiffusionLocalWorkspace, step)
    eMCMC._set_local_state°(updt, gws, lws, step)
    θ° = Float64[state°(lws)...]
    critical = any(lws.critical_param_change) || !all(step.same_layout)
    nblk = length(lws.sub_ws_diff°.ll)
    ok = Ref{UInt8}(0)
    check(gws.ctx, ccall((:dmcmc_param_loglik, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Ref{UInt8}),
                         gws.ctx.h, θ°, length(θ°), critical, lws.sub_ws_diff°.ll, ok))
    ok[] == 0 && fill!(lws.sub_ws_diff°.ll, -Inf)
end

# swap_laws_and_paths! (src/run!_alterations.jl:328-334)
swap_laws_and_paths!(gws::DiffusionGlobalWorkspace, accepted::Bool) =
    check(gws.ctx, ccall((:dmcmc_param_commit, LIB), Cint, (Ptr{Cvoid}, Int32), gws.ctx.h, accepted))

# save! (src/path_saving_buffer.jl:52-64): glued accepted path of recording `rec`
function download_path!(x::Matrix{Float64}, ctx, chain, rec)
    check(ctx, ccall((:dmcmc_download_path, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}), ctx.h, chain - 1, rec - 1, x))
    x
end

end # module
