# DiffusionMCMCB200.jl -- the reference-side binding of libdmcmc_b200.so (include/dmcmc.h).
#
# NOT EXECUTED IN THIS REPOSITORY: Julia is not installed in the build image (SURVEY.md §0).  It is the
# binding a DiffusionMCMC.jl maintainer would add; the Python package `diffusionmcmc.jl_b200/` mirrors
# exactly this control flow over the same symbols and is what the tests drive, and tests/abi_smoke.c
# drives the same sequence of calls from plain C.
#
# The backend tag `B200Backend` takes the place of `DiffusionMCMCBackend()` in
#     MCMC(updates; backend=DiffusionMCMCB200.B200Backend())
# ExtensibleMCMC's `run!`, schedule, callbacks, parameter chain and chain statistics stay as they are; the
# methods below are the ones DiffusionMCMC.jl defines for its own backend (file:line of each is cited), with
# the GuidedProposals.jl / DiffusionDefinition.jl arithmetic replaced by `ccall`s.
module DiffusionMCMCB200

using ExtensibleMCMC, ObservationSchemes, DiffusionMCMC, OrderedCollections
const eMCMC = ExtensibleMCMC
const OBS = ObservationSchemes
const DM = DiffusionMCMC

const LIB = get(ENV, "DMCMC_B200_LIB", "libdmcmc_b200.so")
const ABI_VERSION = Int32(5)

struct B200Backend <: eMCMC.MCMCBackend end

# dmcmc_config (include/dmcmc.h): 56 bytes, same field order
struct Config
    abi_version::Int32; model::Int32; d::Int32; n_chains::Int32
    chain_offset::Int32; interval_offset::Int32; device::Int32; seed::UInt64; pin_eps::Float64; store_noise::Int32
end

# the opaque context plus what the host side needs to translate the reference's 1-based containers
mutable struct Ctx
    h::Ptr{Cvoid}
    rec_nobs::Vector{Int32}     # observations per recording
    n_steps::Vector{Int32}      # Euler steps per interval (all recordings concatenated)
    d::Int
    n_chains::Int
    num_updates::Int
end

last_error(h) = unsafe_string(ccall((:dmcmc_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
check(ctx::Ctx, rc) = rc == 0 || error(last_error(ctx.h))

function Ctx(model, d, n_chains; device=0, seed=UInt64(1), chain_offset=0, pin_eps=1e-4)
    cfg = Ref(Config(ABI_VERSION, model, d, n_chains, chain_offset, 0, device, seed, pin_eps, 1))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:dmcmc_create, LIB), Cint, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, h)
    rc == 0 || error(last_error(C_NULL))
    ctx = Ctx(h[], Int32[], Int32[], d, n_chains, 1)
    finalizer(c -> ccall((:dmcmc_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.h), ctx)
end

# ---- DiffusionGlobalSubworkspace's constructor inputs (src/workspaces.jl:55-71, :164-170) -------------------
function upload_problem!(ctx::Ctx, data::AllObservations, tts, θ, x0s)
    ctx.rec_nobs = Int32[length(r.obs) for r in data.recordings]
    ctx.n_steps  = Int32[length(tt) - 1 for rec_tts in tts for tt in rec_tts]
    t            = Float64[s for rec_tts in tts for tt in rec_tts for s in tt]
    check(ctx, ccall((:dmcmc_set_grid, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
                     ctx.h, length(ctx.rec_nobs), ctx.rec_nobs, ctx.n_steps, t))
    obs = [o for r in data.recordings for o in r.obs]
    p = length(obs[1].obs)
    L = Float64[obs[j].L[a, b] for j in eachindex(obs) for a in 1:p for b in 1:size(obs[j].L, 2)]
    Σ = Float64[obs[j].Σ[a, b] for j in eachindex(obs) for a in 1:p for b in 1:p]
    # LinearGsnObs carries a mean μ (v ~ N(Lx + μ, Σ)); the ABI takes v - μ (documented at dmcmc_set_obs)
    v = Float64[obs[j].obs[a] - obs[j].μ[a] for j in eachindex(obs) for a in 1:p]
    check(ctx, ccall((:dmcmc_set_obs, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                     ctx.h, p, L, Σ, v))
    check(ctx, ccall((:dmcmc_set_start, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, x0s))  # [n_chains][n_recordings][d]
    check(ctx, ccall((:dmcmc_set_theta, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32), ctx.h, θ, length(θ)))
end

# ---- block layouts: ((rec_idx, obs_range), ...) 1-based (src/workspaces.jl:242-261; src/schedule.jl:97-100) ----
function layout_arrays(ctx::Ctx, layout)
    offs = cumsum([0; ctx.rec_nobs[1:end-1]])
    i1 = Int32[offs[r] + first(rg) - 1 for (r, rg) in layout]
    i2 = Int32[offs[r] + last(rg) - 1 for (r, rg) in layout]
    pinned = Int32[last(rg) < ctx.rec_nobs[r] for (r, rg) in layout]
    i1, i2, pinned
end

# change_laws_for_blocking! + recompute_loglikhd! (called but undefined in the reference, src/workspaces.jl:469-470)
function relayout!(ctx::Ctx, layout)
    i1, i2, pinned = layout_arrays(ctx, layout)
    check(ctx, ccall((:dmcmc_relayout, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                     ctx.h, length(i1), i1, i2, pinned))
end

# BlockingChequerboardSwitch(block_half_len) -> layout tuple (the `update_layout` the reference leaves as a TODO)
function chequerboard_layout(ctx::Ctx, half_len, pattern)
    J = sum(ctx.rec_nobs)
    i1, i2, pin = zeros(Int32, J), zeros(Int32, J), zeros(Int32, J)
    nb = ccall((:dmcmc_chequerboard_layout, LIB), Cint,
               (Int32, Ptr{Int32}, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
               length(ctx.rec_nobs), ctx.rec_nobs, half_len, pattern, i1, i2, pin)
    offs = cumsum([0; ctx.rec_nobs])
    rec_of(j) = findlast(o -> o <= j, offs[1:end-1])
    Tuple((rec_of(i1[b]), (i1[b] - offs[rec_of(i1[b])] + 1):(i2[b] - offs[rec_of(i1[b])] + 1)) for b in 1:nb)
end

# ---- global workspace: the reference's fields (src/workspaces.jl:119-127) with the device context in place of the
#      two DiffusionGlobalSubworkspaces (sub_ws_diff, sub_ws_diff°: both live on the device) -------------------------
struct B200GlobalWorkspace{T,SW,TD,TX} <: eMCMC.GlobalWorkspace{T}
    sub_ws::SW
    ctx::Ctx
    data::TD
    stats::eMCMC.GenericChainStats{T}
    XX_buffer::TX              # Vector of glued trajectories per recording: (t = ..., x = Matrix [points][d])
    pnames::Vector{Symbol}
end

# eMCMC.init_global_workspace (src/workspaces.jl:129-191)
function eMCMC.init_global_workspace(backend::B200Backend, num_mcmc_steps, updates::Vector{<:eMCMC.MCMCUpdate},
                                     data, θinit::Nothing; kwargs...)
    eMCMC.init_global_workspace(backend, num_mcmc_steps, updates, data, OrderedDict(:none => -Inf); kwargs...)
end

function eMCMC.init_global_workspace(::B200Backend, num_mcmc_steps, updates::Vector{<:eMCMC.MCMCUpdate},
                                     data, θinit::OrderedDict{Symbol,T}; kwargs...) where T
    dkwargs = Dict(kwargs)
    M, NU = num_mcmc_steps, length(updates)
    _θ, _pnames = collect(values(θinit)), collect(keys(θinit))
    sub_ws = eMCMC.StandardGlobalSubworkspace(M, NU, data, _θ)
    tts = OBS.setup_time_grids(data, get(dkwargs, :dt, 0.01), get(dkwargs, :τ, identity),
                               get(dkwargs, :Ttime, Float64), get(dkwargs, :tt, nothing))      # src/workspaces.jl:164-170
    C = get(dkwargs, :n_chains, 1)
    d = length(data.recordings[1].x0_prior.y)                     # KnownStartingPt
    ctx = Ctx(get(dkwargs, :model, 0), d, C; device=get(dkwargs, :device, 0), seed=UInt64(get(dkwargs, :seed, 1)))
    ctx.num_updates = NU
    x0s = Float64[r.x0_prior.y[i] for _ in 1:C for r in data.recordings for i in 1:d]
    upload_problem!(ctx, data, tts, Float64.(_θ), x0s)
    # no-blocking layout (src/schedule.jl:97-100), tables (build_guid_prop, :60), paths (init_paths!, :80-92), ll (init_ll!, :425-431)
    relayout_none = Tuple((i, 1:length(r.obs)) for (i, r) in enumerate(data.recordings))
    i1, i2, pinned = layout_arrays(ctx, relayout_none)
    check(ctx, ccall((:dmcmc_set_layout, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                     ctx.h, length(i1), i1, i2, pinned))
    check(ctx, ccall((:dmcmc_backward_filter, LIB), Cint, (Ptr{Cvoid}, Int32), ctx.h, 0))
    check(ctx, ccall((:dmcmc_init_paths, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, C_NULL))   # runs init_ll! as well
    # PathSavingBuffer (src/path_saving_buffer.jl:29-44): ring of N glued paths per recording
    N = get(dkwargs, :path_buffer_size, 1)
    npts(r) = sum(ctx.n_steps[(sum(ctx.rec_nobs[1:r-1]) + 1):sum(ctx.rec_nobs[1:r])]) + 1
    buf = (XX = [[zeros(Float64, d, npts(r)) for r in eachindex(data.recordings)] for _ in 1:N], next = Ref(1))
    B200GlobalWorkspace{T,typeof(sub_ws),typeof(data),typeof(buf)}(
        sub_ws, ctx, data, eMCMC.GenericChainStats(_θ, NU, M), buf, _pnames)
end

# ---- local workspaces (src/workspaces.jl:339-423, :448-458): histories only; ll, ll° and the paths live on the device
struct B200LocalWorkspace{T,SW} <: eMCMC.LocalWorkspace{T}
    sub_ws::SW
    sub_ws°::SW
    layout::Any
    ll::Vector{Float64}                      # [n_chains * nblk], chain-major like the ABI
    ll°::Vector{Float64}
    ll_history::Vector{Vector{Float64}}
    ll°_history::Vector{Vector{Float64}}
    acceptance_history::Vector{Vector{Bool}}
    critical::Bool
end

function eMCMC.create_workspaces(::B200Backend, mcmc::eMCMC.MCMC)
    gws = mcmc.workspace
    C = gws.ctx.n_chains
    map(1:length(mcmc.updates)) do i
        u, layout, M = mcmc.updates[i], mcmc.schedule.extra_info.blocking_layout[i], mcmc.schedule.num_mcmc_steps
        isparam = typeof(u) <: eMCMC.MCMCParamUpdate
        _state = isparam ? copy(eMCMC.state(gws, u)) : Float64[]
        sub_ws = eMCMC.StandardLocalSubworkspace(_state, M)
        n = C * length(layout)
        isparam || DM.init_update!(u, layout)                                    # src/updates.jl:53-74
        B200LocalWorkspace{eltype(eMCMC.state(gws)),typeof(sub_ws)}(
            sub_ws, deepcopy(sub_ws), layout, fill(-Inf, n), fill(-Inf, n),
            [zeros(n) for _ in 1:M], [zeros(n) for _ in 1:M], [fill(false, n) for _ in 1:M], true)
    end
end

# update_workspaces! (src/workspaces.jl:460-476): a layout change re-conditions the laws, recovers the noise and
# recomputes ll on the device; with an unchanged layout ll is already there (it never left the device)
function eMCMC.update_workspaces!(updt, gws::B200GlobalWorkspace, lws::B200LocalWorkspace, step, prev_ws)
    all(step.same_layout) || relayout!(gws.ctx, lws.layout)
    lws, updt
end

# ---- eMCMC.update!(::PathImputation)  (src/run!_alterations.jl:5-26) + register (:36-75) + swap (:81-91) ----------
function eMCMC.update!(update::DM.PathImputation, gws::B200GlobalWorkspace, lws::B200LocalWorkspace, step)
    ctx = gws.ctx
    ρ = Float64[r for ρs in update.ρs for r in ρs]                                # one value per interval (src/updates.jl:53-74)
    check(ctx, ccall((:dmcmc_set_rho, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, ρ))
    n = length(lws.ll)                                                            # n_chains * nblk
    acc = Vector{UInt8}(undef, n)
    iter = UInt32(step.mcmciter * ctx.num_updates + step.pidx)                    # distinct Philox counters per update
    check(ctx, ccall((:dmcmc_impute, LIB), Cint,
                     (Ptr{Cvoid}, UInt32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
                     ctx.h, iter, C_NULL, C_NULL, lws.ll°, lws.ll, acc))          # the swap happened on the device
    lws.acceptance_history[step.mcmciter] .= acc .!= 0                            # src/run!_alterations.jl:65-74
    lws.ll°_history[step.mcmciter] .= lws.ll°
    lws.ll_history[step.mcmciter] .= lws.ll
    step.mcmciter % 100 == 0 && step.pidx == 1 && save_path!(gws)                 # src/run!_alterations.jl:51-54
    for (i, a) in enumerate(update.adpt isa DM.AdaptationMultiPathImputation ? update.adpt.v : ())
        # one adaptation per block (src/adaptation.jl:128-137): chains of a block pool their decisions
        nb = length(lws.layout)
        for c in 0:ctx.n_chains-1
            eMCMC.register!(update, a, acc[c * nb + i] != 0, nothing)
        end
    end
end

# ---- eMCMC.set_proposal!(::MCMCParamUpdate)  (src/run!_alterations.jl:177-211), compute_ll! (:258-265) -----------
function eMCMC.set_proposal!(updt::eMCMC.MCMCParamUpdate, gws::B200GlobalWorkspace, lws::B200LocalWorkspace, step)
    eMCMC._set_local_state°(updt, gws, lws, step)
    θ° = Float64[eMCMC.state°(lws)...]
    critical = lws.critical || !all(step.same_layout)
    ok = Vector{UInt8}(undef, gws.ctx.n_chains)                                   # out_success is [n_chains]
    check(gws.ctx, ccall((:dmcmc_param_loglik, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Ptr{UInt8}),
                         gws.ctx.h, θ°, length(θ°), critical, lws.ll°, ok))
    all(ok .!= 0) || fill!(lws.ll°, -Inf)
    lws.ll°_history[step.mcmciter] .= lws.ll°
end

# register_accept_reject_results! for parameter updates (src/run!_alterations.jl:296-326): swap_laws_and_paths! (:328-334)
function eMCMC.register_accept_reject_results!(accepted::Bool, updt::eMCMC.MCMCParamUpdate, gws::B200GlobalWorkspace,
                                               lws::B200LocalWorkspace, step)
    check(gws.ctx, ccall((:dmcmc_param_commit, LIB), Cint, (Ptr{Cvoid}, Int32), gws.ctx.h, accepted))
    accepted && (lws.ll .= lws.ll°)
    lws.ll_history[step.mcmciter] .= lws.ll
    accepted && eMCMC.set_chain_param!(accepted, updt, gws, lws, step)            # src/run!_alterations.jl:306
end

# accessors the reference imports (src/workspaces.jl:478-483)
DM.ll(ws::B200LocalWorkspace) = ws.ll
DM.ll°(ws::B200LocalWorkspace) = ws.ll°
DM.ll(ws::B200LocalWorkspace, i::Int) = ws.ll_history[i]
DM.ll°(ws::B200LocalWorkspace, i::Int) = ws.ll°_history[i]
eMCMC.accepted(ws::B200LocalWorkspace, i::Int) = ws.acceptance_history[i]

# ---- save! (src/path_saving_buffer.jl:52-64): glued accepted path of chain 1 of every recording into the ring -------
function save_path!(gws::B200GlobalWorkspace)
    buf, ctx = gws.XX_buffer, gws.ctx
    slot = buf.XX[buf.next[]]
    for rec in eachindex(slot)
        # the ABI fills x_out [points][d] row-major = a Julia Matrix [d][points] column-major
        check(ctx, ccall((:dmcmc_download_path, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}),
                         ctx.h, 0, rec - 1, slot[rec]))
    end
    buf.next[] = mod1(buf.next[] + 1, length(buf.XX))
end

# ---- one long recording over the GPUs of a box (no counterpart in the reference; SURVEY 8e, INTEGRATION.md "Multi-GPU") ----
# One process and one context per GPU; `id` = 128 bytes from comm_unique_id() on one rank, handed to the others by the
# host's own means.  The collectives live in the library: nothing here needs CUDA.jl, NCCL.jl or MPI.
comm_unique_id() = (id = Vector{UInt8}(undef, 128); ccall((:dmcmc_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id) == 0 || error("NCCL not available"); id)
comm_init!(ctx, id::Vector{UInt8}, rank::Integer, world::Integer) =
    check(ctx, ccall((:dmcmc_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), ctx.h, id, rank, world))
# end-point exchange over NVLink peer memory instead of ncclSend / ncclRecv (collective); true when every rank has it
function comm_enable_p2p!(ctx, mailbox_bytes::Integer)
    check(ctx, ccall((:dmcmc_comm_enable_p2p, LIB), Cint, (Ptr{Cvoid}, Int64), ctx.h, mailbox_bytes))
    ccall((:dmcmc_comm_p2p_enabled, LIB), Cint, (Ptr{Cvoid},), ctx.h) != 0
end
halo_exchange!(ctx, after_pattern::Integer, n_first::Integer, n_halo::Integer, left_edge_steps::Integer) =
    check(ctx, ccall((:dmcmc_halo_exchange, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Int32, Int32),
                     ctx.h, after_pattern, n_first, n_halo, left_edge_steps))
# the sweep and the exchange that follows it in one call (fused into the sweep kernels when peer memory is on)
impute_exchange!(ctx, iter::Integer, pattern::Integer, n_first::Integer, n_halo::Integer, left_edge_steps::Integer) =
    check(ctx, ccall((:dmcmc_impute_exchange_async, LIB), Cint, (Ptr{Cvoid}, UInt32, Int32, Int32, Int32, Int32),
                     ctx.h, iter, pattern, n_first, n_halo, left_edge_steps))
comm_destroy!(ctx) = check(ctx, ccall((:dmcmc_comm_destroy, LIB), Cint, (Ptr{Cvoid},), ctx.h))

end # module
