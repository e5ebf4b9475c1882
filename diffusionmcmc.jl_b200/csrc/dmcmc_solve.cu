// dmcmc_solve.cu -- the imputation hot path as hand-written CUDA for sm_100a.
//
// One thread owns one (chain, block) unit -- the parallel axis the reference itself names
// ("this loop should be entirely parallelizable", /root/reference src/run!_alterations.jl:15-16)
// times the chain batch axis.  Consecutive lanes are consecutive chains of the same block, so
//   * path/noise traffic is 64 contiguous bytes per lane per 8 steps (256-bit LDG/STG), 2 KB per
//     warp request, and
//   * the backward-filter table record of a step is the same address for the whole warp (one
//     broadcast transaction).
// Inside a unit the Euler recursion is serial (x_{k+1} needs x_k); everything that does not
// depend on x -- accepted-noise loads, Philox/Box-Muller, table records -- is issued a tile ahead.
//
// Reference rows:
//   a2 pCN refresh            forward_guide!(..., rho, ...; Wnr)   src/run!_alterations.jl:18-21
//   a3 guiding term lookup    r = F[k] - H[k] x
//   a4 Euler-Maruyama solve   GP.solve_and_ll!                     src/run!_alterations.jl:206-208
//   a5 Girsanov ll            loglikhd / loglikhd_obs              src/workspaces.jl:429
//   a6 accept                 eMCMC.accept_reject!                 src/run!_alterations.jl:22-24
//   a7 register + swap        swap_paths!                          src/run!_alterations.jl:36-47, :81-91
//   a9 parameter-step solve   set_proposal!                        src/run!_alterations.jl:177-211
#include <cmath>

#include "dmcmc_device.cuh"
#include "dmcmc_internal.h"

namespace dmcmc {

// ---- one guided step: Girsanov integrand G*dt and the drift part (b + a r) dt -----------------
template <class MD, bool PSI>
__device__ __forceinline__ void guided_terms(const Theta &th, const double *__restrict__ rec,
                                             const double *xb, const double *aB,
                                             const double *abeta, const double *aatil,
                                             const double *x, double dt, double *drift_dt,
                                             double &dll) {
    constexpr int D = MD::D;
    using R = Rec<D, PSI>;
    double F[D], r[D], b[D], ar[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double f = rec[R::OFF_F + i];
        if constexpr (PSI) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) s += rec[R::OFF_PSI + i * D + k] * xb[k];
            f += s;
        }
        F[i] = f;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s += rec[R::OFF_H + hidx(D, i, k)] * x[k];
        r[i] = F[i] - s;
    }
    MD::drift(th, x, b);
    double g = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s += aB[i * D + k] * x[k];
        g += (b[i] - (s + abeta[i])) * r[i];
    }
    if constexpr (!MD::CONST_SIGMA) {
        double a[D * D];
        MD::a_full(th, x, a);
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int k = 0; k < D; ++k)
                tr += (a[i * D + k] - aatil[i * D + k]) * (r[k] * r[i] - rec[R::OFF_H + hidx(D, k, i)]);
        g += 0.5 * tr;
    }
    dll = g * dt;
    MD::a_mul(th, x, r, ar);
#pragma unroll
    for (int i = 0; i < D; ++i) drift_dt[i] = (b[i] + ar[i]) * dt;
}

template <int REC>
__device__ __forceinline__ void load_rec(const double *__restrict__ p, double *r) {
#pragma unroll
    for (int i = 0; i < REC / 2; ++i) {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(p) + i);
        r[2 * i] = v.x;
        r[2 * i + 1] = v.y;
    }
}

// end point of interval j (value after its last Euler step) in the accepted container
template <int D>
__device__ __forceinline__ void load_end(const SolveParams &p, int chain, int j, double *x) {
    const int s = p.selX_in[(size_t)j * p.C + chain];
    const int k = p.iv_N[j] - 1;
    const size_t tile = (size_t)p.iv_tile_off[j] + (k >> 3);
    const double *X = p.X[s];
#pragma unroll
    for (int c = 0; c < D; ++c)
        x[c] = X[(((size_t)c * p.TT + tile) * p.C + chain) * TILE + (k & 7)];
}

template <int D>
__device__ __forceinline__ void load_start(const SolveParams &p, int chain, int j, double *x) {
    if (p.iv_first[j]) {
        const int rec = p.iv_rec[j];
#pragma unroll
        for (int c = 0; c < D; ++c) x[c] = p.x0[((size_t)rec * D + c) * p.C + chain];
    } else {
        load_end<D>(p, chain, j - 1, x);
    }
}

template <class MD, bool PSI, int MODE, bool PHILOX>
__global__ void __launch_bounds__(128) solve_kernel(const SolveParams p) {
    constexpr int D = MD::D, M = MD::M;
    using R = Rec<D, PSI>;
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (long long)p.nblk * p.C) return;
    const int blk = (int)(u / p.C), chain = (int)(u - (long long)blk * p.C);
    const size_t uo = (size_t)chain * p.nblk + blk;  // [C][nblk] host layout of ll / histories
    const int i1 = p.blk_i1[blk], i2 = p.blk_i2[blk];

    if (MODE == MODE_IMPUTE && p.unit_mask && !p.unit_mask[uo]) {
        for (int j = i1; j <= i2; ++j) {
            const size_t o = (size_t)j * p.C + chain;
            p.selX_out[o] = p.selX_in[o];
            p.selW_out[o] = p.selW_in[o];
        }
        return;
    }

    const bool pinned = PSI && p.blk_pinned[blk];
    double x[D], xb[D];
    load_start<D>(p, chain, i1, x);
#pragma unroll
    for (int c = 0; c < D; ++c) xb[c] = 0.0;
    if (pinned) load_end<D>(p, chain, i2, xb);

    // log h~(t_a, x_a) = -c - 1/2 x'Hx + F'x at the block's first grid point (loglikhd_obs)
    double llp;
    {
        double rec[R::SIZE];
        load_rec<R::SIZE>(p.table + (size_t)p.iv_tile_off[i1] * TILE * R::SIZE, rec);
        double c = p.blk_c0[blk];
        if (pinned) {
            double gx = 0.0, q = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                gx += p.blk_g[(size_t)blk * D + i] * xb[i];
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) s += p.blk_Qc[(size_t)blk * D * D + i * D + k] * xb[k];
                q += xb[i] * s;
            }
            c += gx + 0.5 * q;
        }
        double xHx = 0.0, Fx = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double f = rec[R::OFF_F + i];
            if constexpr (PSI) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) s += rec[R::OFF_PSI + i * D + k] * xb[k];
                f += s;
            }
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) s += rec[R::OFF_H + hidx(D, i, k)] * x[k];
            xHx += x[i] * s;
            Fx += f * x[i];
        }
        llp = -c - 0.5 * xHx + Fx;
    }

    bool ok = true;
    const size_t cstride = (size_t)p.TT * p.C * TILE;  // component stride in X / W / Z
    const uint32_t gchain = (uint32_t)(chain + p.chain_offset);

    for (int j = i1; j <= i2; ++j) {
        const int N = p.iv_N[j];
        const size_t toff = (size_t)p.iv_tile_off[j];
        const int nt = (N + TILE - 1) >> 3;
        const int sx = p.selX_in[(size_t)j * p.C + chain], sw = p.selW_in[(size_t)j * p.C + chain];
        double aB[D * D], abeta[D], aatil[MD::CONST_SIGMA ? 1 : D * D];
#pragma unroll
        for (int i = 0; i < D * D; ++i) aB[i] = p.auxB[(size_t)j * D * D + i];
#pragma unroll
        for (int i = 0; i < D; ++i) abeta[i] = p.auxbeta[(size_t)j * D + i];
        if constexpr (!MD::CONST_SIGMA) {
#pragma unroll
            for (int i = 0; i < D * D; ++i) aatil[i] = p.auxatil[(size_t)j * D * D + i];
        }
        double lsum = 0.0;

        if constexpr (MODE == MODE_IMPUTE || MODE == MODE_PARAM) {
            const double rho = (MODE == MODE_IMPUTE) ? p.rho[j] : 1.0;
            const double lam = (MODE == MODE_IMPUTE) ? sqrt(1.0 - rho * rho) : 0.0;
            const double *Wa = p.W[sw];
            double *Wp = p.W[1 - sw];
            double *Xp = p.X[1 - sx];
            double dwa[M][TILE], dwn[M][TILE];
            {
                const size_t base = (toff * p.C + chain) * TILE;
#pragma unroll
                for (int w = 0; w < M; ++w) ld_tile(Wa + w * cstride + base, dwa[w]);
            }
            for (int tl = 0; tl < nt; ++tl) {
                const size_t tile = toff + tl;
                const size_t base = (tile * p.C + chain) * TILE;
                if (tl + 1 < nt) {  // accepted noise of the next tile: independent of x, fetch early
#pragma unroll
                    for (int w = 0; w < M; ++w)
                        ld_tile(Wa + w * cstride + base + (size_t)p.C * TILE, dwn[w]);
                }
                double z[M][TILE];
                if constexpr (MODE == MODE_IMPUTE) {
                    if constexpr (PHILOX) {
#pragma unroll
                        for (int w = 0; w < M; ++w)
#pragma unroll
                            for (int q = 0; q < TILE / 2; ++q)
                                philox_normal_pair(p.seed, p.iter, (uint32_t)j, gchain, (uint32_t)w,
                                                   (uint32_t)(tl * (TILE / 2) + q), z[w][2 * q],
                                                   z[w][2 * q + 1]);
                    } else {
#pragma unroll
                        for (int w = 0; w < M; ++w) ld_tile_stream(p.Z + w * cstride + base, z[w]);
                    }
                }
                double xo[D][TILE], dwo[M][TILE];
                const double *trec = p.table + tile * TILE * R::SIZE;
#pragma unroll
                for (int s = 0; s < TILE; ++s) {
                    double rec[R::SIZE];
                    load_rec<R::SIZE>(trec + s * R::SIZE, rec);
                    double dw[M];
#pragma unroll
                    for (int w = 0; w < M; ++w) {
                        if constexpr (MODE == MODE_IMPUTE)
                            dw[w] = lam * (rec[1] * z[w][s]) + rho * dwa[w][s];
                        else
                            dw[w] = dwa[w][s];
                        dwo[w][s] = dw[w];
                    }
                    if (tl * TILE + s < N) {
                        double dr[D], sdw[D], dll;
                        guided_terms<MD, PSI>(p.th, rec, xb, aB, abeta, aatil, x, rec[0], dr, dll);
                        lsum += dll;
                        MD::sigma_mul(p.th, x, dw, sdw);
#pragma unroll
                        for (int c = 0; c < D; ++c) x[c] = x[c] + dr[c] + sdw[c];
                        ok = ok && MD::bound_ok(x);
                    }
#pragma unroll
                    for (int c = 0; c < D; ++c) xo[c][s] = x[c];
                }
#pragma unroll
                for (int c = 0; c < D; ++c) st_tile(Xp + c * cstride + base, xo[c]);
                if (pinned && j == i2 && tl == nt - 1) {  // pin the block's right end to x_b
#pragma unroll
                    for (int c = 0; c < D; ++c) Xp[c * cstride + base + ((N - 1) & 7)] = xb[c];
                }
                if constexpr (MODE == MODE_IMPUTE) {
#pragma unroll
                    for (int w = 0; w < M; ++w) st_tile(Wp + w * cstride + base, dwo[w]);
                }
#pragma unroll
                for (int w = 0; w < M; ++w)
#pragma unroll
                    for (int s = 0; s < TILE; ++s) dwa[w][s] = dwn[w][s];
            }
        } else {
            // MODE_RELAYOUT / MODE_LOGLIK: walk the accepted path, recover dW, accumulate ll
            const double *Xa = p.X[sx];
            double *Wacc = p.W[sw];
            for (int tl = 0; tl < nt; ++tl) {
                const size_t tile = toff + tl;
                const size_t base = (tile * p.C + chain) * TILE;
                double xa[D][TILE], dwo[M][TILE];
#pragma unroll
                for (int c = 0; c < D; ++c) ld_tile(Xa + c * cstride + base, xa[c]);
                const double *trec = p.table + tile * TILE * R::SIZE;
#pragma unroll
                for (int s = 0; s < TILE; ++s) {
                    double dw[M];
#pragma unroll
                    for (int w = 0; w < M; ++w) dw[w] = 0.0;
                    if (tl * TILE + s < N) {
                        double rec[R::SIZE];
                        load_rec<R::SIZE>(trec + s * R::SIZE, rec);
                        double dr[D], res[D], dll;
                        guided_terms<MD, PSI>(p.th, rec, xb, aB, abeta, aatil, x, rec[0], dr, dll);
                        lsum += dll;
#pragma unroll
                        for (int c = 0; c < D; ++c) res[c] = xa[c][s] - x[c] - dr[c];
                        MD::sigma_solve(p.th, x, res, dw);
#pragma unroll
                        for (int c = 0; c < D; ++c) x[c] = xa[c][s];
                    }
#pragma unroll
                    for (int w = 0; w < M; ++w) dwo[w][s] = dw[w];
                }
                if constexpr (MODE == MODE_RELAYOUT) {
#pragma unroll
                    for (int w = 0; w < M; ++w) st_tile(Wacc + w * cstride + base, dwo[w]);
                }
            }
        }
        llp += lsum;
    }
    if (!ok) llp = -INFINITY;

    if constexpr (MODE == MODE_IMPUTE) {
        double llcur = p.ll[uo];
        const double e = PHILOX ? philox_exponential(p.seed, p.iter, (uint32_t)i1, gchain) : p.E[uo];
        // accepted = rand(Exponential(1)) > -(ll deg - ll)   (eMCMC.accept_reject!)
        const bool acc = p.force_accept ? ok : (e > -(llp - llcur));
        for (int j = i1; j <= i2; ++j) {  // swap_paths!: flip the accepted-container selectors
            const size_t o = (size_t)j * p.C + chain;
            p.selX_out[o] = p.selX_in[o] ^ (uint8_t)acc;
            p.selW_out[o] = p.selW_in[o] ^ (uint8_t)acc;
        }
        if (acc) {
            llcur = llp;
            p.ll[uo] = llp;
            if (p.acc_count) atomicAdd(p.acc_count + blk, 1ULL);
        }
        if (p.prop_count && chain == 0) atomicAdd(p.prop_count + blk, (unsigned long long)p.C);
        if (p.out_llprop) p.out_llprop[uo] = llp;
        if (p.out_ll) p.out_ll[uo] = llcur;
        if (p.out_acc) p.out_acc[uo] = (uint8_t)acc;
    } else if constexpr (MODE == MODE_PARAM) {
        if (p.out_llprop) p.out_llprop[uo] = llp;
        if (p.out_acc) p.out_acc[uo] = (uint8_t)ok;
    } else {
        p.ll[uo] = llp;
    }
}

// ---- launch dispatch ---------------------------------------------------------------------------
template <class MD, bool PSI>
static cudaError_t launch_model(int mode, const SolveParams &p, int threads, cudaStream_t s) {
    const long long nunits = (long long)p.nblk * p.C;
    if (nunits == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((nunits + threads - 1) / threads);
    const bool philox = (p.Z == nullptr);
    switch (mode) {
    case MODE_IMPUTE:
        if (philox) solve_kernel<MD, PSI, MODE_IMPUTE, true><<<grid, threads, 0, s>>>(p);
        else solve_kernel<MD, PSI, MODE_IMPUTE, false><<<grid, threads, 0, s>>>(p);
        break;
    case MODE_PARAM: solve_kernel<MD, PSI, MODE_PARAM, false><<<grid, threads, 0, s>>>(p); break;
    case MODE_RELAYOUT: solve_kernel<MD, PSI, MODE_RELAYOUT, false><<<grid, threads, 0, s>>>(p); break;
    case MODE_LOGLIK: solve_kernel<MD, PSI, MODE_LOGLIK, false><<<grid, threads, 0, s>>>(p); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <class MD>
static cudaError_t launch_psi(int mode, const SolveParams &p, int threads, cudaStream_t s) {
    return p.has_psi ? launch_model<MD, true>(mode, p, threads, s)
                     : launch_model<MD, false>(mode, p, threads, s);
}

cudaError_t launch_solve(int model, int mode, const SolveParams &p, int threads, cudaStream_t s) {
    switch (model) {
    case 0: return launch_psi<ModelFHN>(mode, p, threads, s);
    case 1: return launch_psi<ModelLV>(mode, p, threads, s);
    case 2: return launch_psi<ModelL63>(mode, p, threads, s);
    case 4: return launch_psi<ModelLIN2>(mode, p, threads, s);
    case 5: return launch_psi<ModelLVM>(mode, p, threads, s);
    }
    return cudaErrorInvalidValue;
}

// ---- read-out: reference-shaped containers ([C][P][D], per-interval N_j+1 points; cumulative W) --
__global__ void gather_kernel(const GatherParams g) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (long long)g.J * g.C) return;
    const int j = (int)(u / g.C), chain = (int)(u - (long long)j * g.C);
    const int N = g.iv_N[j];
    const size_t cs = (size_t)g.TT * g.C * TILE;
    const size_t o = (size_t)j * g.C + chain;
    const int sx = g.selX[o] ^ g.which, sw = g.selW[o] ^ g.which;
    const size_t pt0 = (size_t)chain * g.P + g.iv_pt_off[j];
    // point 0 of the container
    for (int c = 0; c < g.D; ++c) {
        double v;
        if (g.iv_first[j]) {
            v = g.x0[((size_t)g.iv_rec[j] * g.D + c) * g.C + chain];
        } else {
            // proposal containers chain from the proposal of the previous interval of the same
            // block, and from the accepted path at the block's first interval
            const int jp = j - 1;
            const bool same_blk = g.which && g.iv_blk[jp] == g.iv_blk[j];
            const int sp = g.selX[(size_t)jp * g.C + chain] ^ (same_blk ? 1 : 0);
            const int k = g.iv_N[jp] - 1;
            const size_t tile = (size_t)g.iv_tile_off[jp] + (k >> 3);
            v = g.X[sp][(((size_t)c * g.TT + tile) * g.C + chain) * TILE + (k & 7)];
        }
        g.outX[(pt0)*g.D + c] = v;
    }
    for (int w = 0; w < g.M; ++w) g.outW[pt0 * g.M + w] = 0.0;
    for (int k = 0; k < N; ++k) {
        const size_t tile = (size_t)g.iv_tile_off[j] + (k >> 3);
        const size_t e = (tile * g.C + chain) * TILE + (k & 7);
        for (int c = 0; c < g.D; ++c) g.outX[(pt0 + k + 1) * g.D + c] = g.X[sx][c * cs + e];
        for (int w = 0; w < g.M; ++w)
            g.outW[(pt0 + k + 1) * g.M + w] = g.outW[(pt0 + k) * g.M + w] + g.W[sw][w * cs + e];
    }
}

cudaError_t launch_gather(const GatherParams &g, cudaStream_t s) {
    const long long n = (long long)g.J * g.C;
    if (n == 0) return cudaSuccess;
    gather_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g);
    return cudaGetLastError();
}

}  // namespace dmcmc
