// dmcmc_solve_big.cu -- the imputation hot path for large state dimension (Lorenz-96, d = 40).
//
// One WARP owns four (chain, block) units of the same layout block: lane l works for chain l / 8 and
// owns the five state components g, g+8, .., g+32 (g = l % 8), so the warp's 160 rows sit on 32
// lanes with none idle.  The per-step work is three dense 40 x 40 mat-vecs (H x, H x_acc when the
// accepted noise is recovered on the fly, Psi' x_b), done as one sweep over the columns in which
// every H / Psi' element read from shared memory serves the four chains (multicast LDS).  The CTA
// (1-8 warps, all on the same block) walks the Euler steps in lockstep: the step's table record
// (H, F0, Psi': 26 KB) is brought into shared memory ONCE per CTA by a TMA bulk copy, double-buffered
// one step ahead.  Path I/O is per step and coalesced (a chain's state is 320 contiguous bytes).
// Log-likelihood terms are accumulated per lane and reduced over the chain's eight lanes once per
// block.  Bound: FP64 pipe (3 x 1600 FMA per chain and step); shared-memory bandwidth was the limit
// of the first mapping (one chain per warp, 11 wavefronts per chain and column; now 3.3).
//
// Same rows of SURVEY.md section 8a as dmcmc_solve.cu (a1-a7, a9, layout switch, init_ll).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "dmcmc_device.cuh"
#include "dmcmc_internal.h"

namespace dmcmc {

__device__ __forceinline__ uint32_t smem_u32b(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_b(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32b(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_b(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32b(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_b(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32b(dst)), "l"(src), "r"(bytes), "r"(smem_u32b(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait_b(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAITB_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra WAITB_DONE;\n bra WAITB_LOOP;\n WAITB_DONE:\n}" ::"r"(smem_u32b(b)), "r"(parity) : "memory");
}

constexpr int RPL = 5;        // rows (state components) per lane: D / 8
#ifndef DMCMC_BIG_UNROLL
#define DMCMC_BIG_UNROLL 2
#endif
constexpr int BIG_UNROLL = DMCMC_BIG_UNROLL;  // columns per iteration of the NC = 2 mat-vec loop

__device__ __forceinline__ double group_sum(double v) {  // over the 8 lanes of one chain group
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// NC chains per lane group: a warp's four groups of eight lanes carry 4 NC chains.  Every H / Psi' element a lane reads
// from shared memory is applied to NC chains, and the chains' state vectors are interleaved ([component][NC]) so that
// one vector load brings a component of all NC chains: per column 10 table loads + 3 state loads feed 15 NC FMAs
// (NC = 1: the LSU is the busiest pipe, 43 % against 30 % FP64, profiles/r1_v5_l96_kernel_ncu_summary.txt).
template <int NC> struct VecN;
template <> struct VecN<1> { double a; };
template <> struct alignas(16) VecN<2> { double a, b; };
// component n of a group vector (n is a compile-time constant once the chain loops are unrolled; named members keep
// the vector in registers, an array member indexed by the loop variable sent it to local memory)
__device__ __forceinline__ double vget(const VecN<1> &v, int) { return v.a; }
__device__ __forceinline__ double vget(const VecN<2> &v, int n) { return n == 0 ? v.a : v.b; }
__device__ __forceinline__ void vset(VecN<1> &v, int, double x) { v.a = x; }
__device__ __forceinline__ void vset(VecN<2> &v, int n, double x) { if (n == 0) v.a = x; else v.b = x; }

template <int D, bool PSI, int MODE, bool PHILOX, bool REINV, int NC>
__global__ void __launch_bounds__(256, NC == 1 ? 2 : 1) solve_big_kernel(const SolveParams p) {
    static_assert(D == 8 * RPL, "five rows per lane");
    constexpr int CH = 4 * NC;  // chains per warp
    constexpr int REC = rec_size_big(D, PSI);
    constexpr int OFF_H = 2, OFF_F = 2 + D * D, OFF_PSI = 2 + D * D + D;
    constexpr bool SOLVE = (MODE == MODE_IMPUTE || MODE == MODE_PARAM);
    constexpr bool NEED_XA = !SOLVE || REINV;
    constexpr bool NEED_WA = SOLVE && !REINV;
    constexpr bool STORE_W_MODE = ((MODE == MODE_IMPUTE && !REINV) || MODE == MODE_RELAYOUT);
    using V = VecN<NC>;
    extern __shared__ __align__(128) double dsm[];
    __shared__ __align__(8) uint64_t full[2];
    double *stage0 = dsm, *stage1 = dsm + REC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cg = lane >> 3, g = lane & 7;  // chain group within the warp, row group
    V *vec = reinterpret_cast<V *>(dsm + 2 * REC) + (size_t)(warp * 4 + cg) * 3 * D;  // this group's [x | x_acc | x_b], [component][NC]
    V *xs = vec, *xas = vec + D, *xbs = vec + 2 * D;
    // NC = 2: the four steps' normals of this lane live in shared memory (40 registers otherwise): [step][chain][row][lane]
    constexpr bool ZSM = NC > 1 && MODE == MODE_IMPUTE && PHILOX;
    float *zsm = reinterpret_cast<float *>(reinterpret_cast<V *>(dsm + 2 * REC) + (size_t)nwarps * 4 * 3 * D) + (size_t)warp * (4 * NC * RPL * 32) + lane;
    // NC = 2: the accepted state after the current step arrives by cp.async, one step ahead, in a lane-private stage
    // [2 stages][chain][row][lane] behind the normals (registers are short; a late plain load exposed the DRAM latency
    // every step: long-scoreboard was the top stall, profiles/r2_l96_nc2_first_ncu_summary.txt)
    constexpr bool XSM = NC > 1 && NEED_XA;
    double *xst = reinterpret_cast<double *>(reinterpret_cast<float *>(reinterpret_cast<V *>(dsm + 2 * REC) + (size_t)nwarps * 4 * 3 * D) +
                                             (size_t)nwarps * (ZSM ? 4 * NC * RPL * 32 : 0)) + (size_t)warp * (2 * NC * RPL * 32) + lane;

    const int cpc = nwarps * CH;  // chains per CTA
    const int groups = (p.C + cpc - 1) / cpc;
    const int blk = (int)(blockIdx.x / groups);
    const int i1 = p.blk_i1[blk], i2 = p.blk_i2[blk];
    int chain[NC];
    bool in_range[NC], masked[NC], active[NC];
    size_t uo[NC];
    uint32_t gchain[NC];
#pragma unroll
    for (int n = 0; n < NC; ++n) {
        const int chain_raw = (int)(blockIdx.x % groups) * cpc + (warp * 4 + cg) * NC + n;
        in_range[n] = chain_raw < p.C;
        chain[n] = in_range[n] ? chain_raw : p.C - 1;  // surplus lanes shadow the last chain (stores off)
        uo[n] = (size_t)chain[n] * p.nblk + blk;
        masked[n] = (MODE == MODE_IMPUTE) && ((p.unit_mask && !p.unit_mask[uo[n]]) || (p.blk_active && !p.blk_active[blk]));
        active[n] = in_range[n] && !masked[n];
        gchain[n] = (uint32_t)(chain[n] + p.chain_offset);
    }
    int row[RPL];
#pragma unroll
    for (int m = 0; m < RPL; ++m) row[m] = g + 8 * m;

    const double Fth = p.th.v[0], sig = p.th.v[1], caux = p.th.v[2];
    const double sig2 = sig * sig, isig = 1.0 / sig;
    const bool fast_aux = p.fast_ok != 0;
    const bool pinned = PSI && p.blk_pinned[blk];
    PhiloxKeys pkeys;
    pkeys.set(p.seed);

    if (threadIdx.x == 0) {
        mbar_init_b(&full[0], 1);
        mbar_init_b(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](size_t step_index, int q) {  // thread 0: stage the record of one Euler step
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx_b(&full[q & 1], (uint32_t)(REC * sizeof(double)));
        bulk_g2s_b((q & 1) ? stage1 : stage0, p.table + step_index * REC, (uint32_t)(REC * sizeof(double)), &full[q & 1]);
    };
    if (threadIdx.x == 0) issue((size_t)p.iv_tile_off[i1] * TILE, 0);

    // ---- start point and pinned end point -----------------------------------------------------
    auto end_value = [&](int n, int j, int c) {
        const int s = p.selX_in[(size_t)j * p.C + chain[n]];
        return p.X[s][p.iv_xoff[j] + (size_t)chain[n] * p.iv_xs[j] + (size_t)(p.iv_N[j] - 1) * D + c];
    };
#pragma unroll
    for (int n = 0; n < NC; ++n)
#pragma unroll
        for (int m = 0; m < RPL; ++m) {
            const int c = row[m];
            const double v = p.iv_first[i1] ? p.x0[((size_t)p.iv_rec[i1] * D + c) * p.C + chain[n]] : end_value(n, i1 - 1, c);
            vset(xs[c], n, v);
            vset(xas[c], n, v);
            vset(xbs[c], n, pinned ? end_value(n, i2, c) : 0.0);
        }
    __syncwarp();

    double llp[NC], lla[NC], lsum[NC], lsum_a[NC];  // lsum*: per-lane partial sums of the Girsanov integrals
#pragma unroll
    for (int n = 0; n < NC; ++n) { llp[n] = 0.0; lla[n] = 0.0; lsum[n] = 0.0; lsum_a[n] = 0.0; }
    int q = 0;                        // Euler steps staged so far (pipeline position)

    // Lorenz-96 drift and its built-in auxiliary drift, row i, chain n, from a group's interleaved state
    auto drift = [&](const V *x, int i, int n) {
        const int ip = (i + 1 == D) ? 0 : i + 1, im1 = (i == 0) ? D - 1 : i - 1, im2 = (i < 2) ? i + D - 2 : i - 2;
        return (vget(x[ip], n) - vget(x[im2], n)) * vget(x[im1], n) - vget(x[i], n) + Fth;
    };
    auto aux_drift = [&](const V *x, int i, int n) {
        const int ip = (i + 1 == D) ? 0 : i + 1, im2 = (i < 2) ? i + D - 2 : i - 2;
        return ((-vget(x[i], n)) + caux * vget(x[ip], n) + (-caux) * vget(x[im2], n)) + Fth;  // B = -I + c (S_{+1} - S_{-2}), beta = F
    };

    for (int j = i1; j <= i2; ++j) {
        const int N = p.iv_N[j];
        const double rho = (MODE == MODE_IMPUTE) ? p.rho[j] : 1.0;
        const double lam = (MODE == MODE_IMPUTE) ? sqrt(1.0 - rho * rho) : 0.0;
        const double *Wa[NC], *Xa[NC];
        double *Wdst[NC], *Xp[NC];
#pragma unroll
        for (int n = 0; n < NC; ++n) {
            const int sx = p.selX_in[(size_t)j * p.C + chain[n]], sw = p.selW_in[(size_t)j * p.C + chain[n]];
            Wa[n] = p.W[sw];
            Wdst[n] = (MODE == MODE_RELAYOUT) ? p.W[sw] : p.W[1 - sw];
            Xa[n] = p.X[sx] + p.iv_xoff[j] + (size_t)chain[n] * p.iv_xs[j];
            Xp[n] = p.X[1 - sx] + p.iv_xoff[j] + (size_t)chain[n] * p.iv_xs[j];
        }
        const double *aB = p.auxB + (size_t)j * D * D, *abeta = p.auxbeta + (size_t)j * D;
        const size_t wt0 = (size_t)p.iv_tile_off[j];

        // accepted state after the current step, fetched one step ahead of its use (NC = 1: registers; NC = 2: shared stage)
        double xan[NC == 1 ? 1 : 1][RPL];
        if constexpr (NC == 1) {
#pragma unroll
            for (int m = 0; m < RPL; ++m) xan[0][m] = NEED_XA ? Xa[0][row[m]] : 0.0;
        }
        auto fetch_acc = [&](int k) {  // X_a[k] of both chains -> stage k & 1 (8-byte asynchronous copies)
#pragma unroll
            for (int n = 0; n < NC; ++n)
#pragma unroll
                for (int m = 0; m < RPL; ++m)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32b(xst + (((k & 1) * NC + n) * RPL + m) * 32)),
                                 "l"(Xa[n] + (size_t)k * D + row[m]) : "memory");
        };
        if constexpr (XSM) {
            fetch_acc(0);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }

        for (int k4 = 0; k4 < N; k4 += 4) {
            // ---- noise of the next four steps: one Philox call per owned component and chain ----
            float zf[(MODE == MODE_IMPUTE && PHILOX && !ZSM) ? NC : 1][(MODE == MODE_IMPUTE && PHILOX && !ZSM) ? RPL : 1][4];
            if constexpr (MODE == MODE_IMPUTE && PHILOX) {
#pragma unroll
                for (int n = 0; n < NC; ++n)
#pragma unroll
                    for (int m = 0; m < RPL; ++m) {
                        if constexpr (ZSM) {
                            float z4[4];
                            philox_normal4f(pkeys, p.iter, (uint32_t)(j + p.interval_offset), gchain[n], (uint32_t)row[m], (uint32_t)(k4 >> 2), z4);
#pragma unroll
                            for (int e = 0; e < 4; ++e) zsm[((e * NC + n) * RPL + m) * 32] = z4[e];
                        } else {
                            philox_normal4f(pkeys, p.iter, (uint32_t)(j + p.interval_offset), gchain[n], (uint32_t)row[m], (uint32_t)(k4 >> 2), zf[n][m]);
                        }
                    }
            }
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
                const int k = k4 + s4;
                if (k >= N) break;  // CTA-uniform
                // ---- pipeline: everyone is done with the other stage, refill it one step ahead ----
                __syncthreads();
                if (threadIdx.x == 0) {
                    if (k + 1 < N) issue((wt0 * TILE) + k + 1, q + 1);
                    else if (j + 1 <= i2) issue((size_t)p.iv_tile_off[j + 1] * TILE, q + 1);
                }
                mbar_wait_b(&full[q & 1], (uint32_t)((q >> 1) & 1));
                const double *rec = (q & 1) ? stage1 : stage0;
                ++q;
                if (q == 1) {
                    // log h~(t_a, x_a) = -c - 1/2 x'Hx + F'x at the block's first grid point
#pragma unroll
                    for (int n = 0; n < NC; ++n) {
                        double part = 0.0;
#pragma unroll
                        for (int m = 0; m < RPL; ++m) {
                            const int c = row[m];
                            double hx = 0.0, px = 0.0, qx = 0.0;
                            for (int kk = 0; kk < D; ++kk) {
                                hx += rec[OFF_H + kk * D + c] * vget(xs[kk], n);
                                if (PSI) px += rec[OFF_PSI + kk * D + c] * vget(xbs[kk], n);
                                if (pinned) qx += p.blk_Qc[(size_t)blk * D * D + (size_t)kk * D + c] * vget(xbs[kk], n);
                            }
                            double f = rec[OFF_F + c];
                            if (PSI) f += px;
                            part += f * vget(xs[c], n) - 0.5 * vget(xs[c], n) * hx;
                            if (pinned) part -= p.blk_g[(size_t)blk * D + c] * vget(xbs[c], n) + 0.5 * vget(xbs[c], n) * qx;
                        }
                        llp[n] = group_sum(part) - p.blk_c0[blk];
                        lla[n] = llp[n];
                    }
                }
                const double dt = rec[0], sq = rec[1];
                double xacur[NC][RPL];
                if constexpr (NC == 1) {
#pragma unroll
                    for (int n = 0; n < NC; ++n)
#pragma unroll
                        for (int m = 0; m < RPL; ++m) xacur[n][m] = xan[n][m];
                    if constexpr (NEED_XA) {
                        if (k + 1 < N) {
#pragma unroll
                            for (int n = 0; n < NC; ++n)
#pragma unroll
                                for (int m = 0; m < RPL; ++m) xan[n][m] = Xa[n][(size_t)(k + 1) * D + row[m]];
                        }
                    }
                } else if constexpr (NEED_XA) {
                    if (k + 1 < N) fetch_acc(k + 1);
                    asm volatile("cp.async.commit_group;" ::: "memory");  // always: the wait below counts groups
                }
                // ---- the three mat-vecs share one sweep over the columns; every H / Psi' element read
                //      from shared memory serves the 4 NC chains of the warp ----
                double Hx[NC][RPL], Hxa[NC][RPL], Px[NC][RPL];
#pragma unroll
                for (int n = 0; n < NC; ++n)
#pragma unroll
                    for (int m = 0; m < RPL; ++m) { Hx[n][m] = 0.0; Hxa[n][m] = 0.0; Px[n][m] = 0.0; }
                if constexpr (NC == 1) {
#pragma unroll 4
                    for (int kk = 0; kk < D; ++kk) {
                        const V xk = xs[kk], xak = xas[kk], xbk = xbs[kk];
                        const double *hrow = rec + OFF_H + kk * D + g, *prow = rec + OFF_PSI + kk * D + g;
#pragma unroll
                        for (int m = 0; m < RPL; ++m) {
                            const double h = hrow[8 * m];
                            const double pv = PSI ? prow[8 * m] : 0.0;
#pragma unroll
                            for (int n = 0; n < NC; ++n) {
                                if (SOLVE) Hx[n][m] += h * vget(xk, n);
                                if (NEED_XA) Hxa[n][m] += h * vget(xak, n);
                                if (PSI) Px[n][m] += pv * vget(xbk, n);
                            }
                        }
                    }
                } else {
                    // three running shared-memory addresses, bumped per column: nothing in the loop depends on the lane id
                    // (under register pressure the compiler re-read %tid.x in every iteration: S2R + short-scoreboard stall)
                    uint32_t ha = smem_u32b(rec + OFF_H + g), pa = smem_u32b(rec + OFF_PSI + g), va = smem_u32b(xs);
#pragma unroll BIG_UNROLL
                    for (int kk = 0; kk < D; ++kk) {
                        double xk[2], xak[2], xbk[2], h[RPL], pv[RPL];
                        if (SOLVE) asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(xk[0]), "=d"(xk[1]) : "r"(va));
                        if (NEED_XA) asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(xak[0]), "=d"(xak[1]) : "r"(va + (uint32_t)(D * sizeof(V))));
                        if (PSI) asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(xbk[0]), "=d"(xbk[1]) : "r"(va + (uint32_t)(2 * D * sizeof(V))));
#pragma unroll
                        for (int m = 0; m < RPL; ++m) {
                            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(h[m]) : "r"(ha + (uint32_t)(64 * m)));
                            if (PSI) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(pv[m]) : "r"(pa + (uint32_t)(64 * m)));
                        }
#pragma unroll
                        for (int m = 0; m < RPL; ++m)
#pragma unroll
                            for (int n = 0; n < NC; ++n) {
                                if (SOLVE) Hx[n][m] += h[m] * xk[n];
                                if (NEED_XA) Hxa[n][m] += h[m] * xak[n];
                                if (PSI) Px[n][m] += pv[m] * xbk[n];
                            }
                        ha += (uint32_t)(D * 8); pa += (uint32_t)(D * 8); va += (uint32_t)sizeof(V);
                    }
                }
                if constexpr (XSM) {
                    asm volatile("cp.async.wait_group 1;" ::: "memory");  // everything but the copy issued this step has landed
#pragma unroll
                    for (int n = 0; n < NC; ++n)
#pragma unroll
                        for (int m = 0; m < RPL; ++m) xacur[n][m] = xst[(((k & 1) * NC + n) * RPL + m) * 32];
                }
                double xnew[NC][RPL], dwv[NC][RPL];
#pragma unroll
                for (int n = 0; n < NC; ++n)
#pragma unroll
                    for (int m = 0; m < RPL; ++m) {
                        const int c = row[m];
                        double F = rec[OFF_F + c];
                        if (PSI) F += Px[n][m];
                        double dw = 0.0;
                        if constexpr (NEED_XA) {  // recover the accepted noise increment
                            const double r = F - Hxa[n][m];
                            const double b = drift(xas, c, n);
                            double bt;
                            if (fast_aux) bt = aux_drift(xas, c, n);
                            else {
                                double a = 0.0;
                                for (int kk = 0; kk < D; ++kk) a += aB[(size_t)c * D + kk] * vget(xas[kk], n);
                                bt = a + abeta[c];
                            }
                            lsum_a[n] += ((b - bt) * r) * dt;
                            const double res = xacur[n][m] - vget(xas[c], n) - (b + sig2 * r) * dt;
                            dw = res * isig;
                        } else if constexpr (NEED_WA) {
                            dw = Wa[n][w_off(wt0 + (size_t)(k >> 3), chain[n], c, p.C, D) + (k & 7)];
                        }
                        if constexpr (MODE == MODE_IMPUTE) {
                            double z;
                            if constexpr (ZSM) z = (double)zsm[((s4 * NC + n) * RPL + m) * 32];
                            else if constexpr (PHILOX) z = (double)zf[n][m][s4];
                            else z = p.Z[w_off(wt0 + (size_t)(k >> 3), chain[n], c, p.C, D) + (k & 7)];
                            dw = lam * (sq * z) + rho * dw;
                        }
                        dwv[n][m] = dw;
                        xnew[n][m] = 0.0;
                        if constexpr (SOLVE) {
                            const double r = F - Hx[n][m];
                            const double b = drift(xs, c, n);
                            double bt;
                            if (fast_aux) bt = aux_drift(xs, c, n);
                            else {
                                double a = 0.0;
                                for (int kk = 0; kk < D; ++kk) a += aB[(size_t)c * D + kk] * vget(xs[kk], n);
                                bt = a + abeta[c];
                            }
                            lsum[n] += ((b - bt) * r) * dt;
                            xnew[n][m] = vget(xs[c], n) + (b + sig2 * r) * dt + sig * dw;
                        }
                    }
                __syncwarp();  // everyone has read the old state
#pragma unroll
                for (int m = 0; m < RPL; ++m) {
                    if constexpr (SOLVE) {
                        V t;
#pragma unroll
                        for (int n = 0; n < NC; ++n) vset(t, n, xnew[n][m]);
                        xs[row[m]] = t;
                    }
                    if constexpr (NEED_XA) {
                        V t;
#pragma unroll
                        for (int n = 0; n < NC; ++n) vset(t, n, xacur[n][m]);
                        xas[row[m]] = t;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int n = 0; n < NC; ++n)
                    if (active[n]) {
                        if constexpr (SOLVE) {
                            const bool pin_here = pinned && j == i2 && k == N - 1;  // the block's right end stays x_b
#pragma unroll
                            for (int m = 0; m < RPL; ++m) Xp[n][(size_t)k * D + row[m]] = pin_here ? vget(xbs[row[m]], n) : xnew[n][m];
                        }
                        if constexpr (STORE_W_MODE) {
#pragma unroll
                            for (int m = 0; m < RPL; ++m) Wdst[n][w_off(wt0 + (size_t)(k >> 3), chain[n], row[m], p.C, D) + (k & 7)] = dwv[n][m];
                        }
                    }
            }
        }
    }
#pragma unroll
    for (int n = 0; n < NC; ++n) {
        llp[n] += group_sum(lsum[n]);
        lla[n] += group_sum(lsum_a[n]);
    }
#pragma unroll
    for (int n = 0; n < NC; ++n) {
        if (!in_range[n]) continue;
        const int ch = chain[n];
        if constexpr (MODE == MODE_IMPUTE) {
            if (masked[n]) {
                for (int j = i1 + g; j <= i2; j += 8) {
                    const size_t o = (size_t)j * p.C + ch;
                    p.selX_out[o] = p.selX_in[o];
                    p.selW_out[o] = p.selW_in[o];
                }
                if (g == 0 && p.blk_active && !p.blk_active[blk]) {
                    const double nan = __longlong_as_double(0x7ff8000000000000LL);
                    if (p.out_llprop) p.out_llprop[(size_t)ch * p.out_stride + blk] = nan;
                    if (p.out_ll) p.out_ll[(size_t)ch * p.out_stride + blk] = nan;
                    if (p.out_acc) p.out_acc[(size_t)ch * p.out_stride + blk] = 0;
                }
                continue;
            }
            double llcur = REINV ? lla[n] : p.ll[uo[n]];
            const double e = PHILOX ? philox_exponential(p.seed, p.iter, (uint32_t)(i1 + p.interval_offset), gchain[n]) : p.E[uo[n]];
            const bool ok = isfinite(llp[n]);
            const bool acc = p.force_accept ? ok : (e > -(llp[n] - llcur));  // identical on the eight lanes of a group
            for (int j = i1 + g; j <= i2; j += 8) {
                const size_t o = (size_t)j * p.C + ch;
                p.selX_out[o] = p.selX_in[o] ^ (uint8_t)acc;
                p.selW_out[o] = p.selW_in[o] ^ (uint8_t)(acc && STORE_W_MODE);
            }
            if (g == 0) {
                if (acc) {
                    llcur = llp[n];
                    if (p.acc_count) atomicAdd(p.acc_count + blk, 1ULL);
                }
                p.ll[uo[n]] = llcur;
                if (p.prop_count) atomicAdd(p.prop_count + blk, 1ULL);
                if (p.out_llprop) p.out_llprop[(size_t)ch * p.out_stride + blk] = llp[n];
                if (p.out_ll) p.out_ll[(size_t)ch * p.out_stride + blk] = llcur;
                if (p.out_acc) p.out_acc[(size_t)ch * p.out_stride + blk] = (uint8_t)acc;
            }
        } else if constexpr (MODE == MODE_PARAM) {
            if (g == 0) {
                if (p.out_llprop) p.out_llprop[uo[n]] = llp[n];
                if (p.out_acc) p.out_acc[uo[n]] = (uint8_t)isfinite(llp[n]);
            }
        } else {
            if (g == 0) p.ll[uo[n]] = lla[n];
        }
    }
}

template <int D, bool PSI, int NC>
static cudaError_t launch_big_nc(int mode, const SolveParams &p, cudaStream_t s) {
    constexpr int CH = 4 * NC;
    // warps per CTA (all on one block: the per-step table record is staged once per CTA).  Every CTA of a launch runs
    // for about the same time (same number of Euler steps in lockstep), so a grid of 1.19 waves costs two full rounds
    // (measured: 528 CTAs on 444 slots).  Score = warps resident per SM averaged over the rounds the grid needs; ties
    // go to the smaller CTA while the grid does not yet cover every SM, to the larger one otherwise.
    const int regs = NC == 1 ? 128 : 255;
    int warps = 1;
    double best = -1.0;
    for (int w = 1; w <= 8; w <<= 1) {
        const long long g = (long long)p.nblk * ((p.C + w * CH - 1) / (w * CH));
        const size_t sm = (size_t)(2 * rec_size_big(D, PSI) + w * CH * 3 * D) * sizeof(double) + (NC > 1 ? (size_t)w * (4 * NC * RPL * 32 * sizeof(float) + 2 * NC * RPL * 32 * sizeof(double)) : 0) + 1024;
        const int cps = (int)std::max<size_t>(1, std::min<size_t>((size_t)227 * 1024 / sm, (size_t)(65536 / regs / (32 * w))));
        const double slots = 148.0 * cps, waves = (double)g / slots;
        const double score = waves <= 1.0 ? (double)g * w / 148.0 : waves / std::ceil(waves) * cps * w;
        if (score > best * 1.05 || (score >= best * 0.95 && g >= 148)) { best = std::max(best, score); warps = w; }
    }
    const unsigned groups = (unsigned)((p.C + warps * CH - 1) / (warps * CH));
    const unsigned grid = groups * (unsigned)p.nblk;
    const unsigned threads = warps * 32;
    const size_t smem = (size_t)(2 * rec_size_big(D, PSI) + warps * CH * 3 * D) * sizeof(double) + (NC > 1 ? (size_t)warps * (4 * NC * RPL * 32 * sizeof(float) + 2 * NC * RPL * 32 * sizeof(double)) : 0);
    const bool philox = (p.Z == nullptr), reinv = p.reinvert != 0;
#define DMCMC_LAUNCH(MODE_, PH_, RE_)                                                                         \
    do {                                                                                                       \
        cudaError_t e_ = cudaFuncSetAttribute(solve_big_kernel<D, PSI, MODE_, PH_, RE_, NC>,                   \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);        \
        if (e_ != cudaSuccess) return e_;                                                                      \
        solve_big_kernel<D, PSI, MODE_, PH_, RE_, NC><<<grid, threads, smem, s>>>(p);                         \
    } while (0)
    switch (mode) {
    case MODE_IMPUTE:
        if (philox) { if (reinv) DMCMC_LAUNCH(MODE_IMPUTE, true, true); else DMCMC_LAUNCH(MODE_IMPUTE, true, false); }
        else        { if (reinv) DMCMC_LAUNCH(MODE_IMPUTE, false, true); else DMCMC_LAUNCH(MODE_IMPUTE, false, false); }
        break;
    case MODE_PARAM: DMCMC_LAUNCH(MODE_PARAM, false, false); break;
    case MODE_RELAYOUT: DMCMC_LAUNCH(MODE_RELAYOUT, false, false); break;
    case MODE_LOGLIK: DMCMC_LAUNCH(MODE_LOGLIK, false, false); break;
    default: return cudaErrorInvalidValue;
    }
#undef DMCMC_LAUNCH
    return cudaGetLastError();
}

template <int D, bool PSI>
static cudaError_t launch_big_model(int mode, const SolveParams &p, cudaStream_t s) {
    if ((long long)p.nblk * p.C == 0) return cudaSuccess;
    // Two chains per lane group (NC = 2) halve the shared-memory wavefronts per FMA, but need 255 registers (7 warps per
    // SM instead of 16): measured on B200, 256 chains, half_len 16 -- 1024 obs: 30.5 ms per sweep against 31.6 ms;
    // 4096 obs (C5): 119.7 ms against 108.4 ms.  The one-chain mapping stays the default; DMCMC_BIG_NC=2 selects the
    // other (bit-identical results, tests/test_gpu_l96.py runs both).
    const char *env = getenv("DMCMC_BIG_NC");  // read per launch
    const bool two = env && atoi(env) == 2;
    return two ? launch_big_nc<D, PSI, 2>(mode, p, s) : launch_big_nc<D, PSI, 1>(mode, p, s);
}

cudaError_t launch_solve_big(int d, int mode, const SolveParams &p, cudaStream_t s) {
    if (d != 40) return cudaErrorInvalidValue;
    return p.has_psi ? launch_big_model<40, true>(mode, p, s) : launch_big_model<40, false>(mode, p, s);
}

}  // namespace dmcmc
