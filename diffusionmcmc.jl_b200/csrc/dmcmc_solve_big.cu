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

// Lorenz-96 drift and its built-in auxiliary drift, row i, from the state in shared memory
template <int D>
__device__ __forceinline__ double l96_drift(const double *x, int i, double F) {
    const int ip = (i + 1 == D) ? 0 : i + 1, im1 = (i == 0) ? D - 1 : i - 1, im2 = (i < 2) ? i + D - 2 : i - 2;
    return (x[ip] - x[im2]) * x[im1] - x[i] + F;
}
template <int D>
__device__ __forceinline__ double l96_aux_drift(const double *x, int i, double F, double c) {
    const int ip = (i + 1 == D) ? 0 : i + 1, im2 = (i < 2) ? i + D - 2 : i - 2;
    return ((-x[i]) + c * x[ip] + (-c) * x[im2]) + F;  // B = -I + c (S_{+1} - S_{-2}), beta = F
}

constexpr int CH = 4;         // chains per warp
constexpr int RPL = 5;        // rows (state components) per lane: D / 8

__device__ __forceinline__ double group_sum(double v) {  // over the 8 lanes of one chain
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

template <int D, bool PSI, int MODE, bool PHILOX, bool REINV>
__global__ void __launch_bounds__(256, 2) solve_big_kernel(const SolveParams p) {
    static_assert(D == 8 * RPL, "five rows per lane");
    constexpr int REC = rec_size_big(D, PSI);
    constexpr int OFF_H = 2, OFF_F = 2 + D * D, OFF_PSI = 2 + D * D + D;
    constexpr bool SOLVE = (MODE == MODE_IMPUTE || MODE == MODE_PARAM);
    constexpr bool NEED_XA = !SOLVE || REINV;
    constexpr bool NEED_WA = SOLVE && !REINV;
    constexpr bool STORE_W_MODE = ((MODE == MODE_IMPUTE && !REINV) || MODE == MODE_RELAYOUT);
    extern __shared__ __align__(128) double dsm[];
    __shared__ __align__(8) uint64_t full[2];
    double *stage0 = dsm, *stage1 = dsm + REC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cg = lane >> 3, g = lane & 7;  // chain within the warp, row group
    double *vec = dsm + 2 * REC + (size_t)(warp * CH + cg) * 3 * D;  // this chain's [x | x_acc | x_b]
    double *xs = vec, *xas = vec + D, *xbs = vec + 2 * D;

    const int cpc = nwarps * CH;  // chains per CTA
    const int groups = (p.C + cpc - 1) / cpc;
    const int blk = (int)(blockIdx.x / groups);
    const int chain_raw = (int)(blockIdx.x % groups) * cpc + warp * CH + cg;
    const bool in_range = chain_raw < p.C;
    const int chain = in_range ? chain_raw : p.C - 1;  // surplus lanes shadow the last chain (stores off)
    const size_t uo = (size_t)chain * p.nblk + blk;
    const int i1 = p.blk_i1[blk], i2 = p.blk_i2[blk];
    const bool masked = (MODE == MODE_IMPUTE) &&
                        ((p.unit_mask && !p.unit_mask[uo]) || (p.blk_active && !p.blk_active[blk]));
    const bool active = in_range && !masked;
    int row[RPL];
#pragma unroll
    for (int m = 0; m < RPL; ++m) row[m] = g + 8 * m;

    const double Fth = p.th.v[0], sig = p.th.v[1], caux = p.th.v[2];
    const double sig2 = sig * sig, isig = 1.0 / sig;
    const bool fast_aux = p.fast_ok != 0;
    const bool pinned = PSI && p.blk_pinned[blk];
    const uint32_t gchain = (uint32_t)(chain + p.chain_offset);
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
    auto end_value = [&](int j, int c) {
        const int s = p.selX_in[(size_t)j * p.C + chain];
        return p.X[s][p.iv_xoff[j] + (size_t)chain * p.iv_xs[j] + (size_t)(p.iv_N[j] - 1) * D + c];
    };
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
        const int c = row[m];
        const double v = p.iv_first[i1] ? p.x0[((size_t)p.iv_rec[i1] * D + c) * p.C + chain] : end_value(i1 - 1, c);
        xs[c] = v;
        xas[c] = v;
        xbs[c] = pinned ? end_value(i2, c) : 0.0;
    }
    __syncwarp();

    double llp = 0.0, lla = 0.0;
    double lsum = 0.0, lsum_a = 0.0;  // per-lane partial sums of the Girsanov integrals
    int q = 0;                        // Euler steps staged so far (pipeline position)

    for (int j = i1; j <= i2; ++j) {
        const int N = p.iv_N[j];
        const int sx = p.selX_in[(size_t)j * p.C + chain], sw = p.selW_in[(size_t)j * p.C + chain];
        const double rho = (MODE == MODE_IMPUTE) ? p.rho[j] : 1.0;
        const double lam = (MODE == MODE_IMPUTE) ? sqrt(1.0 - rho * rho) : 0.0;
        const double *Wa = p.W[sw];
        double *Wdst = (MODE == MODE_RELAYOUT) ? p.W[sw] : p.W[1 - sw];
        const double *Xa = p.X[sx] + p.iv_xoff[j] + (size_t)chain * p.iv_xs[j];
        double *Xp = p.X[1 - sx] + p.iv_xoff[j] + (size_t)chain * p.iv_xs[j];
        const double *aB = p.auxB + (size_t)j * D * D, *abeta = p.auxbeta + (size_t)j * D;
        const size_t wt0 = (size_t)p.iv_tile_off[j];

        // accepted state after the current step, fetched one step ahead of its use
        double xan[RPL];
#pragma unroll
        for (int m = 0; m < RPL; ++m) xan[m] = NEED_XA ? Xa[row[m]] : 0.0;

        for (int k4 = 0; k4 < N; k4 += 4) {
            // ---- noise of the next four steps: one Philox call per owned component ----
            float zf[(MODE == MODE_IMPUTE && PHILOX) ? RPL : 1][4];
            if constexpr (MODE == MODE_IMPUTE && PHILOX) {
#pragma unroll
                for (int m = 0; m < RPL; ++m)
                    philox_normal4f(pkeys, p.iter, (uint32_t)(j + p.interval_offset), gchain, (uint32_t)row[m], (uint32_t)(k4 >> 2), zf[m]);
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
                    double part = 0.0;
#pragma unroll
                    for (int m = 0; m < RPL; ++m) {
                        const int c = row[m];
                        double hx = 0.0, px = 0.0, qx = 0.0;
                        for (int kk = 0; kk < D; ++kk) {
                            hx += rec[OFF_H + kk * D + c] * xs[kk];
                            if (PSI) px += rec[OFF_PSI + kk * D + c] * xbs[kk];
                            if (pinned) qx += p.blk_Qc[(size_t)blk * D * D + (size_t)kk * D + c] * xbs[kk];
                        }
                        double f = rec[OFF_F + c];
                        if (PSI) f += px;
                        part += f * xs[c] - 0.5 * xs[c] * hx;
                        if (pinned) part -= p.blk_g[(size_t)blk * D + c] * xbs[c] + 0.5 * xbs[c] * qx;
                    }
                    llp = group_sum(part) - p.blk_c0[blk];
                    lla = llp;
                }
                const double dt = rec[0], sq = rec[1];
                double xacur[RPL];
#pragma unroll
                for (int m = 0; m < RPL; ++m) xacur[m] = xan[m];
                if constexpr (NEED_XA) {
                    if (k + 1 < N) {
#pragma unroll
                        for (int m = 0; m < RPL; ++m) xan[m] = Xa[(size_t)(k + 1) * D + row[m]];
                    }
                }
                // ---- the three mat-vecs share one sweep over the columns; every H / Psi' element read
                //      from shared memory serves the four chains of the warp ----
                double Hx[RPL], Hxa[RPL], Px[RPL];
#pragma unroll
                for (int m = 0; m < RPL; ++m) { Hx[m] = 0.0; Hxa[m] = 0.0; Px[m] = 0.0; }
#pragma unroll 4
                for (int kk = 0; kk < D; ++kk) {
                    const double xk = xs[kk], xak = xas[kk], xbk = xbs[kk];
                    const double *hrow = rec + OFF_H + kk * D + g, *prow = rec + OFF_PSI + kk * D + g;
#pragma unroll
                    for (int m = 0; m < RPL; ++m) {
                        const double h = hrow[8 * m];
                        if (SOLVE) Hx[m] += h * xk;
                        if (NEED_XA) Hxa[m] += h * xak;
                        if (PSI) Px[m] += prow[8 * m] * xbk;
                    }
                }
                double xnew[RPL], dwv[RPL];
#pragma unroll
                for (int m = 0; m < RPL; ++m) {
                    const int c = row[m];
                    double F = rec[OFF_F + c];
                    if (PSI) F += Px[m];
                    double dw = 0.0;
                    if constexpr (NEED_XA) {  // recover the accepted noise increment
                        const double r = F - Hxa[m];
                        const double b = l96_drift<D>(xas, c, Fth);
                        double bt;
                        if (fast_aux) bt = l96_aux_drift<D>(xas, c, Fth, caux);
                        else {
                            double a = 0.0;
                            for (int kk = 0; kk < D; ++kk) a += aB[(size_t)c * D + kk] * xas[kk];
                            bt = a + abeta[c];
                        }
                        lsum_a += ((b - bt) * r) * dt;
                        const double res = xacur[m] - xas[c] - (b + sig2 * r) * dt;
                        dw = res * isig;
                    } else if constexpr (NEED_WA) {
                        dw = Wa[w_off(wt0 + (size_t)(k >> 3), chain, c, p.C, D) + (k & 7)];
                    }
                    if constexpr (MODE == MODE_IMPUTE) {
                        double z;
                        if constexpr (PHILOX) z = (double)zf[m][s4];
                        else z = p.Z[w_off(wt0 + (size_t)(k >> 3), chain, c, p.C, D) + (k & 7)];
                        dw = lam * (sq * z) + rho * dw;
                    }
                    dwv[m] = dw;
                    xnew[m] = 0.0;
                    if constexpr (SOLVE) {
                        const double r = F - Hx[m];
                        const double b = l96_drift<D>(xs, c, Fth);
                        double bt;
                        if (fast_aux) bt = l96_aux_drift<D>(xs, c, Fth, caux);
                        else {
                            double a = 0.0;
                            for (int kk = 0; kk < D; ++kk) a += aB[(size_t)c * D + kk] * xs[kk];
                            bt = a + abeta[c];
                        }
                        lsum += ((b - bt) * r) * dt;
                        xnew[m] = xs[c] + (b + sig2 * r) * dt + sig * dw;
                    }
                }
                __syncwarp();  // everyone has read the old state
#pragma unroll
                for (int m = 0; m < RPL; ++m) {
                    if constexpr (SOLVE) xs[row[m]] = xnew[m];
                    if constexpr (NEED_XA) xas[row[m]] = xacur[m];
                }
                __syncwarp();
                if (active) {
                    if constexpr (SOLVE) {
                        const bool pin_here = pinned && j == i2 && k == N - 1;  // the block's right end stays x_b
#pragma unroll
                        for (int m = 0; m < RPL; ++m) Xp[(size_t)k * D + row[m]] = pin_here ? xbs[row[m]] : xnew[m];
                    }
                    if constexpr (STORE_W_MODE) {
#pragma unroll
                        for (int m = 0; m < RPL; ++m) Wdst[w_off(wt0 + (size_t)(k >> 3), chain, row[m], p.C, D) + (k & 7)] = dwv[m];
                    }
                }
            }
        }
    }
    llp += group_sum(lsum);
    lla += group_sum(lsum_a);
    if (!in_range) return;

    if constexpr (MODE == MODE_IMPUTE) {
        if (masked) {
            for (int j = i1 + g; j <= i2; j += 8) {
                const size_t o = (size_t)j * p.C + chain;
                p.selX_out[o] = p.selX_in[o];
                p.selW_out[o] = p.selW_in[o];
            }
            if (g == 0 && p.blk_active && !p.blk_active[blk]) {
                const double nan = __longlong_as_double(0x7ff8000000000000LL);
                if (p.out_llprop) p.out_llprop[(size_t)chain * p.out_stride + blk] = nan;
                if (p.out_ll) p.out_ll[(size_t)chain * p.out_stride + blk] = nan;
                if (p.out_acc) p.out_acc[(size_t)chain * p.out_stride + blk] = 0;
            }
            return;
        }
        double llcur = REINV ? lla : p.ll[uo];
        const double e = PHILOX ? philox_exponential(p.seed, p.iter, (uint32_t)(i1 + p.interval_offset), gchain) : p.E[uo];
        const bool ok = isfinite(llp);
        const bool acc = p.force_accept ? ok : (e > -(llp - llcur));  // identical on the eight lanes of a chain
        for (int j = i1 + g; j <= i2; j += 8) {
            const size_t o = (size_t)j * p.C + chain;
            p.selX_out[o] = p.selX_in[o] ^ (uint8_t)acc;
            p.selW_out[o] = p.selW_in[o] ^ (uint8_t)(acc && STORE_W_MODE);
        }
        if (g == 0) {
            if (acc) {
                llcur = llp;
                if (p.acc_count) atomicAdd(p.acc_count + blk, 1ULL);
            }
            p.ll[uo] = llcur;
            if (p.prop_count) atomicAdd(p.prop_count + blk, 1ULL);
            if (p.out_llprop) p.out_llprop[(size_t)chain * p.out_stride + blk] = llp;
            if (p.out_ll) p.out_ll[(size_t)chain * p.out_stride + blk] = llcur;
            if (p.out_acc) p.out_acc[(size_t)chain * p.out_stride + blk] = (uint8_t)acc;
        }
    } else if constexpr (MODE == MODE_PARAM) {
        if (g == 0) {
            if (p.out_llprop) p.out_llprop[uo] = llp;
            if (p.out_acc) p.out_acc[uo] = (uint8_t)isfinite(llp);
        }
    } else {
        if (g == 0) p.ll[uo] = lla;
    }
}

template <int D, bool PSI>
static cudaError_t launch_big_model(int mode, const SolveParams &p, cudaStream_t s) {
    if ((long long)p.nblk * p.C == 0) return cudaSuccess;
    // warps per CTA (all on one block: the per-step table record is staged once per CTA).  Every CTA of a launch runs
    // for about the same time (same number of Euler steps in lockstep), so a grid of 1.19 waves costs two full rounds
    // (measured: 528 CTAs on 444 slots).  Score = warps resident per SM averaged over the rounds the grid needs; ties
    // go to the smaller CTA while the grid does not yet cover every SM, to the larger one otherwise.
    int warps = 1;
    double best = -1.0;
    for (int w = 1; w <= 8; w <<= 1) {
        const long long g = (long long)p.nblk * ((p.C + w * CH - 1) / (w * CH));
        const size_t sm = (size_t)(2 * rec_size_big(D, PSI) + w * CH * 3 * D) * sizeof(double) + 1024;
        const int cps = (int)std::max<size_t>(1, std::min<size_t>((size_t)227 * 1024 / sm, (size_t)(512 / (32 * w))));  // 128 registers
        const double slots = 148.0 * cps, waves = (double)g / slots;
        const double score = waves <= 1.0 ? (double)g * w / 148.0 : waves / std::ceil(waves) * cps * w;
        if (score > best * 1.05 || (score >= best * 0.95 && g >= 148)) { best = std::max(best, score); warps = w; }
    }
    const unsigned groups = (unsigned)((p.C + warps * CH - 1) / (warps * CH));
    const unsigned grid = groups * (unsigned)p.nblk;
    const unsigned threads = warps * 32;
    const size_t smem = (size_t)(2 * rec_size_big(D, PSI) + warps * CH * 3 * D) * sizeof(double);
    const bool philox = (p.Z == nullptr), reinv = p.reinvert != 0;
#define DMCMC_LAUNCH(MODE_, PH_, RE_)                                                                     \
    do {                                                                                                   \
        cudaError_t e_ = cudaFuncSetAttribute(solve_big_kernel<D, PSI, MODE_, PH_, RE_>,                   \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);    \
        if (e_ != cudaSuccess) return e_;                                                                  \
        solve_big_kernel<D, PSI, MODE_, PH_, RE_><<<grid, threads, smem, s>>>(p);                         \
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

cudaError_t launch_solve_big(int d, int mode, const SolveParams &p, cudaStream_t s) {
    if (d != 40) return cudaErrorInvalidValue;
    return p.has_psi ? launch_big_model<40, true>(mode, p, s) : launch_big_model<40, false>(mode, p, s);
}

}  // namespace dmcmc
