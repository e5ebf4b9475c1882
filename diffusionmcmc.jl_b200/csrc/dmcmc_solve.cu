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

// ---- guiding-term offset of this chain at one grid point: F = F0 + Psi x_b (a3) ---------------
template <int D, bool PSI>
__device__ __forceinline__ void guiding_F(const double *rec, const double *xb, double *F) {
    using R = Rec<D, PSI>;
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
}

// ---- one guided step: Girsanov integrand G*dt and the drift part (b + a r) dt -----------------
template <class MD, bool PSI>
__device__ __forceinline__ void guided_terms(const Theta &th, const double *__restrict__ rec,
                                             const double *F, const double *aB,
                                             const double *abeta, const double *aatil,
                                             const double *x, double dt, double *drift_dt,
                                             double &dll) {
    constexpr int D = MD::D;
    using R = Rec<D, PSI>;
    double r[D], b[D], ar[D];
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

// ---- FitzHugh-Nagumo fused step (built-in FitzHughNagumoAux only) ------------------------------
// With the aux law's second row equal to the target's (B = [., .; gamma, -1], beta~_2 = beta) and
// B_12 = -1/eps, the Girsanov integrand collapses to
//     (b - b~)'r = [ (1/eps - B_11) y - y^3/eps - (beta~_1 - s/eps) ] r_1 ,
// a r = (0, sigma^2 r_2), and the noise recovery only needs the second coordinate.  Same
// mathematics as guided_terms, ~3x fewer FP64 instructions; differences are rounding-level.
struct FhnK {
    double ie, s, gam, bet, sig, sig2, isig;  // per law
    double c1, c2;                            // per interval
    __device__ __forceinline__ void set_law(const Theta &th) {
        ie = 1.0 / th.v[0]; s = th.v[1]; gam = th.v[2]; bet = th.v[3]; sig = th.v[4];
        sig2 = sig * sig; isig = 1.0 / sig;
    }
    __device__ __forceinline__ void set_interval(const double *aB, const double *abeta) {
        c1 = ie - aB[0];
        c2 = abeta[0] - s * ie;
    }
    // Girsanov increment and guiding term r_2 at state (y, x)
    template <bool PSI>
    __device__ __forceinline__ void terms(const double *rec, const double *F, double y, double x,
                                          double dt, double &dll, double &r2) const {
        using R = Rec<2, PSI>;
        const double r1 = fma(-rec[R::OFF_H + 1], x, fma(-rec[R::OFF_H], y, F[0]));
        r2 = fma(-rec[R::OFF_H + 2], x, fma(-rec[R::OFF_H + 1], y, F[1]));
        const double y3 = (y * y) * y;
        const double d1 = fma(-ie, y3, fma(c1, y, -c2));
        dll = (d1 * r1) * dt;
    }
};

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
        x[c] = X[x_off(tile, chain, c, p.C, D) + (k & 7)];
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

// ---- TMA bulk copy + mbarrier plumbing (sm_90+/sm_100a PTX) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (SASS: UBLKCP.S.G), completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra WAIT_DONE;\n bra WAIT_LOOP;\n WAIT_DONE:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

template <int REC>
__device__ __forceinline__ void load_rec_smem(const double *p, double *r) {
#pragma unroll
    for (int i = 0; i < REC / 2; ++i) {
        const double2 v = reinterpret_cast<const double2 *>(p)[i];
        r[2 * i] = v.x;
        r[2 * i + 1] = v.y;
    }
}

constexpr int CHUNK_TILES = 16;  // table records staged per bulk copy: 16 tiles = 128 Euler steps

// One CTA = one layout block x blockDim.x consecutive chains; one thread = one (chain, block)
// unit.  The block's backward-filter table is streamed through shared memory in chunks of
// CHUNK_TILES tiles by TMA bulk copies (double-buffered, mbarrier-signalled), so the per-step
// H/F/Psi record is an LDS broadcast instead of a dependent global load.
//   MODE   : see SolveMode
//   PHILOX : in-register counter-based noise (production) vs explicit Z/E from HBM (parity)
//   REINV  : the accepted noise increment and the accepted ll are recovered on the fly from the
//            accepted path X (fused layout switch; no W traffic) instead of read from W
//   FAST   : model-specific fused step (FitzHugh-Nagumo with its built-in aux law)
template <class MD, bool PSI, int MODE, bool PHILOX, bool REINV, bool FAST>
__global__ void __launch_bounds__(128, 3) solve_kernel(const SolveParams p) {
    constexpr int D = MD::D, M = MD::M;
    using R = Rec<D, PSI>;
    constexpr int STAGE = CHUNK_TILES * TILE * R::SIZE;  // doubles per stage
    __shared__ alignas(128) double stab[2][STAGE];
    __shared__ alignas(8) uint64_t full[2];

    const int groups = (p.C + (int)blockDim.x - 1) / (int)blockDim.x;
    const int blk = (int)(blockIdx.x / groups);
    const int chain_raw = (int)(blockIdx.x % groups) * (int)blockDim.x + (int)threadIdx.x;
    const bool in_range = chain_raw < p.C;
    const int chain = in_range ? chain_raw : p.C - 1;  // out-of-range lanes shadow the last chain, stores off
    const size_t uo = (size_t)chain * p.nblk + blk;    // [C][nblk] host layout of ll
    const int i1 = p.blk_i1[blk], i2 = p.blk_i2[blk];
    const bool masked = (MODE == MODE_IMPUTE) && ((p.unit_mask && !p.unit_mask[uo]) || (p.blk_active && !p.blk_active[blk]));
    const bool active = in_range && !masked;

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int j, int c0, int q) {  // thread 0: stage chunk (j, c0) of the table
        const int nt = (p.iv_N[j] + TILE - 1) >> 3;
        const int ntl = min(CHUNK_TILES, nt - c0);
        const uint32_t bytes = (uint32_t)(ntl * TILE * R::SIZE * sizeof(double));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[q & 1], bytes);
        bulk_g2s(stab[q & 1], p.table + ((size_t)p.iv_tile_off[j] + c0) * TILE * R::SIZE, bytes, &full[q & 1]);
    };
    if (threadIdx.x == 0) issue(i1, 0, 0);

    const bool pinned = PSI && p.blk_pinned[blk];
    double x[D], xb[D], xa[D];  // proposal state, pinned end point, accepted state (REINV / RELAYOUT)
    load_start<D>(p, chain, i1, x);
#pragma unroll
    for (int c = 0; c < D; ++c) { xb[c] = 0.0; xa[c] = x[c]; }
    if (pinned) load_end<D>(p, chain, i2, xb);

    FhnK fk;
    if constexpr (FAST) fk.set_law(p.th);

    double llp = 0.0, lla = 0.0;  // proposal ll, accepted-path ll (recomputed when REINV)
    bool ok = true;
    const uint32_t gchain = (uint32_t)(chain + p.chain_offset);
    int q = 0;

    constexpr bool SOLVE = (MODE == MODE_IMPUTE || MODE == MODE_PARAM);
    constexpr bool NEED_XA = !SOLVE || REINV;   // walk the accepted path
    constexpr bool NEED_WA = SOLVE && !REINV;   // read the stored accepted noise

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
        if constexpr (FAST) fk.set_interval(aB, abeta);
        const double rho = (MODE == MODE_IMPUTE) ? p.rho[j] : 1.0;
        const double lam = (MODE == MODE_IMPUTE) ? sqrt(1.0 - rho * rho) : 0.0;
        const double *Wa = p.W[sw];
        double *Wp = p.W[1 - sw];
        // pair-cooperative access to the X containers (see PairX): lanes 2i / 2i+1 share both chains' pointers
        PairX<D> px;
        {
            const int odd = threadIdx.x & 1;
            const int sx_o = __shfl_xor_sync(0xffffffffu, sx, 1);
            const int chain_o = __shfl_xor_sync(0xffffffffu, chain, 1);
            const int act_o = __shfl_xor_sync(0xffffffffu, (int)active, 1);
            const int sA = odd ? sx_o : sx, sB = odd ? sx : sx_o;
            px.odd = odd; px.C = p.C;
            px.cA = odd ? chain_o : chain; px.cB = odd ? chain : chain_o;
            px.accA = p.X[sA]; px.accB = p.X[sB];
            px.prpA = p.X[1 - sA]; px.prpB = p.X[1 - sB];
            px.storeA = odd ? act_o : (int)active; px.storeB = odd ? (int)active : act_o;
        }
        double lsum = 0.0, lsum_a = 0.0;

        // software prefetch of the x-independent streams, one tile ahead
        double dwn[NEED_WA ? M : 1][TILE];
        double xnA[NEED_XA ? D : 1][4], xnB[NEED_XA ? D : 1][4];  // raw halves of the next accepted-X tile
        {
            const size_t base = (toff * p.C + chain) * TILE;
            if constexpr (NEED_WA) {
#pragma unroll
                for (int w = 0; w < M; ++w) ld_tile(Wa + x_off(toff, chain, w, p.C, M), dwn[w]);
            }
            if constexpr (NEED_XA) px.load(toff, xnA, xnB);
        }

        for (int c0 = 0; c0 < nt; c0 += CHUNK_TILES, ++q) {
            __syncthreads();  // every thread is done with chunk q-1: its stage may be refilled
            if (threadIdx.x == 0) {
                if (c0 + CHUNK_TILES < nt) issue(j, c0 + CHUNK_TILES, q + 1);
                else if (j + 1 <= i2) issue(j + 1, 0, q + 1);
            }
            mbar_wait(&full[q & 1], (uint32_t)((q >> 1) & 1));
            const double *srec = stab[q & 1];

            if (q == 0) {
                // log h~(t_a, x_a) = -c - 1/2 x'Hx + F'x at the block's first grid point (loglikhd_obs)
                double rec[R::SIZE], F[D];
                load_rec_smem<R::SIZE>(srec, rec);
                guiding_F<D, PSI>(rec, xb, F);
                double c = p.blk_c0[blk];
                if (pinned) {
                    double gx = 0.0, qq = 0.0;
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        gx += p.blk_g[(size_t)blk * D + i] * xb[i];
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k < D; ++k) s += p.blk_Qc[(size_t)blk * D * D + i * D + k] * xb[k];
                        qq += xb[i] * s;
                    }
                    c += gx + 0.5 * qq;
                }
                double xHx = 0.0, Fx = 0.0;
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) s += rec[R::OFF_H + hidx(D, i, k)] * x[k];
                    xHx += x[i] * s;
                    Fx += F[i] * x[i];
                }
                llp = -c - 0.5 * xHx + Fx;
                lla = llp;
            }

            const int tl_end = min(nt, c0 + CHUNK_TILES);
            for (int tl = c0; tl < tl_end; ++tl) {
                const size_t tile = toff + tl;
                const size_t base = (tile * p.C + chain) * TILE;
                double dwa[NEED_WA ? M : 1][TILE], xat[NEED_XA ? D : 1][TILE];
                if constexpr (NEED_WA) {
#pragma unroll
                    for (int w = 0; w < M; ++w)
#pragma unroll
                        for (int s = 0; s < TILE; ++s) dwa[w][s] = dwn[w][s];
                    if (tl + 1 < nt) {
#pragma unroll
                        for (int w = 0; w < M; ++w)
                            ld_tile(Wa + x_off(tile + 1, chain, w, p.C, M), dwn[w]);
                    }
                }
                if constexpr (NEED_XA) {
                    px.unpack(xnA, xnB, xat);
                    if (tl + 1 < nt) px.load(tile + 1, xnA, xnB);
                }
                double z[(MODE == MODE_IMPUTE) ? M : 1][TILE];
                if constexpr (MODE == MODE_IMPUTE) {
                    if constexpr (PHILOX) {
#pragma unroll
                        for (int w = 0; w < M; ++w)
#pragma unroll
                            for (int h = 0; h < TILE / 4; ++h)
                                philox_normal4(p.seed, p.iter, (uint32_t)(j + p.interval_offset), gchain, (uint32_t)w,
                                               (uint32_t)(tl * (TILE / 4) + h), &z[w][4 * h]);
                    } else {
#pragma unroll
                        for (int w = 0; w < M; ++w) ld_tile_stream(p.Z + x_off(tile, chain, w, p.C, M), z[w]);
                    }
                }
                double xo[SOLVE ? D : 1][TILE], dwo[M][TILE];
                const double *trec = srec + (size_t)(tl - c0) * TILE * R::SIZE;
#pragma unroll
                for (int s = 0; s < TILE; ++s) {
                    double rec[R::SIZE], F[D];
                    load_rec_smem<R::SIZE>(trec + s * R::SIZE, rec);
                    guiding_F<D, PSI>(rec, xb, F);
                    const double dt = rec[0];
                    const bool live = tl * TILE + s < N;
                    double dw[M];
#pragma unroll
                    for (int w = 0; w < M; ++w) dw[w] = 0.0;
                    if constexpr (NEED_XA) {  // recover the accepted noise increment from the accepted path
                        if (live) {
                            if constexpr (FAST) {
                                double dll, r2;
                                fk.terms<PSI>(rec, F, xa[0], xa[1], dt, dll, r2);
                                lsum_a += dll;
                                const double u = fma(fk.sig2, r2, fma(fk.gam, xa[0], fk.bet - xa[1]));
                                dw[0] = fma(-u, dt, xat[1][s] - xa[1]) * fk.isig;
                            } else {
                                double dr[D], res[D], dll;
                                guided_terms<MD, PSI>(p.th, rec, F, aB, abeta, aatil, xa, dt, dr, dll);
                                lsum_a += dll;
#pragma unroll
                                for (int c = 0; c < D; ++c) res[c] = xat[c][s] - xa[c] - dr[c];
                                MD::sigma_solve(p.th, xa, res, dw);
                            }
#pragma unroll
                            for (int c = 0; c < D; ++c) xa[c] = xat[c][s];
                        }
                    } else if constexpr (NEED_WA) {
#pragma unroll
                        for (int w = 0; w < M; ++w) dw[w] = dwa[w][s];
                    }
                    if constexpr (MODE == MODE_IMPUTE) {  // pCN: W deg = sqrt(1-rho^2) W_fresh + rho W
#pragma unroll
                        for (int w = 0; w < M; ++w) dw[w] = lam * (rec[1] * z[w][s]) + rho * dw[w];
                    }
#pragma unroll
                    for (int w = 0; w < M; ++w) dwo[w][s] = dw[w];
                    if constexpr (SOLVE) {
                        if (live) {
                            if constexpr (FAST) {
                                double dll, r2;
                                fk.terms<PSI>(rec, F, x[0], x[1], dt, dll, r2);
                                lsum += dll;
                                const double y3 = (x[0] * x[0]) * x[0];
                                const double wy = (x[0] - y3) + (fk.s - x[1]);
                                const double u = fma(fk.sig2, r2, fma(fk.gam, x[0], fk.bet - x[1]));
                                const double ynew = fma(wy, dt * fk.ie, x[0]);
                                x[1] = fma(u, dt, fma(fk.sig, dw[0], x[1]));
                                x[0] = ynew;
                            } else {
                                double dr[D], sdw[D], dll;
                                guided_terms<MD, PSI>(p.th, rec, F, aB, abeta, aatil, x, dt, dr, dll);
                                lsum += dll;
                                MD::sigma_mul(p.th, x, dw, sdw);
#pragma unroll
                                for (int c = 0; c < D; ++c) x[c] = x[c] + dr[c] + sdw[c];
                                ok = ok && MD::bound_ok(x);
                            }
                        }
#pragma unroll
                        for (int c = 0; c < D; ++c) xo[c][s] = x[c];
                    }
                }
                if constexpr (SOLVE) {
                    if (pinned && j == i2 && tl == nt - 1) {  // pin the block's right end to x_b
                        const int last = (N - 1) & 7;
#pragma unroll
                        for (int s = 0; s < TILE; ++s)
#pragma unroll
                            for (int c = 0; c < D; ++c) xo[c][s] = (s == last) ? xb[c] : xo[c][s];
                    }
                    px.store(tile, xo);  // every lane takes part (shuffles); stores are predicated per chain
                }
                if (active) {
                    if ((MODE == MODE_IMPUTE && p.store_w) || MODE == MODE_RELAYOUT) {
                        double *Wdst = (MODE == MODE_RELAYOUT) ? p.W[sw] : Wp;
#pragma unroll
                        for (int w = 0; w < M; ++w) st_tile(Wdst + x_off(tile, chain, w, p.C, M), dwo[w]);
                    }
                }
            }
        }
        llp += lsum;
        lla += lsum_a;
    }
    if (!ok) llp = -INFINITY;

    if constexpr (MODE == MODE_IMPUTE) {
        if (!in_range) return;
        if (masked) {
            for (int j = i1; j <= i2; ++j) {
                const size_t o = (size_t)j * p.C + chain;
                p.selX_out[o] = p.selX_in[o];
                p.selW_out[o] = p.selW_in[o];
            }
            if (p.blk_active && !p.blk_active[blk]) {  // block owned by a neighbouring shard
                const double nan = __longlong_as_double(0x7ff8000000000000LL);
                if (p.out_llprop) p.out_llprop[(size_t)chain * p.out_stride + blk] = nan;
                if (p.out_ll) p.out_ll[(size_t)chain * p.out_stride + blk] = nan;
                if (p.out_acc) p.out_acc[(size_t)chain * p.out_stride + blk] = 0;
            }
            return;
        }
        double llcur = REINV ? lla : p.ll[uo];
        const double e = PHILOX ? philox_exponential(p.seed, p.iter, (uint32_t)(i1 + p.interval_offset), gchain) : p.E[uo];
        // accepted = rand(Exponential(1)) > -(ll deg - ll)   (eMCMC.accept_reject!)
        const bool acc = p.force_accept ? ok : (e > -(llp - llcur));
        for (int j = i1; j <= i2; ++j) {  // swap_paths!: flip the accepted-container selectors
            const size_t o = (size_t)j * p.C + chain;
            p.selX_out[o] = p.selX_in[o] ^ (uint8_t)acc;
            p.selW_out[o] = p.selW_in[o] ^ (uint8_t)(acc && p.store_w);
        }
        if (acc) {
            llcur = llp;
            if (p.acc_count) atomicAdd(p.acc_count + blk, 1ULL);
        }
        p.ll[uo] = llcur;
        if (p.prop_count && chain == 0) atomicAdd(p.prop_count + blk, (unsigned long long)p.C);
        if (p.out_llprop) p.out_llprop[(size_t)chain * p.out_stride + blk] = llp;
        if (p.out_ll) p.out_ll[(size_t)chain * p.out_stride + blk] = llcur;
        if (p.out_acc) p.out_acc[(size_t)chain * p.out_stride + blk] = (uint8_t)acc;
    } else if constexpr (MODE == MODE_PARAM) {
        if (!in_range) return;
        if (p.out_llprop) p.out_llprop[uo] = llp;
        if (p.out_acc) p.out_acc[uo] = (uint8_t)ok;
    } else {
        if (in_range) p.ll[uo] = lla;
    }
}

// ---- launch dispatch ---------------------------------------------------------------------------
template <class MD, bool PSI, bool FAST>
static cudaError_t launch_model(int mode, const SolveParams &p, int threads, cudaStream_t s) {
    if ((long long)p.nblk * p.C == 0) return cudaSuccess;
    const unsigned groups = (unsigned)((p.C + threads - 1) / threads);
    const unsigned grid = groups * (unsigned)p.nblk;
    const bool philox = (p.Z == nullptr), reinv = p.reinvert != 0;
#define DMCMC_LAUNCH(MODE_, PH_, RE_) solve_kernel<MD, PSI, MODE_, PH_, RE_, FAST><<<grid, threads, 0, s>>>(p)
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

template <class MD, bool FAST = false>
static cudaError_t launch_psi(int mode, const SolveParams &p, int threads, cudaStream_t s) {
    return p.has_psi ? launch_model<MD, true, FAST>(mode, p, threads, s)
                     : launch_model<MD, false, FAST>(mode, p, threads, s);
}

cudaError_t launch_solve(int model, int mode, const SolveParams &p, int threads, cudaStream_t s) {
    switch (model) {
    case 0: return p.fast_ok ? launch_psi<ModelFHN, true>(mode, p, threads, s)
                             : launch_psi<ModelFHN, false>(mode, p, threads, s);
    case 1: return launch_psi<ModelLV>(mode, p, threads, s);
    case 2: return launch_psi<ModelL63>(mode, p, threads, s);
    case 3: return launch_solve_big(p.D, mode, p, s);
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
            v = g.X[sp][x_off(tile, chain, c, g.C, g.D) + (k & 7)];
        }
        g.outX[(pt0)*g.D + c] = v;
    }
    for (int w = 0; w < g.M; ++w) g.outW[pt0 * g.M + w] = 0.0;
    for (int k = 0; k < N; ++k) {
        const size_t tile = (size_t)g.iv_tile_off[j] + (k >> 3);
        const size_t e = (tile * g.C + chain) * TILE + (k & 7);
        for (int c = 0; c < g.D; ++c) g.outX[(pt0 + k + 1) * g.D + c] = g.X[sx][x_off(tile, chain, c, g.C, g.D) + (k & 7)];
        for (int w = 0; w < g.M; ++w)
            g.outW[(pt0 + k + 1) * g.M + w] = g.outW[(pt0 + k) * g.M + w] + g.W[sw][x_off(tile, chain, w, g.C, g.M) + (k & 7)];
    }
}

__global__ void segment_kernel(const SegmentParams g) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nj = g.j1 - g.j0;
    if (u >= (long long)nj * g.C) return;
    const int j = g.j0 + (int)(u / g.C), chain = (int)(u % g.C);
    const size_t cs = (size_t)g.TT * g.C * TILE;
    double *X = g.X[g.selX[(size_t)j * g.C + chain]];
    const int off = g.iv_st_off[j] - g.iv_st_off[g.j0];
    for (int k = 0; k < g.iv_N[j]; ++k) {
        const size_t tile = (size_t)g.iv_tile_off[j] + (k >> 3);
        const size_t e = (tile * g.C + chain) * TILE + (k & 7);
        for (int c = 0; c < g.D; ++c) {
            double *pk = g.packed + ((size_t)chain * g.steps + off + k) * g.D + c;
            const size_t xe = x_off(tile, chain, c, g.C, g.D) + (k & 7);
            if (g.upload) X[xe] = *pk;
            else *pk = X[xe];
        }
    }
    if (g.upload) {  // pad slots repeat the last value (they are no-ops for the solve kernels)
        const int N = g.iv_N[j];
        for (int k = N; k < ((N + TILE - 1) / TILE) * TILE; ++k) {
            const size_t tile = (size_t)g.iv_tile_off[j] + (k >> 3);
            const size_t e = (tile * g.C + chain) * TILE + (k & 7);
            for (int c = 0; c < g.D; ++c)
                X[x_off(tile, chain, c, g.C, g.D) + (k & 7)] = g.packed[((size_t)chain * g.steps + off + N - 1) * g.D + c];
        }
    }
}

cudaError_t launch_segment(const SegmentParams &g, cudaStream_t s) {
    const long long n = (long long)(g.j1 - g.j0) * g.C;
    if (n <= 0) return cudaSuccess;
    segment_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g);
    return cudaGetLastError();
}

cudaError_t launch_gather(const GatherParams &g, cudaStream_t s) {
    const long long n = (long long)g.J * g.C;
    if (n == 0) return cudaSuccess;
    gather_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g);
    return cudaGetLastError();
}

}  // namespace dmcmc
