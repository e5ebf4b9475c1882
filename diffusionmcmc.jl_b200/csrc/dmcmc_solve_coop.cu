// dmcmc_solve_coop.cu -- the imputation update for FEW, LONG units: one WARP per (chain, block) unit.
//
// solve_kernel (dmcmc_solve.cu) gives every unit one thread; that fills the machine when there are tens
// of thousands of units (1024 chains under chequerboard blocking), but the reference's own shapes -- one
// chain without blocking (BASELINE config 1), one chain with 8 blocks (config 2), 1024 unblocked chains
// (config 3 as the reference would run it), one long recording cut into ~10^3 blocks (config 4) -- have
// 1 .. 10^3 units of 10^2 .. 10^4 serial Euler steps, and then a single lane's ~110-instruction step is the
// whole critical path.  Of that step only the recursion x_{k+1} = f(x_k) is serial.  Here the 32 lanes of a
// warp share one unit, 32 Euler steps per tile:
//   phase P (lane = step): Philox normals, accepted noise in (stored W, or -- fused layout switch -- recovered
//            from the accepted path together with the accepted path's Girsanov terms), pCN mix, W deg out, and
//            the mixed increments left in shared memory;
//   phase S (all lanes redundantly, so no divergence and no broadcast): the recursion and its Girsanov sum,
//            reading one broadcast table record and one increment per step; lane k keeps the state after step k
//            and the tile's 32 states go out as one coalesced store.
// The arithmetic per step is the single-thread kernel's, operation for operation (shared inline functions),
// so the two kernels give bit-identical results and the choice between them (launch_solve: by unit count)
// never changes a sample.  Reference rows: a1-a7 of SURVEY.md section 8a (src/run!_alterations.jl:5-47, :81-91).
#include <cstdlib>

#include "dmcmc_solve_common.cuh"

namespace dmcmc {

constexpr int CT = 32;  // Euler steps per tile = lanes

// SPLIT: every unit gets TWO warps -- a producer that runs phase P one tile ahead and a solver that runs phase S --
// so that phase P (global loads, Philox, Box-Muller) leaves the critical path altogether.
// PARAM: the parameter-step solve (a9, set_proposal! / compute_ll!, src/run!_alterations.jl:177-211, :258-265): fixed
// accepted noise under the proposal law, no mix, no accept -- per-unit ll deg and success flag out.
// MODE_RELAYOUT / MODE_LOGLIK: the accepted path only -- recover its noise (RELAYOUT: store it) and its ll under the
// current layout (update_workspaces!, src/workspaces.jl:467-471; init_ll!, :425-431): phase P alone, plus the serial sum.
// ---- end-point exchange fused into the sweep (HaloFuse, dmcmc_internal.h) --------------------------------------------
__device__ __forceinline__ unsigned long long hf_ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void hf_st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void hf_spin(const unsigned long long *flag, unsigned long long val, int *err) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    while (hf_ld_acquire_sys(flag) < val) {
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if (t - t0 > 3000000000ULL) { *err = 1; break; }  // a dead peer must not hang the GPU
    }
}
// A message holds at most 32 intervals (half_len + 1): lane i keeps interval j0 + i's first step within the message and
// the address of the chain's record in the container to use (sel_flip: the accept decision of the sweep just done), so
// that every global load the copy depends on is issued once, in one round, before the data moves
struct HfMap {
    int off;          // first Euler step of this lane's interval within the message (INT_MAX beyond the message)
    double *rec;      // the chain's record of this lane's interval in the accepted container
};
__device__ __forceinline__ HfMap hf_map(const SolveParams &p, int j0, int j1, int chain, int lane, int sel_flip) {
    HfMap m;
    m.off = 0x7fffffff;
    m.rec = nullptr;
    const int j = j0 + lane;
    if (j < j1) {
        m.off = p.hf.iv_st_off[j] - p.hf.iv_st_off[j0];
        m.rec = p.X[p.selX_in[(size_t)j * p.C + chain] ^ sel_flip] + p.iv_xoff[j] + (size_t)chain * p.iv_xs[j];
    }
    return m;
}
// element e of the chain's slice -> its address in the containers (the owning lane's record, found by shuffles)
__device__ __forceinline__ double *hf_addr(const HfMap &m, int nj, int e, int D) {
    const int step = e / D, c = e - step * D;
    int off = __shfl_sync(0xffffffffu, m.off, 0);
    unsigned long long rec = __shfl_sync(0xffffffffu, (unsigned long long)m.rec, 0);
    for (int i = 1; i < nj; ++i) {
        const int oi = __shfl_sync(0xffffffffu, m.off, i);
        const unsigned long long ri = __shfl_sync(0xffffffffu, (unsigned long long)m.rec, i);
        if (oi <= step) { off = oi; rec = ri; }
    }
    return reinterpret_cast<double *>(rec) + (size_t)(step - off) * D + c;
}
// one warp, its chain: the neighbour's message out of this GPU's mailbox into the accepted containers (and the starting
// point); the last unit of the block to finish acknowledges in the sender's memory.  One chain: no counter, no fence
// beyond the release store (a gpu-scope fence costs more than the hand-over: tools/micro/p2p_latency.cu)
__device__ __forceinline__ void hf_pull_unit(const SolveParams &p, int chain, int lane, int D) {
    const HaloFuse &h = p.hf;
    const int nj = h.pull_j1 - h.pull_j0;
    const HfMap m = hf_map(p, h.pull_j0, h.pull_j1, chain, lane, 0);
    const int n = (h.iv_st_off[h.pull_j1] - h.iv_st_off[h.pull_j0]) * D;
    if (lane == 0) hf_spin(h.pull_flag, h.pull_val, h.err);
    __syncwarp();
    const double *src = h.pull_box + (size_t)chain * h.pull_stride;
    if (h.pull_lead > 0 && lane < D)
        const_cast<double *>(p.x0)[(size_t)lane * p.C + chain] = __ldcg(src + h.pull_lead - D + lane);  // R = 1: record 0
    for (int e0 = 0; e0 < n; e0 += 32) {  // uniform trip count: the address look-up shuffles
        const int e = e0 + lane;
        double *dst = hf_addr(m, nj, min(e, n - 1), D);
        if (e < n) *dst = __ldcg(src + h.pull_lead + e);
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) {
        bool last = true;
        if (p.C > 1) {
            __threadfence();
            last = atomicAdd(h.pull_counter, 1u) == (unsigned)p.C - 1u;
            if (last) { *h.pull_counter = 0; __threadfence(); }
        }
        if (last) hf_st_release_sys(h.pull_ack, h.pull_val);
    }
}
// one warp, its chain, after the accept decision: the accepted path of the message's intervals straight into the
// neighbour's mailbox (stores over NVLink); the last unit of the block raises the neighbour's data flag
__device__ __forceinline__ void hf_push_unit(const SolveParams &p, int chain, int lane, int D, bool acc) {
    const HaloFuse &h = p.hf;
    const int nj = h.push_j1 - h.push_j0;
    const HfMap m = hf_map(p, h.push_j0, h.push_j1, chain, lane, acc ? 1 : 0);
    const int n = (h.iv_st_off[h.push_j1] - h.iv_st_off[h.push_j0]) * D;
    if (lane == 0) hf_spin(h.push_wait, h.push_wait_val, h.err);  // the previous message has been taken out
    __threadfence_block();  // this warp's proposal stores are what is read back below (from L2)
    __syncwarp();
    double *dst = h.push_box + (size_t)chain * h.push_stride;
    for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        const double *src = hf_addr(m, nj, min(e, n - 1), D);
        if (e < n) dst[e] = __ldcg(src);
    }
    __syncwarp();
    if (lane == 0) {
        bool last = true;
        if (p.C > 1) {
            __threadfence();
            last = atomicAdd(h.push_counter, 1u) == (unsigned)p.C - 1u;
            if (last) { *h.push_counter = 0; __threadfence(); }
        }
        if (last) hf_st_release_sys(h.push_flag, h.push_val);
    }
}

// MS: a tile may hold up to MS WHOLE consecutive intervals when they are short (<= 32 steps together): phase P costs the
// same for 10 live lanes as for 32, and with 10-step intervals (BASELINE config 4) it was paid per interval.  MS = 1: every
// tile is a 32-step chunk of one interval.  The tile sequence of a block follows from iv_N alone (fit_intervals), so the
// table pipeline, the producer and the solver agree on it without talking.
template <int MS>
__device__ __forceinline__ int fit_intervals(const SolveParams &p, int j, int i2, int N0, int *b) {
    int ns = 0, tot = 0;
    for (int jj = j; jj <= i2 && ns < MS; ++jj) {
        const int n = jj == j ? N0 : p.iv_N[jj];
        if (n > CT - tot) break;
        b[ns++] = tot;
        tot += n;
    }
    b[ns] = tot;
    return ns;
}

template <class MD, bool PSI, bool PHILOX, bool REINV, bool FAST, bool SPLIT, int MODE, int MS>
__global__ void __launch_bounds__(256) solve_coop_kernel(const SolveParams p) {
    constexpr bool PARAM = MODE == MODE_PARAM;
    constexpr bool ACC_ONLY = MODE == MODE_RELAYOUT || MODE == MODE_LOGLIK;
    static_assert(!ACC_ONLY || (REINV && !PHILOX && !SPLIT), "accepted-path passes walk X, draw nothing, have no recursion");
    constexpr int D = MD::D, M = MD::M;
    using R = Rec<D, PSI>;
    constexpr int RS = R::SIZE;
    constexpr bool STORE_W = MODE == MODE_IMPUTE && !REINV;
    static_assert(!PARAM || (!PHILOX && !REINV), "the parameter step reads the stored accepted noise");
    using ZT = typename std::conditional<PHILOX, float, double>::type;
    static_assert(!FAST || (D == 2 && M == 1), "the fused step is FitzHugh-Nagumo's");

    constexpr int TD = SPLIT ? 4 : 2;   // table ring: SPLIT: tile q-1 (solver), q (producer), q+1 and q+2 in flight
    constexpr int NB = SPLIT ? 2 : 1;   // staging buffers per unit
    extern __shared__ __align__(256) double dsm[];
    __shared__ alignas(8) uint64_t full[TD];
    const int lane = threadIdx.x & 31, nthr = (int)blockDim.x;
    const int wpc = SPLIT ? nthr >> 6 : nthr >> 5;                   // units per CTA
    const int warp_raw = threadIdx.x >> 5;
    const bool do_P = !SPLIT || warp_raw >= wpc, do_S = !SPLIT || warp_raw < wpc;  // producer / solver role of this warp
    const int warp = SPLIT && warp_raw >= wpc ? warp_raw - wpc : warp_raw;          // unit of the CTA
    double *stab = dsm;                                              // [TD][CT][RS]  table records of a tile
    // per step: mixed increments (FAST: e_k, R_k), two factors of the accepted G dt, and (generic) the guiding offset
    // F = F0 + Psi x_b, which does not depend on the proposal state: phase P computes it lane-parallel, so that the serial
    // phase neither multiplies by Psi nor loads it (the same arithmetic on the same inputs: bit-identical)
    constexpr int SW = FAST ? 4 : M + 2 + D;
    double *sdw0 = dsm + TD * CT * RS + (size_t)warp * CT * (NB * SW + D);  // [NB][CT][SW]
    double *sxk = sdw0 + NB * CT * SW;                                       // [CT][D]   the tile's proposal states
    auto cta_sync = [&]() {
        if constexpr (SPLIT) asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");  // called from both roles' code paths
        else __syncthreads();
    };

    const int groups = (p.C + wpc - 1) / wpc;
    const int blk = (int)(blockIdx.x / groups);
    const int chain_raw = (int)(blockIdx.x % groups) * wpc + warp;
    const bool in_range = chain_raw < p.C;
    const int chain = in_range ? chain_raw : p.C - 1;  // surplus warps shadow the last chain, stores off
    const size_t uo = (size_t)chain * p.nblk + blk;
    const int i1 = p.blk_i1[blk], i2 = p.blk_i2[blk];
    const bool masked = MODE == MODE_IMPUTE && ((p.unit_mask && !p.unit_mask[uo]) || (p.blk_active && !p.blk_active[blk]));
    const bool active = in_range && !masked;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TD; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if constexpr (MODE == MODE_IMPUTE) {
        if (p.hf.blk == blk && p.hf.pull_box) {  // CTA-uniform: this is the edge block and a message is due
            if (do_S && in_range) hf_pull_unit(p, chain, lane, D);
            __syncthreads();  // producer warps read the accepted path too
        }
    }

    // ---- table pipeline (thread 0): one bulk copy per 32-step tile, one tile ahead ----
    int tj = i1, ttl = 0, tq = 0, tN = p.iv_N[i1], ttoff = p.iv_tile_off[i1];
    auto issue_table = [&]() {
        if (tj > i2) return;
        if constexpr (MS > 1) {
            if (ttl == 0 && tN <= CT) {  // whole short intervals: one bulk copy each, in lane order, on one barrier
                int b[MS + 1];
                const int ns = fit_intervals<MS>(p, tj, i2, tN, b);
                if (ns > 1) {
                    uint64_t *bar = &full[tq & (TD - 1)];
                    double *dst = stab + (size_t)(tq & (TD - 1)) * CT * RS;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(bar, (uint32_t)(b[ns] * RS * sizeof(double)));
                    for (int g = 0; g < ns; ++g) {
                        const int toff = g == 0 ? ttoff : p.iv_tile_off[tj + g];
                        bulk_g2s(dst + (size_t)b[g] * RS, p.table + (size_t)toff * TILE * RS, (uint32_t)((b[g + 1] - b[g]) * RS * sizeof(double)), bar);
                    }
                    ++tq;
                    tj += ns;
                    if (tj <= i2) { tN = p.iv_N[tj]; ttoff = p.iv_tile_off[tj]; }
                    return;
                }
            }
        }
        const int npad = ((tN + TILE - 1) >> 3) * TILE;  // records the table holds for this interval
        const int nrec = min(CT, npad - ttl * CT);
        const uint32_t bytes = (uint32_t)(nrec * RS * sizeof(double));
        uint64_t *bar = &full[tq & (TD - 1)];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, bytes);
        bulk_g2s(stab + (size_t)(tq & (TD - 1)) * CT * RS, p.table + ((size_t)ttoff * TILE + (size_t)ttl * CT) * RS, bytes, bar);
        ++tq;
        if (++ttl * CT >= tN) {
            ttl = 0;
            if (++tj <= i2) { tN = p.iv_N[tj]; ttoff = p.iv_tile_off[tj]; }
        }
    };
    if (threadIdx.x == 0) {
        issue_table();
        if constexpr (SPLIT) issue_table();
    }

    const bool pinned = PSI && p.blk_pinned[blk];
    double x[D], xb[D];
    load_start<D>(p, chain, i1, x);
#pragma unroll
    for (int c = 0; c < D; ++c) xb[c] = 0.0;
    if (pinned) load_end<D>(p, chain, i2, xb);
    [[maybe_unused]] double xa_prev[D];  // REINV: accepted state before the tile's first step
#pragma unroll
    for (int c = 0; c < D; ++c) xa_prev[c] = x[c];

    FhnK fk;
    if constexpr (FAST) fk.set_law(p.th);

    double llp;
    {
        // log h~(t_a, x_a) = -c - 1/2 x'Hx + F'x at the block's first grid point (loglikhd_obs)
        double rec[RS], F[D];
        const double *r0 = p.table_raw + (size_t)p.iv_tile_off[i1] * TILE * RS;
#pragma unroll
        for (int i = 0; i < RS; ++i) rec[i] = __ldg(r0 + i);
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
    }
    double lla = llp, lsum = 0.0, lsum_a = 0.0;
    bool ok = true;
    const uint32_t gchain = (uint32_t)(chain + p.chain_offset);
    PhiloxKeys pkeys;
    pkeys.set(p.seed);

    int q = 0;
    double *xp_end = nullptr;
    if constexpr (SPLIT) {
        if (do_S) {  // the solver runs one tile behind the producer
            cta_sync();
            if (threadIdx.x == 0) issue_table();
        }
    }
    // context of one interval: selectors, containers, aux law, pCN memory.  Loaded per interval; in a tile of several
    // short intervals every lane loads its own interval's for phase P, and phase S walks the intervals one after another
    int sx = 0, sw = 0;
    double aB[FAST ? 1 : D * D], abeta[FAST ? 1 : D], aatil[MD::CONST_SIGMA ? 1 : D * D];
    double rho = 1.0, lam = 0.0;
    const double *Xa = nullptr, *Wa = nullptr;
    double *Xp = nullptr, *Wp = nullptr;
    size_t wtile0 = 0;
    auto load_iv = [&](int jj) {
        sx = p.selX_in[(size_t)jj * p.C + chain];
        sw = p.selW_in[(size_t)jj * p.C + chain];
        if constexpr (FAST) {
            fk.set_interval(p.auxB + (size_t)jj * D * D, p.auxbeta + (size_t)jj * D);
        } else {
#pragma unroll
            for (int i = 0; i < D * D; ++i) aB[i] = p.auxB[(size_t)jj * D * D + i];
#pragma unroll
            for (int i = 0; i < D; ++i) abeta[i] = p.auxbeta[(size_t)jj * D + i];
        }
        if constexpr (!MD::CONST_SIGMA) {
#pragma unroll
            for (int i = 0; i < D * D; ++i) aatil[i] = p.auxatil[(size_t)jj * D * D + i];
        }
        rho = MODE != MODE_IMPUTE ? 1.0 : p.rho[jj];
        lam = MODE != MODE_IMPUTE ? 0.0 : sqrt(1.0 - rho * rho);
        Xa = p.X[sx] + p.iv_xoff[jj] + (size_t)chain * p.iv_xs[jj];      // accepted record
        Xp = p.X[1 - sx] + p.iv_xoff[jj] + (size_t)chain * p.iv_xs[jj];  // proposal record
        Wa = p.W[sw];
        Wp = p.W[1 - sw];
        wtile0 = (size_t)p.iv_tile_off[jj];
    };
    for (int j = i1; j <= i2; ++j) {
        const int N = p.iv_N[j], nt = (N + CT - 1) / CT;
        load_iv(j);
        // whole short intervals j .. j + ns - 1 in one tile (lanes bnd[g] .. bnd[g + 1] - 1 work on interval j + g)
        int ns = 1, bnd[MS + 1];
        bnd[0] = 0; bnd[1] = min(CT, N);
        if constexpr (MS > 1) {
            if (N <= CT) ns = fit_intervals<MS>(p, j, i2, N, bnd);
        }

        for (int tl = 0; tl < nt; ++tl, ++q) {
            cta_sync();  // every warp is done with its previous tile: that table stage and staging buffer are free
            if (threadIdx.x == 0) issue_table();
            mbar_wait(&full[q & (TD - 1)], (uint32_t)((q / TD) & 1));
            const double *srec = stab + (size_t)(q & (TD - 1)) * CT * RS;
            double *sdw = sdw0 + (SPLIT ? (q & 1) * CT * SW : 0);
            // this lane's interval and Euler step in phase P
            const bool multi = MS > 1 && ns > 1;
            int gl = 0;
            if constexpr (MS > 1) {
                if (multi) {
#pragma unroll
                    for (int i = 1; i < MS; ++i) gl += (i < ns && lane >= bnd[i]) ? 1 : 0;
                    if (gl > 0) load_iv(j + gl);  // the lanes of the first interval hold its context already
                }
            }
            int lane0 = 0;  // first lane of this lane's interval
            if constexpr (MS > 1) {
#pragma unroll
                for (int i = 1; i < MS; ++i) lane0 = (multi && i < ns && lane >= bnd[i]) ? bnd[i] : lane0;
            }
            const int k = multi ? lane - lane0 : tl * CT + lane;
            const int nlive = multi ? bnd[ns] : min(CT, N - tl * CT);
            const bool live = lane < nlive;
            double *const Xp_l = Xp;  // phase S walks the intervals one after another: the tile's states go out from here

            // ================= phase P: lane = step =================
            if (do_P) {
                const double *rp = srec + lane * RS;
                double dw[M];
#pragma unroll
                for (int w = 0; w < M; ++w) dw[w] = 0.0;
                double ga = 0.0, gb = 0.0;  // accepted-path Girsanov increment = ga * gb
                [[maybe_unused]] double xn[D];
                if constexpr (REINV) {
                    // accepted state before and after this lane's step; recover the accepted increment and G dt
                    double xa[D];
                    if (live) {
#pragma unroll
                        for (int c = 0; c < D; ++c) xn[c] = Xa[(size_t)k * D + c];
                    } else {
#pragma unroll
                        for (int c = 0; c < D; ++c) xn[c] = 0.0;
                    }
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        const double up = __shfl_up_sync(0xffffffffu, xn[c], 1);
                        xa[c] = lane == 0 ? xa_prev[c] : up;
                    }
                    if (live) {
                        if constexpr (FAST) {
                            double rec[RS];
                            load_rec_smem<RS>(rp, rec);
                            double Cf = rec[FD_C1], R0c = rec[FD_R0];
                            if constexpr (PSI) {
                                Cf = fma(rec[FD_Q10], xb[0], fma(rec[FD_Q11], xb[1], Cf));
                                R0c = fma(rec[FD_T00], xb[0], fma(rec[FD_T01], xb[1], R0c));
                            }
                            double g;
                            fhn_dll_terms(fk, rec, R0c, xa[0], xa[1], g, ga, gb);
                            const double pred = fma(rec[FD_A1], xa[1], fma(rec[FD_B1], xa[0], Cf));
                            dw[0] = (xn[1] - pred) * fk.isig;
                        } else {
                            double rec[RS], F[D], dr[D], res[D];
                            load_rec_smem<RS>(rp, rec);
                            guiding_F<D, PSI>(rec, xb, F);
                            guided_terms<MD, PSI>(p.th, rec, F, aB, abeta, aatil, xa, rec[0], dr, ga);
                            gb = rec[0];
#pragma unroll
                            for (int c = 0; c < D; ++c) res[c] = xn[c] - xa[c] - dr[c];
                            MD::sigma_solve(p.th, xa, res, dw);
                        }
                    }
                    // the next tile's first step starts from this tile's last accepted state
                    const int last = nlive - 1;
#pragma unroll
                    for (int c = 0; c < D; ++c) xa_prev[c] = __shfl_sync(0xffffffffu, xn[c], last);
                } else {
                    if (live) {
#pragma unroll
                        for (int w = 0; w < M; ++w) dw[w] = Wa[w_off(wtile0 + (size_t)(k >> 3), chain, w, p.C, M) + (k & 7)];
                    }
                }
                // pCN: W deg = sqrt(1 - rho^2) W_fresh + rho W   (a2)
                [[maybe_unused]] ZT z[M];
#pragma unroll
                for (int w = 0; w < M; ++w) {
                    if constexpr (MODE != MODE_IMPUTE) {
                        z[w] = 0.0;
                    } else if constexpr (PHILOX) {
                        z[w] = philox_normal1f(pkeys, p.iter, (uint32_t)(j + gl + p.interval_offset), gchain, (uint32_t)w, (uint32_t)k);
                    } else {
                        z[w] = live ? p.Z[w_off(wtile0 + (size_t)(k >> 3), chain, w, p.C, M) + (k & 7)] : 0.0;
                    }
                }
                if (live) {
                    if constexpr (MODE != MODE_IMPUTE) {
                        // the accepted increments as they are
                        if (MODE == MODE_RELAYOUT && in_range) {  // materialise them in the accepted noise container
#pragma unroll
                            for (int w = 0; w < M; ++w) p.W[sw][w_off(wtile0 + (size_t)(k >> 3), chain, w, p.C, M) + (k & 7)] = dw[w];
                        }
                    } else if constexpr (FAST) {
                        dw[0] = fma(lam * rp[FD_SQ], (double)z[0], rho * dw[0]);
                    } else {
#pragma unroll
                        for (int w = 0; w < M; ++w) dw[w] = lam * (rp[1] * (double)z[w]) + rho * dw[w];
                    }
                    if (STORE_W && active) {
#pragma unroll
                        for (int w = 0; w < M; ++w) Wp[w_off(wtile0 + (size_t)(k >> 3), chain, w, p.C, M) + (k & 7)] = dw[w];
                    }
                }
                if constexpr (FAST) {
                    // everything of the step that does not depend on the proposal state: e_k = sigma dW deg_k + C_k, R_k
                    double Cf = rp[FD_C1], R0c = rp[FD_R0];
                    if constexpr (PSI) {
                        Cf = fma(rp[FD_Q10], xb[0], fma(rp[FD_Q11], xb[1], Cf));
                        R0c = fma(rp[FD_T00], xb[0], fma(rp[FD_T01], xb[1], R0c));
                    }
                    sdw[lane * SW] = fma(fk.sig, dw[0], Cf);
                    sdw[lane * SW + 1] = R0c;
                    if constexpr (REINV) { sdw[lane * SW + 2] = ga; sdw[lane * SW + 3] = gb; }
                } else {
#pragma unroll
                    for (int w = 0; w < M; ++w) sdw[lane * SW + w] = dw[w];
                    if constexpr (REINV) { sdw[lane * SW + M] = ga; sdw[lane * SW + M + 1] = gb; }
                    if constexpr (!ACC_ONLY) {
                        double recf[RS], Fl[D];
                        load_rec_smem<RS>(rp, recf);
                        guiding_F<D, PSI>(recf, xb, Fl);
#pragma unroll
                        for (int c = 0; c < D; ++c) sdw[lane * SW + M + 2 + c] = Fl[c];
                    }
                }
            }
            if constexpr (!SPLIT) __syncwarp();
            if (!do_S) continue;

            if constexpr (ACC_ONLY) {  // no proposal: only the accepted path's Girsanov sum, in serial order
                for (int s = 0; s < nlive; ++s) lsum_a = fma(sdw[s * SW + (FAST ? 2 : M)], sdw[s * SW + (FAST ? 3 : M + 1)], lsum_a);
                __syncwarp();
                continue;
            }
            // ================= phase S: the recursion, all lanes in step =================
            // The operands of step s+1 (table record, staged increments) are loaded BEFORE step s's state is stored:
            // shared-memory loads are never moved above a shared store whose address they might alias, so in source
            // order their ~30-cycle latency would sit in the recurrence of every step.
            const uint32_t sxk_a = smem_u32(sxk);
            // FAST: A1, B1, dt H11, dt H12, dt/eps, (sqrt dt) is all the recursion reads; generic: dt, sqrt dt and H
            constexpr int NREC = FAST ? 6 : ((R::OFF_F + 1) & ~1);
            double rec[NREC], sv[SW];
            auto load_ops = [&](int s, double *r, double *v) {
                load_rec_smem<NREC>(srec + s * RS, r);
                if constexpr (SW % 2 == 0) {
#pragma unroll
                    for (int i = 0; i < SW / 2; ++i) {
                        const double2 t = *reinterpret_cast<const double2 *>(sdw + s * SW + 2 * i);
                        v[2 * i] = t.x; v[2 * i + 1] = t.y;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < SW; ++i) v[i] = sdw[s * SW + i];
                }
            };
            // FAST: two steps ahead -- one step of the fused recursion is ~30 issue cycles, about one shared-load latency;
            // the generic step is long enough for one step ahead (and its 20-double records would not fit twice)
            constexpr int AHEAD = FAST ? 2 : 1;
            [[maybe_unused]] double rec1[AHEAD == 2 ? NREC : 1], sv1[AHEAD == 2 ? SW : 1];
            load_ops(0, rec, sv);
            if constexpr (AHEAD == 2) load_ops(1, rec1, sv1);
            int s_beg = 0;
            // aux law of the interval the recursion walks (uniform); the next interval's is fetched while this one runs
            [[maybe_unused]] double aBn[FAST ? 2 : D * D], abn[FAST ? 1 : D], aan[MD::CONST_SIGMA ? 1 : D * D];
            auto fetch_aux = [&](int jj) {
                if constexpr (FAST) {
                    aBn[0] = p.auxB[(size_t)jj * D * D];
                    aBn[1] = p.auxbeta[(size_t)jj * D];
                } else {
#pragma unroll
                    for (int i = 0; i < D * D; ++i) aBn[i] = p.auxB[(size_t)jj * D * D + i];
#pragma unroll
                    for (int i = 0; i < D; ++i) abn[i] = p.auxbeta[(size_t)jj * D + i];
                }
                if constexpr (!MD::CONST_SIGMA) {
#pragma unroll
                    for (int i = 0; i < D * D; ++i) aan[i] = p.auxatil[(size_t)jj * D * D + i];
                }
            };
            auto take_aux = [&]() {
                if constexpr (FAST) {
                    fk.aB0 = aBn[0];
                    fk.c2 = aBn[1] - fk.s * fk.ie;
                } else {
#pragma unroll
                    for (int i = 0; i < D * D; ++i) aB[i] = aBn[i];
#pragma unroll
                    for (int i = 0; i < D; ++i) abeta[i] = abn[i];
                }
                if constexpr (!MD::CONST_SIGMA) {
#pragma unroll
                    for (int i = 0; i < D * D; ++i) aatil[i] = aan[i];
                }
            };
            if constexpr (MS > 1) {
                if (multi) fetch_aux(j);
            }
            for (int gs = 0; gs < (multi ? ns : 1); ++gs) {
            if constexpr (MS > 1) {
                if (multi) {
                    take_aux();
                    if (gs + 1 < ns) fetch_aux(j + gs + 1);
                }
            }
            const int s_end = multi ? s_beg + p.iv_N[j + gs] : nlive;
#pragma unroll 4
            for (int s = s_beg; s < s_end; ++s) {
                double recn[NREC], svn[SW];
                load_ops(min(s + AHEAD, CT - 1), recn, svn);
                if constexpr (FAST) {
                    if constexpr (REINV) lsum_a = fma(sv[2], sv[3], lsum_a);
                    fhn_step(fk, rec, sv[1], sv[0], x[0], x[1], lsum);  // sv = e_k, R_k, (two factors of the accepted G dt)
                } else {
                    if constexpr (REINV) lsum_a = fma(sv[M], sv[M + 1], lsum_a);
                    double F[D], dr[D], sdwv[D], gg;
#pragma unroll
                    for (int c = 0; c < D; ++c) F[c] = sv[M + 2 + c];
                    guided_terms<MD, PSI>(p.th, rec, F, aB, abeta, aatil, x, rec[0], dr, gg);
                    lsum = fma(gg, rec[0], lsum);
                    MD::sigma_mul(p.th, x, sv, sdwv);
#pragma unroll
                    for (int c = 0; c < D; ++c) x[c] = x[c] + dr[c] + sdwv[c];
                    ok = ok && MD::bound_ok(x);
                }
                if constexpr (AHEAD == 2) {
#pragma unroll
                    for (int i = 0; i < NREC; ++i) { rec[i] = rec1[i]; rec1[i] = recn[i]; }
#pragma unroll
                    for (int i = 0; i < SW; ++i) { sv[i] = sv1[i]; sv1[i] = svn[i]; }
                } else {
#pragma unroll
                    for (int i = 0; i < NREC; ++i) rec[i] = recn[i];
#pragma unroll
                    for (int i = 0; i < SW; ++i) sv[i] = svn[i];
                }
                // every lane holds the same state: one (same-address) shared store per step.  No "memory" clobber: the
                // stage is read only after the loop (behind __syncwarp), and the loop's table / increment loads must stay
                // free to move ahead of these stores
                if constexpr (D == 2) {
                    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(sxk_a + (uint32_t)s * 16u), "d"(x[0]), "d"(x[1]));
                } else {
#pragma unroll
                    for (int c = 0; c < D; ++c)
                        asm volatile("st.shared.f64 [%0], %1;" ::"r"(sxk_a + (uint32_t)(s * D + c) * 8u), "d"(x[c]));
                }
            }
            s_beg = s_end;
            }
            __syncwarp();
            if (live && active) {  // the tile's states: one coalesced store (per interval)
                double *dst = Xp_l + (size_t)k * D;
                if constexpr (D == 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(sxk + lane * D);
                    st_v2(dst, v.x, v.y);
                } else {
#pragma unroll
                    for (int c = 0; c < D; ++c) dst[c] = sxk[lane * D + c];
                }
            }
        }
        j += ns - 1;  // the tile has dealt with ns whole intervals
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if constexpr (SPLIT) {
        if (!do_S) {  // producer: the solver's last tile is still to come
            cta_sync();
            return;
        }
    }
    llp += lsum;
    lla += lsum_a;
    if (!ok) llp = -INFINITY;
    if constexpr (ACC_ONLY) {
        if (in_range && lane == 0) p.ll[uo] = lla;
        return;
    }
    if (pinned && active && lane == 0) {  // pin the proposal's last point to x_b (glued paths stay continuous)
        xp_end = p.X[1 - p.selX_in[(size_t)i2 * p.C + chain]] + p.iv_xoff[i2] + (size_t)chain * p.iv_xs[i2] + (size_t)(p.iv_N[i2] - 1) * D;
#pragma unroll
        for (int c = 0; c < D; ++c) xp_end[c] = xb[c];
    }

    if constexpr (PARAM) {
        if (in_range && lane == 0) {
            if (p.out_llprop) p.out_llprop[uo] = llp;
            if (p.out_acc) p.out_acc[uo] = (uint8_t)ok;
        }
        return;
    }
    double llcur = REINV ? lla : p.ll[uo];
    const double e = PHILOX ? philox_exponential(p.seed, p.iter, (uint32_t)(i1 + p.interval_offset), gchain) : p.E[uo];
    // accepted = rand(Exponential(1)) > -(ll deg - ll)   (eMCMC.accept_reject!)
    const bool acc = active && (p.force_accept ? ok : (e > -(llp - llcur)));
    if (in_range) {
        for (int j = i1 + lane; j <= i2; j += 32) {  // swap_paths!: flip the accepted-container selectors
            const size_t o = (size_t)j * p.C + chain;
            p.selX_out[o] = p.selX_in[o] ^ (uint8_t)acc;
            p.selW_out[o] = p.selW_in[o] ^ (uint8_t)(acc && STORE_W);
        }
        if (p.hf.blk == blk && p.hf.push_box) hf_push_unit(p, chain, lane, D, acc);  // the neighbour's halo / start point
    }
    if (lane == 0) {
        if (p.acc_count && acc) atomicAdd(p.acc_count + blk, 1ull);
        if (p.prop_count && active) atomicAdd(p.prop_count + blk, 1ull);
        if (active) {
            if (acc) llcur = llp;
            p.ll[uo] = llcur;
            if (p.out_llprop) p.out_llprop[(size_t)chain * p.out_stride + blk] = llp;
            if (p.out_ll) p.out_ll[(size_t)chain * p.out_stride + blk] = llcur;
            if (p.out_acc) p.out_acc[(size_t)chain * p.out_stride + blk] = (uint8_t)acc;
        } else if (in_range && p.blk_active && !p.blk_active[blk]) {  // block owned by a neighbouring shard
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            if (p.out_llprop) p.out_llprop[(size_t)chain * p.out_stride + blk] = nan;
            if (p.out_ll) p.out_ll[(size_t)chain * p.out_stride + blk] = nan;
            if (p.out_acc) p.out_acc[(size_t)chain * p.out_stride + blk] = 0;
        }
    }
}

template <class MD, bool PSI, bool PHILOX, bool REINV, bool FAST, bool SPLIT, int MODE, int MS>
static cudaError_t launch_coop_ms(const SolveParams &p, cudaStream_t s) {
    constexpr int RS = Rec<MD::D, PSI>::SIZE;
    constexpr int TD = SPLIT ? 4 : 2, NB = SPLIT ? 2 : 1, SW = FAST ? 4 : MD::M + 2 + MD::D;
    const int wpc = std::min(4, (int)p.C);  // units of one block (consecutive chains) share the table stage
    const unsigned groups = (unsigned)((p.C + wpc - 1) / wpc);
    const unsigned grid = groups * (unsigned)p.nblk;
    const size_t smem = sizeof(double) * ((size_t)TD * CT * RS + (size_t)wpc * CT * (NB * SW + MD::D));
    auto kern = solve_coop_kernel<MD, PSI, PHILOX, REINV, FAST, SPLIT, MODE, MS>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3((unsigned)((SPLIT ? 64 : 32) * wpc));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

// short intervals (fewer than 16 Euler steps on average): up to four whole intervals per tile
template <class MD, bool PSI, bool PHILOX, bool REINV, bool FAST, bool SPLIT, int MODE>
static cudaError_t launch_coop_split(const SolveParams &p, cudaStream_t s) {
    const char *env = getenv("DMCMC_COOP_MULTI");  // A/B override, read per launch (the tests switch it)
    const int env_ms = env ? atoi(env) : -1;
    const bool multi = env_ms >= 0 ? env_ms != 0 : p.mean_N <= 16;
    return multi ? launch_coop_ms<MD, PSI, PHILOX, REINV, FAST, SPLIT, MODE, 4>(p, s)
                 : launch_coop_ms<MD, PSI, PHILOX, REINV, FAST, SPLIT, MODE, 1>(p, s);
}

template <class MD, bool PSI, bool PHILOX, bool REINV, bool FAST, int MODE = MODE_IMPUTE>
static cudaError_t launch_coop_one(const SolveParams &p, cudaStream_t s) {
    // a producer warp per unit pays off when tiles are reasonably full (intervals of >= 24 Euler steps) or when there
    // are so few units that nothing else hides phase P; with ~10^3 units of short intervals (Lorenz-63 at 10 steps per
    // interval) the other warps already hide it and the extra hand-over per tile costs more than it saves
    static const int env_split = getenv("DMCMC_COOP_SPLIT") ? atoi(getenv("DMCMC_COOP_SPLIT")) : -1;  // A/B override
    const bool split = env_split >= 0 ? env_split != 0 : (p.mean_N >= 24 || (long long)p.C * p.nblk <= 2 * 148);
    if constexpr (MODE == MODE_RELAYOUT || MODE == MODE_LOGLIK) {
        return launch_coop_split<MD, PSI, PHILOX, REINV, FAST, false, MODE>(p, s);
    } else {
        return split ? launch_coop_split<MD, PSI, PHILOX, REINV, FAST, true, MODE>(p, s)
                     : launch_coop_split<MD, PSI, PHILOX, REINV, FAST, false, MODE>(p, s);
    }
}

template <class MD, bool FAST>
static cudaError_t launch_coop_model(int mode, const SolveParams &p, cudaStream_t s) {
    if (mode == MODE_PARAM)
        return p.has_psi ? launch_coop_one<MD, true, false, false, FAST, MODE_PARAM>(p, s)
                         : launch_coop_one<MD, false, false, false, FAST, MODE_PARAM>(p, s);
    if (mode == MODE_RELAYOUT)
        return p.has_psi ? launch_coop_one<MD, true, false, true, FAST, MODE_RELAYOUT>(p, s)
                         : launch_coop_one<MD, false, false, true, FAST, MODE_RELAYOUT>(p, s);
    if (mode == MODE_LOGLIK)
        return p.has_psi ? launch_coop_one<MD, true, false, true, FAST, MODE_LOGLIK>(p, s)
                         : launch_coop_one<MD, false, false, true, FAST, MODE_LOGLIK>(p, s);
    const bool philox = (p.Z == nullptr), reinv = p.reinvert != 0;
#define DMCMC_COOP(PSI_)                                                                              \
    if (philox) return reinv ? launch_coop_one<MD, PSI_, true, true, FAST>(p, s)                       \
                             : launch_coop_one<MD, PSI_, true, false, FAST>(p, s);                     \
    return reinv ? launch_coop_one<MD, PSI_, false, true, FAST>(p, s)                                  \
                 : launch_coop_one<MD, PSI_, false, false, FAST>(p, s);
    if (p.has_psi) { DMCMC_COOP(true) }
    DMCMC_COOP(false)
#undef DMCMC_COOP
}

// every solve mode; the caller (launch_solve) has decided that the launch has few units
cudaError_t launch_solve_coop(int model, int mode, const SolveParams &p, cudaStream_t s) {
    if ((long long)p.nblk * p.C == 0) return cudaSuccess;
    switch (model) {
    case 0: return p.fast_ok ? launch_coop_model<ModelFHN, true>(mode, p, s) : launch_coop_model<ModelFHN, false>(mode, p, s);
    case 1: return launch_coop_model<ModelLV, false>(mode, p, s);
    case 2: return launch_coop_model<ModelL63, false>(mode, p, s);
    case 4: return launch_coop_model<ModelLIN2, false>(mode, p, s);
    case 5: return launch_coop_model<ModelLVM, false>(mode, p, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace dmcmc
