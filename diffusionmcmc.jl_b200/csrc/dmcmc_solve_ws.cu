// dmcmc_solve_ws.cu -- warp-specialised imputation kernel for the production blocked sweep of
// FitzHugh-Nagumo (fused step, Philox noise, layout switch fused in: solve_kernel<ModelFHN, PSI,
// MODE_IMPUTE, PHILOX, REINV, FAST> is the single-role statement of the same arithmetic).
//
// The blocked C3 sweep has only chains x blocks = 5.2e4 units (11 warps per SM): the one-thread-per-unit
// kernel is bounded by the latency of ONE warp's 110-instruction step, not by a pipe or by HBM.  Of
// that step only the Euler recursion itself (x_{k+1} from x_k) is serial.  Here every unit is worked on
// by two lanes of two different warps of the CTA:
//   producer warp : walks the accepted path (one 128-byte TMA bulk copy per lane and tile), recovers the accepted
//                   noise increment and the accepted-path Girsanov sum, draws the Philox normals of the tile's
//                   first four steps (the last four come from the solver, see below), applies the pCN mix and
//                   leaves, IN PLACE of the 16-byte accepted state it has just consumed,
//                       ( e_k = sigma * dW deg_k + C_k ,  R_k = dt (F0_1 + Psi_1. x_b) )
//                   one tile ahead of
//   solver warp   : runs the recursion and its Girsanov sum from (e_k, R_k) and the first six table
//                   coefficients, stores the proposal, takes the accept decision; in the shadow of the
//                   recursion's dependency chain it draws the normals of steps 4..7 of the tile the producer
//                   will work on next.
// Per step the roles execute ~50 and ~40 instructions instead of 110 in one warp, both fit 80 registers, so
// the 22 warps per SM the blocked sweep now needs are all resident; one CTA barrier per 8-step tile is the
// only hand-over.  Shared memory per warp pair: a ring of three 4.5 KB slots (32 chains x (8 steps x 16 B + 16 B
// pad: conflict-free LDS.128/STS.128)): slot q-1 is read by the solver, q is rewritten by the producer, q+1 is
// in flight.  Results are bit-identical to the single-role kernel (same operations in the same order;
// tests/test_gpu_parity.py::test_warp_specialised_matches_single_role).
//
// Reference rows: a2-a7 of SURVEY.md section 8a (src/run!_alterations.jl:5-47, :81-91).
#include <cstdlib>

#include "dmcmc_solve_common.cuh"

namespace dmcmc {

constexpr int WS_TD = 4;    // table ring: tile q-1 (solver), q (producer), q+1 and q+2 in flight
constexpr int WS_XD = 3;    // path ring
constexpr int WTS = 8;      // Euler steps per tile: 128 contiguous bytes per chain
constexpr int XREC = WTS * 16 + 16;   // bytes per chain in a slot (padded: lane r, step k -> bank group (r + k) mod 8)
constexpr int XSLOT = 32 * XREC;      // bytes per ring slot (32 chains)
constexpr int WS_MAXNP = 4;

__device__ __forceinline__ void cta_bar(int nthreads) {
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

template <bool PSI>
__global__ void __launch_bounds__(128, 6) solve_fhn_ws_kernel(const SolveParams p) {
    constexpr int D = 2;
    using R = Rec<2, PSI>;
    constexpr int RS = R::SIZE;

    extern __shared__ __align__(256) double dsm[];
    __shared__ alignas(8) uint64_t full[WS_TD];
    __shared__ alignas(8) uint64_t xfull_all[WS_MAXNP][WS_XD];
    const int nthr = (int)blockDim.x, np = nthr >> 6;  // np solver warps, then np producer warps
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool producer = warp >= np;
    const int pair = producer ? warp - np : warp;
    double *stab = dsm;                                                    // [WS_TD][WTS][RS]
    unsigned char *pbytes = reinterpret_cast<unsigned char *>(dsm + WS_TD * WTS * RS) + (size_t)pair * (WS_XD * XSLOT + 1024 + 256);
    const uint32_t xring = smem_u32(pbytes);                               // [WS_XD][32][XREC]
    float *zbuf = reinterpret_cast<float *>(pbytes + WS_XD * XSLOT);       // [2][4][32]
    double *llsh = reinterpret_cast<double *>(pbytes + WS_XD * XSLOT + 1024);  // [32]
    uint64_t *xfull = xfull_all[pair];

    const int cpc = 32 * np;  // chains per CTA
    const int groups = (p.C + cpc - 1) / cpc;
    const int blk = (int)(blockIdx.x / groups);
    const int chain_raw = (int)(blockIdx.x % groups) * cpc + pair * 32 + lane;
    const bool in_range = chain_raw < p.C;
    const int chain = in_range ? chain_raw : p.C - 1;  // out-of-range lanes shadow the last chain, stores off
    const size_t uo = (size_t)chain * p.nblk + blk;
    const int i1 = p.blk_i1[blk], i2 = p.blk_i2[blk];
    const bool masked = (p.unit_mask && !p.unit_mask[uo]) || (p.blk_active && !p.blk_active[blk]);
    const bool active = in_range && !masked;

    if (threadIdx.x == 0) {
        for (int i = 0; i < WS_TD; ++i) mbar_init(&full[i], 1);
        for (int i = 0; i < np * WS_XD; ++i) mbar_init(&xfull_all[0][0] + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const bool pinned = PSI && p.blk_pinned[blk];
    double xb[D] = {0.0, 0.0};
    if (pinned) load_end<D>(p, chain, i2, xb);
    FhnK fk;
    fk.set_law(p.th);
    const uint32_t gchain = (uint32_t)(chain + p.chain_offset);
    PhiloxKeys pkeys;
    pkeys.set(p.seed);

    if (producer) {
        // ================= producer: accepted path -> accepted noise, pCN mix, accepted ll =================
        double xa[D];
        load_start<D>(p, chain, i1, xa);
        double lsum_a = 0.0;

        // input pipeline (issue side): every lane fetches its own chain's 128 bytes of the tile with one bulk copy
        const double *psrc = nullptr;
        int ij = i1, itl = 0, iN = 0, ist = 0;
        auto setup_issue = [&](int j) {
            iN = p.iv_N[j];
            const int sel = p.selX_in[(size_t)j * p.C + chain];
            psrc = p.X[sel] + (size_t)p.iv_xoff[j] + (size_t)chain * p.iv_xs[j];
        };
        auto issue_in = [&]() {
            if (ij <= i2) {
                const uint32_t bytes = (uint32_t)min(WTS, iN - itl * WTS) * 16u;
                uint64_t *bar = &xfull[ist];
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (lane == 0) mbar_expect_tx(bar, bytes * 32u);
                bulk_g2s(pbytes + (size_t)ist * XSLOT + (size_t)lane * XREC, psrc, bytes, bar);
                psrc += WTS * D;
                if (++itl * WTS >= iN) {
                    itl = 0;
                    if (++ij <= i2) setup_issue(ij);
                }
            }
            ist = ist + 1 == WS_XD ? 0 : ist + 1;
        };
        setup_issue(i1);
        issue_in();  // tile 0; tile q + 1 is issued when tile q is taken up

        int q = 0, cst = 0;
        uint32_t cph = 0;  // slot and phase parity of the tile being consumed
        for (int j = i1; j <= i2; ++j) {
            const int N = p.iv_N[j], nt = (N + WTS - 1) / WTS;
            fk.aB0 = p.auxB[(size_t)j * D * D];
            fk.c2 = p.auxbeta[(size_t)j * D] - fk.s * fk.ie;
            const double rho = p.rho[j], lam = sqrt(1.0 - rho * rho);
            for (int tl = 0; tl < nt; ++tl, ++q) {
                cta_bar(nthr);  // the solver has left slot (q+1) % 3 (tile q-2) and has written this tile's normals
                issue_in();
                mbar_wait(&full[q & (WS_TD - 1)], (uint32_t)((q / WS_TD) & 1));
                const double *srec = stab + (size_t)(q & (WS_TD - 1)) * WTS * RS;
                mbar_wait(&xfull[cst], cph);
                const uint32_t cur = xring + (uint32_t)cst * XSLOT + (uint32_t)lane * XREC;
                if (++cst == WS_XD) { cst = 0; cph ^= 1u; }
                const float *zb = zbuf + (q & 1) * 128 + lane;

                auto tile_fn = [&](auto check_tag) {
                    constexpr bool CHECK = decltype(check_tag)::value;
                    const int kh = tl * WTS;
                    float z[WTS];
                    philox_normal4f(pkeys, p.iter, (uint32_t)(j + p.interval_offset), gchain, 0u, (uint32_t)(kh >> 2), z);
#pragma unroll
                    for (int s = 4; s < WTS; ++s) z[s] = zb[(s - 4) * 32];
#pragma unroll
                    for (int s = 0; s < WTS; ++s) {
                        const bool live = !CHECK || (kh + s < N);
                        if (live) {
                            double rec[RS];
                            load_rec_smem<RS>(srec + s * RS, rec);
                            double Cf = rec[FD_C1], R0c = rec[FD_R0];
                            if constexpr (PSI) {
                                Cf = fma(rec[FD_Q10], xb[0], fma(rec[FD_Q11], xb[1], Cf));
                                R0c = fma(rec[FD_T00], xb[0], fma(rec[FD_T01], xb[1], R0c));
                            }
                            const uint32_t a = cur + (uint32_t)s * 16u;
                            double xn[2], g;
                            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(xn[0]), "=d"(xn[1]) : "r"(a));
                            fhn_dll(fk, rec, R0c, xa[0], xa[1], g, lsum_a);
                            const double pred = fma(rec[FD_A1], xa[1], fma(rec[FD_B1], xa[0], Cf));
                            const double dwa = (xn[1] - pred) * fk.isig;
                            xa[0] = xn[0]; xa[1] = xn[1];
                            const double dw = fma(lam * rec[FD_SQ], (double)z[s], rho * dwa);  // pCN
                            const double ev = fma(fk.sig, dw, Cf);
                            asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(ev), "d"(R0c) : "memory");
                        }
                    }
                };
                if (tl * WTS + WTS <= N) tile_fn(std::false_type{});
                else tile_fn(std::true_type{});
            }
        }
        llsh[lane] = lsum_a;
        cta_bar(nthr);  // the solver's last tile
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        return;
    }

    // ================= solver: table pipeline, recursion, proposal ll, stores, accept/swap =================
    int tj = i1, ttl = 0, tq = 0, tN = p.iv_N[i1], ttoff = p.iv_tile_off[i1];
    auto issue_table = [&]() {  // thread 0 only
        if (tj > i2) return;
        constexpr uint32_t bytes = (uint32_t)(WTS * RS * sizeof(double));  // the table holds whole 8-step tiles
        uint64_t *bar = &full[tq & (WS_TD - 1)];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, bytes);
        bulk_g2s(stab + (size_t)(tq & (WS_TD - 1)) * WTS * RS, p.table + ((size_t)ttoff + ttl) * WTS * RS, bytes, bar);
        ++tq;
        if (++ttl * WTS >= tN) {
            ttl = 0;
            if (++tj <= i2) { tN = p.iv_N[tj]; ttoff = p.iv_tile_off[tj]; }
        }
    };
    if (threadIdx.x == 0) { issue_table(); issue_table(); }

    // normals of steps 4..7 of the tile the producer works on next (Philox quad 2 tl + 1 of the interval)
    int zj = i1, ztl = 0, znt = (p.iv_N[i1] + WTS - 1) / WTS, zq = 0;
    auto make_normals = [&]() {
        if (zj <= i2) {
            float z[4];
            philox_normal4f(pkeys, p.iter, (uint32_t)(zj + p.interval_offset), gchain, 0u, (uint32_t)(2 * ztl + 1), z);
            float *zb = zbuf + (zq & 1) * 128 + lane;
#pragma unroll
            for (int k = 0; k < 4; ++k) zb[k * 32] = z[k];
            if (++ztl == znt) {
                ztl = 0;
                if (++zj <= i2) znt = (p.iv_N[zj] + WTS - 1) / WTS;
            }
        }
        ++zq;
    };
    make_normals();  // tile 0

    double x[D];
    load_start<D>(p, chain, i1, x);
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
    double lla = llp, lsum = 0.0;
    uint32_t sxbits = 0;
    double *xp_end = nullptr;

    cta_bar(nthr);  // producers work on tile 0
    if (threadIdx.x == 0) issue_table();
    make_normals();  // tile 1
    int q = 0, cst = 0;
    for (int j = i1; j <= i2; ++j) {
        const int N = p.iv_N[j], nt = (N + WTS - 1) / WTS;
        const int sx = p.selX_in[(size_t)j * p.C + chain];
        if (j - i1 < 32) sxbits |= (uint32_t)sx << (j - i1);
        fk.aB0 = p.auxB[(size_t)j * D * D];
        fk.c2 = p.auxbeta[(size_t)j * D] - fk.s * fk.ie;
        double *Xp = p.X[1 - sx] + p.iv_xoff[j] + (size_t)chain * p.iv_xs[j];
        xp_end = Xp + (size_t)(N - 1) * D;
        for (int tl = 0; tl < nt; ++tl, ++q) {
            cta_bar(nthr);  // the producer has finished tile q
            if (threadIdx.x == 0) issue_table();
            make_normals();  // tile q + 2
            mbar_wait(&full[q & (WS_TD - 1)], (uint32_t)((q / WS_TD) & 1));
            const double *srec = stab + (size_t)(q & (WS_TD - 1)) * WTS * RS;
            const uint32_t cur = xring + (uint32_t)cst * XSLOT + (uint32_t)lane * XREC;
            cst = cst + 1 == WS_XD ? 0 : cst + 1;

            auto tile_fn = [&](auto check_tag) {
                constexpr bool CHECK = decltype(check_tag)::value;
                const int kh = tl * WTS;
                double *xph = Xp + (size_t)kh * D;
                double xprev[D] = {0.0, 0.0};
#pragma unroll
                for (int s = 0; s < WTS; ++s) {
                    const bool live = !CHECK || (kh + s < N);
                    if (live) {
                        double rec[6];
                        load_rec_smem<6>(srec + s * RS, rec);  // A1, B1, dt H11, dt H12, dt/eps, (sqrt dt)
                        double ev, R0c, g;
                        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(ev), "=d"(R0c) : "r"(cur + (uint32_t)s * 16u));
                        fhn_dll(fk, rec, R0c, x[0], x[1], g, lsum);
                        const double ynew = fma(rec[FD_KIE], g + (fk.s - x[1]), x[0]);
                        x[1] = fma(rec[FD_A1], x[1], fma(rec[FD_B1], x[0], ev));
                        x[0] = ynew;
                    }
                    if (s & 1) {
                        if (live && active) st_v4(xph + (s - 1) * D, xprev[0], xprev[1], x[0], x[1]);
                    } else {
                        if (CHECK && live && kh + s == N - 1 && active) {  // odd N: the last state goes out alone
                            xph[s * D] = x[0];
                            xph[s * D + 1] = x[1];
                        }
                        xprev[0] = x[0]; xprev[1] = x[1];
                    }
                }
            };
            if (tl * WTS + WTS <= N) tile_fn(std::false_type{});
            else tile_fn(std::true_type{});
        }
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    llp += lsum;
    lla += llsh[lane];  // written before the producers' last barrier
    if (pinned && active) {  // pin the proposal's last point to x_b (glued paths stay continuous)
        xp_end[0] = xb[0];
        xp_end[1] = xb[1];
    }
    const double e = philox_exponential(p.seed, p.iter, (uint32_t)(i1 + p.interval_offset), gchain);
    // accepted = rand(Exponential(1)) > -(ll deg - ll)   (eMCMC.accept_reject!)
    const bool acc = active && (p.force_accept ? true : (e > -(llp - lla)));
    if (in_range) {
        for (int j = i1; j <= i2; ++j) {  // swap_paths!: flip the accepted-container selectors
            const size_t o = (size_t)j * p.C + chain;
            const int b = j - i1;
            const uint8_t sxj = b < 32 ? (uint8_t)((sxbits >> b) & 1u) : p.selX_in[o];
            p.selX_out[o] = sxj ^ (uint8_t)acc;
            p.selW_out[o] = p.selW_in[o];
        }
    }
    const unsigned accm = __ballot_sync(0xffffffffu, acc), actm = __ballot_sync(0xffffffffu, active);
    if (lane == 0) {
        if (p.acc_count && accm) atomicAdd(p.acc_count + blk, (unsigned long long)__popc(accm));
        if (p.prop_count && actm) atomicAdd(p.prop_count + blk, (unsigned long long)__popc(actm));
    }
    if (active) {
        const double llcur = acc ? llp : lla;
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

template <bool PSI>
static cudaError_t launch_ws(const SolveParams &p, int np, cudaStream_t s) {
    constexpr int RS = Rec<2, PSI>::SIZE;
    const int cpc = 32 * np;
    const unsigned groups = (unsigned)((p.C + cpc - 1) / cpc);
    const unsigned grid = groups * (unsigned)p.nblk;
    const size_t smem = sizeof(double) * (size_t)WS_TD * WTS * RS + (size_t)np * (WS_XD * XSLOT + 1024 + 256);
    auto kern = solve_fhn_ws_kernel<PSI>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3((unsigned)(64 * np));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

// threads = the caller's CTA size for the single-role kernel (chains per CTA); here it selects the number
// of solver/producer warp pairs: 32 -> one pair (64 threads), otherwise two pairs (128 threads)
cudaError_t launch_solve_fhn_ws(const SolveParams &p, int threads, cudaStream_t s) {
    if ((long long)p.nblk * p.C == 0) return cudaSuccess;
    const long long units = (long long)p.C * p.nblk;
    static const int env_np = getenv("DMCMC_WS_NP") ? atoi(getenv("DMCMC_WS_NP")) : 0;  // tuning override
    const int np = env_np ? std::min(env_np, WS_MAXNP) : ((threads <= 32 || units < 148LL * 64 * 2) ? 1 : 2);
    return p.has_psi ? launch_ws<true>(p, np, s) : launch_ws<false>(p, np, s);
}

}  // namespace dmcmc
