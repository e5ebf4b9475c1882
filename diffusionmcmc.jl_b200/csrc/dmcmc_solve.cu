// dmcmc_solve.cu -- the imputation hot path as hand-written CUDA for sm_100a.
//
// One thread owns one (chain, block) unit -- the parallel axis the reference itself names
// ("this loop should be entirely parallelizable", /root/reference src/run!_alterations.jl:15-16)
// times the chain batch axis.  One CTA = one layout block x blockDim.x consecutive chains, so the
// backward-filter table of the block is shared by the whole CTA.  Inside a unit the Euler
// recursion is serial (x_{k+1} needs x_k); everything that does not depend on x is moved off
// that chain:
//   * the accepted path is brought into shared memory by 16-byte asynchronous copies (LDGSTS), 16
//     consecutive lanes per (chain, 16-step tile) = 256 contiguous bytes, one tile ahead, no
//     registers involved; each lane then reads its own chain's states with conflict-free LDS.128;
//   * the block's table records are streamed by TMA bulk copies (UBLKCP + mbarrier) one tile
//     ahead, and for FitzHugh-Nagumo rewritten once per CTA into the per-step coefficients that
//     are the same for every chain (fhn_derive), so the per-chain step is ~30 FP64 instructions;
//   * Philox4x32-10 normals are generated in-register, four per call;
//   * the proposal is stored straight from registers, 32 bytes per two steps, into the chain's
//     proposal container (whole 128-byte lines are assembled in L2 before they reach HBM).
//
// Reference rows:
//   a2 pCN refresh            forward_guide!(..., rho, ...; Wnr)   src/run!_alterations.jl:18-21
//   a3 guiding term lookup    r = F[k] - H[k] x
//   a4 Euler-Maruyama solve   GP.solve_and_ll!                     src/run!_alterations.jl:206-208
//   a5 Girsanov ll            loglikhd / loglikhd_obs              src/workspaces.jl:429
//   a6 accept                 eMCMC.accept_reject!                 src/run!_alterations.jl:22-24
//   a7 register + swap        swap_paths!                          src/run!_alterations.jl:36-47, :81-91
//   a9 parameter-step solve   set_proposal!                        src/run!_alterations.jl:177-211
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "dmcmc_solve_common.cuh"

namespace dmcmc {


// One CTA = one layout block x blockDim.x consecutive chains; one thread = one (chain, block) unit.
//   MODE   : see SolveMode
//   PHILOX : in-register counter-based noise (production) vs explicit Z/E from HBM (parity)
//   REINV  : the accepted noise increment and the accepted ll are recovered on the fly from the
//            accepted path X (fused layout switch; no W traffic) instead of read from W
//   FAST   : model-specific fused step (FitzHugh-Nagumo with its built-in aux law)
// p.depth (2, 4 or 8) = tiles in flight of the table and input pipelines: 2 when the grid fills the
// machine (latency is hidden by the other warps), deeper when a few long units run alone.
template <class MD, bool PSI, int MODE, bool PHILOX, bool REINV, bool FAST>
__global__ void __launch_bounds__(128, MD::D == 2 ? DMCMC_MINB : 2) solve_kernel(const SolveParams p) {
    constexpr int D = MD::D, M = MD::M;
    using R = Rec<D, PSI>;
    constexpr int RS = R::SIZE;
    constexpr bool SOLVE = (MODE == MODE_IMPUTE || MODE == MODE_PARAM);
    constexpr bool NEED_XA = !SOLVE || REINV;   // walk the accepted path
    constexpr bool NEED_WA = SOLVE && !REINV;   // read the stored accepted noise
    using IS = InStage<D, M, NEED_XA>;          // exactly one of the two streams is staged
    // the proposal noise is kept exactly when it is not recovered from X later (store_w == !reinvert)
    constexpr bool STORE_W_MODE = ((MODE == MODE_IMPUTE && !REINV) || MODE == MODE_RELAYOUT);
    using ZT = typename std::conditional<PHILOX, float, double>::type;  // Philox normals carry float resolution
    static_assert(!FAST || (D == 2 && M == 1), "the fused step is FitzHugh-Nagumo's");

    extern __shared__ __align__(256) double dsm[];
    __shared__ alignas(8) uint64_t full[MAX_DEPTH];
    const int depth = p.depth, dmask = depth - 1, dlog = 31 - __clz(depth);  // depth is 2, 4 or 8
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *stab = dsm;                                                            // [depth][TS][RS]
    double *sinp = dsm + (size_t)depth * TS * RS + (size_t)warp * depth * IS::WARP;  // [depth][NPT][32][2]

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
        for (int i = 0; i < depth; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // everything below reads what the previous kernel in the stream wrote (selectors, paths, ll)
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---- table pipeline (issue side, thread 0): TMA bulk copy of one tile's records per stage ----
    IvMeta mc = load_meta(p, i1, chain, D), mn = mc;  // current interval of the solve, and the next one
    if (i1 < i2) mn = load_meta(p, i1 + 1, chain, D);
    int jc = i1;                                      // interval the solve is in
    int tj = i1, ttl = 0, tq = 0;
    auto issue_table = [&]() {
        if (tj > i2) return;
        const int Nj = tj == jc ? mc.N : (tj == jc + 1 ? mn.N : p.iv_N[tj]);
        const int toffj = tj == jc ? mc.tile_off : (tj == jc + 1 ? mn.tile_off : p.iv_tile_off[tj]);
        const int npad = ((Nj + TILE - 1) >> 3) * TILE;  // records the table holds for this interval
        const int nrec = min(TS, npad - ttl * TS);
        const uint32_t bytes = (uint32_t)(nrec * RS * sizeof(double));
        uint64_t *bar = &full[tq & dmask];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, bytes);
        bulk_g2s(stab + (size_t)(tq & dmask) * TS * RS, p.table + ((size_t)toffj * TILE + (size_t)ttl * TS) * RS, bytes, bar);
        ++tq;
        if (++ttl * TS >= Nj) { ttl = 0; ++tj; }
    };
    if (threadIdx.x == 0)
        for (int i = 0; i < depth - 1; ++i) issue_table();

    const bool pinned = PSI && p.blk_pinned[blk];
    double x[D], xb[D], xa[D];  // proposal state, pinned end point, accepted state (REINV / RELAYOUT)
    load_start<D>(p, chain, i1, x);
#pragma unroll
    for (int c = 0; c < D; ++c) { xb[c] = 0.0; xa[c] = x[c]; }
    if (pinned) load_end<D>(p, chain, i2, xb);

    FhnK fk;
    if constexpr (FAST) fk.set_law(p.th);

    double llp, lla;  // proposal ll, accepted-path ll (recomputed when REINV)
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
        lla = llp;
    }
    double lsum = 0.0, lsum_a = 0.0;
    bool ok = true;
    const uint32_t gchain = (uint32_t)(chain + p.chain_offset);
    PhiloxKeys pkeys;
    pkeys.set(p.seed);

    // ---- input pipeline (issue side): LDGSTS, runs depth-1 tiles ahead of the solve, across intervals ----
    // Lane l copies piece pc of chain r = q * RPI + l / NPT in its q-th request (RPI chains per request),
    // so the piece index, the live test and the swizzle are per-lane constants; src[q] walks the record.
    constexpr bool PO2 = (32 % IS::NPT == 0);          // d = 2 path, m = 1, 2 noise
    constexpr int RPI = PO2 ? 32 / IS::NPT : 1;
    const double *src[IS::NPT];  // per request: its chain's accepted record of the interval, this lane's 16 bytes
    int ij = i1, itl = 0, int_ = 0, ilim = 0, iq = 0;
    size_t iadv = 0;
    const int chain0 = chain_raw - lane;  // first chain of this warp
    const int pcl = lane % IS::NPT, rl = lane / IS::NPT;  // PO2: this lane's piece and first chain
    auto piece_of = [&](int q, int &r, int &pc) {
        if constexpr (PO2) { r = q * RPI + rl; pc = pcl; }
        else { const int g = q * 32 + lane; r = g / IS::NPT; pc = g % IS::NPT; }
    };
    auto setup_issue = [&](int j) {
        const IvMeta mi = j == jc ? mc : (j == jc + 1 ? mn : load_meta(p, j, chain, D));
        const int Nj = mi.N;
        int_ = (Nj + TS - 1) / TS;
        const unsigned selmask = __ballot_sync(0xffffffffu, NEED_XA ? mi.sx : mi.sw);
        if constexpr (NEED_XA) {
            const size_t xo = (size_t)mi.xoff;
            const int xs = mi.xs;
            ilim = Nj * D;  // live doubles of the chain's stream in this interval
            iadv = (size_t)TS * D;
#pragma unroll
            for (int q = 0; q < IS::NPT; ++q) {
                int r, pc;
                piece_of(q, r, pc);
                const int cr = min(chain0 + r, p.C - 1);
                src[q] = p.X[(selmask >> r) & 1] + xo + (size_t)cr * xs + pc * 2;
            }
        } else {
            const size_t t0 = (size_t)mi.tile_off;
            ilim = (Nj + TILE - 1) >> 3;  // W tiles of this interval
            iadv = (size_t)(TS / TILE) * p.C * M * TILE;
#pragma unroll
            for (int q = 0; q < IS::NPT; ++q) {
                int r, pc;
                piece_of(q, r, pc);
                const int cr = min(chain0 + r, p.C - 1);
                const int t8 = pc / (4 * M), rem = pc % (4 * M);
                src[q] = p.W[(selmask >> r) & 1] + ((t0 + t8) * p.C + cr) * (size_t)(M * TILE) + rem * 2;
            }
        }
    };
    // per-lane destination offset inside a stage (bytes).  d = 2 swizzle: request q copies piece pcl of chain
    // r = RPI q + rl into slot pcl ^ (r & 7) = (pcl ^ rl) ^ ((RPI q) & 7) of its record (rl < RPI, a power of two: no
    // carry), i.e. byte offset q * RPI * record + (dst0 ^ (((RPI q) & 7) << 4)): one register, XORs with immediates
    const uint32_t dst0 = IS::SWZ ? (uint32_t)(rl * IS::RECW * 8 + ((pcl ^ rl) << 4)) : (uint32_t)IS::piece(rl, pcl) * 8u;
    auto issue_in = [&]() {  // copy tile (ij, itl) into its stage, advance; always commits a group
        if (ij <= i2) {
            const uint32_t sbase = smem_u32(sinp + (size_t)(iq & dmask) * IS::WARP);
            if constexpr (PO2) {
                bool lv;
                if constexpr (NEED_XA) lv = itl * TS * D + pcl * 2 < ilim;
                else lv = itl * (TS / TILE) + pcl / (4 * M) < ilim;
                const uint32_t da = sbase + dst0;  // stage bases are 256-byte aligned: the XOR below never carries
#pragma unroll
                for (int q = 0; q < IS::NPT; ++q) {
                    const uint32_t dst = IS::SWZ ? (da ^ (uint32_t)(((q * RPI) & 7) << 4)) + (uint32_t)(q * RPI * IS::RECW * 8)
                                                 : da + (uint32_t)(q * RPI * IS::RECW * 8);
#ifndef DMCMC_EXP_NOLOAD
                    if (lv) cp_async16(dst, src[q]);
#endif
                    src[q] += iadv;
                }
            } else {
                const size_t toff = (size_t)itl * iadv;
#pragma unroll
                for (int q = 0; q < IS::NPT; ++q) {
                    int r, pc;
                    piece_of(q, r, pc);
                    bool lv;
                    if constexpr (NEED_XA) lv = itl * TS * D + pc * 2 < ilim;
                    else lv = itl * (TS / TILE) + pc / (4 * M) < ilim;
                    if (lv) cp_async16(sbase + (uint32_t)IS::piece(r, pc) * 8u, src[q] + toff);
                }
            }
            if (__builtin_expect(++itl == int_, 0)) {  // next interval (rare: keep it a branch, not predicated code)
                itl = 0;
                if (++ij <= i2) setup_issue(ij);
            }
        }
        ++iq;
        cp_async_commit();
    };
    setup_issue(i1);
    for (int i = 0; i < depth - 1; ++i) issue_in();

    int q = 0;  // tiles consumed so far
    uint32_t sxbits = 0, swbits = 0;  // selectors of the block's first 32 intervals (re-used by the swap below)
    double *xp_end = nullptr;         // last point of the block in the proposal container
    for (int j = i1; j <= i2; ++j) {
        if (j > i1) {  // the metadata of this interval arrived while the previous one was solved
            mc = mn;
            jc = j;
            if (j < i2) mn = load_meta(p, j + 1, chain, D);  // in flight for the whole interval
        }
        const int N = mc.N;
        const int nt = (N + TS - 1) / TS;
        const int sx = mc.sx, sw = mc.sw;
        if (j - i1 < 32) { sxbits |= (uint32_t)sx << (j - i1); swbits |= (uint32_t)sw << (j - i1); }
        double aB[FAST ? 1 : D * D], abeta[FAST ? 1 : D], aatil[MD::CONST_SIGMA ? 1 : D * D];
        if constexpr (FAST) {
            fk.aB0 = mc.aB0;
            fk.c2 = mc.ab0 - fk.s * fk.ie;
        } else {
#pragma unroll
            for (int i = 0; i < D * D; ++i) aB[i] = p.auxB[(size_t)j * D * D + i];
#pragma unroll
            for (int i = 0; i < D; ++i) abeta[i] = p.auxbeta[(size_t)j * D + i];
        }
        if constexpr (!MD::CONST_SIGMA) {
#pragma unroll
            for (int i = 0; i < D * D; ++i) aatil[i] = p.auxatil[(size_t)j * D * D + i];
        }
        const double rho = (MODE == MODE_IMPUTE) ? mc.rho : 1.0;
        const double lam = (MODE == MODE_IMPUTE) ? sqrt(1.0 - rho * rho) : 0.0;
        double *Wdst = (MODE == MODE_RELAYOUT) ? p.W[sw] : p.W[1 - sw];
        double *Xp = p.X[1 - sx] + mc.xoff + (size_t)chain * mc.xs;  // this chain's proposal record
        const bool store_w = active && STORE_W_MODE;
        xp_end = Xp + (size_t)(N - 1) * D;
        const size_t wtile0 = (size_t)mc.tile_off;

        for (int tl = 0; tl < nt; ++tl, ++q) {
            // ---- table records of this tile: wait, refill the stage freed by the previous tile ----
            __syncthreads();  // every thread is done with tile q-1 (a full/empty mbarrier handshake was measured: no faster)
            if (threadIdx.x == 0) issue_table();
            mbar_wait(&full[q & dmask], (uint32_t)((q >> dlog) & 1));
            const double *srec = stab + (size_t)(q & dmask) * TS * RS;
            // ---- this tile's input has landed; refill the stage freed by the previous tile ----
            cp_async_wait_depth(depth);
            __syncwarp();  // all lanes' pieces are visible, and all lanes are done with tile q-1's stage
            issue_in();
            const double *sst = sinp + (size_t)(q & dmask) * IS::WARP;

            // eight Euler steps; CHECK = the half-tile may run past the interval's last step
            auto eight = [&](auto check_tag, int half) {
                constexpr bool CHECK = decltype(check_tag)::value;
                const int kh = tl * TS + half * TILE;
                [[maybe_unused]] const size_t wtile = wtile0 + (size_t)(kh >> 3);
                const double *hrec = srec + (size_t)half * TILE * RS;
                const uint32_t hst = IS::cursor(sst, lane);
                [[maybe_unused]] double *xph = Xp + (size_t)kh * D;
                // The fused FitzHugh-Nagumo step keeps all eight steps in one basic block.  The generic step is ~450
                // instructions: eight copies (times two CHECK variants) made the loop 126 KB of code against a 32 KB
                // instruction cache (ncu on Lorenz-63: no_instruction 0.5 warps per issue), so it is unrolled by FOUR
                // steps -- one Philox call per Wiener component -- inside a rolled loop of two.
                constexpr int QN = FAST ? 1 : 2, SL = TILE / QN;
#pragma unroll 1
              for (int hq = 0; hq < QN; ++hq) {
                const int s0 = hq * SL;
                [[maybe_unused]] double dwo[STORE_W_MODE ? M : 1][SL];
                [[maybe_unused]] ZT z[(MODE == MODE_IMPUTE) ? M : 1][SL];
                if constexpr (MODE == MODE_IMPUTE) {
                    if constexpr (PHILOX) {
#pragma unroll
                        for (int w = 0; w < M; ++w)
#pragma unroll
                            for (int h = 0; h < SL / 4; ++h)
                                philox_normal4f(pkeys, p.iter, (uint32_t)(j + p.interval_offset), gchain, (uint32_t)w,
                                                (uint32_t)(((kh + s0) >> 2) + h), (float *)&z[w][4 * h]);
                    } else {
#pragma unroll
                        for (int w = 0; w < M; ++w) {
                            const double *zsrc = p.Z + w_off(wtile, chain, w, p.C, M) + s0;
                            if constexpr (SL == TILE) ld_tile_stream(zsrc, (double *)z[w]);
                            else asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(z[w][0]), "=d"(z[w][1]), "=d"(z[w][2]), "=d"(z[w][3]) : "l"(zsrc));
                        }
                    }
                }
                double xprev[D];  // proposal state after the previous (even) step: stored together with the odd one
#pragma unroll
                for (int c = 0; c < D; ++c) xprev[c] = 0.0;
#pragma unroll
                for (int si = 0; si < SL; ++si) {
                    const int s = s0 + si;
                    const int ks = half * TILE + s;  // step within the tile
                    const bool live = !CHECK || (kh + s < N);  // CTA-uniform
                    const double *rp = hrec + s * RS;
                    double dw[M];
#pragma unroll
                    for (int w = 0; w < M; ++w) dw[w] = 0.0;
                    if (live) {
                        if constexpr (FAST) {
                            double rec[RS];
                            load_rec_smem<RS>(rp, rec);
                            double Cf = rec[FD_C1], R0c = rec[FD_R0];
                            if constexpr (PSI) {
                                Cf = fma(rec[FD_Q10], xb[0], fma(rec[FD_Q11], xb[1], Cf));
                                R0c = fma(rec[FD_T00], xb[0], fma(rec[FD_T01], xb[1], R0c));
                            }
                            if constexpr (NEED_XA) {  // recover the accepted noise increment from the accepted path
                                double xn[2], g;
                                IS::state(hst, ks, xn);
                                fhn_dll(fk, rec, R0c, xa[0], xa[1], g, lsum_a);
                                const double pred = fma(rec[FD_A1], xa[1], fma(rec[FD_B1], xa[0], Cf));
                                dw[0] = (xn[1] - pred) * fk.isig;
                                xa[0] = xn[0]; xa[1] = xn[1];
                            } else if constexpr (NEED_WA) {
                                IS::noise(hst, ks, dw);
                            }
                            if constexpr (MODE == MODE_IMPUTE) dw[0] = fma(lam * rec[FD_SQ], (double)z[0][si], rho * dw[0]);  // pCN
                            if constexpr (SOLVE) {
                                fhn_step(fk, rec, R0c, fma(fk.sig, dw[0], Cf), x[0], x[1], lsum);
                            }
                        } else {
                            double rec[RS], F[D];
                            load_rec_smem<RS>(rp, rec);
                            guiding_F<D, PSI>(rec, xb, F);
                            const double dt = rec[0];
                            if constexpr (NEED_XA) {
                                double xn[D], dr[D], res[D], gg;
                                IS::state(hst, ks, xn);
                                guided_terms<MD, PSI>(p.th, rec, F, aB, abeta, aatil, xa, dt, dr, gg);
                                lsum_a = fma(gg, dt, lsum_a);
#pragma unroll
                                for (int c = 0; c < D; ++c) res[c] = xn[c] - xa[c] - dr[c];
                                MD::sigma_solve(p.th, xa, res, dw);
#pragma unroll
                                for (int c = 0; c < D; ++c) xa[c] = xn[c];
                            } else if constexpr (NEED_WA) {
                                IS::noise(hst, ks, dw);
                            }
                            if constexpr (MODE == MODE_IMPUTE) {  // pCN: W deg = sqrt(1-rho^2) W_fresh + rho W
#pragma unroll
                                for (int w = 0; w < M; ++w) dw[w] = lam * (rec[1] * (double)z[w][si]) + rho * dw[w];
                            }
                            if constexpr (SOLVE) {
                                double dr[D], sdw[D], gg;
                                guided_terms<MD, PSI>(p.th, rec, F, aB, abeta, aatil, x, dt, dr, gg);
                                lsum = fma(gg, dt, lsum);
                                MD::sigma_mul(p.th, x, dw, sdw);
#pragma unroll
                                for (int c = 0; c < D; ++c) x[c] = x[c] + dr[c] + sdw[c];
                                ok = ok && MD::bound_ok(x);
                            }
                        }
                    }
                    if constexpr (STORE_W_MODE) {
#pragma unroll
                        for (int w = 0; w < M; ++w) dwo[w][si] = dw[w];
                    }
                    if constexpr (SOLVE) {
                        if (si & 1) {  // two steps per store: 16-byte aligned (32 for d = 2)
                            if (live && active) {
                                double *dst = xph + (s - 1) * D;
                                if constexpr (D == 2) {
#ifndef DMCMC_EXP_NOSTORE
                                    st_v4(dst, xprev[0], xprev[1], x[0], x[1]);
#else
                                    if (x[0] == 1.2345e300) st_v4(dst, xprev[0], xprev[1], x[0], x[1]);
#endif
                                } else {
                                    double pr[2 * D];
#pragma unroll
                                    for (int c = 0; c < D; ++c) { pr[c] = xprev[c]; pr[D + c] = x[c]; }
#pragma unroll
                                    for (int c = 0; c < D; ++c) st_v2(dst + 2 * c, pr[2 * c], pr[2 * c + 1]);
                                }
                            }
                        } else {
                            if (CHECK && live && kh + s == N - 1 && active) {  // odd N: the last state goes out alone
#pragma unroll
                                for (int c = 0; c < D; ++c) xph[s * D + c] = x[c];
                            }
#pragma unroll
                            for (int c = 0; c < D; ++c) xprev[c] = x[c];
                        }
                    }
                }
                if constexpr (STORE_W_MODE) {
                    if (store_w) {
#pragma unroll
                        for (int w = 0; w < M; ++w) {
                            double *wd = Wdst + w_off(wtile, chain, w, p.C, M) + s0;
                            if constexpr (SL == TILE) st_tile(wd, dwo[w]);
                            else st_v4(wd, dwo[w][0], dwo[w][1], dwo[w][2], dwo[w][3]);
                        }
                    }
                }
              }
            };
#pragma unroll 1
            for (int half = 0; half < TS / TILE; ++half) {
                const int kh = tl * TS + half * TILE;
                // generic models: one copy of the eight-step body (the live test is CTA-uniform); two copies of the ~450
                // instructions per step put the loop at 126 KB, far beyond the instruction cache (ncu: no_instruction 0.5
                // warps per issue on Lorenz-63)
                if (FAST && kh + TILE <= N) eight(std::false_type{}, half);
                else if (kh < N) eight(std::true_type{}, half);
            }
        }
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    llp += lsum;
    lla += lsum_a;
    if (!ok) llp = -INFINITY;
    cp_async_wait<0>();  // nothing of ours may still be in flight when the CTA retires
    if constexpr (SOLVE) {
        if (pinned && active) {  // pin the proposal's last point to x_b (glued paths stay continuous)
#pragma unroll
            for (int c = 0; c < D; ++c) xp_end[c] = xb[c];
        }
    }

    if constexpr (MODE == MODE_IMPUTE) {
        double llcur = REINV ? lla : p.ll[uo];
        const double e = PHILOX ? philox_exponential(p.seed, p.iter, (uint32_t)(i1 + p.interval_offset), gchain) : p.E[uo];
        // accepted = rand(Exponential(1)) > -(ll deg - ll)   (eMCMC.accept_reject!)
        const bool acc = active && (p.force_accept ? ok : (e > -(llp - llcur)));
        if (in_range) {
            for (int j = i1; j <= i2; ++j) {  // swap_paths!: flip the accepted-container selectors
                const size_t o = (size_t)j * p.C + chain;
                const int b = j - i1;
                const uint8_t sxj = b < 32 ? (uint8_t)((sxbits >> b) & 1u) : p.selX_in[o];
                const uint8_t swj = b < 32 ? (uint8_t)((swbits >> b) & 1u) : p.selW_in[o];
                p.selX_out[o] = sxj ^ (uint8_t)acc;
                p.selW_out[o] = swj ^ (uint8_t)(acc && STORE_W_MODE);
            }
        }
        // acceptance counters of the block: one atomic per warp
        const unsigned accm = __ballot_sync(0xffffffffu, acc), actm = __ballot_sync(0xffffffffu, active);
        if (lane == 0) {
            if (p.acc_count && accm) atomicAdd(p.acc_count + blk, (unsigned long long)__popc(accm));
            if (p.prop_count && actm) atomicAdd(p.prop_count + blk, (unsigned long long)__popc(actm));
        }
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
    } else if constexpr (MODE == MODE_PARAM) {
        if (in_range) {
            if (p.out_llprop) p.out_llprop[uo] = llp;
            if (p.out_acc) p.out_acc[uo] = (uint8_t)ok;
        }
    } else {
        if (in_range) p.ll[uo] = lla;
    }
}

// ---- FitzHugh-Nagumo: rewrite a (layout, law) table into the fused step's coefficients -----------
template <bool PSI>
__global__ void fhn_derive_kernel(const double *__restrict__ raw, double *__restrict__ out, long long nrec, Theta th) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrec) return;
    constexpr int RS = Rec<2, PSI>::SIZE;
    FhnK fk;
    fk.set_law(th);
    double r[RS], o[RS];
#pragma unroll
    for (int k = 0; k < RS; ++k) { r[k] = raw[i * RS + k]; o[k] = 0.0; }
    fhn_derive<PSI>(r, o, fk);
#pragma unroll
    for (int k = 0; k < RS; ++k) out[i * RS + k] = o[k];
}

cudaError_t launch_fhn_derive(const double *raw, double *out, long long nrec, int has_psi, const Theta &th, cudaStream_t s) {
    if (nrec <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((nrec + 127) / 128);
    if (has_psi) fhn_derive_kernel<true><<<grid, 128, 0, s>>>(raw, out, nrec, th);
    else fhn_derive_kernel<false><<<grid, 128, 0, s>>>(raw, out, nrec, th);
    return cudaGetLastError();
}

// ---- launch dispatch ---------------------------------------------------------------------------
template <class MD, bool PSI, int MODE, bool PHILOX, bool REINV, bool FAST>
static cudaError_t launch_one(const SolveParams &p0, int threads, cudaStream_t s) {
    constexpr bool SOLVE = (MODE == MODE_IMPUTE || MODE == MODE_PARAM);
    constexpr bool NEED_XA = !SOLVE || REINV;
    using IS = InStage<MD::D, MD::M, NEED_XA>;
    SolveParams p = p0;
    // few long units (e.g. one block per chain): one warp per CTA so that more SMs take part
    const long long units = (long long)p.C * p.nblk;
    // CTA size when the caller did not fix it (dmcmc_set_threads): four warps share a table stage when the grid still
    // fills the machine (C3: 76.7 us per sweep against 77.7 us with two warps), two warps otherwise
    if (threads <= 0) threads = (MD::D == 2 && units >= 148LL * 128 * 2) ? 128 : 64;
    if (units < 148LL * 64 * 2) threads = 32;
    const unsigned groups = (unsigned)((p.C + threads - 1) / threads);
    const unsigned grid = groups * (unsigned)p.nblk;
    auto smem_for = [&](int depth) {
        return sizeof(double) * (size_t)depth * ((size_t)TS * Rec<MD::D, PSI>::SIZE + (size_t)(threads / 32) * IS::WARP);
    };
    // pipeline depth: 2 when the grid fills the machine; deeper when few CTAs run alone per SM
    const int cta_per_sm = (int)std::min<long long>(6, (grid + 147) / 148);
    int depth = 2;
    if (cta_per_sm <= 4)
        for (int d = MAX_DEPTH; d > 2; d >>= 1)
            if ((smem_for(d) + 1024) * cta_per_sm <= 200 * 1024) { depth = d; break; }
    p.depth = depth;
    auto kern = solve_kernel<MD, PSI, MODE, PHILOX, REINV, FAST>;
    static bool configured = false;  // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    // programmatic dependent launch: the next sweep's CTAs may be scheduled (and set up their shared
    // memory) while this grid drains; they block in griddepcontrol.wait before touching global memory
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem_for(depth);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

template <class MD, bool PSI, bool FAST>
static cudaError_t launch_model(int mode, const SolveParams &p, int threads, cudaStream_t s) {
    if ((long long)p.nblk * p.C == 0) return cudaSuccess;
    if (threads > 0) threads = ((threads + 31) / 32) * 32;
    const bool philox = (p.Z == nullptr), reinv = p.reinvert != 0;
    switch (mode) {
    case MODE_IMPUTE:
        if (philox) return reinv ? launch_one<MD, PSI, MODE_IMPUTE, true, true, FAST>(p, threads, s)
                                 : launch_one<MD, PSI, MODE_IMPUTE, true, false, FAST>(p, threads, s);
        return reinv ? launch_one<MD, PSI, MODE_IMPUTE, false, true, FAST>(p, threads, s)
                     : launch_one<MD, PSI, MODE_IMPUTE, false, false, FAST>(p, threads, s);
    case MODE_PARAM: return launch_one<MD, PSI, MODE_PARAM, false, false, FAST>(p, threads, s);
    case MODE_RELAYOUT: return launch_one<MD, PSI, MODE_RELAYOUT, false, false, FAST>(p, threads, s);
    case MODE_LOGLIK: return launch_one<MD, PSI, MODE_LOGLIK, false, false, FAST>(p, threads, s);
    }
    return cudaErrorInvalidValue;
}

template <class MD, bool FAST = false>
static cudaError_t launch_psi(int mode, const SolveParams &p, int threads, cudaStream_t s) {
    return p.has_psi ? launch_model<MD, true, FAST>(mode, p, threads, s)
                     : launch_model<MD, false, FAST>(mode, p, threads, s);
}

// launches with at most this many (chain, block) units go to the warp-per-unit kernel: below it the
// thread-per-unit kernel leaves most of the machine idle and one lane's instruction stream is the
// critical path; above it the cooperative kernel's 32 x redundant recursion costs more than it hides
constexpr long long COOP_MAX_UNITS = 2048;  // measured crossover on B200: ~3000 units (FitzHugh-Nagumo fused step), ~4000 (Lorenz-63)

bool solve_uses_warp_per_unit(int model, const SolveParams &p) {
    const long long units = (long long)p.C * p.nblk;
    return model != 3 && (p.coop > 0 || (p.coop < 0 && units <= COOP_MAX_UNITS));
}

cudaError_t launch_solve(int model, int mode, const SolveParams &p, int threads, cudaStream_t s) {
    if (solve_uses_warp_per_unit(model, p)) return launch_solve_coop(model, mode, p, s);
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
            v = g.X[sp][g.iv_xoff[jp] + (size_t)chain * g.iv_xs[jp] + (size_t)(g.iv_N[jp] - 1) * g.D + c];
        }
        g.outX[(pt0)*g.D + c] = v;
    }
    for (int w = 0; w < g.M; ++w) g.outW[pt0 * g.M + w] = 0.0;
    const double *X = g.X[sx] + g.iv_xoff[j] + (size_t)chain * g.iv_xs[j];
    for (int k = 0; k < N; ++k) {
        const size_t tile = (size_t)g.iv_tile_off[j] + (k >> 3);
        for (int c = 0; c < g.D; ++c) g.outX[(pt0 + k + 1) * g.D + c] = X[(size_t)k * g.D + c];
        for (int w = 0; w < g.M; ++w)
            g.outW[(pt0 + k + 1) * g.M + w] = g.outW[(pt0 + k) * g.M + w] + g.W[sw][w_off(tile, chain, w, g.C, g.M) + (k & 7)];
    }
}

__global__ void segment_kernel(const SegmentParams g) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nj = g.j1 - g.j0, nc = g.nc > 0 ? g.nc : g.C;
    if (u >= (long long)nj * nc) return;
    const int j = g.j0 + (int)(u / nc), ci = (int)(u % nc), chain = g.c0 + ci;
    double *X = g.X[g.selX[(size_t)j * g.C + chain]] + g.iv_xoff[j] + (size_t)chain * g.iv_xs[j];
    const int off = g.iv_st_off[j] - g.iv_st_off[g.j0];
    for (int k = 0; k < g.iv_N[j]; ++k)
        for (int c = 0; c < g.D; ++c) {
            double *pk = g.packed + (size_t)ci * g.pk_stride + (size_t)(off + k) * g.D + c;
            if (g.upload) X[(size_t)k * g.D + c] = *pk;
            else *pk = X[(size_t)k * g.D + c];
        }
}

// x0[(rec * D + comp) * C + chain] = src[chain * chain_stride + rec * D + comp]
__global__ void set_start_kernel(double *x0, const double *src, long long chain_stride, int C, int R, int D) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (long long)C * R * D) return;
    const int chain = (int)(u % C), rc = (int)(u / C);
    x0[(size_t)rc * C + chain] = src[(size_t)chain * chain_stride + rc];
}

cudaError_t launch_set_start(double *x0, const double *src, int64_t chain_stride, int C, int R, int D, cudaStream_t s) {
    const long long n = (long long)C * R * D;
    if (n <= 0) return cudaSuccess;
    set_start_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(x0, src, (long long)chain_stride, C, R, D);
    return cudaGetLastError();
}

cudaError_t launch_segment(const SegmentParams &g, cudaStream_t s) {
    const long long n = (long long)(g.j1 - g.j0) * (g.nc > 0 ? g.nc : g.C);
    if (n <= 0) return cudaSuccess;
    segment_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g);
    return cudaGetLastError();
}

// ---- halo exchange over peer memory --------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// every CTA: thread 0 spins on the flag (three seconds at most: a dead peer must not hang the GPU), then the CTA goes on
__device__ __forceinline__ void p2p_wait(const P2PSync &y) {
    if (y.wait_flag && threadIdx.x == 0) {
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(y.wait_flag) < y.wait_val) {
            if (global_ns() - t0 > 3000000000ULL) { *y.err = 1; break; }
        }
    }
    __syncthreads();
}
// the last CTA (of the nb CTAs in this role) to finish raises the flag in the peer's memory, after every CTA's writes.
// ONE st.release.sys by one thread: the release is cumulative over what the CTA barrier (and, across CTAs, the gpu-scope
// fence around the counter) has ordered before it.  Measured between two B200s (tools/micro/p2p_latency.cu): 4.2 us per
// hand-over of 150 doubles this way, 5.8 us with an extra __threadfence_system() by the signalling thread, 9.2 us with
// one by every thread (the first version here).
__device__ __forceinline__ void p2p_signal(const P2PSync &y, unsigned nb) {
    if (!y.signal_flag) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        bool last = true;
        if (nb > 1) {
            __threadfence();
            last = atomicAdd(y.counter, 1u) == nb - 1;
            if (last) { *y.counter = 0; __threadfence(); }
        }
        if (last) st_release_sys(y.signal_flag, y.signal_val);
    }
}
// element u of the packed message -> its address in the accepted container and (pk) in the message
__device__ __forceinline__ double *halo_element(const SegmentParams &g, long long u, size_t &pk) {
    const long long per = (long long)g.steps * g.D;
    const int ci = (int)(u / per);
    const int rem = (int)(u - (long long)ci * per), step = rem / g.D, c = rem - step * g.D;
    int j = g.j0;
    const int base = g.iv_st_off[g.j0];
    while (j + 1 < g.j1 && g.iv_st_off[j + 1] - base <= step) ++j;
    const int k = step - (g.iv_st_off[j] - base), chain = g.c0 + ci;
    pk = (size_t)ci * g.pk_stride + (size_t)step * g.D + c;
    return g.X[g.selX[(size_t)j * g.C + chain]] + g.iv_xoff[j] + (size_t)chain * g.iv_xs[j] + (size_t)k * g.D + c;
}

// One launch per rank and exchange.  CTAs [0, nb_push) SEND: wait until the neighbour has taken the previous message out
// of its mailbox, store the accepted X of intervals [gs.j0, gs.j1) straight into that mailbox (peer memory: the stores go
// out over NVLink), raise its data flag.  The other CTAs RECEIVE: spin on the own data flag, unpack the own mailbox into
// intervals [gr.j0, gr.j1) (and the new starting point), acknowledge in the sender's memory.  Launched with programmatic
// stream serialisation like the sweeps around it, so its launch latency hides behind the sweep's tail.
__global__ void __launch_bounds__(128) halo_xchg_kernel(const SegmentParams gs, const P2PSync ys, const unsigned nb_push,
                                                        const SegmentParams gr, double *x0, const double *start_src,
                                                        const long long start_stride, const int R, const P2PSync yr) {
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the sweep before this has written the paths and selectors
    if (blockIdx.x < nb_push) {
        p2p_wait(ys);
        const long long n = (long long)(gs.nc > 0 ? gs.nc : gs.C) * gs.steps * gs.D;
        for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += (long long)nb_push * blockDim.x) {
            size_t pk;
            const double *x = halo_element(gs, u, pk);
            gs.packed[pk] = *x;
        }
        p2p_signal(ys, nb_push);
    } else {
        const unsigned nb = gridDim.x - nb_push, bid = blockIdx.x - nb_push;
        p2p_wait(yr);
        const long long n = (long long)(gr.nc > 0 ? gr.nc : gr.C) * gr.steps * gr.D;
        const long long ns = start_src ? (long long)gr.C * R * gr.D : 0;
        for (long long u = (long long)bid * blockDim.x + threadIdx.x; u < n + ns; u += (long long)nb * blockDim.x) {
            if (u < n) {
                size_t pk;
                double *x = halo_element(gr, u, pk);
                *x = __ldcg(gr.packed + pk);
            } else {
                const long long v = u - n;
                const int chain = (int)(v % gr.C), rc = (int)(v / gr.C);
                x0[(size_t)rc * gr.C + chain] = __ldcg(start_src + (size_t)chain * start_stride + rc);
            }
        }
        p2p_signal(yr, nb);  // acknowledge: the mailbox may be overwritten
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

cudaError_t launch_halo_exchange(const SegmentParams *push, const P2PSync &ys, const SegmentParams *pull, double *x0,
                                 const double *start_src, int64_t start_stride, int R, const P2PSync &yr, cudaStream_t s) {
    auto ctas = [](long long n) { return (unsigned)std::min<long long>(std::max<long long>((n + 127) / 128, 1), 148 * 2); };
    SegmentParams gs{}, gr{};
    unsigned nb_push = 0, nb_pull = 0;
    if (push) { gs = *push; nb_push = ctas((long long)(gs.nc > 0 ? gs.nc : gs.C) * gs.steps * gs.D); }
    if (pull) { gr = *pull; nb_pull = ctas((long long)(gr.nc > 0 ? gr.nc : gr.C) * gr.steps * gr.D + (start_src ? (long long)gr.C * R * gr.D : 0)); }
    if (nb_push + nb_pull == 0) return cudaSuccess;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nb_push + nb_pull);
    cfg.blockDim = dim3(128);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, halo_xchg_kernel, gs, ys, nb_push, gr, x0, start_src, (long long)start_stride, R, yr);
}

cudaError_t launch_gather(const GatherParams &g, cudaStream_t s) {
    const long long n = (long long)g.J * g.C;
    if (n == 0) return cudaSuccess;
    gather_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g);
    return cudaGetLastError();
}

}  // namespace dmcmc
