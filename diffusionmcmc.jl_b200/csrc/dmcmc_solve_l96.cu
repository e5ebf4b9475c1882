// dmcmc_solve_l96.cu -- the imputation hot path for large state dimension (Lorenz-96, d = 40) on the FP64 tensor cores.
//
// Per Euler step a chain needs three dense 40 x 40 mat-vecs with matrices every chain of a layout block shares
// (H x, H x_acc when the accepted noise is recovered on the fly, Psi' x_b): batched over chains they are GEMMs,
// [8 chains x 40] x [40 x 40], and run as mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; measured 36.8 TFLOP/s on B200 against
// 34.2 for plain DFMA, tools/micro/dmma.cu) -- one shared-memory load of a matrix fragment now feeds 512 FMAs instead
// of 15 (the DFMA mapping of round 1/2 was shared-memory-wavefront bound: LSU 43 %, FP64 30 %).
//
// One WARP owns eight (chain, block) units of the same layout block.  Lane l works for chain l / 4 and owns the ten
// state components 8 nt + 2 (l % 4) + {0, 1}, nt = 0..4: exactly the elements of the m8n8k4 accumulator fragment it
// holds when the product is formed as  Y' = X' T  (rows = chains, T[kk][c] the table's matrix).  The summation index is
// free to be permuted, and with kk = 8 kp + 2 (l % 4) + e for k-step (kp, e) the A operand a lane must supply is its OWN
// state value x[kp][e]: the state never leaves its registers between steps.  The B operands (matrix fragments) are
// stored in the table record in fragment order (big_frag, dmcmc_device.cuh), so a lane fetches the two k-steps of a
// (kp, nt) tile with one conflict-free 16-byte shared-memory load.
//
// The CTA's warps (all on the same block) share the step's table record (H, F0, Psi': 26 KB), brought in by TMA bulk
// copies into a ring of stages guarded by full / empty mbarriers; no CTA barrier in the loop.  The accepted path
// arrives by cp.async two states ahead in a warp-private ring (neighbour components of the Lorenz-96 drift are read
// from there), the proposal's neighbours go through a warp-private double buffer.  Path I/O is per step, 16 bytes per
// lane, 320 contiguous bytes per chain.
//
// Same rows of SURVEY.md section 8a as dmcmc_solve.cu (a1-a7, a9, layout switch, init_ll).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "dmcmc_device.cuh"
#include "dmcmc_internal.h"

namespace dmcmc {

namespace {

__device__ __forceinline__ uint32_t sm_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sm_u32(dst)), "l"(src), "r"(bytes), "r"(sm_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sm_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAITL_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra WAITL_DONE;\n bra WAITL_LOOP;\n WAITL_DONE:\n}" ::"r"(sm_u32(b)), "r"(parity) : "memory");
}
// D (8 chains x 8 components) += A (8 chains x 4) * B (4 x 8 components), FP64 tensor core
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ double quad_sum(double v) {  // over the four lanes of one chain
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

constexpr int D = 40;
constexpr int NT = 5;         // tiles of eight components
constexpr int NST = 2;        // table stages
constexpr int AST = 3;        // accepted-state stages (A_t neighbours, A_{t+1}, A_{t+2} in flight)
constexpr int ROW = 8 * D;    // doubles per (stage, warp): eight chains' states
constexpr int WARP_DOUBLES = AST * ROW + 2 * ROW + (4 * 10 * 32) / 2;  // + proposal double buffer + four steps of normals (float)

// auxiliary drift of a custom law (dmcmc_set_aux), row c: kept out of line so that the unrolled step body of the built-in
// law stays inside the instruction cache (the loop body was 80 KB with this inlined 20 times)
__device__ __noinline__ double aux_drift_general(const double *aB, const double *abeta, int c, const double *xrow) {
    double a = 0.0;
#pragma unroll 1
    for (int kk = 0; kk < 40; ++kk) a += aB[(size_t)c * 40 + kk] * xrow[kk];
    return a + abeta[c];
}

struct Cursor {  // (interval, step) walking a layout block
    int j, k, N;
};

}  // namespace

size_t l96_smem_bytes(bool psi, int warps) {
    return (size_t)(NST * rec_size_big(D, psi) + warps * WARP_DOUBLES) * sizeof(double);
}

// FAST: the law uses the model's built-in auxiliary law (a compile-time switch: a run-time branch per component cut the
// unrolled element-wise part into small basic blocks and dragged generic-address arithmetic for the call into the loop)
template <bool PSI, int MODE, bool PHILOX, bool REINV, bool FAST>
__global__ void __launch_bounds__(256, 1) solve_l96_kernel(const SolveParams p) {
    constexpr int REC = rec_size_big(D, PSI);
    constexpr int OFF_H = 2, OFF_F = 2 + D * D, OFF_PSI = 2 + D * D + D;
    constexpr bool SOLVE = (MODE == MODE_IMPUTE || MODE == MODE_PARAM);
    constexpr bool NEED_XA = !SOLVE || REINV;
    constexpr bool NEED_WA = SOLVE && !REINV;
    constexpr bool STORE_W_MODE = ((MODE == MODE_IMPUTE && !REINV) || MODE == MODE_RELAYOUT);
    constexpr bool ZSM = (MODE == MODE_IMPUTE) && PHILOX;
    extern __shared__ __align__(128) double dsm[];
    __shared__ __align__(8) uint64_t full[NST], empty[NST];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int c8 = lane >> 2, q = lane & 3;  // chain within the warp, position in the chain's quad
    double *wbase = dsm + NST * REC + (size_t)warp * WARP_DOUBLES;
    double *ast = wbase + c8 * D;                 // accepted states: [stage][chain][component], this lane's chain row
    double *xbuf = wbase + AST * ROW + c8 * D;    // proposal state: [2][chain][component]
    float *zsm = reinterpret_cast<float *>(wbase + AST * ROW + 2 * ROW) + lane;  // [step][own component][lane]

    const int cpc = nwarps * 8;  // chains per CTA
    const int groups = (p.C + cpc - 1) / cpc;
    const int blk = (int)(blockIdx.x / groups);
    const int i1 = p.blk_i1[blk], i2 = p.blk_i2[blk];
    const int chain_raw = (int)(blockIdx.x % groups) * cpc + warp * 8 + c8;
    const bool in_range = chain_raw < p.C;
    const int chain = in_range ? chain_raw : p.C - 1;  // surplus lanes shadow the last chain (stores off)
    const size_t uo = (size_t)chain * p.nblk + blk;
    const bool masked = (MODE == MODE_IMPUTE) && ((p.unit_mask && !p.unit_mask[uo]) || (p.blk_active && !p.blk_active[blk]));
    const bool active = in_range && !masked;
    const uint32_t gchain = (uint32_t)(chain + p.chain_offset);

    const double Fth = p.th.v[0], sig = p.th.v[1], caux = p.th.v[2];
    const double sig2 = sig * sig, isig = 1.0 / sig;
    const bool pinned = PSI && p.blk_pinned[blk];
    PhiloxKeys pkeys;
    pkeys.set(p.seed);

    // neighbour offsets of the Lorenz-96 drift inside a chain row: pair (i-2, i-1) left of the own pair (i, i+1), and
    // i+2 right of it; only tile 0 of quad lane 0 and tile 4 of quad lane 3 wrap around
    const int relL0 = q ? -2 : D - 2;       // relative to the own pair of tile 0 (component 2 q)
    const int relR4 = (q == 3) ? -6 : 34;   // relative to component 2 q: component 34 + 2 q, or 0 for quad lane 3

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (uint32_t)nwarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto advance = [&](Cursor &c) {
        if (++c.k == c.N) { ++c.j; c.k = 0; c.N = (c.j <= i2) ? p.iv_N[c.j] : 0; }
    };
    auto issue_table = [&](const Cursor &c, int s) {  // one thread: stage the record of one Euler step
        mbar_expect_tx(&full[s], (uint32_t)(REC * sizeof(double)));
        bulk_g2s(dsm + s * REC, p.table + ((size_t)p.iv_tile_off[c.j] * TILE + c.k) * REC, (uint32_t)(REC * sizeof(double)), &full[s]);
    };
    Cursor tc{i1, 0, p.iv_N[i1]};  // table cursor: the step NST ahead of the one being computed
#pragma unroll
    for (int s = 0; s < NST; ++s) {
        if (tc.j <= i2) {
            if (threadIdx.x == 0) issue_table(tc, s);
            advance(tc);
        }
    }

    // ---- start point and pinned end point -----------------------------------------------------
    auto end_value2 = [&](int j, int c, double &v0, double &v1) {
        const int s = p.selX_in[(size_t)j * p.C + chain];
        const double *src = p.X[s] + p.iv_xoff[j] + (size_t)chain * p.iv_xs[j] + (size_t)(p.iv_N[j] - 1) * D + c;
        v0 = src[0]; v1 = src[1];
    };
    double x[NT][2], xa[NT][2], xb[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int c = 8 * nt + 2 * q;
        if (p.iv_first[i1]) {
            x[nt][0] = p.x0[((size_t)p.iv_rec[i1] * D + c) * p.C + chain];
            x[nt][1] = p.x0[((size_t)p.iv_rec[i1] * D + c + 1) * p.C + chain];
        } else {
            end_value2(i1 - 1, c, x[nt][0], x[nt][1]);
        }
        xa[nt][0] = x[nt][0]; xa[nt][1] = x[nt][1];
        xb[nt][0] = 0.0; xb[nt][1] = 0.0;
        if (pinned) end_value2(i2, c, xb[nt][0], xb[nt][1]);
        *reinterpret_cast<double2 *>(xbuf + c) = make_double2(x[nt][0], x[nt][1]);
        *reinterpret_cast<double2 *>(ast + c) = make_double2(x[nt][0], x[nt][1]);   // A_0
        *reinterpret_cast<double2 *>(xbuf + ROW + c) = make_double2(xb[nt][0], xb[nt][1]);  // x_b, for the block constant
    }
    __syncwarp();
    // pinned block: g' x_b + 1/2 x_b' Qc x_b of the block constant (once per block, plain FMAs)
    double pin_part = 0.0;
    if (pinned) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = 8 * nt + 2 * q + e;
                double qx = 0.0;
#pragma unroll 1
                for (int kk = 0; kk < D; ++kk) qx += p.blk_Qc[(size_t)blk * D * D + (size_t)kk * D + c] * xbuf[ROW + kk];
                pin_part += p.blk_g[(size_t)blk * D + c] * xb[nt][e] + 0.5 * xb[nt][e] * qx;
            }
    }
    __syncwarp();

    // accepted-path cursor: the state after step (sc.j, sc.k) is A_{t+2} when step t is computed
    Cursor sc{i1, 0, p.iv_N[i1]};
    const double *sXa = nullptr;
    auto state_base = [&](int j) {
        const int s = p.selX_in[(size_t)j * p.C + chain];
        return p.X[s] + p.iv_xoff[j] + (size_t)chain * p.iv_xs[j] + 2 * q;
    };
    auto fetch_state = [&](int stage) {  // own five pairs of the state after step (sc.j, sc.k) -> ring stage
        if (sc.j <= i2) {
            if (sc.k == 0) sXa = state_base(sc.j);
            const double *src = sXa + (size_t)sc.k * D;
            const uint32_t dst = sm_u32(ast + stage * ROW + 2 * q);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) cp_async16(dst + nt * 64, src + 8 * nt);
            advance(sc);
        }
        cp_async_commit();  // always: the waits count groups
    };
    if constexpr (NEED_XA) fetch_state(1);  // A_1

    double llp = 0.0, lla = 0.0, lsum = 0.0, lsum_a = 0.0;  // lsum*: per-lane partial sums of the Girsanov integrals
    int t = 0, ts = 0, tpar = 0;   // step within the block, its table stage and that stage's parity
    int a0 = 0;                    // ring stage of A_t

    for (int j = i1; j <= i2; ++j) {
        const int N = p.iv_N[j];
        const double rho = (MODE == MODE_IMPUTE) ? p.rho[j] : 1.0;
        const double lam = (MODE == MODE_IMPUTE) ? sqrt(1.0 - rho * rho) : 0.0, rho_isig = rho * isig;
        const int sx = p.selX_in[(size_t)j * p.C + chain], sw = p.selW_in[(size_t)j * p.C + chain];
        const double *Wa = p.W[sw];
        double *Wdst = (MODE == MODE_RELAYOUT) ? p.W[sw] : p.W[1 - sw];
        double *Xp = p.X[1 - sx] + p.iv_xoff[j] + (size_t)chain * p.iv_xs[j] + 2 * q;
        const double *aB = p.auxB + (size_t)j * D * D, *abeta = p.auxbeta + (size_t)j * D;
        const size_t wt0 = (size_t)p.iv_tile_off[j];

        for (int k = 0; k < N; ++k) {
            // ---- noise of the next four steps: one Philox call per owned component -> shared memory ----
            if constexpr (ZSM) {
                if ((k & 3) == 0) {
#pragma unroll 1
                    for (int e = 0; e < 2; ++e)  // half rolled: five independent Philox / Box-Muller chains in flight, and
#pragma unroll                                   // the step body still fits the instruction cache
                        for (int nt = 0; nt < NT; ++nt) {
                            float z4[4];
                            philox_normal4f(pkeys, p.iter, (uint32_t)(j + p.interval_offset), gchain, (uint32_t)(8 * nt + 2 * q + e), (uint32_t)(k >> 2), z4);
#pragma unroll
                            for (int s4 = 0; s4 < 4; ++s4) zsm[(s4 * 10 + nt * 2 + e) * 32] = z4[s4];
                        }
                }
            }
            __syncwarp();  // the proposal state written by the previous step is visible; everyone is done with A_{t-1}
            const int a1 = (a0 + 1 == AST) ? 0 : a0 + 1, a2 = (a1 + 1 == AST) ? 0 : a1 + 1;
            if constexpr (NEED_XA) {
                fetch_state(a2);
                cp_async_wait<1>();  // A_{t+1} has landed (this lane's copies) ...
                __syncwarp();        // ... and every lane's
            }
            mbar_wait(&full[ts], (uint32_t)tpar);
            const double *rec = dsm + ts * REC;

            // ---- the three mat-vecs on the tensor cores; the accumulator of Psi' x_b starts from F0, so it ends as
            //      the guiding term's F = F0 + Psi' x_b ----
            double Hx[NT][2], Hxa[NT][2], Fv[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                Hx[nt][0] = Hx[nt][1] = 0.0; Hxa[nt][0] = Hxa[nt][1] = 0.0;
                const double2 f = *reinterpret_cast<const double2 *>(rec + OFF_F + 8 * nt + 2 * q);
                Fv[nt][0] = f.x; Fv[nt][1] = f.y;
            }
            {
                const double2 *hf = reinterpret_cast<const double2 *>(rec + OFF_H) + lane;
                const double2 *pf = reinterpret_cast<const double2 *>(rec + OFF_PSI) + lane;
#pragma unroll
                for (int kp = 0; kp < NT; ++kp)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        // the two k-steps of a tile go to the same accumulator: keep them apart (DMMA latency 26 cycles,
                        // issue interval 16)
                        const double2 h = hf[(kp * NT + nt) * 32];
                        double2 pp;
                        if constexpr (PSI) pp = pf[(kp * NT + nt) * 32];
                        if constexpr (SOLVE) dmma(Hx[nt], x[kp][0], h.x);
                        if constexpr (NEED_XA) dmma(Hxa[nt], xa[kp][0], h.x);
                        if constexpr (PSI) dmma(Fv[nt], xb[kp][0], pp.x);
                        if constexpr (SOLVE) dmma(Hx[nt], x[kp][1], h.y);
                        if constexpr (NEED_XA) dmma(Hxa[nt], xa[kp][1], h.y);
                        if constexpr (PSI) dmma(Fv[nt], xb[kp][1], pp.y);
                    }
            }
            const double dt = rec[0], lamsq = lam * rec[1];
            // ---- done with the table stage (warp 0 refills it at the end of the step, when every warp has said so) ----
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[ts]);
            if (t == 0) {
                // log h~(t_a, x_a) = -c - 1/2 x'Hx + F'x at the block's first grid point (x = x_acc there)
                double part = 0.0;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double hx = SOLVE ? Hx[nt][e] : Hxa[nt][e];
                        part += Fv[nt][e] * x[nt][e] - 0.5 * x[nt][e] * hx;
                    }
                part -= pin_part;
                llp = quad_sum(part) - p.blk_c0[blk];
                lla = llp;
            }

            // ---- element-wise part, one tile of eight components at a time (short live ranges).  Neighbours of the
            //      Lorenz-96 drift: accepted state A_t from its ring stage, proposal from the double buffer ----
            const double *arow = ast + a0 * ROW, *nrow = ast + a1 * ROW + 2 * q;   // chain row; own pair of tile 0
            const double *xrow = xbuf + (t & 1) * ROW;
            double *xout = xbuf + ((t + 1) & 1) * ROW + 2 * q;
            const double *aq = arow + 2 * q, *xq = xrow + 2 * q;
            const double *aL0 = aq + relL0, *aR4 = aq + relR4, *xL0 = xq + relL0, *xR4 = xq + relR4;
            const bool pin_here = pinned && j == i2 && k == N - 1;  // the block's right end stays x_b
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                double dwv[2] = {0.0, 0.0}, xnew[2] = {0.0, 0.0};
                if constexpr (NEED_XA) {  // recover the accepted noise increments
                    const double2 l = *reinterpret_cast<const double2 *>(nt == 0 ? aL0 : aq + 8 * nt - 2);
                    const double rgt = reinterpret_cast<const double2 *>(nt == NT - 1 ? aR4 : aq + 8 * nt + 2)->x;  // 16-byte access: no bank conflict
                    const double2 nx = *reinterpret_cast<const double2 *>(nrow + 8 * nt);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = 8 * nt + 2 * q + e;
                        const double r = Fv[nt][e] - Hxa[nt][e];
                        // Lorenz-96: b_i = (x_{i+1} - x_{i-2}) x_{i-1} - x_i + F
                        const double xi = xa[nt][e];
                        const double xip = e == 0 ? xa[nt][1] : rgt;
                        const double xim1 = e == 0 ? l.y : xa[nt][0];
                        const double xim2 = e == 0 ? l.x : l.y;
                        const double dd = xip - xim2;
                        const double b = dd * xim1 - xi + Fth;
                        // built-in aux law B = -I + c (S_{+1} - S_{-2}), beta = F:  b - b~ = (x_{i+1} - x_{i-2}) (x_{i-1} - c)
                        const double bmbt = FAST ? dd * (xim1 - caux) : b - aux_drift_general(aB, abeta, c, arow);
                        lsum_a += (bmbt * r) * dt;
                        const double res = (e == 0 ? nx.x : nx.y) - xi - (b + sig2 * r) * dt;
                        dwv[e] = (MODE == MODE_IMPUTE) ? res : res * isig;  // imputation: 1 / sigma is folded into the pCN mix
                    }
                    xa[nt][0] = nx.x; xa[nt][1] = nx.y;
                } else if constexpr (NEED_WA) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) dwv[e] = Wa[w_off(wt0 + (size_t)(k >> 3), chain, 8 * nt + 2 * q + e, p.C, D) + (k & 7)];
                }
                if constexpr (MODE == MODE_IMPUTE) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double z;
                        if constexpr (ZSM) z = (double)zsm[((k & 3) * 10 + nt * 2 + e) * 32];
                        else z = p.Z[w_off(wt0 + (size_t)(k >> 3), chain, 8 * nt + 2 * q + e, p.C, D) + (k & 7)];
                        dwv[e] = lamsq * z + (NEED_XA ? rho_isig : rho) * dwv[e];  // dW deg = lambda sqrt(dt) Z + rho dW
                    }
                }
                if constexpr (SOLVE) {
                    const double2 l = *reinterpret_cast<const double2 *>(nt == 0 ? xL0 : xq + 8 * nt - 2);
                    const double rgt = reinterpret_cast<const double2 *>(nt == NT - 1 ? xR4 : xq + 8 * nt + 2)->x;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = 8 * nt + 2 * q + e;
                        const double r = Fv[nt][e] - Hx[nt][e];
                        const double xi = x[nt][e];
                        const double xip = e == 0 ? x[nt][1] : rgt;
                        const double xim1 = e == 0 ? l.y : x[nt][0];
                        const double xim2 = e == 0 ? l.x : l.y;
                        const double dd = xip - xim2;
                        const double b = dd * xim1 - xi + Fth;
                        const double bmbt = FAST ? dd * (xim1 - caux) : b - aux_drift_general(aB, abeta, c, xrow);
                        lsum += (bmbt * r) * dt;
                        xnew[e] = xi + (b + sig2 * r) * dt + sig * dwv[e];
                    }
                    // new state: registers (next step's A operands), the other half of the proposal buffer, global
                    x[nt][0] = xnew[0]; x[nt][1] = xnew[1];
                    *reinterpret_cast<double2 *>(xout + 8 * nt) = make_double2(xnew[0], xnew[1]);
                    if (active) {
                        if (pin_here) st_v2(Xp + (size_t)k * D + 8 * nt, xb[nt][0], xb[nt][1]);
                        else st_v2(Xp + (size_t)k * D + 8 * nt, xnew[0], xnew[1]);
                    }
                }
                if constexpr (STORE_W_MODE) {
                    if (active) {
#pragma unroll
                        for (int e = 0; e < 2; ++e)
                            Wdst[w_off(wt0 + (size_t)(k >> 3), chain, 8 * nt + 2 * q + e, p.C, D) + (k & 7)] = dwv[e];
                    }
                }
            }
            // ---- warp 0: the record NST steps ahead goes into the stage this step has used.  No fence and no atomics
            //      (a membar here also waits for the step's global stores and prefetches);
            //      by now the other warps are normally past their mat-vecs, so the wait does not spin ----
            if (tc.j <= i2) {
                if (threadIdx.x == 0) {
                    mbar_wait(&empty[ts], (uint32_t)tpar);
                    issue_table(tc, ts);
                }
                advance(tc);
            }
            ++t;
            if (++ts == NST) { ts = 0; tpar ^= 1; }
            a0 = a1;
        }
    }
    if constexpr (NEED_XA) cp_async_wait<0>();
    llp += quad_sum(lsum);
    lla += quad_sum(lsum_a);

    if (!in_range) return;
    if constexpr (MODE == MODE_IMPUTE) {
        if (masked) {
            for (int j = i1 + q; j <= i2; j += 4) {
                const size_t o = (size_t)j * p.C + chain;
                p.selX_out[o] = p.selX_in[o];
                p.selW_out[o] = p.selW_in[o];
            }
            if (q == 0 && p.blk_active && !p.blk_active[blk]) {
                const double nan = __longlong_as_double(0x7ff8000000000000LL);
                if (p.out_llprop) p.out_llprop[(size_t)chain * p.out_stride + blk] = nan;
                if (p.out_ll) p.out_ll[(size_t)chain * p.out_stride + blk] = nan;
                if (p.out_acc) p.out_acc[(size_t)chain * p.out_stride + blk] = 0;
            }
            return;
        }
        double llcur = REINV ? lla : p.ll[uo];
        const double ex = PHILOX ? philox_exponential(p.seed, p.iter, (uint32_t)(i1 + p.interval_offset), gchain) : p.E[uo];
        const bool ok = isfinite(llp);
        const bool acc = p.force_accept ? ok : (ex > -(llp - llcur));  // identical on the four lanes of a chain
        for (int j = i1 + q; j <= i2; j += 4) {
            const size_t o = (size_t)j * p.C + chain;
            p.selX_out[o] = p.selX_in[o] ^ (uint8_t)acc;
            p.selW_out[o] = p.selW_in[o] ^ (uint8_t)(acc && STORE_W_MODE);
        }
        if (q == 0) {
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
        if (q == 0) {
            if (p.out_llprop) p.out_llprop[uo] = llp;
            if (p.out_acc) p.out_acc[uo] = (uint8_t)isfinite(llp);
        }
    } else {
        if (q == 0) p.ll[uo] = lla;
    }
}

template <bool PSI>
static cudaError_t launch_l96(int mode, const SolveParams &p, cudaStream_t s) {
    if ((long long)p.nblk * p.C == 0) return cudaSuccess;
    // warps per CTA (eight chains each, all on one block: the per-step table record is staged once per CTA): the largest
    // CTA whose grid still covers every SM; DMCMC_L96_WARPS pins it
    // (one CTA per SM: 255 registers, 52 KB of table stages + 18 KB per warp).  Smaller CTAs only while all of them are
    // resident at once (measured, 33 blocks x 256 chains: 132 CTAs of 8 warps 20.6 ms, 264 CTAs of 4 warps = two rounds 25 ms)
    int warps = 8;
    while (warps > 1 && (warps / 2) * 8 >= p.C) warps >>= 1;  // no CTA of nothing but surplus warps
    while (warps > 1 && (long long)p.nblk * ((p.C + warps * 4 - 1) / (warps * 4)) <= 148) warps >>= 1;
    if (const char *env = getenv("DMCMC_L96_WARPS")) {
        const int w = atoi(env);
        if (w >= 1 && w <= 8) warps = w;
    }
    const unsigned groups = (unsigned)((p.C + warps * 8 - 1) / (warps * 8));
    const unsigned grid = groups * (unsigned)p.nblk;
    const unsigned threads = warps * 32;
    const size_t smem = l96_smem_bytes(PSI, warps);
    const bool philox = (p.Z == nullptr), reinv = p.reinvert != 0;
#define DMCMC_LAUNCH(MODE_, PH_, RE_)                                                                         \
    do {                                                                                                       \
        auto kern_ = p.fast_ok ? solve_l96_kernel<PSI, MODE_, PH_, RE_, true> : solve_l96_kernel<PSI, MODE_, PH_, RE_, false>; \
        cudaError_t e_ = cudaFuncSetAttribute(kern_, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); \
        if (e_ != cudaSuccess) return e_;                                                                      \
        kern_<<<grid, threads, smem, s>>>(p);                                                                  \
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
    return p.has_psi ? launch_l96<true>(mode, p, s) : launch_l96<false>(mode, p, s);
}

}  // namespace dmcmc
