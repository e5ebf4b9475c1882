// dmcmc_solve_common.cuh -- building blocks of the small-d solve kernel (dmcmc_solve.cu):
// guided step, FitzHugh-Nagumo fused step, TMA / mbarrier plumbing, the LDGSTS input stage.
#pragma once
#include <algorithm>
#include <cmath>
#include <type_traits>

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

// ---- one guided step: Girsanov integrand G and the drift part (b + a r) dt -------------------
// (callers accumulate ll = fma(G, dt, ll) explicitly, so that every kernel rounds the sum the same way)
template <class MD, bool PSI>
__device__ __forceinline__ void guided_terms(const Theta &th, const double *__restrict__ rec,
                                             const double *F, const double *aB,
                                             const double *abeta, const double *aatil,
                                             const double *x, double dt, double *drift_dt,
                                             double &gout) {
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
    gout = g;
    MD::a_mul(th, x, r, ar);
#pragma unroll
    for (int i = 0; i < D; ++i) drift_dt[i] = (b[i] + ar[i]) * dt;
}

// ---- FitzHugh-Nagumo fused step (built-in FitzHughNagumoAux only) ------------------------------
// With the aux law's second row equal to the target's (B = [., .; gamma, -1], beta~_2 = beta) and
// B_12 = -1/eps, the Girsanov integrand collapses to
//     (b - b~)'r = [ (y - y^3)/eps - B_11 y - (beta~_1 - s/eps) ] r_1 ,
// and a r = (0, sigma^2 r_2), so the second coordinate is an affine map of the state,
//     x' = A1 x + B1 y + Cf + sigma dW ,   A1 = 1 - dt (1 + sigma^2 H_22),  B1 = dt (gamma - sigma^2 H_12),
//     Cf = dt (beta + sigma^2 F_2) ,       F = F0 + Psi x_b .
// A1, B1 and the dt-scaled rows of H, F0, Psi are the same for every chain: fhn_derive rewrites
// the table once per (layout, law) into p.table (fhn_derive_kernel); p.table_raw keeps H, F0, Psi.  Same mathematics as guided_terms, ~3x fewer FP64
// instructions per chain and step; differences are rounding-level.
struct FhnK {
    double ie, s, gam, bet, sig, sig2, isig;  // per law
    double aB0, c2;                            // per interval
    __device__ __forceinline__ void set_law(const Theta &th) {
        ie = 1.0 / th.v[0]; s = th.v[1]; gam = th.v[2]; bet = th.v[3]; sig = th.v[4];
        sig2 = sig * sig; isig = 1.0 / sig;
    }
    __device__ __forceinline__ void set_interval(const double *auxB, const double *auxbeta) {
        aB0 = auxB[0];
        c2 = auxbeta[0] - s * ie;
    }
};

// derived record: [A1, B1, dt H11, dt H12, dt/eps, sqrt(dt) | C1, dt F0_1 | dt s2 Psi21, dt s2 Psi22, dt Psi11, dt Psi12]
// (the first six are all the recursion itself reads)
enum { FD_A1 = 0, FD_B1, FD_G00, FD_G01, FD_KIE, FD_SQ, FD_C1, FD_R0, FD_Q10, FD_Q11, FD_T00, FD_T01 };

template <bool PSI>
__device__ __forceinline__ void fhn_derive(const double *raw, double *r, const FhnK &fk) {
    using R = Rec<2, PSI>;
    const double dt = raw[0], sq = raw[1];
    const double H00 = raw[R::OFF_H], H01 = raw[R::OFF_H + 1], H11 = raw[R::OFF_H + 2];
    const double F0 = raw[R::OFF_F], F1 = raw[R::OFF_F + 1];
    double P00 = 0.0, P01 = 0.0, P10 = 0.0, P11 = 0.0;
    if constexpr (PSI) { P00 = raw[R::OFF_PSI]; P01 = raw[R::OFF_PSI + 1]; P10 = raw[R::OFF_PSI + 2]; P11 = raw[R::OFF_PSI + 3]; }
    const double ds2 = dt * fk.sig2;
    r[FD_A1] = fma(-ds2, H11, 1.0 - dt);
    r[FD_B1] = fma(-ds2, H01, dt * fk.gam);
    r[FD_C1] = fma(ds2, F1, dt * fk.bet);
    r[FD_KIE] = dt * fk.ie;
    r[FD_G00] = dt * H00;
    r[FD_G01] = dt * H01;
    r[FD_R0] = dt * F0;
    r[FD_SQ] = sq;
    if constexpr (PSI) {
        r[FD_Q10] = ds2 * P10; r[FD_Q11] = ds2 * P11;
        r[FD_T00] = dt * P00;  r[FD_T01] = dt * P01;
    }
}

// the two factors of the Girsanov increment G dt = d1 * r1dt at state (y, x); p = 1 - y^2 is returned for the drift
__device__ __forceinline__ void fhn_dll_terms(const FhnK &fk, const double *rec, double R0c, double y, double x, double &p,
                                              double &d1, double &r1dt) {
    p = fma(-y, y, 1.0);
    const double g = y * p;  // y - y^3
    d1 = fma(fk.ie, g, fma(-fk.aB0, y, -fk.c2));
    r1dt = fma(-rec[FD_G00], y, fma(-rec[FD_G01], x, R0c));
}
// adds the Girsanov increment G dt at state (y, x) to ll
__device__ __forceinline__ void fhn_dll(const FhnK &fk, const double *rec, double R0c, double y, double x, double &p, double &ll) {
    double d1, r1dt;
    fhn_dll_terms(fk, rec, R0c, y, x, p, d1, r1dt);
    ll = fma(d1, r1dt, ll);
}
// one fused step from (y, x): ll += G dt, then
//     y' = y + (dt/eps) (y - y^3 + s - x) = [y + (dt/eps)(s - x)] + ((dt/eps) y)(1 - y^2)
//     x' = A1 x + B1 y + ev ,   ev = sigma dW + C
// written so that the recurrence y -> y' is two dependent FMAs deep (1 - y^2, then the sum): this chain is the
// critical path of every latency-bound launch (few units; the warp-per-unit kernel)
__device__ __forceinline__ void fhn_step(const FhnK &fk, const double *rec, double R0c, double ev, double &y, double &x, double &ll) {
    double p;
    fhn_dll(fk, rec, R0c, y, x, p, ll);
    const double a = fma(rec[FD_KIE], fk.s - x, y);
    const double ynew = fma(rec[FD_KIE] * y, p, a);
    x = fma(rec[FD_A1], x, fma(rec[FD_B1], y, ev));
    y = ynew;
}

// ---- TMA bulk copy + mbarrier plumbing (sm_90+/sm_100a PTX) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
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

// state after the last Euler step of interval j in the accepted container
template <int D>
__device__ __forceinline__ void load_end(const SolveParams &p, int chain, int j, double *x) {
    const int s = p.selX_in[(size_t)j * p.C + chain];
    const double *X = p.X[s] + p.iv_xoff[j] + (size_t)chain * p.iv_xs[j] + (size_t)(p.iv_N[j] - 1) * D;
#pragma unroll
    for (int c = 0; c < D; ++c) x[c] = X[c];
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

// A/B knobs (defaults are what ships).  Measured on the C3 sweep, B200, CTA of 128 threads: TS 16 / 168 registers 76.2 us;
// TS 16 / 128 registers 77.1; TS 8 / 160 registers 76.5; TS 8 / 128 registers 77.1 -- the sweep has 51,200 units, i.e.
// 10.8 warps per SM whatever the register or shared-memory budget allows, and a warp's own in-order dependent-issue
// latency (~7 cycles per instruction) sets the time, so neither knob moves it.
#ifndef DMCMC_TS
#define DMCMC_TS 16
#endif
#ifndef DMCMC_MINB
#define DMCMC_MINB 3
#endif
constexpr int TS = DMCMC_TS;        // Euler steps per staged tile: 256 contiguous bytes per chain for d = 2
constexpr int MAX_DEPTH = 8;  // deepest software pipeline (tiles in flight) of the table / input stages

// ---- shared-memory stage of a chain's input stream (accepted path, or stored accepted noise) ------
// The stream of one (chain, tile) is cut into 16-byte pieces that 16 (d = 2) consecutive lanes copy
// with one LDGSTS each: 256 contiguous bytes per chain in HBM and in shared memory.
//   XA = true : accepted path,  stream index of (step k, component c) = k * D + c
//   XA = false: accepted noise, stream index of (step k, Wiener comp w) = ((k / 8) * M + w) * 8 + k % 8
//               (two tiles of the W container)
// d = 2 path stages are unpadded (256 B per chain, 8 KB per warp: six CTAs per SM) and XOR-swizzled:
// piece pc of chain r sits in 16-byte slot pc ^ (r & 7) of its record, so that the eight lanes of an
// LDS.128 phase (same step, eight chains) hit eight different bank groups.  Other shapes pad the
// record by 16 bytes instead.
template <int D, int M, bool XA>
struct InStage {
    static constexpr int NPT = XA ? TS * D / 2 : (TS / TILE) * M * 4;  // pieces per (tile, chain)
    static constexpr bool SWZ = XA && D == 2;
    static constexpr int RECW = 2 * NPT + (SWZ ? 0 : 2);               // doubles per chain
    static constexpr int WARP = 32 * RECW;                             // doubles per warp and stage
    __device__ static __forceinline__ int piece(int r, int pc) {
        return SWZ ? r * RECW + ((pc ^ (r & 7)) << 1) : r * RECW + pc * 2;
    }
    // per-lane read cursor of a stage: byte address with the swizzle folded in (stage is 256-byte aligned)
    __device__ static __forceinline__ uint32_t cursor(const double *stage, int lane) {
        const uint32_t a = smem_u32(stage) + (uint32_t)lane * RECW * 8u;
        return SWZ ? a ^ ((uint32_t)(lane & 7) << 4) : a;
    }
    // state after step k of the tile (XA)
    __device__ static __forceinline__ void state(uint32_t cur, int k, double *x) {
        if constexpr (SWZ) {
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x[0]), "=d"(x[1]) : "r"(cur ^ ((uint32_t)k << 4)));
        } else {
#pragma unroll
            for (int c = 0; c < D; ++c)
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x[c]) : "r"(cur + (uint32_t)(k * D + c) * 8u));
        }
    }
    // accepted noise increment of step k (!XA)
    __device__ static __forceinline__ void noise(uint32_t cur, int k, double *dw) {
#pragma unroll
        for (int w = 0; w < M; ++w)
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(dw[w]) : "r"(cur + (uint32_t)(((k >> 3) * M + w) * 8 + (k & 7)) * 8u));
    }
};

// per-interval metadata of one unit, held in registers one interval ahead: under a loaded memory
// system every dependent global load at an interval switch costs microseconds
struct IvMeta {
    int N, tile_off, xs, sx, sw;
    long long xoff;
    double rho, aB0, ab0;
};
__device__ __forceinline__ IvMeta load_meta(const SolveParams &p, int j, int chain, int D) {
    IvMeta m;
    m.N = p.iv_N[j]; m.tile_off = p.iv_tile_off[j]; m.xs = p.iv_xs[j]; m.xoff = p.iv_xoff[j];
    m.sx = p.selX_in[(size_t)j * p.C + chain]; m.sw = p.selW_in[(size_t)j * p.C + chain];
    m.rho = p.rho[j]; m.aB0 = p.auxB[(size_t)j * D * D]; m.ab0 = p.auxbeta[(size_t)j * D];
    return m;
}

__device__ __forceinline__ void cp_async_wait_depth(int depth) {  // allow depth - 2 groups in flight
    if (depth == 2) cp_async_wait<0>();
    else if (depth == 4) cp_async_wait<2>();
    else cp_async_wait<6>();
}

}  // namespace dmcmc
