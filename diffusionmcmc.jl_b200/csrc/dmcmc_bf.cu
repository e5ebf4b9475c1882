// dmcmc_bf.cu -- a8: the backward filter (backward_filter!, /root/reference
// src/run!_alterations.jl:203; recompute_guiding_term!, src/workspaces.jl:428) as a batched
// small-matrix recursion, one thread per layout block.
//
// The ODEs  H' = -B'H - HB + H a~ H,  F' = -B'F + H a~ F + H beta,  c' = beta'F + F'a~F/2 - tr(H a~)/2
// with (B, beta, a~) constant on an interval have the exact one-step flow
//     K = (I + H+ Q)^-1,  H* = K H+,  F* = K F+,
//     H = Phi' H* Phi,    F = Phi'(F* - H* m),
//     c = c+ + log det(I + H+ Q)/2 + m'H* m/2 - m'F* - F+' Q F*/2
// where x(t+dt) | x(t) ~ N(Phi x + m, Q) is the aux law's transition.  For pinned blocks the
// dependence on the pinned end point x_b is carried as F = F0 + Psi x_b, c = c0 + g'x_b + x_b'Qc x_b/2
// so that one table serves every chain.  Same operation order as oracle/dmcmc_oracle.c.
#include <cmath>
#include <cstdlib>

#include "dmcmc_device.cuh"
#include "dmcmc_internal.h"

namespace dmcmc {

// Every loop below has compile-time bounds and is unrolled, and the row exchanges of the pivoted LU are written as
// predicated swaps over compile-time indices: with a run-time row index the matrices would live in local memory,
// and this recursion is serial per layout block -- its latency IS the run time of a critical parameter step.
template <int D>
struct SmallMat {
    __device__ static __forceinline__ void mm(const double *A, const double *B, double *C) {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) s += A[i * D + k] * B[k * D + j];
                C[i * D + j] = s;
            }
    }
    __device__ static __forceinline__ void mmT(const double *A, const double *B, double *C) {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) s += A[i * D + k] * B[j * D + k];
                C[i * D + j] = s;
            }
    }
    __device__ static __forceinline__ void mTm(const double *A, const double *B, double *C) {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) s += A[k * D + i] * B[k * D + j];
                C[i * D + j] = s;
            }
    }
    __device__ static __forceinline__ void mv(const double *A, const double *x, double *y) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) s += A[i * D + k] * x[k];
            y[i] = s;
        }
    }
    __device__ static __forceinline__ void mTv(const double *A, const double *x, double *y) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) s += A[k * D + i] * x[k];
            y[i] = s;
        }
    }
    __device__ static __forceinline__ double dot(const double *a, const double *b) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) s += a[i] * b[i];
        return s;
    }
    __device__ static __forceinline__ void sym(double *A) {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = i + 1; j < D; ++j) {
                const double s = 0.5 * (A[i * D + j] + A[j * D + i]);
                A[i * D + j] = s;
                A[j * D + i] = s;
            }
    }
    // LU with partial pivoting in place (pivot = first row of maximal modulus, as in the oracle).  invd = reciprocals
    // of the pivots (each division is done once and shared by every solve against this factorisation); lad = log|det|
    // taken from the product of the pivots (one log per factorisation instead of D)
    __device__ static __forceinline__ void lu(double *A, int *piv, double *invd, double &lad) {
        double prod = 1.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            int p = k;
            double best = fabs(A[k * D + k]);
#pragma unroll
            for (int i = k + 1; i < D; ++i)
                if (fabs(A[i * D + k]) > best) { best = fabs(A[i * D + k]); p = i; }
            piv[k] = p;
#pragma unroll
            for (int i = k + 1; i < D; ++i)
                if (p == i) {
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        const double t = A[k * D + j]; A[k * D + j] = A[i * D + j]; A[i * D + j] = t;
                    }
                }
            prod *= fabs(A[k * D + k]);
            const double inv = 1.0 / A[k * D + k];
            invd[k] = inv;
#pragma unroll
            for (int i = k + 1; i < D; ++i) {
                const double l = A[i * D + k] * inv;
                A[i * D + k] = l;
#pragma unroll
                for (int j = k + 1; j < D; ++j) A[i * D + j] -= l * A[k * D + j];
            }
        }
        lad = log(prod);
    }
    template <int NR>
    __device__ static __forceinline__ void solve(const double *LU, const int *piv, const double *invd, double *Bm) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
#pragma unroll
            for (int i = k + 1; i < D; ++i)
                if (piv[k] == i) {
#pragma unroll
                    for (int j = 0; j < NR; ++j) {
                        const double t = Bm[k * NR + j]; Bm[k * NR + j] = Bm[i * NR + j]; Bm[i * NR + j] = t;
                    }
                }
#pragma unroll
            for (int i = k + 1; i < D; ++i) {
                const double l = LU[i * D + k];
#pragma unroll
                for (int j = 0; j < NR; ++j) Bm[i * NR + j] -= l * Bm[k * NR + j];
            }
        }
#pragma unroll
        for (int k = D - 1; k >= 0; --k) {
            const double inv = invd[k];
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                double s = Bm[k * NR + j];
#pragma unroll
                for (int i = k + 1; i < D; ++i) s -= LU[k * D + i] * Bm[i * NR + j];
                Bm[k * NR + j] = s * inv;
            }
        }
    }
};

constexpr int TAYLOR_TERMS = 16;

// transition (Phi, Q, m) of dx = (Bx + beta)dt + sigma~ dW over dt: Taylor series of
// Phi = e^{B dt}, m = int_0^dt e^{Bs} beta ds, Q = int_0^dt e^{Bs} a~ e^{B's} ds with
// scaling-and-doubling.  Mirrors orc_transition.
template <int D>
__device__ void transition(const double *B, const double *beta, const double *atil, double dt,
                           double *Phi, double *Q, double *mvec) {
    using S = SmallMat<D>;
    double nrm = 0.0;
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
        for (int j = 0; j < D; ++j) s += fabs(B[i * D + j]);
        if (s > nrm) nrm = s;
    }
    int sq = 0;
    double h = dt;
    while (nrm * h > 0.25 && sq < 40) { h *= 0.5; ++sq; }
    double Pn[D * D], An[D * D], T1[D * D], T2[D * D], bn[D], tv[D];
    for (int i = 0; i < D * D; ++i) { Phi[i] = 0.0; Pn[i] = 0.0; }
    for (int i = 0; i < D; ++i) { Phi[i * D + i] = 1.0; Pn[i * D + i] = 1.0; }
    for (int i = 0; i < D * D; ++i) An[i] = atil[i];
    for (int i = 0; i < D; ++i) bn[i] = beta[i];
    double coef = h;
    for (int i = 0; i < D * D; ++i) Q[i] = coef * An[i];
    for (int i = 0; i < D; ++i) mvec[i] = coef * bn[i];
    for (int n = 1; n <= TAYLOR_TERMS; ++n) {
        S::mm(B, Pn, T1);
        for (int i = 0; i < D * D; ++i) { Pn[i] = T1[i] * (h / n); Phi[i] += Pn[i]; }
        S::mm(B, An, T1);
        S::mmT(An, B, T2);
        for (int i = 0; i < D * D; ++i) An[i] = T1[i] + T2[i];
        S::mv(B, bn, tv);
        for (int i = 0; i < D; ++i) bn[i] = tv[i];
        coef = coef * h / (n + 1);
        for (int i = 0; i < D * D; ++i) Q[i] += coef * An[i];
        for (int i = 0; i < D; ++i) mvec[i] += coef * bn[i];
    }
    for (int s = 0; s < sq; ++s) {
        S::mm(Phi, Q, T1);
        S::mmT(T1, Phi, T2);
        for (int i = 0; i < D * D; ++i) Q[i] += T2[i];
        S::mv(Phi, mvec, tv);
        for (int i = 0; i < D; ++i) mvec[i] += tv[i];
        S::mm(Phi, Phi, T1);
        for (int i = 0; i < D * D; ++i) Phi[i] = T1[i];
    }
    S::sym(Q);
}

template <int D>
struct BFState {
    double H[D * D], F0[D], Psi[D * D], c0, g[D], Qc[D * D];
};

template <int D>
__device__ void bf_step(bool pinned, const double *Phi, const double *Q, const double *mvec,
                        BFState<D> &s) {
    using S = SmallMat<D>;
    double A[D * D], Hs[D * D], T1[D * D], Psis[D * D], Fs[D], tv[D], tv2[D];
    int piv[D];
    double lad, invd[D];
    S::mm(s.H, Q, A);
    for (int i = 0; i < D; ++i) A[i * D + i] += 1.0;
    S::lu(A, piv, invd, lad);
    for (int i = 0; i < D * D; ++i) Hs[i] = s.H[i];
    S::template solve<D>(A, piv, invd, Hs);
    S::sym(Hs);
    for (int i = 0; i < D; ++i) Fs[i] = s.F0[i];
    S::template solve<1>(A, piv, invd, Fs);
    S::mv(Hs, mvec, tv);
    S::mv(Q, Fs, tv2);
    s.c0 = s.c0 + 0.5 * lad + 0.5 * S::dot(mvec, tv) - S::dot(mvec, Fs) - 0.5 * S::dot(s.F0, tv2);
    if (pinned) {
        for (int i = 0; i < D * D; ++i) Psis[i] = s.Psi[i];
        S::template solve<D>(A, piv, invd, Psis);
        double a1[D], a2[D];
        S::mTv(Psis, mvec, a1);
        S::mTv(s.Psi, tv2, a2);
        for (int i = 0; i < D; ++i) s.g[i] = s.g[i] - a1[i] - a2[i];
        S::mm(Q, Psis, T1);
        S::mTm(s.Psi, T1, A);
        for (int i = 0; i < D * D; ++i) s.Qc[i] -= A[i];
        S::sym(s.Qc);
        S::mTm(Phi, Psis, s.Psi);
    }
    for (int i = 0; i < D; ++i) tv2[i] = Fs[i] - tv[i];
    S::mTv(Phi, tv2, s.F0);
    S::mm(Hs, Phi, T1);
    S::mTm(Phi, T1, s.H);
    S::sym(s.H);
}

template <int D>
__device__ void store_point(const BFParams &b, const BFState<D> &s, size_t pt) {
    if (b.ptH)
        for (int i = 0; i < D * D; ++i) b.ptH[pt * D * D + i] = s.H[i];
    if (b.ptF0)
        for (int i = 0; i < D; ++i) b.ptF0[pt * D + i] = s.F0[i];
    if (b.ptPsi)
        for (int i = 0; i < D * D; ++i) b.ptPsi[pt * D * D + i] = s.Psi[i];
    if (b.ptc0) b.ptc0[pt] = s.c0;
}

template <int D>
__device__ void store_rec(const BFParams &b, const BFState<D> &s, size_t step, double dt) {
    double *r = b.table + step * b.rec;
    r[0] = dt;
    r[1] = sqrt(dt);
    int o = 2;
    for (int i = 0; i < D; ++i)
        for (int k = i; k < D; ++k) r[o++] = s.H[i * D + k];
    for (int i = 0; i < D; ++i) r[o++] = s.F0[i];
    if (b.has_psi)
        for (int i = 0; i < D * D; ++i) r[o++] = s.Psi[i];
}

template <int D>
__global__ void bf_kernel(const BFParams b) {
    const int blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= b.nblk) return;
    BFState<D> s;
    for (int i = 0; i < D * D; ++i) { s.H[i] = 0.0; s.Psi[i] = 0.0; s.Qc[i] = 0.0; }
    for (int i = 0; i < D; ++i) { s.F0[i] = 0.0; s.g[i] = 0.0; }
    s.c0 = 0.0;
    const bool pinned = b.blk_pinned[blk] != 0;
    const int i1 = b.blk_i1[blk], i2 = b.blk_i2[blk];
    double Phi[D * D], Q[D * D], mvec[D];
    for (int j = i2; j >= i1; --j) {
        if (j == i2 && pinned) {
            const double ie = 1.0 / b.pin_eps;
            for (int i = 0; i < D; ++i) { s.H[i * D + i] = ie; s.Psi[i * D + i] = ie; s.Qc[i * D + i] = ie; }
            s.c0 = 0.5 * D * log(2.0 * M_PI * b.pin_eps);
        } else {
            for (int i = 0; i < D * D; ++i) s.H[i] += b.Hobs[(size_t)j * D * D + i];
            for (int i = 0; i < D; ++i) s.F0[i] += b.Fobs[(size_t)j * D + i];
            s.c0 += b.cobs[j];
        }
        const int N = b.iv_N[j];
        const size_t base = (size_t)b.iv_pt_off[j];
        const size_t sbase = (size_t)b.iv_tile_off[j] * TILE;
        store_point<D>(b, s, base + N);
        double dt_cached = -1.0;
        for (int k = N - 1; k >= 0; --k) {
            const double dt = b.t[base + k + 1] - b.t[base + k];
            if (!(fabs(dt - dt_cached) <= 1e-9 * fabs(dt_cached))) {
                transition<D>(b.auxB + (size_t)j * D * D, b.auxbeta + (size_t)j * D,
                              b.auxatil + (size_t)j * D * D, dt, Phi, Q, mvec);
                dt_cached = dt;
            }
            bf_step<D>(pinned, Phi, Q, mvec, s);
            store_point<D>(b, s, base + k);
            store_rec<D>(b, s, sbase + k, dt);
        }
        // pad slots of the interval's last tile: benign record (dt = 0)
        for (int k = N; k < ((N + TILE - 1) / TILE) * TILE; ++k) {
            double *r = b.table + (sbase + k) * b.rec;
            for (int i = 0; i < b.rec; ++i) r[i] = 0.0;
        }
    }
    b.blk_c0[blk] = s.c0;
    for (int i = 0; i < D; ++i) b.blk_g[(size_t)blk * D + i] = s.g[i];
    for (int i = 0; i < D * D; ++i) b.blk_Qc[(size_t)blk * D * D + i] = s.Qc[i];
}

// ---------------------------------------------------------------------------------------------
// The same filter, parallel in time inside an interval.  On an interval the aux law (B, beta, a~) is constant,
// so the Gaussian transition from grid point k to the interval's right end is the one-step flow over
// tau_k = t_N - t_k, and the value at EVERY grid point follows from the right-end value by one application of
// bf_step with that transition: the N grid points of an interval are independent, only the chain over intervals is
// serial (J links per block instead of J N).  One CTA per layout block, one thread per grid point of the current
// interval.  Mathematically identical to the step-by-step recursion (both are exact flows of the H, F, c ODEs);
// rounding differs at the 1e-13 level.  A critical parameter update on an unblocked recording -- one block, where
// bf_kernel runs 10^4 steps on one thread -- drops from milliseconds to the latency of J transitions.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) bf_par_kernel(const BFParams b) {
    const int blk = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    __shared__ BFState<D> send[2];  // right-end state of the current interval / of the next one to the left
    const bool pinned = b.blk_pinned[blk] != 0;
    const int i1 = b.blk_i1[blk], i2 = b.blk_i2[blk];
    if (tid == 0) {
        BFState<D> &s = send[0];
        for (int i = 0; i < D * D; ++i) { s.H[i] = 0.0; s.Psi[i] = 0.0; s.Qc[i] = 0.0; }
        for (int i = 0; i < D; ++i) { s.F0[i] = 0.0; s.g[i] = 0.0; }
        s.c0 = 0.0;
    }
    int cur = 0;
    for (int j = i2; j >= i1; --j) {
        const int N = b.iv_N[j];
        const size_t base = (size_t)b.iv_pt_off[j];
        const size_t sbase = (size_t)b.iv_tile_off[j] * TILE;
        if (tid == 0) {  // observation at the interval's right end (or the pin of the block's last point)
            BFState<D> &s = send[cur];
            if (j == i2 && pinned) {
                const double ie = 1.0 / b.pin_eps;
                for (int i = 0; i < D; ++i) { s.H[i * D + i] = ie; s.Psi[i * D + i] = ie; s.Qc[i * D + i] = ie; }
                s.c0 = 0.5 * D * log(2.0 * M_PI * b.pin_eps);
            } else {
                for (int i = 0; i < D * D; ++i) s.H[i] += b.Hobs[(size_t)j * D * D + i];
                for (int i = 0; i < D; ++i) s.F0[i] += b.Fobs[(size_t)j * D + i];
                s.c0 += b.cobs[j];
            }
            store_point<D>(b, s, base + N);
        }
        __syncthreads();
        const double tN = b.t[base + N];
        for (int k = tid; k < N; k += nt) {
            double Phi[D * D], Q[D * D], mvec[D];
            transition<D>(b.auxB + (size_t)j * D * D, b.auxbeta + (size_t)j * D, b.auxatil + (size_t)j * D * D,
                          tN - b.t[base + k], Phi, Q, mvec);
            BFState<D> s = send[cur];
            bf_step<D>(pinned, Phi, Q, mvec, s);
            store_point<D>(b, s, base + k);
            store_rec<D>(b, s, sbase + k, b.t[base + k + 1] - b.t[base + k]);
            if (k == 0) send[cur ^ 1] = s;
        }
        // pad slots of the interval's last tile: benign record (dt = 0)
        const int npad = ((N + TILE - 1) / TILE) * TILE;
        for (int e = N * b.rec + tid; e < npad * b.rec; e += nt) b.table[sbase * b.rec + e] = 0.0;
        __syncthreads();
        cur ^= 1;
    }
    if (tid == 0) {
        const BFState<D> &s = send[cur];
        b.blk_c0[blk] = s.c0;
        for (int i = 0; i < D; ++i) b.blk_g[(size_t)blk * D + i] = s.g[i];
        for (int i = 0; i < D * D; ++i) b.blk_Qc[(size_t)blk * D * D + i] = s.Qc[i];
    }
}

// ---------------------------------------------------------------------------------------------
// Large state dimension (Lorenz-96, D = 40): one CTA per layout block, matrices in shared memory,
// every matrix operation of the small-D recursion spread over the CTA's threads.  Each output
// element is still produced by ONE thread that sums in the same order as the scalar code (and the
// oracle), the LU pivot is chosen by the same first-maximum rule, so the result differs from the
// oracle only by FMA contraction.
// ---------------------------------------------------------------------------------------------
template <int D>
struct Coop {
    int tid, nt;
    __device__ void mm(const double *A, const double *B, double *C) const {   // C = A B
        for (int e = tid; e < D * D; e += nt) {
            const int i = e / D, j = e % D;
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += A[i * D + k] * B[k * D + j];
            C[e] = s;
        }
        __syncthreads();
    }
    __device__ void mmT(const double *A, const double *B, double *C) const {  // C = A B'
        for (int e = tid; e < D * D; e += nt) {
            const int i = e / D, j = e % D;
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += A[i * D + k] * B[j * D + k];
            C[e] = s;
        }
        __syncthreads();
    }
    __device__ void mTm(const double *A, const double *B, double *C) const {  // C = A' B
        for (int e = tid; e < D * D; e += nt) {
            const int i = e / D, j = e % D;
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += A[k * D + i] * B[k * D + j];
            C[e] = s;
        }
        __syncthreads();
    }
    __device__ void mv(const double *A, const double *x, double *y) const {
        for (int i = tid; i < D; i += nt) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += A[i * D + k] * x[k];
            y[i] = s;
        }
        __syncthreads();
    }
    __device__ void mTv(const double *A, const double *x, double *y) const {
        for (int i = tid; i < D; i += nt) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += A[k * D + i] * x[k];
            y[i] = s;
        }
        __syncthreads();
    }
    __device__ double dot(const double *a, const double *b) const {  // every thread sums in order
        double s = 0.0;
        for (int i = 0; i < D; ++i) s += a[i] * b[i];
        return s;
    }
    __device__ void sym(double *A) const {
        for (int e = tid; e < D * D; e += nt) {
            const int i = e / D, j = e % D;
            if (j > i) {
                const double s = 0.5 * (A[i * D + j] + A[j * D + i]);
                A[i * D + j] = s;
                A[j * D + i] = s;
            }
        }
        __syncthreads();
    }
    __device__ void copy(double *dst, const double *src, int n) const {
        for (int e = tid; e < n; e += nt) dst[e] = src[e];
        __syncthreads();
    }
    // LU with partial pivoting in place; returns log|det| (same value on every thread)
    __device__ double lu(double *A, int *piv) const {
        double lad = 0.0;
        for (int k = 0; k < D; ++k) {
            if (tid < 32) {  // first maximum of |A[i][k]|, i >= k (the oracle's rule), by one warp: strided scan + reduction
                int p = k;
                double best = -1.0;
                for (int i = k + tid; i < D; i += 32) {
                    const double v = fabs(A[i * D + k]);
                    if (v > best) { best = v; p = i; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double ob = __shfl_down_sync(0xffffffffu, best, off);
                    const int op = __shfl_down_sync(0xffffffffu, p, off);
                    if (ob > best || (ob == best && op < p)) { best = ob; p = op; }
                }
                if (tid == 0) piv[k] = p;
            }
            __syncthreads();
            const int p = piv[k];
            if (p != k) {
                for (int j = tid; j < D; j += nt) {
                    const double t = A[k * D + j]; A[k * D + j] = A[p * D + j]; A[p * D + j] = t;
                }
                __syncthreads();
            }
            const double pivv = A[k * D + k];
            lad += log(fabs(pivv));
            const double inv = 1.0 / pivv;
            __syncthreads();
            for (int i = k + 1 + tid; i < D; i += nt) A[i * D + k] = A[i * D + k] * inv;
            __syncthreads();
            const int m = D - k - 1;
            for (int e = tid; e < m * m; e += nt) {
                const int i = k + 1 + e / m, j = k + 1 + e % m;
                A[i * D + j] -= A[i * D + k] * A[k * D + j];
            }
            __syncthreads();
        }
        return lad;
    }
    // column j of A X = Bm (Bm row-major D x NR, overwritten) by one thread
    __device__ void solve_col(const double *LU, const int *piv, double *Bm, int NR, int j) const {
        for (int k = 0; k < D; ++k) {
            const int p = piv[k];
            if (p != k) { const double t = Bm[k * NR + j]; Bm[k * NR + j] = Bm[p * NR + j]; Bm[p * NR + j] = t; }
            const double bk = Bm[k * NR + j];
            for (int i = k + 1; i < D; ++i) Bm[i * NR + j] -= LU[i * D + k] * bk;
        }
        for (int k = D - 1; k >= 0; --k) {
            double sacc = Bm[k * NR + j];
            for (int i = k + 1; i < D; ++i) sacc -= LU[k * D + i] * Bm[i * NR + j];
            Bm[k * NR + j] = sacc * (1.0 / LU[k * D + k]);
        }
    }
    // solve A X = Bm for NR columns: one thread per column
    __device__ void solve(const double *LU, const int *piv, double *Bm, int NR) const {
        for (int j = tid; j < NR; j += nt) solve_col(LU, piv, Bm, NR, j);
        __syncthreads();
    }
    // the step's three right-hand sides (H: D columns, F: one, Psi: D columns when the block is pinned) in ONE pass, a
    // thread per column in separate warps -- each column's arithmetic is what solve() does, so the tables do not change;
    // one after the other the three passes were about half of a backward step (a column is 3200 serial multiply-subtracts)
    __device__ void solve3(const double *LU, const int *piv, double *Hm, double *Fv, double *Pm) const {
        if (tid < D) solve_col(LU, piv, Hm, D, tid);
        else if (tid == 64) solve_col(LU, piv, Fv, 1, 0);
        else if (Pm && tid >= 96 && tid < 96 + D) solve_col(LU, piv, Pm, D, tid - 96);
        __syncthreads();
    }
};

template <int D>
__global__ void __launch_bounds__(256) bf_big_kernel(const BFParams b) {
    extern __shared__ double smem[];
    constexpr int DD = D * D;
    double *H = smem, *Psi = H + DD, *Qc = Psi + DD, *Phi = Qc + DD, *Q = Phi + DD, *A = Q + DD;
    double *Hs = A + DD, *T1 = Hs + DD, *Psis = T1 + DD, *T2 = Psis + DD;
    double *F0 = T2 + DD, *g = F0 + D, *mvec = g + D, *Fs = mvec + D, *tv = Fs + D, *tv2 = tv + D;
    double *a1 = tv2 + D, *a2 = a1 + D, *bn = a2 + D;
    __shared__ int piv[D];
    __shared__ double c0s;
    const Coop<D> co{(int)threadIdx.x, (int)blockDim.x};
    const int tid = threadIdx.x, nt = blockDim.x;
    const int blk = blockIdx.x;
    for (int e = tid; e < DD; e += nt) { H[e] = 0.0; Psi[e] = 0.0; Qc[e] = 0.0; }
    for (int e = tid; e < D; e += nt) { F0[e] = 0.0; g[e] = 0.0; }
    if (tid == 0) c0s = 0.0;
    __syncthreads();
    const bool pinned = b.blk_pinned[blk] != 0;
    const int i1 = b.blk_i1[blk], i2 = b.blk_i2[blk];

    auto store = [&](size_t pt, long long step, double dt) {
        if (b.ptH) for (int e = tid; e < DD; e += nt) b.ptH[pt * DD + e] = H[e];
        if (b.ptF0) for (int e = tid; e < D; e += nt) b.ptF0[pt * D + e] = F0[e];
        if (b.ptPsi) for (int e = tid; e < DD; e += nt) b.ptPsi[pt * DD + e] = Psi[e];
        if (b.ptc0 && tid == 0) b.ptc0[pt] = c0s;
        if (step >= 0) {
            double *r = b.table + (size_t)step * b.rec;
            if (tid == 0) { r[0] = dt; r[1] = sqrt(dt); }
            for (int e = tid; e < DD; e += nt) r[2 + big_frag(e / D, e % D)] = H[e];  // fragment order (dmcmc_device.cuh)
            for (int e = tid; e < D; e += nt) r[2 + DD + e] = F0[e];
            if (b.has_psi)
                for (int e = tid; e < DD; e += nt) r[2 + DD + D + big_frag(e / D, e % D)] = Psi[(e % D) * D + (e / D)];  // transposed
        }
        __syncthreads();
    };

    for (int j = i2; j >= i1; --j) {
        if (j == i2 && pinned) {
            const double ie = 1.0 / b.pin_eps;
            for (int i = tid; i < D; i += nt) { H[i * D + i] = ie; Psi[i * D + i] = ie; Qc[i * D + i] = ie; }
            if (tid == 0) c0s = 0.5 * D * log(2.0 * M_PI * b.pin_eps);
        } else {
            for (int e = tid; e < DD; e += nt) H[e] += b.Hobs[(size_t)j * DD + e];
            for (int e = tid; e < D; e += nt) F0[e] += b.Fobs[(size_t)j * D + e];
            if (tid == 0) c0s += b.cobs[j];
        }
        __syncthreads();
        const int N = b.iv_N[j];
        const size_t base = (size_t)b.iv_pt_off[j];
        const long long sbase = (long long)b.iv_tile_off[j] * TILE;
        store(base + N, -1, 0.0);
        const double *Bj = b.auxB + (size_t)j * DD, *betaj = b.auxbeta + (size_t)j * D, *atj = b.auxatil + (size_t)j * DD;
        double dt_cached = -1.0;
        for (int k = N - 1; k >= 0; --k) {
            const double dt = b.t[base + k + 1] - b.t[base + k];
            if (!(fabs(dt - dt_cached) <= 1e-9 * fabs(dt_cached))) {
                // transition (Phi, Q, mvec): Taylor series with scaling-and-doubling (mirrors transition<D>)
                double nrm = 0.0;
                for (int i = 0; i < D; ++i) {
                    double sacc = 0.0;
                    for (int jj = 0; jj < D; ++jj) sacc += fabs(Bj[i * D + jj]);
                    if (sacc > nrm) nrm = sacc;
                }
                int sq = 0;
                double h = dt;
                while (nrm * h > 0.25 && sq < 40) { h *= 0.5; ++sq; }
                double *Pn = A, *An = Hs;  // scratch: A and Hs are free between steps
                for (int e = tid; e < DD; e += nt) {
                    const double idm = (e / D == e % D) ? 1.0 : 0.0;
                    Phi[e] = idm; Pn[e] = idm; An[e] = atj[e];
                }
                for (int e = tid; e < D; e += nt) bn[e] = betaj[e];
                __syncthreads();
                double coef = h;
                for (int e = tid; e < DD; e += nt) Q[e] = coef * An[e];
                for (int e = tid; e < D; e += nt) mvec[e] = coef * bn[e];
                __syncthreads();
                for (int n = 1; n <= TAYLOR_TERMS; ++n) {
                    co.mm(Bj, Pn, T1);
                    for (int e = tid; e < DD; e += nt) { Pn[e] = T1[e] * (h / n); Phi[e] += Pn[e]; }
                    __syncthreads();
                    co.mm(Bj, An, T1);
                    co.mmT(An, Bj, T2);
                    for (int e = tid; e < DD; e += nt) An[e] = T1[e] + T2[e];
                    co.mv(Bj, bn, tv);
                    for (int e = tid; e < D; e += nt) bn[e] = tv[e];
                    coef = coef * h / (n + 1);
                    __syncthreads();
                    for (int e = tid; e < DD; e += nt) Q[e] += coef * An[e];
                    for (int e = tid; e < D; e += nt) mvec[e] += coef * bn[e];
                    __syncthreads();
                }
                for (int s2 = 0; s2 < sq; ++s2) {
                    co.mm(Phi, Q, T1);
                    co.mmT(T1, Phi, T2);
                    for (int e = tid; e < DD; e += nt) Q[e] += T2[e];
                    co.mv(Phi, mvec, tv);
                    for (int e = tid; e < D; e += nt) mvec[e] += tv[e];
                    __syncthreads();
                    co.mm(Phi, Phi, T1);
                    co.copy(Phi, T1, DD);
                }
                co.sym(Q);
                dt_cached = dt;
            }
            // ---- one exact backward step (mirrors bf_step<D>) ----
            co.mm(H, Q, A);
            for (int i = tid; i < D; i += nt) A[i * D + i] += 1.0;
            __syncthreads();
            const double lad = co.lu(A, piv);
            co.copy(Hs, H, DD);
            co.copy(Fs, F0, D);
            if (pinned) co.copy(Psis, Psi, DD);
            co.solve3(A, piv, Hs, Fs, pinned ? Psis : nullptr);
            co.sym(Hs);
            co.mv(Hs, mvec, tv);
            co.mv(Q, Fs, tv2);
            if (tid == 0)
                c0s = c0s + 0.5 * lad + 0.5 * co.dot(mvec, tv) - co.dot(mvec, Fs) - 0.5 * co.dot(F0, tv2);
            __syncthreads();
            if (pinned) {
                co.mTv(Psis, mvec, a1);
                co.mTv(Psi, tv2, a2);
                for (int i = tid; i < D; i += nt) g[i] = g[i] - a1[i] - a2[i];
                co.mm(Q, Psis, T1);
                co.mTm(Psi, T1, A);
                for (int e = tid; e < DD; e += nt) Qc[e] -= A[e];
                __syncthreads();
                co.sym(Qc);
                co.mTm(Phi, Psis, Psi);
            }
            for (int i = tid; i < D; i += nt) tv2[i] = Fs[i] - tv[i];
            __syncthreads();
            co.mTv(Phi, tv2, F0);
            co.mm(Hs, Phi, T1);
            co.mTm(Phi, T1, H);
            co.sym(H);
            store(base + k, sbase + k, dt);
        }
        for (int k = N; k < ((N + TILE - 1) / TILE) * TILE; ++k) {  // pad slots: benign record
            double *r = b.table + (size_t)(sbase + k) * b.rec;
            for (int e = tid; e < b.rec; e += nt) r[e] = 0.0;
        }
        __syncthreads();
    }
    if (tid == 0) b.blk_c0[blk] = c0s;
    for (int e = tid; e < D; e += nt) b.blk_g[(size_t)blk * D + e] = g[e];
    for (int e = tid; e < DD; e += nt) b.blk_Qc[(size_t)blk * DD + e] = Qc[e];
}

// Built-in auxiliary law of every interval from theta and the interval's end observation, on the device: a critical
// parameter proposal (DD.set_parameters! + backward_filter!, src/run!_alterations.jl:188-204) no longer rebuilds and
// uploads J aux laws on the host (measured on C4, 10^4 intervals: 0.57 of the 0.68 ms parameter step).  Lives in this
// translation unit because it is compiled without FMA contraction: the aux law feeds the filter, and the CPU restatement
// builds it with separately rounded operations.
__global__ void aux_build_kernel(int model, int d, int m, int p, int J, Theta th, const double *__restrict__ v,
                                 double *__restrict__ B, double *__restrict__ beta, double *__restrict__ atil) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J) return;
    double *Bj = B + (size_t)j * d * d, *bj = beta + (size_t)j * d, *aj = atil + (size_t)j * d * d;
    if (d <= 3) {
        double sig[9];
        aux_build_host(model, d, m, th.v, v + (size_t)j * p, p, Bj, bj, sig);
        for (int i = 0; i < d; ++i)
            for (int k = 0; k < d; ++k) {
                double s = 0.0;
                for (int c = 0; c < m; ++c) s += sig[i * m + c] * sig[k * m + c];
                aj[i * d + k] = s;
            }
    } else {  // Lorenz-96: sigma~ = sigma I (the d x m scratch would not fit a thread's registers)
        for (int i = 0; i < d * d; ++i) { Bj[i] = 0.0; aj[i] = 0.0; }
        const double c = th.v[2];
        for (int i = 0; i < d; ++i) {
            Bj[i * d + i] += -1.0;
            Bj[i * d + (i + 1) % d] += c;
            Bj[i * d + (i + d - 2) % d] += -c;
            bj[i] = th.v[0];
            aj[i * d + i] = th.v[1] * th.v[1];
        }
    }
}

cudaError_t launch_aux_build(int model, int d, int m, int p, int J, const Theta &th, const double *v, double *B, double *beta,
                             double *atil, cudaStream_t s) {
    if (J <= 0) return cudaSuccess;
    aux_build_kernel<<<(unsigned)((J + 127) / 128), 128, 0, s>>>(model, d, m, p, J, th, v, B, beta, atil);
    return cudaGetLastError();
}

cudaError_t launch_backward_filter(const BFParams &b, cudaStream_t s) {
    if (b.nblk == 0) return cudaSuccess;
    if (b.D == 40) {
        constexpr int D = 40;
        const size_t smem = (size_t)(10 * D * D + 9 * D) * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(bf_big_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        bf_big_kernel<D><<<b.nblk, 256, smem, s>>>(b);
        return cudaGetLastError();
    }
    static const int env_serial = getenv("DMCMC_BF_SERIAL") ? atoi(getenv("DMCMC_BF_SERIAL")) : 0;  // A/B: step-by-step recursion
    if (env_serial) {
        const int threads = 32;
        const unsigned grid = (unsigned)((b.nblk + threads - 1) / threads);
        switch (b.D) {
        case 2: bf_kernel<2><<<grid, threads, 0, s>>>(b); break;
        case 3: bf_kernel<3><<<grid, threads, 0, s>>>(b); break;
        default: return cudaErrorInvalidValue;
        }
        return cudaGetLastError();
    }
    // one thread per grid point of an interval: 32 .. 128 threads by the mean interval length
    const int threads = b.mean_N <= 32 ? 32 : (b.mean_N <= 64 ? 64 : 128);
    switch (b.D) {
    case 2: bf_par_kernel<2><<<(unsigned)b.nblk, threads, 0, s>>>(b); break;
    case 3: bf_par_kernel<3><<<(unsigned)b.nblk, threads, 0, s>>>(b); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace dmcmc
