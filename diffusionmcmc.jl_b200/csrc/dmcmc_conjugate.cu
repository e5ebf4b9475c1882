// dmcmc_conjugate.cu -- path integrals of the conjugate Gaussian parameter update.
//
// Reference: compute_mu_and_W, /root/reference src/updates.jl:241-262.  For a drift whose rough
// (non-hypoelliptic) coordinates are linear in the parameters, b_rough(x) = phi(x) theta + phi_c(x),
//     mu += phi' Gamma^-1 (dy - phi_c dt),     W += phi' Gamma^-1 phi dt      over every Euler step
// of the accepted path, with dy the increment of the rough coordinates and Gamma = sigma sigma'
// restricted to them (diagonal for every law here).  The reference walks the path on the host; here
// the accepted containers stay in HBM and two small kernels reduce them: one thread per
// (chain, interval) walks its contiguous record, a second pass adds the intervals of a chain in order
// (deterministic, independent of the launch shape).  Reads the path once: 8 d bytes per step.
#include "dmcmc_device.cuh"
#include "dmcmc_internal.h"

namespace dmcmc {

// Conjugate structure of each law: NC parameters (theta indices IDX), NR rough coordinates (ROW).
struct ConjFHN {   // dX2 = (gamma X1 - X2 + beta) dt + sigma dW : theta = (gamma, beta)
    static constexpr int D = 2, NC = 2, NR = 1;
    __device__ static void idx(int *i) { i[0] = 2; i[1] = 3; }
    __device__ static void terms(const Theta &th, const double *x, int *row, double (*P)[NC], double *pc, double *g) {
        row[0] = 1;
        P[0][0] = x[0]; P[0][1] = 1.0;
        pc[0] = -x[1];
        g[0] = 1.0 / (th.v[4] * th.v[4]);
    }
};
template <bool MULT>
struct ConjLV {    // theta = (alpha, beta, gamma, delta)
    static constexpr int D = 2, NC = 4, NR = 2;
    __device__ static void idx(int *i) { i[0] = 0; i[1] = 1; i[2] = 2; i[3] = 3; }
    __device__ static void terms(const Theta &th, const double *x, int *row, double (*P)[NC], double *pc, double *g) {
        row[0] = 0; row[1] = 1;
        P[0][0] = x[0]; P[0][1] = -x[0] * x[1]; P[0][2] = 0.0; P[0][3] = 0.0;
        P[1][0] = 0.0; P[1][1] = 0.0; P[1][2] = -x[1]; P[1][3] = x[0] * x[1];
        pc[0] = 0.0; pc[1] = 0.0;
        const double s0 = MULT ? th.v[4] * x[0] : th.v[4], s1 = MULT ? th.v[5] * x[1] : th.v[5];
        g[0] = 1.0 / (s0 * s0); g[1] = 1.0 / (s1 * s1);
    }
};
struct ConjL63 {   // theta = (theta1, theta2, theta3)
    static constexpr int D = 3, NC = 3, NR = 3;
    __device__ static void idx(int *i) { i[0] = 0; i[1] = 1; i[2] = 2; }
    __device__ static void terms(const Theta &th, const double *x, int *row, double (*P)[NC], double *pc, double *g) {
        row[0] = 0; row[1] = 1; row[2] = 2;
        for (int r = 0; r < 3; ++r) for (int a = 0; a < 3; ++a) P[r][a] = 0.0;
        P[0][0] = x[1] - x[0]; P[1][1] = x[0]; P[2][2] = -x[2];
        pc[0] = 0.0; pc[1] = -x[1] - x[0] * x[2]; pc[2] = x[0] * x[1];
        g[0] = g[1] = g[2] = 1.0 / (th.v[3] * th.v[3]);
    }
};

// one thread per (interval, chain): sums over the interval's Euler steps -> partial[j][chain][NC + NC*NC]
template <class CJ>
__global__ void conj_partial_kernel(const ConjParams c) {
    constexpr int D = CJ::D, NC = CJ::NC, NR = CJ::NR, NOUT = NC + NC * NC;
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (long long)c.J * c.C) return;
    const int j = (int)(u / c.C), chain = (int)(u - (long long)j * c.C);
    const int N = c.iv_N[j];
    const double *X = c.X[c.selX[(size_t)j * c.C + chain]] + c.iv_xoff[j] + (size_t)chain * c.iv_xs[j];
    const double *t = c.t + c.iv_pt_off[j];
    double x[D], xn[D], acc[NOUT];
    for (int o = 0; o < NOUT; ++o) acc[o] = 0.0;
    if (c.iv_first[j]) {
        for (int i = 0; i < D; ++i) x[i] = c.x0[((size_t)c.iv_rec[j] * D + i) * c.C + chain];
    } else {
        const int jp = j - 1;
        const double *Xp = c.X[c.selX[(size_t)jp * c.C + chain]] + c.iv_xoff[jp] + (size_t)chain * c.iv_xs[jp] + (size_t)(c.iv_N[jp] - 1) * D;
        for (int i = 0; i < D; ++i) x[i] = Xp[i];
    }
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < D; ++i) xn[i] = X[(size_t)k * D + i];
        const double dt = t[k + 1] - t[k];
        int row[NR];
        double P[NR][NC], pc[NR], g[NR];
        CJ::terms(c.th, x, row, P, pc, g);
        for (int r = 0; r < NR; ++r) {
            const double res = (xn[row[r]] - x[row[r]]) - pc[r] * dt;
            for (int a = 0; a < NC; ++a) {
                const double pg = P[r][a] * g[r];
                acc[a] += pg * res;
                for (int b = 0; b < NC; ++b) acc[NC + a * NC + b] += pg * P[r][b] * dt;
            }
        }
        for (int i = 0; i < D; ++i) x[i] = xn[i];
    }
    double *out = c.partial + ((size_t)j * c.C + chain) * NOUT;
    for (int o = 0; o < NOUT; ++o) out[o] = acc[o];
}

// out[chain][o] = sum over intervals j (in order) of partial[j][chain][o]
__global__ void conj_reduce_kernel(const double *partial, double *out, int C, int J, int nout) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (long long)C * nout) return;
    const int chain = (int)(u / nout), o = (int)(u % nout);
    double s = 0.0;
    for (int j = 0; j < J; ++j) s += partial[((size_t)j * C + chain) * nout + o];
    out[(size_t)chain * nout + o] = s;
}

int conjugate_dim(int model, int32_t *idx) {
    switch (model) {
    case 0: if (idx) { idx[0] = 2; idx[1] = 3; } return 2;
    case 1: case 5: if (idx) for (int i = 0; i < 4; ++i) idx[i] = i; return 4;
    case 2: if (idx) for (int i = 0; i < 3; ++i) idx[i] = i; return 3;
    }
    return 0;
}

cudaError_t launch_conjugate(int model, const ConjParams &c, double *out, cudaStream_t s) {
    const long long n = (long long)c.J * c.C;
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 127) / 128);
    switch (model) {
    case 0: conj_partial_kernel<ConjFHN><<<grid, 128, 0, s>>>(c); break;
    case 1: conj_partial_kernel<ConjLV<false>><<<grid, 128, 0, s>>>(c); break;
    case 5: conj_partial_kernel<ConjLV<true>><<<grid, 128, 0, s>>>(c); break;
    case 2: conj_partial_kernel<ConjL63><<<grid, 128, 0, s>>>(c); break;
    default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int nc = conjugate_dim(model, nullptr), nout = nc + nc * nc;
    const long long m = (long long)c.C * nout;
    conj_reduce_kernel<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(c.partial, out, c.C, c.J, nout);
    return cudaGetLastError();
}

}  // namespace dmcmc
