// dmcmc_models.cuh -- target diffusion laws and their linear auxiliary laws (sm_100a device code
// plus the host-side auxiliary-law builders).
//
// The reference takes these from DiffusionDefinition.jl's `@load_diffusion` (not on disk; SURVEY
// section 8c): drift b(t,x;theta), diffusion coefficient sigma(t,x;theta), `constdiff`,
// `bound_satisfied`, and the auxiliary B(t), beta(t), sigma~ (call sites: forward_guide! at
// /root/reference src/run!_alterations.jl:18-21, GP.solve_and_ll! at :206-208).
//
// Every law is a struct of static __host__ __device__ functions over register arrays, so the
// solve kernels are fully unrolled per model (state dims 2..3; Lorenz-96 d=40 has its own kernel).
#pragma once
#include <cmath>
#include <cstdint>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

#define DMCMC_MAX_THETA 8

namespace dmcmc {

struct Theta {
    double v[DMCMC_MAX_THETA];
};

// ---------------------------------------------------------------------------------------------
// FitzHugh-Nagumo: dY = (Y - Y^3 - X + s)/eps dt ; dX = (gamma Y - X + beta) dt + sigma dW
// (docs/src/tutorials/conjugate_updates.md:9-12; theta order of simple_inference.md:30)
// ---------------------------------------------------------------------------------------------
struct ModelFHN {
    static constexpr int D = 2, M = 1, NTH = 5;
    static constexpr bool CONST_SIGMA = true;
    __host__ __device__ static __forceinline__ void drift(const Theta &th, const double *x, double *b) {
        b[0] = (x[0] - x[0] * x[0] * x[0] - x[1] + th.v[1]) * (1.0 / th.v[0]);
        b[1] = th.v[2] * x[0] - x[1] + th.v[3];
    }
    // out = sigma(x) * w          (w has M entries)
    __host__ __device__ static __forceinline__ void sigma_mul(const Theta &th, const double *, const double *w, double *out) {
        out[0] = 0.0;
        out[1] = th.v[4] * w[0];
    }
    // out = sigma sigma' r
    __host__ __device__ static __forceinline__ void a_mul(const Theta &th, const double *, const double *r, double *out) {
        out[0] = 0.0;
        out[1] = th.v[4] * th.v[4] * r[1];
    }
    // least-squares noise recovery: w = argmin |sigma w - res|
    __host__ __device__ static __forceinline__ void sigma_solve(const Theta &th, const double *, const double *res, double *w) {
        w[0] = res[1] * (1.0 / th.v[4]);  // loop-invariant reciprocal: no FP64 division (and its branches) per step
    }
    __host__ __device__ static __forceinline__ void a_full(const Theta &th, const double *, double *a) {
        a[0] = a[1] = a[2] = 0.0;
        a[3] = th.v[4] * th.v[4];
    }
    __host__ __device__ static __forceinline__ bool bound_ok(const double *) { return true; }
};

// Lotka-Volterra with additive noise: dX=(aX-bXY)dt+s1 dW1 ; dY=(dXY-cY)dt+s2 dW2, X,Y>0
template <bool MULT>
struct ModelLVT {
    static constexpr int D = 2, M = 2, NTH = 6;
    static constexpr bool CONST_SIGMA = !MULT;
    __host__ __device__ static __forceinline__ void drift(const Theta &th, const double *x, double *b) {
        b[0] = th.v[0] * x[0] - th.v[1] * x[0] * x[1];
        b[1] = th.v[3] * x[0] * x[1] - th.v[2] * x[1];
    }
    __host__ __device__ static __forceinline__ void sig(const Theta &th, const double *x, double *s) {
        s[0] = MULT ? th.v[4] * x[0] : th.v[4];
        s[1] = MULT ? th.v[5] * x[1] : th.v[5];
    }
    __host__ __device__ static __forceinline__ void sigma_mul(const Theta &th, const double *x, const double *w, double *out) {
        double s[2];
        sig(th, x, s);
        out[0] = s[0] * w[0];
        out[1] = s[1] * w[1];
    }
    __host__ __device__ static __forceinline__ void a_mul(const Theta &th, const double *x, const double *r, double *out) {
        double s[2];
        sig(th, x, s);
        out[0] = s[0] * s[0] * r[0];
        out[1] = s[1] * s[1] * r[1];
    }
    __host__ __device__ static __forceinline__ void sigma_solve(const Theta &th, const double *x, const double *res, double *w) {
        double s[2];
        sig(th, x, s);
        w[0] = res[0] * (1.0 / s[0]);
        w[1] = res[1] * (1.0 / s[1]);
    }
    __host__ __device__ static __forceinline__ void a_full(const Theta &th, const double *x, double *a) {
        double s[2];
        sig(th, x, s);
        a[0] = s[0] * s[0];
        a[1] = a[2] = 0.0;
        a[3] = s[1] * s[1];
    }
    __host__ __device__ static __forceinline__ bool bound_ok(const double *x) { return x[0] > 0.0 && x[1] > 0.0; }
};
using ModelLV = ModelLVT<false>;
using ModelLVM = ModelLVT<true>;

// Lorenz-63 with isotropic additive noise
struct ModelL63 {
    static constexpr int D = 3, M = 3, NTH = 4;
    static constexpr bool CONST_SIGMA = true;
    __host__ __device__ static __forceinline__ void drift(const Theta &th, const double *x, double *b) {
        b[0] = th.v[0] * (x[1] - x[0]);
        b[1] = th.v[1] * x[0] - x[1] - x[0] * x[2];
        b[2] = x[0] * x[1] - th.v[2] * x[2];
    }
    __host__ __device__ static __forceinline__ void sigma_mul(const Theta &th, const double *, const double *w, double *out) {
        for (int i = 0; i < 3; ++i) out[i] = th.v[3] * w[i];
    }
    __host__ __device__ static __forceinline__ void a_mul(const Theta &th, const double *, const double *r, double *out) {
        for (int i = 0; i < 3; ++i) out[i] = th.v[3] * th.v[3] * r[i];
    }
    __host__ __device__ static __forceinline__ void sigma_solve(const Theta &th, const double *, const double *res, double *w) {
        const double is = 1.0 / th.v[3];
        for (int i = 0; i < 3; ++i) w[i] = res[i] * is;
    }
    __host__ __device__ static __forceinline__ void a_full(const Theta &th, const double *, double *a) {
        for (int i = 0; i < 9; ++i) a[i] = 0.0;
        a[0] = a[4] = a[8] = th.v[3] * th.v[3];
    }
    __host__ __device__ static __forceinline__ bool bound_ok(const double *) { return true; }
};

// linear 2-d law (target == its own auxiliary law: analytic known-answer case)
struct ModelLIN2 {
    static constexpr int D = 2, M = 2, NTH = 8;
    static constexpr bool CONST_SIGMA = true;
    __host__ __device__ static __forceinline__ void drift(const Theta &th, const double *x, double *b) {
        b[0] = th.v[0] * x[0] + th.v[1] * x[1] + th.v[4];
        b[1] = th.v[2] * x[0] + th.v[3] * x[1] + th.v[5];
    }
    __host__ __device__ static __forceinline__ void sigma_mul(const Theta &th, const double *, const double *w, double *out) {
        out[0] = th.v[6] * w[0];
        out[1] = th.v[7] * w[1];
    }
    __host__ __device__ static __forceinline__ void a_mul(const Theta &th, const double *, const double *r, double *out) {
        out[0] = th.v[6] * th.v[6] * r[0];
        out[1] = th.v[7] * th.v[7] * r[1];
    }
    __host__ __device__ static __forceinline__ void sigma_solve(const Theta &th, const double *, const double *res, double *w) {
        w[0] = res[0] * (1.0 / th.v[6]);
        w[1] = res[1] * (1.0 / th.v[7]);
    }
    __host__ __device__ static __forceinline__ void a_full(const Theta &th, const double *, double *a) {
        a[0] = th.v[6] * th.v[6];
        a[1] = a[2] = 0.0;
        a[3] = th.v[7] * th.v[7];
    }
    __host__ __device__ static __forceinline__ bool bound_ok(const double *) { return true; }
};

// ---------------------------------------------------------------------------------------------
// Host-side builders of the auxiliary law of one interval from theta and the interval's end
// observation v (p entries).  FitzHughNagumoAux / LotkaVolterraAux follow upstream
// DiffusionDefinition.jl [UPSTREAM-RECALL]; Lorenz aux laws are this project's definition
// (DESIGN.md "auxiliary laws").  B is d x d row-major, sig is d x m row-major.
// ---------------------------------------------------------------------------------------------
inline int model_dims(int model, int d_in, int *d, int *m, int *nth) {
    switch (model) {
    case 0: *d = 2; *m = 1; *nth = 5; return 0;
    case 1: *d = 2; *m = 2; *nth = 6; return 0;
    case 2: *d = 3; *m = 3; *nth = 4; return 0;
    case 3: *d = d_in; *m = d_in; *nth = 3; return (d_in == 40) ? 0 : -1;
    case 4: *d = 2; *m = 2; *nth = 8; return 0;
    case 5: *d = 2; *m = 2; *nth = 6; return 0;
    }
    return -1;
}

__host__ __device__ inline void aux_build_host(int model, int d, int m, const double *th, const double *v, int p,
                                               double *B, double *beta, double *sig) {
    for (int i = 0; i < d * d; ++i) B[i] = 0.0;
    for (int i = 0; i < d; ++i) beta[i] = 0.0;
    for (int i = 0; i < d * m; ++i) sig[i] = 0.0;
    switch (model) {
    case 0: {
        const double vv = v[0], eps = th[0];
        B[0] = (1.0 - 3.0 * vv * vv) / eps;
        B[1] = -1.0 / eps;
        B[2] = th[2];
        B[3] = -1.0;
        beta[0] = (th[1] + 2.0 * vv * vv * vv) / eps;
        beta[1] = th[3];
        sig[1] = th[4];
        break;
    }
    case 1:
    case 5: {
        const double al = th[0], be = th[1], ga = th[2], de = th[3];
        B[1] = -be * ga / de;
        B[2] = al * de / be;
        beta[0] = ga * al / de;
        beta[1] = -al * ga / be;
        if (model == 1) {
            sig[0] = th[4];
            sig[3] = th[5];
        } else {
            sig[0] = th[4] * (ga / de);
            sig[3] = th[5] * (al / be);
        }
        break;
    }
    case 2: {
        const double x = v[0], y = (p > 1 ? v[1] : 0.0), z = (p > 2 ? v[2] : 0.0);
        const double vv[3] = {x, y, z};
        B[0] = -th[0]; B[1] = th[0]; B[2] = 0.0;
        B[3] = th[1] - z; B[4] = -1.0; B[5] = -x;
        B[6] = y; B[7] = x; B[8] = -th[2];
        Theta t{};
        for (int i = 0; i < 4; ++i) t.v[i] = th[i];
        double bv[3];
        ModelL63::drift(t, vv, bv);
        for (int i = 0; i < 3; ++i) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += B[i * 3 + k] * vv[k];
            beta[i] = bv[i] - s;
        }
        sig[0] = sig[4] = sig[8] = th[3];
        break;
    }
    case 3: {
        const double c = th[2];
        for (int i = 0; i < d; ++i) {
            B[i * d + i] += -1.0;
            B[i * d + (i + 1) % d] += c;
            B[i * d + (i + d - 2) % d] += -c;
            beta[i] = th[0];
            sig[i * m + i] = th[1];
        }
        break;
    }
    case 4:
        B[0] = th[0]; B[1] = th[1]; B[2] = th[2]; B[3] = th[3];
        beta[0] = th[4]; beta[1] = th[5];
        sig[0] = th[6]; sig[3] = th[7];
        break;
    }
}

}  // namespace dmcmc
