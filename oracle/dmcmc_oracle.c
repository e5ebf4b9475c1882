/*
 * dmcmc_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See dmcmc_oracle.h.
 *
 * PARITY STATUS: parity unpinned for the arithmetic (no reference golden vectors exist and
 * the arithmetic's upstream sources are not on disk); the formulas follow SURVEY.md 8a rows
 * a2-a9 / Appendix B, which restate how the reference calls GuidedProposals.jl:
 *   forward_guide! (pCN form)   src/run!_alterations.jl:18-21
 *   accept_reject!              src/run!_alterations.jl:22-24  (log_prior = log_q = 0, :28-30)
 *   register / swap             src/run!_alterations.jl:36-47, :57-75, :81-91
 *   GP.solve_and_ll! (fixed W)  src/run!_alterations.jl:206-208
 *   backward_filter!            src/run!_alterations.jl:203
 *   init_ll! / loglikhd         src/workspaces.jl:425-431
 *   block start points          src/workspaces.jl:267-273, :433-446
 *   path gluing                 src/path_saving_buffer.jl:46-64
 *
 * Compile with -ffp-contract=off so the evaluation order below is the arithmetic.
 */
#include "dmcmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MD ORC_MAXD

/* ------------------------------------------------------------------------------------------
 * Models (DiffusionDefinition.jl is not on disk; equations from
 * docs/src/tutorials/conjugate_updates.md:9-12 and SURVEY 8c/8d).
 * ---------------------------------------------------------------------------------------- */
int orc_model_dims(int model, int d_in, int *d, int *m, int *nth) {
    switch (model) {
    case ORC_MODEL_FHN: *d = 2; *m = 1; *nth = 5; return 0;
    case ORC_MODEL_LV: *d = 2; *m = 2; *nth = 6; return 0;
    case ORC_MODEL_L63: *d = 3; *m = 3; *nth = 4; return 0;
    case ORC_MODEL_L96: *d = d_in; *m = d_in; *nth = 3; return (d_in >= 4 && d_in <= MD) ? 0 : -1;
    case ORC_MODEL_LIN2: *d = 2; *m = 2; *nth = 8; return 0;
    case ORC_MODEL_LVM: *d = 2; *m = 2; *nth = 6; return 0;
    }
    return -1;
}

void orc_drift(int model, int d, const double *th, double t, const double *x, double *b) {
    (void)t;
    switch (model) {
    case ORC_MODEL_FHN: /* dY=(Y-Y^3-X+s)/eps dt ; dX=(gamma Y - X + beta)dt + sigma dW */
        b[0] = (x[0] - x[0] * x[0] * x[0] - x[1] + th[1]) / th[0];
        b[1] = th[2] * x[0] - x[1] + th[3];
        break;
    case ORC_MODEL_LV:
    case ORC_MODEL_LVM:
        b[0] = th[0] * x[0] - th[1] * x[0] * x[1];
        b[1] = th[3] * x[0] * x[1] - th[2] * x[1];
        break;
    case ORC_MODEL_L63:
        b[0] = th[0] * (x[1] - x[0]);
        b[1] = th[1] * x[0] - x[1] - x[0] * x[2];
        b[2] = x[0] * x[1] - th[2] * x[2];
        break;
    case ORC_MODEL_L96:
        for (int i = 0; i < d; ++i) {
            int ip = (i + 1) % d, im1 = (i + d - 1) % d, im2 = (i + d - 2) % d;
            b[i] = (x[ip] - x[im2]) * x[im1] - x[i] + th[0];
        }
        break;
    case ORC_MODEL_LIN2:
        b[0] = th[0] * x[0] + th[1] * x[1] + th[4];
        b[1] = th[2] * x[0] + th[3] * x[1] + th[5];
        break;
    }
}

/* sig is d x m row-major */
void orc_sigma(int model, int d, int m, const double *th, const double *x, double *sig) {
    memset(sig, 0, sizeof(double) * (size_t)d * m);
    switch (model) {
    case ORC_MODEL_FHN: sig[1] = th[4]; break; /* [0; sigma] */
    case ORC_MODEL_LV: sig[0] = th[4]; sig[3] = th[5]; break;
    case ORC_MODEL_LVM: sig[0] = th[4] * x[0]; sig[3] = th[5] * x[1]; break;
    case ORC_MODEL_L63: sig[0] = sig[4] = sig[8] = th[3]; break;
    case ORC_MODEL_L96: for (int i = 0; i < d; ++i) sig[i * m + i] = th[1]; break;
    case ORC_MODEL_LIN2: sig[0] = th[6]; sig[3] = th[7]; break;
    }
}

int orc_constdiff(int model) { return model != ORC_MODEL_LVM; }

int orc_bound_satisfied(int model, int d, const double *x) {
    (void)d;
    if (model == ORC_MODEL_LV || model == ORC_MODEL_LVM) return x[0] > 0.0 && x[1] > 0.0;
    return 1;
}

/* Auxiliary (linear) laws.  [UPSTREAM-RECALL] for FHN (FitzHughNagumoAux: linearisation of the
 * cubic at the interval's end observation v) and LV (linearisation at the fixed point); the L63
 * and L96 aux laws are OUR definitions.  v has p entries. */
void orc_aux_build(int model, int d, int m, const double *th, const double *v, int p, double *B,
                   double *beta, double *sigt) {
    memset(B, 0, sizeof(double) * (size_t)d * d);
    memset(beta, 0, sizeof(double) * (size_t)d);
    memset(sigt, 0, sizeof(double) * (size_t)d * m);
    switch (model) {
    case ORC_MODEL_FHN: {
        double vv = v[0], eps = th[0];
        B[0] = (1.0 - 3.0 * vv * vv) / eps; B[1] = -1.0 / eps;
        B[2] = th[2];                       B[3] = -1.0;
        beta[0] = (th[1] + 2.0 * vv * vv * vv) / eps;
        beta[1] = th[3];
        sigt[1] = th[4];
        break;
    }
    case ORC_MODEL_LV:
    case ORC_MODEL_LVM: {
        double al = th[0], be = th[1], ga = th[2], de = th[3];
        B[1] = -be * ga / de;
        B[2] = al * de / be;
        beta[0] = ga * al / de;
        beta[1] = -al * ga / be;
        if (model == ORC_MODEL_LV) { sigt[0] = th[4]; sigt[3] = th[5]; }
        else { sigt[0] = th[4] * (ga / de); sigt[3] = th[5] * (al / be); }
        break;
    }
    case ORC_MODEL_L63: {
        /* Jacobian at the (fully observed) end point v; beta = b(v) - J v */
        double x = v[0], y = (p > 1 ? v[1] : 0.0), z = (p > 2 ? v[2] : 0.0);
        double vv[3] = {x, y, z}, bv[3];
        B[0] = -th[0];     B[1] = th[0]; B[2] = 0.0;
        B[3] = th[1] - z;  B[4] = -1.0;  B[5] = -x;
        B[6] = y;          B[7] = x;     B[8] = -th[2];
        orc_drift(model, 3, th, 0.0, vv, bv);
        for (int i = 0; i < 3; ++i) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += B[i * 3 + k] * vv[k];
            beta[i] = bv[i] - s;
        }
        sigt[0] = sigt[4] = sigt[8] = th[3];
        break;
    }
    case ORC_MODEL_L96: {
        /* Jacobian at the uniform state c*1: -I + c (S_{+1} - S_{-2}); beta = b(c 1) - J c 1 = F */
        double c = th[2];
        for (int i = 0; i < d; ++i) {
            B[i * d + i] += -1.0;
            B[i * d + (i + 1) % d] += c;
            B[i * d + (i + d - 2) % d] += -c;
            beta[i] = th[0];
            sigt[i * m + i] = th[1];
        }
        break;
    }
    case ORC_MODEL_LIN2:
        B[0] = th[0]; B[1] = th[1]; B[2] = th[2]; B[3] = th[3];
        beta[0] = th[4]; beta[1] = th[5];
        sigt[0] = th[6]; sigt[3] = th[7];
        break;
    }
}

/* ------------------------------------------------------------------------------------------
 * small dense helpers (row-major, n <= MD)
 * ---------------------------------------------------------------------------------------- */
static void mm(int n, const double *A, const double *B, double *C) { /* C = A B */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += A[i * n + k] * B[k * n + j];
            C[i * n + j] = s;
        }
}
static void mmT(int n, const double *A, const double *B, double *C) { /* C = A B' */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += A[i * n + k] * B[j * n + k];
            C[i * n + j] = s;
        }
}
static void mTm(int n, const double *A, const double *B, double *C) { /* C = A' B */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += A[k * n + i] * B[k * n + j];
            C[i * n + j] = s;
        }
}
static void mv(int n, const double *A, const double *x, double *y) {
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += A[i * n + k] * x[k];
        y[i] = s;
    }
}
static void mTv(int n, const double *A, const double *x, double *y) {
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += A[k * n + i] * x[k];
        y[i] = s;
    }
}
static double dotv(int n, const double *a, const double *b) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}
static void symmetrize(int n, double *A) {
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            double s = 0.5 * (A[i * n + j] + A[j * n + i]);
            A[i * n + j] = s;
            A[j * n + i] = s;
        }
}
/* LU with partial pivoting in place; returns log|det| and sign; piv[n] */
static int lu_factor(int n, double *A, int *piv, double *logabsdet) {
    double lad = 0.0;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (fabs(A[i * n + k]) > best) { best = fabs(A[i * n + k]); p = i; }
        piv[k] = p;
        if (best == 0.0) return -1;
        if (p != k)
            for (int j = 0; j < n; ++j) {
                double tmp = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = tmp;
            }
        lad += log(fabs(A[k * n + k]));
        double inv = 1.0 / A[k * n + k];
        for (int i = k + 1; i < n; ++i) {
            double l = A[i * n + k] * inv;
            A[i * n + k] = l;
            for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
        }
    }
    *logabsdet = lad;
    return 0;
}
/* solve A X = Bm for nrhs columns (Bm is n x nrhs row-major, overwritten) */
static void lu_solve(int n, const double *LU, const int *piv, double *Bm, int nrhs) {
    for (int k = 0; k < n; ++k) {
        int p = piv[k];
        if (p != k)
            for (int j = 0; j < nrhs; ++j) {
                double tmp = Bm[k * nrhs + j]; Bm[k * nrhs + j] = Bm[p * nrhs + j]; Bm[p * nrhs + j] = tmp;
            }
        for (int i = k + 1; i < n; ++i) {
            double l = LU[i * n + k];
            for (int j = 0; j < nrhs; ++j) Bm[i * nrhs + j] -= l * Bm[k * nrhs + j];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        double inv = 1.0 / LU[k * n + k];
        for (int j = 0; j < nrhs; ++j) {
            double s = Bm[k * nrhs + j];
            for (int i = k + 1; i < n; ++i) s -= LU[k * n + i] * Bm[i * nrhs + j];
            Bm[k * nrhs + j] = s * inv;
        }
    }
}

/* Observation at the end of an interval: v ~ N(L x, Sigma).  Contribution to (H,F,c):
 * H += L' S^-1 L ; F += L' S^-1 v ; c += 1/2 v' S^-1 v + 1/2 log det(2 pi S)   (SURVEY a8). */
int orc_obs_terms(int d, int p, const double *L, const double *Sigma, const double *v,
                  double *Hobs, double *Fobs, double *cobs) {
    double S[MD * MD], SiL[MD * MD], Siv[MD];
    int piv[MD];
    double lad;
    if (p > MD || d > MD) return -1;
    memcpy(S, Sigma, sizeof(double) * (size_t)p * p);
    if (lu_factor(p, S, piv, &lad)) return -2;
    /* SiL = S^-1 L  (p x d) */
    memcpy(SiL, L, sizeof(double) * (size_t)p * d);
    lu_solve(p, S, piv, SiL, d);
    memcpy(Siv, v, sizeof(double) * (size_t)p);
    lu_solve(p, S, piv, Siv, 1);
    for (int i = 0; i < d; ++i) {
        for (int j = 0; j < d; ++j) {
            double s = 0.0;
            for (int k = 0; k < p; ++k) s += L[k * d + i] * SiL[k * d + j];
            Hobs[i * d + j] = s;
        }
        double s = 0.0;
        for (int k = 0; k < p; ++k) s += L[k * d + i] * Siv[k];
        Fobs[i] = s;
    }
    symmetrize(d, Hobs);
    *cobs = 0.5 * dotv(p, v, Siv) + 0.5 * (p * log(2.0 * M_PI) + lad);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a8  Backward filter.
 * The continuous-time ODEs
 *     H' = -B'H - HB + H a~ H ,  F' = -B'F + H a~ F + H beta ,  c' = beta'F + 1/2 F' a~ F - 1/2 tr(H a~)
 * with B, beta, a~ constant on each interval have the exact one-step solution
 *     x(t+dt) | x(t) ~ N(Phi x + m, Q),   Phi = e^{B dt}, m = int e^{Bs} beta ds, Q = int e^{Bs} a~ e^{B's} ds
 *     K = (I + H+ Q)^-1 ; H* = K H+ ; F* = K F+ ;
 *     H = Phi' H* Phi ; F = Phi'(F* - H* m) ;
 *     c = c+ + 1/2 log det(I + H+ Q) + 1/2 m'H* m - m'F* - 1/2 F+' Q F*
 * which we use as the integrator (upstream integrates the same ODEs with OrdinaryDiffEq Tsit5 and
 * saves at the grid; this recursion is the exact flow of those ODEs, see DESIGN.md "deviations").
 * Pinned blocks (OUR definition of the blocking TODOs, src/workspaces.jl:469, src/schedule.jl:51-60):
 * the right end is an exact observation L=I, Sigma=eps I, v = x_b of the *current path*.  Since
 * F and c depend on x_b affinely / quadratically, we tabulate F = F0 + Psi x_b and
 * c = c0 + g'x_b + 1/2 x_b' Qc x_b so the tables are shared by every chain.
 * ---------------------------------------------------------------------------------------- */
#define ORC_TAYLOR_TERMS 16

void orc_transition(int d, const double *B, const double *beta, const double *atil, double dt,
                    double *Phi, double *Q, double *mvec) {
    /* scaling: halve dt until ||B dt||_inf <= 0.25, then double back */
    double nrm = 0.0;
    for (int i = 0; i < d; ++i) {
        double s = 0.0;
        for (int j = 0; j < d; ++j) s += fabs(B[i * d + j]);
        if (s > nrm) nrm = s;
    }
    int sq = 0;
    double h = dt;
    while (nrm * h > 0.25 && sq < 40) { h *= 0.5; ++sq; }

    double *Pn = (double *)malloc(sizeof(double) * 4 * MD * MD + sizeof(double) * 2 * MD);
    double *An = Pn + MD * MD, *T1 = An + MD * MD, *T2 = T1 + MD * MD;
    double *bn = T2 + MD * MD, *tv = bn + MD;
    /* n = 0 terms */
    for (int i = 0; i < d * d; ++i) { Phi[i] = 0.0; Pn[i] = 0.0; }
    for (int i = 0; i < d; ++i) { Phi[i * d + i] = 1.0; Pn[i * d + i] = 1.0; }
    memcpy(An, atil, sizeof(double) * (size_t)d * d);
    memcpy(bn, beta, sizeof(double) * (size_t)d);
    double coef = h; /* h^{n+1}/(n+1)! */
    for (int i = 0; i < d * d; ++i) Q[i] = coef * An[i];
    for (int i = 0; i < d; ++i) mvec[i] = coef * bn[i];
    for (int n = 1; n <= ORC_TAYLOR_TERMS; ++n) {
        /* Pn <- B Pn * h/n  (= (hB)^n/n!) */
        mm(d, B, Pn, T1);
        for (int i = 0; i < d * d; ++i) { Pn[i] = T1[i] * (h / n); Phi[i] += Pn[i]; }
        /* An <- B An + An B' ; bn <- B bn */
        mm(d, B, An, T1);
        mmT(d, An, B, T2);
        for (int i = 0; i < d * d; ++i) An[i] = T1[i] + T2[i];
        mv(d, B, bn, tv);
        memcpy(bn, tv, sizeof(double) * (size_t)d);
        coef = coef * h / (n + 1);
        for (int i = 0; i < d * d; ++i) Q[i] += coef * An[i];
        for (int i = 0; i < d; ++i) mvec[i] += coef * bn[i];
    }
    for (int s = 0; s < sq; ++s) { /* doubling: Q2 = Phi Q Phi' + Q ; m2 = Phi m + m ; Phi2 = Phi Phi */
        mm(d, Phi, Q, T1);
        mmT(d, T1, Phi, T2);
        for (int i = 0; i < d * d; ++i) Q[i] += T2[i];
        mv(d, Phi, mvec, tv);
        for (int i = 0; i < d; ++i) mvec[i] += tv[i];
        mm(d, Phi, Phi, T1);
        memcpy(Phi, T1, sizeof(double) * (size_t)d * d);
    }
    symmetrize(d, Q);
    free(Pn);
}

typedef struct {
    double H[MD * MD], F0[MD], Psi[MD * MD], c0, g[MD], Qc[MD * MD];
} bf_state;

/* one exact backward step through a transition (Phi,Q,m); pinned => also carry Psi,g,Qc */
static void bf_step(int d, int pinned, const double *Phi, const double *Q, const double *mvec,
                    bf_state *s, double *work) {
    double *A = work, *Hs = A + MD * MD, *T1 = Hs + MD * MD, *Psis = T1 + MD * MD;
    double Fs[MD], tv[MD], tv2[MD];
    int piv[MD];
    double lad;
    /* A = I + H+ Q */
    mm(d, s->H, Q, A);
    for (int i = 0; i < d; ++i) A[i * d + i] += 1.0;
    lu_factor(d, A, piv, &lad);
    /* Hs = K H+ ; Fs = K F0+ ; Psis = K Psi+ */
    memcpy(Hs, s->H, sizeof(double) * (size_t)d * d);
    lu_solve(d, A, piv, Hs, d);
    symmetrize(d, Hs);
    memcpy(Fs, s->F0, sizeof(double) * (size_t)d);
    lu_solve(d, A, piv, Fs, 1);
    /* c0 = c0+ + 1/2 logdet + 1/2 m'Hs m - m'Fs - 1/2 F0+' Q Fs */
    mv(d, Hs, mvec, tv);
    mv(d, Q, Fs, tv2);
    s->c0 = s->c0 + 0.5 * lad + 0.5 * dotv(d, mvec, tv) - dotv(d, mvec, Fs) - 0.5 * dotv(d, s->F0, tv2);
    if (pinned) {
        memcpy(Psis, s->Psi, sizeof(double) * (size_t)d * d);
        lu_solve(d, A, piv, Psis, d);
        /* g = g+ - Psis' m - Psi+' Q Fs ; Qc = Qc+ - Psi+' Q Psis */
        double a1[MD], a2[MD];
        mTv(d, Psis, mvec, a1);
        mTv(d, s->Psi, tv2, a2);
        for (int i = 0; i < d; ++i) s->g[i] = s->g[i] - a1[i] - a2[i];
        mm(d, Q, Psis, T1);
        mTm(d, s->Psi, T1, A); /* A reused */
        for (int i = 0; i < d * d; ++i) s->Qc[i] -= A[i];
        symmetrize(d, s->Qc);
        /* Psi = Phi' Psis */
        mTm(d, Phi, Psis, s->Psi);
    }
    /* F0 = Phi'(Fs - Hs m) */
    for (int i = 0; i < d; ++i) tv2[i] = Fs[i] - tv[i];
    mTv(d, Phi, tv2, s->F0);
    /* H = Phi' Hs Phi */
    mm(d, Hs, Phi, T1);
    mTm(d, Phi, T1, s->H);
    symmetrize(d, s->H);
}

/* scheme 0: step by step, grid point k from k+1 by the one-step transition (dt_k);
 * scheme 1: every grid point of an interval directly from the interval's right-end value by ONE
 *           transition over tau_k = t_N - t_k (the aux law is constant on the interval, so this is
 *           the same exact flow; it is what the device kernel bf_par_kernel computes for d <= 3,
 *           where the N grid points of an interval are independent and run in parallel).
 * Both are exact solutions of the H, F, c ODEs; they differ by rounding only
 * (tests/test_oracle.py::test_backward_filter_schemes_agree). */
void orc_backward_filter_scheme(int d, int J, const int32_t *N, const int32_t *pt_off, const double *t,
                                const double *auxB, const double *auxbeta, const double *auxatil,
                                const double *Hobs, const double *Fobs, const double *cobs, int nblk,
                                const int32_t *blk_i1, const int32_t *blk_i2, const int32_t *blk_pinned,
                                double pin_eps, int scheme, double *H, double *F0, double *Psi, double *c0,
                                double *blk_g, double *blk_Qc) {
    (void)J;
    for (int b = 0; b < nblk; ++b) {
        bf_state *s = (bf_state *)calloc(1, sizeof(bf_state));
        bf_state *send = (bf_state *)calloc(1, sizeof(bf_state));
        double *work = (double *)malloc(sizeof(double) * 4 * MD * MD);
        double *Phi = (double *)malloc(sizeof(double) * 2 * MD * MD + sizeof(double) * MD);
        double *Q = Phi + MD * MD, *mvec = Q + MD * MD;
        int pinned = blk_pinned[b];
        for (int j = blk_i2[b]; j >= blk_i1[b]; --j) {
            /* observation update at the interval's right end (zero state beyond the block) */
            if (j == blk_i2[b] && pinned) {
                double ie = 1.0 / pin_eps;
                for (int i = 0; i < d; ++i) { s->H[i * d + i] = ie; s->Psi[i * d + i] = ie; s->Qc[i * d + i] = ie; }
                s->c0 = 0.5 * d * log(2.0 * M_PI * pin_eps);
            } else {
                for (int i = 0; i < d * d; ++i) s->H[i] += Hobs[(size_t)j * d * d + i];
                for (int i = 0; i < d; ++i) s->F0[i] += Fobs[(size_t)j * d + i];
                s->c0 += cobs[j];
            }
            int Nj = N[j];
            size_t base = (size_t)pt_off[j];
            /* store terminal values at k = N */
            memcpy(H + (base + Nj) * d * d, s->H, sizeof(double) * (size_t)d * d);
            memcpy(F0 + (base + Nj) * d, s->F0, sizeof(double) * (size_t)d);
            if (Psi) memcpy(Psi + (base + Nj) * d * d, s->Psi, sizeof(double) * (size_t)d * d);
            c0[base + Nj] = s->c0;
            if (scheme == 1) {
                *send = *s;
                double tN = t[base + Nj];
                for (int k = Nj - 1; k >= 0; --k) {
                    orc_transition(d, auxB + (size_t)j * d * d, auxbeta + (size_t)j * d,
                                   auxatil + (size_t)j * d * d, tN - t[base + k], Phi, Q, mvec);
                    *s = *send;
                    bf_step(d, pinned, Phi, Q, mvec, s, work);
                    memcpy(H + (base + k) * d * d, s->H, sizeof(double) * (size_t)d * d);
                    memcpy(F0 + (base + k) * d, s->F0, sizeof(double) * (size_t)d);
                    if (Psi) memcpy(Psi + (base + k) * d * d, s->Psi, sizeof(double) * (size_t)d * d);
                    c0[base + k] = s->c0;
                }
                continue; /* s holds the value at k = 0: the next interval (to the left) continues from it */
            }
            double dt_cached = -1.0;
            for (int k = Nj - 1; k >= 0; --k) {
                double dt = t[base + k + 1] - t[base + k];
                /* grids are uniform inside an interval up to rounding: reuse the transition
                 * unless dt moved by more than 1e-9 relative (deterministic rule, same on device): on a uniform grid
                 * t_k = t_0 + k dt the differences t_{k+1} - t_k carry rounding noise of ulp(t)/dt, 2e-12 at t = 10 */
                if (!(fabs(dt - dt_cached) <= 1e-9 * fabs(dt_cached))) {
                    orc_transition(d, auxB + (size_t)j * d * d, auxbeta + (size_t)j * d,
                                   auxatil + (size_t)j * d * d, dt, Phi, Q, mvec);
                    dt_cached = dt;
                }
                bf_step(d, pinned, Phi, Q, mvec, s, work);
                memcpy(H + (base + k) * d * d, s->H, sizeof(double) * (size_t)d * d);
                memcpy(F0 + (base + k) * d, s->F0, sizeof(double) * (size_t)d);
                if (Psi) memcpy(Psi + (base + k) * d * d, s->Psi, sizeof(double) * (size_t)d * d);
                c0[base + k] = s->c0;
            }
        }
        memcpy(blk_g + (size_t)b * d, s->g, sizeof(double) * (size_t)d);
        memcpy(blk_Qc + (size_t)b * d * d, s->Qc, sizeof(double) * (size_t)d * d);
        free(Phi);
        free(work);
        free(send);
        free(s);
    }
}

void orc_backward_filter(int d, int J, const int32_t *N, const int32_t *pt_off, const double *t,
                         const double *auxB, const double *auxbeta, const double *auxatil,
                         const double *Hobs, const double *Fobs, const double *cobs, int nblk,
                         const int32_t *blk_i1, const int32_t *blk_i2, const int32_t *blk_pinned,
                         double pin_eps, double *H, double *F0, double *Psi, double *c0,
                         double *blk_g, double *blk_Qc) {
    orc_backward_filter_scheme(d, J, N, pt_off, t, auxB, auxbeta, auxatil, Hobs, Fobs, cobs, nblk, blk_i1, blk_i2,
                               blk_pinned, pin_eps, 0, H, F0, Psi, c0, blk_g, blk_Qc);
}

/* ------------------------------------------------------------------------------------------
 * a2  pCN Wiener refresh  (GP rand!(Wnr, W deg) + crank_nicolson!, via forward_guide! at
 *     src/run!_alterations.jl:18-21)
 * ---------------------------------------------------------------------------------------- */
void orc_wiener_refresh(int N, int m, const double *t, const double *Z, double *Wp) {
    for (int c = 0; c < m; ++c) Wp[c] = 0.0;
    for (int k = 0; k < N; ++k) {
        double sq = sqrt(t[k + 1] - t[k]);
        for (int c = 0; c < m; ++c) Wp[(k + 1) * m + c] = Wp[k * m + c] + sq * Z[k * m + c];
    }
}

void orc_crank_nicolson(int N, int m, double rho, const double *Wacc, double *Wp) {
    double lam = sqrt(1.0 - rho * rho);
    for (int k = 0; k <= N; ++k)
        for (int c = 0; c < m; ++c) Wp[k * m + c] = lam * Wp[k * m + c] + rho * Wacc[k * m + c];
}

/* a3: F[k] for this chain = F0[k] + Psi[k] x_b  (x_b == NULL in un-pinned blocks) */
static void guiding_F(const orc_problem *p, size_t pt, const double *xb, double *F) {
    int d = p->d;
    const double *F0 = p->F0 + pt * d;
    if (!xb) { memcpy(F, F0, sizeof(double) * (size_t)d); return; }
    const double *Ps = p->Psi + pt * d * d;
    for (int i = 0; i < d; ++i) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += Ps[i * d + k] * xb[k];
        F[i] = F0[i] + s;
    }
}

/* Girsanov integrand and one guided Euler-Maruyama step at grid point pt of interval j.
 * Returns G*dt in *dll, new state in xn (without the noise term added: drift part only is
 * returned in xdrift so the inverse map can reuse it). */
static void guided_step_terms(const orc_problem *p, int j, size_t pt, const double *xb,
                              const double *x, double dt, double *drift_dt, double *sig,
                              double *dll) {
    int d = p->d, m = p->m;
    double F[MD], r[MD], b[MD], bt[MD], a[MD * MD], ar[MD];
    const double *H = p->H + pt * d * d;
    guiding_F(p, pt, xb, F);
    for (int i = 0; i < d; ++i) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += H[i * d + k] * x[k];
        r[i] = F[i] - s;
    }
    orc_drift(p->model, d, p->th, p->t[pt], x, b);
    const double *B = p->auxB + (size_t)j * d * d, *be = p->auxbeta + (size_t)j * d;
    for (int i = 0; i < d; ++i) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += B[i * d + k] * x[k];
        bt[i] = s + be[i];
    }
    orc_sigma(p->model, d, m, p->th, x, sig);
    for (int i = 0; i < d; ++i)
        for (int k = 0; k < d; ++k) {
            double s = 0.0;
            for (int c = 0; c < m; ++c) s += sig[i * m + c] * sig[k * m + c];
            a[i * d + k] = s;
        }
    double g = 0.0;
    for (int i = 0; i < d; ++i) g += (b[i] - bt[i]) * r[i];
    if (!orc_constdiff(p->model)) {
        /* + 1/2 tr((a - a~)(r r' - H)) */
        const double *st = p->auxsig + (size_t)j * d * m;
        double tr = 0.0;
        for (int i = 0; i < d; ++i)
            for (int k = 0; k < d; ++k) {
                double at = 0.0;
                for (int c = 0; c < m; ++c) at += st[i * m + c] * st[k * m + c];
                tr += (a[i * d + k] - at) * (r[k] * r[i] - H[k * d + i]);
            }
        g += 0.5 * tr;
    }
    *dll = g * dt;
    mv(d, a, r, ar);
    for (int i = 0; i < d; ++i) drift_dt[i] = (b[i] + ar[i]) * dt;
}

/* a4 + a5: GP.solve_and_ll!(X, W, P, y1)  (call sites src/run!_alterations.jl:18-21, :206-208).
 * W is the cumulative Wiener path of interval j (N_j+1 points), X gets N_j+1 points. */
int orc_solve_and_ll(const orc_problem *p, int j, const double *xb, const double *W,
                     const double *y1, double *X, double *ll_out) {
    int d = p->d, m = p->m, N = p->N[j];
    size_t base = (size_t)p->pt_off[j];
    double x[MD], dr[MD], sig[MD * MD];
    double ll = 0.0;
    memcpy(x, y1, sizeof(double) * (size_t)d);
    memcpy(X, x, sizeof(double) * (size_t)d);
    for (int k = 0; k < N; ++k) {
        double dt = p->t[base + k + 1] - p->t[base + k], dll;
        guided_step_terms(p, j, base + k, xb, x, dt, dr, sig, &dll);
        ll += dll;
        for (int i = 0; i < d; ++i) {
            double s = 0.0;
            for (int c = 0; c < m; ++c) s += sig[i * m + c] * (W[(k + 1) * m + c] - W[k * m + c]);
            x[i] = x[i] + dr[i] + s;
        }
        memcpy(X + (size_t)(k + 1) * d, x, sizeof(double) * (size_t)d);
        if (!orc_bound_satisfied(p->model, d, x)) { *ll_out = -INFINITY; return 0; }
    }
    *ll_out = ll;
    return 1;
}

/* log h~(t_a, x_a) = -c - 1/2 x'Hx + F'x at the block's first grid point (GP loglikhd_obs) */
double orc_loglik_obs(const orc_problem *p, int blk, const double *x0, const double *xb) {
    int d = p->d;
    size_t pt = (size_t)p->pt_off[p->blk_i1[blk]];
    double F[MD], Hx[MD];
    guiding_F(p, pt, xb, F);
    mv(d, p->H + pt * d * d, x0, Hx);
    double c = p->c0[pt];
    if (xb) {
        double Qx[MD];
        mv(d, p->blk_Qc + (size_t)blk * d * d, xb, Qx);
        c += dotv(d, p->blk_g + (size_t)blk * d, xb) + 0.5 * dotv(d, xb, Qx);
    }
    return -c - 0.5 * dotv(d, x0, Hx) + dotv(d, F, x0);
}

static inline double *Xbuf(const orc_problem *p, const orc_state *s, int buf, int chain, int j) {
    return s->X[buf] + ((size_t)chain * p->pt_off[p->J] + p->pt_off[j]) * p->d;
}
static inline double *Wbuf(const orc_problem *p, const orc_state *s, int buf, int chain, int j) {
    return s->W[buf] + ((size_t)chain * p->pt_off[p->J] + p->pt_off[j]) * p->m;
}

/* loglikhd(P, XX) of the accepted path of one block (src/workspaces.jl:429) */
double orc_loglikhd_block(const orc_problem *p, const orc_state *s, int chain, int blk) {
    int d = p->d, i1 = p->blk_i1[blk], i2 = p->blk_i2[blk];
    const uint8_t *selX = s->selX + (size_t)chain * p->J;
    const double *x0 = Xbuf(p, s, selX[i1], chain, i1);
    double xbv[MD];
    const double *xb = NULL;
    if (p->blk_pinned[blk]) {
        memcpy(xbv, Xbuf(p, s, selX[i2], chain, i2) + (size_t)p->N[i2] * d, sizeof(double) * (size_t)d);
        xb = xbv;
    }
    double ll = orc_loglik_obs(p, blk, x0, xb);
    for (int j = i1; j <= i2; ++j) {
        const double *X = Xbuf(p, s, selX[j], chain, j);
        size_t base = (size_t)p->pt_off[j];
        double l = 0.0, dr[MD], sig[MD * MD];
        for (int k = 0; k < p->N[j]; ++k) {
            double dt = p->t[base + k + 1] - p->t[base + k], dll;
            guided_step_terms(p, j, base + k, xb, X + (size_t)k * d, dt, dr, sig, &dll);
            l += dll;
        }
        ll += l;
    }
    return ll;
}

/* ------------------------------------------------------------------------------------------
 * Counter-based noise (production mode).  Philox4x32-10 (Salmon et al. 2011), pinned by the
 * Random123 known-answer vectors in tests/test_oracle.py.
 *   key = (seed_lo ^ stream, seed_hi)      stream: 0 = Wiener normals, 0x9E3779B9 = accept draw
 *   ctr = ((wcomp << 24) | quad, interval, chain, iter)
 * Each call gives four N(0,1) (two single-precision Box-Muller pairs): steps 4*quad .. 4*quad+3.
 * ---------------------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double u01_53(uint32_t hi, uint32_t lo) {
    uint64_t v = (((uint64_t)hi << 32) | lo) >> 11; /* 53 bits */
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

/* Single-precision Box-Muller written ONLY with IEEE fmaf/mul/add/sqrtf so that the CUDA kernel
 * (which uses __fmaf_rn/__fmul_rn/__fadd_rn/__fsqrt_rn in the same order) is bit-identical.
 * The normals carry float resolution (|err| ~ 1e-7), then widen exactly to double; the Euler
 * recursion itself is FP64.  Polynomials: log1p on [-1/3,1/3] (degree 7 in f^3 Q(f)),
 * sin(pi r)/r and cos(pi r) on |r| <= 1/4 (degree 3 in r^2), least-squares/Chebyshev fits. */
static inline void bm_pair_f32(uint32_t oa, uint32_t ob, float *z0, float *z1) {
    const float u1 = (float)(((oa >> 9) << 1) | 1u) * 0x1p-24f; /* (0,1), exact */
    const float t = (float)(ob >> 8) * 0x1p-23f;                /* 2*u2 in [0,2), exact */
    union { float f; uint32_t u; } cv;
    cv.f = u1;
    int e = (int)((cv.u >> 23) & 0xffu) - 127;
    cv.u = (cv.u & 0x7fffffu) | 0x3f800000u;
    float m = cv.f;
    if (m > 1.33333337f) { m = m * 0.5f; e += 1; }
    const float f = m - 1.0f;
    float q = -0.124298662f;
    q = fmaf(q, f, 0.134337738f);
    q = fmaf(q, f, -0.122872368f);
    q = fmaf(q, f, 0.141169786f);
    q = fmaf(q, f, -0.16673398f);
    q = fmaf(q, f, 0.200038359f);
    q = fmaf(q, f, -0.249999434f);
    q = fmaf(q, f, 0.333333194f);
    const float s = f * f, s3 = s * f;
    const float r = fmaf(q, s3, fmaf(-0.5f, s, f));
    const float lnu = fmaf((float)e, 0.693147182f, r);
    const float rad = sqrtf(-2.0f * lnu);
    const float qf = rintf(t * 2.0f);
    const int qi = (int)qf;
    const float r2 = fmaf(-0.5f, qf, t);
    const float ss = r2 * r2;
    float sp = fmaf(ss, -0.589076877f, 2.54976702f);
    sp = fmaf(sp, ss, -5.16770792f);
    sp = fmaf(sp, ss, 3.14159274f);
    const float sn = sp * r2;
    float cp = fmaf(ss, -1.30612838f, 4.05757761f);
    cp = fmaf(cp, ss, -4.93478823f);
    cp = fmaf(cp, ss, 1.0f);
    float sv, cv2;
    switch (qi & 3) {
    case 0: sv = sn; cv2 = cp; break;
    case 1: sv = cp; cv2 = -sn; break;
    case 2: sv = -sn; cv2 = -cp; break;
    default: sv = -cp; cv2 = sn; break;
    }
    *z0 = rad * cv2;
    *z1 = rad * sv;
}

/* One Philox call -> four N(0,1): Euler steps 4*quad .. 4*quad+3 of Wiener component wcomp. */
void orc_philox_normal4(uint64_t seed, uint32_t iter, uint32_t interval, uint32_t chain,
                        uint32_t wcomp, uint32_t quad, double *z) {
    uint32_t ctr[4] = {(wcomp << 24) | quad, interval, chain, iter};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
    float a, b, c, d;
    orc_philox4x32_10(ctr, key, o);
    bm_pair_f32(o[0], o[1], &a, &b);
    bm_pair_f32(o[2], o[3], &c, &d);
    z[0] = a; z[1] = b; z[2] = c; z[3] = d;
}

double orc_philox_exponential(uint64_t seed, uint32_t iter, uint32_t first_interval, uint32_t chain) {
    uint32_t ctr[4] = {0u, first_interval, chain, iter};
    uint32_t key[2] = {(uint32_t)seed ^ 0x9E3779B9u, (uint32_t)(seed >> 32)}, o[4];
    orc_philox4x32_10(ctr, key, o);
    return -log(u01_53(o[0], o[1]));
}

/* ------------------------------------------------------------------------------------------
 * a1 / a6 / a7  eMCMC.update!(::PathImputation)  src/run!_alterations.jl:5-26 and the
 * register/swap that follows (:36-47, :57-75, :81-91), for every chain.
 * Z is [C][S][m] (S = st_off[J]), E is [C][nblk]; NULL with use_philox.
 * ---------------------------------------------------------------------------------------- */
static void impute_unit(const orc_problem *p, orc_state *s, int c, int b, const double *Z,
                        const double *E, uint64_t seed, uint32_t iter, int use_philox,
                        int force_accept, double *out_llprop, double *out_ll, uint8_t *out_acc) {
    int d = p->d, m = p->m, J = p->J;
    int i1 = p->blk_i1[b], i2 = p->blk_i2[b];
    uint8_t *selX = s->selX + (size_t)c * J, *selW = s->selW + (size_t)c * J;
    size_t S = (size_t)p->st_off[J];
    double y1[MD], xbv[MD];
    const double *xb = NULL;
    /* start point: ws.XX[i][1].x[1]  (src/run!_alterations.jl:20) */
    memcpy(y1, Xbuf(p, s, selX[i1], c, i1), sizeof(double) * (size_t)d);
    if (p->blk_pinned[b]) {
        memcpy(xbv, Xbuf(p, s, selX[i2], c, i2) + (size_t)p->N[i2] * d, sizeof(double) * (size_t)d);
        xb = xbv;
    }
    double llp = orc_loglik_obs(p, b, y1, xb);
    int ok = 1;
    double *zbuf = NULL;
    for (int j = i1; j <= i2 && ok; ++j) {
        int N = p->N[j];
        double *Wp = Wbuf(p, s, 1 - selW[j], c, j);
        const double *Wa = Wbuf(p, s, selW[j], c, j);
        double *Xp = Xbuf(p, s, 1 - selX[j], c, j);
        const double *z;
        if (use_philox) {
            zbuf = (double *)realloc(zbuf, sizeof(double) * (size_t)(N + 1) * m);
            for (int w = 0; w < m; ++w)
                for (int q = 0; 4 * q < N; ++q) {
                    double z4[4];
                    orc_philox_normal4(seed, iter, (uint32_t)(j + p->interval_offset), (uint32_t)(c + p->chain_offset), (uint32_t)w, (uint32_t)q, z4);
                    for (int i = 0; i < 4 && 4 * q + i < N; ++i) zbuf[(size_t)(4 * q + i) * m + w] = z4[i];
                }
            z = zbuf;
        } else {
            z = Z + ((size_t)c * S + p->st_off[j]) * m;
        }
        orc_wiener_refresh(N, m, p->t + p->pt_off[j], z, Wp);
        orc_crank_nicolson(N, m, p->rho[j], Wa, Wp);
        double l;
        ok = orc_solve_and_ll(p, j, xb, Wp, y1, Xp, &l);
        if (!ok) { llp = -INFINITY; break; }
        llp += l;
        memcpy(y1, Xp + (size_t)N * d, sizeof(double) * (size_t)d);
        if (j == i2 && xb) memcpy(Xp + (size_t)N * d, xb, sizeof(double) * (size_t)d); /* pin */
    }
    free(zbuf);
    double *ll = s->ll + (size_t)c * p->nblk + b;
    double e = use_philox ? orc_philox_exponential(seed, iter, (uint32_t)(i1 + p->interval_offset), (uint32_t)(c + p->chain_offset))
                          : E[(size_t)c * p->nblk + b];
    /* accepted = rand(Exponential(1)) > -(ll deg - ll)   (eMCMC.accept_reject!) */
    int acc = force_accept ? ok : (e > -(llp - *ll));
    if (acc) { /* swap_paths! + ll = ll deg  (src/run!_alterations.jl:44, :71, :87-91) */
        for (int j = i1; j <= i2; ++j) { selX[j] ^= 1; selW[j] ^= 1; }
        *ll = llp;
    }
    size_t u = (size_t)c * p->nblk + b;
    out_llprop[u] = llp;
    out_ll[u] = *ll;
    out_acc[u] = (uint8_t)acc;
}

void orc_impute_sweep(const orc_problem *p, orc_state *s, const double *Z, const double *E,
                      uint64_t seed, uint32_t iter, int use_philox, int force_accept,
                      const uint8_t *unit_mask, double *out_llprop, double *out_ll,
                      uint8_t *out_acc, int nthreads) {
    long nunits = (long)p->C * p->nblk;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (long u = 0; u < nunits; ++u) {
        int c = (int)(u / p->nblk), b = (int)(u % p->nblk);
        if (unit_mask && !unit_mask[u]) continue;
        impute_unit(p, s, c, b, Z, E, seed, iter, use_philox, force_accept, out_llprop, out_ll, out_acc);
    }
}

/* ------------------------------------------------------------------------------------------
 * a9  set_proposal!(::MCMCParamUpdate)  src/run!_alterations.jl:177-211: proposal law p_prop
 * (theta deg, tables deg), solve with the FIXED accepted W into the proposal X containers.
 * The reference returns early on the first failed block (:209, a quirk that leaves later ll deg
 * stale); here a failed block reports success=0 and ll deg=-Inf and the others are still computed.
 * out_llprop [C][nblk], out_success [C].
 * ---------------------------------------------------------------------------------------- */
void orc_param_solve(const orc_problem *p, orc_state *s, double *out_llprop, uint8_t *out_success,
                     int nthreads) {
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int c = 0; c < p->C; ++c) {
        int d = p->d, J = p->J, all_ok = 1;
        const uint8_t *selX = s->selX + (size_t)c * J, *selW = s->selW + (size_t)c * J;
        for (int b = 0; b < p->nblk; ++b) {
            int i1 = p->blk_i1[b], i2 = p->blk_i2[b], ok = 1;
            double y1[MD], xbv[MD];
            const double *xb = NULL;
            memcpy(y1, Xbuf(p, s, selX[i1], c, i1), sizeof(double) * (size_t)d);
            if (p->blk_pinned[b]) {
                memcpy(xbv, Xbuf(p, s, selX[i2], c, i2) + (size_t)p->N[i2] * d, sizeof(double) * (size_t)d);
                xb = xbv;
            }
            double llp = orc_loglik_obs(p, b, y1, xb);
            for (int j = i1; j <= i2; ++j) {
                double *Xp = Xbuf(p, s, 1 - selX[j], c, j), l;
                ok = orc_solve_and_ll(p, j, xb, Wbuf(p, s, selW[j], c, j), y1, Xp, &l);
                if (!ok) { llp = -INFINITY; break; }
                llp += l;
                memcpy(y1, Xp + (size_t)p->N[j] * d, sizeof(double) * (size_t)d);
                if (j == i2 && xb) memcpy(Xp + (size_t)p->N[j] * d, xb, sizeof(double) * (size_t)d);
            }
            out_llprop[(size_t)c * p->nblk + b] = llp;
            all_ok &= ok;
        }
        out_success[c] = (uint8_t)all_ok;
    }
}

/* swap_laws_and_paths!  src/run!_alterations.jl:328-334 (X swapped, W kept) + ll = ll deg (:322).
 * The law swap (P deg <-> P) is the caller's: the proposal problem becomes the current one. */
void orc_param_commit(const orc_problem *p, orc_state *s, const uint8_t *accepted_per_chain,
                      const double *llprop) {
    for (int c = 0; c < p->C; ++c) {
        if (!accepted_per_chain[c]) continue;
        for (int j = 0; j < p->J; ++j) s->selX[(size_t)c * p->J + j] ^= 1;
        for (int b = 0; b < p->nblk; ++b) s->ll[(size_t)c * p->nblk + b] = llprop[(size_t)c * p->nblk + b];
    }
}

/* ------------------------------------------------------------------------------------------
 * Layout change (OUR definition of change_laws_for_blocking! / recompute_loglikhd!, both
 * undefined in the reference: src/workspaces.jl:469-470).  Under the new layout's law, recover
 * the Wiener path that reproduces the accepted X through the guided Euler scheme
 *      sigma dW_k = x_{k+1} - x_k - (b + a r)(t_k, x_k) dt_k      (least squares when m < d)
 * and recompute ll = loglikhd of the accepted path.  W is rewritten in the accepted container.
 * ---------------------------------------------------------------------------------------- */
void orc_relayout(const orc_problem *p, orc_state *s, int nthreads) {
    long nunits = (long)p->C * p->nblk;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (long u = 0; u < nunits; ++u) {
        int c = (int)(u / p->nblk), b = (int)(u % p->nblk);
        int d = p->d, m = p->m, i1 = p->blk_i1[b], i2 = p->blk_i2[b];
        const uint8_t *selX = s->selX + (size_t)c * p->J, *selW = s->selW + (size_t)c * p->J;
        double xbv[MD];
        const double *xb = NULL;
        if (p->blk_pinned[b]) {
            memcpy(xbv, Xbuf(p, s, selX[i2], c, i2) + (size_t)p->N[i2] * d, sizeof(double) * (size_t)d);
            xb = xbv;
        }
        double ll = orc_loglik_obs(p, b, Xbuf(p, s, selX[i1], c, i1), xb);
        for (int j = i1; j <= i2; ++j) {
            const double *X = Xbuf(p, s, selX[j], c, j);
            double *W = Wbuf(p, s, selW[j], c, j);
            size_t base = (size_t)p->pt_off[j];
            double l = 0.0;
            for (int w = 0; w < m; ++w) W[w] = 0.0;
            for (int k = 0; k < p->N[j]; ++k) {
                double dt = p->t[base + k + 1] - p->t[base + k], dll, dr[MD], sig[MD * MD];
                double res[MD], StS[MD * MD], Str[MD];
                int piv[MD];
                double lad;
                guided_step_terms(p, j, base + k, xb, X + (size_t)k * d, dt, dr, sig, &dll);
                l += dll;
                for (int i = 0; i < d; ++i) res[i] = X[(size_t)(k + 1) * d + i] - X[(size_t)k * d + i] - dr[i];
                /* normal equations (sig' sig) dW = sig' res */
                for (int a = 0; a < m; ++a) {
                    for (int e = 0; e < m; ++e) {
                        double q = 0.0;
                        for (int i = 0; i < d; ++i) q += sig[i * m + a] * sig[i * m + e];
                        StS[a * m + e] = q;
                    }
                    double q = 0.0;
                    for (int i = 0; i < d; ++i) q += sig[i * m + a] * res[i];
                    Str[a] = q;
                }
                lu_factor(m, StS, piv, &lad);
                lu_solve(m, StS, piv, Str, 1);
                for (int w = 0; w < m; ++w) W[(size_t)(k + 1) * m + w] = W[(size_t)k * m + w] + Str[w];
            }
            ll += l;
        }
        s->ll[(size_t)c * p->nblk + b] = ll;
    }
}

/* init_ll!  src/workspaces.jl:425-431 */
void orc_init_ll(const orc_problem *p, orc_state *s, int nthreads) {
    long nunits = (long)p->C * p->nblk;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (long u = 0; u < nunits; ++u) {
        int c = (int)(u / p->nblk), b = (int)(u % p->nblk);
        s->ll[u] = orc_loglikhd_block(p, s, c, b);
    }
}

/* _copy_path!  src/path_saving_buffer.jl:54-64: every interval's x[1:N-1 (all but last)] glued,
 * then the final end point.  out has (sum_j N_j) + 1 points of d doubles. */
void orc_glue_path(const orc_problem *p, const orc_state *s, int chain, int rec_i1, int rec_i2,
                   double *out) {
    int d = p->d;
    size_t off = 0;
    const uint8_t *selX = s->selX + (size_t)chain * p->J;
    for (int j = rec_i1; j <= rec_i2; ++j) {
        const double *X = Xbuf(p, s, selX[j], chain, j);
        memcpy(out + off * d, X, sizeof(double) * (size_t)p->N[j] * d);
        off += (size_t)p->N[j];
    }
    const double *Xl = Xbuf(p, s, selX[rec_i2], chain, rec_i2);
    memcpy(out + off * d, Xl + (size_t)p->N[rec_i2] * d, sizeof(double) * (size_t)d);
}

/* ------------------------------------------------------------------------------------------
 * Block layouts.  No-blocking layout: init_block_layout(data, nothing), src/schedule.jl:97-100
 * = one block (rec, 1:n_obs) per recording.  Chequerboard layout (OUR definition; the reference
 * has only the decorator struct BlockingChequerboardSwitch(block_half_len), src/blocking.jl:7-9):
 *   pattern 0: delimiters at multiples of 2h  -> [0,2h) [2h,4h) ...
 *   pattern 1: delimiters shifted by h        -> [0,h) [h,3h) [3h,5h) ...
 * A block whose right end is not the recording's last interval is pinned.  Ranges are 0-based
 * inclusive global interval indices.  Returns the number of blocks (out arrays sized >= J).
 * ---------------------------------------------------------------------------------------- */
int orc_chequerboard_layout(int R, const int32_t *rec_nobs, int half_len, int pattern,
                            int32_t *blk_i1, int32_t *blk_i2, int32_t *blk_pinned) {
    int nb = 0, base = 0;
    for (int r = 0; r < R; ++r) {
        int n = rec_nobs[r];
        if (half_len <= 0) {
            blk_i1[nb] = base; blk_i2[nb] = base + n - 1; blk_pinned[nb] = 0; ++nb;
        } else {
            int start = 0;
            int first_len = pattern ? half_len : 2 * half_len;
            int len = first_len;
            while (start < n) {
                int end = start + len;
                if (end > n) end = n;
                blk_i1[nb] = base + start; blk_i2[nb] = base + end - 1;
                blk_pinned[nb] = (end < n) ? 1 : 0;
                ++nb;
                start = end;
                len = 2 * half_len;
            }
        }
        base += n;
    }
    return nb;
}
