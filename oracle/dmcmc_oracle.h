/*
 * dmcmc_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C, FP64, restatement of the path-imputation hot path of
 * JuliaDiffusionBayes/DiffusionMCMC.jl (reference mounted at /root/reference).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library; the product path
 * (diffusionmcmc.jl_b200/csrc) never links or calls it.
 *
 * PARITY STATUS: **parity unpinned** for the arithmetic.  The reference repo only
 * orchestrates (src/run!_alterations.jl:5-91); the arithmetic lives in un-vendored
 * packages (GuidedProposals.jl v0.1.0 tree 61855a13..., DiffusionDefinition.jl
 * v0.1.0 tree 529e7251..., ExtensibleMCMC.jl v0.1.0 tree 4e4360c9..., pinned in
 * Manifest.toml:199-205, :120-126, :156-162) and the reference has no golden
 * vector for it (test/runtests.jl:56-58 is an empty testset).  Julia is not
 * installed, so the reference cannot be run here.  What IS pinned:
 *   - the schedule known-answer test (test/runtests.jl:8-54) -> tests/test_schedule.py
 *   - Philox4x32-10 against the published Random123 known-answer vectors
 *   - analytic identities (linear target == aux law  =>  G == 0, ll = log h~(x_a);
 *     rho = 1 => X deg == X;  OU closed forms for the backward recursion).
 *
 * Each function cites the reference file:line whose behaviour it restates.
 */
#ifndef DMCMC_ORACLE_H
#define DMCMC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXD 40

/* model ids (shared with include/dmcmc.h) */
enum {
    ORC_MODEL_FHN = 0,  /* FitzHugh-Nagumo d=2 m=1  th=(eps,s,gamma,beta,sigma)          */
    ORC_MODEL_LV = 1,   /* Lotka-Volterra  d=2 m=2  th=(alpha,beta,gamma,delta,s1,s2)     */
    ORC_MODEL_L63 = 2,  /* Lorenz-63       d=3 m=3  th=(th1,th2,th3,sigma)                */
    ORC_MODEL_L96 = 3,  /* Lorenz-96       d=any m=d th=(F,sigma,c_aux)                   */
    ORC_MODEL_LIN2 = 4, /* linear d=2 m=2  th=(B00,B01,B10,B11,b0,b1,s0,s1) (analytic KAT) */
    ORC_MODEL_LVM = 5   /* LV, multiplicative noise sigma(x)=diag(s1 x, s2 y) (non-const a) */
};

typedef struct {
    int32_t model, d, m, nth;
    const double *th;        /* [nth]                                                     */
    int32_t J;               /* number of inter-observation intervals (all recordings)    */
    int32_t R;               /* number of recordings                                      */
    const int32_t *N;        /* [J]   Euler steps in interval j                           */
    const int32_t *pt_off;   /* [J+1] offset of interval j in grid-point arrays (N_j+1)   */
    const int32_t *st_off;   /* [J+1] offset of interval j in step arrays (N_j)           */
    const double *t;         /* [P]   grid times, P = pt_off[J]                           */
    const double *auxB;      /* [J][d*d] aux drift matrix, constant on the interval       */
    const double *auxbeta;   /* [J][d]                                                    */
    const double *auxsig;    /* [J][d*m] aux diffusion coefficient                        */
    int32_t nblk;            /* blocks in the current layout                              */
    const int32_t *blk_i1;   /* [nblk] first interval (0-based, inclusive)                */
    const int32_t *blk_i2;   /* [nblk] last interval (0-based, inclusive)                 */
    const int32_t *blk_pinned; /* [nblk] 1: right end conditioned on the current path value */
    const double *H;         /* [P][d*d]  backward-filter tables for this layout          */
    const double *F0;        /* [P][d]                                                    */
    const double *Psi;       /* [P][d*d]  (only read in pinned blocks; may be NULL)       */
    const double *c0;        /* [P]                                                       */
    const double *blk_g;     /* [nblk][d]   (pinned blocks)                               */
    const double *blk_Qc;    /* [nblk][d*d] (pinned blocks)                               */
    const double *rho;       /* [J] pCN memory per interval                               */
    int32_t C;               /* chains                                                    */
    int32_t chain_offset;    /* global index of chain 0 / interval 0: Philox streams are keyed   */
    int32_t interval_offset; /* by GLOBAL indices so shards reproduce the unsharded run         */
} orc_problem;

typedef struct {
    double *X[2];   /* [C][P][d]  two containers per (chain, interval); sel picks "accepted" */
    double *W[2];   /* [C][P][m]  cumulative Wiener path per interval, W[pt_off[j]] = 0       */
    uint8_t *selX;  /* [C][J]  which container holds the accepted X of (chain, interval)      */
    uint8_t *selW;  /* [C][J]  same for W (the parameter step swaps X but keeps W)            */
    double *ll;     /* [C][nblk]  accepted log-likelihood per block                           */
} orc_state;

/* ---- model definitions (restating DiffusionDefinition.jl) ---- */
int orc_model_dims(int model, int d_in, int *d, int *m, int *nth);
void orc_drift(int model, int d, const double *th, double t, const double *x, double *b);
void orc_sigma(int model, int d, int m, const double *th, const double *x, double *sig);
int orc_constdiff(int model);
int orc_bound_satisfied(int model, int d, const double *x);
void orc_aux_build(int model, int d, int m, const double *th, const double *v, int p,
                   double *B, double *beta, double *sigt);

/* ---- observation terms: H += L'S^-1 L etc. ---- */
int orc_obs_terms(int d, int p, const double *L, const double *Sigma, const double *v,
                  double *Hobs, double *Fobs, double *cobs);

/* ---- a8 backward filter, one layout ---- */
void orc_transition(int d, const double *B, const double *beta, const double *atil, double dt,
                    double *Phi, double *Q, double *mvec);
void orc_backward_filter(int d, int J, const int32_t *N, const int32_t *pt_off, const double *t,
                         const double *auxB, const double *auxbeta, const double *auxatil,
                         const double *Hobs, const double *Fobs, const double *cobs,
                         int nblk, const int32_t *blk_i1, const int32_t *blk_i2,
                         const int32_t *blk_pinned, double pin_eps,
                         double *H, double *F0, double *Psi, double *c0,
                         double *blk_g, double *blk_Qc);
/* scheme 0 = step by step (above); scheme 1 = every grid point of an interval from the interval's right-end value by
 * one transition over t_N - t_k (what the device computes for d <= 3).  Same exact flow, different rounding. */
void orc_backward_filter_scheme(int d, int J, const int32_t *N, const int32_t *pt_off, const double *t,
                                const double *auxB, const double *auxbeta, const double *auxatil,
                                const double *Hobs, const double *Fobs, const double *cobs,
                                int nblk, const int32_t *blk_i1, const int32_t *blk_i2,
                                const int32_t *blk_pinned, double pin_eps, int scheme,
                                double *H, double *F0, double *Psi, double *c0,
                                double *blk_g, double *blk_Qc);

/* ---- a2 pCN refresh, a4+a5 solve, a3 guiding term ---- */
void orc_wiener_refresh(int N, int m, const double *t, const double *Z, double *Wp);
void orc_crank_nicolson(int N, int m, double rho, const double *Wacc, double *Wp);
int orc_solve_and_ll(const orc_problem *p, int j, const double *xb, const double *W,
                     const double *y1, double *X, double *ll_out);
double orc_loglik_obs(const orc_problem *p, int blk, const double *x0, const double *xb);
double orc_loglikhd_block(const orc_problem *p, const orc_state *s, int chain, int blk);

/* ---- a1,a6,a7 one PathImputation update over every (chain, block) ---- */
void orc_impute_sweep(const orc_problem *p, orc_state *s, const double *Z, const double *E,
                      uint64_t seed, uint32_t iter, int use_philox, int force_accept,
                      const uint8_t *unit_mask, double *out_llprop, double *out_ll,
                      uint8_t *out_acc, int nthreads);

/* ---- a9 parameter-step solve with fixed accepted W ---- */
void orc_param_solve(const orc_problem *p_prop, orc_state *s, double *out_llprop,
                     uint8_t *out_success, int nthreads);
void orc_param_commit(const orc_problem *p, orc_state *s, const uint8_t *accepted_per_chain,
                      const double *llprop);

/* ---- layout change: W re-inversion + ll recompute (OUR definition of the TODOs at
 *      src/workspaces.jl:469-470) ---- */
void orc_relayout(const orc_problem *p_new, orc_state *s, int nthreads);

/* ---- init_ll! (src/workspaces.jl:425-431) ---- */
void orc_init_ll(const orc_problem *p, orc_state *s, int nthreads);

/* ---- layouts (src/schedule.jl:97-100; chequerboard = OUR definition) ---- */
int orc_chequerboard_layout(int R, const int32_t *rec_nobs, int half_len, int pattern,
                            int32_t *blk_i1, int32_t *blk_i2, int32_t *blk_pinned);

/* ---- counter-based noise (production mode) ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_philox_normal4(uint64_t seed, uint32_t iter, uint32_t interval, uint32_t chain,
                        uint32_t wcomp, uint32_t quad, double *z);
double orc_philox_exponential(uint64_t seed, uint32_t iter, uint32_t first_interval,
                              uint32_t chain);

/* ---- a12 glue a chain's accepted path (src/path_saving_buffer.jl:46-64) ---- */
void orc_glue_path(const orc_problem *p, const orc_state *s, int chain, int rec_i1, int rec_i2,
                   double *out);

#ifdef __cplusplus
}
#endif
#endif
