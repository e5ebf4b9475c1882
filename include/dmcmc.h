/*
 * dmcmc.h -- C ABI of the B200-native path-imputation backend (libdmcmc_b200.so).
 *
 * Drop-in boundary for ONE hot path of JuliaDiffusionBayes/DiffusionMCMC.jl: every entry point
 * below is what the reference's Julia methods for that path would `ccall` instead of calling
 * GuidedProposals.jl / DiffusionDefinition.jl.  File:line citations are relative to the
 * reference checkout (/root/reference).  Plain pointers and sizes only; the caller owns every
 * host buffer, the library copies in/out before returning unless stated otherwise; device
 * memory is owned by the context.  All floating point is IEEE binary64.
 *
 * Return value: 0 on success, <0 on error (text via dmcmc_last_error).  A path that leaves the
 * model's domain is NOT an error: it reports success = 0 and ll = -Inf for that (chain, block),
 * mirroring the `(success, ll)` tuples at src/run!_alterations.jl:18 and :206-209.
 *
 * Index conventions: intervals are the inter-observation intervals of all recordings
 * concatenated (0-based); interval j ends at observation j.  A layout block is an inclusive
 * range [i1, i2] of intervals of one recording -- the `(rec_idx, obs_range)` tuples of
 * src/schedule.jl:97-100.  Chains are a batch axis that the reference does not have
 * (SURVEY.md section 0): every chain owns a full copy of the path/noise state.
 */
#ifndef DMCMC_H
#define DMCMC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMCMC_ABI_VERSION 5

/* Built-in diffusion laws (DiffusionDefinition.jl `@load_diffusion`, src/DiffusionMCMC.jl:4). */
enum dmcmc_model {
    DMCMC_MODEL_FHN = 0,  /* FitzHugh-Nagumo           d=2  m=1  theta=(eps,s,gamma,beta,sigma)        */
    DMCMC_MODEL_LV = 1,   /* Lotka-Volterra            d=2  m=2  theta=(alpha,beta,gamma,delta,s1,s2)  */
    DMCMC_MODEL_L63 = 2,  /* Lorenz-63                 d=3  m=3  theta=(th1,th2,th3,sigma)             */
    DMCMC_MODEL_L96 = 3,  /* Lorenz-96                 d=40 m=40 theta=(F,sigma,c_aux)                 */
    DMCMC_MODEL_LIN2 = 4, /* linear 2-d (analytic KAT) d=2  m=2  theta=(B00,B01,B10,B11,b0,b1,s0,s1)   */
    DMCMC_MODEL_LVM = 5   /* Lotka-Volterra, multiplicative noise (state-dependent sigma)              */
};

typedef struct dmcmc_ctx dmcmc_ctx;

typedef struct {
    int32_t abi_version;   /* DMCMC_ABI_VERSION                                                    */
    int32_t model;         /* enum dmcmc_model                                                     */
    int32_t d;             /* state dimension (fixed by the model except L96)                      */
    int32_t n_chains;      /* chains held by THIS context (one context per GPU)                    */
    int32_t chain_offset;  /* global index of local chain 0 (Philox streams are per global chain,  */
                           /* so results do not depend on how chains are sharded over GPUs)        */
    int32_t interval_offset; /* global index of local interval 0 when one long recording is sharded    */
                           /* over GPUs by contiguous interval ranges (same purpose)     */
    int32_t device;        /* CUDA device ordinal                                                  */
    uint64_t seed;         /* Philox key                                                           */
    double pin_eps;        /* variance of the (almost exact) end-point observation of pinned blocks */
    int32_t store_noise;   /* 1: keep the accepted Wiener increments W in HBM like the reference's    */
                           /* WW containers; 0: never store W, always recover it from the accepted X  */
                           /* (halves the state: needed for Lorenz-96 at full size)                   */
} dmcmc_config;

/* -- lifetime -------------------------------------------------------------------------------- */
int dmcmc_create(const dmcmc_config *cfg, dmcmc_ctx **out);
void dmcmc_destroy(dmcmc_ctx *ctx);
const char *dmcmc_last_error(const dmcmc_ctx *ctx); /* ctx may be NULL: last create() error */

/* -- problem definition (replaces DiffusionGlobalSubworkspace's constructor inputs,
 *    src/workspaces.jl:55-71, :164-170) ------------------------------------------------------ */
/* Imputation grids from OBS.setup_time_grids: n_steps[j] Euler steps in interval j and the
 * n_steps[j]+1 grid times of each interval concatenated in t. */
int dmcmc_set_grid(dmcmc_ctx *ctx, int32_t n_recordings, const int32_t *rec_nobs,
                   const int32_t *n_steps, const double *t);
/* Linear-Gaussian observations v_j ~ N(L_j x, Sigma_j) at the right end of interval j
 * (LinearGsnObs, docs/src/tutorials/simple_inference.md:52-60).  L [J][p][d], Sigma [J][p][p],
 * v [J][p].  LinearGsnObs also carries a mean mu (v ~ N(L x + mu, Sigma)): the filter only ever sees
 * v - mu, so the caller passes v - mu here (julia/DiffusionMCMCB200.jl does). */
int dmcmc_set_obs(dmcmc_ctx *ctx, int32_t p, const double *L, const double *Sigma,
                  const double *v);
/* Known starting points (KnownStartingPt), x0 [n_chains][n_recordings][d]. */
int dmcmc_set_start(dmcmc_ctx *ctx, const double *x0);
/* Parameters of the CURRENT law; rebuilds the built-in auxiliary laws (DD.set_parameters!). */
int dmcmc_set_theta(dmcmc_ctx *ctx, const double *theta, int32_t ntheta);
/* Optional user-supplied auxiliary law, constant on each interval: B [J][d][d], beta [J][d],
 * sigma_aux [J][d][m].  which: 0 = current law, 1 = proposal law. */
int dmcmc_set_aux(dmcmc_ctx *ctx, int32_t which, const double *B, const double *beta,
                  const double *sigma_aux);
/* Block layout of the update about to run (src/workspaces.jl:242-261).  pinned[b] = 1 when the
 * block's right end is conditioned on the current path value (blocking). */
int dmcmc_set_layout(dmcmc_ctx *ctx, int32_t nblk, const int32_t *i1, const int32_t *i2,
                     const int32_t *pinned);
/* Optional per-block activity mask of the CURRENT layout (NULL = all active).  Inactive blocks are
 * skipped by dmcmc_impute / dmcmc_run_sweeps (their histories read NaN).  Used when a recording is
 * sharded over GPUs: halo intervals are present in the context but updated by the neighbour. */
int dmcmc_set_block_mask(dmcmc_ctx *ctx, const int32_t *active);
/* Chequerboard layout helper for BlockingChequerboardSwitch(block_half_len), src/blocking.jl:7-9.
 * pattern 0: delimiters at multiples of 2h; pattern 1: shifted by h.  half_len <= 0: no
 * blocking (src/schedule.jl:97-100).  Arrays must hold n_intervals entries; returns nblk. */
int dmcmc_chequerboard_layout(int32_t n_recordings, const int32_t *rec_nobs, int32_t half_len,
                              int32_t pattern, int32_t *i1, int32_t *i2, int32_t *pinned);
/* pCN memory per interval, update.rho_s[i][j] (src/updates.jl:53-74). */
int dmcmc_set_rho(dmcmc_ctx *ctx, const double *rho);

/* -- a8: backward filter (backward_filter!, src/run!_alterations.jl:203;
 *    recompute_guiding_term!, src/workspaces.jl:428) ------------------------------------------ */
/* Device kernel: H, F, c tables of law `which` (0 current, 1 proposal) for the current layout. */
int dmcmc_backward_filter(dmcmc_ctx *ctx, int32_t which);
/* Host-computed tables instead (bit-exact inputs for parity runs).  Per grid point (N_j+1 per
 * interval): H [P][d][d], F0 [P][d], Psi [P][d][d] (may be NULL when no block is pinned),
 * c0 [P]; per block: g [nblk][d], Qc [nblk][d][d] (may be NULL likewise). */
int dmcmc_upload_tables(dmcmc_ctx *ctx, int32_t which, const double *H, const double *F0,
                        const double *Psi, const double *c0, const double *blk_g,
                        const double *blk_Qc);
int dmcmc_download_tables(dmcmc_ctx *ctx, int32_t which, double *H, double *F0, double *Psi,
                          double *c0, double *blk_g, double *blk_Qc);

/* -- a13: init_paths! + init_ll! (src/workspaces.jl:80-92, :425-431) -------------------------- */
/* Z [n_chains][S][m] explicit standard normals (S = total steps) or NULL for Philox noise. */
int dmcmc_init_paths(dmcmc_ctx *ctx, const double *Z);
int dmcmc_init_ll(dmcmc_ctx *ctx);

/* -- a1-a7: one eMCMC.update!(::PathImputation) (src/run!_alterations.jl:5-26, :36-91) over every
 *    (chain, block): pCN refresh, guided Euler-Maruyama solve, Girsanov ll, accept, swap.
 *    Parity mode: Z [n_chains][S][m], E [n_chains][nblk] (Exp(1) draws).  Production mode:
 *    Z = E = NULL, noise = Philox4x32-10(seed; iter, interval, global chain, step).
 *    Outputs (may be NULL) are the history rows of src/run!_alterations.jl:65-74, shape
 *    [n_chains][nblk]. -------------------------------------------------------------------------- */
int dmcmc_impute(dmcmc_ctx *ctx, uint32_t iter, const double *Z, const double *E,
                 double *out_ll_prop, double *out_ll, uint8_t *out_accepted);

/* Production loop (what `run!` drives for a smoothing chain): n_sweeps consecutive
 * PathImputation updates with Philox noise, iterations iter0, iter0+1, ...  half_len > 0
 * alternates the two chequerboard layouts (pattern first_pattern first), each sweep = layout
 * switch + imputation; half_len <= 0 keeps the current layout.  rho_sched [n_sweeps][J] (may be
 * NULL) goes host->device inside the call.  History rows (any may be NULL) are
 * [n_sweeps][n_chains][nblk_stride] -- the rows of src/run!_alterations.jl:65-74; they stream back
 * while later sweeps run.  Page-locked history buffers (dmcmc_host_alloc, or the caller's own
 * arrays after dmcmc_host_register) are filled by DMA directly; pageable ones are served through
 * the library's pinned ring and a pool of copy threads.  Returns after everything has landed. */
int dmcmc_run_sweeps(dmcmc_ctx *ctx, uint32_t iter0, int32_t n_sweeps, int32_t half_len,
                     int32_t first_pattern, const double *rho_sched, int32_t nblk_stride,
                     double *hist_ll_prop, double *hist_ll, uint8_t *hist_accepted);
/* Device-side adaptation of the pCN memory (AdaptationPathImputation, src/adaptation.jl:22-41, :82-149; same
 * field names and meaning): once set, dmcmc_run_sweeps keeps rho per block on the device -- one vector per
 * chequerboard pattern, started from rho_pattern0 / rho_pattern1 ([J]; NULL: the current dmcmc_set_rho values) --
 * and after every sweep registers the accept counts and, for every block whose counter has reached
 * adapt_every_k_steps proposals, takes the logit-space step of readjust! (delta = scale / sqrt(max(1, iter /
 * adapt_every_k_steps - offset)), iter = MCMC iteration = sweep index / sweeps_per_iteration + 1).  rho_sched must
 * then be NULL.  cfg = NULL switches it off.  dmcmc_get_rho reads the current vector of a pattern. */
typedef struct {
    double target_accpt_rate; /* 0.234 */
    double scale;             /* 1.0   */
    double min, max;          /* 1e-12, 1 - 1e-12 */
    double offset;            /* 1e2   */
    int32_t adapt_every_k_steps; /* 100: proposals per block between adjustments */
} dmcmc_adapt;
int dmcmc_set_adaptation(dmcmc_ctx *ctx, const dmcmc_adapt *cfg, int32_t sweeps_per_iteration,
                         const double *rho_pattern0, const double *rho_pattern1);
int dmcmc_get_rho(dmcmc_ctx *ctx, int32_t pattern, double *rho);
/* Pinned host memory for the streaming buffers above; or page-lock memory the caller already owns
 * (a Julia Array: register once when the workspace is built, unregister before it is freed). */
void *dmcmc_host_alloc(size_t bytes);
void dmcmc_host_free(void *p);
int dmcmc_host_register(void *p, size_t bytes);
int dmcmc_host_unregister(void *p);

/* -- layout switch (change_laws_for_blocking! + recompute_loglikhd!, called but undefined in the
 *    reference: src/workspaces.jl:469-470): new tables, Wiener re-inversion, ll recompute. ------ */
int dmcmc_relayout(dmcmc_ctx *ctx, int32_t nblk, const int32_t *i1, const int32_t *i2,
                   const int32_t *pinned);

/* -- a9: parameter step (set_proposal!, src/run!_alterations.jl:177-211; swap_laws_and_paths!,
 *    :328-334).  theta_prop -> proposal law; critical != 0 re-runs the backward filter; solve
 *    with the FIXED accepted Wiener paths.  out_ll_prop [n_chains][nblk], out_success [n_chains]. */
int dmcmc_param_loglik(dmcmc_ctx *ctx, const double *theta_prop, int32_t ntheta, int32_t critical,
                       double *out_ll_prop, uint8_t *out_success);
int dmcmc_param_commit(dmcmc_ctx *ctx, int32_t accepted);

/* -- read-outs -------------------------------------------------------------------------------- */
/* a12: accepted path of one chain/recording glued as _copy_path! does
 * (src/path_saving_buffer.jl:54-64): x_out [(sum_j N_j)+1][d]. */
int dmcmc_download_path(dmcmc_ctx *ctx, int32_t chain, int32_t rec, double *x_out);
/* The same without stalling the sampler (save!, src/path_saving_buffer.jl:52; the reference saves every
 * 100th iteration, src/run!_alterations.jl:51-54): the glued accepted paths of n_sel chosen chains are
 * snapshotted on the device in stream order and copied out while later sweeps run.  *ticket identifies
 * the save (at most four in flight); dmcmc_save_path_wait blocks until it has landed and fills
 * x_out [n_sel][(sum_j N_j)+1][d]. */
int dmcmc_save_path_async(dmcmc_ctx *ctx, int32_t n_sel, const int32_t *chains, int32_t rec,
                          int32_t *ticket);
int dmcmc_save_path_wait(dmcmc_ctx *ctx, int32_t ticket, double *x_out);
/* Full containers in the reference's per-interval shape (N_j+1 points per interval, duplicated
 * joins): X [n_chains][P][d], W [n_chains][P][m] cumulative.  which: 0 accepted, 1 proposal. */
int dmcmc_download_state(dmcmc_ctx *ctx, int32_t which, double *X, double *W);
/* Number of blocks of the current layout (-1: none). */
int32_t dmcmc_n_blocks(const dmcmc_ctx *ctx);
/* Accepted path values of intervals [j0, j1) (points 1..N_j of each, i.e. without the start point),
 * X [n_chains][steps][d]: the halo exchange of a block-sharded recording.  Uploading
 * overwrites the accepted containers and invalidates the stored noise (it is re-inverted). */
int dmcmc_download_segment(dmcmc_ctx *ctx, int32_t j0, int32_t j1, double *X);
int dmcmc_upload_segment(dmcmc_ctx *ctx, int32_t j0, int32_t j1, const double *X);
int dmcmc_get_ll(dmcmc_ctx *ctx, double *ll /* [n_chains][nblk] */);
int dmcmc_set_ll(dmcmc_ctx *ctx, const double *ll);
/* Device-side accept counters per block since the last call (src/adaptation.jl:100-103). */
int dmcmc_accept_counts(dmcmc_ctx *ctx, int64_t *accepted /* [nblk] */, int64_t *proposed);

/* -- conjugate Gaussian parameter update (src/updates.jl:83-105, :225-262) ----------------------------
 * For a drift whose rough coordinates are linear in some parameters, b = phi(x) theta_c + phi_c(x), the
 * path integrals  mu = sum phi' Gamma^-1 (dy - phi_c dt),  W = sum phi' Gamma^-1 phi dt  over the accepted
 * path (compute_mu_and_W, :241-262) are reduced on the device; the posterior draw stays on the host.
 * dmcmc_conjugate_dim: number nc of such parameters of the context's law and their theta indices
 * (FHN: gamma, beta; Lotka-Volterra: alpha..delta; Lorenz-63: theta1..theta3; 0 = none).
 * Parameters of that set that are held fixed enter as  mu_S - W[S,F] theta_F  on the host. */
int32_t dmcmc_conjugate_dim(const dmcmc_ctx *ctx, int32_t *theta_idx /* [nc] or NULL */);
int dmcmc_conjugate_sums(dmcmc_ctx *ctx, double *mu /* [C][nc] */, double *W /* [C][nc][nc] */);

/* -- device-side exchange for a block-sharded recording ------------------------------
 * The halo / end-point messages between neighbouring GPUs never have to touch the host: the caller
 * (NCCL send/recv, or a peer copy) works on DEVICE buffers, ordered on the context's stream.
 * Element (chain, step, comp) of a segment buffer sits at dev[chain * chain_stride + step * d + comp]. */
int dmcmc_segment_device(dmcmc_ctx *ctx, int32_t j0, int32_t j1, double *dev, int64_t chain_stride,
                         int32_t upload /* 0: accepted X -> dev, 1: dev -> accepted X */);
/* Starting points from a device buffer: x0 of (chain, recording r) at dev[chain * chain_stride + r * d + comp]. */
int dmcmc_set_start_device(dmcmc_ctx *ctx, const double *dev, int64_t chain_stride);
/* The context's CUDA stream (cudaStream_t), for ordering the caller's collectives with the kernels. */
void *dmcmc_stream(dmcmc_ctx *ctx);
/* One PathImputation update with in-register Philox noise under the current layout and block mask:
 * enqueued, not awaited; ll, selectors and accept counters stay on the device
 * (dmcmc_get_ll / dmcmc_accept_counts read them later). */
int dmcmc_impute_async(dmcmc_ctx *ctx, uint32_t iter);

/* -- collectives of a block-sharded recording, inside the library -----------------------
 * One context per GPU/process; the ranks that share one long recording form an NCCL communicator.  The
 * unique id (128 bytes) is made by one rank and handed to the others by the host (file, socket, MPI ...).
 * NCCL is bound at run time (libnccl.so.2); a process that already carries one (PyTorch) is reused. */
int dmcmc_comm_unique_id(void *id128);
int dmcmc_comm_init(dmcmc_ctx *ctx, const void *id128, int32_t rank, int32_t world);
int dmcmc_comm_destroy(dmcmc_ctx *ctx);
/* End-point / halo exchange after a sweep (ncclSend / ncclRecv on the context's stream, packed by the
 * library, no host synchronisation).  The context holds its rank's contiguous interval range followed by
 * a halo of n_halo intervals owned by the right neighbour (0 on the last rank); n_first = my first
 * intervals that the LEFT neighbour holds as its halo.
 *   after_pattern 0: rank r -> r-1 accepted path of my first n_first intervals
 *   after_pattern 1: rank r -> r+1 my last owned interval (its end = the neighbour's new start point) + halo
 * left_edge_steps: Euler steps of the left neighbour's last owned interval (0 on rank 0). */
int dmcmc_halo_exchange(dmcmc_ctx *ctx, int32_t after_pattern, int32_t n_first, int32_t n_halo,
                        int32_t left_edge_steps);
/* The same exchange over peer memory instead of NCCL (collective over the communicator, once after dmcmc_comm_init):
 * every rank allocates a mailbox per direction, the CUDA IPC handles travel by one ncclAllGather, each rank maps its two
 * neighbours.  From then on dmcmc_halo_exchange is two small kernels per rank: the sender's pack kernel stores straight
 * into the neighbour's mailbox over NVLink and raises a sequence flag there, the receiver's unpack kernel spins on the
 * flag, unpacks and acknowledges -- no NCCL call, no host synchronisation (a ncclSend / ncclRecv pair of a few hundred
 * bytes costs ~25 us; this is what a single chain's block sharding is bound by).  Messages larger than mailbox_bytes,
 * and every exchange when IPC is unavailable on any rank, still go through NCCL; results are identical either way.
 * dmcmc_comm_p2p_enabled: 1 when peer memory is in use. */
int dmcmc_comm_enable_p2p(dmcmc_ctx *ctx, int64_t mailbox_bytes);
/* dmcmc_impute_async + the dmcmc_halo_exchange that follows it under `pattern`, FUSED: with peer memory on and a launch
 * small enough for the warp-per-unit kernel (a single chain, or a few), the units of the layout's edge block store their
 * freshly accepted intervals into the neighbour's mailbox right after their accept decision, inside the sweep kernel, and
 * the neighbour's next such sweep takes the message out of its mailbox before that block's units read anything -- no
 * kernel boundary and no pack / unpack pass between a sweep and its exchange.  A message still in the mailbox when
 * anything else touches the paths (read-outs, parameter step, layout switch, dmcmc_halo_exchange, dmcmc_comm_destroy) is
 * taken out first.  Otherwise (no peer memory, many chains, d = 40) it is the two calls one after the other; results
 * are bit-identical either way. */
int dmcmc_impute_exchange_async(dmcmc_ctx *ctx, uint32_t iter, int32_t pattern, int32_t n_first, int32_t n_halo,
                                int32_t left_edge_steps);
int dmcmc_comm_p2p_enabled(const dmcmc_ctx *ctx);
/* a9 on a block-sharded recording (set_proposal!, src/run!_alterations.jl:177-211; ONE accept decision
 * for all blocks, :367): every rank solves its blocks, the per-block ll deg / ll are all-reduced over the
 * GLOBAL block index (blk_global[b] = global index of local block b, -1 = owned by another rank) and
 * `success` is AND-reduced.  Outputs are identical on all ranks and bit-identical to the unsharded
 * context's: out_ll_prop_global, out_ll_global [n_chains][nblk_global], out_success [n_chains].
 * Follow with dmcmc_param_commit(accepted) on every rank. */
int dmcmc_param_loglik_sharded(dmcmc_ctx *ctx, const double *theta_prop, int32_t ntheta, int32_t critical,
                               int32_t nblk_global, const int32_t *blk_global, double *out_ll_prop_global,
                               double *out_ll_global, uint8_t *out_success);

/* -- measurement helpers (bench.py) ----------------------------------------------------------- */
/* Device time (ms, CUDA events on the context's stream) and launch count of the solve kernels
 * since the previous call, per kernel mode: [0] imputation, [1] parameter-step solve,
 * [2] layout switch (W re-inversion + ll), [3] ll only. */
int dmcmc_kernel_times(dmcmc_ctx *ctx, double ms[4], int64_t launches[4]);
/* The event pair costs about 5 us of stream time per launch, so only every n-th solve launch is
 * timed (default 8; 1 = every launch, 0 = never).  `launches` of dmcmc_kernel_times counts the timed
 * launches only; dmcmc_launch_count counts all solve launches since the previous call. */
int dmcmc_set_timing(dmcmc_ctx *ctx, int32_t every);
int64_t dmcmc_launch_count(dmcmc_ctx *ctx);
/* Stop-watch on the context's stream: CUDA events bracketing everything enqueued between the two
 * calls (dmcmc_timer_stop synchronises the stream). */
int dmcmc_timer_start(dmcmc_ctx *ctx);
int dmcmc_timer_stop(dmcmc_ctx *ctx, double *ms);
int64_t dmcmc_device_bytes(const dmcmc_ctx *ctx);
/* Measured FP64 FMA throughput of a device in TFLOP/s (independent DFMA chains on every SM, best of five): the
 * denominator of the FP64-pipe view of the Lorenz-96 kernel (bench.py). */
double dmcmc_fp64_peak(int32_t device);
/* The same through the FP64 tensor cores (mma.sync m8n8k4.f64, SASS DMMA.8x8x4), the instruction the Lorenz-96 kernel's
 * 40 x 40 mat-vecs run on; on B200 both share one FP64 datapath, so this is that kernel's ceiling. */
double dmcmc_fp64_tensor_peak(int32_t device);
/* Threads per CTA of the solve kernels (32..128, multiple of 32; default: chosen per launch by unit count). */
int dmcmc_set_threads(dmcmc_ctx *ctx, int32_t threads);
/* Model-specific fused step kernels (FitzHugh-Nagumo with its built-in aux law); on by default.
 * Off = the generic step for every model (same mathematics, different rounding). */
int dmcmc_set_fast_step(dmcmc_ctx *ctx, int32_t on);
/* Which imputation kernel serves dmcmc_impute / dmcmc_run_sweeps: one thread per (chain, block) unit, or --
 * for launches with few, long units (the reference's own shapes: one chain, no or few blocks,
 * src/run!_alterations.jl:15-16 loops over them serially) -- one warp per unit.  -1: chosen by unit count
 * (default), 0: never, 1: always.  The two kernels give bit-identical results. */
int dmcmc_set_cooperative(dmcmc_ctx *ctx, int32_t mode);

#ifdef __cplusplus
}
#endif
#endif /* DMCMC_H */
