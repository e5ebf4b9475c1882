// dmcmc_internal.h -- structures shared by the C-ABI layer (dmcmc_capi.cu) and the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "dmcmc_models.cuh"

namespace dmcmc {

enum SolveMode {
    MODE_IMPUTE = 0,    // a1-a7: pCN refresh + guided solve + ll + accept/swap
    MODE_PARAM = 1,     // a9: solve with the fixed accepted noise under the proposal law
    MODE_RELAYOUT = 2,  // layout switch: recover dW from the accepted X, recompute ll
    MODE_LOGLIK = 3     // init_ll!: ll of the accepted path only
};

// End-point exchange of a block-sharded recording FUSED into the imputation sweep (warp-per-unit kernel, peer memory):
// the units of the layout's edge block pull the neighbour's message out of this GPU's mailbox before they read anything,
// and push their freshly accepted intervals into the neighbour's mailbox right after their accept decision -- no kernel
// boundary and no separate pack / unpack pass between a sweep and its exchange (dmcmc_impute_exchange_async).
struct HaloFuse {
    int32_t blk;                       // the edge block under this sweep's layout; -1: nothing is fused
    const int32_t *iv_st_off;          // [J + 1] cumulative Euler steps
    // pull: own mailbox -> accepted containers of intervals [pull_j0, pull_j1) (+ starting point: the last D of the
    // pull_lead doubles that precede them in every chain's slice)
    const double *pull_box;            // nullptr: no pull
    int32_t pull_j0, pull_j1, pull_lead;
    int64_t pull_stride;
    const unsigned long long *pull_flag;   // this GPU's memory, raised by the sender
    unsigned long long pull_val;
    unsigned long long *pull_ack;          // the sender's memory
    unsigned int *pull_counter;
    // push: accepted path of intervals [push_j0, push_j1) after this sweep -> the neighbour's mailbox
    double *push_box;                  // nullptr: no push
    int32_t push_j0, push_j1;
    int64_t push_stride;
    const unsigned long long *push_wait;   // this GPU's memory: the neighbour's acknowledgement of the previous message
    unsigned long long push_wait_val;
    unsigned long long *push_flag;         // the neighbour's memory
    unsigned long long push_val;
    unsigned int *push_counter;
    int *err;                          // host-mapped: 1 when a wait timed out
};

// Everything a solve kernel needs.  Pointers are device pointers.
struct SolveParams {
    int32_t C, nblk, J, D;
    int32_t chain_offset, interval_offset;
    int64_t TT;  // total tiles over all intervals
    // layout
    const int32_t *blk_i1, *blk_i2, *blk_pinned;
    const int32_t *blk_active;  // nullptr: every block is updated
    // intervals
    const int32_t *iv_N, *iv_tile_off, *iv_first, *iv_rec;
    const int64_t *iv_xoff; // [J] first element of interval j in an X container
    const int32_t *iv_xs;   // [J] per-chain stride of interval j in an X container: even(N_j * D)
    const double *rho;      // [J]
    const double *auxB;     // [J][D*D]
    const double *auxbeta;  // [J][D]
    const double *auxatil;  // [J][D*D]
    // tables of the law the kernel solves under
    const double *table;    // [TT*8][REC] records the kernel streams (FitzHugh-Nagumo fused step: derived coefficients)
    const double *table_raw; // [TT*8][REC] dt, sqrt(dt), H, F0, Psi (== table unless the fused step is used)
    const double *blk_c0;   // [nblk] c0 at the block's first grid point
    const double *blk_g;    // [nblk][D]
    const double *blk_Qc;   // [nblk][D*D]
    int32_t has_psi;
    // state
    double *X[2];           // [iv_xoff[j] + chain * iv_xs[j] + k * D + comp]  one record per (interval, chain)
    double *W[2];           // [TT][C][M][8]
    // accepted-container selectors [J][C]: read from *_in, written (flipped on accept) to *_out;
    // the two arrays are swapped by the host after the sweep, so no unit ever reads a selector
    // that a neighbouring block's unit is flipping.
    const uint8_t *selX_in, *selW_in;
    uint8_t *selX_out, *selW_out;
    const double *x0;       // [R][D][C]
    double *ll;             // [C][nblk]
    // noise
    const double *Z;        // [TT][C][M][8] explicit normals or nullptr (Philox)
    const double *E;        // [C][nblk] explicit Exp(1) or nullptr
    uint64_t seed;
    uint32_t iter;
    int32_t force_accept;
    int32_t reinvert;   // 1: accepted noise is recovered from the accepted X (stored W is stale)
    int32_t store_w;    // 1: write the proposal noise W deg (0 when the next update re-inverts anyway)
    int32_t fast_ok;    // 1: the law uses the model's built-in aux law (fused model-specific step allowed)
    int32_t coop;       // warp-per-unit kernel for launches with few units: -1 auto (by unit count), 0 never, 1 always
    int32_t mean_N;     // mean Euler steps per interval (host-side hint for kernel variants)
    int32_t depth;      // tiles in flight of the table / input pipelines (2, 4 or 8; set by the launcher)
    int32_t out_stride; // row stride (per chain) of out_llprop / out_ll / out_acc
    const uint8_t *unit_mask;  // [C][nblk] or nullptr
    // outputs
    double *out_llprop, *out_ll;  // [C][nblk]
    uint8_t *out_acc;             // [C][nblk]  (MODE_PARAM: per-unit success)
    unsigned long long *acc_count, *prop_count;  // [nblk]
    Theta th;
    HaloFuse hf;        // hf.blk < 0 unless the launch is dmcmc_impute_exchange_async's (warp-per-unit kernel only)
};

// launchers (dmcmc_solve.cu); return the CUDA error of the launch
cudaError_t launch_solve(int model, int mode, const SolveParams &p, int threads, cudaStream_t s);
bool solve_uses_warp_per_unit(int model, const SolveParams &p);  // the kernel launch_solve will pick for p
cudaError_t launch_solve_big(int d, int mode, const SolveParams &p, cudaStream_t s);  // dmcmc_solve_l96.cu
cudaError_t launch_solve_coop(int model, int mode, const SolveParams &p, cudaStream_t s);  // dmcmc_solve_coop.cu (MODE_IMPUTE, MODE_PARAM)
cudaError_t launch_fhn_derive(const double *raw, double *out, long long nrec, int has_psi, const Theta &th, cudaStream_t s);

// state gather for read-outs (dmcmc_solve.cu)
struct GatherParams {
    int32_t C, J, D, M;
    int64_t TT;
    const int32_t *iv_N, *iv_tile_off, *iv_first, *iv_rec, *iv_pt_off, *iv_blk;
    const int64_t *iv_xoff;
    const int32_t *iv_xs;
    const int32_t *blk_i1;
    const double *X[2], *W[2];
    const uint8_t *selX, *selW;
    const double *x0;
    int32_t which;  // 0 accepted, 1 proposal
    double *outX;   // [C][P][D]
    double *outW;   // [C][P][M]
    int64_t P;
};
cudaError_t launch_gather(const GatherParams &g, cudaStream_t s);

// halo exchange of a block-sharded recording: accepted X of intervals [j0, j1), packed [C][steps][D]
struct SegmentParams {
    int32_t C, D, j0, j1, steps;
    int32_t c0, nc;   // chains [c0, c0 + nc) are packed (nc = 0: all)
    int64_t TT;
    const int32_t *iv_N, *iv_tile_off, *iv_st_off;
    const int64_t *iv_xoff;
    const int32_t *iv_xs;
    double *X[2];
    const uint8_t *selX;
    double *packed;   // device: element (chain, step, comp) at packed[chain * pk_stride + step * D + comp]
    int64_t pk_stride;
    int32_t upload;   // 0: X -> packed, 1: packed -> X
};
cudaError_t launch_segment(const SegmentParams &g, cudaStream_t s);

// Halo exchange over peer memory (NVLink P2P, CUDA IPC mailboxes; dmcmc_comm_enable_p2p): the pack kernel of the sending
// rank writes straight into the neighbour's mailbox and raises a sequence flag there, the unpack kernel of the receiving
// rank spins on that flag, unpacks and acknowledges -- no NCCL call, no host involvement (one ncclSend / ncclRecv pair
// of a few hundred bytes costs ~25 us, measured).
struct P2PSync {
    const unsigned long long *wait_flag;  // memory of THIS GPU, written by the peer; nullptr: nothing to wait for
    unsigned long long wait_val;          // proceed when *wait_flag >= wait_val
    unsigned long long *signal_flag;      // memory of the PEER (mapped through CUDA IPC); nullptr: nothing to signal
    unsigned long long signal_val;
    unsigned int *counter;                // this GPU: CTAs of the launch that are done (reset by the last one)
    int *err;                             // host-mapped: set to 1 when a wait times out (the peer died)
};
// one launch per rank and exchange: `push` (accepted X of intervals [j0, j1) -> push->packed = the peer's mailbox) and / or
// `pull` (pull->packed = own mailbox -> accepted X; with start_src the new starting point
// x0[(rec * D + comp) * C + chain] = start_src[chain * start_stride + rec * D + comp] as well); either may be null
cudaError_t launch_halo_exchange(const SegmentParams *push, const P2PSync &ys, const SegmentParams *pull, double *x0,
                                 const double *start_src, int64_t start_stride, int R, const P2PSync &yr, cudaStream_t s);
cudaError_t launch_set_start(double *x0, const double *src, int64_t chain_stride, int C, int R, int D, cudaStream_t s);

// noise transposition [C][S][M] -> [M][TT][C][8] is done on the host in the C-ABI layer.

// conjugate Gaussian update path integrals (dmcmc_conjugate.cu)
struct ConjParams {
    int32_t C, J;
    const int32_t *iv_N, *iv_first, *iv_rec, *iv_xs, *iv_pt_off;
    const int64_t *iv_xoff;
    const double *t;       // [P] grid times
    const double *X[2];
    const uint8_t *selX;
    const double *x0;
    double *partial;       // [J][C][nc + nc*nc]
    Theta th;
};
int conjugate_dim(int model, int32_t *idx);  // number of conjugate parameters of the law (0: none) and their theta indices
cudaError_t launch_conjugate(int model, const ConjParams &c, double *out /* device [C][nc + nc*nc] */, cudaStream_t s);

// backward filter (dmcmc_bf.cu)
struct BFParams {
    int32_t D, J, nblk;
    const int32_t *blk_i1, *blk_i2, *blk_pinned;
    const int32_t *iv_N, *iv_tile_off, *iv_pt_off;
    const double *t;        // [P]
    const double *auxB, *auxbeta, *auxatil;
    const double *Hobs;     // [J][D*D]
    const double *Fobs;     // [J][D]
    const double *cobs;     // [J]
    double pin_eps;
    int32_t has_psi, rec;
    int32_t mean_N;         // mean Euler steps per interval (launch-shape hint)
    double *table;          // [TT*8][rec]
    double *ptH, *ptF0, *ptPsi, *ptc0;  // per-grid-point copies (download / parity), may be null
    double *blk_c0, *blk_g, *blk_Qc;
};
cudaError_t launch_backward_filter(const BFParams &b, cudaStream_t s);
// built-in aux law (B, beta, a~ = sigma~ sigma~') of every interval from theta and the end observations v [J][p]
cudaError_t launch_aux_build(int model, int d, int m, int p, int J, const Theta &th, const double *v, double *B, double *beta,
                             double *atil, cudaStream_t s);

}  // namespace dmcmc
