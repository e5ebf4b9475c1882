// dmcmc_capi.cu -- the C ABI of include/dmcmc.h: context, problem upload, launches, read-outs.
// Host-side logic only; the arithmetic of the hot path is in dmcmc_solve.cu / dmcmc_bf.cu.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/prctl.h>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include "../../include/dmcmc.h"
#include "dmcmc_device.cuh"
#include "dmcmc_internal.h"

using namespace dmcmc;

namespace {
std::string g_create_error;

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// tables and block constants of one (law, layout) pair
struct Tables {
    DevBuf<double> table, blk_c0, blk_g, blk_Qc;
    DevBuf<double> ftable;   // FitzHugh-Nagumo fused step: per-step coefficients derived from `table` and theta
    bool fvalid = false;
    DevBuf<double> ptH, ptF0, ptPsi, ptc0;  // per grid point, for download / parity
    bool valid = false;
};

struct Layout {
    std::vector<int32_t> i1, i2, pinned, iv_blk;
    int32_t nblk = 0;
    bool has_psi = false;
    DevBuf<int32_t> d_i1, d_i2, d_pinned, d_iv_blk, d_active;
    bool has_mask = false, has_mask_copy = false;
    std::vector<int32_t> active_host;
    Tables tab[2];  // per law slot
    uint64_t last_used = 0;
    bool used = false;
};

struct Law {
    double theta[DMCMC_MAX_THETA] = {0};
    std::vector<double> auxB, auxbeta, auxsig, auxatil;
    DevBuf<double> d_auxB, d_auxbeta, d_auxatil;
    bool have_theta = false, custom_aux = false;
    bool host_aux_valid = false;  // auxB / auxbeta / auxsig hold this law (false: built on the device only)
};
}  // namespace

// staging rings and streams of dmcmc_run_sweeps (allocated on first use, kept for the context's lifetime)
constexpr int SWEEP_RING = 16;   // ring slots: sweeps whose history rows may be in flight
constexpr int NDRAIN = 4;        // drainer threads (pageable caller buffers only)

// what the drainer threads of one dmcmc_run_sweeps call copy: pinned ring -> the caller's pageable arrays
struct DrainJob {
    int n_sweeps = 0;
    size_t row = 0, slot_bytes = 0;
    double *llp = nullptr, *ll = nullptr;
    uint8_t *acc = nullptr;
};

struct SweepIO {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t kernel_done[SWEEP_RING] = {}, copied[SWEEP_RING] = {};
    DevBuf<uint8_t> d_hist;   // [R][ll deg row | ll row | accepted row]
    DevBuf<double> d_rho;     // [chunk][J]
    uint8_t *h_hist = nullptr;  // pinned mirror of d_hist (pageable caller buffers only)
    size_t h_hist_bytes = 0;
    // persistent drainer pool: started by the first call that needs it, parked on cv_job between calls
    std::thread drainer[NDRAIN];
    bool pool_started = false, stop = false;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    uint64_t gen = 0;
    int done = 0;
    DrainJob job;
    std::atomic<int> recorded{0};
    std::atomic<bool> failed{false};
    std::atomic<int> drained[NDRAIN];
};

// device snapshots of glued accepted paths on their way to the host (dmcmc_save_path_async / _wait)
struct PathSave {
    DevBuf<double> d_buf;
    double *h_buf = nullptr;   // pinned
    size_t h_n = 0, n = 0;
    cudaEvent_t snap = nullptr, landed = nullptr;
    bool busy = false;
};

struct dmcmc_ctx {
    dmcmc_config cfg{};
    int d = 0, m = 0, nth = 0, C = 0, R = 0, J = 0, p = 0;
    int64_t P = 0, S = 0, TT = 0;
    std::string err;
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> ev;  // pairs (start, stop) of timed solve launches
    std::vector<int> ev_mode;     // SolveMode of each pair
    size_t ev_used = 0;
    bool timing = true;
    int time_every = 8;   // kernel-duration events around every n-th solve launch (0: never)
    uint64_t launches = 0, launches_reported = 0;
    cudaEvent_t sw0 = nullptr, sw1 = nullptr;  // stop-watch (dmcmc_timer_start / stop)
    // host copies
    std::vector<int32_t> rec_nobs, N, pt_off, st_off, tile_off, iv_first, iv_rec, xs;
    std::vector<int64_t> xoff;  // [J+1] X container layout: interval j starts at xoff[j], chain stride xs[j]
    std::vector<double> t, L, Sigma, v, Hobs, Fobs, cobs, rho, x0;
    bool have_grid = false, have_obs = false, have_start = false;
    Law law[2];
    int law_cur = 0;
    Layout layout[2];
    int lay_cur = -1;
    uint64_t lay_clock = 0;
    // device
    DevBuf<int32_t> d_N, d_tile_off, d_first, d_rec, d_pt_off, d_st_off, d_xs;
    DevBuf<int64_t> d_xoff;
    DevBuf<double> d_t, d_rho, d_Hobs, d_Fobs, d_cobs, d_x0, d_ll, d_Z, d_E, d_v;
    DevBuf<double> d_X[2], d_W[2];
    DevBuf<uint8_t> d_selX[2], d_selW[2], d_mask, d_out_acc;
    int sel_cur = 0;
    DevBuf<double> d_out_llprop, d_out_ll, d_param_llprop;
    DevBuf<unsigned long long> d_counts;  // [2][J]: accepted, proposed
    std::vector<double> h_param_llprop;
    bool param_pending = false;
    bool w_valid = false;  // stored accepted W is consistent with the accepted X under the current law/layout
    int threads = 0;      // threads per CTA of the thread-per-unit solve kernel (0: chosen by the launcher)
    bool fast_step = true;  // allow model-specific fused steps (dmcmc_set_fast_step)
    int coop = -1;          // warp-per-unit imputation kernel: -1 auto, 0 never, 1 always (dmcmc_set_cooperative)
    SweepIO io;
    PathSave saves[4];
    cudaStream_t save_stream = nullptr;
    // device-side adaptation of the pCN memory inside dmcmc_run_sweeps (src/adaptation.jl:82-149): per chequerboard
    // pattern a rho vector [J] and accept / proposal counters per block [2][J]
    bool adapt_on = false;
    dmcmc_adapt adapt{};
    int adapt_spi = 1;
    DevBuf<double> d_rho_pat[2];
    DevBuf<unsigned long long> d_cnt_pat[2];
    // block-sharded recording: NCCL communicator of the ranks that share it, and its message buffers
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    DevBuf<double> cb_send, cb_recv, d_gl_llp, d_gl_ll;
    // halo exchange over peer memory (dmcmc_comm_enable_p2p): one IPC-shared region per rank = [flags 256 B | mailbox of
    // messages going left | mailbox of messages going right]; flags: data_seq[2] (written by the sender), ack_seq[2]
    // (written by the receiver into the SENDER's region)
    bool p2p_on = false;
    size_t p2p_cap = 0;                        // bytes per mailbox
    char *p2p_region = nullptr;                // mine
    char *p2p_peer[2] = {nullptr, nullptr};    // left (rank - 1) and right (rank + 1) neighbour's region, mapped here
    unsigned long long p2p_sent[2] = {0, 0}, p2p_recvd[2] = {0, 0};   // per direction (0: leftwards, 1: rightwards)
    unsigned int *p2p_counter = nullptr;       // device: [4] CTA counters (push / pull per direction)
    int *p2p_err = nullptr, *p2p_err_dev = nullptr;  // host-mapped time-out flag
    // exchange fused into the sweeps (dmcmc_impute_exchange_async): a message the neighbour pushes at the end of ITS sweep
    // is taken out of my mailbox by my next fused sweep, or by flush_pending_pull before anything else touches the paths
    bool pull_pending = false;
    int pull_dir = 0, pull_j0 = 0, pull_j1 = 0;
    size_t pull_lead = 0;
    HaloFuse hf_next{};        // what the next imputation launch fuses (blk < 0: nothing)
    bool hf_armed = false;
    DevBuf<int32_t> d_gl_ok, d_gidx;
};

// a message pushed by the neighbour's fused sweep waits in my mailbox: take it out (stand-alone kernel) before anything
// but my own next fused sweep reads or writes the path containers
static int flush_pending_pull(dmcmc_ctx *ctx);

namespace {

int fail(dmcmc_ctx *c, const std::string &msg) {
    if (c) c->err = msg;
    else g_create_error = msg;
    return -1;
}

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return fail(ctx, std::string(#call) + ": " + cudaGetErrorString(e__));      \
    } while (0)

#define REQUIRE(cond, msg)                 \
    do {                                   \
        if (!(cond)) return fail(ctx, msg); \
    } while (0)

template <class T>
int upload(dmcmc_ctx *ctx, DevBuf<T> &b, const T *src, size_t n) {
    CK(b.ensure(n));
    if (n) CK(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
template <class T>
int upload(dmcmc_ctx *ctx, DevBuf<T> &b, const std::vector<T> &v) {
    return upload(ctx, b, v.data(), v.size());
}

int sync(dmcmc_ctx *ctx) {
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// LU helpers for the (tiny) observation algebra on the host
bool lu_factor(int n, std::vector<double> &A, std::vector<int> &piv, double &lad) {
    lad = 0.0;
    piv.assign(n, 0);
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = std::fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(A[i * n + k]) > best) { best = std::fabs(A[i * n + k]); p = i; }
        piv[k] = p;
        if (best == 0.0) return false;
        if (p != k)
            for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[p * n + j]);
        lad += std::log(std::fabs(A[k * n + k]));
        const double inv = 1.0 / A[k * n + k];
        for (int i = k + 1; i < n; ++i) {
            const double l = A[i * n + k] * inv;
            A[i * n + k] = l;
            for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
        }
    }
    return true;
}
void lu_solve(int n, const std::vector<double> &LU, const std::vector<int> &piv, double *B, int nrhs) {
    for (int k = 0; k < n; ++k) {
        const int p = piv[k];
        if (p != k)
            for (int j = 0; j < nrhs; ++j) std::swap(B[k * nrhs + j], B[p * nrhs + j]);
        for (int i = k + 1; i < n; ++i) {
            const double l = LU[i * n + k];
            for (int j = 0; j < nrhs; ++j) B[i * nrhs + j] -= l * B[k * nrhs + j];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        const double inv = 1.0 / LU[k * n + k];
        for (int j = 0; j < nrhs; ++j) {
            double s = B[k * nrhs + j];
            for (int i = k + 1; i < n; ++i) s -= LU[k * n + i] * B[i * nrhs + j];
            B[k * nrhs + j] = s * inv;
        }
    }
}

// H += L'S^-1 L, F += L'S^-1 v, c += v'S^-1 v/2 + log det(2 pi S)/2 at every observation
int compute_obs_terms(dmcmc_ctx *ctx) {
    const int d = ctx->d, p = ctx->p, J = ctx->J;
    ctx->Hobs.assign((size_t)J * d * d, 0.0);
    ctx->Fobs.assign((size_t)J * d, 0.0);
    ctx->cobs.assign(J, 0.0);
    std::vector<double> S, SiL, Siv;
    std::vector<int> piv;
    for (int j = 0; j < J; ++j) {
        const double *L = &ctx->L[(size_t)j * p * d], *v = &ctx->v[(size_t)j * p];
        S.assign(&ctx->Sigma[(size_t)j * p * p], &ctx->Sigma[(size_t)j * p * p] + p * p);
        double lad;
        REQUIRE(lu_factor(p, S, piv, lad), "dmcmc_set_obs: singular observation covariance");
        SiL.assign(L, L + p * d);
        lu_solve(p, S, piv, SiL.data(), d);
        Siv.assign(v, v + p);
        lu_solve(p, S, piv, Siv.data(), 1);
        double *H = &ctx->Hobs[(size_t)j * d * d], *F = &ctx->Fobs[(size_t)j * d];
        for (int i = 0; i < d; ++i) {
            for (int k = 0; k < d; ++k) {
                double s = 0.0;
                for (int q = 0; q < p; ++q) s += L[q * d + i] * SiL[q * d + k];
                H[i * d + k] = s;
            }
            double s = 0.0;
            for (int q = 0; q < p; ++q) s += L[q * d + i] * Siv[q];
            F[i] = s;
        }
        for (int i = 0; i < d; ++i)
            for (int k = i + 1; k < d; ++k) {
                const double s = 0.5 * (H[i * d + k] + H[k * d + i]);
                H[i * d + k] = H[k * d + i] = s;
            }
        double vs = 0.0;
        for (int q = 0; q < p; ++q) vs += v[q] * Siv[q];
        ctx->cobs[j] = 0.5 * vs + 0.5 * (p * std::log(2.0 * M_PI) + lad);
    }
    if (upload(ctx, ctx->d_Hobs, ctx->Hobs)) return -1;
    if (upload(ctx, ctx->d_Fobs, ctx->Fobs)) return -1;
    if (upload(ctx, ctx->d_cobs, ctx->cobs)) return -1;
    return 0;
}

int build_aux(dmcmc_ctx *ctx, int slot) {
    Law &law = ctx->law[slot];
    const int d = ctx->d, m = ctx->m, J = ctx->J;
    if (!law.custom_aux && ctx->d_v.p) {
        REQUIRE(ctx->have_obs, "auxiliary laws need the observations: call dmcmc_set_obs first");
        CK(law.d_auxB.ensure((size_t)J * d * d));
        CK(law.d_auxbeta.ensure((size_t)J * d));
        CK(law.d_auxatil.ensure((size_t)J * d * d));
        Theta th{};
        for (int i = 0; i < DMCMC_MAX_THETA; ++i) th.v[i] = law.theta[i];
        CK(launch_aux_build(ctx->cfg.model, d, m, ctx->p, J, th, ctx->d_v.p, law.d_auxB.p, law.d_auxbeta.p, law.d_auxatil.p, ctx->stream));
        law.host_aux_valid = false;
        for (auto &lay : ctx->layout) lay.tab[slot].valid = lay.tab[slot].fvalid = false;
        return 0;
    }
    if (!law.custom_aux) {
        REQUIRE(ctx->have_obs, "auxiliary laws need the observations: call dmcmc_set_obs first");
        law.auxB.assign((size_t)J * d * d, 0.0);
        law.auxbeta.assign((size_t)J * d, 0.0);
        law.auxsig.assign((size_t)J * d * m, 0.0);
        for (int j = 0; j < J; ++j)
            aux_build_host(ctx->cfg.model, d, m, law.theta, &ctx->v[(size_t)j * ctx->p], ctx->p,
                           &law.auxB[(size_t)j * d * d], &law.auxbeta[(size_t)j * d],
                           &law.auxsig[(size_t)j * d * m]);
    }
    law.auxatil.assign((size_t)J * d * d, 0.0);
    for (int j = 0; j < J; ++j)
        for (int i = 0; i < d; ++i)
            for (int k = 0; k < d; ++k) {
                double s = 0.0;
                for (int c = 0; c < m; ++c)
                    s += law.auxsig[((size_t)j * d + i) * m + c] * law.auxsig[((size_t)j * d + k) * m + c];
                law.auxatil[((size_t)j * d + i) * d + k] = s;
            }
    if (upload(ctx, law.d_auxB, law.auxB)) return -1;
    if (upload(ctx, law.d_auxbeta, law.auxbeta)) return -1;
    if (upload(ctx, law.d_auxatil, law.auxatil)) return -1;
    law.host_aux_valid = true;
    for (auto &lay : ctx->layout) lay.tab[slot].valid = lay.tab[slot].fvalid = false;
    return 0;
}

int ensure_tables_alloc(dmcmc_ctx *ctx, Layout &lay, int slot) {
    Tables &tb = lay.tab[slot];
    const int d = ctx->d;
    const size_t rec = rec_size(d, lay.has_psi);
    CK(tb.table.ensure((size_t)ctx->TT * TILE * rec));
    CK(tb.blk_c0.ensure(lay.nblk));
    CK(tb.blk_g.ensure((size_t)lay.nblk * d));
    CK(tb.blk_Qc.ensure((size_t)lay.nblk * d * d));
    // per-grid-point copies exist for dmcmc_download_tables / parity runs; skipped when they would
    // not be small (Lorenz-96 at full size: 5 GB each)
    if ((size_t)ctx->P * d * d * sizeof(double) <= ((size_t)512 << 20)) {
        CK(tb.ptH.ensure((size_t)ctx->P * d * d));
        CK(tb.ptPsi.ensure((size_t)ctx->P * d * d));
    }
    CK(tb.ptF0.ensure((size_t)ctx->P * d));
    CK(tb.ptc0.ensure((size_t)ctx->P));
    return 0;
}

int run_backward_filter(dmcmc_ctx *ctx, int slot) {
    REQUIRE(ctx->lay_cur >= 0, "no layout set");
    REQUIRE(ctx->have_obs && ctx->have_grid, "grid/observations missing");
    REQUIRE(ctx->law[slot].have_theta || ctx->law[slot].custom_aux, "law has no parameters");
    Layout &lay = ctx->layout[ctx->lay_cur];
    if (ensure_tables_alloc(ctx, lay, slot)) return -1;
    Tables &tb = lay.tab[slot];
    Law &law = ctx->law[slot];
    BFParams b{};
    b.D = ctx->d; b.J = ctx->J; b.nblk = lay.nblk;
    b.blk_i1 = lay.d_i1.p; b.blk_i2 = lay.d_i2.p; b.blk_pinned = lay.d_pinned.p;
    b.iv_N = ctx->d_N.p; b.iv_tile_off = ctx->d_tile_off.p; b.iv_pt_off = ctx->d_pt_off.p;
    b.t = ctx->d_t.p;
    b.auxB = law.d_auxB.p; b.auxbeta = law.d_auxbeta.p; b.auxatil = law.d_auxatil.p;
    b.Hobs = ctx->d_Hobs.p; b.Fobs = ctx->d_Fobs.p; b.cobs = ctx->d_cobs.p;
    b.pin_eps = ctx->cfg.pin_eps;
    b.has_psi = lay.has_psi; b.rec = rec_size(ctx->d, lay.has_psi);
    b.table = tb.table.p;
    b.ptH = tb.ptH.p; b.ptF0 = tb.ptF0.p; b.ptPsi = tb.ptPsi.p; b.ptc0 = tb.ptc0.p;
    b.mean_N = ctx->J > 0 ? (int32_t)(ctx->st_off[ctx->J] / ctx->J) : 0;
    b.blk_c0 = tb.blk_c0.p; b.blk_g = tb.blk_g.p; b.blk_Qc = tb.blk_Qc.p;
    CK(launch_backward_filter(b, ctx->stream));
    tb.valid = true;
    tb.fvalid = false;
    return 0;
}

int fill_params(dmcmc_ctx *ctx, SolveParams &sp, int slot) {
    REQUIRE(ctx->lay_cur >= 0, "no layout set");
    REQUIRE(ctx->have_grid && ctx->have_start, "grid / starting points missing");
    Layout &lay = ctx->layout[ctx->lay_cur];
    Tables &tb = lay.tab[slot];
    if (!tb.valid && run_backward_filter(ctx, slot)) return -1;  // tables follow the layout lazily
    Law &law = ctx->law[slot];
    sp = SolveParams{};
    sp.hf.blk = -1;
    sp.C = ctx->C; sp.nblk = lay.nblk; sp.J = ctx->J; sp.D = ctx->d; sp.chain_offset = ctx->cfg.chain_offset;
    sp.interval_offset = ctx->cfg.interval_offset;
    sp.blk_active = lay.has_mask ? lay.d_active.p : nullptr;
    sp.TT = ctx->TT;
    sp.blk_i1 = lay.d_i1.p; sp.blk_i2 = lay.d_i2.p; sp.blk_pinned = lay.d_pinned.p;
    sp.iv_N = ctx->d_N.p; sp.iv_tile_off = ctx->d_tile_off.p; sp.iv_first = ctx->d_first.p;
    sp.iv_rec = ctx->d_rec.p;
    sp.iv_xoff = ctx->d_xoff.p; sp.iv_xs = ctx->d_xs.p;
    sp.rho = ctx->d_rho.p;
    sp.auxB = law.d_auxB.p; sp.auxbeta = law.d_auxbeta.p; sp.auxatil = law.d_auxatil.p;
    sp.table = sp.table_raw = tb.table.p; sp.blk_c0 = tb.blk_c0.p; sp.blk_g = tb.blk_g.p; sp.blk_Qc = tb.blk_Qc.p;
    sp.has_psi = lay.has_psi;
    for (int b = 0; b < 2; ++b) { sp.X[b] = ctx->d_X[b].p; sp.W[b] = ctx->d_W[b].p; }
    sp.selX_in = ctx->d_selX[ctx->sel_cur].p; sp.selW_in = ctx->d_selW[ctx->sel_cur].p;
    sp.selX_out = ctx->d_selX[1 - ctx->sel_cur].p; sp.selW_out = ctx->d_selW[1 - ctx->sel_cur].p;
    sp.x0 = ctx->d_x0.p; sp.ll = ctx->d_ll.p;
    sp.seed = ctx->cfg.seed;
    sp.out_stride = lay.nblk;
    sp.reinvert = 0;
    sp.store_w = 1;
    sp.fast_ok = (!law.custom_aux && ctx->fast_step) ? 1 : 0;
    sp.coop = ctx->coop;
    sp.mean_N = ctx->J > 0 ? (int32_t)(ctx->st_off[ctx->J] / ctx->J) : 0;
    if (ctx->cfg.model == DMCMC_MODEL_FHN && sp.fast_ok) {  // fused step streams the derived coefficient table
        if (!tb.fvalid) {
            Theta th{};
            for (int i = 0; i < DMCMC_MAX_THETA; ++i) th.v[i] = law.theta[i];
            CK(tb.ftable.ensure(tb.table.n));
            CK(launch_fhn_derive(tb.table.p, tb.ftable.p, (long long)ctx->TT * TILE, lay.has_psi, th, ctx->stream));
            tb.fvalid = true;
        }
        sp.table = tb.ftable.p;
    }
    for (int i = 0; i < DMCMC_MAX_THETA; ++i) sp.th.v[i] = law.theta[i];
    return 0;
}

int timed_launch(dmcmc_ctx *ctx, int mode, const SolveParams &sp) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool timed = ctx->timing && ctx->time_every > 0 && (ctx->launches % ctx->time_every) == 0;
    ++ctx->launches;
    if (timed) {
        if (ctx->ev_used >= 65536) ctx->ev_used = 0;  // nobody is reading the timings
        if (ctx->ev_used + 2 > ctx->ev.size()) {
            for (int i = 0; i < 2; ++i) {
                cudaEvent_t e;
                CK(cudaEventCreate(&e));
                ctx->ev.push_back(e);
            }
        }
        e0 = ctx->ev[ctx->ev_used];
        e1 = ctx->ev[ctx->ev_used + 1];
        if (ctx->ev_mode.size() < ctx->ev.size() / 2) ctx->ev_mode.resize(ctx->ev.size() / 2);
        ctx->ev_mode[ctx->ev_used / 2] = mode;
        ctx->ev_used += 2;
        CK(cudaEventRecord(e0, ctx->stream));
    }
    if (ctx->hf_armed && mode == MODE_IMPUTE) {  // dmcmc_impute_exchange_async: this launch carries the exchange
        SolveParams spf = sp;
        spf.hf = ctx->hf_next;
        ctx->hf_armed = false;
        CK(launch_solve(ctx->cfg.model, mode, spf, ctx->threads, ctx->stream));
    } else {
        if (flush_pending_pull(ctx)) return -1;
        CK(launch_solve(ctx->cfg.model, mode, sp, ctx->threads, ctx->stream));
    }
    if (timed) CK(cudaEventRecord(e1, ctx->stream));
    return 0;
}

// [C][S][m] host normals -> [m][TT][C][8] device tiles
int stage_noise(dmcmc_ctx *ctx, const double *Z) {
    const int C = ctx->C, m = ctx->m, J = ctx->J;
    const size_t n = (size_t)m * ctx->TT * C * TILE;
    std::vector<double> buf(n, 0.0);
    for (int c = 0; c < C; ++c)
        for (int j = 0; j < J; ++j)
            for (int k = 0; k < ctx->N[j]; ++k) {
                const size_t tile = (size_t)ctx->tile_off[j] + (k >> 3);
                for (int w = 0; w < m; ++w)
                    buf[w_off(tile, c, w, C, m) + (k & 7)] =
                        Z[((size_t)c * ctx->S + ctx->st_off[j] + k) * m + w];
            }
    CK(ctx->d_Z.ensure(n));
    CK(cudaMemcpy(ctx->d_Z.p, buf.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

int alloc_state(dmcmc_ctx *ctx) {
    const size_t nx = (size_t)ctx->xoff[ctx->J], nw = (size_t)ctx->m * ctx->TT * ctx->C * TILE;
    const size_t ns = (size_t)ctx->J * ctx->C;
    for (int b = 0; b < 2; ++b) {
        CK(ctx->d_X[b].ensure(nx));
        CK(cudaMemsetAsync(ctx->d_X[b].p, 0, nx * sizeof(double), ctx->stream));
        if (ctx->cfg.store_noise) {
            CK(ctx->d_W[b].ensure(nw));
            CK(cudaMemsetAsync(ctx->d_W[b].p, 0, nw * sizeof(double), ctx->stream));
        }
        CK(ctx->d_selX[b].ensure(ns));
        CK(ctx->d_selW[b].ensure(ns));
        CK(cudaMemsetAsync(ctx->d_selX[b].p, 0, ns, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_selW[b].p, 0, ns, ctx->stream));
    }
    ctx->sel_cur = 0;
    CK(ctx->d_counts.ensure((size_t)2 * ctx->J));
    CK(cudaMemsetAsync(ctx->d_counts.p, 0, ctx->d_counts.bytes(), ctx->stream));
    return 0;
}

int ensure_unit_buffers(dmcmc_ctx *ctx) {
    const size_t n = (size_t)ctx->C * ctx->layout[ctx->lay_cur].nblk;
    CK(ctx->d_out_llprop.ensure(n));
    CK(ctx->d_out_ll.ensure(n));
    CK(ctx->d_out_acc.ensure(n));
    CK(ctx->d_E.ensure(n));
    CK(ctx->d_mask.ensure(n));
    return 0;
}

int select_layout(dmcmc_ctx *ctx, int32_t nblk, const int32_t *i1, const int32_t *i2,
                  const int32_t *pinned) {
    REQUIRE(ctx->have_grid, "dmcmc_set_layout: call dmcmc_set_grid first");
    REQUIRE(nblk > 0, "dmcmc_set_layout: empty layout");
    // validate: blocks tile [0, J) in order, never straddle recordings
    int expect = 0;
    for (int b = 0; b < nblk; ++b) {
        REQUIRE(i1[b] == expect && i2[b] >= i1[b] && i2[b] < ctx->J, "dmcmc_set_layout: blocks must partition the intervals in order");
        REQUIRE(ctx->iv_rec[i1[b]] == ctx->iv_rec[i2[b]], "dmcmc_set_layout: a block straddles two recordings");
        const bool last_of_rec = (i2[b] == ctx->J - 1) || ctx->iv_first[i2[b] + 1];
        // a pinned LAST block is allowed: it is the halo of a recording sharded over GPUs (its end
        // point is conditioned on the value the right neighbour holds)
        REQUIRE(pinned[b] || last_of_rec, "dmcmc_set_layout: an interior block must be pinned");
        expect = i2[b] + 1;
    }
    REQUIRE(expect == ctx->J, "dmcmc_set_layout: blocks must cover every interval");
    for (int s = 0; s < 2; ++s) {
        Layout &l = ctx->layout[s];
        if (l.used && l.nblk == nblk && std::equal(i1, i1 + nblk, l.i1.begin()) &&
            std::equal(i2, i2 + nblk, l.i2.begin()) && std::equal(pinned, pinned + nblk, l.pinned.begin())) {
            ctx->lay_cur = s;
            l.last_used = ++ctx->lay_clock;
            return 0;
        }
    }
    int s = 0;
    if (ctx->layout[0].used && (!ctx->layout[1].used || ctx->layout[1].last_used < ctx->layout[0].last_used)) s = 1;
    Layout &l = ctx->layout[s];
    l.used = true;
    l.nblk = nblk;
    l.has_mask = l.has_mask_copy = false;
    l.i1.assign(i1, i1 + nblk);
    l.i2.assign(i2, i2 + nblk);
    l.pinned.assign(pinned, pinned + nblk);
    l.has_psi = std::any_of(pinned, pinned + nblk, [](int32_t x) { return x != 0; });
    l.iv_blk.assign(ctx->J, 0);
    for (int b = 0; b < nblk; ++b)
        for (int j = i1[b]; j <= i2[b]; ++j) l.iv_blk[j] = b;
    if (upload(ctx, l.d_i1, l.i1) || upload(ctx, l.d_i2, l.i2) || upload(ctx, l.d_pinned, l.pinned) ||
        upload(ctx, l.d_iv_blk, l.iv_blk))
        return -1;
    l.tab[0].valid = l.tab[1].valid = false;
    l.tab[0].fvalid = l.tab[1].fvalid = false;
    l.has_mask = false;
    l.last_used = ++ctx->lay_clock;
    ctx->lay_cur = s;
    return 0;
}

int resize_ll(dmcmc_ctx *ctx) {
    const size_t n = (size_t)ctx->C * ctx->layout[ctx->lay_cur].nblk;
    CK(ctx->d_ll.ensure(n));
    return 0;
}

}  // namespace

// ---- NCCL, bound at run time ---------------------------------------------------------------------------------------
// The library has no link-time dependency on NCCL: a process that already carries one (PyTorch's) is reused
// (RTLD_NOLOAD), otherwise the system libnccl.so.2 is loaded on the first dmcmc_comm_* call.
namespace {
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_m;

NcclApi *nccl_api(std::string &err) {
    std::lock_guard<std::mutex> lk(g_nccl_m);
    if (g_nccl.h) return &g_nccl;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { err = std::string("NCCL not found: ") + dlerror(); return nullptr; }
#define DMCMC_NCCL_SYM(field, name)                                             \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));     \
    if (!g_nccl.field) { err = std::string("NCCL symbol missing: ") + name; return nullptr; }
    DMCMC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    DMCMC_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    DMCMC_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    DMCMC_NCCL_SYM(Send, "ncclSend")
    DMCMC_NCCL_SYM(Recv, "ncclRecv")
    DMCMC_NCCL_SYM(AllReduce, "ncclAllReduce")
    DMCMC_NCCL_SYM(AllGather, "ncclAllGather")
    DMCMC_NCCL_SYM(GroupStart, "ncclGroupStart")
    DMCMC_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    DMCMC_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef DMCMC_NCCL_SYM
    g_nccl.h = h;
    return &g_nccl;
}

// readjust! on the device (src/adaptation.jl:82-90 with eMCMC.compute_delta / compute_eps [UPSTREAM-RECALL], the host
// mirror's AdaptationPathImputation.readjust): one thread per block of the sweep's layout; a block whose counter has
// reached adapt_every_k_steps proposals takes a step of size delta in logit space, DOWN when its acceptance rate is
// above the target, and starts counting afresh
__global__ void adapt_rho_kernel(int nblk, const int32_t *i1, const int32_t *i2, unsigned long long *acc,
                                 unsigned long long *prop, double *rho, dmcmc_adapt a, double mcmciter) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const unsigned long long np = prop[b];
    if (np < (unsigned long long)a.adapt_every_k_steps) return;
    const double rate = (double)acc[b] / (double)np;
    const double delta = a.scale / sqrt(fmax(1.0, mcmciter / (double)a.adapt_every_k_steps - a.offset));
    const double r = fmin(fmax(rho[i1[b]], a.min), a.max);
    const double step = rate > a.target_accpt_rate ? -delta : delta;
    const double lg = log(r) - log(1.0 - r) + step;
    const double nr = fmin(fmax(1.0 / (1.0 + exp(-lg)), a.min), a.max);
    for (int j = i1[b]; j <= i2[b]; ++j) rho[j] = nr;
    acc[b] = 0;
    prop[b] = 0;
}

static void p2p_release(dmcmc_ctx *ctx) {
    for (int i = 0; i < 2; ++i) {
        if (ctx->p2p_peer[i]) cudaIpcCloseMemHandle(ctx->p2p_peer[i]);
        ctx->p2p_peer[i] = nullptr;
    }
    if (ctx->p2p_region) cudaFree(ctx->p2p_region);
    if (ctx->p2p_counter) cudaFree(ctx->p2p_counter);
    if (ctx->p2p_err) cudaFreeHost(ctx->p2p_err);
    ctx->p2p_region = nullptr; ctx->p2p_counter = nullptr; ctx->p2p_err = nullptr; ctx->p2p_err_dev = nullptr;
    ctx->p2p_on = false; ctx->p2p_cap = 0;
    ctx->pull_pending = false; ctx->hf_armed = false;
    ctx->p2p_sent[0] = ctx->p2p_sent[1] = ctx->p2p_recvd[0] = ctx->p2p_recvd[1] = 0;
}

void comm_release(dmcmc_ctx *ctx) {
    p2p_release(ctx);
    if (ctx->comm && g_nccl.h) g_nccl.CommDestroy(ctx->comm);
    ctx->comm = nullptr;
    ctx->cb_send.release(); ctx->cb_recv.release(); ctx->d_gl_llp.release(); ctx->d_gl_ll.release();
    ctx->d_gl_ok.release(); ctx->d_gidx.release();
}

#define NK(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess) return fail(ctx, std::string(#call) + ": " + nc->GetErrorString(r__));   \
    } while (0)

// owned blocks of a rank scattered into vectors over the GLOBAL block index (zero elsewhere): summing the ranks'
// vectors then reproduces every block's value bit for bit (x + 0 is exact), whatever the number of ranks
__global__ void scatter_blocks_kernel(const double *llp, const double *ll, const uint8_t *ok, const int32_t *gidx, int C,
                                      int nblk, int nblk_g, double *g_llp, double *g_ll, int32_t *g_ok) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= C * nblk) return;
    const int c = u / nblk, b = u % nblk, g = gidx[b];
    if (g < 0) return;
    g_llp[(size_t)c * nblk_g + g] = llp[u];
    g_ll[(size_t)c * nblk_g + g] = ll[u];
    if (!ok[u]) atomicMin(g_ok + c, 0);
}
__global__ void fill_i32_kernel(int32_t *p, int n, int32_t v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
}  // namespace

// =================================================================================================
extern "C" {

int dmcmc_create(const dmcmc_config *cfg, dmcmc_ctx **out) {
    dmcmc_ctx *ctx = nullptr;
    if (!cfg || !out) return fail(nullptr, "dmcmc_create: null argument");
    if (cfg->abi_version != DMCMC_ABI_VERSION) return fail(nullptr, "dmcmc_create: ABI version mismatch");
    int d, m, nth;
    if (model_dims(cfg->model, cfg->d, &d, &m, &nth)) return fail(nullptr, "dmcmc_create: unknown model / unsupported dimension");
    if (cfg->n_chains <= 0) return fail(nullptr, "dmcmc_create: n_chains must be positive");
    if (!(cfg->pin_eps > 0.0)) return fail(nullptr, "dmcmc_create: pin_eps must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, std::string("dmcmc_create: no CUDA device (") + cudaGetErrorString(e) +
                                 "); this backend has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, "dmcmc_create: bad device ordinal");
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return fail(nullptr, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major != 10)
        return fail(nullptr, "dmcmc_create: kernels are built for sm_100a (B200) only; found sm_" +
                                 std::to_string(prop.major) + std::to_string(prop.minor));
    ctx = new dmcmc_ctx();
    ctx->cfg = *cfg;
    ctx->d = d; ctx->m = m; ctx->nth = nth; ctx->C = cfg->n_chains;
    if (const char *t = getenv("DMCMC_TIMING_EVERY")) ctx->time_every = std::max(0, atoi(t));
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return fail(nullptr, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    *out = ctx;
    return 0;
}

void dmcmc_destroy(dmcmc_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    cudaStreamSynchronize(ctx->stream);
    for (auto e : ctx->ev) cudaEventDestroy(e);
    if (ctx->sw0) { cudaEventDestroy(ctx->sw0); cudaEventDestroy(ctx->sw1); }
    {
        SweepIO &io = ctx->io;
        if (io.pool_started) {
            {
                std::lock_guard<std::mutex> lk(io.m);
                io.stop = true;
            }
            io.cv_job.notify_all();
            for (auto &t : io.drainer)
                if (t.joinable()) t.join();
        }
        if (io.s_in) {
            cudaStreamSynchronize(io.s_in); cudaStreamSynchronize(io.s_out);
            for (int i = 0; i < SWEEP_RING; ++i) { cudaEventDestroy(io.kernel_done[i]); cudaEventDestroy(io.copied[i]); }
            cudaStreamDestroy(io.s_in); cudaStreamDestroy(io.s_out);
        }
        io.d_hist.release(); io.d_rho.release();
        if (io.h_hist) cudaFreeHost(io.h_hist);
        if (ctx->save_stream) { cudaStreamSynchronize(ctx->save_stream); cudaStreamDestroy(ctx->save_stream); }
        for (auto &sv : ctx->saves) {
            sv.d_buf.release();
            if (sv.h_buf) cudaFreeHost(sv.h_buf);
            if (sv.snap) { cudaEventDestroy(sv.snap); cudaEventDestroy(sv.landed); }
        }
    }
    comm_release(ctx);
    auto rel_law = [](Law &l) { l.d_auxB.release(); l.d_auxbeta.release(); l.d_auxatil.release(); };
    rel_law(ctx->law[0]);
    rel_law(ctx->law[1]);
    for (auto &l : ctx->layout) {
        l.d_i1.release(); l.d_i2.release(); l.d_pinned.release(); l.d_iv_blk.release(); l.d_active.release();
        for (auto &t : l.tab) {
            t.table.release(); t.ftable.release(); t.blk_c0.release(); t.blk_g.release(); t.blk_Qc.release();
            t.ptH.release(); t.ptF0.release(); t.ptPsi.release(); t.ptc0.release();
        }
    }
    ctx->d_N.release(); ctx->d_tile_off.release(); ctx->d_first.release(); ctx->d_rec.release();
    ctx->d_xoff.release(); ctx->d_xs.release();
    ctx->d_pt_off.release(); ctx->d_st_off.release(); ctx->d_t.release(); ctx->d_rho.release(); ctx->d_Hobs.release();
    ctx->d_Fobs.release(); ctx->d_cobs.release(); ctx->d_x0.release(); ctx->d_ll.release();
    ctx->d_Z.release(); ctx->d_E.release(); ctx->d_v.release();
    for (int b = 0; b < 2; ++b) {
        ctx->d_X[b].release(); ctx->d_W[b].release(); ctx->d_selX[b].release(); ctx->d_selW[b].release();
    }
    ctx->d_mask.release(); ctx->d_out_acc.release(); ctx->d_out_llprop.release();
    ctx->d_out_ll.release(); ctx->d_param_llprop.release(); ctx->d_counts.release();
    for (int i = 0; i < 2; ++i) { ctx->d_rho_pat[i].release(); ctx->d_cnt_pat[i].release(); }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *dmcmc_last_error(const dmcmc_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int dmcmc_set_grid(dmcmc_ctx *ctx, int32_t n_recordings, const int32_t *rec_nobs,
                   const int32_t *n_steps, const double *t) {
    REQUIRE(ctx && rec_nobs && n_steps && t && n_recordings > 0, "dmcmc_set_grid: bad argument");
    CK(cudaSetDevice(ctx->cfg.device));
    ctx->R = n_recordings;
    ctx->rec_nobs.assign(rec_nobs, rec_nobs + n_recordings);
    int J = 0;
    for (int r = 0; r < n_recordings; ++r) {
        REQUIRE(rec_nobs[r] > 0, "dmcmc_set_grid: a recording has no observations");
        J += rec_nobs[r];
    }
    ctx->J = J;
    ctx->N.assign(n_steps, n_steps + J);
    ctx->pt_off.assign(J + 1, 0);
    ctx->st_off.assign(J + 1, 0);
    ctx->tile_off.assign(J + 1, 0);
    ctx->iv_first.assign(J, 0);
    ctx->iv_rec.assign(J, 0);
    ctx->xoff.assign(J + 1, 0);
    ctx->xs.assign(J, 0);
    int j = 0;
    for (int r = 0; r < n_recordings; ++r)
        for (int i = 0; i < rec_nobs[r]; ++i, ++j) {
            ctx->iv_first[j] = (i == 0);
            ctx->iv_rec[j] = r;
        }
    for (j = 0; j < J; ++j) {
        REQUIRE(n_steps[j] > 0 && n_steps[j] < (1 << 24), "dmcmc_set_grid: n_steps out of range");
        ctx->pt_off[j + 1] = ctx->pt_off[j] + n_steps[j] + 1;
        ctx->st_off[j + 1] = ctx->st_off[j] + n_steps[j];
        ctx->tile_off[j + 1] = ctx->tile_off[j] + (n_steps[j] + TILE - 1) / TILE;
        ctx->xs[j] = x_stride(n_steps[j], ctx->d);
        ctx->xoff[j + 1] = ctx->xoff[j] + (int64_t)ctx->C * ctx->xs[j];
    }
    ctx->P = ctx->pt_off[J];
    ctx->S = ctx->st_off[J];
    ctx->TT = ctx->tile_off[J];
    ctx->t.assign(t, t + ctx->P);
    for (j = 0; j < J; ++j)
        for (int k = 0; k < n_steps[j]; ++k)
            REQUIRE(t[ctx->pt_off[j] + k + 1] > t[ctx->pt_off[j] + k], "dmcmc_set_grid: grid times must increase");
    if (upload(ctx, ctx->d_N, ctx->N) || upload(ctx, ctx->d_tile_off, ctx->tile_off) ||
        upload(ctx, ctx->d_first, ctx->iv_first) || upload(ctx, ctx->d_rec, ctx->iv_rec) ||
        upload(ctx, ctx->d_pt_off, ctx->pt_off) || upload(ctx, ctx->d_st_off, ctx->st_off) ||
        upload(ctx, ctx->d_xoff, ctx->xoff) || upload(ctx, ctx->d_xs, ctx->xs) || upload(ctx, ctx->d_t, ctx->t))
        return -1;
    ctx->rho.assign(J, 0.0);
    if (upload(ctx, ctx->d_rho, ctx->rho)) return -1;
    ctx->have_grid = true;
    ctx->have_obs = ctx->have_start = false;
    ctx->lay_cur = -1;
    ctx->layout[0].used = ctx->layout[1].used = false;
    if (alloc_state(ctx)) return -1;
    return sync(ctx);
}

int dmcmc_set_obs(dmcmc_ctx *ctx, int32_t p, const double *L, const double *Sigma, const double *v) {
    REQUIRE(ctx && L && Sigma && v && p > 0, "dmcmc_set_obs: bad argument");
    REQUIRE(ctx->have_grid, "dmcmc_set_obs: call dmcmc_set_grid first");
    CK(cudaSetDevice(ctx->cfg.device));
    ctx->p = p;
    const size_t J = ctx->J;
    ctx->L.assign(L, L + J * p * ctx->d);
    ctx->Sigma.assign(Sigma, Sigma + J * p * p);
    ctx->v.assign(v, v + J * p);
    if (compute_obs_terms(ctx)) return -1;
    if (upload(ctx, ctx->d_v, ctx->v)) return -1;
    ctx->have_obs = true;
    for (int s = 0; s < 2; ++s)
        if (ctx->law[s].have_theta && build_aux(ctx, s)) return -1;
    return sync(ctx);
}

int dmcmc_set_start(dmcmc_ctx *ctx, const double *x0) {
    REQUIRE(ctx && x0, "dmcmc_set_start: bad argument");
    REQUIRE(ctx->have_grid, "dmcmc_set_start: call dmcmc_set_grid first");
    CK(cudaSetDevice(ctx->cfg.device));
    const int C = ctx->C, R = ctx->R, d = ctx->d;
    ctx->x0.assign(x0, x0 + (size_t)C * R * d);
    std::vector<double> tr((size_t)R * d * C);
    for (int c = 0; c < C; ++c)
        for (int r = 0; r < R; ++r)
            for (int i = 0; i < d; ++i) tr[((size_t)r * d + i) * C + c] = x0[((size_t)c * R + r) * d + i];
    if (upload(ctx, ctx->d_x0, tr)) return -1;
    ctx->have_start = true;
    ctx->w_valid = false;
    return sync(ctx);
}

int dmcmc_set_theta(dmcmc_ctx *ctx, const double *theta, int32_t ntheta) {
    REQUIRE(ctx && theta, "dmcmc_set_theta: bad argument");
    REQUIRE(ntheta == ctx->nth, "dmcmc_set_theta: wrong number of parameters for this model");
    CK(cudaSetDevice(ctx->cfg.device));
    Law &law = ctx->law[ctx->law_cur];
    std::copy(theta, theta + ntheta, law.theta);
    law.have_theta = true;
    if (ctx->have_obs && build_aux(ctx, ctx->law_cur)) return -1;
    return sync(ctx);
}

int dmcmc_set_aux(dmcmc_ctx *ctx, int32_t which, const double *B, const double *beta, const double *sigma_aux) {
    REQUIRE(ctx && B && beta && sigma_aux && (which == 0 || which == 1), "dmcmc_set_aux: bad argument");
    REQUIRE(ctx->have_grid, "dmcmc_set_aux: call dmcmc_set_grid first");
    CK(cudaSetDevice(ctx->cfg.device));
    const int slot = ctx->law_cur ^ which;
    Law &law = ctx->law[slot];
    const size_t J = ctx->J, d = ctx->d, m = ctx->m;
    law.auxB.assign(B, B + J * d * d);
    law.auxbeta.assign(beta, beta + J * d);
    law.auxsig.assign(sigma_aux, sigma_aux + J * d * m);
    law.custom_aux = true;
    if (build_aux(ctx, slot)) return -1;
    return sync(ctx);
}

int dmcmc_chequerboard_layout(int32_t R, const int32_t *rec_nobs, int32_t half_len, int32_t pattern,
                              int32_t *i1, int32_t *i2, int32_t *pinned) {
    if (!rec_nobs || !i1 || !i2 || !pinned || R <= 0) return -1;
    int nb = 0, base = 0;
    for (int r = 0; r < R; ++r) {
        const int n = rec_nobs[r];
        if (half_len <= 0) {
            i1[nb] = base; i2[nb] = base + n - 1; pinned[nb] = 0; ++nb;
        } else {
            int start = 0, len = pattern ? half_len : 2 * half_len;
            while (start < n) {
                const int end = std::min(start + len, n);
                i1[nb] = base + start; i2[nb] = base + end - 1; pinned[nb] = end < n; ++nb;
                start = end;
                len = 2 * half_len;
            }
        }
        base += n;
    }
    return nb;
}

int dmcmc_set_layout(dmcmc_ctx *ctx, int32_t nblk, const int32_t *i1, const int32_t *i2, const int32_t *pinned) {
    REQUIRE(ctx && i1 && i2 && pinned, "dmcmc_set_layout: bad argument");
    CK(cudaSetDevice(ctx->cfg.device));
    const int before = ctx->lay_cur;
    if (select_layout(ctx, nblk, i1, i2, pinned)) return -1;
    if (resize_ll(ctx)) return -1;
    if (before != ctx->lay_cur) ctx->w_valid = false;
    return 0;  // stream-ordered: no host synchronisation
}

int dmcmc_set_block_mask(dmcmc_ctx *ctx, const int32_t *active) {
    REQUIRE(ctx && ctx->lay_cur >= 0, "dmcmc_set_block_mask: no layout set");
    CK(cudaSetDevice(ctx->cfg.device));
    Layout &l = ctx->layout[ctx->lay_cur];
    if (!active) {
        l.has_mask = false;
        return 0;
    }
    std::vector<int32_t> a(active, active + l.nblk);
    if (!(l.has_mask_copy && a == l.active_host)) {  // masks alternate with the layouts: upload only when changed
        if (upload(ctx, l.d_active, a)) return -1;
        l.active_host = a;
        l.has_mask_copy = true;
    }
    l.has_mask = true;
    return 0;  // stream-ordered: no host synchronisation
}

static int segment_copy(dmcmc_ctx *ctx, int32_t j0, int32_t j1, double *X, bool up, int c0 = 0, int nc = 0) {
    REQUIRE(ctx && X && j0 >= 0 && j1 > j0 && j1 <= ctx->J, "segment: bad interval range");
    REQUIRE(ctx->have_grid, "segment: call dmcmc_set_grid first");
    CK(cudaSetDevice(ctx->cfg.device));
    const int steps = ctx->st_off[j1] - ctx->st_off[j0];
    const size_t n = (size_t)(nc > 0 ? nc : ctx->C) * steps * ctx->d;
    DevBuf<double> pk;
    CK(pk.ensure(n));
    if (up) CK(cudaMemcpyAsync(pk.p, X, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SegmentParams g{};
    g.C = ctx->C; g.D = ctx->d; g.j0 = j0; g.j1 = j1; g.steps = steps; g.TT = ctx->TT;
    g.iv_N = ctx->d_N.p; g.iv_tile_off = ctx->d_tile_off.p; g.iv_st_off = ctx->d_st_off.p;
    g.iv_xoff = ctx->d_xoff.p; g.iv_xs = ctx->d_xs.p;
    g.X[0] = ctx->d_X[0].p; g.X[1] = ctx->d_X[1].p;
    g.selX = ctx->d_selX[ctx->sel_cur].p;
    g.packed = pk.p; g.upload = up; g.pk_stride = (int64_t)steps * ctx->d; g.c0 = c0; g.nc = nc;
    if (flush_pending_pull(ctx)) return -1;  // a fused exchange's message may still sit in the mailbox
    CK(launch_segment(g, ctx->stream));
    if (!up) CK(cudaMemcpyAsync(X, pk.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (sync(ctx)) return -1;
    pk.release();
    if (up) ctx->w_valid = false;
    return 0;
}

// device-side variant: `dev` is a caller-owned DEVICE buffer, element (chain, step, comp) at
// dev[chain * chain_stride + step * d + comp]; enqueued on the context's stream, no host sync
int dmcmc_segment_device(dmcmc_ctx *ctx, int32_t j0, int32_t j1, double *dev, int64_t chain_stride, int32_t upload_) {
    REQUIRE(ctx && dev && j0 >= 0 && j1 > j0 && j1 <= ctx->J, "dmcmc_segment_device: bad interval range");
    REQUIRE(ctx->have_grid, "dmcmc_segment_device: call dmcmc_set_grid first");
    CK(cudaSetDevice(ctx->cfg.device));
    const int steps = ctx->st_off[j1] - ctx->st_off[j0];
    REQUIRE(chain_stride >= (int64_t)steps * ctx->d, "dmcmc_segment_device: chain_stride too small");
    SegmentParams g{};
    g.C = ctx->C; g.D = ctx->d; g.j0 = j0; g.j1 = j1; g.steps = steps; g.TT = ctx->TT;
    g.iv_N = ctx->d_N.p; g.iv_tile_off = ctx->d_tile_off.p; g.iv_st_off = ctx->d_st_off.p;
    g.iv_xoff = ctx->d_xoff.p; g.iv_xs = ctx->d_xs.p;
    g.X[0] = ctx->d_X[0].p; g.X[1] = ctx->d_X[1].p;
    g.selX = ctx->d_selX[ctx->sel_cur].p;
    g.packed = dev; g.upload = upload_; g.pk_stride = chain_stride;
    if (flush_pending_pull(ctx)) return -1;  // a fused exchange's message may still sit in the mailbox
    CK(launch_segment(g, ctx->stream));
    if (upload_) ctx->w_valid = false;
    return 0;
}

// starting points from a DEVICE buffer: x0 of (chain, recording r) at dev[chain * chain_stride + r * d + comp]
int dmcmc_set_start_device(dmcmc_ctx *ctx, const double *dev, int64_t chain_stride) {
    REQUIRE(ctx && dev && ctx->have_start, "dmcmc_set_start_device: context not initialised");
    CK(cudaSetDevice(ctx->cfg.device));
    if (flush_pending_pull(ctx)) return -1;  // a fused exchange's message may still sit in the mailbox
    CK(launch_set_start(ctx->d_x0.p, dev, chain_stride, ctx->C, ctx->R, ctx->d, ctx->stream));
    ctx->w_valid = false;
    return 0;
}

void *dmcmc_stream(dmcmc_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

// one imputation update with Philox noise, current layout and block mask: enqueued, not awaited;
// results stay on the device (ll, selectors, accept counters)
int dmcmc_impute_async(dmcmc_ctx *ctx, uint32_t iter) {
    REQUIRE(ctx, "dmcmc_impute_async: null context");
    CK(cudaSetDevice(ctx->cfg.device));
    Layout &lay0 = ctx->layout[ctx->lay_cur];
    if (!lay0.tab[ctx->law_cur].valid && run_backward_filter(ctx, ctx->law_cur)) return -1;
    SolveParams sp;
    if (fill_params(ctx, sp, ctx->law_cur)) return -1;
    if (ensure_unit_buffers(ctx)) return -1;
    sp.iter = iter;
    sp.reinvert = (!ctx->w_valid || !ctx->cfg.store_noise) ? 1 : 0;
    sp.store_w = !sp.reinvert;
    sp.out_llprop = ctx->d_out_llprop.p; sp.out_ll = ctx->d_out_ll.p; sp.out_acc = ctx->d_out_acc.p;
    sp.acc_count = ctx->d_counts.p;
    sp.prop_count = ctx->d_counts.p + ctx->J;
    if (timed_launch(ctx, MODE_IMPUTE, sp)) return -1;
    ctx->sel_cur ^= 1;
    return 0;
}

int dmcmc_download_segment(dmcmc_ctx *ctx, int32_t j0, int32_t j1, double *X) {
    return segment_copy(ctx, j0, j1, X, false);
}

int dmcmc_upload_segment(dmcmc_ctx *ctx, int32_t j0, int32_t j1, const double *X) {
    return segment_copy(ctx, j0, j1, const_cast<double *>(X), true);
}

int dmcmc_set_rho(dmcmc_ctx *ctx, const double *rho) {
    REQUIRE(ctx && rho, "dmcmc_set_rho: bad argument");
    REQUIRE(ctx->have_grid, "dmcmc_set_rho: call dmcmc_set_grid first");
    CK(cudaSetDevice(ctx->cfg.device));
    for (int j = 0; j < ctx->J; ++j)
        REQUIRE(rho[j] >= 0.0 && rho[j] <= 1.0, "dmcmc_set_rho: rho must lie in [0, 1]");
    ctx->rho.assign(rho, rho + ctx->J);
    if (upload(ctx, ctx->d_rho, ctx->rho)) return -1;
    return sync(ctx);
}

int dmcmc_backward_filter(dmcmc_ctx *ctx, int32_t which) {
    REQUIRE(ctx && (which == 0 || which == 1), "dmcmc_backward_filter: bad argument");
    CK(cudaSetDevice(ctx->cfg.device));
    if (run_backward_filter(ctx, ctx->law_cur ^ which)) return -1;
    return sync(ctx);
}

int dmcmc_upload_tables(dmcmc_ctx *ctx, int32_t which, const double *H, const double *F0,
                        const double *Psi, const double *c0, const double *blk_g, const double *blk_Qc) {
    REQUIRE(ctx && H && F0 && c0 && (which == 0 || which == 1), "dmcmc_upload_tables: bad argument");
    REQUIRE(ctx->lay_cur >= 0, "dmcmc_upload_tables: no layout set");
    CK(cudaSetDevice(ctx->cfg.device));
    Layout &lay = ctx->layout[ctx->lay_cur];
    REQUIRE(!lay.has_psi || (Psi && blk_g && blk_Qc), "dmcmc_upload_tables: pinned layout needs Psi, g, Qc");
    const int slot = ctx->law_cur ^ which;
    if (ensure_tables_alloc(ctx, lay, slot)) return -1;
    Tables &tb = lay.tab[slot];
    const int d = ctx->d, rec = rec_size(d, lay.has_psi);
    std::vector<double> tab((size_t)ctx->TT * TILE * rec, 0.0);
    for (int j = 0; j < ctx->J; ++j)
        for (int k = 0; k < ctx->N[j]; ++k) {
            const size_t pt = (size_t)ctx->pt_off[j] + k;
            double *r = &tab[((size_t)ctx->tile_off[j] * TILE + k) * rec];
            const double dt = ctx->t[pt + 1] - ctx->t[pt];
            r[0] = dt;
            r[1] = std::sqrt(dt);
            int o = 2;
            if (d > 3) {  // large-D record: full H, F0, Psi transposed; matrices in mma fragment order (dmcmc_device.cuh)
                for (int k = 0; k < d; ++k)
                    for (int i = 0; i < d; ++i) r[o + big_frag(k, i)] = H[pt * d * d + (size_t)k * d + i];
                o += d * d;
                for (int i = 0; i < d; ++i) r[o++] = F0[pt * d + i];
                if (lay.has_psi)
                    for (int k = 0; k < d; ++k)
                        for (int i = 0; i < d; ++i) r[o + big_frag(k, i)] = Psi[pt * d * d + (size_t)i * d + k];
            } else {
                for (int i = 0; i < d; ++i)
                    for (int q = i; q < d; ++q) r[o++] = H[pt * d * d + i * d + q];
                for (int i = 0; i < d; ++i) r[o++] = F0[pt * d + i];
                if (lay.has_psi)
                    for (int i = 0; i < d * d; ++i) r[o++] = Psi[pt * d * d + i];
            }
        }
    std::vector<double> c0b(lay.nblk), g((size_t)lay.nblk * d, 0.0), Qc((size_t)lay.nblk * d * d, 0.0);
    for (int b = 0; b < lay.nblk; ++b) c0b[b] = c0[ctx->pt_off[lay.i1[b]]];
    if (blk_g) g.assign(blk_g, blk_g + (size_t)lay.nblk * d);
    if (blk_Qc) Qc.assign(blk_Qc, blk_Qc + (size_t)lay.nblk * d * d);
    CK(cudaMemcpy(tb.table.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(tb.blk_c0.p, c0b.data(), c0b.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(tb.blk_g.p, g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(tb.blk_Qc.p, Qc.data(), Qc.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (tb.ptH.p) CK(cudaMemcpy(tb.ptH.p, H, (size_t)ctx->P * d * d * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(tb.ptF0.p, F0, (size_t)ctx->P * d * sizeof(double), cudaMemcpyHostToDevice));
    if (tb.ptPsi.p) {
        if (Psi) CK(cudaMemcpy(tb.ptPsi.p, Psi, (size_t)ctx->P * d * d * sizeof(double), cudaMemcpyHostToDevice));
        else CK(cudaMemset(tb.ptPsi.p, 0, (size_t)ctx->P * d * d * sizeof(double)));
    }
    CK(cudaMemcpy(tb.ptc0.p, c0, (size_t)ctx->P * sizeof(double), cudaMemcpyHostToDevice));
    tb.valid = true;
    tb.fvalid = false;
    return 0;
}

int dmcmc_download_tables(dmcmc_ctx *ctx, int32_t which, double *H, double *F0, double *Psi,
                          double *c0, double *blk_g, double *blk_Qc) {
    REQUIRE(ctx && (which == 0 || which == 1), "dmcmc_download_tables: bad argument");
    REQUIRE(ctx->lay_cur >= 0, "dmcmc_download_tables: no layout set");
    CK(cudaSetDevice(ctx->cfg.device));
    Layout &lay = ctx->layout[ctx->lay_cur];
    Tables &tb = lay.tab[ctx->law_cur ^ which];
    REQUIRE(tb.valid, "dmcmc_download_tables: tables not computed");
    if (sync(ctx)) return -1;
    const size_t d = ctx->d, P = ctx->P;
    REQUIRE(tb.ptH.p || (!H && !Psi), "dmcmc_download_tables: per-point H/Psi are not kept for this problem size");
    if (H) CK(cudaMemcpy(H, tb.ptH.p, P * d * d * sizeof(double), cudaMemcpyDeviceToHost));
    if (F0) CK(cudaMemcpy(F0, tb.ptF0.p, P * d * sizeof(double), cudaMemcpyDeviceToHost));
    if (Psi) CK(cudaMemcpy(Psi, tb.ptPsi.p, P * d * d * sizeof(double), cudaMemcpyDeviceToHost));
    if (c0) CK(cudaMemcpy(c0, tb.ptc0.p, P * sizeof(double), cudaMemcpyDeviceToHost));
    if (blk_g) CK(cudaMemcpy(blk_g, tb.blk_g.p, lay.nblk * d * sizeof(double), cudaMemcpyDeviceToHost));
    if (blk_Qc) CK(cudaMemcpy(blk_Qc, tb.blk_Qc.p, lay.nblk * d * d * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

// one PathImputation update; internal form shared by impute / init_paths
static int impute_once(dmcmc_ctx *ctx, uint32_t iter, const double *Z, const double *E,
                       bool force_accept, const uint8_t *mask_host, bool zero_rho,
                       double *out_ll_prop, double *out_ll, uint8_t *out_acc) {
    SolveParams sp;
    if (fill_params(ctx, sp, ctx->law_cur)) return -1;
    if (ensure_unit_buffers(ctx)) return -1;
    const size_t n = (size_t)ctx->C * sp.nblk;
    REQUIRE((Z == nullptr) == (E == nullptr) || force_accept, "dmcmc_impute: pass both Z and E, or neither");
    if (Z) {
        if (stage_noise(ctx, Z)) return -1;
        sp.Z = ctx->d_Z.p;
        if (E) {
            CK(cudaMemcpyAsync(ctx->d_E.p, E, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        } else {
            CK(cudaMemsetAsync(ctx->d_E.p, 0, n * sizeof(double), ctx->stream));
        }
        sp.E = ctx->d_E.p;
    }
    if (mask_host) {
        CK(cudaMemcpyAsync(ctx->d_mask.p, mask_host, n, cudaMemcpyHostToDevice, ctx->stream));
        sp.unit_mask = ctx->d_mask.p;
    }
    DevBuf<double> zero_rho_buf;
    if (zero_rho) {
        CK(zero_rho_buf.ensure(ctx->J));
        CK(cudaMemsetAsync(zero_rho_buf.p, 0, ctx->J * sizeof(double), ctx->stream));
        sp.rho = zero_rho_buf.p;
    }
    sp.iter = iter;
    sp.force_accept = force_accept;
    // without stored noise every update recovers it from the accepted X (for init_paths, rho = 0
    // makes the recovered value irrelevant)
    sp.reinvert = ((!force_accept && !ctx->w_valid) || !ctx->cfg.store_noise) ? 1 : 0;
    sp.store_w = !sp.reinvert;
    sp.out_llprop = ctx->d_out_llprop.p;
    sp.out_ll = ctx->d_out_ll.p;
    sp.out_acc = ctx->d_out_acc.p;
    if (!force_accept) {
        sp.acc_count = ctx->d_counts.p;
        sp.prop_count = ctx->d_counts.p + ctx->J;
    }
    if (mask_host) {  // masked-out units keep their previous outputs: clear first
        CK(cudaMemsetAsync(ctx->d_out_acc.p, 0, n, ctx->stream));
    }
    if (timed_launch(ctx, MODE_IMPUTE, sp)) return -1;
    ctx->sel_cur ^= 1;
    if (force_accept) ctx->w_valid = ctx->cfg.store_noise != 0;
    if (out_ll_prop) CK(cudaMemcpyAsync(out_ll_prop, ctx->d_out_llprop.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_ll) CK(cudaMemcpyAsync(out_ll, ctx->d_out_ll.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_acc) CK(cudaMemcpyAsync(out_acc, ctx->d_out_acc.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (sync(ctx)) return -1;
    zero_rho_buf.release();
    return 0;
}

int dmcmc_init_ll(dmcmc_ctx *ctx) {
    REQUIRE(ctx, "dmcmc_init_ll: null context");
    CK(cudaSetDevice(ctx->cfg.device));
    SolveParams sp;
    if (fill_params(ctx, sp, ctx->law_cur)) return -1;
    if (timed_launch(ctx, MODE_LOGLIK, sp)) return -1;
    return sync(ctx);
}

int dmcmc_init_paths(dmcmc_ctx *ctx, const double *Z) {
    REQUIRE(ctx, "dmcmc_init_paths: null context");
    CK(cudaSetDevice(ctx->cfg.device));
    REQUIRE(ctx->lay_cur >= 0, "dmcmc_init_paths: no layout set");
    const size_t n = (size_t)ctx->C * ctx->layout[ctx->lay_cur].nblk;
    std::vector<uint8_t> done(n, 0), mask(n, 1), acc(n, 0);
    // init_paths!: un-preconditioned forward_guide! until success (src/workspaces.jl:86-90)
    for (int tr = 0; tr < 100; ++tr) {
        for (size_t i = 0; i < n; ++i) mask[i] = !done[i];
        if (impute_once(ctx, 0xFFFF0000u + (uint32_t)tr, Z, nullptr, true, mask.data(), true, nullptr, nullptr, acc.data()))
            return -1;
        bool all = true;
        for (size_t i = 0; i < n; ++i) {
            done[i] = done[i] || (mask[i] && acc[i]);
            all = all && done[i];
        }
        if (all) return dmcmc_init_ll(ctx);
        REQUIRE(Z == nullptr, "dmcmc_init_paths: the supplied noise leaves the model's domain");
    }
    return fail(ctx, "dmcmc_init_paths: no admissible path after 100 tries");
}

int dmcmc_impute(dmcmc_ctx *ctx, uint32_t iter, const double *Z, const double *E,
                 double *out_ll_prop, double *out_ll, uint8_t *out_accepted) {
    REQUIRE(ctx, "dmcmc_impute: null context");
    CK(cudaSetDevice(ctx->cfg.device));
    REQUIRE((Z == nullptr) == (E == nullptr), "dmcmc_impute: pass both Z and E, or neither");
    return impute_once(ctx, iter, Z, E, false, nullptr, false, out_ll_prop, out_ll, out_accepted);
}

// select a layout on the stream (no host sync): tables are cached per (layout, law)
static int switch_layout_async(dmcmc_ctx *ctx, int32_t nblk, const int32_t *i1, const int32_t *i2,
                               const int32_t *pinned) {
    const int before = ctx->lay_cur;
    if (select_layout(ctx, nblk, i1, i2, pinned)) return -1;
    if (resize_ll(ctx)) return -1;
    Layout &lay = ctx->layout[ctx->lay_cur];
    if (!lay.tab[ctx->law_cur].valid && run_backward_filter(ctx, ctx->law_cur)) return -1;
    if (before != ctx->lay_cur) ctx->w_valid = false;  // W and ll belong to the previous layout's law
    return 0;
}

// materialise W (re-inversion) and ll under the current law/layout
static int materialize_async(dmcmc_ctx *ctx) {
    REQUIRE(ctx->cfg.store_noise, "this context was created with store_noise = 0: the accepted Wiener increments are not kept (parameter steps and dmcmc_relayout need them)");
    SolveParams sp;
    if (fill_params(ctx, sp, ctx->law_cur)) return -1;
    if (timed_launch(ctx, MODE_RELAYOUT, sp)) return -1;
    ctx->w_valid = true;
    return 0;
}

static int relayout_async(dmcmc_ctx *ctx, int32_t nblk, const int32_t *i1, const int32_t *i2,
                          const int32_t *pinned) {
    if (switch_layout_async(ctx, nblk, i1, i2, pinned)) return -1;
    return materialize_async(ctx);
}

// is [p, p + bytes) page-locked host memory (cudaHostAlloc / cudaHostRegister)?  Then DMA can land in it directly.
static bool is_pinned(const void *p) {
    if (!p) return true;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static void drainer_main(dmcmc_ctx *ctx, int t) {
    SweepIO &io = ctx->io;
    cudaSetDevice(ctx->cfg.device);
    prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0);  // 1 us instead of the default 50 us on this thread's sleeps
    uint64_t seen = 0;
    for (;;) {
        DrainJob j;
        {
            std::unique_lock<std::mutex> lk(io.m);
            io.cv_job.wait(lk, [&] { return io.stop || io.gen != seen; });
            if (io.stop) return;
            seen = io.gen;
            j = io.job;
        }
        for (int it = t; it < j.n_sweeps; it += NDRAIN) {
            bool bail = false;
            while (io.recorded.load(std::memory_order_acquire) <= it) {  // sleep, do not spin: one process per GPU
                if (io.failed.load(std::memory_order_relaxed)) { bail = true; break; }  // shares the host cores with 7 others
                std::this_thread::sleep_for(std::chrono::microseconds(20));
            }
            if (bail) break;
            const int slot = it % SWEEP_RING;
            if (cudaEventSynchronize(io.copied[slot]) != cudaSuccess) { io.failed = true; break; }
            const uint8_t *src = io.h_hist + j.slot_bytes * slot;
            if (j.llp) std::memcpy(j.llp + j.row * it, src, j.row * sizeof(double));
            if (j.ll) std::memcpy(j.ll + j.row * it, src + j.row * sizeof(double), j.row * sizeof(double));
            if (j.acc) std::memcpy(j.acc + j.row * it, src + 2 * j.row * sizeof(double), j.row);
            io.drained[t].fetch_add(1, std::memory_order_release);
        }
        {
            std::lock_guard<std::mutex> lk(io.m);
            ++io.done;
        }
        io.cv_done.notify_one();
    }
}

// Host side of the sweep loop (what `run!` drives for a smoothing chain):
//   compute stream : [rho schedule of the chunk, one H2D]  kernel 0, kernel 1, ...   (kernel k waits for D2H k-R)
//   copy-out stream: history rows of sweep k: device ring -> host                   (waits for kernel k)
// PINNED caller buffers (dmcmc_host_alloc / dmcmc_host_register): the rows are DMA'd straight into the caller's
// arrays -- every byte crosses host memory once and no host thread touches it; the call only enqueues and then waits.
// PAGEABLE caller buffers: the rows land in the library's small pinned ring (which stays hot in the IOMMU/TLB) and
// a pool of drainer threads, kept parked between calls, moves them into the caller's arrays.
int dmcmc_run_sweeps(dmcmc_ctx *ctx, uint32_t iter0, int32_t n_sweeps, int32_t half_len,
                     int32_t first_pattern, const double *rho_sched, int32_t nblk_stride,
                     double *hist_ll_prop, double *hist_ll, uint8_t *hist_accepted) {
    REQUIRE(ctx && n_sweeps >= 0, "dmcmc_run_sweeps: bad argument");
    REQUIRE(ctx->lay_cur >= 0, "dmcmc_run_sweeps: no layout set");
    REQUIRE(!(ctx->adapt_on && rho_sched), "dmcmc_run_sweeps: with dmcmc_set_adaptation the pCN memory lives on the device; pass rho_sched = NULL");
    CK(cudaSetDevice(ctx->cfg.device));
    const bool want_hist = hist_ll_prop || hist_ll || hist_accepted;
    // the two chequerboard layouts (integer ranges; src/blocking.jl:7-9)
    std::vector<int32_t> li1[2], li2[2], lpin[2];
    int nb[2] = {0, 0};
    if (half_len > 0) {
        for (int pat = 0; pat < 2; ++pat) {
            li1[pat].resize(ctx->J); li2[pat].resize(ctx->J); lpin[pat].resize(ctx->J);
            nb[pat] = dmcmc_chequerboard_layout(ctx->R, ctx->rec_nobs.data(), half_len, pat,
                                                li1[pat].data(), li2[pat].data(), lpin[pat].data());
        }
    } else {
        nb[0] = nb[1] = ctx->layout[ctx->lay_cur].nblk;
    }
    const int nbmax = std::max(nb[0], nb[1]);
    REQUIRE(!want_hist || nblk_stride >= nbmax, "dmcmc_run_sweeps: nblk_stride smaller than the number of blocks");
    const size_t row = (size_t)ctx->C * (want_hist ? nblk_stride : nbmax);  // units per history row
    constexpr int R = SWEEP_RING;
    const size_t slot_bytes = 2 * row * sizeof(double) + ((row + 15) & ~(size_t)15);  // [ll deg | ll | accepted], 16-byte aligned
    const size_t J = (size_t)ctx->J;
    SweepIO &io = ctx->io;
    const bool direct = want_hist && is_pinned(hist_ll_prop) && is_pinned(hist_ll) && is_pinned(hist_accepted);
    const bool drain = want_hist && !direct && n_sweeps > 0;
    if (want_hist) CK(io.d_hist.ensure(slot_bytes * R));
    if (drain && io.h_hist_bytes < slot_bytes * R) {
        if (io.h_hist) cudaFreeHost(io.h_hist);
        io.h_hist = nullptr; io.h_hist_bytes = 0;
        CK(cudaHostAlloc((void **)&io.h_hist, slot_bytes * R, cudaHostAllocDefault));
        io.h_hist_bytes = slot_bytes * R;
    }
    if (!io.s_in) {
        CK(cudaStreamCreateWithFlags(&io.s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&io.s_out, cudaStreamNonBlocking));
        for (int i = 0; i < R; ++i) {
            CK(cudaEventCreateWithFlags(&io.kernel_done[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&io.copied[i], cudaEventDisableTiming | cudaEventBlockingSync));
        }
    }
    if (drain) {
        // Drainers: pinned ring -> caller's arrays.  One core copies ~10 GB/s, which is just the rate the C3 histories
        // arrive at (879 KB per 80 us sweep), and every hand-over costs two wake-ups: NDRAIN threads take every
        // NDRAIN-th sweep; they sleep, they do not spin.  The pool outlives the call (run! with adaptation every k
        // steps issues many short calls).
        static_assert(SWEEP_RING % NDRAIN == 0, "a ring slot is always drained by the same thread");
        if (!io.pool_started) {
            for (auto &d : io.drained) d.store(0);
            for (int t = 0; t < NDRAIN; ++t) io.drainer[t] = std::thread(drainer_main, ctx, t);
            io.pool_started = true;
        }
        {
            std::lock_guard<std::mutex> lk(io.m);
            io.job.n_sweeps = n_sweeps; io.job.row = row; io.job.slot_bytes = slot_bytes;
            io.job.llp = hist_ll_prop; io.job.ll = hist_ll; io.job.acc = hist_accepted;
            io.recorded.store(0);
            io.failed.store(false);
            for (auto &d : io.drained) d.store(0);
            io.done = 0;
            ++io.gen;
        }
        io.cv_job.notify_all();
    }
    // rho schedule: one H2D per chunk of sweeps, on the compute stream (ordered before the chunk's first kernel and
    // after the previous chunk's last one)
    const int CHUNK = (int)std::max<size_t>(1, std::min<size_t>((size_t)1 << 14, ((size_t)32 << 20) / (J * sizeof(double))));
    if (rho_sched) CK(io.d_rho.ensure(J * (size_t)std::min(n_sweeps, CHUNK)));
    int rc = 0;
    auto body = [&]() -> int {
        for (int it = 0; it < n_sweeps; ++it) {
            const int pat = (first_pattern + it) & 1;
            const int slot = it % R;
            if (drain) {  // the slot's previous rows have left the pinned ring (and hence the device ring)
                const int prev = it - R;  // the sweep that used this slot before
                while (prev >= 0 && io.drained[prev % NDRAIN].load(std::memory_order_acquire) < prev / NDRAIN + 1) {
                    if (io.failed.load(std::memory_order_relaxed)) return fail(ctx, "dmcmc_run_sweeps: history copy failed");
                    std::this_thread::sleep_for(std::chrono::microseconds(20));
                }
            }
            if (rho_sched && it % CHUNK == 0) {
                const size_t n = (size_t)std::min(CHUNK, n_sweeps - it);
                CK(cudaMemcpyAsync(io.d_rho.p, rho_sched + J * it, n * J * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            }
            if (half_len > 0 && switch_layout_async(ctx, nb[pat], li1[pat].data(), li2[pat].data(), lpin[pat].data()))
                return -1;
            SolveParams sp;
            if (fill_params(ctx, sp, ctx->law_cur)) return -1;
            if (rho_sched) sp.rho = io.d_rho.p + J * (size_t)(it % CHUNK);
            const int aslot = half_len > 0 ? pat : 0;
            if (ctx->adapt_on) sp.rho = ctx->d_rho_pat[aslot].p;
            sp.iter = iter0 + (uint32_t)it;
            sp.reinvert = !ctx->w_valid;  // fused layout switch: noise and ll recovered from the accepted X
            sp.store_w = ctx->w_valid;
            sp.out_stride = want_hist ? nblk_stride : sp.nblk;
            sp.acc_count = ctx->adapt_on ? ctx->d_cnt_pat[aslot].p : ctx->d_counts.p;
            sp.prop_count = sp.acc_count + ctx->J;
            if (want_hist) {
                uint8_t *d = io.d_hist.p + slot_bytes * slot;
                sp.out_llprop = hist_ll_prop ? (double *)d : nullptr;
                sp.out_ll = hist_ll ? (double *)(d + row * sizeof(double)) : nullptr;
                sp.out_acc = hist_accepted ? d + 2 * row * sizeof(double) : nullptr;
                if (it >= R) CK(cudaStreamWaitEvent(ctx->stream, io.copied[slot], 0));  // device ring slot is free
            }
            if (timed_launch(ctx, MODE_IMPUTE, sp)) return -1;
            ctx->sel_cur ^= 1;
            if (ctx->adapt_on) {  // register! + time_to_update + readjust!, stream-ordered behind the sweep
                const Layout &lay = ctx->layout[ctx->lay_cur];
                const double mcmciter = (double)((iter0 + (uint32_t)it) / (uint32_t)ctx->adapt_spi + 1u);
                adapt_rho_kernel<<<(unsigned)((lay.nblk + 127) / 128), 128, 0, ctx->stream>>>(
                    lay.nblk, lay.d_i1.p, lay.d_i2.p, ctx->d_cnt_pat[aslot].p, ctx->d_cnt_pat[aslot].p + ctx->J,
                    ctx->d_rho_pat[aslot].p, ctx->adapt, mcmciter);
                CK(cudaGetLastError());
            }
            if (want_hist) {
                CK(cudaEventRecord(io.kernel_done[slot], ctx->stream));
                CK(cudaStreamWaitEvent(io.s_out, io.kernel_done[slot], 0));
                const uint8_t *d = io.d_hist.p + slot_bytes * slot;
                if (direct) {  // straight into the caller's pinned arrays
                    if (hist_ll_prop) CK(cudaMemcpyAsync(hist_ll_prop + row * it, d, row * sizeof(double), cudaMemcpyDeviceToHost, io.s_out));
                    if (hist_ll) CK(cudaMemcpyAsync(hist_ll + row * it, d + row * sizeof(double), row * sizeof(double), cudaMemcpyDeviceToHost, io.s_out));
                    if (hist_accepted) CK(cudaMemcpyAsync(hist_accepted + row * it, d + 2 * row * sizeof(double), row, cudaMemcpyDeviceToHost, io.s_out));
                } else {       // the kernel wrote its rows with the caller's stride: one contiguous copy per sweep
                    CK(cudaMemcpyAsync(io.h_hist + slot_bytes * slot, d, slot_bytes, cudaMemcpyDeviceToHost, io.s_out));
                }
                CK(cudaEventRecord(io.copied[slot], io.s_out));
                if (drain) io.recorded.store(it + 1, std::memory_order_release);
            }
        }
        return 0;
    };
    const int old_slack = prctl(PR_GET_TIMERSLACK, 0, 0, 0, 0);  // this thread's waits for a drained ring slot
    if (drain) prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0);
    rc = body();
    if (drain && old_slack > 0) prctl(PR_SET_TIMERSLACK, (unsigned long)old_slack, 0, 0, 0);
    if (drain) {
        if (rc) io.failed = true;
        std::unique_lock<std::mutex> lk(io.m);
        io.cv_done.wait(lk, [&] { return io.done == NDRAIN; });
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc) rc = fail(ctx, "dmcmc_run_sweeps: stream synchronisation failed");
    if (io.s_out && cudaStreamSynchronize(io.s_out) != cudaSuccess && !rc) rc = fail(ctx, "dmcmc_run_sweeps: history copy failed");
    if (!rc && drain && io.failed) rc = fail(ctx, "dmcmc_run_sweeps: history copy failed");
    return rc;
}

void *dmcmc_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void dmcmc_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int dmcmc_host_register(void *p, size_t bytes) {
    if (!p || !bytes) return -1;
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return -1; }
    return 0;
}

int dmcmc_host_unregister(void *p) {
    if (!p) return -1;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return -1; }
    return 0;
}

int dmcmc_relayout(dmcmc_ctx *ctx, int32_t nblk, const int32_t *i1, const int32_t *i2, const int32_t *pinned) {
    REQUIRE(ctx && i1 && i2 && pinned, "dmcmc_relayout: bad argument");
    CK(cudaSetDevice(ctx->cfg.device));
    if (relayout_async(ctx, nblk, i1, i2, pinned)) return -1;
    return sync(ctx);
}

// set_proposal! up to and including the solve (src/run!_alterations.jl:177-208): theta deg into the proposal law, aux law
// and backward filter when critical, fixed-noise solve of every (chain, block) -- enqueued; per-unit ll deg and success
// are left on the device (d_param_llprop, d_out_acc)
static int param_solve_async(dmcmc_ctx *ctx, const double *theta_prop, int32_t ntheta, int32_t critical, int *nblk_out) {
    REQUIRE(ctx && theta_prop, "dmcmc_param_loglik: bad argument");
    REQUIRE(ntheta == ctx->nth, "dmcmc_param_loglik: wrong number of parameters for this model");
    REQUIRE(ctx->lay_cur >= 0, "dmcmc_param_loglik: no layout set");
    CK(cudaSetDevice(ctx->cfg.device));
    const int cur = ctx->law_cur, prop = cur ^ 1;
    Law &lp = ctx->law[prop];
    Layout &lay = ctx->layout[ctx->lay_cur];
    std::copy(theta_prop, theta_prop + ntheta, lp.theta);
    lp.have_theta = true;
    if (critical) {
        // critical_change && backward_filter!(ws deg.P[i])  (src/run!_alterations.jl:199-204)
        lp.custom_aux = false;
        if (build_aux(ctx, prop)) return -1;
        if (run_backward_filter(ctx, prop)) return -1;
    } else {
        // non-critical: aux law and tables are those of the current law
        Law &lc = ctx->law[cur];
        lp.custom_aux = lc.custom_aux;
        if (lc.custom_aux) { lp.auxB = lc.auxB; lp.auxbeta = lc.auxbeta; lp.auxsig = lc.auxsig; }
        if (build_aux(ctx, prop)) return -1;
        Tables &tc = lay.tab[cur], &tp = lay.tab[prop];
        REQUIRE(tc.valid, "dmcmc_param_loglik: current tables missing");
        if (ensure_tables_alloc(ctx, lay, prop)) return -1;
        auto cp = [&](DevBuf<double> &dst, DevBuf<double> &src) {
            return cudaMemcpyAsync(dst.p, src.p, src.bytes(), cudaMemcpyDeviceToDevice, ctx->stream);
        };
        CK(cp(tp.table, tc.table)); CK(cp(tp.blk_c0, tc.blk_c0)); CK(cp(tp.blk_g, tc.blk_g));
        CK(cp(tp.blk_Qc, tc.blk_Qc)); CK(cp(tp.ptH, tc.ptH)); CK(cp(tp.ptF0, tc.ptF0));
        CK(cp(tp.ptPsi, tc.ptPsi)); CK(cp(tp.ptc0, tc.ptc0));
        tp.valid = true;
        tp.fvalid = false;
    }
    // the solve needs the accepted W under the CURRENT law: recover it first when it is stale
    if (!ctx->w_valid && materialize_async(ctx)) return -1;
    SolveParams sp;
    if (fill_params(ctx, sp, prop)) return -1;
    if (ensure_unit_buffers(ctx)) return -1;
    const size_t n = (size_t)ctx->C * sp.nblk;
    CK(ctx->d_param_llprop.ensure(n));
    sp.out_llprop = ctx->d_param_llprop.p;
    sp.out_acc = ctx->d_out_acc.p;
    if (timed_launch(ctx, MODE_PARAM, sp)) return -1;
    *nblk_out = sp.nblk;
    return 0;
}

int dmcmc_param_loglik(dmcmc_ctx *ctx, const double *theta_prop, int32_t ntheta, int32_t critical,
                       double *out_ll_prop, uint8_t *out_success) {
    int nblk = 0;
    if (param_solve_async(ctx, theta_prop, ntheta, critical, &nblk)) return -1;
    const size_t n = (size_t)ctx->C * nblk;
    ctx->h_param_llprop.resize(n);
    std::vector<uint8_t> okb(n);
    CK(cudaMemcpyAsync(ctx->h_param_llprop.data(), ctx->d_param_llprop.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(okb.data(), ctx->d_out_acc.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (sync(ctx)) return -1;
    if (out_ll_prop) std::copy(ctx->h_param_llprop.begin(), ctx->h_param_llprop.end(), out_ll_prop);
    if (out_success)
        for (int c = 0; c < ctx->C; ++c) {
            uint8_t ok = 1;
            for (int b = 0; b < nblk; ++b) ok &= okb[(size_t)c * nblk + b];
            out_success[c] = ok;
        }
    ctx->param_pending = true;
    return 0;
}

__global__ void flip_sel_kernel(uint8_t *sel, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sel[i] ^= 1;
}

int dmcmc_param_commit(dmcmc_ctx *ctx, int32_t accepted) {
    REQUIRE(ctx, "dmcmc_param_commit: null context");
    REQUIRE(ctx->param_pending, "dmcmc_param_commit: no pending dmcmc_param_loglik");
    CK(cudaSetDevice(ctx->cfg.device));
    ctx->param_pending = false;
    if (!accepted) return 0;
    // swap_laws_and_paths!: P deg <-> P, XX deg <-> XX, W kept  (src/run!_alterations.jl:328-334)
    const size_t n = (size_t)ctx->J * ctx->C;
    flip_sel_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_selX[ctx->sel_cur].p, n);
    CK(cudaGetLastError());
    ctx->law_cur ^= 1;
    const size_t nu = (size_t)ctx->C * ctx->layout[ctx->lay_cur].nblk;
    CK(cudaMemcpyAsync(ctx->d_ll.p, ctx->d_param_llprop.p, nu * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    // tables of the other layout slot belong to the old law now
    for (int s = 0; s < 2; ++s)
        if (s != ctx->lay_cur) ctx->layout[s].tab[ctx->law_cur].valid = false;
    return sync(ctx);
}

int dmcmc_download_state(dmcmc_ctx *ctx, int32_t which, double *X, double *W) {
    REQUIRE(ctx && X && W && (which == 0 || which == 1), "dmcmc_download_state: bad argument");
    REQUIRE(ctx->lay_cur >= 0 && ctx->have_start, "dmcmc_download_state: context not initialised");
    CK(cudaSetDevice(ctx->cfg.device));
    Layout &lay = ctx->layout[ctx->lay_cur];
    if (ctx->cfg.store_noise && !ctx->w_valid && which == 0 && lay.tab[ctx->law_cur].valid && materialize_async(ctx)) return -1;
    REQUIRE(ctx->cfg.store_noise, "dmcmc_download_state: store_noise = 0 keeps no W; use dmcmc_download_path / dmcmc_download_segment");
    DevBuf<double> oX, oW;
    const size_t nX = (size_t)ctx->C * ctx->P * ctx->d, nW = (size_t)ctx->C * ctx->P * ctx->m;
    CK(oX.ensure(nX));
    CK(oW.ensure(nW));
    GatherParams g{};
    g.C = ctx->C; g.J = ctx->J; g.D = ctx->d; g.M = ctx->m; g.TT = ctx->TT; g.P = ctx->P;
    g.iv_N = ctx->d_N.p; g.iv_tile_off = ctx->d_tile_off.p; g.iv_first = ctx->d_first.p;
    g.iv_rec = ctx->d_rec.p; g.iv_pt_off = ctx->d_pt_off.p; g.iv_blk = lay.d_iv_blk.p;
    g.blk_i1 = lay.d_i1.p;
    g.iv_xoff = ctx->d_xoff.p; g.iv_xs = ctx->d_xs.p;
    for (int b = 0; b < 2; ++b) { g.X[b] = ctx->d_X[b].p; g.W[b] = ctx->d_W[b].p; }
    g.selX = ctx->d_selX[ctx->sel_cur].p; g.selW = ctx->d_selW[ctx->sel_cur].p;
    g.x0 = ctx->d_x0.p; g.which = which; g.outX = oX.p; g.outW = oW.p;
    if (flush_pending_pull(ctx)) return -1;  // a fused exchange's message may still sit in the mailbox
    CK(launch_gather(g, ctx->stream));
    CK(cudaMemcpyAsync(X, oX.p, nX * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(W, oW.p, nW * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (sync(ctx)) return -1;
    oX.release();
    oW.release();
    return 0;
}

int dmcmc_download_path(dmcmc_ctx *ctx, int32_t chain, int32_t rec, double *x_out) {
    REQUIRE(ctx && x_out, "dmcmc_download_path: bad argument");
    REQUIRE(chain >= 0 && chain < ctx->C && rec >= 0 && rec < ctx->R, "dmcmc_download_path: index out of range");
    REQUIRE(ctx->have_start, "dmcmc_download_path: context not initialised");
    // glue: the starting point, then every interval's points 1..N -- exactly what _copy_path!
    // builds from the per-interval containers (src/path_saving_buffer.jl:54-64)
    int j0 = 0;
    for (int r = 0; r < rec; ++r) j0 += ctx->rec_nobs[r];
    const int j1 = j0 + ctx->rec_nobs[rec], d = ctx->d;
    const int steps = ctx->st_off[j1] - ctx->st_off[j0];
    // only this chain's records leave the device (a chain's interval record is contiguous in HBM)
    if (segment_copy(ctx, j0, j1, x_out + d, false, chain, 1)) return -1;
    // starting point from the device copy (dmcmc_set_start_device may have replaced it): x0[(rec * d + i) * C + chain]
    CK(cudaMemcpy2D(x_out, sizeof(double), ctx->d_x0.p + (size_t)rec * d * ctx->C + chain, (size_t)ctx->C * sizeof(double),
                    sizeof(double), d, cudaMemcpyDeviceToHost));
    (void)steps;
    return 0;
}

// a12 without stalling the sampler (save!, src/path_saving_buffer.jl:52; trigger src/run!_alterations.jl:51-54):
// the glued accepted paths of the chosen chains are snapshotted into a device buffer on the compute stream (so the
// sweeps enqueued afterwards cannot touch what is saved) and copied to a pinned buffer on the copy-out stream while
// those sweeps run; dmcmc_save_path_wait collects them.  Up to four saves may be in flight.
int dmcmc_save_path_async(dmcmc_ctx *ctx, int32_t n_sel, const int32_t *chains, int32_t rec, int32_t *ticket) {
    REQUIRE(ctx && chains && ticket && n_sel > 0, "dmcmc_save_path_async: bad argument");
    REQUIRE(rec >= 0 && rec < ctx->R && ctx->have_start, "dmcmc_save_path_async: bad recording / context not initialised");
    for (int i = 0; i < n_sel; ++i) REQUIRE(chains[i] >= 0 && chains[i] < ctx->C, "dmcmc_save_path_async: chain out of range");
    CK(cudaSetDevice(ctx->cfg.device));
    int tk = -1;
    for (int i = 0; i < 4; ++i)
        if (!ctx->saves[i].busy) { tk = i; break; }
    REQUIRE(tk >= 0, "dmcmc_save_path_async: four saves already in flight (call dmcmc_save_path_wait)");
    PathSave &sv = ctx->saves[tk];
    int j0 = 0;
    for (int r = 0; r < rec; ++r) j0 += ctx->rec_nobs[r];
    const int j1 = j0 + ctx->rec_nobs[rec], d = ctx->d;
    const int steps = ctx->st_off[j1] - ctx->st_off[j0];
    const size_t per = (size_t)(steps + 1) * d;   // one glued path: start point + every step
    sv.n = per * n_sel;
    CK(sv.d_buf.ensure(sv.n));
    if (sv.h_n < sv.n) {
        if (sv.h_buf) cudaFreeHost(sv.h_buf);
        sv.h_buf = nullptr; sv.h_n = 0;
        CK(cudaHostAlloc((void **)&sv.h_buf, sv.n * sizeof(double), cudaHostAllocDefault));
        sv.h_n = sv.n;
    }
    if (!sv.snap) {
        CK(cudaEventCreateWithFlags(&sv.snap, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&sv.landed, cudaEventDisableTiming | cudaEventBlockingSync));
    }
    if (!ctx->save_stream) CK(cudaStreamCreateWithFlags(&ctx->save_stream, cudaStreamNonBlocking));
    for (int i = 0; i < n_sel; ++i) {
        SegmentParams g{};
        g.C = ctx->C; g.D = d; g.j0 = j0; g.j1 = j1; g.steps = steps; g.TT = ctx->TT;
        g.iv_N = ctx->d_N.p; g.iv_tile_off = ctx->d_tile_off.p; g.iv_st_off = ctx->d_st_off.p;
        g.iv_xoff = ctx->d_xoff.p; g.iv_xs = ctx->d_xs.p;
        g.X[0] = ctx->d_X[0].p; g.X[1] = ctx->d_X[1].p;
        g.selX = ctx->d_selX[ctx->sel_cur].p;
        g.packed = sv.d_buf.p + per * i + d; g.upload = 0; g.pk_stride = (int64_t)per; g.c0 = chains[i]; g.nc = 1;
        if (flush_pending_pull(ctx)) return -1;  // a fused exchange's message may still sit in the mailbox
        CK(launch_segment(g, ctx->stream));
        // starting point x0[(rec * d + c) * C + chain] -> first d doubles of the glued path
        CK(cudaMemcpy2DAsync(sv.d_buf.p + per * i, sizeof(double), ctx->d_x0.p + (size_t)rec * d * ctx->C + chains[i],
                             (size_t)ctx->C * sizeof(double), sizeof(double), d, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CK(cudaEventRecord(sv.snap, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->save_stream, sv.snap, 0));
    CK(cudaMemcpyAsync(sv.h_buf, sv.d_buf.p, sv.n * sizeof(double), cudaMemcpyDeviceToHost, ctx->save_stream));
    CK(cudaEventRecord(sv.landed, ctx->save_stream));
    sv.busy = true;
    *ticket = tk;
    return 0;
}

int dmcmc_save_path_wait(dmcmc_ctx *ctx, int32_t ticket, double *x_out) {
    REQUIRE(ctx && x_out && ticket >= 0 && ticket < 4 && ctx->saves[ticket].busy, "dmcmc_save_path_wait: no such save in flight");
    CK(cudaSetDevice(ctx->cfg.device));
    PathSave &sv = ctx->saves[ticket];
    CK(cudaEventSynchronize(sv.landed));
    std::memcpy(x_out, sv.h_buf, sv.n * sizeof(double));
    sv.busy = false;
    return 0;
}

int dmcmc_get_ll(dmcmc_ctx *ctx, double *ll) {
    REQUIRE(ctx && ll && ctx->lay_cur >= 0, "dmcmc_get_ll: bad argument");
    CK(cudaSetDevice(ctx->cfg.device));
    if (sync(ctx)) return -1;
    CK(cudaMemcpy(ll, ctx->d_ll.p, (size_t)ctx->C * ctx->layout[ctx->lay_cur].nblk * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int dmcmc_set_ll(dmcmc_ctx *ctx, const double *ll) {
    REQUIRE(ctx && ll && ctx->lay_cur >= 0, "dmcmc_set_ll: bad argument");
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaMemcpy(ctx->d_ll.p, ll, (size_t)ctx->C * ctx->layout[ctx->lay_cur].nblk * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

int dmcmc_accept_counts(dmcmc_ctx *ctx, int64_t *accepted, int64_t *proposed) {
    REQUIRE(ctx && ctx->lay_cur >= 0, "dmcmc_accept_counts: bad argument");
    CK(cudaSetDevice(ctx->cfg.device));
    if (sync(ctx)) return -1;
    std::vector<unsigned long long> h((size_t)2 * ctx->J);
    CK(cudaMemcpy(h.data(), ctx->d_counts.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    const int nblk = ctx->layout[ctx->lay_cur].nblk;
    for (int b = 0; b < nblk; ++b) {
        if (accepted) accepted[b] = (int64_t)h[b];
        if (proposed) proposed[b] = (int64_t)h[ctx->J + b];
    }
    CK(cudaMemset(ctx->d_counts.p, 0, ctx->d_counts.bytes()));
    return 0;
}

int32_t dmcmc_n_blocks(const dmcmc_ctx *ctx) {
    return (ctx && ctx->lay_cur >= 0) ? ctx->layout[ctx->lay_cur].nblk : -1;
}

int dmcmc_kernel_times(dmcmc_ctx *ctx, double *ms, int64_t *launches) {
    REQUIRE(ctx, "dmcmc_kernel_times: null context");
    CK(cudaSetDevice(ctx->cfg.device));
    if (sync(ctx)) return -1;
    double tot[4] = {0, 0, 0, 0};
    int64_t cnt[4] = {0, 0, 0, 0};
    for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
        float f = 0.f;
        CK(cudaEventElapsedTime(&f, ctx->ev[i], ctx->ev[i + 1]));
        const int m = ctx->ev_mode[i / 2] & 3;
        tot[m] += f;
        cnt[m] += 1;
    }
    for (int m = 0; m < 4; ++m) {
        if (ms) ms[m] = tot[m];
        if (launches) launches[m] = cnt[m];
    }
    ctx->ev_used = 0;
    return 0;
}

int64_t dmcmc_device_bytes(const dmcmc_ctx *ctx) {
    if (!ctx) return 0;
    int64_t b = 0;
    for (int i = 0; i < 2; ++i)
        b += ctx->d_X[i].bytes() + ctx->d_W[i].bytes() + ctx->d_selX[i].bytes() + ctx->d_selW[i].bytes();
    for (auto &l : ctx->layout)
        for (auto &t : l.tab) b += t.table.bytes() + t.ptH.bytes() + t.ptF0.bytes() + t.ptPsi.bytes() + t.ptc0.bytes();
    b += ctx->d_Z.bytes() + ctx->d_ll.bytes();
    return b;
}

}  // extern "C"

extern "C" int dmcmc_set_fast_step(dmcmc_ctx *ctx, int32_t on) {
    REQUIRE(ctx, "dmcmc_set_fast_step: null context");
    ctx->fast_step = on != 0;
    return 0;
}

extern "C" int dmcmc_set_cooperative(dmcmc_ctx *ctx, int32_t mode) {
    REQUIRE(ctx && mode >= -1 && mode <= 1, "dmcmc_set_cooperative: -1 (auto), 0 (never) or 1 (always)");
    ctx->coop = mode;
    return 0;
}

extern "C" int32_t dmcmc_conjugate_dim(const dmcmc_ctx *ctx, int32_t *theta_idx) {
    return ctx ? conjugate_dim(ctx->cfg.model, theta_idx) : -1;
}

// mu [C][nc] and W [C][nc][nc] of the conjugate Gaussian update over the accepted path under the
// current law (compute_mu_and_W, src/updates.jl:241-262); nc = dmcmc_conjugate_dim
extern "C" int dmcmc_conjugate_sums(dmcmc_ctx *ctx, double *mu, double *W) {
    REQUIRE(ctx && mu && W, "dmcmc_conjugate_sums: bad argument");
    REQUIRE(ctx->have_grid && ctx->have_start, "dmcmc_conjugate_sums: context not initialised");
    const int nc = conjugate_dim(ctx->cfg.model, nullptr);
    REQUIRE(nc > 0, "dmcmc_conjugate_sums: this law has no conjugate parameters implemented");
    CK(cudaSetDevice(ctx->cfg.device));
    const int nout = nc + nc * nc;
    DevBuf<double> partial, out;
    CK(partial.ensure((size_t)ctx->J * ctx->C * nout));
    CK(out.ensure((size_t)ctx->C * nout));
    ConjParams c{};
    c.C = ctx->C; c.J = ctx->J;
    c.iv_N = ctx->d_N.p; c.iv_first = ctx->d_first.p; c.iv_rec = ctx->d_rec.p; c.iv_xs = ctx->d_xs.p;
    c.iv_pt_off = ctx->d_pt_off.p; c.iv_xoff = ctx->d_xoff.p; c.t = ctx->d_t.p;
    c.X[0] = ctx->d_X[0].p; c.X[1] = ctx->d_X[1].p;
    c.selX = ctx->d_selX[ctx->sel_cur].p;
    c.x0 = ctx->d_x0.p; c.partial = partial.p;
    for (int i = 0; i < DMCMC_MAX_THETA; ++i) c.th.v[i] = ctx->law[ctx->law_cur].theta[i];
    if (flush_pending_pull(ctx)) return -1;  // a fused exchange's message may still sit in the mailbox
    CK(launch_conjugate(ctx->cfg.model, c, out.p, ctx->stream));
    std::vector<double> h((size_t)ctx->C * nout);
    CK(cudaMemcpyAsync(h.data(), out.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (sync(ctx)) return -1;
    for (int ch = 0; ch < ctx->C; ++ch) {
        std::copy(h.begin() + (size_t)ch * nout, h.begin() + (size_t)ch * nout + nc, mu + (size_t)ch * nc);
        std::copy(h.begin() + (size_t)ch * nout + nc, h.begin() + (size_t)(ch + 1) * nout, W + (size_t)ch * nc * nc);
    }
    partial.release();
    out.release();
    return 0;
}

extern "C" int dmcmc_set_timing(dmcmc_ctx *ctx, int32_t every) {
    REQUIRE(ctx && every >= 0, "dmcmc_set_timing: bad argument");
    ctx->time_every = every;
    return 0;
}

extern "C" int64_t dmcmc_launch_count(dmcmc_ctx *ctx) {
    if (!ctx) return -1;
    const int64_t n = (int64_t)(ctx->launches - ctx->launches_reported);
    ctx->launches_reported = ctx->launches;
    return n;
}

extern "C" int dmcmc_timer_start(dmcmc_ctx *ctx) {
    REQUIRE(ctx, "dmcmc_timer_start: null context");
    CK(cudaSetDevice(ctx->cfg.device));
    if (!ctx->sw0) { CK(cudaEventCreate(&ctx->sw0)); CK(cudaEventCreate(&ctx->sw1)); }
    CK(cudaEventRecord(ctx->sw0, ctx->stream));
    return 0;
}

extern "C" int dmcmc_timer_stop(dmcmc_ctx *ctx, double *ms) {
    REQUIRE(ctx && ms && ctx->sw0, "dmcmc_timer_stop: no running timer");
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaEventRecord(ctx->sw1, ctx->stream));
    CK(cudaEventSynchronize(ctx->sw1));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, ctx->sw0, ctx->sw1));
    *ms = (double)f;
    return 0;
}

extern "C" int dmcmc_set_threads(dmcmc_ctx *ctx, int32_t threads) {
    REQUIRE(ctx && threads >= 32 && threads <= 128 && threads % 32 == 0, "dmcmc_set_threads: 32, 64, 96 or 128");
    ctx->threads = threads;
    return 0;
}

// ---- measurement helper: FP64 FMA throughput of this device (the denominator of the Lorenz-96 kernel's FP64 view) ----
// Every thread runs eight independent DFMA chains (the dependent-issue latency of one chain is ~8 cycles, measured in
// tools/micro/fp64lat.cu), 32 warps per SM resident: nothing but the FP64 pipe limits it.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double b, double c) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-9 * (double)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" double dmcmc_fp64_peak(int32_t device) {
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *out = nullptr;
    if (cudaMalloc((void **)&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {  // first pass warms the clocks
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}

// FP64 tensor-core peak (mma.sync m8n8k4.f64 = SASS DMMA.8x8x4, the instruction the Lorenz-96 solve kernel's mat-vecs run
// on): eight independent accumulator pairs per warp, 16 warps per SM.  On B200 DMMA and DFMA share the FP64 datapath
// (36.8 against 34.2 TFLOP/s; a DFMA stream next to a DMMA stream is throttled), so this is the ceiling of that kernel.
__global__ void __launch_bounds__(256) fp64_dmma_peak_kernel(double *out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k][0] = k; c[k][1] = 1e-9 * threadIdx.x; }
    const double av = a + 1e-9 * threadIdx.x, bv = b + 1e-12 * threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(av), "d"(bv));
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" double dmcmc_fp64_tensor_peak(int32_t device) {
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    const int blocks = prop.multiProcessorCount * 2, threads = 256, iters = 4096;
    double *out = nullptr;
    if (cudaMalloc((void **)&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fp64_dmma_peak_kernel<<<blocks, threads>>>(out, iters, 1e-3, 1e-3);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 256 * 4 * 8 * (double)iters * blocks * (threads / 32);  // 8 x 8 x 4 FMAs per warp instruction
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}

// =================================================================================================
// Block-sharded recording over the GPUs of one box: the collectives live in the library, so a host without
// torch.distributed (the Julia plugin) gets the end-point exchange and the parameter-step all-reduce through ccall.
// =================================================================================================
extern "C" int dmcmc_comm_unique_id(void *id128) {
    dmcmc_ctx *ctx = nullptr;
    std::string err;
    NcclApi *nc = nccl_api(err);
    if (!nc) return fail(nullptr, err);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    NK(nc->GetUniqueId(reinterpret_cast<ncclUniqueId *>(id128)));
    return 0;
}

extern "C" int dmcmc_comm_init(dmcmc_ctx *ctx, const void *id128, int32_t rank, int32_t world) {
    REQUIRE(ctx && id128 && world >= 1 && rank >= 0 && rank < world, "dmcmc_comm_init: bad argument");
    CK(cudaSetDevice(ctx->cfg.device));
    std::string err;
    NcclApi *nc = nccl_api(err);
    REQUIRE(nc, "dmcmc_comm_init: " + err);
    if (ctx->comm) { nc->CommDestroy(ctx->comm); ctx->comm = nullptr; }
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    NK(nc->CommInitRank(&ctx->comm, world, id, rank));
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    return 0;
}

extern "C" int dmcmc_comm_destroy(dmcmc_ctx *ctx) {
    REQUIRE(ctx, "dmcmc_comm_destroy: null context");
    CK(cudaSetDevice(ctx->cfg.device));
    if (flush_pending_pull(ctx)) return -1;
    if (sync(ctx)) return -1;
    comm_release(ctx);
    ctx->comm_rank = 0;
    ctx->comm_world = 1;
    return 0;
}

// Collective over the communicator: every rank allocates its mailbox region, the CUDA IPC handles travel by one
// ncclAllGather, each rank maps its two neighbours' regions.  Peer memory is used only if EVERY rank got that far (the
// ranks agree through an all-reduce); otherwise dmcmc_halo_exchange stays on ncclSend / ncclRecv.  mailbox_bytes: the
// largest halo message (n_chains x Euler steps of half_len + 1 intervals x d doubles); larger messages go through NCCL.
extern "C" int dmcmc_comm_enable_p2p(dmcmc_ctx *ctx, int64_t mailbox_bytes) {
    REQUIRE(ctx && ctx->comm, "dmcmc_comm_enable_p2p: call dmcmc_comm_init first");
    REQUIRE(mailbox_bytes > 0, "dmcmc_comm_enable_p2p: bad mailbox size");
    CK(cudaSetDevice(ctx->cfg.device));
    if (sync(ctx)) return -1;
    std::string err;
    NcclApi *nc = nccl_api(err);
    REQUIRE(nc, "dmcmc_comm_enable_p2p: " + err);
    p2p_release(ctx);
    const int r = ctx->comm_rank, w = ctx->comm_world;
    if (w == 1) return 0;
    const size_t cap = ((size_t)mailbox_bytes + 255) & ~(size_t)255;
    struct Rec { cudaIpcMemHandle_t h; int32_t ok; int32_t pad[15]; };
    static_assert(sizeof(Rec) == 128, "IPC record");
    Rec mine{};
    int ok = 1;
    if (cudaMalloc((void **)&ctx->p2p_region, 256 + 2 * cap) != cudaSuccess) ok = 0;
    if (ok && cudaMemset(ctx->p2p_region, 0, 256 + 2 * cap) != cudaSuccess) ok = 0;
    if (ok && cudaMalloc((void **)&ctx->p2p_counter, 4 * sizeof(unsigned int)) != cudaSuccess) ok = 0;
    if (ok && cudaMemset(ctx->p2p_counter, 0, 4 * sizeof(unsigned int)) != cudaSuccess) ok = 0;
    if (ok && cudaHostAlloc((void **)&ctx->p2p_err, sizeof(int), cudaHostAllocMapped) != cudaSuccess) ok = 0;
    if (ok) { *ctx->p2p_err = 0; if (cudaHostGetDevicePointer((void **)&ctx->p2p_err_dev, ctx->p2p_err, 0) != cudaSuccess) ok = 0; }
    if (ok && cudaIpcGetMemHandle(&mine.h, ctx->p2p_region) != cudaSuccess) ok = 0;
    cudaGetLastError();
    mine.ok = ok;
    // all-gather the records (device buffers: NCCL)
    Rec *d_all = nullptr, *d_mine = nullptr;
    CK(cudaMalloc((void **)&d_all, sizeof(Rec) * w));
    CK(cudaMalloc((void **)&d_mine, sizeof(Rec)));
    CK(cudaMemcpyAsync(d_mine, &mine, sizeof(Rec), cudaMemcpyHostToDevice, ctx->stream));
    NK(nc->AllGather(d_mine, d_all, sizeof(Rec), ncclChar, ctx->comm, ctx->stream));
    std::vector<Rec> all(w);
    CK(cudaMemcpyAsync(all.data(), d_all, sizeof(Rec) * w, cudaMemcpyDeviceToHost, ctx->stream));
    if (sync(ctx)) return -1;
    for (int i = 0; i < w; ++i) ok = ok && all[i].ok;
    if (ok) {
        for (int side = 0; side < 2 && ok; ++side) {
            const int peer = side == 0 ? r - 1 : r + 1;
            if (peer < 0 || peer >= w) continue;
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[peer].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
            ctx->p2p_peer[side] = (char *)ptr;
        }
    }
    // every rank must have mapped its neighbours
    int32_t *d_ok = reinterpret_cast<int32_t *>(d_mine);
    const int32_t okv = ok;
    CK(cudaMemcpyAsync(d_ok, &okv, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    NK(nc->AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream));
    int32_t all_ok = 0;
    CK(cudaMemcpyAsync(&all_ok, d_ok, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (sync(ctx)) return -1;
    cudaFree(d_all);
    cudaFree(d_mine);
    if (!all_ok) { p2p_release(ctx); return 0; }
    ctx->p2p_on = true;
    ctx->p2p_cap = cap;
    return 0;
}

extern "C" int dmcmc_comm_p2p_enabled(const dmcmc_ctx *ctx) { return ctx && ctx->p2p_on ? 1 : 0; }

static SegmentParams seg_params(dmcmc_ctx *ctx, int j0, int j1, double *dev, int64_t stride, int up) {
    SegmentParams g{};
    g.C = ctx->C; g.D = ctx->d; g.j0 = j0; g.j1 = j1; g.steps = ctx->st_off[j1] - ctx->st_off[j0]; g.TT = ctx->TT;
    g.iv_N = ctx->d_N.p; g.iv_tile_off = ctx->d_tile_off.p; g.iv_st_off = ctx->d_st_off.p;
    g.iv_xoff = ctx->d_xoff.p; g.iv_xs = ctx->d_xs.p;
    g.X[0] = ctx->d_X[0].p; g.X[1] = ctx->d_X[1].p;
    g.selX = ctx->d_selX[ctx->sel_cur].p;
    g.packed = dev; g.upload = up; g.pk_stride = stride;
    return g;
}

// one direction of the exchange over peer memory.  dir 0: messages go to rank - 1, dir 1: to rank + 1.
//   send: intervals [sj0, sj1) -> the neighbour's mailbox[dir];   receive: own mailbox[dir] -> intervals [rj0, rj1),
//   preceded (dir 1) by `lead` doubles per chain whose last d are the new starting point
static int halo_p2p(dmcmc_ctx *ctx, int dir, bool send, int sj0, int sj1, bool recv, int rj0, int rj1, size_t lead) {
    if (!send && !recv) return 0;
    REQUIRE(*ctx->p2p_err == 0, "dmcmc_halo_exchange: a wait on the neighbouring GPU timed out (peer gone?)");
    char *flags = ctx->p2p_region;
    auto data_seq = [](char *region, int d) { return reinterpret_cast<unsigned long long *>(region) + d; };
    auto ack_seq = [](char *region, int d) { return reinterpret_cast<unsigned long long *>(region) + 2 + d; };
    auto mailbox = [&](char *region, int d) { return reinterpret_cast<double *>(region + 256 + (size_t)d * ctx->p2p_cap); };
    SegmentParams gs{}, gr{};
    P2PSync ys{}, yr{};
    const double *start_src = nullptr;
    int64_t stride = 0;
    if (send) {
        char *peer = ctx->p2p_peer[dir == 0 ? 0 : 1];
        gs = seg_params(ctx, sj0, sj1, mailbox(peer, dir), (int64_t)(ctx->st_off[sj1] - ctx->st_off[sj0]) * ctx->d, 0);
        ys = P2PSync{ack_seq(flags, dir), ctx->p2p_sent[dir], data_seq(peer, dir), ctx->p2p_sent[dir] + 1,
                     ctx->p2p_counter + dir, ctx->p2p_err_dev};
        ++ctx->p2p_sent[dir];
    }
    if (recv) {
        char *peer = ctx->p2p_peer[dir == 0 ? 1 : 0];   // the sender: right neighbour for leftward messages
        stride = (int64_t)lead + (int64_t)(ctx->st_off[rj1] - ctx->st_off[rj0]) * ctx->d;
        double *mb = mailbox(flags, dir);
        gr = seg_params(ctx, rj0, rj1, mb + lead, stride, 1);
        yr = P2PSync{data_seq(flags, dir), ctx->p2p_recvd[dir] + 1, ack_seq(peer, dir), ctx->p2p_recvd[dir] + 1,
                     ctx->p2p_counter + 2 + dir, ctx->p2p_err_dev};
        start_src = lead ? mb + lead - ctx->d : nullptr;
        ++ctx->p2p_recvd[dir];
        ctx->w_valid = false;
    }
    CK(launch_halo_exchange(send ? &gs : nullptr, ys, recv ? &gr : nullptr, ctx->d_x0.p, start_src, stride, ctx->R, yr, ctx->stream));
    return 0;
}

namespace {
struct P2PAddr {
    dmcmc_ctx *c;
    unsigned long long *data_seq(char *region, int d) const { return reinterpret_cast<unsigned long long *>(region) + d; }
    unsigned long long *ack_seq(char *region, int d) const { return reinterpret_cast<unsigned long long *>(region) + 2 + d; }
    double *mailbox(char *region, int d) const { return reinterpret_cast<double *>(region + 256 + (size_t)d * c->p2p_cap); }
};
}  // namespace

static int flush_pending_pull(dmcmc_ctx *ctx) {
    if (!ctx->pull_pending) return 0;
    const int dir = ctx->pull_dir;
    ctx->pull_pending = false;
    return halo_p2p(ctx, dir, false, 0, 0, true, ctx->pull_j0, ctx->pull_j1, ctx->pull_lead);
}

// One imputation sweep (Philox noise, current layout and block mask, enqueued like dmcmc_impute_async) WITH the end-point
// exchange that follows it under `pattern` (the arguments of dmcmc_halo_exchange) fused into the sweep kernels: the edge
// block's units push their accepted intervals into the neighbour's mailbox right after their accept decision, and the
// neighbour's next fused sweep (or the first other call that touches the paths) takes them out.  Needs
// dmcmc_comm_enable_p2p and a launch small enough for the warp-per-unit kernel; otherwise it is dmcmc_impute_async
// followed by dmcmc_halo_exchange.  Results are identical either way.
extern "C" int dmcmc_impute_exchange_async(dmcmc_ctx *ctx, uint32_t iter, int32_t pattern, int32_t n_first, int32_t n_halo,
                                           int32_t left_edge_steps) {
    REQUIRE(ctx && ctx->comm, "dmcmc_impute_exchange_async: call dmcmc_comm_init first");
    REQUIRE(pattern == 0 || pattern == 1, "dmcmc_impute_exchange_async: pattern is 0 or 1");
    REQUIRE(ctx->R == 1 && ctx->lay_cur >= 0, "dmcmc_impute_exchange_async: one recording, layout set");
    const int r = ctx->comm_rank, w = ctx->comm_world, J = ctx->J, d = ctx->d;
    REQUIRE(n_first > 0 && n_first <= J - n_halo && n_halo >= 0 && n_halo < J, "dmcmc_impute_exchange_async: bad interval counts");
    REQUIRE((r < w - 1) == (n_halo > 0) || w == 1, "dmcmc_impute_exchange_async: every rank but the last holds a halo");
    const int n_own = J - n_halo;
    const size_t C = (size_t)ctx->C;
    // the messages of this pattern's exchange (as in dmcmc_halo_exchange)
    int sj0, sj1, rj0, rj1;
    size_t lead = 0;
    bool send, recv;
    if (pattern == 0) { send = r > 0; sj0 = 0; sj1 = n_first; recv = r < w - 1; rj0 = n_own; rj1 = J; }
    else { send = r < w - 1; sj0 = n_own - 1; sj1 = J; recv = r > 0; rj0 = 0; rj1 = n_first; lead = (size_t)left_edge_steps * d; }
    const size_t send_b = send ? C * (size_t)(ctx->st_off[sj1] - ctx->st_off[sj0]) * d * sizeof(double) : 0;
    const size_t recv_b = recv ? C * ((size_t)(ctx->st_off[rj1] - ctx->st_off[rj0]) * d + lead) * sizeof(double) : 0;
    Layout &lay = ctx->layout[ctx->lay_cur];
    SolveParams probe{};
    probe.C = ctx->C; probe.nblk = lay.nblk; probe.coop = ctx->coop;
    const bool fuse = ctx->p2p_on && send_b <= ctx->p2p_cap && recv_b <= ctx->p2p_cap &&
                      solve_uses_warp_per_unit(ctx->cfg.model, probe) && (pattern == 0 || left_edge_steps > 0 || r == 0) &&
                      sj1 - sj0 <= 32 && rj1 - rj0 <= 32;  // a warp maps one interval of the message per lane
    if (!fuse) {
        if (dmcmc_impute_async(ctx, iter)) return -1;
        return dmcmc_halo_exchange(ctx, pattern, n_first, n_halo, left_edge_steps);
    }
    REQUIRE(*ctx->p2p_err == 0, "dmcmc_impute_exchange_async: a wait on the neighbouring GPU timed out (peer gone?)");
    // the block both messages belong to: the one that holds interval 0 (pattern 0) or the last interval (pattern 1)
    const int target = pattern == 0 ? 0 : J - 1;
    int eb = -1;
    for (int b = 0; b < lay.nblk; ++b)
        if (lay.i1[b] <= target && target <= lay.i2[b]) eb = b;
    REQUIRE(eb >= 0, "dmcmc_impute_exchange_async: no block holds the edge interval");
    // a pending message is this sweep's to take out only if it was sent under the OTHER pattern (it feeds the same block)
    if (ctx->pull_pending && ctx->pull_dir == pattern) { if (flush_pending_pull(ctx)) return -1; }
    const P2PAddr A{ctx};
    char *mine = ctx->p2p_region;
    HaloFuse h{};
    h.blk = eb;
    h.iv_st_off = ctx->d_st_off.p;
    h.err = ctx->p2p_err_dev;
    if (ctx->pull_pending) {
        const int pd = ctx->pull_dir;                       // == 1 - pattern
        char *sender = ctx->p2p_peer[pd == 0 ? 1 : 0];
        REQUIRE(lay.i1[eb] <= ctx->pull_j0 && ctx->pull_j1 - 1 <= lay.i2[eb], "dmcmc_impute_exchange_async: the pending message is not the edge block's");
        h.pull_box = A.mailbox(mine, pd);
        h.pull_j0 = ctx->pull_j0; h.pull_j1 = ctx->pull_j1; h.pull_lead = (int32_t)ctx->pull_lead;
        h.pull_stride = (int64_t)ctx->pull_lead + (int64_t)(ctx->st_off[ctx->pull_j1] - ctx->st_off[ctx->pull_j0]) * d;
        h.pull_flag = A.data_seq(mine, pd); h.pull_val = ctx->p2p_recvd[pd] + 1;
        h.pull_ack = A.ack_seq(sender, pd); h.pull_counter = ctx->p2p_counter + 2 + pd;
        ++ctx->p2p_recvd[pd];
        ctx->pull_pending = false;
        ctx->w_valid = false;
    }
    if (send) {
        REQUIRE(lay.i1[eb] <= sj0 && sj1 - 1 <= lay.i2[eb], "dmcmc_impute_exchange_async: the message is not the edge block's");
        char *peer = ctx->p2p_peer[pattern == 0 ? 0 : 1];
        h.push_box = A.mailbox(peer, pattern);
        h.push_j0 = sj0; h.push_j1 = sj1;
        h.push_stride = (int64_t)(ctx->st_off[sj1] - ctx->st_off[sj0]) * d;
        h.push_wait = A.ack_seq(mine, pattern); h.push_wait_val = ctx->p2p_sent[pattern];
        h.push_flag = A.data_seq(peer, pattern); h.push_val = ctx->p2p_sent[pattern] + 1;
        h.push_counter = ctx->p2p_counter + pattern;
        ++ctx->p2p_sent[pattern];
    }
    ctx->hf_next = h;
    ctx->hf_armed = true;
    if (dmcmc_impute_async(ctx, iter)) { ctx->hf_armed = false; return -1; }
    if (recv) {  // my neighbour pushes at the end of its sweep: mine to take out next
        ctx->pull_pending = true;
        ctx->pull_dir = pattern; ctx->pull_j0 = rj0; ctx->pull_j1 = rj1; ctx->pull_lead = lead;
    }
    return 0;
}

static int seg_on_stream(dmcmc_ctx *ctx, int j0, int j1, double *dev, int64_t stride, int up) {
    const int steps = ctx->st_off[j1] - ctx->st_off[j0];
    SegmentParams g{};
    g.C = ctx->C; g.D = ctx->d; g.j0 = j0; g.j1 = j1; g.steps = steps; g.TT = ctx->TT;
    g.iv_N = ctx->d_N.p; g.iv_tile_off = ctx->d_tile_off.p; g.iv_st_off = ctx->d_st_off.p;
    g.iv_xoff = ctx->d_xoff.p; g.iv_xs = ctx->d_xs.p;
    g.X[0] = ctx->d_X[0].p; g.X[1] = ctx->d_X[1].p;
    g.selX = ctx->d_selX[ctx->sel_cur].p;
    g.packed = dev; g.upload = up; g.pk_stride = stride;
    CK(launch_segment(g, ctx->stream));
    return 0;
}

// The exchange that follows a sweep of a block-sharded chequerboard sampler.  This context holds the contiguous
// interval range its rank owns followed by a halo of n_halo intervals owned by the right neighbour (0 on the last rank).
//   after_pattern 0 (blocks end on rank edges):  rank r -> r-1 : accepted path of my first n_first intervals
//                                                (the left neighbour's halo: n_first there == its n_halo)
//   after_pattern 1 (blocks straddle the edges): rank r -> r+1 : my last owned interval (its end point is the
//                                                neighbour's new starting point) followed by the halo I just updated
// left_edge_steps = Euler steps of the left neighbour's last owned interval (0 on rank 0).  Everything is enqueued on the
// context's stream (pack kernel, ncclSend/ncclRecv, unpack kernel): no host synchronisation.
extern "C" int dmcmc_halo_exchange(dmcmc_ctx *ctx, int32_t after_pattern, int32_t n_first, int32_t n_halo,
                                   int32_t left_edge_steps) {
    REQUIRE(ctx && ctx->comm, "dmcmc_halo_exchange: call dmcmc_comm_init first");
    REQUIRE(ctx->R == 1, "dmcmc_halo_exchange: block sharding is defined for a single recording");
    const int r = ctx->comm_rank, w = ctx->comm_world, J = ctx->J, d = ctx->d;
    REQUIRE(n_first > 0 && n_first <= J - n_halo && n_halo >= 0 && n_halo < J, "dmcmc_halo_exchange: bad interval counts");
    REQUIRE((r < w - 1) == (n_halo > 0) || w == 1, "dmcmc_halo_exchange: every rank but the last holds a halo");
    CK(cudaSetDevice(ctx->cfg.device));
    if (flush_pending_pull(ctx)) return -1;
    std::string err;
    NcclApi *nc = nccl_api(err);
    REQUIRE(nc, "dmcmc_halo_exchange: " + err);
    const int n_own = J - n_halo;
    const size_t C = (size_t)ctx->C;
    if (after_pattern == 0) {
        size_t send_n = r > 0 ? C * (size_t)(ctx->st_off[n_first] - ctx->st_off[0]) * d : 0;
        size_t recv_n = r < w - 1 ? C * (size_t)(ctx->st_off[J] - ctx->st_off[n_own]) * d : 0;
        // per link: a message that fits the mailbox goes over peer memory (both ends see the same size), else NCCL
        const bool ps = ctx->p2p_on && send_n * sizeof(double) <= ctx->p2p_cap, pr = ctx->p2p_on && recv_n * sizeof(double) <= ctx->p2p_cap;
        if (halo_p2p(ctx, 0, ps && send_n > 0, 0, n_first, pr && recv_n > 0, n_own, J, 0)) return -1;
        if (ps) send_n = 0;
        if (pr) recv_n = 0;
        if (send_n == 0 && recv_n == 0) return 0;
        CK(ctx->cb_send.ensure(std::max<size_t>(send_n, 1)));
        CK(ctx->cb_recv.ensure(std::max<size_t>(recv_n, 1)));
        if (send_n && seg_on_stream(ctx, 0, n_first, ctx->cb_send.p, (int64_t)(send_n / C), 0)) return -1;
        NK(nc->GroupStart());
        if (send_n) NK(nc->Send(ctx->cb_send.p, send_n, ncclDouble, r - 1, ctx->comm, ctx->stream));
        if (recv_n) NK(nc->Recv(ctx->cb_recv.p, recv_n, ncclDouble, r + 1, ctx->comm, ctx->stream));
        NK(nc->GroupEnd());
        if (recv_n && seg_on_stream(ctx, n_own, J, ctx->cb_recv.p, (int64_t)(recv_n / C), 1)) return -1;
        if (recv_n) ctx->w_valid = false;
    } else {
        size_t send_n = r < w - 1 ? C * (size_t)(ctx->st_off[J] - ctx->st_off[n_own - 1]) * d : 0;
        const size_t first_steps = (size_t)(ctx->st_off[n_first] - ctx->st_off[0]);
        size_t recv_n = r > 0 ? C * ((size_t)left_edge_steps + first_steps) * d : 0;
        REQUIRE(r == 0 || left_edge_steps > 0, "dmcmc_halo_exchange: left_edge_steps missing");
        const bool ps = ctx->p2p_on && send_n * sizeof(double) <= ctx->p2p_cap, pr = ctx->p2p_on && recv_n * sizeof(double) <= ctx->p2p_cap;
        if (halo_p2p(ctx, 1, ps && send_n > 0, n_own - 1, J, pr && recv_n > 0, 0, n_first, (size_t)left_edge_steps * d)) return -1;
        if (ps) send_n = 0;
        if (pr) recv_n = 0;
        if (send_n == 0 && recv_n == 0) return 0;
        CK(ctx->cb_send.ensure(std::max<size_t>(send_n, 1)));
        CK(ctx->cb_recv.ensure(std::max<size_t>(recv_n, 1)));
        if (send_n && seg_on_stream(ctx, n_own - 1, J, ctx->cb_send.p, (int64_t)(send_n / C), 0)) return -1;
        NK(nc->GroupStart());
        if (send_n) NK(nc->Send(ctx->cb_send.p, send_n, ncclDouble, r + 1, ctx->comm, ctx->stream));
        if (recv_n) NK(nc->Recv(ctx->cb_recv.p, recv_n, ncclDouble, r - 1, ctx->comm, ctx->stream));
        NK(nc->GroupEnd());
        if (recv_n) {
            const int64_t stride = (int64_t)(recv_n / C);
            // the edge interval's last point is my new starting point; the rest are my first n_first intervals
            CK(launch_set_start(ctx->d_x0.p, ctx->cb_recv.p + (size_t)(left_edge_steps - 1) * d, stride, ctx->C, ctx->R, d, ctx->stream));
            if (seg_on_stream(ctx, 0, n_first, ctx->cb_recv.p + (size_t)left_edge_steps * d, stride, 1)) return -1;
            ctx->w_valid = false;
        }
    }
    return 0;
}

// a9 for a block-sharded recording: set_proposal! on every rank's blocks, then ONE decision for all blocks
// (src/run!_alterations.jl:177-211, :367) needs the ll deg of every block of every rank.  blk_global[b] = global index of
// local block b under the current layout, or -1 when another rank owns it (halo).  Out (identical on every rank, and
// bit for bit what the unsharded context gives): ll deg and current ll of every global block [n_chains][nblk_global],
// success AND-reduced over all blocks [n_chains].  Follow with dmcmc_param_commit(accepted) on every rank.
extern "C" int dmcmc_param_loglik_sharded(dmcmc_ctx *ctx, const double *theta_prop, int32_t ntheta, int32_t critical,
                                          int32_t nblk_global, const int32_t *blk_global, double *out_ll_prop_global,
                                          double *out_ll_global, uint8_t *out_success) {
    REQUIRE(ctx && ctx->comm, "dmcmc_param_loglik_sharded: call dmcmc_comm_init first");
    REQUIRE(blk_global && nblk_global > 0, "dmcmc_param_loglik_sharded: bad argument");
    std::string err;
    NcclApi *nc = nccl_api(err);
    REQUIRE(nc, "dmcmc_param_loglik_sharded: " + err);
    int nblk = 0;
    if (param_solve_async(ctx, theta_prop, ntheta, critical, &nblk)) return -1;
    for (int b = 0; b < nblk; ++b) REQUIRE(blk_global[b] < nblk_global, "dmcmc_param_loglik_sharded: global block index out of range");
    const size_t ng = (size_t)ctx->C * nblk_global;
    CK(ctx->d_gl_llp.ensure(ng));
    CK(ctx->d_gl_ll.ensure(ng));
    CK(ctx->d_gl_ok.ensure(ctx->C));
    CK(ctx->d_gidx.ensure(nblk));
    CK(cudaMemcpyAsync(ctx->d_gidx.p, blk_global, nblk * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_gl_llp.p, 0, ng * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_gl_ll.p, 0, ng * sizeof(double), ctx->stream));
    fill_i32_kernel<<<(unsigned)((ctx->C + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_gl_ok.p, ctx->C, 1);
    const int nu = ctx->C * nblk;
    scatter_blocks_kernel<<<(unsigned)((nu + 127) / 128), 128, 0, ctx->stream>>>(
        ctx->d_param_llprop.p, ctx->d_ll.p, ctx->d_out_acc.p, ctx->d_gidx.p, ctx->C, nblk, nblk_global, ctx->d_gl_llp.p,
        ctx->d_gl_ll.p, ctx->d_gl_ok.p);
    CK(cudaGetLastError());
    NK(nc->GroupStart());
    NK(nc->AllReduce(ctx->d_gl_llp.p, ctx->d_gl_llp.p, ng, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    NK(nc->AllReduce(ctx->d_gl_ll.p, ctx->d_gl_ll.p, ng, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    NK(nc->AllReduce(ctx->d_gl_ok.p, ctx->d_gl_ok.p, (size_t)ctx->C, ncclInt32, ncclMin, ctx->comm, ctx->stream));
    NK(nc->GroupEnd());
    std::vector<int32_t> okh(ctx->C);
    if (out_ll_prop_global) CK(cudaMemcpyAsync(out_ll_prop_global, ctx->d_gl_llp.p, ng * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_ll_global) CK(cudaMemcpyAsync(out_ll_global, ctx->d_gl_ll.p, ng * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(okh.data(), ctx->d_gl_ok.p, ctx->C * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (sync(ctx)) return -1;
    if (out_success)
        for (int c = 0; c < ctx->C; ++c) out_success[c] = (uint8_t)(okh[c] != 0);
    ctx->param_pending = true;
    return 0;
}

// =================================================================================================
// Device-side adaptation of the pCN memory (src/adaptation.jl:82-149): with it dmcmc_run_sweeps runs
// straight through the adaptation points -- register!, time_to_update and readjust! happen in a one-thread-per-block
// kernel behind every sweep -- instead of returning to the host every adapt_every_k_steps iterations.
// =================================================================================================
extern "C" int dmcmc_set_adaptation(dmcmc_ctx *ctx, const dmcmc_adapt *cfg, int32_t sweeps_per_iteration,
                                    const double *rho_pattern0, const double *rho_pattern1) {
    REQUIRE(ctx && ctx->have_grid, "dmcmc_set_adaptation: call dmcmc_set_grid first");
    CK(cudaSetDevice(ctx->cfg.device));
    if (!cfg) {
        ctx->adapt_on = false;
        return 0;
    }
    REQUIRE(cfg->adapt_every_k_steps > 0 && cfg->min > 0.0 && cfg->max < 1.0 && cfg->min < cfg->max && sweeps_per_iteration > 0,
            "dmcmc_set_adaptation: bad configuration");
    const size_t J = ctx->J;
    for (int pat = 0; pat < 2; ++pat) {
        const double *src = pat == 0 ? rho_pattern0 : (rho_pattern1 ? rho_pattern1 : rho_pattern0);
        CK(ctx->d_rho_pat[pat].ensure(J));
        CK(ctx->d_cnt_pat[pat].ensure(2 * J));
        if (src) CK(cudaMemcpyAsync(ctx->d_rho_pat[pat].p, src, J * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        else CK(cudaMemcpyAsync(ctx->d_rho_pat[pat].p, ctx->d_rho.p, J * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_cnt_pat[pat].p, 0, 2 * J * sizeof(unsigned long long), ctx->stream));
    }
    ctx->adapt = *cfg;
    ctx->adapt_spi = sweeps_per_iteration;
    ctx->adapt_on = true;
    return sync(ctx);
}

extern "C" int dmcmc_get_rho(dmcmc_ctx *ctx, int32_t pattern, double *rho) {
    REQUIRE(ctx && rho && (pattern == 0 || pattern == 1), "dmcmc_get_rho: bad argument");
    CK(cudaSetDevice(ctx->cfg.device));
    if (sync(ctx)) return -1;
    const double *src = ctx->adapt_on ? ctx->d_rho_pat[pattern].p : ctx->d_rho.p;
    CK(cudaMemcpy(rho, src, (size_t)ctx->J * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
