// Access-pattern microbenchmark: what HBM throughput does the (chain, block)-per-thread tiled
// streaming pattern reach with no compute at all?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ void ld32(const double* p, double* v) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st32(double* p, const double* v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

// pattern kernel: thread = (chain, block); block owns TPB consecutive tiles; per tile read 64 B x 2 comps, write same
template <int DEPTH>
__global__ void pattern(const double* __restrict__ X0, const double* __restrict__ X1, double* __restrict__ Y0, double* __restrict__ Y1,
                        const unsigned char* __restrict__ sel, int C, int nblk, int TPB, long long TT, int do_write) {
    long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (long long)C * nblk) return;
    int blk = u / C, chain = u % C;
    long long cs = TT * C * 8;
    double acc = 0;
    double buf[DEPTH][2][8];
    int s = sel[u];
    const double* X = s ? X1 : X0;
    double* Y = s ? Y0 : Y1;
    long long t0 = (long long)blk * TPB;
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
        long long base = ((t0 + d) * C + chain) * 8;
        for (int c = 0; c < 2; ++c) { ld32(X + c * cs + base, buf[d][c]); ld32(X + c * cs + base + 4, buf[d][c] + 4); }
    }
    for (int t = 0; t < TPB; t += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            if (t + d >= TPB) break;
            long long base = ((t0 + t + d) * C + chain) * 8;
            double v[2][8];
            for (int c = 0; c < 2; ++c) for (int k = 0; k < 8; ++k) { v[c][k] = buf[d][c][k]; acc += v[c][k]; }
            if (t + d + DEPTH < TPB) {
                long long nb = ((t0 + t + d + DEPTH) * C + chain) * 8;
                for (int c = 0; c < 2; ++c) { ld32(X + c * cs + nb, buf[d][c]); ld32(X + c * cs + nb + 4, buf[d][c] + 4); }
            }
            // a little dependent math so the loop is not free
            for (int c = 0; c < 2; ++c) for (int k = 0; k < 8; ++k) v[c][k] = v[c][k] * 1.0000001 + acc * 1e-30;
            if (do_write) for (int c = 0; c < 2; ++c) { st32(Y + c * cs + base, v[c]); st32(Y + c * cs + base + 4, v[c] + 4); }
        }
    }
    if (acc == 123.456) Y[0] = acc;
}

// 128-byte ownership: one lane owns a whole line (16 steps) of one chain.  HALF=0: read/write the
// whole line at once; HALF=1: touch it as two 64-byte halves in consecutive iterations.
template <int HALF>
__global__ void pattern128(const double* __restrict__ X0, const double* __restrict__ X1, double* __restrict__ Y0, double* __restrict__ Y1,
                           const unsigned char* __restrict__ sel, int C, int nblk, int LPB, long long TL, int do_write) {
    long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (long long)C * nblk) return;
    int blk = u / C, chain = u % C;
    long long cs = TL * C * 16;
    double acc = 0;
    int s = sel[u];
    const double* X = s ? X1 : X0;
    double* Y = s ? Y0 : Y1;
    long long l0 = (long long)blk * LPB;
    constexpr int W = HALF ? 8 : 16;
    double cur[2][W], nxt[2][W];
    auto addr = [&](int it) { return HALF ? ((l0 + it / 2) * C + chain) * 16 + (it & 1) * 8 : ((l0 + it) * C + chain) * 16; };
    const int NIT = HALF ? 2 * LPB : LPB;
    for (int c = 0; c < 2; ++c) for (int k = 0; k < W; k += 4) ld32(X + c * cs + addr(0) + k, nxt[c] + k);
    for (int it = 0; it < NIT; ++it) {
        for (int c = 0; c < 2; ++c) for (int k = 0; k < W; ++k) { cur[c][k] = nxt[c][k]; acc += cur[c][k]; }
        if (it + 1 < NIT) for (int c = 0; c < 2; ++c) for (int k = 0; k < W; k += 4) ld32(X + c * cs + addr(it + 1) + k, nxt[c] + k);
        for (int c = 0; c < 2; ++c) for (int k = 0; k < W; ++k) cur[c][k] = cur[c][k] * 1.0000001 + acc * 1e-30;
        if (do_write) for (int c = 0; c < 2; ++c) for (int k = 0; k < W; k += 4) st32(Y + c * cs + addr(it) + k, cur[c] + k);
    }
    if (acc == 123.456) Y[0] = acc;
}

// pair-cooperative pattern: layout [tile][chain][comp(2)][8] = one 128-byte line per (chain, tile);
// lanes (2i, 2i+1) load chain 2i's line together (64 contiguous bytes per instruction), then chain
// 2i+1's, and swap halves with shuffles; same for the stores.
__device__ __forceinline__ double shx(double v) { return __shfl_xor_sync(0xffffffffu, v, 1); }
__global__ void pattern_pair(const double* X0, const double* X1, double* Y0, double* Y1,
                             const unsigned char* __restrict__ sel, int C, int nblk, int TPB, long long TT, int rs = 16) {
    long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (long long)C * nblk) return;
    int blk = u / C, chain = u % C, odd = chain & 1;
    int s_own = sel[u], s_oth = __shfl_xor_sync(0xffffffffu, s_own, 1);
    int sA = odd ? s_oth : s_own, sB = odd ? s_own : s_oth;     // A = even chain of the pair, B = odd chain
    const double* XA = sA ? X1 : X0; const double* XB = sB ? X1 : X0;
    double* YA = sA ? Y0 : Y1; double* YB = sB ? Y0 : Y1;
    long long cA = chain - odd, cB = cA + 1;
    double acc = 0;
    long long t0 = (long long)blk * TPB;
    double nA[2][4], nB[2][4];
    auto lineoff = [&](long long t, long long c) { return (t * C + c) * rs; };
    auto loadpair = [&](long long t) {
        for (int c = 0; c < 2; ++c) { ld32(XA + lineoff(t, cA) + c * 8 + odd * 4, nA[c]); }
        for (int c = 0; c < 2; ++c) { ld32(XB + lineoff(t, cB) + c * 8 + odd * 4, nB[c]); }
    };
    loadpair(t0);
    for (int t = 0; t < TPB; ++t) {
        double v[2][8];
        // own chain's data: even lane owns A (has A[.][0:4], needs A[.][4:8] from partner's nA); odd lane owns B
        for (int c = 0; c < 2; ++c) for (int k = 0; k < 4; ++k) {
            double mineA = nA[c][k], mineB = nB[c][k];
            double send = odd ? mineA : mineB;          // what the partner needs
            double got = shx(send);
            if (!odd) { v[c][k] = mineA; v[c][4 + k] = got; } else { v[c][k] = got; v[c][4 + k] = mineB; }
        }
        if (t + 1 < TPB) loadpair(t0 + t + 1);
        for (int c = 0; c < 2; ++c) for (int k = 0; k < 8; ++k) { acc += v[c][k]; v[c][k] = v[c][k] * 1.0000001 + acc * 1e-30; }
        // stores: swap halves back so that each instruction writes 64 contiguous bytes of one chain
        double oA[2][4], oB[2][4];
        for (int c = 0; c < 2; ++c) for (int k = 0; k < 4; ++k) {
            double send = odd ? v[c][k] : v[c][4 + k];  // even lane sends its upper half of A; odd lane sends lower half of B
            double got = shx(send);
            if (!odd) { oA[c][k] = v[c][k]; oB[c][k] = got; } else { oA[c][k] = got; oB[c][k] = v[c][4 + k]; }
        }
        for (int c = 0; c < 2; ++c) st32(YA + lineoff(t0 + t, cA) + c * 8 + odd * 4, oA[c]);
        for (int c = 0; c < 2; ++c) st32(YB + lineoff(t0 + t, cB) + c * 8 + odd * 4, oB[c]);
    }
    if (acc == 123.456) Y0[0] = acc;
}


// cp.async (LDGSTS) cooperative loads: 8 consecutive lanes fetch one 128-byte record (16 B each) into a
// padded per-warp shared-memory stage, NST stages deep; every lane then reads its own record with
// LDS.128.  COOP_ST=0: each lane stores its own record with 4 x 32-byte STG; COOP_ST=1: the record
// goes back through shared memory and is stored 8 lanes per record (fully coalesced lines).
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gsrc) : "memory");
}
template <int NST, int COOP_ST>
__global__ void pattern_cpasync(const double* __restrict__ X0, const double* __restrict__ X1, double* __restrict__ Y0, double* __restrict__ Y1,
                                const unsigned char* __restrict__ sel, int C, int nblk, int TPB, long long TT) {
    constexpr int RECP = 18;  // padded record: 144 B (odd multiple of 16 B => conflict-free LDS.128)
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* xin = sm + (size_t)warp * (NST + COOP_ST) * 32 * RECP;
    double* xout = xin + (size_t)NST * 32 * RECP;
    long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int blk = u / C, chain = u % C;
    const int chain0 = chain - lane;
    const int s_own = sel[u];
    const unsigned selmask = __ballot_sync(0xffffffffu, s_own);
    const long long t0 = (long long)blk * TPB;
    const long long stride = (long long)C * 16;
    const double* src[8]; double* dst[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int r = q * 4 + (lane >> 3), pc = lane & 7;
        const int s = (selmask >> r) & 1;
        const long long off = (t0 * C + chain0 + r) * 16 + pc * 2;
        src[q] = (s ? X1 : X0) + off;
        dst[q] = (s ? Y0 : Y1) + off;
    }
    double* yown = (s_own ? Y0 : Y1) + (t0 * C + chain) * 16;
    auto issue = [&](int t) {
        double* st = xin + (size_t)(t % NST) * 32 * RECP;
#pragma unroll
        for (int q = 0; q < 8; ++q) cp_async16(st + (q * 4 + (lane >> 3)) * RECP + (lane & 7) * 2, src[q] + t * stride);
    };
    for (int t = 0; t < NST - 1; ++t) { if (t < TPB) issue(t); asm volatile("cp.async.commit_group;" ::: "memory"); }
    double acc = 0;
    for (int t = 0; t < TPB; ++t) {
        asm volatile("cp.async.wait_group %0;" :: "n"(NST - 2) : "memory");
        __syncwarp();
        if (t + NST - 1 < TPB) issue(t + NST - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double* rec = xin + (size_t)(t % NST) * 32 * RECP + lane * RECP;
        double v[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) { const double2 d = *reinterpret_cast<const double2*>(rec + 2 * k); v[2 * k] = d.x; v[2 * k + 1] = d.y; }
#pragma unroll
        for (int k = 0; k < 16; ++k) { acc += v[k]; v[k] = v[k] * 1.0000001 + acc * 1e-30; }
        if (COOP_ST) {
            double* o = xout + lane * RECP;
#pragma unroll
            for (int k = 0; k < 8; ++k) *reinterpret_cast<double2*>(o + 2 * k) = make_double2(v[2 * k], v[2 * k + 1]);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double2 d = *reinterpret_cast<const double2*>(xout + (q * 4 + (lane >> 3)) * RECP + (lane & 7) * 2);
                *reinterpret_cast<double2*>(dst[q] + t * stride) = d;
            }
            __syncwarp();
        } else {
#pragma unroll
            for (int k = 0; k < 16; k += 4) st32(yown + t * stride + k, v + k);
        }
    }
    if (acc == 123.456) Y0[0] = acc;
}

// own-record pattern: every lane reads and writes its own 128-byte record with 4 x 32-byte accesses
template <int DEPTH>
__global__ void pattern_own(const double* __restrict__ X0, const double* __restrict__ X1, double* __restrict__ Y0, double* __restrict__ Y1,
                            const unsigned char* __restrict__ sel, int C, int nblk, int TPB, long long TT) {
    long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (long long)C * nblk) return;
    int blk = u / C, chain = u % C;
    const int s = sel[u];
    const long long t0 = (long long)blk * TPB, stride = (long long)C * 16;
    const double* x = (s ? X1 : X0) + (t0 * C + chain) * 16;
    double* y = (s ? Y0 : Y1) + (t0 * C + chain) * 16;
    double buf[DEPTH][16];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
#pragma unroll
        for (int k = 0; k < 16; k += 4) ld32(x + d * stride + k, buf[d] + k);
    double acc = 0;
    for (int t = 0; t < TPB; t += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            if (t + d >= TPB) break;
            double v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) { v[k] = buf[d][k]; acc += v[k]; }
            if (t + d + DEPTH < TPB) {
#pragma unroll
                for (int k = 0; k < 16; k += 4) ld32(x + (t + d + DEPTH) * stride + k, buf[d] + k);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = v[k] * 1.0000001 + acc * 1e-30;
#pragma unroll
            for (int k = 0; k < 16; k += 4) st32(y + (t + d) * stride + k, v + k);
        }
    }
    if (acc == 123.456) Y0[0] = acc;
}

// generalised record size: RB doubles per (tile, chain) record (16 -> 128 B, 32 -> 256 B, 64 -> 512 B)
template <int RB>
__global__ void pattern_cpasync_rb(const double* X0, const double* X1, double* Y0, double* Y1,
                                   const unsigned char* __restrict__ sel, int C, int nblk, int TPB, long long TT) {
    constexpr int RECP = RB + 2, NP = RB / 2, NST = 2;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* xin = sm + (size_t)warp * (NST + 1) * 32 * RECP;
    double* xout = xin + (size_t)NST * 32 * RECP;
    long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int blk = u / C, chain = u % C;
    const int chain0 = chain - lane;
    const int s_own = sel[u];
    const unsigned selmask = __ballot_sync(0xffffffffu, s_own);
    const long long t0 = (long long)blk * TPB;
    const long long stride = (long long)C * RB;
    auto issue = [&](int t) {
        double* st = xin + (size_t)(t % NST) * 32 * RECP;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int g = q * 32 + lane, r = g / NP, pc = g % NP;
            const int s = (selmask >> r) & 1;
            cp_async16(st + r * RECP + pc * 2, (s ? X1 : X0) + ((t0 + t) * C + chain0 + r) * RB + pc * 2);
        }
    };
    issue(0); asm volatile("cp.async.commit_group;" ::: "memory");
    double acc = 0;
    for (int t = 0; t < TPB; ++t) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (t + 1 < TPB) issue(t + 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double* rec = xin + (size_t)(t % NST) * 32 * RECP + lane * RECP;
        double* o = xout + lane * RECP;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            double2 d = *reinterpret_cast<const double2*>(rec + 2 * k);
            acc += d.x + d.y; d.x = d.x * 1.0000001 + acc * 1e-30; d.y = d.y * 1.0000001 + acc * 1e-30;
            *reinterpret_cast<double2*>(o + 2 * k) = d;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int g = q * 32 + lane, r = g / NP, pc = g % NP;
            const int s = (selmask >> r) & 1;
            const double2 d = *reinterpret_cast<const double2*>(xout + r * RECP + pc * 2);
            *reinterpret_cast<double2*>((s ? Y0 : Y1) + ((t0 + t) * C + chain0 + r) * RB + pc * 2) = d;
        }
        __syncwarp();
    }
    if (acc == 123.456) Y0[0] = acc;
}

__global__ void copyk(const double4* __restrict__ a, double4* __restrict__ b, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += st) b[i] = a[i];
}

int main() {
    const int C = 1024, nblk = 51, TPB = 26;
    const long long TT = (long long)nblk * TPB;
    const size_t n = (size_t)2 * TT * C * 8;
    double *X0, *X1, *Y0, *Y1; unsigned char* sel;
    CK(cudaMalloc(&X0, 2 * n * 8)); CK(cudaMemset(X0, 0, 2 * n * 8)); CK(cudaMalloc(&X1, n * 8)); CK(cudaMalloc(&Y0, n * 8)); CK(cudaMalloc(&Y1, n * 8));
    CK(cudaMemset(X0, 0, n * 8)); CK(cudaMemset(X1, 0, n * 8));
    CK(cudaMalloc(&sel, (size_t)C * nblk));
    unsigned char* hs = (unsigned char*)malloc((size_t)C * nblk);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes_r = (double)n * 8, bytes_w = (double)n * 8;
    for (int mode = 0; mode < 3; ++mode) {           // 0: uniform sel, write; 1: random sel, write; 2: random sel, read only
        for (size_t i = 0; i < (size_t)C * nblk; ++i) hs[i] = (mode == 0) ? 0 : (rand() & 1);
        CK(cudaMemcpy(sel, hs, (size_t)C * nblk, cudaMemcpyHostToDevice));
        for (int threads = 64; threads <= 128; threads *= 2) {
            int grid = (C * nblk + threads - 1) / threads;
            for (int depth = 1; depth <= 4; depth *= 2) {
                float best = 1e9;
                for (int rep = 0; rep < 6; ++rep) {
                    cudaEventRecord(e0);
                    if (depth == 1) pattern<1><<<grid, threads>>>(X0, X1, Y0, Y1, sel, C, nblk, TPB, TT, mode != 2);
                    if (depth == 2) pattern<2><<<grid, threads>>>(X0, X1, Y0, Y1, sel, C, nblk, TPB, TT, mode != 2);
                    if (depth == 4) pattern<4><<<grid, threads>>>(X0, X1, Y0, Y1, sel, C, nblk, TPB, TT, mode != 2);
                    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
                }
                double b = bytes_r + (mode != 2 ? bytes_w : 0);
                printf("mode %d threads %3d depth %d : %7.1f us  %6.0f GB/s (algorithmic)\n", mode, threads, depth, best * 1e3, b / best / 1e6);
            }
        }
    }
    for (int mode = 0; mode < 2; ++mode) {
        for (size_t i = 0; i < (size_t)C * nblk; ++i) hs[i] = (mode == 0) ? 0 : (rand() & 1);
        CK(cudaMemcpy(sel, hs, (size_t)C * nblk, cudaMemcpyHostToDevice));
        const int LPB = TPB / 2; const long long TL = (long long)nblk * LPB;
        for (int half = 0; half < 2; ++half) {
            int threads = 64, grid = (C * nblk + threads - 1) / threads;
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                if (half) pattern128<1><<<grid, threads>>>(X0, X1, Y0, Y1, sel, C, nblk, LPB, TL, 1);
                else pattern128<0><<<grid, threads>>>(X0, X1, Y0, Y1, sel, C, nblk, LPB, TL, 1);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
            }
            printf("128B-line layout, sel %s, %s: %7.1f us %6.0f GB/s (algorithmic)\n", mode ? "random" : "uniform",
                   half ? "two 64B halves" : "whole line", best * 1e3, 2.0 * n * 8 / best / 1e6);
        }
    }
    for (int mode = 0; mode < 2; ++mode) {
        for (size_t i = 0; i < (size_t)C * nblk; ++i) hs[i] = (mode == 0) ? 0 : (rand() & 1);
        CK(cudaMemcpy(sel, hs, (size_t)C * nblk, cudaMemcpyHostToDevice));
        for (int threads = 64; threads <= 128; threads *= 2) {
            int grid = (C * nblk + threads - 1) / threads;
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                pattern_pair<<<grid, threads>>>(X0, X1, Y0, Y1, sel, C, nblk, TPB, TT);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
            }
            printf("pair-cooperative 128B line, sel %s, threads %d: %7.1f us %6.0f GB/s (algorithmic)\n", mode ? "random" : "uniform",
                   threads, best * 1e3, 2.0 * n * 8 / best / 1e6);
        }
    }

    for (int mode = 0; mode < 2; ++mode) {
        for (size_t i = 0; i < (size_t)C * nblk; ++i) hs[i] = (mode == 0) ? 0 : (rand() & 1);
        CK(cudaMemcpy(sel, hs, (size_t)C * nblk, cudaMemcpyHostToDevice));
        for (int variant = 0; variant < 6; ++variant) {
            const int threads = 64, grid = (C * nblk + threads - 1) / threads;
            const int nst = (variant % 3) + 2, coop = variant / 3;
            const size_t smem = (size_t)(threads / 32) * (nst + coop) * 32 * 18 * 8;
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
#define RUNCP(N_, C_) do { CK(cudaFuncSetAttribute(pattern_cpasync<N_, C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                           pattern_cpasync<N_, C_><<<grid, threads, smem>>>(X0, X1, Y0, Y1, sel, C, nblk, TPB, TT); } while (0)
                if (variant == 0) RUNCP(2, 0); if (variant == 1) RUNCP(3, 0); if (variant == 2) RUNCP(4, 0);
                if (variant == 3) RUNCP(2, 1); if (variant == 4) RUNCP(3, 1); if (variant == 5) RUNCP(4, 1);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
            }
            printf("cp.async coop loads NST=%d, %s stores, sel %s: %7.1f us %6.0f GB/s (algorithmic)\n", nst, coop ? "coop" : "own-record",
                   mode ? "random" : "uniform", best * 1e3, 2.0 * n * 8 / best / 1e6);
        }
        for (int depth = 1; depth <= 2; ++depth) {
            const int threads = 64, grid = (C * nblk + threads - 1) / threads;
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                if (depth == 1) pattern_own<1><<<grid, threads>>>(X0, X1, Y0, Y1, sel, C, nblk, TPB, TT);
                else pattern_own<2><<<grid, threads>>>(X0, X1, Y0, Y1, sel, C, nblk, TPB, TT);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
            }
            printf("own-record 4x32B loads depth %d + own-record stores, sel %s: %7.1f us %6.0f GB/s (algorithmic)\n", depth,
                   mode ? "random" : "uniform", best * 1e3, 2.0 * n * 8 / best / 1e6);
        }
    }

    // in-place ping-pong (read container s, write container 1-s of the SAME allocation), containers
    // either split (two arrays) or interleaved per record ([tile][chain][container][16])
    for (int mode = 0; mode < 2; ++mode) {
        for (size_t i = 0; i < (size_t)C * nblk; ++i) hs[i] = (mode == 0) ? 0 : (rand() & 1);
        CK(cudaMemcpy(sel, hs, (size_t)C * nblk, cudaMemcpyHostToDevice));
        for (int inter = 0; inter < 2; ++inter) {
            const int threads = 64, grid = (C * nblk + threads - 1) / threads;
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                if (inter) pattern_pair<<<grid, threads>>>(X0, X0 + 16, X0, X0 + 16, sel, C, nblk, TPB, TT, 32);   // X0 and X1 are adjacent allocations? use X0 (2n doubles needed)
                else pattern_pair<<<grid, threads>>>(X0, X1, X0, X1, sel, C, nblk, TPB, TT, 16);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
            }
            printf("in-place pair-cooperative, %s containers, sel %s: %7.1f us %6.0f GB/s (algorithmic)\n", inter ? "interleaved" : "split",
                   mode ? "random" : "uniform", best * 1e3, 2.0 * n * 8 / best / 1e6);
        }
    }

    for (int mode = 0; mode < 2; ++mode) {
        for (size_t i = 0; i < (size_t)C * nblk; ++i) hs[i] = (mode == 0) ? 0 : (rand() & 1);
        CK(cudaMemcpy(sel, hs, (size_t)C * nblk, cudaMemcpyHostToDevice));
        for (int rb = 16; rb <= 64; rb *= 2) {
            const int threads = 64, grid = (C * nblk + threads - 1) / threads;
            const size_t smem = (size_t)(threads / 32) * 3 * 32 * (rb + 2) * 8;
            const int tpb = TPB * 16 / rb;   // 26, 13, 6 (6*64 < 26*16: slightly less data; bytes adjusted below)
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
#define RUNRB(R_) do { CK(cudaFuncSetAttribute(pattern_cpasync_rb<R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                       pattern_cpasync_rb<R_><<<grid, threads, smem>>>(X0, X1, X0, X1, sel, C, nblk, tpb, TT); } while (0)
                if (rb == 16) RUNRB(16); if (rb == 32) RUNRB(32); if (rb == 64) RUNRB(64);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
            }
            const double by = 2.0 * (double)C * nblk * tpb * rb * 8;
            printf("in-place cp.async coop, record %3d B, sel %s: %7.1f us %6.0f GB/s (algorithmic; %.0f MB)\n", rb * 8,
                   mode ? "random" : "uniform", best * 1e3, by / best / 1e6, by / 1e6);
        }
    }
    {
        float best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            copyk<<<148 * 16, 256>>>((const double4*)X0, (double4*)Y0, (long long)n / 4);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
        }
        printf("plain copy %zu MB: %7.1f us %6.0f GB/s\n", n * 8 / 1000000, best * 1e3, 2.0 * n * 8 / best / 1e6);
    }
    return 0;
}
