// Latency of a flag-and-payload hand-over between two GPUs of one box through peer memory (NVLink / NVSwitch), the
// primitive under dmcmc_halo_exchange's peer-memory path, and of a chain of tiny dependent kernels in one stream.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o p2p_latency p2p_latency.cu && ./p2p_latency
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// MODE 0: every thread __threadfence_system, then thread 0 fence + st.release (what the library did first)
// MODE 1: __syncthreads, thread 0: one __threadfence_system + relaxed store; poll relaxed, one fence after
// MODE 2: thread 0: st.release.sys only (release is cumulative over the CTA barrier); poll acquire
template <int MODE>
__global__ void pingpong(int me, int rounds, int payload, double *peer_box, const double *my_box, unsigned long long *peer_flag,
                         const unsigned long long *my_flag, double *sink) {
    double acc = 0.0;
    for (int i = 1; i <= rounds; ++i) {
        if (me == 0 || i > 0) {
            if (me == 1) {  // wait for the ping first
                if (threadIdx.x == 0) {
                    if (MODE == 1) { while (ld_relaxed_sys(my_flag) < (unsigned long long)i) {} __threadfence_system(); }
                    else { while (ld_acquire_sys(my_flag) < (unsigned long long)i) {} }
                }
                __syncthreads();
                for (int k = threadIdx.x; k < payload; k += blockDim.x) acc += __ldcg(my_box + k);
            }
            for (int k = threadIdx.x; k < payload; k += blockDim.x) peer_box[k] = (double)i + k;
            if (MODE == 0) __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0) {
                if (MODE == 0) { __threadfence_system(); st_release_sys(peer_flag, (unsigned long long)i); }
                if (MODE == 1) { __threadfence_system(); st_relaxed_sys(peer_flag, (unsigned long long)i); }
                if (MODE == 2) st_release_sys(peer_flag, (unsigned long long)i);
            }
            if (me == 0) {  // wait for the pong
                if (threadIdx.x == 0) {
                    if (MODE == 1) { while (ld_relaxed_sys(my_flag) < (unsigned long long)i) {} __threadfence_system(); }
                    else { while (ld_acquire_sys(my_flag) < (unsigned long long)i) {} }
                }
                __syncthreads();
                for (int k = threadIdx.x; k < payload; k += blockDim.x) acc += __ldcg(my_box + k);
            }
        }
    }
    sink[threadIdx.x] = acc;
}

__global__ void tiny(double *x) { if (threadIdx.x == 0) x[0] += 1.0; }
__global__ void tiny_pdl(double *x) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) x[0] += 1.0;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs two GPUs\n"); return 0; }
    double *box[2], *sink[2];
    unsigned long long *flag[2];
    cudaStream_t st[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&box[d], 1 << 20));
        CK(cudaMalloc(&sink[d], 4096));
        CK(cudaMalloc(&flag[d], 256));
        CK(cudaStreamCreate(&st[d]));
    }
    const int rounds = 2000;
    for (int mode = 0; mode < 3; ++mode)
        for (int payload : {1, 150, 1500, 15000}) {
            for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaMemset(flag[d], 0, 256)); CK(cudaDeviceSynchronize()); }
            cudaEvent_t e0, e1;
            CK(cudaSetDevice(0));
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            CK(cudaEventRecord(e0, st[0]));
            for (int d = 0; d < 2; ++d) {
                CK(cudaSetDevice(d));
                if (mode == 0) pingpong<0><<<1, 128, 0, st[d]>>>(d, rounds, payload, box[1 - d], box[d], flag[1 - d], flag[d], sink[d]);
                if (mode == 1) pingpong<1><<<1, 128, 0, st[d]>>>(d, rounds, payload, box[1 - d], box[d], flag[1 - d], flag[d], sink[d]);
                if (mode == 2) pingpong<2><<<1, 128, 0, st[d]>>>(d, rounds, payload, box[1 - d], box[d], flag[1 - d], flag[d], sink[d]);
            }
            CK(cudaSetDevice(0));
            CK(cudaEventRecord(e1, st[0]));
            CK(cudaEventSynchronize(e1));
            CK(cudaSetDevice(1));
            CK(cudaDeviceSynchronize());
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("mode %d payload %6d doubles: one-way hand-over %.2f us (round trip / 2, %d rounds, one CTA of 128 threads per GPU)\n", mode, payload,
                   ms * 1e3 / rounds / 2, rounds);
        }
    // chains of tiny dependent kernels in one stream
    CK(cudaSetDevice(0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int pdl = 0; pdl < 2; ++pdl) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0, st[0]));
            for (int i = 0; i < 1000; ++i) {
                if (!pdl) tiny<<<1, 128, 0, st[0]>>>(sink[0]);
                else {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(1); cfg.blockDim = dim3(128); cfg.stream = st[0];
                    cudaLaunchAttribute attr[1];
                    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    attr[0].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = attr; cfg.numAttrs = 1;
                    CK(cudaLaunchKernelEx(&cfg, tiny_pdl, sink[0]));
                }
            }
            CK(cudaEventRecord(e1, st[0]));
            CK(cudaEventSynchronize(e1));
        }
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("1000 dependent tiny kernels in one stream, %s: %.2f us per kernel\n", pdl ? "programmatic dependent launch" : "plain launches", ms);
    }
    return 0;
}
