// FP64 tensor-core (DMMA, mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 .f64) throughput and latency on this GPU, next
// to plain DFMA, so the Lorenz-96 kernel's roofline has a measured denominator for both pipes.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma dmma.cu && ./dmma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double *c, const double *a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double *c, const double *a, const double *b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double *c, const double *a, const double *b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// KIND 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16, 4: DFMA.  NACC independent accumulator sets per warp.
template <int KIND, int NACC>
__global__ void __launch_bounds__(1024) pipe(double *out, long long *cyc, int iters) {
    double c[NACC][4], a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    for (int i = 0; i < 4; ++i) b[i] = 1e-9 * (threadIdx.x + 3 * i);
    for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) c[n][i] = n + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < NACC; ++n) {
            if (KIND == 0) dmma884(c[n][0], c[n][1], a[0], b[0]);
            if (KIND == 1) dmma1684(c[n], a, b[0]);
            if (KIND == 2) dmma1688(c[n], a, b);
            if (KIND == 3) dmma16816(c[n], a, b);
            if (KIND == 4) {
#pragma unroll
                for (int i = 0; i < 4; ++i) c[n][i] = fma(c[n][i], b[0], a[0]);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) s += c[n][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int KIND, int NACC>
static void run(const char *name, int warps_per_sm, double fma_per_instr, double *out, long long *cyc) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        pipe<KIND, NACC><<<sms, warps_per_sm * 32>>>(out, cyc, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double instr = (double)sms * warps_per_sm * (double)iters * NACC;  // warp-level instructions (DFMA: 4 per slot)
    const double flops = instr * fma_per_instr * 2.0;
    printf("%-10s acc=%2d warps/SM=%2d  %8.2f TFLOP/s   %7.2f cycles per warp-instr (one warp's view)  err=%s\n", name, NACC, warps_per_sm,
           flops / (ms * 1e-3) / 1e12, (double)h / ((double)iters * NACC * (KIND == 4 ? 4 : 1)), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8 * 2); cudaMalloc(&cyc, 8);
    // latency: one warp, one accumulator
    run<0, 1>("m8n8k4", 1, 256, out, cyc);
    run<1, 1>("m16n8k4", 1, 512, out, cyc);
    run<2, 1>("m16n8k8", 1, 1024, out, cyc);
    run<3, 1>("m16n8k16", 1, 2048, out, cyc);
    run<4, 1>("dfma", 1, 4 * 32, out, cyc);
    for (int w : {4, 8, 16, 32}) {
        if (w == 4)  { run<0, 8>("m8n8k4", 4, 256, out, cyc); run<1, 8>("m16n8k4", 4, 512, out, cyc); run<2, 8>("m16n8k8", 4, 1024, out, cyc); run<3, 8>("m16n8k16", 4, 2048, out, cyc); run<4, 8>("dfma", 4, 128, out, cyc); }
        if (w == 8)  { run<0, 8>("m8n8k4", 8, 256, out, cyc); run<1, 8>("m16n8k4", 8, 512, out, cyc); run<2, 8>("m16n8k8", 8, 1024, out, cyc); run<3, 8>("m16n8k16", 8, 2048, out, cyc); run<4, 8>("dfma", 8, 128, out, cyc); }
        if (w == 16) { run<0, 8>("m8n8k4", 16, 256, out, cyc); run<1, 8>("m16n8k4", 16, 512, out, cyc); run<2, 8>("m16n8k8", 16, 1024, out, cyc); run<3, 8>("m16n8k16", 16, 2048, out, cyc); run<4, 8>("dfma", 16, 128, out, cyc); }
        if (w == 32) { run<0, 8>("m8n8k4", 32, 256, out, cyc); run<1, 8>("m16n8k4", 32, 512, out, cyc); run<2, 8>("m16n8k8", 32, 1024, out, cyc); run<3, 8>("m16n8k16", 32, 2048, out, cyc); run<4, 8>("dfma", 32, 128, out, cyc); }
    }
    // few accumulators (dependent chains) at 8 warps per SM: how many independent chains the pipe needs
    run<0, 1>("m8n8k4", 8, 256, out, cyc);
    run<0, 2>("m8n8k4", 8, 256, out, cyc);
    run<0, 4>("m8n8k4", 8, 256, out, cyc);
    run<0, 15>("m8n8k4", 8, 256, out, cyc);
    return 0;
}
