// Dependent-issue latency of FP64 FMA / ADD / MUL on this GPU (one warp, one SM): cycles per dependent op.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64lat fp64lat.cu && ./fp64lat
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double a, double b, double *out, long long *cyc) {
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 4096; ++i) {
        if (OP == 0) x = fma(x, b, a);
        else if (OP == 1) x = x + b;
        else x = x * b;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    const char *names[] = {"DFMA", "DADD", "DMUL"};
    for (int op = 0; op < 3; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            if (op == 0) chain<0><<<1, 32>>>(1.0, 0.999999, out, cyc);
            if (op == 1) chain<1><<<1, 32>>>(1.0, 0.999999, out, cyc);
            if (op == 2) chain<2><<<1, 32>>>(1.0, 0.999999, out, cyc);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        }
        printf("%s dependent chain: %.2f cycles per op\n", names[op], h / 4096.0);
    }
    return 0;
}
