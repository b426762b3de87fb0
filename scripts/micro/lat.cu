// Latency / throughput microbenchmarks for the FP64 pipe of B200 (dependent chains, 1..8 independent chains, warps per SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double *out, int iters, double b, double c, long long *cyc) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void rsq_chain(double *out, int iters, long long *cyc) {
    double a = 1.0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            double y;
            asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
            a = y + 1.0;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
    const int iters = 1000;
#define RUN(ILP, warps) { chain<ILP><<<1, 32 * warps>>>(out, iters, 1.0000001, 1e-9, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("DFMA ILP=%d warps/SM=%2d: %.2f cycles per dependent step (%.2f DFMA/cycle/SM)\n", ILP, warps, (double)h / (iters * 16.0), ILP * warps * iters * 16.0 / h); }
    RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1)
    RUN(1, 4) RUN(2, 4) RUN(4, 4) RUN(8, 4)
    RUN(1, 8) RUN(2, 8) RUN(1, 16) RUN(2, 16) RUN(4, 16) RUN(1, 32)
    rsq_chain<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("MUFU.RSQ64H + DADD dependent: %.2f cycles per step\n", (double)h / (iters * 16.0));
    return 0;
}
