// shared-memory FP64 atomicAdd vs plain read-modify-write, distinct address per lane (B200): nvcc -arch=sm_100a -O3 satom.cu -o satom
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *out, int iters) {
    extern __shared__ double sm[];
    const int t = threadIdx.x;
    for (int i = t; i < 150 * 32; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    const int lane = t & 31, warp = t >> 5;
    double v = 1.0 + lane * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int a = 0; a < 150; a += 3) {
            const int row = (a + 7 * warp) % 150;
            double *p = sm + row * 32 + lane;
            if (MODE == 0) { *p += v; }                      // plain RMW (racy across warps: timing only)
            else atomicAdd(p, v);
            v = v * 1.0000001;
        }
    }
    __syncthreads();
    if (t < 32) out[blockIdx.x * 32 + t] = sm[t] + v;
}
int main() {
    double *out; cudaMalloc(&out, 148 * 32 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, 512, 150 * 32 * 8>>>(out, 2000); else k<1><<<148, 512, 150 * 32 * 8>>>(out, 2000);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("%s: %.3f ms  -> %.2f ns per warp-op per SM\n", mode ? "atomicAdd(double) shared" : "plain RMW", ms, ms * 1e6 / (2000.0 * 50 * 16));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
