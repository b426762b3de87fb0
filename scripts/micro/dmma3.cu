// DMMA.8x8x4: does guarding (warp-uniform predicates) or feeding operands from shared memory cost throughput?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma3 scripts/micro/dmma3.cu
#include <cstdio>
#include <cuda_runtime.h>
#define DMMA(c, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b))

template <int MODE>
__global__ void kern(double *out, int iters, const double *in, int len0, int len1) {
    __shared__ double sm[32 * 68];
    for (int i = threadIdx.x; i < 32 * 68; i += blockDim.x) sm[i] = in[i & 255];
    __syncthreads();
    double c0[8][2], c1[4][2];
    for (int i = 0; i < 8; ++i) c0[i][0] = c0[i][1] = 0.0;
    for (int i = 0; i < 4; ++i) c1[i][0] = c1[i][1] = 0.0;
    const int lane = threadIdx.x & 31;
    const double *fb0 = sm + (lane & 3) * 68 + (lane >> 2);
    double a0 = in[lane], a1 = in[lane + 32], bf[8];
    for (int i = 0; i < 8; ++i) bf[i] = in[lane + 8 * i];
    for (int it = 0; it < iters; ++it) {
        const double *fb = fb0 + 4 * 68 * (it & 7);
        if (MODE >= 2) {
            a0 = fb[0];
            a1 = fb[56];
#pragma unroll
            for (int i = 0; i < 8; ++i) bf[i] = fb[8 * i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 2) {  // unguarded: 8 + 4 DMMAs
                DMMA(c0[i], a0, bf[i]);
                if (i < 4) DMMA(c1[i & 3], a1, bf[i]);
            } else {  // guarded as in the RESP kernel
                if (i < len0) DMMA(c0[i], a0, bf[i]);
                if (i < 4 && i < len1) DMMA(c1[i & 3], a1, bf[i]);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c0[i][0] + c0[i][1];
    for (int i = 0; i < 4; ++i) s += c1[i][0] + c1[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
void run(const char *name, K k, int ctas_per_sm, const double *in, int len0, int len1) {
    int nsm;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    const int threads = 128, iters = 20000;
    cudaMalloc(&out, sizeof(double) * nsm * ctas_per_sm * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<<<nsm * ctas_per_sm, threads>>>(out, iters, in, len0, len1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<nsm * ctas_per_sm, threads>>>(out, iters, in, len0, len1);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warps = (double)nsm * ctas_per_sm * threads / 32;
    printf("%-40s warps/SM=%2d len=%d+%d  %.3f ms  %.2f TFLOP/s  %s\n", name, ctas_per_sm * 4, len0, len1, ms,
           warps * (len0 + len1) * 512.0 * iters / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    double h[256];
    for (int i = 0; i < 256; ++i) h[i] = 1e-3 * (i + 1);
    double *in;
    cudaMalloc(&in, sizeof(h));
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int cps : {2, 5}) {
        run("registers, unguarded", kern<0>, cps, in, 8, 4);
        run("registers, guarded", kern<1>, cps, in, 8, 4);
        run("registers, guarded", kern<1>, cps, in, 6, 3);
        run("shared-memory operands, unguarded", kern<2>, cps, in, 8, 4);
        run("shared-memory operands, guarded", kern<3>, cps, in, 8, 4);
        run("shared-memory operands, guarded", kern<3>, cps, in, 6, 3);
        run("shared-memory operands, guarded", kern<3>, cps, in, 5, 4);
    }
    return 0;
}
