// DMMA.8x8x4 throughput versus operand pattern (same / distinct A and B registers per instruction, accumulator count).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma2 scripts/micro/dmma2.cu
#include <cstdio>
#include <cuda_runtime.h>

#define DMMA(c, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b))

template <int NACC, int MODE>
__global__ void kern(double *out, int iters, const double *in) {
    double c[NACC][2];
    double a[NACC], b[NACC];
    for (int i = 0; i < NACC; ++i) {
        c[i][0] = c[i][1] = 0.0;
        a[i] = in[(threadIdx.x + i) & 255];
        b[i] = in[(threadIdx.x + 2 * i + 7) & 255];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) DMMA(c[i], a[0], b[0]);        // same A, same B
            if (MODE == 1) DMMA(c[i], a[0], b[i]);        // same A, distinct B
            if (MODE == 2) DMMA(c[i], a[i], b[i]);        // distinct A and B
            if (MODE == 3) DMMA(c[i], a[i & 1], b[i >> 1]);  // the RESP pattern: two A rows share each B
        }
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
void run(const char *name, K k, int threads, int ctas_per_sm, int nacc, const double *in) {
    int nsm;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * nsm * ctas_per_sm * threads);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<<<nsm * ctas_per_sm, threads>>>(out, iters, in);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<nsm * ctas_per_sm, threads>>>(out, iters, in);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warps = (double)nsm * ctas_per_sm * threads / 32;
    printf("%-34s warps/SM=%2d  %.3f ms  %.2f TFLOP/s  %s\n", name, ctas_per_sm * threads / 32, ms, warps * nacc * 512.0 * iters / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    double h[256];
    for (int i = 0; i < 256; ++i) h[i] = 1e-3 * (i + 1);
    double *in;
    cudaMalloc(&in, sizeof(h));
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int cps : {1, 2, 4}) {
        run("x8  same A same B", kern<8, 0>, 128, cps, 8, in);
        run("x8  same A distinct B", kern<8, 1>, 128, cps, 8, in);
        run("x8  distinct A distinct B", kern<8, 2>, 128, cps, 8, in);
        run("x8  2 A x 4 B (RESP pattern)", kern<8, 3>, 128, cps, 8, in);
        run("x12 same A distinct B", kern<12, 1>, 128, cps, 12, in);
        run("x2  distinct", kern<2, 2>, 128, cps, 2, in);
        run("x1", kern<1, 0>, 128, cps, 1, in);
    }
    return 0;
}
