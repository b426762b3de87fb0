// FP64 tensor-core (DMMA) throughput on sm_100a: mma.sync m8n8k4 / m16n8k8 / m16n8k16 f64, NACC independent accumulator sets.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma scripts/micro/dmma.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k884(double *out, int iters, double a0, double b0) {
    double c[NACC][2];
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k1688(double *out, int iters, double a0, double b0) {
    double c[NACC][4];
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    double a[4] = {a0 + threadIdx.x, a0, a0 * 2, a0 - 1}, b[2] = {b0 - threadIdx.x, b0};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k16816(double *out, int iters, double a0, double b0) {
    double c[NACC][4];
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = a0 + i + threadIdx.x;
    for (int i = 0; i < 4; ++i) b[i] = b0 - i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void kfma(double *out, int iters, double a0, double b0) {
    double c[NACC];
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = a0 + threadIdx.x, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(a, c[i], b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
void run(const char *name, K kern, int threads, int ctas_per_sm, double flop_per_warp_iter, int iters) {
    int nsm;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * nsm * ctas_per_sm * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<nsm * ctas_per_sm, threads>>>(out, iters, 1e-9, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<nsm * ctas_per_sm, threads>>>(out, iters, 1e-9, 1e-9);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warps = (double)nsm * ctas_per_sm * threads / 32;
    printf("%-28s threads=%4d ctas/sm=%d  %.3f ms  %.2f TFLOP/s  err=%s\n", name, threads, ctas_per_sm, ms, warps * flop_per_warp_iter * iters / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    const int it = 20000;
    for (int thr : {128, 256, 512}) {
        run("m8n8k4  x1", k884<1>, thr, 2, 1 * 2.0 * 8 * 8 * 4, it);
        run("m8n8k4  x4", k884<4>, thr, 2, 4 * 2.0 * 8 * 8 * 4, it);
        run("m8n8k4  x8", k884<8>, thr, 2, 8 * 2.0 * 8 * 8 * 4, it);
        run("m16n8k8 x4", k1688<4>, thr, 2, 4 * 2.0 * 16 * 8 * 8, it);
        run("m16n8k8 x8", k1688<8>, thr, 2, 8 * 2.0 * 16 * 8 * 8, it);
        run("m16n8k16 x4", k16816<4>, thr, 2, 4 * 2.0 * 16 * 8 * 16, it);
        run("m16n8k16 x8", k16816<8>, thr, 2, 8 * 2.0 * 16 * 8 * 16, it);
        run("dfma x8", kfma<8>, thr, 2, 8 * 2.0 * 32, it);
        run("dfma x16", kfma<16>, thr, 2, 16 * 2.0 * 32, it);
    }
    return 0;
}
