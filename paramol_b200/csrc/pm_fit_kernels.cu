// pm_fit_kernels.cu -- sm_100a kernels for the two linear fitting paths of ParaMol, C ABI in include/paramol_b200.h.
//
//   K5  pm_resp_assemble   RESP normal equations: A = sum_m w_m R_m R_m^T, B = sum_m w_m R_m V_m, C = sum_m w_m V_m.V_m with
//                          R_m[j][i] = 1 / |x_mj - g_mi| generated on the fly (never stored in HBM).
//                          Replaces the Python loops of ParaMol/MM_engines/resp.py:99-134 (inverse distances),
//                          :381-421 (A_aux), :463-496 (B_aux).
//       pm_esp_potential   V_mm[m][i] = sum_j q_j / r_ij   (ParaMolSystem.get_esp_ensemble, ParaMol/System/system.py:356-382)
//   K3  pm_lls_internal    bond lengths / angles / dihedrals of every conformation with the conventions of
//                          ParaMol/Utils/geometry.py:4-107 (used by LinearLeastSquare._calculate_a, least_square.py:469-649)
//       pm_lls_columns     design-matrix columns 1/2 (q - q0)^2 and 1 + cos(n phi - delta)  (least_square.py:492-645)
//
// All FP64 (the reference is numpy float64 throughout).  The RESP kernel is a batched rank-G update of an N x N matrix:
// each CTA keeps the accumulators of the upper triangle in registers (8x8 DMMA tiles, two tile rows per warp) across all
// the (conformation, grid chunk) work items it processes and writes ONE partial at the end; a second kernel sums the
// partials in a fixed order (deterministic, no atomics).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/paramol_b200.h"
#include "pm_device.cuh"

int pm_fail_msg(const std::string &msg);  // pm_kernels.cu

#define PMF_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) return pm_fail_msg(std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define RESP_GT 32          // grid points per shared-memory tile (8 DMMA k-steps of 4)
#define RESP_CHUNK 512      // grid points per work item
#define RESP_MAX_CTAS_PER_SM 6
#ifndef RESP_SKIP_GEN
#define RESP_SKIP_GEN 0  // timing experiments only
#endif
#ifndef RESP_SKIP_UPD
#define RESP_SKIP_UPD 0
#endif
#ifndef RESP_GEN_UNROLL
#define RESP_GEN_UNROLL 4
#endif
constexpr int kRespGenUnroll = RESP_GEN_UNROLL;
#ifndef RESP_WARPS_PER_SM
#define RESP_WARPS_PER_SM 20  // resident warps per SM the variants for up to 63 atoms are compiled for (register cap 96)
#endif

// ---------------------------------------------------------------------------------------------------------------------
// K5
// ---------------------------------------------------------------------------------------------------------------------
// A = sum R R^T is a symmetric rank-k update, so it runs on the FP64 tensor-core path: mma.sync m8n8k4 (SASS DMMA.8x8x4), which
// on B200 has the same peak as scalar DFMA (measured 37.1 TFLOP/s both, scripts/micro/dmma.cu) but needs one operand double per
// lane per 8x8x4 product instead of eight per 16 FMAs -- the scalar 4x4 register-tile version of this kernel was bound by
// shared-memory bandwidth, not by the FP64 pipe.
//
// The padded matrix ((N+1) rows: N atoms + the reference potential V as an extra row, so that column N of the product is
// B = sum R V and entry [N][N] is C = sum V.V) is cut into 8x8 tiles; only tiles (I, J >= I) are computed.  Warp u of a k-slice
// owns tile rows i0 = u and i1 = NB8-1-u (a short row paired with a long one: every warp has NB8+1 tiles), keeps their
// accumulators in registers for the whole kernel and walks the columns from the right, so that one B fragment feeds both rows
// and all register indices are static.  For both operands the fragment of tile row I at k-step k4 is
// Rt[4 k4 + (lane & 3)][8 I + (lane >> 2)]; the row stride of Rt is = 4 (mod 8) doubles, which makes that load conflict-free.
__device__ __forceinline__ void pm_dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Writes (first) or adds one 8x8 accumulator tile into a CTA's partial [N*N A | N B | C]; kept out of line so that the unrolled
// epilogue does not inflate the register budget of the main loop.
__device__ __noinline__ void pm_resp_emit(double c0, double c1, int I, int J, bool first, double *__restrict__ out, int N, int lane) {
    const int r = 8 * I + (lane >> 2);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int col = 8 * J + 2 * (lane & 3) + e;
        if (r > col) continue;  // lower part of a diagonal tile
        double *p0 = nullptr, *p1 = nullptr;
        if (col < N) {
            p0 = out + (size_t)r * N + col;
            if (r != col) p1 = out + (size_t)col * N + r;
        } else if (col == N) {
            p0 = out + (size_t)N * N + r;  // r < N: B[r];  r == N: C
        }
        if (p0) {
            const double c = e ? c1 : c0;
            const double v = first ? c : *p0 + c;
            *p0 = v;
            if (p1) *p1 = v;
        }
    }
}

template <int NB8, int KS>  // NB8 8-wide tile columns ((N + 1) padded to 8 * NB8 rows); KS k-slices (warps per tile row pair)
__global__ void __launch_bounds__(32 * ((NB8 + 1) / 2) * KS, NB8 <= 8 ? RESP_WARPS_PER_SM / (((NB8 + 1) / 2) * KS) : 1)
    pm_resp_kernel(int N, const double *__restrict__ coords, const double *__restrict__ grid, const double *__restrict__ vref,
                   const double *__restrict__ wt, long S, int G, long n_items, int chunks_per_conf, double *__restrict__ partial) {
    constexpr int NP = 8 * NB8, RS = NP + 4, NS = NB8 + 1;  // NS accumulator tiles ("slots") per warp
    constexpr int U = (NB8 + 1) / 2;                        // warps per k-slice
    const int XS = (3 * N + 1) & ~1;
    extern __shared__ __align__(16) double sm[];
    double *xs = sm;                       // [2][XS]          atom coordinates / sqrt(w), double-buffered over work items
    double *gs = xs + 2 * XS;              // [2][RESP_GT][4]  grid point (x, y, z)/sqrt(w) and sqrt(w) V, double-buffered over steps
    double *Rt = gs + 2 * 4 * RESP_GT;     // [2][RESP_GT][RS] the R tile, double-buffered over steps
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int ks = warp / U, u = warp - ks * U;
    // Slot s < len0 is tile (i0, NB8-1-s); slot s >= len0 is tile (i1, NB8-1-(s-len0)).  For the middle row of an odd NB8
    // (i1 == i0) the second half of the slots is fed zeros and never written out.
    const int i0 = u, i1 = NB8 - 1 - u;
    const int len0 = NB8 - i0;
    const bool has1 = i1 > i0;

    double acc[NS][2];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s][0] = acc[s][1] = 0.0;

    // generation: thread -> (atom row j, grid-point group kg); rows are strided when there are more rows than threads
    const int kgroups = T >= N ? T / N : 1;
    const int j0 = T >= N ? tid % N : tid, kg = T >= N ? tid / N : 0, jstride = T >= N ? N : T;
    const bool gen_active = kg < kgroups;
    const int frag_off = (lane & 3) * RS + (lane >> 2);

    // A work item is (conformation m, chunk of RESP_CHUNK grid points); a step is RESP_GT grid points of an item.  Items whose
    // weight is (numerically) zero are skipped; the scaling by 1/sqrt(w) of both coordinate sets makes rsqrt return sqrt(w)/r.
    auto next_item = [&](long it) {
        for (; it < n_items; it += gridDim.x) {
            const double w = wt ? wt[it / chunks_per_conf] : 1.0;
            if (!(w < 1e-280)) break;  // NaN weights are processed (and poison the sums, as they do in the reference)
        }
        return it;
    };
    auto stage_x = [&](long it, int xb) {
        const long m = it / chunks_per_conf;
        const double isw = wt ? 1.0 / sqrt(wt[m]) : 1.0;
        for (int i = tid; i < 3 * N; i += T) xs[xb * XS + i] = isw * coords[(size_t)m * 3 * N + i];
    };
    // threads 0..RESP_GT-1 fetch one grid point each (raw values: nothing may depend on the loads yet) ...
    auto load_g = [&](long it, int t0, double (&r)[4]) {
        const long m = it / chunks_per_conf;
        const int g1 = min(G, ((int)(it - m * chunks_per_conf) + 1) * RESP_CHUNK);
        const int i = min(t0 + tid, g1 - 1);
        const double *gp = grid + ((size_t)m * G + i) * 3;
        r[0] = gp[0];
        r[1] = gp[1];
        r[2] = gp[2];
        r[3] = vref[(size_t)m * G + i];
    };
    // ... and publish it, scaled, after the loads have had a generation phase to land; points beyond the chunk are pushed
    // 1e150 away (R = 0 for them)
    auto store_g = [&](long it, int t0, const double (&r)[4], int gb) {
        const long m = it / chunks_per_conf;
        const int g1 = min(G, ((int)(it - m * chunks_per_conf) + 1) * RESP_CHUNK);
        const double w = wt ? wt[m] : 1.0, sw = sqrt(w), isw = 1.0 / sw;
        const bool in = t0 + tid < g1;
        double *d = gs + gb * 4 * RESP_GT + 4 * tid;
        *reinterpret_cast<double2 *>(d) = make_double2(in ? isw * r[0] : 1e150, in ? isw * r[1] : 1e150);
        *reinterpret_cast<double2 *>(d + 2) = make_double2(in ? isw * r[2] : 1e150, in ? sw * r[3] : 0.0);
    };

    for (int i = tid; i < 2 * RESP_GT * RS; i += T) Rt[i] = 0.0;  // the padding rows stay zero for the whole kernel
    long item = next_item(blockIdx.x);
    int t0 = 0, g1 = 0, xb = 0, b = 0;
    if (item < n_items) {
        const long m = item / chunks_per_conf;
        t0 = (int)(item - m * chunks_per_conf) * RESP_CHUNK;
        g1 = min(G, t0 + RESP_CHUNK);
        stage_x(item, 0);
        if (tid < RESP_GT) {
            double r[4];
            load_g(item, t0, r);
            store_g(item, t0, r, 0);
        }
    }
    __syncthreads();
    while (item < n_items) {
        // where the next step is; its grid points are requested now and stored after the generation phase
        long n_item = item;
        int n_t0 = t0 + RESP_GT, n_g1 = g1;
        if (n_t0 >= g1) {
            n_item = next_item(item + gridDim.x);
            if (n_item < n_items) {
                const long m = n_item / chunks_per_conf;
                n_t0 = (int)(n_item - m * chunks_per_conf) * RESP_CHUNK;
                n_g1 = min(G, n_t0 + RESP_CHUNK);
            }
        }
        double gnext[4];
        if (n_item < n_items && tid < RESP_GT) load_g(n_item, n_t0, gnext);
        // (A) generate the R tile of this step: atom rows by everybody, the V row by the first warp
        {
            const double *g = gs + b * 4 * RESP_GT, *x = xs + xb * XS;
            double *rt = Rt + b * RESP_GT * RS;
            if (gen_active && !RESP_SKIP_GEN) {
#pragma unroll 1
                for (int j = j0; j < N; j += jstride) {
                    double *dst = rt + j;
                    const double xj = x[3 * j], yj = x[3 * j + 1], zj = x[3 * j + 2];
#pragma unroll kRespGenUnroll
                    for (int k = kg; k < RESP_GT; k += kgroups) {
                        const double2 gxy = *reinterpret_cast<const double2 *>(g + 4 * k);
                        const double dx = xj - gxy.x, dy = yj - gxy.y, dz = zj - g[4 * k + 2];
                        dst[k * RS] = pm_rsqrt(fma(dz, dz, fma(dy, dy, dx * dx)));  // = sqrt(w) / |x - g|
                    }
                }
            }
            if (tid < RESP_GT) rt[tid * RS + N] = g[4 * tid + 3];
        }
        // (B) stage the inputs of the next step into the other buffers
        if (n_item < n_items) {
            if (n_item != item) stage_x(n_item, xb ^ 1);
            if (tid < RESP_GT) store_g(n_item, n_t0, gnext, b ^ 1);
        }
        __syncthreads();
        // (C) rank-RESP_GT update of this warp's two tile rows
        // Every warp issues exactly NS unguarded DMMAs per k-step (a predicated-off DMMA costs as much as a live one:
        // scripts/micro/dmma3.cu); which of its two A fragments and which column a slot uses is warp-uniform arithmetic.
        if (!RESP_SKIP_UPD) {
            constexpr int Q = (RESP_GT / 4) / KS;  // k-steps of this warp per tile
            const double *fb0 = Rt + b * RESP_GT * RS + frag_off + 4 * ks * RS;
#pragma unroll NB8 <= 8 ? 2 : 1
            for (int q = 0; q < Q; ++q) {
                const double *fb = fb0 + 4 * KS * RS * q;
                const double a0 = fb[8 * i0], a1 = has1 ? fb[8 * i1] : 0.0;
                double bf[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) bf[s] = fb[8 * (NB8 - 1 - s) + (s >= len0 ? 8 * len0 : 0)];
#pragma unroll
                for (int s = 0; s < NS; ++s) pm_dmma(acc[s], s >= len0 ? a1 : a0, bf[s]);
            }
        }
        if (n_item != item) xb ^= 1;
        item = n_item;
        t0 = n_t0;
        g1 = n_g1;
        b ^= 1;
    }
    // one partial per CTA: [N*N] A (both triangles), [N] B, [1] C.  The k-slices add theirs one after the other (fixed order).
    double *out = partial + (size_t)blockIdx.x * ((size_t)N * N + N + 1);
    for (int round = 0; round < KS; ++round) {
        if (round) __syncthreads();
        if (ks == round) {
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) {
                const bool second = sl >= len0;
                if (!second || has1) pm_resp_emit(acc[sl][0], acc[sl][1], second ? i1 : i0, NB8 - 1 - (second ? sl - len0 : sl), round == 0, out, N, lane);
            }
        }
    }
}

__global__ void pm_sum_partials_kernel(const double *__restrict__ partial, int n_partial, long stride, double *__restrict__ out) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= stride) return;
    double s = 0.0;
    for (int p = 0; p < n_partial; ++p) s += partial[(size_t)p * stride + e];
    out[e] = s;
}

__global__ void pm_esp_potential_kernel(int N, const double *__restrict__ coords, const double *__restrict__ grid,
                                        const double *__restrict__ charges, long S, int G, double *__restrict__ vmm) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * (long)G) return;
    const long m = idx / G;
    const double gx = grid[3 * idx], gy = grid[3 * idx + 1], gz = grid[3 * idx + 2];
    const double *x = coords + (size_t)m * 3 * N;
    double v = 0.0;
    for (int j = 0; j < N; ++j) {
        const double dx = x[3 * j] - gx, dy = x[3 * j + 1] - gy, dz = x[3 * j + 2] - gz;
        v = fma(charges[j], pm_rsqrt(fma(dz, dz, fma(dy, dy, dx * dx))), v);
    }
    vmm[idx] = v;
}

// ---------------------------------------------------------------------------------------------------------------------
// K3
// ---------------------------------------------------------------------------------------------------------------------
__global__ void pm_lls_internal_kernel(int N, const double *__restrict__ coords, long S, int n_terms, const int *__restrict__ kind,
                                       const int *__restrict__ atoms, double *__restrict__ q_out) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * (long)n_terms) return;
    const long s = idx / n_terms;
    const int t = (int)(idx - s * n_terms);
    const double *x = coords + (size_t)s * 3 * N;
    const int *a = atoms + 4 * t;
    auto ld = [&](int i) { return pm_d3{x[3 * i], x[3 * i + 1], x[3 * i + 2]}; };
    double v;
    if (kind[t] == 0) {  // calculate_distance, geometry.py:4-41
        const pm_d3 d = pm_sub(ld(a[1]), ld(a[0]));
        v = sqrt(pm_dot(d, d));
    } else if (kind[t] == 1) {  // calculate_angle(v1, v2) with unit vectors and clipping, geometry.py:43-63
        const pm_d3 v1 = pm_sub(ld(a[0]), ld(a[1])), v2 = pm_sub(ld(a[2]), ld(a[1]));
        const double n1 = sqrt(pm_dot(v1, v1)), n2 = sqrt(pm_dot(v2, v2));
        const pm_d3 u1{v1.x / n1, v1.y / n1, v1.z / n1}, u2{v2.x / n2, v2.y / n2, v2.z / n2};
        v = acos(fmin(1.0, fmax(-1.0, pm_dot(u1, u2))));
    } else {  // calculate_dihedral ("praxeolitic" atan2 formula), geometry.py:66-107
        const pm_d3 p0 = ld(a[0]), p1 = ld(a[1]), p2 = ld(a[2]), p3 = ld(a[3]);
        const pm_d3 b0 = pm_sub(p0, p1);
        pm_d3 b1 = pm_sub(p2, p1);
        const pm_d3 b2 = pm_sub(p3, p2);
        const double n1 = sqrt(pm_dot(b1, b1));
        b1 = pm_d3{b1.x / n1, b1.y / n1, b1.z / n1};
        const double d0 = pm_dot(b0, b1), d2 = pm_dot(b2, b1);
        const pm_d3 vv{b0.x - d0 * b1.x, b0.y - d0 * b1.y, b0.z - d0 * b1.z};
        const pm_d3 ww{b2.x - d2 * b1.x, b2.y - d2 * b1.y, b2.z - d2 * b1.z};
        const double xx = pm_dot(vv, ww);
        const double yy = pm_dot(pm_cross(b1, vv), ww);
        v = atan2(yy, xx);
    }
    q_out[idx] = v;
}

__global__ void pm_lls_columns_kernel(const double *__restrict__ q, long S, int n_terms, int n_cols, const int *__restrict__ col_term,
                                      const int *__restrict__ col_kind, const double *__restrict__ col_p0,
                                      const int *__restrict__ col_per, double *__restrict__ A) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * (long)n_cols) return;
    const long s = idx / n_cols;
    const int c = (int)(idx - s * n_cols);
    const double v = q[(size_t)s * n_terms + col_term[c]];
    const double p0 = col_p0[c];
    double out;
    if (col_kind[c] == 2)
        out = 1.0 + cos((double)col_per[c] * v - p0);
    else
        out = 0.5 * (v - p0) * (v - p0);
    A[idx] = out;
}

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
struct RespLaunch {
    int ks, threads, grid;
    size_t smem;
    const void *fn;
};

// Launch geometry for an N-atom problem (independent of the data, so that pm_resp_scratch_doubles can size the partials).
static int resp_launch(int N, long n_items, RespLaunch *L) {
    const int NP = (N + 1 + 7) & ~7, NB8 = NP / 8, U = (NB8 + 1) / 2;
    if (NB8 > 22) return -1;
    L->ks = U >= 4 ? 1 : U == 1 ? 4 : 2;  // small molecules: split the 8 k-steps of a tile over several warps per tile row
    L->threads = 32 * U * L->ks;
    L->smem = sizeof(double) * 2 * (((3 * N + 1) & ~1) + 4 * RESP_GT + (size_t)RESP_GT * (NP + 4));
    switch (NB8) {
#define RESP_CASE(B, K) case B: L->fn = (const void *)pm_resp_kernel<B, K>; break;
        RESP_CASE(1, 4) RESP_CASE(2, 4) RESP_CASE(3, 2) RESP_CASE(4, 2) RESP_CASE(5, 2) RESP_CASE(6, 2) RESP_CASE(7, 1) RESP_CASE(8, 1)
        RESP_CASE(9, 1) RESP_CASE(10, 1) RESP_CASE(11, 1) RESP_CASE(12, 1) RESP_CASE(13, 1) RESP_CASE(14, 1) RESP_CASE(15, 1)
        RESP_CASE(16, 1) RESP_CASE(17, 1) RESP_CASE(18, 1) RESP_CASE(19, 1) RESP_CASE(20, 1) RESP_CASE(21, 1) RESP_CASE(22, 1)
#undef RESP_CASE
        default: return -1;
    }
    int dev = 0, n_sm = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (L->smem > 48 * 1024 && cudaFuncSetAttribute(L->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L->smem) != cudaSuccess) return -2;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, L->fn, L->threads, L->smem) != cudaSuccess || per_sm < 1) return -2;
    if (per_sm > RESP_MAX_CTAS_PER_SM) per_sm = RESP_MAX_CTAS_PER_SM;
    long g = (long)n_sm * per_sm;
    if (g > n_items) g = n_items;
    L->grid = (int)(g < 1 ? 1 : g);
    return 0;
}

extern "C" long pm_resp_scratch_doubles(int n_atoms, long n_conf, int n_grid) {
    const long chunks = (n_grid + RESP_CHUNK - 1) / RESP_CHUNK;
    RespLaunch L;
    if (n_atoms < 1 || resp_launch(n_atoms, n_conf * chunks, &L) != 0) return 0;
    return (long)L.grid * ((long)n_atoms * n_atoms + n_atoms + 1);
}

extern "C" int pm_resp_assemble(int n_atoms, const double *coords_dev, const double *grid_dev, const double *vref_dev,
                                const double *wt_dev, long n_conf, int n_grid, double *out_dev, double *scratch_dev, void *stream) {
    if (!coords_dev || !grid_dev || !vref_dev || !out_dev || !scratch_dev) return pm_fail_msg("pm_resp_assemble: null argument");
    if (n_atoms < 1 || n_conf < 1 || n_grid < 1) return pm_fail_msg("pm_resp_assemble: empty problem");
    int N = n_atoms, G = n_grid;
    long S = n_conf;
    const int chunks = (n_grid + RESP_CHUNK - 1) / RESP_CHUNK;
    long n_items = n_conf * chunks;
    RespLaunch L;
    const int rc = resp_launch(N, n_items, &L);
    if (rc == -1) return pm_fail_msg("pm_resp_assemble: too many atoms for this kernel generation (the limit is 175)");
    if (rc != 0) return pm_fail_msg(std::string("pm_resp_assemble: launch configuration failed: ") + cudaGetErrorString(cudaGetLastError()));
    const long stride = (long)N * N + N + 1;
    cudaStream_t st = (cudaStream_t)stream;
    int cpc = chunks;
    // the partials of CTAs that find all their items weightless must still be defined
    PMF_CUDA(cudaMemsetAsync(scratch_dev, 0, sizeof(double) * (size_t)L.grid * stride, st));
    void *args[] = {&N, &coords_dev, &grid_dev, &vref_dev, &wt_dev, &S, &G, &n_items, &cpc, &scratch_dev};
    PMF_CUDA(cudaLaunchKernel(L.fn, dim3(L.grid), dim3(L.threads), args, L.smem, st));
    pm_sum_partials_kernel<<<(unsigned)((stride + 255) / 256), 256, 0, st>>>(scratch_dev, L.grid, stride, out_dev);
    PMF_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_esp_potential(int n_atoms, const double *coords_dev, const double *grid_dev, const double *charges_dev,
                                long n_conf, int n_grid, double *vmm_dev, void *stream) {
    if (!coords_dev || !grid_dev || !charges_dev || !vmm_dev) return pm_fail_msg("pm_esp_potential: null argument");
    if (n_atoms < 1 || n_conf < 1 || n_grid < 1) return pm_fail_msg("pm_esp_potential: empty problem");
    const long total = n_conf * (long)n_grid;
    pm_esp_potential_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_atoms, coords_dev, grid_dev,
                                                                                              charges_dev, n_conf, n_grid, vmm_dev);
    PMF_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_lls_internal(int n_atoms, const double *coords_dev, long n_conf, int n_terms, const int *kind_dev,
                               const int *atoms_dev, double *q_dev, void *stream) {
    if (!coords_dev || !kind_dev || !atoms_dev || !q_dev) return pm_fail_msg("pm_lls_internal: null argument");
    if (n_atoms < 1 || n_conf < 1 || n_terms < 1) return pm_fail_msg("pm_lls_internal: empty problem");
    const long total = n_conf * (long)n_terms;
    pm_lls_internal_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_atoms, coords_dev, n_conf, n_terms,
                                                                                             kind_dev, atoms_dev, q_dev);
    PMF_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_lls_columns(const double *q_dev, long n_conf, int n_terms, int n_cols, const int *col_term_dev,
                              const int *col_kind_dev, const double *col_p0_dev, const int *col_per_dev, double *a_dev, void *stream) {
    if (!q_dev || !col_term_dev || !col_kind_dev || !col_p0_dev || !col_per_dev || !a_dev)
        return pm_fail_msg("pm_lls_columns: null argument");
    if (n_conf < 1 || n_terms < 1 || n_cols < 1) return pm_fail_msg("pm_lls_columns: empty problem");
    const long total = n_conf * (long)n_cols;
    pm_lls_columns_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q_dev, n_conf, n_terms, n_cols, col_term_dev,
                                                                                            col_kind_dev, col_p0_dev, col_per_dev, a_dev);
    PMF_CUDA(cudaGetLastError());
    return 0;
}
