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
// each CTA keeps its share of the accumulators in registers (4x4 micro-tiles of the upper triangle per thread) across all
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

#ifndef RESP_CTAS_PER_SM
#define RESP_CTAS_PER_SM 4
#endif
#define RESP_MAX_WARPS 8   // warps per CTA
#define RESP_MAX_TPT 4     // 4x4 micro-tiles per thread
#define RESP_GT 32         // grid points per shared-memory tile
#define RESP_CHUNK 512     // grid points per work item

// ---------------------------------------------------------------------------------------------------------------------
// K5
// ---------------------------------------------------------------------------------------------------------------------
// Which 4x4 micro-tile of the upper triangle of the padded (N+1) x (N+1) matrix every thread owns.  Built on the host
// (resp_make_plan): micro-tiles are grouped by 4 x 8 super-blocks and packed 32 to a warp, so that the 32 lanes of a warp read
// at most ~8 distinct 32-byte row pieces and ~8 distinct column pieces of R per grid point -- shared-memory traffic, which
// bounded the first version of this kernel, drops to ~4 wavefronts per 16 warp-wide DFMAs.
struct RespPlan {
    unsigned char bi[RESP_MAX_TPT * RESP_MAX_WARPS * 32];  // 0xFF = no tile
    unsigned char bj[RESP_MAX_TPT * RESP_MAX_WARPS * 32];
};

// R is kept in shared memory as Rt[k][half][block][2]: element j of grid point k lives at rt_off(j); the two 16-byte halves
// of a micro-tile row piece are then contiguous across neighbouring blocks (conflict-free LDS.128).
__device__ __forceinline__ int rt_off(int j, int NB) { return ((j >> 1) & 1) * (2 * NB) + 2 * (j >> 2) + (j & 1); }

// The reference potential V rides along as an extra "atom row" (row N of the padded R tile), so that the same rank-32 update
// that forms A = sum R R^T also yields B = sum R V (column N) and C = sum V.V (entry [N][N]).
template <int TPT>
__global__ void __launch_bounds__(RESP_MAX_WARPS * 32)
    pm_resp_kernel(const __grid_constant__ RespPlan plan, int N, const double *__restrict__ coords, const double *__restrict__ grid,
                   const double *__restrict__ vref, const double *__restrict__ wt, long S, int G, long n_items, int chunks_per_conf,
                   double *__restrict__ partial) {
    const int NP = (N + 1 + 3) & ~3, NB = NP >> 2;  // N atom rows + the V row, padded to a multiple of 4
    extern __shared__ __align__(16) double sm[];
    double *xs = sm;                       // [N][3]
    double *gs = xs + ((3 * N + 1) & ~1);  // [RESP_GT][4]: grid point (x, y, z) and sqrt(w) V
    double *Rt = gs + 4 * RESP_GT;         // [RESP_GT][NP]
    const int tid = threadIdx.x, T = blockDim.x;

    int bi[TPT], bj[TPT];
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
        const int e = t * T + tid;
        bi[t] = plan.bi[e] == 0xFF ? -1 : plan.bi[e];
        bj[t] = plan.bj[e];
    }
    double acc[TPT][16];
#pragma unroll
    for (int t = 0; t < TPT; ++t)
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[t][e] = 0.0;

    // generation: thread -> (row j, grid-point group kg); with T >= NP every thread keeps one row, else rows are strided
    const int kgroups = T >= NP ? T / NP : 1;
    const int j0 = T >= NP ? tid % NP : tid, kg = T >= NP ? tid / NP : 0, jstride = T >= NP ? NP : T;
    const bool gen_active = kg < kgroups;

    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long m = item / chunks_per_conf;
        const int g0 = (int)(item - m * chunks_per_conf) * RESP_CHUNK;
        const int g1 = min(G, g0 + RESP_CHUNK);
        const double sw = wt ? sqrt(wt[m]) : 1.0;
        __syncthreads();
        for (int i = tid; i < 3 * N; i += T) xs[i] = coords[(size_t)m * 3 * N + i];
        const double *gm = grid + (size_t)m * 3 * G;
        const double *vm = vref + (size_t)m * G;
        for (int t0 = g0; t0 < g1; t0 += RESP_GT) {
            __syncthreads();  // previous update done with Rt / gs
            const int nk = min(RESP_GT, g1 - t0);
            if (tid < RESP_GT) {
                const bool in = tid < nk;
                double4 g4;
                g4.x = in ? gm[3 * (t0 + tid)] : 0.0;
                g4.y = in ? gm[3 * (t0 + tid) + 1] : 0.0;
                g4.z = in ? gm[3 * (t0 + tid) + 2] : 0.0;
                g4.w = in ? sw * vm[t0 + tid] : 0.0;
                *reinterpret_cast<double2 *>(gs + 4 * tid) = make_double2(g4.x, g4.y);
                *reinterpret_cast<double2 *>(gs + 4 * tid + 2) = make_double2(g4.z, g4.w);
            }
            __syncthreads();
            if (gen_active) {
                for (int j = j0; j < NP; j += jstride) {
                    double *dst = Rt + rt_off(j, NB);
                    if (j < N) {
                        const double xj = xs[3 * j], yj = xs[3 * j + 1], zj = xs[3 * j + 2];
#pragma unroll 4
                        for (int k = kg; k < RESP_GT; k += kgroups) {
                            const double2 gxy = *reinterpret_cast<const double2 *>(gs + 4 * k);
                            const double dx = xj - gxy.x, dy = yj - gxy.y, dz = zj - gs[4 * k + 2];
                            const double r = sw * pm_rsqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
                            dst[k * NP] = k < nk ? r : 0.0;
                        }
                    } else {
                        for (int k = kg; k < RESP_GT; k += kgroups) dst[k * NP] = j == N ? gs[4 * k + 3] : 0.0;
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                if (bi[t] >= 0) {
                    const double *ra = Rt + 2 * bi[t], *rb = Rt + 2 * bj[t];
#pragma unroll 4
                    for (int k = 0; k < RESP_GT; ++k) {
                        const double2 a01 = *reinterpret_cast<const double2 *>(ra + k * NP),
                                      a23 = *reinterpret_cast<const double2 *>(ra + k * NP + 2 * NB);
                        const double2 b01 = *reinterpret_cast<const double2 *>(rb + k * NP),
                                      b23 = *reinterpret_cast<const double2 *>(rb + k * NP + 2 * NB);
                        const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int v = 0; v < 4; ++v) acc[t][4 * u + v] = fma(a[u], b[v], acc[t][4 * u + v]);
                    }
                }
            }
        }
    }
    // one partial per CTA: [N*N] A (both triangles), [N] B, [1] C
    double *out = partial + (size_t)blockIdx.x * ((size_t)N * N + N + 1);
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
        if (bi[t] >= 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int r = 4 * bi[t] + u, c = 4 * bj[t] + v;
                    if (bi[t] == bj[t] && r > c) continue;  // diagonal micro-tiles hold both triangles: take the upper one
                    const double val = acc[t][4 * u + v];
                    if (c < N) {  // r <= c < N: the A block
                        out[(size_t)r * N + c] = val;
                        out[(size_t)c * N + r] = val;
                    } else if (c == N) {
                        if (r < N) out[(size_t)N * N + r] = val;  // B
                        else if (r == N) out[(size_t)N * N + N] = val;  // C
                    }
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
static int resp_grid_size(long n_items) {
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    long g = (long)n_sm * RESP_CTAS_PER_SM;
    if (g > n_items) g = n_items;
    if (g < 1) g = 1;
    return (int)g;
}

extern "C" long pm_resp_scratch_doubles(int n_atoms, long n_conf, int n_grid) {
    const long chunks = (n_grid + RESP_CHUNK - 1) / RESP_CHUNK;
    return (long)resp_grid_size(n_conf * chunks) * ((long)n_atoms * n_atoms + n_atoms + 1);
}

// Packs the micro-tiles of the upper triangle into warp slots of 32 (see RespPlan).  Returns the number of slots, or -1.
static int resp_make_plan(int NB, RespPlan *plan, int *n_warps, int *tpt) {
    struct Sb {
        std::vector<std::pair<int, int>> tiles;
    };
    std::vector<Sb> sbs;
    for (int I = 0; 4 * I < NB; ++I)
        for (int J = 0; 8 * J < NB; ++J) {
            Sb sb;
            for (int r = 4 * I; r < 4 * I + 4 && r < NB; ++r)
                for (int c = 8 * J; c < 8 * J + 8 && c < NB; ++c)
                    if (r <= c) sb.tiles.push_back({r, c});
            if (!sb.tiles.empty()) sbs.push_back(sb);
        }
    std::stable_sort(sbs.begin(), sbs.end(), [](const Sb &x, const Sb &y) { return x.tiles.size() > y.tiles.size(); });
    std::vector<std::vector<std::pair<int, int>>> slots;  // first-fit decreasing
    for (const Sb &sb : sbs) {
        size_t s = 0;
        while (s < slots.size() && slots[s].size() + sb.tiles.size() > 32) ++s;
        if (s == slots.size()) slots.emplace_back();
        slots[s].insert(slots[s].end(), sb.tiles.begin(), sb.tiles.end());
    }
    const int ns = (int)slots.size();
    if (ns > RESP_MAX_TPT * RESP_MAX_WARPS) return -1;
    const int per = (ns + RESP_MAX_WARPS - 1) / RESP_MAX_WARPS;  // tiles per thread
    const int W = (ns + per - 1) / per;                          // warps per CTA
    memset(plan->bi, 0xFF, sizeof(plan->bi));
    memset(plan->bj, 0, sizeof(plan->bj));
    for (int s = 0; s < ns; ++s) {
        const int t = s / W, w = s % W;
        for (size_t l = 0; l < slots[s].size(); ++l) {
            const int e = t * W * 32 + w * 32 + (int)l;
            plan->bi[e] = (unsigned char)slots[s][l].first;
            plan->bj[e] = (unsigned char)slots[s][l].second;
        }
    }
    *n_warps = W;
    *tpt = per;
    return ns;
}

extern "C" int pm_resp_assemble(int n_atoms, const double *coords_dev, const double *grid_dev, const double *vref_dev,
                                const double *wt_dev, long n_conf, int n_grid, double *out_dev, double *scratch_dev, void *stream) {
    if (!coords_dev || !grid_dev || !vref_dev || !out_dev || !scratch_dev) return pm_fail_msg("pm_resp_assemble: null argument");
    if (n_atoms < 1 || n_conf < 1 || n_grid < 1) return pm_fail_msg("pm_resp_assemble: empty problem");
    const int N = n_atoms, NP = (N + 1 + 3) & ~3, NB = NP / 4;
    RespPlan plan;
    int W = 0, tpt = 0;
    if (NB > 255 || resp_make_plan(NB, &plan, &W, &tpt) < 0)
        return pm_fail_msg("pm_resp_assemble: too many atoms for this kernel generation (the limit is about 160)");
    const int chunks = (n_grid + RESP_CHUNK - 1) / RESP_CHUNK;
    const long n_items = n_conf * chunks;
    const int gridsz = resp_grid_size(n_items);
    const size_t smem = sizeof(double) * (((3 * N + 1) & ~1) + 4 * RESP_GT + (size_t)RESP_GT * NP);
    const long stride = (long)N * N + N + 1;
    cudaStream_t st = (cudaStream_t)stream;
#define RESP_LAUNCH(K)                                                                                                        \
    do {                                                                                                                      \
        if (smem > 48 * 1024) PMF_CUDA(cudaFuncSetAttribute(pm_resp_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        pm_resp_kernel<K><<<gridsz, W * 32, smem, st>>>(plan, N, coords_dev, grid_dev, vref_dev, wt_dev, n_conf, n_grid, n_items, \
                                                       chunks, scratch_dev);                                                  \
    } while (0)
    switch (tpt) {
        case 1: RESP_LAUNCH(1); break;
        case 2: RESP_LAUNCH(2); break;
        case 3: RESP_LAUNCH(3); break;
        default: RESP_LAUNCH(4); break;
    }
#undef RESP_LAUNCH
    PMF_CUDA(cudaGetLastError());
    pm_sum_partials_kernel<<<(unsigned)((stride + 255) / 256), 256, 0, st>>>(scratch_dev, gridsz, stride, out_dev);
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
