// pm_lls_kernels.cu -- K3 + K4 of the linear least-squares fit of the bonded force constants as ONE pass: the design matrix of
// ParaMol/MM_engines/least_square.py:469-649 is generated from the resident conformation tiles inside the kernel that contracts it,
// and never exists in HBM.
//
//   reference                                                                  here
//   ---------------------------------------------------------------------     ------------------------------------------------------
//   _calculate_a: for every column a Python loop over the conformations        pm_lls_colstats   column sums (-> means), min / max of the
//     (calculate_distance / calculate_angle / calculate_dihedral,               internal coordinates (-> the two-column expansion of
//     Utils/geometry.py:4-107), then column centring (:641-645)                 angle_eq, :583-586), A^T W e for a residual e
//   A * sqrt(w) / sqrt(var), A / scaling constants (:104-117)                  row / column scales applied while the tile is generated
//   np.linalg.lstsq(A, B) on the [S, M] matrix (:135)                          pm_lls_normal     G = [A b]^T W [A b] on mma.sync.m8n8k4.f64
//                                                                                                (SASS DMMA.8x8x4), (M+1)^2 doubles out
//                                                                              pm_lls_residual   e = b - A x for iterative refinement of the
//                                                                                                normal-equation solution (host, M x M)
//
// pm_lls_normal: the (M+1) x (M+1) Gram matrix (column M is the right-hand side, so G[:, M] = A^T W b and G[M][M] = b^T W b) is cut
// into 128 x 128 blocks (I, J >= I).  A CTA (16 warps) owns ONE block and a slice of the conformations: it keeps the block's
// accumulators in registers (a warp holds 32 x 32 = 4 x 4 DMMA tiles), generates the two 128-column panels of 32 conformations at a
// time into shared memory ([column][conformation], row stride 36 doubles: fragment loads and generation stores are both conflict
// free), and overlaps the generation of the next 32 conformations with the rank-32 update of the current ones (double buffer, one
// CTA barrier per 32 conformations).  21 blocks x 7 slices fill the 148 SMs for the 742 columns of a 60-atom ligand; every CTA
// writes one partial block and pm_lls_block_sum adds the slices in a fixed order (deterministic, no atomics).
//
// All FP64: the reference is numpy float64 and the fitted parameters are compared at 1e-6 relative.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/paramol_b200.h"
#include "pm_device.cuh"

int pm_fail_msg(const std::string &msg);  // pm_kernels.cu

#define PML_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) return pm_fail_msg(std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define LLS_PANEL 128      // columns per panel (block edge of the Gram matrix)
#define LLS_KC 32          // conformations per generated chunk (8 DMMA k-steps of 4)
#define LLS_RS 36          // row stride of a panel in shared memory, doubles (= 4 mod 8: conflict-free fragment loads)
#define LLS_MAX_PANELS 16  // up to 2047 columns
#define LLS_NORMAL_THREADS 512
#define LLS_STAT_CHUNK 32  // conformations staged per step of the column-statistics kernel

struct LlsTerm {  // one internal coordinate (= one optimizable force constant)
    int a0, a1, a2, a3;
    int kind;  // 0 bond, 1 angle, 2 torsion
    int per;   // periodicity (torsions)
    int col;   // first column
    int ncol;  // 1, or 2 when the equilibrium value / phase is fitted as well (two-column expansion)
};

struct LlsDev {  // device view of a plan
    const LlsTerm *terms;
    const double *p0;     // [M] centre r0 / theta0 (bonds, angles)
    const double *cd;     // [M] cos(delta) (torsions)
    const double *sd;     // [M] sin(delta)
    const double *mean;   // [M] subtracted from the raw column
    const double *scale;  // [M] multiplies the centred column
    const double2 *acos_tab;  // (cos, sin)(q pi / PM_ACOS_N): pm_acos_cs
    int n_atoms, cpt, n_terms, n_cols;
};

struct LlsPanels {
    int t0[LLS_MAX_PANELS], t1[LLS_MAX_PANELS];  // terms [t0, t1) have a column in the panel
};

struct pm_lls_s {
    int device, n_atoms, cpt, n_terms, n_cols, n_sm, n_panels;
    LlsTerm *terms_dev;
    double *cols_dev;  // [5][M]: p0, cos delta, sin delta, mean, scale
    double2 *acos_dev;
    std::vector<LlsTerm> terms;
    LlsPanels panels;
};

// ---------------------------------------------------------------------------------------------------------------------
// geometry: the raw column values of one term for one conformation
// ---------------------------------------------------------------------------------------------------------------------
// LD(atom) -> pm_d3.  q is the internal coordinate (bonds, angles; 0 for torsions), v0 / v1 the raw (uncentred) columns:
// 1/2 (q - p0)^2 (least_square.py:516-517, 566-567) or 1 + cos(n phi - delta) (:614-615).  Conventions of Utils/geometry.py: the angle
// is arccos of the clipped cosine of the unit vectors (here recovered from cosine AND sine, pm_acos_cs: the same angle, well
// conditioned near 0 and pi); the dihedral is the atan2 formula of calculate_dihedral -- cos phi and sin phi are its normalised arguments, cos n phi and
// sin n phi follow by n complex multiplications, so no inverse trigonometric function is needed at all.
template <class LD>
__device__ __forceinline__ void lls_term_values(const LlsTerm &t, LD ld, const LlsDev &P, double &q, double &v0, double &v1) {
    const int c0 = t.col, c1 = t.col + (t.ncol > 1 ? 1 : 0);
    if (t.kind == 0) {
        const pm_d3 d = pm_sub(ld(t.a1), ld(t.a0));
        const double r2 = pm_dot(d, d);
        q = r2 > 0.0 ? r2 * pm_rsqrt(r2) : 0.0;
        const double e0 = q - P.p0[c0], e1 = q - P.p0[c1];
        v0 = 0.5 * e0 * e0;
        v1 = 0.5 * e1 * e1;
    } else if (t.kind == 1) {
        const pm_d3 x1 = ld(t.a1);
        const pm_d3 u = pm_sub(ld(t.a0), x1), w = pm_sub(ld(t.a2), x1);
        const pm_d3 cr = pm_cross(u, w);
        // theta from cos = u.w / (|u||w|) and sin = |u x w| / (|u||w|) through the (cos, sin) table (pm_device.cuh): no libm atan2 / acos
        const double n2 = pm_dot(u, u) * pm_dot(w, w), c2 = pm_dot(cr, cr);
        const double inv = n2 > 0.0 ? pm_rsqrt(n2) : 0.0;
        const double sn = c2 > 0.0 ? c2 * pm_rsqrt(c2) * inv : 0.0;
        q = pm_acos_cs(fmin(1.0, fmax(-1.0, pm_dot(u, w) * inv)), sn, P.acos_tab);
        const double e0 = q - P.p0[c0], e1 = q - P.p0[c1];
        v0 = 0.5 * e0 * e0;
        v1 = 0.5 * e1 * e1;
    } else {
        const pm_d3 p1 = ld(t.a1), p2 = ld(t.a2);
        const pm_d3 b0 = pm_sub(ld(t.a0), p1), b2 = pm_sub(ld(t.a3), p2);
        pm_d3 b1 = pm_sub(p2, p1);
        const double l2 = pm_dot(b1, b1);
        const double il = l2 > 0.0 ? pm_rsqrt(l2) : 0.0;
        b1 = pm_d3{b1.x * il, b1.y * il, b1.z * il};
        const double d0 = pm_dot(b0, b1), d2 = pm_dot(b2, b1);
        const pm_d3 vv{fma(-d0, b1.x, b0.x), fma(-d0, b1.y, b0.y), fma(-d0, b1.z, b0.z)};
        const pm_d3 ww{fma(-d2, b1.x, b2.x), fma(-d2, b1.y, b2.y), fma(-d2, b1.z, b2.z)};
        const double xx = pm_dot(vv, ww), yy = pm_dot(pm_cross(b1, vv), ww);
        const double h2 = fma(xx, xx, yy * yy);
        double c = 1.0, s = 0.0;  // atan2(0, 0) = 0
        if (h2 > 0.0) {
            const double ih = pm_rsqrt(h2);
            c = xx * ih;
            s = yy * ih;
        }
        double cn = 1.0, sn = 0.0;
        for (int k = 0; k < t.per; ++k) {
            const double tcn = fma(cn, c, -(sn * s));
            sn = fma(sn, c, cn * s);
            cn = tcn;
        }
        q = 0.0;
        v0 = 1.0 + fma(cn, P.cd[c0], sn * P.sd[c0]);
        v1 = 1.0 + fma(cn, P.cd[c1], sn * P.sd[c1]);
    }
}

// conformation s of a tiled ensemble ([tile][3N][cpt], pm_pack_tiles layout; cpt = 1 is the plain [S][N][3] array)
struct LlsGlobalLoad {
    const double *base;
    int cpt;
    __device__ __forceinline__ pm_d3 operator()(int a) const {
        const double *p = base + (size_t)(3 * a) * cpt;
        return pm_d3{__ldg(p), __ldg(p + cpt), __ldg(p + 2 * cpt)};
    }
};
__device__ __forceinline__ LlsGlobalLoad lls_conf(const double *coords, long s, int n_atoms, int cpt) {
    const long tile = s / cpt;
    return LlsGlobalLoad{coords + (size_t)tile * 3 * n_atoms * cpt + (s - tile * cpt), cpt};
}

// ---------------------------------------------------------------------------------------------------------------------
// column statistics: y_j = sum_s w_s e_s (a_sj - mean_j),  min / max of the internal coordinate of every term
// ---------------------------------------------------------------------------------------------------------------------
// One thread per term (terms are ordered bonds, angles, torsions, so warps are almost uniform); the CTA stages 32 conformations
// at a time in shared memory (a contiguous, coalesced copy: whole tiles) and every thread walks them with its sums in registers.
// No cross-thread reduction: the CTA writes its partial row, pm_lls_stat_sum adds the rows in a fixed order.
__global__ void __launch_bounds__(1024, 1)
    pm_lls_colstat_kernel(LlsDev P, const double *__restrict__ coords, long S, const double *__restrict__ e, const double *__restrict__ w,
                          double *__restrict__ part_y, double *__restrict__ part_mm) {
    extern __shared__ __align__(16) double lsm[];
    const int chunk_d = 3 * P.n_atoms * LLS_STAT_CHUNK;
    double *xs = lsm;             // [chunk_d]
    double *we = lsm + chunk_d;   // [32] w_s e_s of the chunk (0 beyond S)
    const long n_chunks = (S + LLS_STAT_CHUNK - 1) / LLS_STAT_CHUNK;
    // registers for up to 2 terms per thread (n_terms <= 2 * blockDim is guaranteed by the host)
    double y[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, qmin[2] = {INFINITY, INFINITY}, qmax[2] = {-INFINITY, -INFINITY};
    LlsTerm tm[2];
    int nt = 0;
    for (int t = threadIdx.x; t < P.n_terms && nt < 2; t += blockDim.x) tm[nt++] = P.terms[t];
    double mean[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int k = 0; k < nt; ++k) {
        mean[k][0] = P.mean[tm[k].col];
        mean[k][1] = P.mean[tm[k].col + (tm[k].ncol > 1 ? 1 : 0)];
    }
    for (long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const long s0 = ch * LLS_STAT_CHUNK;
        __syncthreads();
        // whole tiles: the chunk is contiguous in the tiled layout (chunk = 32 conformations is a multiple of cpt); the tail of
        // the last tile is padding written by pm_pack_tiles (cpt > 1) or simply not read (cpt = 1)
        const long first = s0 / P.cpt * (long)(3 * P.n_atoms * P.cpt);
        long avail = ((S + P.cpt - 1) / P.cpt) * (long)(3 * P.n_atoms * P.cpt) - first;
        if (avail > chunk_d) avail = chunk_d;
        for (int i = threadIdx.x; i < (int)avail; i += blockDim.x) xs[i] = coords[first + i];
        if (threadIdx.x < LLS_STAT_CHUNK) {
            const long s = s0 + threadIdx.x;
            we[threadIdx.x] = s < S ? (w ? w[s] : 1.0) * (e ? e[s] : 1.0) : 0.0;
        }
        __syncthreads();
        const int n_here = (int)((S - s0) < LLS_STAT_CHUNK ? (S - s0) : LLS_STAT_CHUNK);
        for (int k = 0; k < nt; ++k) {
            for (int c = 0; c < n_here; ++c) {
                const int tile = c / P.cpt;
                const double *base = xs + (size_t)tile * 3 * P.n_atoms * P.cpt + (c - tile * P.cpt);
                const int cpt = P.cpt;
                auto ld = [&](int a) {
                    const double *p = base + (3 * a) * cpt;
                    return pm_d3{p[0], p[cpt], p[2 * cpt]};
                };
                double q, v0, v1;
                lls_term_values(tm[k], ld, P, q, v0, v1);
                const double f = we[c];
                y[k][0] = fma(f, v0 - mean[k][0], y[k][0]);
                y[k][1] = fma(f, v1 - mean[k][1], y[k][1]);
                qmin[k] = fmin(qmin[k], q);
                qmax[k] = fmax(qmax[k], q);
            }
        }
    }
    int k = 0;
    for (int t = threadIdx.x; t < P.n_terms && k < 2; t += blockDim.x, ++k) {
        double *py = part_y + (size_t)blockIdx.x * P.n_cols;
        py[tm[k].col] = y[k][0];
        if (tm[k].ncol > 1) py[tm[k].col + 1] = y[k][1];
        double *pm = part_mm + ((size_t)blockIdx.x * P.n_terms + t) * 2;
        pm[0] = qmin[k];
        pm[1] = qmax[k];
    }
}

__global__ void pm_lls_stat_sum(const double *__restrict__ part_y, const double *__restrict__ part_mm, int n_part, int n_cols, int n_terms,
                                double *__restrict__ y, double *__restrict__ mm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_cols) {
        double s = 0.0;
        for (int p = 0; p < n_part; ++p) s += part_y[(size_t)p * n_cols + i];
        y[i] = s;
    }
    if (i < n_terms && mm != nullptr) {
        double lo = INFINITY, hi = -INFINITY;
        for (int p = 0; p < n_part; ++p) {
            lo = fmin(lo, part_mm[((size_t)p * n_terms + i) * 2]);
            hi = fmax(hi, part_mm[((size_t)p * n_terms + i) * 2 + 1]);
        }
        mm[2 * i] = lo;
        mm[2 * i + 1] = hi;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// residual: e_s = b_s - sum_j (a_sj - mean_j) z_j     (one conformation per thread, terms from shared memory)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    pm_lls_residual_kernel(LlsDev P, const double *__restrict__ coords, long S, const double *__restrict__ z, const double *__restrict__ b,
                           double *__restrict__ e_out) {
    extern __shared__ __align__(16) double lsm[];
    double *zs = lsm;                                                            // [M] z_j
    double *ms = lsm + P.n_cols;                                                 // [M] mean_j
    LlsTerm *ts = reinterpret_cast<LlsTerm *>(lsm + 2 * (size_t)P.n_cols);       // [T]
    for (int i = threadIdx.x; i < P.n_cols; i += blockDim.x) {
        zs[i] = z[i];
        ms[i] = P.mean[i];
    }
    for (int i = threadIdx.x; i < P.n_terms; i += blockDim.x) ts[i] = P.terms[i];
    __syncthreads();
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const LlsGlobalLoad ld = lls_conf(coords, s, P.n_atoms, P.cpt);
    double acc = 0.0;
    for (int t = 0; t < P.n_terms; ++t) {
        const LlsTerm tm = ts[t];
        double q, v0, v1;
        lls_term_values(tm, ld, P, q, v0, v1);
        acc = fma(v0 - ms[tm.col], zs[tm.col], acc);
        if (tm.ncol > 1) acc = fma(v1 - ms[tm.col + 1], zs[tm.col + 1], acc);
    }
    e_out[s] = (b ? b[s] : 0.0) - acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// K3 + K4: G = [A b]^T W [A b]
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pml_dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Columns [128 panel, 128 panel + 128) of conformations [s0, s0 + 32) into pt[column][conformation]: lane = conformation, the warps
// share the panel's terms.  Entry = rowscale_s * scale_j * (raw_sj - mean_j); column M is rowscale_s * b_s; conformations beyond S
// and columns beyond M are zero (the latter are zeroed once, when the kernel starts).
__device__ __forceinline__ void lls_generate_panel(const LlsDev &P, const LlsTerm *__restrict__ terms, int n_t, int panel,
                                                   const double *__restrict__ coords, long s0, long S, const double *__restrict__ rs,
                                                   const double *__restrict__ b, double *__restrict__ pt, int warp, int n_warps, int lane) {
    const long s = s0 + lane;
    const bool live = s < S;
    const LlsGlobalLoad ld = lls_conf(coords, live ? s : S - 1, P.n_atoms, P.cpt);
    const double r = live ? (rs ? rs[s] : 1.0) : 0.0;
    const int lo = panel * LLS_PANEL;
    for (int t = warp; t < n_t; t += n_warps) {
        const LlsTerm tm = terms[t];
        double q, v0, v1;
        lls_term_values(tm, ld, P, q, v0, v1);
        const int j0 = tm.col - lo, j1 = j0 + 1;
        if (j0 >= 0) pt[j0 * LLS_RS + lane] = r * P.scale[tm.col] * (v0 - P.mean[tm.col]);
        if (tm.ncol > 1 && j1 < LLS_PANEL) pt[j1 * LLS_RS + lane] = r * P.scale[tm.col + 1] * (v1 - P.mean[tm.col + 1]);
    }
    const int jb = P.n_cols - lo;  // the right-hand side rides along as column M
    if (jb >= 0 && jb < LLS_PANEL && warp == n_warps - 1) pt[jb * LLS_RS + lane] = live && b ? r * b[s] : 0.0;
}

__global__ void __launch_bounds__(LLS_NORMAL_THREADS, 1)
    pm_lls_normal_kernel(LlsDev P, LlsPanels R, int n_panels, int n_split, const double *__restrict__ coords, long S,
                         const double *__restrict__ rs, const double *__restrict__ b, double *__restrict__ partial) {
    extern __shared__ __align__(16) double lsm[];
    constexpr int PANEL_D = LLS_PANEL * LLS_RS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = LLS_NORMAL_THREADS / 32;
    // block (I, J >= I) and conformation slice z of this CTA
    const int blk = blockIdx.x / n_split, z = blockIdx.x - blk * n_split;
    int I = 0, rem = blk;
    while (rem >= n_panels - I) {
        rem -= n_panels - I;
        ++I;
    }
    const int J = I + rem;
    const bool diag = I == J;
    // buffers: [2][panel I | panel J], then the metadata of the two panels: terms, column parameters (one column of margin on the
    // left for a term that straddles the panel boundary), the (cos, sin) table
    double *buf[2] = {lsm, lsm + 2 * PANEL_D};
    double *meta = lsm + 4 * PANEL_D;                       // [2 panels][5][LLS_PANEL + 2]
    double2 *s_acos = reinterpret_cast<double2 *>(meta + 2 * 5 * (LLS_PANEL + 2));
    LlsTerm *s_terms = reinterpret_cast<LlsTerm *>(s_acos + PM_ACOS_N + 1);  // [2][LLS_PANEL + 1]
    for (int i = tid; i < 4 * PANEL_D; i += LLS_NORMAL_THREADS) lsm[i] = 0.0;
    for (int i = tid; i <= PM_ACOS_N; i += LLS_NORMAL_THREADS) s_acos[i] = P.acos_tab[i];
    LlsDev PP[2];
    int n_t[2];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const int pnl = side ? J : I, lo = pnl * LLS_PANEL - 1;
        double *m = meta + side * 5 * (LLS_PANEL + 2);
        const double *src[5] = {P.p0, P.cd, P.sd, P.mean, P.scale};
        for (int i = tid; i < 5 * (LLS_PANEL + 2); i += LLS_NORMAL_THREADS) {
            const int a = i / (LLS_PANEL + 2), c = lo + (i - a * (LLS_PANEL + 2));
            m[i] = (c >= 0 && c < P.n_cols) ? src[a][c] : 0.0;
        }
        n_t[side] = R.t1[pnl] - R.t0[pnl];
        for (int i = tid; i < n_t[side]; i += LLS_NORMAL_THREADS) s_terms[side * (LLS_PANEL + 1) + i] = P.terms[R.t0[pnl] + i];
        PP[side] = P;  // column arrays re-based so that the absolute column index still works
        PP[side].p0 = m - lo;
        PP[side].cd = m + (LLS_PANEL + 2) - lo;
        PP[side].sd = m + 2 * (LLS_PANEL + 2) - lo;
        PP[side].mean = m + 3 * (LLS_PANEL + 2) - lo;
        PP[side].scale = m + 4 * (LLS_PANEL + 2) - lo;
        PP[side].acos_tab = s_acos;
    }
    __syncthreads();

    const int wr = warp >> 2, wc = warp & 3;  // 4 x 4 warps, 32 x 32 outputs each
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const long n_chunks = (S + LLS_KC - 1) / LLS_KC;
    long ch = z;
    if (ch < n_chunks) {
        lls_generate_panel(PP[0], s_terms, n_t[0], I, coords, ch * LLS_KC, S, rs, b, buf[0], warp, n_warps, lane);
        if (!diag) lls_generate_panel(PP[1], s_terms + LLS_PANEL + 1, n_t[1], J, coords, ch * LLS_KC, S, rs, b, buf[0] + PANEL_D, warp, n_warps, lane);
    }
    __syncthreads();
    int cur = 0;
    for (; ch < n_chunks; ch += n_split) {
        const long nxt = ch + n_split;
        if (nxt < n_chunks) {  // the next 32 conformations into the other buffer while this one is contracted
            lls_generate_panel(PP[0], s_terms, n_t[0], I, coords, nxt * LLS_KC, S, rs, b, buf[cur ^ 1], warp, n_warps, lane);
            if (!diag) lls_generate_panel(PP[1], s_terms + LLS_PANEL + 1, n_t[1], J, coords, nxt * LLS_KC, S, rs, b, buf[cur ^ 1] + PANEL_D, warp, n_warps, lane);
        }
        const double *pa = buf[cur] + (32 * wr + (lane >> 2)) * LLS_RS + (lane & 3);
        const double *pb = buf[cur] + (diag ? 0 : PANEL_D) + (32 * wc + (lane >> 2)) * LLS_RS + (lane & 3);
#pragma unroll
        for (int k4 = 0; k4 < LLS_KC / 4; ++k4) {
            double a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = pa[(8 * i) * LLS_RS + 4 * k4];
                bb[i] = pb[(8 * i) * LLS_RS + 4 * k4];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) pml_dmma(acc[i][j], a[i], bb[j]);
        }
        __syncthreads();
        cur ^= 1;
    }
    // one partial block per CTA: [z][blk][128][128]
    double *out = partial + ((size_t)z * (n_panels * (n_panels + 1) / 2) + blk) * (LLS_PANEL * LLS_PANEL);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = 32 * wr + 8 * i + (lane >> 2), c = 32 * wc + 8 * j + 2 * (lane & 3);
            *reinterpret_cast<double2 *>(out + (size_t)r * LLS_PANEL + c) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
}


// ---------------------------------------------------------------------------------------------------------------------
// K3 -> K4 through an L2-resident slab (the default route of pm_lls_normal)
// ---------------------------------------------------------------------------------------------------------------------
// In the fused kernel above a panel is regenerated by every block that needs it (6-7 times for 742 columns) and its generation --
// latency-bound scattered coordinate loads, ~100 instructions per value -- keeps the FP64 pipe 15 % busy (ncu,
// profiles/r2_lls_normal_summary.csv: 13.4 ms).  Here the columns of a GROUP of conformations are generated ONCE into a slab sized
// to stay in the 126 MB L2 (<= 48 MB; layout [chunk of 32 conformations][panel][128 columns][36], i.e. already in the padded
// shared-memory layout of the contraction), and the contraction streams the two panels of its block with one TMA bulk copy each
// per chunk (double-buffered, mbarrier completion) -- its inner loop is loads of fragments and DMMA only.  The slab is overwritten
// by the next group, so the [S][M] matrix still never exists; what reaches HBM is the eviction traffic of a buffer that fits L2.
__global__ void __launch_bounds__(256)
    pm_lls_slab_kernel(LlsDev P, int n_panels, const double *__restrict__ coords, long S, long conf0, int n_chunks,
                       const double *__restrict__ rs, const double *__restrict__ b, double *__restrict__ slab) {
    extern __shared__ __align__(16) double lsm[];
    double2 *s_acos = reinterpret_cast<double2 *>(lsm);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    for (int i = tid; i <= PM_ACOS_N; i += blockDim.x) s_acos[i] = P.acos_tab[i];
    __syncthreads();
    LlsDev Q = P;
    Q.acos_tab = s_acos;
    constexpr size_t PANEL_D = (size_t)LLS_PANEL * LLS_RS;
    for (int ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const long s = conf0 + (long)ch * LLS_KC + lane;
        const bool live = s < S;
        const LlsGlobalLoad ld = lls_conf(coords, live ? s : S - 1, P.n_atoms, P.cpt);
        const double r = live ? (rs ? rs[s] : 1.0) : 0.0;
        double *out = slab + (size_t)ch * n_panels * PANEL_D + lane;
        for (int t = warp; t < P.n_terms; t += n_warps) {
            const LlsTerm tm = P.terms[t];
            double q, v0, v1;
            lls_term_values(tm, ld, Q, q, v0, v1);
            out[(size_t)tm.col * LLS_RS] = r * P.scale[tm.col] * (v0 - P.mean[tm.col]);   // column j of panel j / 128 sits at j * 36
            if (tm.ncol > 1) out[(size_t)(tm.col + 1) * LLS_RS] = r * P.scale[tm.col + 1] * (v1 - P.mean[tm.col + 1]);
        }
        // the right-hand side as column M, zeros behind it
        for (int j = P.n_cols + warp; j < n_panels * LLS_PANEL; j += n_warps)
            out[(size_t)j * LLS_RS] = (j == P.n_cols && live && b) ? r * b[s] : 0.0;
    }
}

__global__ void __launch_bounds__(LLS_NORMAL_THREADS, 1)
    pm_lls_syrk_kernel(int n_panels, int n_split, const double *__restrict__ slab, int n_chunks, int accumulate, double *__restrict__ partial) {
    extern __shared__ __align__(128) double lsm[];
    constexpr int PANEL_D = LLS_PANEL * LLS_RS;
    constexpr uint32_t PANEL_BYTES = PANEL_D * sizeof(double);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int blk = blockIdx.x / n_split, z = blockIdx.x - blk * n_split;
    int I = 0, rem = blk;
    while (rem >= n_panels - I) {
        rem -= n_panels - I;
        ++I;
    }
    const int J = I + rem;
    const bool diag = I == J;
    uint64_t *bar = reinterpret_cast<uint64_t *>(lsm + 4 * PANEL_D);
    if (tid == 0) {
        pm_mbar_init(&bar[0], 1);
        pm_mbar_init(&bar[1], 1);
        pm_fence_mbar_init();
    }
    __syncthreads();
    const int wr = warp >> 2, wc = warp & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    auto issue = [&](int ch, int stage) {  // one thread: the two panels of chunk ch into buffer `stage`
        const double *src = slab + (size_t)ch * n_panels * PANEL_D;
        double *dst = lsm + (size_t)stage * 2 * PANEL_D;
        pm_mbar_expect_tx(&bar[stage], diag ? PANEL_BYTES : 2 * PANEL_BYTES);
        pm_bulk_g2s(dst, src + (size_t)I * PANEL_D, PANEL_BYTES, &bar[stage]);
        if (!diag) pm_bulk_g2s(dst + PANEL_D, src + (size_t)J * PANEL_D, PANEL_BYTES, &bar[stage]);
    };
    uint32_t parity = 0;  // bit s: phase of barrier s
    int ch = z;
    if (ch < n_chunks && tid == 0) issue(ch, 0);
    int cur = 0;
    for (; ch < n_chunks; ch += n_split) {
        if (ch + n_split < n_chunks && tid == 0) {
            pm_fence_proxy_async();  // the generic-proxy reads of that buffer (previous iteration) are ordered before the copy
            issue(ch + n_split, cur ^ 1);
        }
        pm_mbar_wait(&bar[cur], (parity >> cur) & 1u);
        parity ^= 1u << cur;
        const double *bcur = lsm + (size_t)cur * 2 * PANEL_D;
        const double *pa = bcur + (32 * wr + (lane >> 2)) * LLS_RS + (lane & 3);
        const double *pb = bcur + (diag ? 0 : PANEL_D) + (32 * wc + (lane >> 2)) * LLS_RS + (lane & 3);
#pragma unroll
        for (int k4 = 0; k4 < LLS_KC / 4; ++k4) {
            double a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = pa[(8 * i) * LLS_RS + 4 * k4];
                bb[i] = pb[(8 * i) * LLS_RS + 4 * k4];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) pml_dmma(acc[i][j], a[i], bb[j]);
        }
        __syncthreads();
        cur ^= 1;
    }
    double *out = partial + ((size_t)z * (n_panels * (n_panels + 1) / 2) + blk) * (LLS_PANEL * LLS_PANEL);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = 32 * wr + 8 * i + (lane >> 2), c = 32 * wc + 8 * j + 2 * (lane & 3);
            double2 *dst = reinterpret_cast<double2 *>(out + (size_t)r * LLS_PANEL + c);
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            if (accumulate) {
                const double2 old = *dst;
                v.x += old.x;
                v.y += old.y;
            }
            *dst = v;
        }
}

// G[(M+1)][(M+1)] (full, symmetric) from the partial blocks, slices added in order
__global__ void pm_lls_block_sum(const double *__restrict__ partial, int n_panels, int n_split, int n, double *__restrict__ g) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n * n) return;
    int r = (int)(idx / n), c = (int)(idx - (long)r * n);
    if (r > c) {
        const int t = r;
        r = c;
        c = t;
    }
    const int I = r / LLS_PANEL, J = c / LLS_PANEL;
    const int blk = I * n_panels - I * (I - 1) / 2 + (J - I);
    const int n_blk = n_panels * (n_panels + 1) / 2;
    const size_t off = (size_t)(r - I * LLS_PANEL) * LLS_PANEL + (c - J * LLS_PANEL);
    double s = 0.0;
    for (int z = 0; z < n_split; ++z) s += partial[((size_t)z * n_blk + blk) * (LLS_PANEL * LLS_PANEL) + off];
    g[idx] = s;
}

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
static LlsDev lls_dev(pm_lls p) {
    LlsDev d;
    d.terms = p->terms_dev;
    d.p0 = p->cols_dev;
    d.cd = p->cols_dev + p->n_cols;
    d.sd = p->cols_dev + 2 * (size_t)p->n_cols;
    d.mean = p->cols_dev + 3 * (size_t)p->n_cols;
    d.scale = p->cols_dev + 4 * (size_t)p->n_cols;
    d.acos_tab = p->acos_dev;
    d.n_atoms = p->n_atoms;
    d.cpt = p->cpt;
    d.n_terms = p->n_terms;
    d.n_cols = p->n_cols;
    return d;
}

static int lls_stat_grid(pm_lls p, long n_conf) {
    long g = 2L * p->n_sm;
    const long n_chunks = (n_conf + LLS_STAT_CHUNK - 1) / LLS_STAT_CHUNK;
    if (g > n_chunks) g = n_chunks;
    return (int)(g < 1 ? 1 : g);
}

static int lls_split(pm_lls p, long n_conf) {
    const int n_blk = p->n_panels * (p->n_panels + 1) / 2;
    long z = p->n_sm / n_blk;
    const long n_chunks = (n_conf + LLS_KC - 1) / LLS_KC;
    if (z > n_chunks) z = n_chunks;
    return (int)(z < 1 ? 1 : z);
}

extern "C" int pm_lls_create(int device, int n_atoms, int conf_per_tile, int n_terms, const int *term_kind, const int *term_atoms,
                             const int *term_per, const int *term_ncol, pm_lls *out) {
    if (!out || !term_kind || !term_atoms || !term_per || !term_ncol) return pm_fail_msg("pm_lls_create: null argument");
    if (n_atoms < 1 || n_terms < 1) return pm_fail_msg("pm_lls_create: empty problem");
    if (conf_per_tile != 1 && conf_per_tile != 4 && conf_per_tile != 8 && conf_per_tile != 16 && conf_per_tile != 32)
        return pm_fail_msg("pm_lls_create: conf_per_tile must be 1 (plain [S][N][3] coordinates), 4, 8, 16 or 32");
    if (n_terms > 2048) return pm_fail_msg("pm_lls_create: more than 2048 terms");
    if ((size_t)(3 * n_atoms * LLS_STAT_CHUNK + LLS_STAT_CHUNK) * sizeof(double) > 200 * 1024)
        return pm_fail_msg("pm_lls_create: too many atoms for this kernel generation (the limit is 266)");
    PML_CUDA(cudaSetDevice(device));
    pm_lls_s *p = new pm_lls_s();
    p->device = device;
    p->n_atoms = n_atoms;
    p->cpt = conf_per_tile;
    p->n_terms = n_terms;
    p->terms_dev = nullptr;
    p->cols_dev = nullptr;
    p->acos_dev = nullptr;
    p->n_sm = 148;
    cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, device);
    int col = 0;
    for (int t = 0; t < n_terms; ++t) {
        LlsTerm tm;
        tm.kind = term_kind[t];
        tm.per = term_per[t];
        tm.ncol = term_ncol[t];
        tm.col = col;
        const int need = tm.kind == 0 ? 2 : tm.kind == 1 ? 3 : 4;
        int a[4] = {0, 0, 0, 0};
        bool ok = tm.kind >= 0 && tm.kind <= 2 && (tm.ncol == 1 || tm.ncol == 2) && tm.per >= 0 && tm.per <= 64;
        for (int k = 0; k < need && ok; ++k) {
            a[k] = term_atoms[4 * t + k];
            ok = a[k] >= 0 && a[k] < n_atoms;
        }
        if (!ok) {
            delete p;
            return pm_fail_msg("pm_lls_create: bad term " + std::to_string(t) + " (kind 0..2, 1 or 2 columns, atoms inside the molecule)");
        }
        tm.a0 = a[0];
        tm.a1 = a[1];
        tm.a2 = a[2];
        tm.a3 = a[3];
        col += tm.ncol;
        p->terms.push_back(tm);
    }
    p->n_cols = col;
    p->n_panels = (col + 1 + LLS_PANEL - 1) / LLS_PANEL;
    if (p->n_panels > LLS_MAX_PANELS) {
        delete p;
        return pm_fail_msg("pm_lls_create: more than 2047 columns");
    }
    for (int q = 0; q < LLS_MAX_PANELS; ++q) p->panels.t0[q] = p->panels.t1[q] = 0;
    for (int q = 0; q < p->n_panels; ++q) {
        int t0 = n_terms, t1 = 0;
        for (int t = 0; t < n_terms; ++t) {
            const LlsTerm &tm = p->terms[t];
            if (tm.col + tm.ncol > q * LLS_PANEL && tm.col < (q + 1) * LLS_PANEL) {
                if (t < t0) t0 = t;
                t1 = t + 1;
            }
        }
        p->panels.t0[q] = t0 < t1 ? t0 : 0;
        p->panels.t1[q] = t0 < t1 ? t1 : 0;
    }
    const std::vector<double> tab = pm_acos_table_host();
    if (cudaMalloc(&p->acos_dev, sizeof(double) * tab.size()) != cudaSuccess ||
        cudaMemcpy(p->acos_dev, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc(&p->terms_dev, sizeof(LlsTerm) * n_terms) != cudaSuccess ||
        cudaMalloc(&p->cols_dev, sizeof(double) * 5 * (size_t)col) != cudaSuccess ||
        cudaMemcpy(p->terms_dev, p->terms.data(), sizeof(LlsTerm) * n_terms, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(p->cols_dev, 0, sizeof(double) * 5 * (size_t)col) != cudaSuccess) {
        const std::string why = cudaGetErrorString(cudaGetLastError());
        if (p->terms_dev) cudaFree(p->terms_dev);
        if (p->cols_dev) cudaFree(p->cols_dev);
        if (p->acos_dev) cudaFree(p->acos_dev);
        delete p;
        return pm_fail_msg("pm_lls_create: " + why);
    }
    *out = p;
    return 0;
}

extern "C" int pm_lls_n_cols(pm_lls p) { return p ? p->n_cols : -1; }

extern "C" int pm_lls_set_columns(pm_lls p, const double *p0, const double *mean, const double *scale) {
    if (!p || !p0) return pm_fail_msg("pm_lls_set_columns: null argument");
    PML_CUDA(cudaSetDevice(p->device));
    const int M = p->n_cols;
    std::vector<double> h(5 * (size_t)M);
    for (const LlsTerm &tm : p->terms)
        for (int k = 0; k < tm.ncol; ++k) {
            const int j = tm.col + k;
            h[j] = p0[j];
            h[M + j] = tm.kind == 2 ? cos(p0[j]) : 0.0;
            h[2 * (size_t)M + j] = tm.kind == 2 ? sin(p0[j]) : 0.0;
            h[3 * (size_t)M + j] = mean ? mean[j] : 0.0;
            h[4 * (size_t)M + j] = scale ? scale[j] : 1.0;
        }
    PML_CUDA(cudaMemcpy(p->cols_dev, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
    return 0;
}

// chunks of 32 conformations per slab group: the slab stays below LLS_SLAB_BYTES (L2-resident)
#define LLS_SLAB_BYTES (48L << 20)
static long lls_group_chunks(pm_lls p, long n_conf) {
    const long n_chunks = (n_conf + LLS_KC - 1) / LLS_KC;
    long g = LLS_SLAB_BYTES / ((long)p->n_panels * LLS_PANEL * LLS_RS * (long)sizeof(double));
    if (g < 8) g = 8;
    return g < n_chunks ? g : n_chunks;
}

static bool lls_fused_route() {
    const char *e = getenv("PM_LLS_FUSED");
    return e && e[0] == '1';
}

extern "C" long pm_lls_scratch_doubles(pm_lls p, long n_conf) {
    if (!p || n_conf < 1) return -1;
    const long stat = (long)lls_stat_grid(p, n_conf) * ((long)p->n_cols + 2L * p->n_terms);
    const long n_blk = (long)p->n_panels * (p->n_panels + 1) / 2;
    const long normal = (long)lls_split(p, n_conf) * n_blk * LLS_PANEL * LLS_PANEL +
                        lls_group_chunks(p, n_conf) * p->n_panels * LLS_PANEL * LLS_RS;
    return stat > normal ? stat : normal;
}

extern "C" int pm_lls_colstats(pm_lls p, const double *coords_dev, long n_conf, const double *e_dev, const double *w_dev, double *y_dev,
                               double *minmax_dev, double *scratch_dev, void *stream) {
    if (!p || !coords_dev || !y_dev || !scratch_dev) return pm_fail_msg("pm_lls_colstats: null argument");
    if (n_conf < 1) return pm_fail_msg("pm_lls_colstats: no conformations");
    PML_CUDA(cudaSetDevice(p->device));
    const int grid = lls_stat_grid(p, n_conf);
    int threads = (p->n_terms + 31) / 32 * 32;
    if (threads > 1024) threads = 1024;
    const size_t smem = sizeof(double) * (size_t)(3 * p->n_atoms * LLS_STAT_CHUNK + LLS_STAT_CHUNK);
    PML_CUDA(cudaFuncSetAttribute((const void *)pm_lls_colstat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double *part_y = scratch_dev, *part_mm = scratch_dev + (size_t)grid * p->n_cols;
    cudaStream_t st = (cudaStream_t)stream;
    pm_lls_colstat_kernel<<<grid, threads, smem, st>>>(lls_dev(p), coords_dev, n_conf, e_dev, w_dev, part_y, part_mm);
    PML_CUDA(cudaGetLastError());
    const int n = p->n_cols > p->n_terms ? p->n_cols : p->n_terms;
    pm_lls_stat_sum<<<(n + 255) / 256, 256, 0, st>>>(part_y, part_mm, grid, p->n_cols, p->n_terms, y_dev, minmax_dev);
    PML_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_lls_residual(pm_lls p, const double *coords_dev, long n_conf, const double *z_dev, const double *b_dev, double *e_dev,
                               void *stream) {
    if (!p || !coords_dev || !z_dev || !e_dev) return pm_fail_msg("pm_lls_residual: null argument");
    if (n_conf < 1) return pm_fail_msg("pm_lls_residual: no conformations");
    PML_CUDA(cudaSetDevice(p->device));
    const size_t smem = sizeof(double) * 2 * (size_t)p->n_cols + sizeof(LlsTerm) * (size_t)p->n_terms;
    PML_CUDA(cudaFuncSetAttribute((const void *)pm_lls_residual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pm_lls_residual_kernel<<<(unsigned)((n_conf + 255) / 256), 256, smem, (cudaStream_t)stream>>>(lls_dev(p), coords_dev, n_conf, z_dev, b_dev,
                                                                                              e_dev);
    PML_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_lls_normal(pm_lls p, const double *coords_dev, long n_conf, const double *rowscale_dev, const double *b_dev,
                             double *g_dev, double *scratch_dev, void *stream) {
    if (!p || !coords_dev || !g_dev || !scratch_dev) return pm_fail_msg("pm_lls_normal: null argument");
    if (n_conf < 1) return pm_fail_msg("pm_lls_normal: no conformations");
    PML_CUDA(cudaSetDevice(p->device));
    const int n_blk = p->n_panels * (p->n_panels + 1) / 2;
    const int n_split = lls_split(p, n_conf);
    cudaStream_t st = (cudaStream_t)stream;
    if (lls_fused_route()) {  // the single-kernel route (generation inside the contraction), kept for comparison
        const size_t smem = sizeof(double) * (4 * (size_t)LLS_PANEL * LLS_RS + 2 * 5 * (LLS_PANEL + 2)) + sizeof(double2) * (PM_ACOS_N + 1) +
                            sizeof(LlsTerm) * 2 * (LLS_PANEL + 1);
        PML_CUDA(cudaFuncSetAttribute((const void *)pm_lls_normal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pm_lls_normal_kernel<<<n_blk * n_split, LLS_NORMAL_THREADS, smem, st>>>(lls_dev(p), p->panels, p->n_panels, n_split, coords_dev, n_conf,
                                                                               rowscale_dev, b_dev, scratch_dev);
        PML_CUDA(cudaGetLastError());
    } else {
        const long n_chunks = (n_conf + LLS_KC - 1) / LLS_KC, per_group = lls_group_chunks(p, n_conf);
        double *slab = scratch_dev + (size_t)n_split * n_blk * LLS_PANEL * LLS_PANEL;
        const size_t smem_syrk = sizeof(double) * 4 * (size_t)LLS_PANEL * LLS_RS + 64;
        PML_CUDA(cudaFuncSetAttribute((const void *)pm_lls_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_syrk));
        const size_t smem_gen = sizeof(double2) * (PM_ACOS_N + 1);
        for (long c0 = 0, group = 0; c0 < n_chunks; c0 += per_group, ++group) {
            const int n_here = (int)(n_chunks - c0 < per_group ? n_chunks - c0 : per_group);
            const int gen_grid = n_here < 8 * p->n_sm ? n_here : 8 * p->n_sm;
            pm_lls_slab_kernel<<<gen_grid, 256, smem_gen, st>>>(lls_dev(p), p->n_panels, coords_dev, n_conf, c0 * LLS_KC, n_here, rowscale_dev,
                                                               b_dev, slab);
            PML_CUDA(cudaGetLastError());
            pm_lls_syrk_kernel<<<n_blk * n_split, LLS_NORMAL_THREADS, smem_syrk, st>>>(p->n_panels, n_split, slab, n_here, group > 0 ? 1 : 0,
                                                                                      scratch_dev);
            PML_CUDA(cudaGetLastError());
        }
    }
    const int n = p->n_cols + 1;
    const long total = (long)n * n;
    pm_lls_block_sum<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(scratch_dev, p->n_panels, n_split, n, g_dev);
    PML_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_lls_destroy(pm_lls p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    cudaFree(p->terms_dev);
    cudaFree(p->cols_dev);
    cudaFree(p->acos_dev);
    delete p;
    return 0;
}
