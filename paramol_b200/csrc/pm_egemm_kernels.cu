// pm_egemm_kernels.cu -- K2 for ENERGY-ONLY batches: geometry is evaluated once per conformation and reused by every parameter vector.
//
// What is replaced (file:line under /root/reference): the optimisers that evaluate MANY trial vectors on the same conformations --
// Monte Carlo (ParaMol/Optimizers/monte_carlo.py:89-101), simulated annealing (simulated_annealing.py:90-131), the finite-difference
// gradient (gradient_descent.py:104-122, scipy "2-point") and the WHAM energy passes (ParaMol/System/system.py) -- call
// get_energies_ensemble once per vector: S x (setPositions + getState) each.  Every MM energy term is LINEAR in coefficients that
// depend only on the parameters once the geometry is fixed:
//     bond   1/2 k (r - r0)^2          = (k/2) r^2 + (-k r0) r + k r0^2 / 2
//     angle  1/2 k (theta - theta0)^2  = likewise in theta
//     torsion k (1 + cos(n phi - d))   = (k cos d) cos(n phi) + (k sin d) sin(n phi) + k
//     pair   c12 r^-12 - c6 r^-6 + K qq / r
// so  E[p][s] = sum_t C[p][t] G[t][s]  with T = 2 n_bond + 2 n_angle + 2 n_torsion + 3 n_pair + 1 features (the constants share
// one feature): a dense FP64 contraction.  pm_egemm_kernel generates G for 64 conformations at a time from the resident tiles into
// shared memory (16 features per step, double-buffered; never stored in HBM), streams the coefficient block of 128 vectors through
// shared memory and contracts on mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4); the epilogue turns E into the five weighted sums of
// pm_reduce_objective per vector (rows, summed -- across GPUs too -- by pm_reduce_rows) and optionally writes E itself.
// Cost: 2 P S T flop on the FP64 tensor path instead of P full term passes; forces are not covered (the force residual is quadratic
// in the forces), so E+F batches keep K1.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/paramol_b200.h"
#include "pm_device.cuh"
#include "pm_internal.hpp"

#define PMG_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) return pm_fail_msg(std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define EG_PB 128      // parameter vectors per CTA block
#define EG_ST 64       // conformations per tile
// features per step: KC = 16 (4 DMMA k-steps) by default, 32 as a sweep option;
// row stride of the coefficient chunk in shared memory = KC + 4 (= 4 mod 8 doubles: conflict-free A fragments)
#define EG_GS 68       // row stride of the feature chunk (= 4 mod 16 doubles: the four feature rows a half-warp reads as B fragments
                       // fall into disjoint bank ranges; 72 measured 42 % excess wavefronts)
#define EG_THREADS 512

struct EgTerm {  // 32 bytes
    int a0, a1, a2, a3;
    int kind;  // 0 bond (r^2, r), 1 angle (theta^2, theta), 2 torsion (cos n phi, sin n phi), 3 pair (r^-12, r^-6, 1/r), 4 constant (1)
    int per;   // periodicity (torsions)
    int row;   // first feature row inside its chunk
    int pad;   // index of that chunk
};

struct EgFeat {  // how a coefficient is made from the derived table
    int kind;    // 0: k/2  1: -k x0  (bond / angle, derived (x0, k) at off)   2: k cos d  3: k sin d  (torsion, derived (k, kn, cos d, sin d) at off)
                 // 4: derived[off]  5: -derived[off]  6: zero (padding)  7: the constant (sum over all terms)
    int off;
};

struct pm_egemm_s {
    int kc;                         // features per step (16 or 32)
    int n_terms, n_feat, n_chunks;  // n_feat = kc * n_chunks
    EgTerm *terms_dev;
    int *chunk_ptr_dev;  // [n_chunks + 1] first term of every chunk
    EgFeat *feat_dev;    // [n_feat]
    int n_bonds, n_angles, n_torsions;
};

static void pm_egemm_free(void *p) {
    pm_egemm_s *e = (pm_egemm_s *)p;
    if (!e) return;
    cudaFree(e->terms_dev);
    cudaFree(e->chunk_ptr_dev);
    cudaFree(e->feat_dev);
    delete e;
}

// ---------------------------------------------------------------------------------------------------------------------
// coefficients C[Ppad][n_feat] from the derived tables [P][n_derived]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    pm_egemm_coef_kernel(const double *__restrict__ derived, int n_vec, int n_derived, const EgFeat *__restrict__ feat, int n_feat,
                         int doff_bond, int n_bonds, int doff_angle, int n_angles, int doff_tors, int n_tors, double *__restrict__ C) {
    const int p = blockIdx.x;
    double *c = C + (size_t)p * n_feat;
    if (p >= n_vec) {  // padding rows of the last block
        for (int f = threadIdx.x; f < n_feat; f += blockDim.x) c[f] = 0.0;
        return;
    }
    const double *d = derived + (size_t)p * n_derived;
    int f_const = -1;
    for (int f = threadIdx.x; f < n_feat; f += blockDim.x) {
        const EgFeat ft = feat[f];
        double v = 0.0;
        switch (ft.kind) {
            case 0: v = 0.5 * d[ft.off + 1]; break;
            case 1: v = -d[ft.off + 1] * d[ft.off]; break;
            case 2: v = d[ft.off] * d[ft.off + 2]; break;
            case 3: v = d[ft.off] * d[ft.off + 3]; break;
            case 4: v = d[ft.off]; break;
            case 5: v = -d[ft.off]; break;
            case 7: f_const = f; break;
            default: break;
        }
        if (ft.kind != 7) c[f] = v;
    }
    // the constant: sum of k x0^2 / 2 over bonds and angles and of k over torsions (fixed order: thread-strided, then a tree)
    double s = 0.0;
    for (int b = threadIdx.x; b < n_bonds; b += blockDim.x) s += 0.5 * d[doff_bond + 2 * b + 1] * d[doff_bond + 2 * b] * d[doff_bond + 2 * b];
    for (int a = threadIdx.x; a < n_angles; a += blockDim.x) s += 0.5 * d[doff_angle + 2 * a + 1] * d[doff_angle + 2 * a] * d[doff_angle + 2 * a];
    for (int t = threadIdx.x; t < n_tors; t += blockDim.x) s += d[doff_tors + 4 * t];
    __shared__ double sh[256];
    __shared__ int sh_f;
    if (threadIdx.x == 0) sh_f = -1;
    __syncthreads();
    sh[threadIdx.x] = s;
    if (f_const >= 0) sh_f = f_const;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0 && sh_f >= 0) c[sh_f] = sh[0];
}

// ---------------------------------------------------------------------------------------------------------------------
// E = C G with the features generated on the fly
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void eg_dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// features of one term for the conformation in column `col` of the staged coordinate tile -> rows tm.row .. of the chunk
__device__ __forceinline__ void eg_features(const EgTerm tm, const double *__restrict__ xs, int n_atoms, int cpt, int conf,
                                            const double2 *__restrict__ s_acos, double *__restrict__ g /* &Gs[0][conf] */) {
    const int tile = conf / cpt;
    const double *base = xs + (size_t)tile * 3 * n_atoms * cpt + (conf - tile * cpt);
    auto ld = [&](int a) {
        const double *p = base + (3 * a) * cpt;
        return pm_d3{p[0], p[cpt], p[2 * cpt]};
    };
    double *out = g + tm.row * EG_GS;
    if (tm.kind == 3) {
        const pm_d3 d = pm_sub(ld(tm.a0), ld(tm.a1));
        const double r2 = pm_dot(d, d);
        const double rinv = pm_rsqrt(r2);
        const double u = rinv * rinv, u3 = u * u * u;
        out[0] = u3 * u3;
        out[EG_GS] = u3;
        out[2 * EG_GS] = rinv;
    } else if (tm.kind == 0) {
        const pm_d3 d = pm_sub(ld(tm.a0), ld(tm.a1));
        const double r2 = pm_dot(d, d);
        out[0] = r2;
        out[EG_GS] = r2 * pm_rsqrt(r2);
    } else if (tm.kind == 1) {
        // theta exactly as the E+F kernel forms it (cosine from the dot product, sine from the Lagrange identity, table arccos)
        const pm_d3 xc = ld(tm.a1);
        const pm_d3 da = pm_sub(xc, ld(tm.a0)), db = pm_sub(xc, ld(tm.a2));
        const double ra2 = pm_dot(da, da), rb2 = pm_dot(db, db), dot = pm_dot(da, db);
        const double rr = pm_rsqrt(ra2) * pm_rsqrt(rb2);
        const double cosine = fmin(1.0, fmax(-1.0, dot * rr));
        const double rp2 = fmax(fma(-dot, dot, ra2 * rb2), 0.0);
        const bool tiny = rp2 <= 1.0e-12;
        const double sine = tiny ? sqrt(rp2) * rr : rp2 * pm_rsqrt(rp2) * rr;
        const double th = pm_acos_cs(cosine, sine, s_acos);
        out[0] = th * th;
        out[EG_GS] = th;
    } else if (tm.kind == 2) {
        // cos phi / sin phi from the two cross products, as in the axis code of the E+F kernel (OpenMM's dihedral sign convention)
        const pm_d3 xj = ld(tm.a1), xk = ld(tm.a2);
        const pm_d3 d0 = pm_sub(ld(tm.a0), xj), d1 = pm_sub(xk, xj), d2 = pm_sub(xk, ld(tm.a3));
        const pm_d3 c1 = pm_cross(d0, d1), c2 = pm_cross(d1, d2);
        const double nb2 = pm_dot(d1, d1);
        const double rb = nb2 * pm_rsqrt(nb2);
        const double inv12 = pm_rsqrt(pm_dot(c1, c1)) * pm_rsqrt(pm_dot(c2, c2));
        const double cphi = pm_dot(c1, c2) * inv12;
        const double sphi = rb * pm_dot(d0, c2) * inv12;
        double cn = tm.per > 0 ? cphi : 1.0, sn = tm.per > 0 ? sphi : 0.0;
        for (int mm = 1; mm < tm.per; ++mm) {
            const double c_new = fma(cn, cphi, -(sn * sphi));
            sn = fma(sn, cphi, cn * sphi);
            cn = c_new;
        }
        out[0] = cn;
        out[EG_GS] = sn;
    } else {
        out[0] = 1.0;
    }
}

template <int EG_KC>
__global__ void __launch_bounds__(EG_THREADS, 1)
    pm_egemm_kernel(int n_atoms, int cpt, const EgTerm *__restrict__ terms, const int *__restrict__ chunk_ptr, int n_chunks, int n_feat,
                    const double *__restrict__ C, const double2 *__restrict__ acos_tab, const double *__restrict__ coords, long S,
                    const double *__restrict__ eref, const double *__restrict__ wts, double d_shift, int n_vec, int n_z,
                    double *__restrict__ rows, double *__restrict__ emm) {
    constexpr int EG_CS = EG_KC + 4;
    extern __shared__ __align__(16) double sm[];
    const int tile_d = 3 * n_atoms * EG_ST;
    double *xs = sm;                                   // [tile_d] coordinates of the 64 conformations (tile layout)
    double *Cs = xs + tile_d;                          // [2][EG_PB][EG_CS]
    double *Gs = Cs + 2 * EG_PB * EG_CS;               // [2][EG_KC][EG_GS]
    double *racc = Gs + 2 * EG_KC * EG_GS;             // [4][EG_PB][4] running sums per (column group of warps, vector)
    double2 *s_acos = reinterpret_cast<double2 *>(racc + 4 * EG_PB * 4);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pb = blockIdx.x / n_z, z = blockIdx.x - pb * n_z;
    const int wr = warp >> 2, wc = warp & 3;           // 4 x 4 warps: 32 vectors x 16 conformations each
    for (int i = tid; i <= PM_ACOS_N; i += EG_THREADS) s_acos[i] = acos_tab[i];
    for (int i = tid; i < 4 * EG_PB * 4; i += EG_THREADS) racc[i] = 0.0;
    for (int i = tid; i < 2 * EG_KC * EG_GS; i += EG_THREADS) Gs[i] = 0.0;
    const double *Cb = C + (size_t)pb * EG_PB * n_feat;
    const long n_stiles = (S + EG_ST - 1) / EG_ST;
    const long n_coord = ((S + cpt - 1) / cpt) * (long)(3 * n_atoms * cpt);  // doubles in the tiled ensemble

    for (long st = z; st < n_stiles; st += n_z) {
        const long s0 = st * EG_ST;
        __syncthreads();  // everybody done with the previous tile's coordinates and buffers
        {
            const long first = s0 / cpt * (long)(3 * n_atoms * cpt);
            long avail = n_coord - first;
            if (avail > tile_d) avail = tile_d;
            for (int i = tid; i < tile_d; i += EG_THREADS) xs[i] = i < avail ? coords[first + i] : 1.0 + 0.37 * (i % 97);  // padding: finite, distinct
        }
        double acc[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        __syncthreads();

        // chunk c: coefficients Cb[:, 16 c .. 16 c + 16) and the features of its terms
        constexpr int CV = EG_KC / 4;  // doubles of the coefficient chunk per thread (128 rows x KC doubles over 512 threads)
        auto load_c = [&](int c, double (&v)[CV]) {
            const int row = tid >> 2, k0 = (tid & 3) * CV;
            const double2 *src = reinterpret_cast<const double2 *>(Cb + (size_t)row * n_feat + c * EG_KC + k0);
#pragma unroll
            for (int q = 0; q < CV / 2; ++q) {
                const double2 a = src[q];
                v[2 * q] = a.x;
                v[2 * q + 1] = a.y;
            }
        };
        auto store_c = [&](int buf, const double (&v)[CV]) {
            const int row = tid >> 2, k0 = (tid & 3) * CV;
            double2 *dst = reinterpret_cast<double2 *>(Cs + (size_t)buf * EG_PB * EG_CS + row * EG_CS + k0);
#pragma unroll
            for (int q = 0; q < CV / 2; ++q) dst[q] = make_double2(v[2 * q], v[2 * q + 1]);
        };
        auto generate = [&](int c, int buf) {
            const int t0 = chunk_ptr[c], t1 = chunk_ptr[c + 1];
            double *g = Gs + (size_t)buf * EG_KC * EG_GS;
            for (int item = warp; item < 2 * (t1 - t0); item += EG_THREADS / 32) {
                const int t = t0 + (item >> 1), half = item & 1;
                const EgTerm tm = terms[t];
                eg_features(tm, xs, n_atoms, cpt, half * 32 + lane, s_acos, g + half * 32 + lane);
            }
        };
        {
            double v[CV];
            load_c(0, v);
            store_c(0, v);
            generate(0, 0);
        }
        __syncthreads();
        for (int c = 0; c < n_chunks; ++c) {
            const int cur = c & 1;
            double v[CV];
            const bool more = c + 1 < n_chunks;
            if (more) {
                load_c(c + 1, v);       // in flight while this chunk is contracted
                generate(c + 1, cur ^ 1);
            }
            const double *pa = Cs + (size_t)cur * EG_PB * EG_CS + (32 * wr + (lane >> 2)) * EG_CS + (lane & 3);
            const double *pg = Gs + (size_t)cur * EG_KC * EG_GS + (lane & 3) * EG_GS + 16 * wc + (lane >> 2);
#pragma unroll
            for (int k4 = 0; k4 < EG_KC / 4; ++k4) {
                double a[4], b[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = pa[(8 * i) * EG_CS + 4 * k4];
#pragma unroll
                for (int j = 0; j < 2; ++j) b[j] = pg[(4 * k4) * EG_GS + 8 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) eg_dmma(acc[i][j], a[i], b[j]);
            }
            if (more) store_c(cur ^ 1, v);
            __syncthreads();
        }
        // epilogue: E[p][s] for p = 128 pb + 32 wr + 8 i + lane / 4, s = s0 + 16 wc + 8 j + 2 (lane % 4) + {0, 1}
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int pl = 32 * wr + 8 * i + (lane >> 2);
            const long p = (long)pb * EG_PB + pl;
            double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const long s = s0 + 16 * wc + 8 * j + 2 * (lane & 3) + e;
                    if (s < S && p < n_vec) {
                        const double en = acc[i][j][e];
                        if (emm != nullptr) emm[(size_t)p * S + s] = en;
                        const double d = (eref != nullptr ? eref[s] : 0.0) - en - d_shift;
                        const double ws = wts != nullptr ? wts[s] : 1.0;
                        q0 += d;
                        q1 = fma(ws, d, q1);
                        q2 = fma(ws * d, d, q2);
                        q3 += ws;
                    }
                }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                q0 += __shfl_xor_sync(0xffffffffu, q0, o);
                q1 += __shfl_xor_sync(0xffffffffu, q1, o);
                q2 += __shfl_xor_sync(0xffffffffu, q2, o);
                q3 += __shfl_xor_sync(0xffffffffu, q3, o);
            }
            if ((lane & 3) == 0) {
                double *r = racc + ((size_t)wc * EG_PB + pl) * 4;
                r[0] += q0;
                r[1] += q1;
                r[2] += q2;
                r[3] += q3;
            }
        }
    }
    __syncthreads();
    // one row of five sums per (vector, conformation slice): rows[p][z][5]
    for (int i = tid; i < EG_PB * 5; i += EG_THREADS) {
        const int pl = i / 5, q = i - pl * 5;
        const long p = (long)pb * EG_PB + pl;
        if (p < n_vec) {
            double v = 0.0;
            if (q < 4)
                for (int g = 0; g < 4; ++g) v += racc[((size_t)g * EG_PB + pl) * 4 + q];
            rows[((size_t)p * n_z + z) * 5 + q] = v;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// the same contraction with generation in bursts and the coefficients streamed by TMA (default)
// ---------------------------------------------------------------------------------------------------------------------
// In pm_egemm_kernel every warp generates, stages and contracts chunk by chunk with one CTA barrier per 16 features: the non-DMMA
// work of a chunk (a few latency-bound items on a few warps) and its DMMAs never overlap, and ncu shows 14 % of the executed
// instructions (the DMMAs) accounting for half of the time.  A producer / consumer split (4 or 8 generating warps feeding 16
// contracting warps through mbarriers) was built and measured SLOWER (18.0 / 14.2 ms against 15.3): the producers' dependent FP64
// chains queue behind the consumers' DMMAs on the shared pipe (a slot every ~64 cycles), so generation becomes the critical path;
// with generation switched off the consumers run at 9.4 ms.  What works is the opposite: generate in BURSTS.  All 16 warps generate
// the features of a BURST of chunks (up to 8: what fits beside the staged coordinates) at once (tens of independent items in flight,
// four interleaved dependency chains per warp, nobody on the pipe: the latency is paid once per burst), one barrier, then they contract the burst chunk by chunk with NO CTA barrier: the coefficient chunks arrive by TMA bulk
// copies (the coefficient kernel writes them chunk-major and padded, 20 KB contiguous each) into a ring of stages handed over
// with mbarriers (full: the copy's bytes; empty: 16 warp arrivals).  The tile's coordinates are staged in shared memory TRANSPOSED
// ([3 N][64], conformation fastest): in the tile layout the eight tiles a warp reads sit 3 N cpt doubles apart, i.e. in the same banks
// (8-way conflicts: 13.6 -> 12.0 ms); molecules too large for that read them from global memory.
// Measured on config 3 (ms): chunk-by-chunk 15.3 | producer / consumer 18.0 (4 producers), 14.2 (8) | bursts 13.5 | bursts overlapped
// with the contraction by two warp groups 15.5 | bursts + staged transposed coordinates 12.0 | the contraction alone 9.5-9.9.
#define EGP_KC 16
#define EGP_CS 20
#define EGB_STAGES 3      // coefficient-chunk ring
#define EGB_THREADS 512

// features of one term for conformation `s` read from the tiled ensemble in global memory -> rows of the burst buffer
__device__ __forceinline__ void eg_features_global(const EgTerm tm, const double *__restrict__ base, int cpt,
                                                   const double2 *__restrict__ s_acos, double *__restrict__ out) {
    auto ld = [&](int a) {
        const double *p = base + (size_t)(3 * a) * cpt;
        return pm_d3{p[0], p[cpt], p[2 * cpt]};
    };
    if (tm.kind == 3) {
        const pm_d3 d = pm_sub(ld(tm.a0), ld(tm.a1));
        const double r2 = pm_dot(d, d);
        const double rinv = pm_rsqrt(r2);
        const double u = rinv * rinv, u3 = u * u * u;
        out[0] = u3 * u3;
        out[EG_GS] = u3;
        out[2 * EG_GS] = rinv;
    } else if (tm.kind == 0) {
        const pm_d3 d = pm_sub(ld(tm.a0), ld(tm.a1));
        const double r2 = pm_dot(d, d);
        out[0] = r2;
        out[EG_GS] = r2 * pm_rsqrt(r2);
    } else if (tm.kind == 1) {
        const pm_d3 xc = ld(tm.a1);
        const pm_d3 da = pm_sub(xc, ld(tm.a0)), db = pm_sub(xc, ld(tm.a2));
        const double ra2 = pm_dot(da, da), rb2 = pm_dot(db, db), dot = pm_dot(da, db);
        const double rr = pm_rsqrt(ra2) * pm_rsqrt(rb2);
        const double cosine = fmin(1.0, fmax(-1.0, dot * rr));
        const double rp2 = fmax(fma(-dot, dot, ra2 * rb2), 0.0);
        const bool tiny = rp2 <= 1.0e-12;
        const double sine = tiny ? sqrt(rp2) * rr : rp2 * pm_rsqrt(rp2) * rr;
        const double th = pm_acos_cs(cosine, sine, s_acos);
        out[0] = th * th;
        out[EG_GS] = th;
    } else if (tm.kind == 2) {
        const pm_d3 xj = ld(tm.a1), xk = ld(tm.a2);
        const pm_d3 d0 = pm_sub(ld(tm.a0), xj), d1 = pm_sub(xk, xj), d2 = pm_sub(xk, ld(tm.a3));
        const pm_d3 c1 = pm_cross(d0, d1), c2 = pm_cross(d1, d2);
        const double nb2 = pm_dot(d1, d1);
        const double rb = nb2 * pm_rsqrt(nb2);
        const double inv12 = pm_rsqrt(pm_dot(c1, c1)) * pm_rsqrt(pm_dot(c2, c2));
        const double cphi = pm_dot(c1, c2) * inv12;
        const double sphi = rb * pm_dot(d0, c2) * inv12;
        double cn = tm.per > 0 ? cphi : 1.0, sn = tm.per > 0 ? sphi : 0.0;
        for (int mm = 1; mm < tm.per; ++mm) {
            const double c_new = fma(cn, cphi, -(sn * sphi));
            sn = fma(sn, cphi, cn * sphi);
            cn = c_new;
        }
        out[0] = cn;
        out[EG_GS] = sn;
    } else {
        out[0] = 1.0;
    }
}

__global__ void __launch_bounds__(EGB_THREADS, 1)
    pm_egemm_burst_kernel(int n_atoms, int cpt, const EgTerm *__restrict__ terms, const int *__restrict__ chunk_ptr, int n_chunks,
                          const double *__restrict__ Cc, const double2 *__restrict__ acos_tab, const double *__restrict__ coords, long S,
                          const double *__restrict__ eref, const double *__restrict__ wts, double d_shift, int n_vec, int n_z,
                          int n_sc, int stage_x, int skip_gen, double *__restrict__ rows, double *__restrict__ emm) {
    extern __shared__ __align__(128) double sm[];
    double *Cs = sm;                                               // [EGB_STAGES][EG_PB][EGP_CS]
    double *Gs = Cs + EGB_STAGES * EG_PB * EGP_CS;                 // [n_sc * EGP_KC][EG_GS]: the features of a burst of n_sc chunks
    double *racc = Gs + (size_t)n_sc * EGP_KC * EG_GS;             // [4][EG_PB][4]
    double2 *s_acos = reinterpret_cast<double2 *>(racc + 4 * EG_PB * 4);
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_acos + PM_ACOS_N + 1);  // full[stage], empty[stage]
    double *xs = reinterpret_cast<double *>(bars + 16);            // [3 N 64] the tile's coordinates (stage_x) -- else they are read from global memory
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pb = blockIdx.x / n_z, z = blockIdx.x - pb * n_z;
    const int wr = warp >> 2, wc = warp & 3;  // 4 x 4 warps: 32 vectors x 16 conformations each
    for (int i = tid; i <= PM_ACOS_N; i += EGB_THREADS) s_acos[i] = acos_tab[i];
    for (int i = tid; i < 4 * EG_PB * 4; i += EGB_THREADS) racc[i] = 0.0;
    for (int i = tid; i < n_sc * EGP_KC * EG_GS; i += EGB_THREADS) Gs[i] = 0.0;  // padding feature rows stay zero (finite)
    if (tid == 0) {
        for (int s = 0; s < EGB_STAGES; ++s) {
            pm_mbar_init(&bars[s], 1);                       // full: the issuing lane (+ the copy's bytes)
            pm_mbar_init(&bars[8 + s], EGB_THREADS / 32);    // empty: one arrival per warp
        }
        pm_fence_mbar_init();
    }
    __syncthreads();
    const long n_stiles = (S + EG_ST - 1) / EG_ST;
    const double *Cb = Cc + (size_t)pb * n_chunks * (EG_PB * EGP_CS);
    long g_issue = 0, g_use = 0;  // chunks whose copy has been issued / that have been contracted (stage = g % EGB_STAGES)
    auto issue = [&](int c) {     // thread 0: coefficient chunk c of this block into the next stage of the ring
        const int s = (int)(g_issue % EGB_STAGES);
        const uint32_t use = (uint32_t)(g_issue / EGB_STAGES);
        if (use > 0) pm_mbar_wait(&bars[8 + s], (use - 1) & 1u);   // all warps are done with what the stage held
        pm_fence_proxy_async();
        pm_mbar_expect_tx(&bars[s], (uint32_t)(EG_PB * EGP_CS * sizeof(double)));
        pm_bulk_g2s(Cs + (size_t)s * EG_PB * EGP_CS, Cb + (size_t)c * (EG_PB * EGP_CS), (uint32_t)(EG_PB * EGP_CS * sizeof(double)), &bars[s]);
        ++g_issue;
    };

    auto conf_base = [&](long sc) {
        if (sc >= S) sc = S - 1;   // columns beyond S are generated from the last conformation and never used
        const long tile = sc / cpt;
        return coords + (size_t)tile * 3 * n_atoms * cpt + (sc - tile * cpt);
    };
    for (long st = z; st < n_stiles; st += n_z) {
        const long s0 = st * EG_ST;
        double acc[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        const double *xb0 = conf_base(s0 + lane), *xb1 = conf_base(s0 + 32 + lane);
        int xstride = cpt;   // doubles between the rows (x, y, z of the atoms) of a conformation
        if (tid == 0)
            for (int c = 0; c < EGB_STAGES - 1 && c < n_chunks; ++c) issue(c);   // the ring runs EGB_STAGES - 1 chunks ahead
        if (stage_x) {
            // the 64 conformations of the tile are contiguous in the tiled layout: stage them once (the generation reads every atom
            // ~N times; from global memory that is L2 traffic, because the shared-memory footprint leaves almost no L1)
            __syncthreads();  // nobody generates from the previous tile's coordinates any more
            // staged TRANSPOSED, [3 N][64] with the conformation fastest: the lanes of a warp (consecutive conformations) then read
            // consecutive doubles -- in the tile layout [3 N][cpt] the 8 tiles of a warp sit 3 N cpt doubles apart, i.e. in the same banks
            const int tile_d = 3 * n_atoms * EG_ST, per_tile = 3 * n_atoms * cpt;
            const long first = s0 / cpt * (long)per_tile;
            long avail = ((S + cpt - 1) / cpt) * (long)per_tile - first;
            if (avail > tile_d) avail = tile_d;
            for (int i = tid; i < tile_d; i += EGB_THREADS) {
                const int t = i / per_tile, rem = i - t * per_tile, r = rem / cpt, c = rem - r * cpt;
                xs[r * EG_ST + t * cpt + c] = i < avail ? coords[first + i] : 1.0 + 0.37 * (i % 97);
            }
            xb0 = xs + lane;
            xb1 = xs + 32 + lane;
            xstride = EG_ST;
        }
        for (int c0 = 0; c0 < n_chunks; c0 += n_sc) {
            const int c1 = c0 + n_sc < n_chunks ? c0 + n_sc : n_chunks;
            __syncthreads();  // every warp has finished contracting the previous burst: its features may be overwritten
            {
                // items = (term, half of the 64 conformations); a warp takes FOUR items per pass and, when all four are pairs (87 % of
                // the features), evaluates them as four interleaved dependency chains
                const int t0 = chunk_ptr[c0], n_items = skip_gen ? 0 : 2 * (chunk_ptr[c1] - t0);
                for (int i0 = 4 * warp; i0 < n_items; i0 += 4 * (EGB_THREADS / 32)) {
                    EgTerm tm[4];
                    bool all_pairs = i0 + 3 < n_items;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        tm[q] = terms[t0 + ((i0 + q < n_items ? i0 + q : i0) >> 1)];
                        all_pairs = all_pairs && tm[q].kind == 3;
                    }
                    if (all_pairs) {
                        double r2[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double *xb = ((i0 + q) & 1) ? xb1 : xb0;
                            const double *pa = xb + (size_t)(3 * tm[q].a0) * xstride, *pc = xb + (size_t)(3 * tm[q].a1) * xstride;
                            const double dx = pa[0] - pc[0], dy = pa[xstride] - pc[xstride], dz = pa[2 * xstride] - pc[2 * xstride];
                            r2[q] = fma(dz, dz, fma(dy, dy, dx * dx));
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double rinv = pm_rsqrt(r2[q]);
                            const double u = rinv * rinv, u3 = u * u * u;
                            double *out = Gs + ((tm[q].pad - c0) * EGP_KC + tm[q].row) * EG_GS + 32 * ((i0 + q) & 1) + lane;
                            out[0] = u3 * u3;
                            out[EG_GS] = u3;
                            out[2 * EG_GS] = rinv;
                        }
                    } else {
                        for (int q = 0; q < 4 && i0 + q < n_items; ++q)
                            eg_features_global(tm[q], ((i0 + q) & 1) ? xb1 : xb0, xstride, s_acos,
                                               Gs + ((tm[q].pad - c0) * EGP_KC + tm[q].row) * EG_GS + 32 * ((i0 + q) & 1) + lane);
                    }
                }
            }
            __syncthreads();  // the burst's features are complete
            for (int c = c0; c < c1; ++c, ++g_use) {
                if (tid == 0 && c + EGB_STAGES - 1 < n_chunks) issue(c + EGB_STAGES - 1);
                const int s = (int)(g_use % EGB_STAGES);
                pm_mbar_wait(&bars[s], (uint32_t)(g_use / EGB_STAGES) & 1u);
                const double *pa = Cs + (size_t)s * EG_PB * EGP_CS + (32 * wr + (lane >> 2)) * EGP_CS + (lane & 3);
                const double *pg = Gs + ((c - c0) * EGP_KC + (lane & 3)) * EG_GS + 16 * wc + (lane >> 2);
#pragma unroll
                for (int k4 = 0; k4 < EGP_KC / 4; ++k4) {
                    double a[4], bq[2];
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = pa[(8 * i) * EGP_CS + 4 * k4];
#pragma unroll
                    for (int j = 0; j < 2; ++j) bq[j] = pg[(4 * k4) * EG_GS + 8 * j];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) eg_dmma(acc[i][j], a[i], bq[j]);
                }
                __syncwarp();
                if (lane == 0) pm_mbar_arrive(&bars[8 + s]);
            }
        }
        // epilogue: E[p][s] for p = 128 pb + 32 wr + 8 i + lane / 4, s = s0 + 16 wc + 8 j + 2 (lane % 4) + {0, 1}
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int pl = 32 * wr + 8 * i + (lane >> 2);
            const long p = (long)pb * EG_PB + pl;
            double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const long sc = s0 + 16 * wc + 8 * j + 2 * (lane & 3) + e;
                    if (sc < S && p < n_vec) {
                        const double en = acc[i][j][e];
                        if (emm != nullptr) emm[(size_t)p * S + sc] = en;
                        const double d = (eref != nullptr ? eref[sc] : 0.0) - en - d_shift;
                        const double ws = wts != nullptr ? wts[sc] : 1.0;
                        q0 += d;
                        q1 = fma(ws, d, q1);
                        q2 = fma(ws * d, d, q2);
                        q3 += ws;
                    }
                }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                q0 += __shfl_xor_sync(0xffffffffu, q0, o);
                q1 += __shfl_xor_sync(0xffffffffu, q1, o);
                q2 += __shfl_xor_sync(0xffffffffu, q2, o);
                q3 += __shfl_xor_sync(0xffffffffu, q3, o);
            }
            if ((lane & 3) == 0) {
                double *r = racc + ((size_t)wc * EG_PB + pl) * 4;
                r[0] += q0;
                r[1] += q1;
                r[2] += q2;
                r[3] += q3;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < EG_PB * 5; i += EGB_THREADS) {
        const int pl = i / 5, q = i - pl * 5;
        const long p = (long)pb * EG_PB + pl;
        if (p < n_vec) {
            double v = 0.0;
            if (q < 4)
                for (int gq = 0; gq < 4; ++gq) v += racc[((size_t)gq * EG_PB + pl) * 4 + q];
            rows[((size_t)p * n_z + z) * 5 + q] = v;
        }
    }
}

// the coefficient matrix chunk-major and padded for the bulk copies of the pipeline: Cc[block][chunk][128][20]
__global__ void pm_egemm_chunk_kernel(const double *__restrict__ C, int n_feat, int n_chunks, double *__restrict__ Cc) {
    const size_t total = (size_t)gridDim.y * n_chunks * EG_PB * EGP_CS;
    (void)total;
    const int pbk = blockIdx.y;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n_chunks * EG_PB * EGP_CS; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % EGP_CS);
        const size_t rc = i / EGP_CS;
        const int row = (int)(rc % EG_PB), c = (int)(rc / EG_PB);
        Cc[(size_t)pbk * n_chunks * EG_PB * EGP_CS + i] = k < EGP_KC ? C[((size_t)pbk * EG_PB + row) * n_feat + (size_t)c * EGP_KC + k] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------------
static size_t pm_egemm_smem(int n_atoms, int kc);

static pm_egemm_s *pm_egemm_plan(pm_handle h) {
    if (h->egemm_state) return (pm_egemm_s *)h->egemm_state;
    const PmProg &g = h->prog;
    // features per step: 32 halves the CTA barriers but measured 3 % slower than 16 on config 3 (16.1 vs 15.6 ms): the step is bound
    // by the imbalance between generating and contracting warps, not by the barrier count; PM_EGEMM_KC=32 selects it for sweeps
    const char *kc_env = getenv("PM_EGEMM_KC");
    const int EG_KC = (kc_env && atoi(kc_env) == 32 && pm_egemm_smem(g.n_atoms, 32) <= (size_t)h->max_smem) ? 32 : 16;
    std::vector<EgTerm> terms;
    std::vector<EgFeat> feats;
    std::vector<int> chunk_ptr(1, 0);
    int row = 0;
    auto add_term = [&](int kind, int a0, int a1, int a2, int a3, int per, int n_f, const EgFeat *f) {
        if (row + n_f > EG_KC) {  // a term never straddles two chunks: pad this one
            for (; row < EG_KC; ++row) feats.push_back(EgFeat{6, 0});
            chunk_ptr.push_back((int)terms.size());
            row = 0;
        }
        EgTerm t;
        t.a0 = a0; t.a1 = a1; t.a2 = a2; t.a3 = a3;
        t.kind = kind; t.per = per; t.row = row; t.pad = (int)chunk_ptr.size() - 1;
        terms.push_back(t);
        for (int k = 0; k < n_f; ++k) feats.push_back(f[k]);
        row += n_f;
    };
    for (int b = 0; b < g.n_bonds; ++b) {
        const EgFeat f[2] = {{0, g.doff_bond + 2 * b}, {1, g.doff_bond + 2 * b}};
        add_term(0, h->bond_idx[2 * b], h->bond_idx[2 * b + 1], 0, 0, 0, 2, f);
    }
    for (int a = 0; a < g.n_angles; ++a) {
        const EgFeat f[2] = {{0, g.doff_angle + 2 * a}, {1, g.doff_angle + 2 * a}};
        add_term(1, h->angle_idx[3 * a], h->angle_idx[3 * a + 1], h->angle_idx[3 * a + 2], 0, 0, 2, f);
    }
    for (int t = 0; t < g.n_torsions; ++t) {
        const EgFeat f[2] = {{2, g.doff_tors + 4 * t}, {3, g.doff_tors + 4 * t}};
        add_term(2, h->torsion_idx[4 * t], h->torsion_idx[4 * t + 1], h->torsion_idx[4 * t + 2], h->torsion_idx[4 * t + 3],
                 h->torsion_per[t], 2, f);
    }
    for (int q = 0; q < g.n_pairs; ++q) {
        const int o = g.doff_pair + g.pair_stride * q;
        const EgFeat f[3] = {{4, o}, {5, o + 1}, {4, o + 2}};
        add_term(3, h->host.pair_map[3 * q], h->host.pair_map[3 * q + 1], 0, 0, 0, 3, f);
    }
    {
        const EgFeat f[1] = {{7, 0}};
        add_term(4, 0, 0, 0, 0, 0, 1, f);
    }
    for (; row < EG_KC; ++row) feats.push_back(EgFeat{6, 0});
    chunk_ptr.push_back((int)terms.size());
    pm_egemm_s *e = new pm_egemm_s();
    e->kc = EG_KC;
    e->n_terms = (int)terms.size();
    e->n_feat = (int)feats.size();
    e->n_chunks = e->n_feat / EG_KC;
    e->n_bonds = g.n_bonds;
    e->n_angles = g.n_angles;
    e->n_torsions = g.n_torsions;
    e->terms_dev = nullptr;
    e->chunk_ptr_dev = nullptr;
    e->feat_dev = nullptr;
    if (cudaMalloc(&e->terms_dev, sizeof(EgTerm) * terms.size()) != cudaSuccess ||
        cudaMalloc(&e->chunk_ptr_dev, sizeof(int) * chunk_ptr.size()) != cudaSuccess ||
        cudaMalloc(&e->feat_dev, sizeof(EgFeat) * feats.size()) != cudaSuccess ||
        cudaMemcpy(e->terms_dev, terms.data(), sizeof(EgTerm) * terms.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(e->chunk_ptr_dev, chunk_ptr.data(), sizeof(int) * chunk_ptr.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(e->feat_dev, feats.data(), sizeof(EgFeat) * feats.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        pm_egemm_free(e);
        return nullptr;
    }
    h->egemm_state = e;
    h->egemm_free = pm_egemm_free;
    return e;
}

static size_t pm_egemm_smem(int n_atoms, int kc) {
    return sizeof(double) * ((size_t)3 * n_atoms * EG_ST + 2 * EG_PB * (kc + 4) + 2 * (size_t)kc * EG_GS + 4 * EG_PB * 4) +
           sizeof(double2) * (PM_ACOS_N + 1);
}

static size_t pm_egemm_burst_smem(int n_sc, int n_atoms_staged) {
    return sizeof(double) * ((size_t)EGB_STAGES * EG_PB * EGP_CS + (size_t)n_sc * EGP_KC * EG_GS + 4 * EG_PB * 4 + (size_t)3 * n_atoms_staged * EG_ST) +
           sizeof(double2) * (PM_ACOS_N + 1) + 16 * sizeof(uint64_t);
}

// burst size (chunks of 16 features) and whether the tile's coordinates are staged in shared memory: the largest burst up to 8 that
// fits beside the staged coordinates; molecules whose tile leaves room for fewer than 2 chunks read coordinates from global memory
static void pm_egemm_burst_plan(pm_handle h, int *n_sc, int *stage_x) {
    for (int sc = 8; sc >= 2; --sc)
        if (pm_egemm_burst_smem(sc, h->prog.n_atoms) <= (size_t)h->max_smem) {
            *n_sc = sc;
            *stage_x = 1;
            return;
        }
    *n_sc = 8;
    *stage_x = 0;
}

// 1: generation in bursts + TMA-streamed coefficients (default); 0: the chunk-by-chunk kernel (PM_EGEMM_BURST=0, sweeps)
static int pm_egemm_use_burst() {
    const char *e = getenv("PM_EGEMM_BURST");
    return !(e && e[0] == '0');
}

static int pm_egemm_nz(pm_handle h, int n_vec, long n_conf) {
    const int n_pb = (n_vec + EG_PB - 1) / EG_PB;
    long z = h->n_sm / n_pb;
    const long n_st = (n_conf + EG_ST - 1) / EG_ST;
    if (z > n_st) z = n_st;
    return (int)(z < 1 ? 1 : z);
}

extern "C" int pm_energy_batch_supported(pm_handle h) {
    if (!h) return 0;
    return !h->rf && h->prog.n_pairs + h->prog.n_bonds + h->prog.n_angles + h->prog.n_torsions > 0 &&
           (pm_egemm_use_burst() || pm_egemm_smem(h->prog.n_atoms, 16) <= (size_t)h->max_smem);
}

extern "C" long pm_energy_batch_scratch_doubles(pm_handle h, int n_vec, long n_conf) {
    if (!h || n_vec < 1 || n_conf < 1) return -1;
    const int n_pb = (n_vec + EG_PB - 1) / EG_PB;
    const long t_max = 2L * h->prog.n_bonds + 2L * h->prog.n_angles + 2L * h->prog.n_torsions + 3L * h->prog.n_pairs + 1;
    const long n_feat = ((t_max * 16 / 15 + 4 * 32) / 32) * 32;  // upper bound of the padded feature count
    // C row-major, its chunk-major padded copy (x 20/16), the rows
    return (long)n_pb * EG_PB * n_feat + (long)n_pb * EG_PB * (n_feat / 16) * EGP_CS + (long)n_vec * pm_egemm_nz(h, n_vec, n_conf) * 5;
}

extern "C" int pm_energy_batch_reduce(pm_handle h, const double *derived_dev, int n_vec, const double *coord_tiles_dev, long n_conf,
                                      const double *eref_dev, const double *weights_dev, double d_shift, double *partials_dev,
                                      double *scratch_dev, pm_comm comm, double *emm_dev, void *stream) {
    if (!h || !derived_dev || !coord_tiles_dev || !partials_dev || !scratch_dev) return pm_fail_msg("pm_energy_batch_reduce: null argument");
    if (n_vec < 1 || n_conf < 1) return pm_fail_msg("pm_energy_batch_reduce: empty batch");
    if (h->rf) return pm_fail_msg("pm_energy_batch_reduce: the reaction-field variant is not linear in the pair coefficients; use pm_eval_reduce");
    if (comm && n_vec > pm_comm_max_vec(comm)) return pm_fail_msg("pm_energy_batch_reduce: more parameter vectors than the communicator was created for");
    if (!pm_egemm_use_burst() && pm_egemm_smem(h->prog.n_atoms, 16) > (size_t)h->max_smem)
        return pm_fail_msg("pm_energy_batch_reduce: molecule too large for the coordinate tile of the chunk-by-chunk kernel");
    PMG_CUDA(cudaSetDevice(h->device));
    pm_egemm_s *e = pm_egemm_plan(h);
    if (!e) return pm_fail_msg("pm_energy_batch_reduce: could not build the feature table");
    const size_t smem = pm_egemm_smem(h->prog.n_atoms, e->kc);
    cudaStream_t st = (cudaStream_t)stream;
    const int n_pb = (n_vec + EG_PB - 1) / EG_PB, n_z = pm_egemm_nz(h, n_vec, n_conf);
    const int n_stages = (e->kc == EGP_KC && pm_egemm_use_burst()) ? EGB_STAGES : 0;
    double *C = scratch_dev;
    double *Cc = C + (size_t)n_pb * EG_PB * e->n_feat;
    double *rows = Cc + (n_stages ? (size_t)n_pb * e->n_chunks * EG_PB * EGP_CS : 0);
    const PmProg &g = h->prog;
    pm_egemm_coef_kernel<<<n_pb * EG_PB, 256, 0, st>>>(derived_dev, n_vec, g.n_derived, e->feat_dev, e->n_feat, g.doff_bond, g.n_bonds, g.doff_angle,
                                                       g.n_angles, g.doff_tors, g.n_torsions, C);
    PMG_CUDA(cudaGetLastError());
    if (n_stages) {
        pm_egemm_chunk_kernel<<<dim3(64, n_pb), 256, 0, st>>>(C, e->n_feat, e->n_chunks, Cc);
        PMG_CUDA(cudaGetLastError());
        int n_sc = 8, stage_x = 0;
        pm_egemm_burst_plan(h, &n_sc, &stage_x);
        const size_t smem_b = pm_egemm_burst_smem(n_sc, stage_x ? g.n_atoms : 0);
        PMG_CUDA(cudaFuncSetAttribute((const void *)pm_egemm_burst_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
        pm_egemm_burst_kernel<<<n_pb * n_z, EGB_THREADS, smem_b, st>>>(g.n_atoms, g.cpw, e->terms_dev, e->chunk_ptr_dev, e->n_chunks, Cc, g.acos_tab,
                                                                       coord_tiles_dev, n_conf, eref_dev, weights_dev, d_shift, n_vec, n_z,
                                                                       n_sc, stage_x, getenv("PM_EGEMM_NOGEN") ? 1 : 0, rows, emm_dev);   // (NOGEN: timing experiments)
        PMG_CUDA(cudaGetLastError());
        return pm_reduce_rows(rows, n_z, n_vec, n_conf, partials_dev, comm, st);
    }
    auto kernel = e->kc == 32 ? pm_egemm_kernel<32> : pm_egemm_kernel<16>;
    PMG_CUDA(cudaFuncSetAttribute((const void *)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<n_pb * n_z, EG_THREADS, smem, st>>>(g.n_atoms, g.cpw, e->terms_dev, e->chunk_ptr_dev, e->n_chunks, e->n_feat, C, g.acos_tab,
                                                 coord_tiles_dev, n_conf, eref_dev, weights_dev, d_shift, n_vec, n_z, rows, emm_dev);
    PMG_CUDA(cudaGetLastError());
    return pm_reduce_rows(rows, n_z, n_vec, n_conf, partials_dev, comm, st);
}
