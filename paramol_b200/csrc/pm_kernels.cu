// pm_kernels.cu -- sm_100a kernels + C ABI of the B200-native ParaMol objective-evaluation path.
//
// What is replaced (all file:line under /root/reference):
//   * the per-conformation  context.setPositions -> context.getState(getEnergy/getForces)  loops
//       ParaMol/System/system.py:324-354, ParaMol/MM_engines/openmm.py:448-502,
//       ParaMol/Objective_function/cpu_objective_function.py:67-108
//   * the per-conformation residual loops of the properties
//       ParaMol/Objective_function/Properties/energy_property.py:99-108, force_property.py:85-117
//   * force.set*Parameters + updateParametersInContext
//       ParaMol/MM_engines/openmm.py:507-588, 661-756
//
// Mapping: ONE LANE = ONE CONFORMATION.  A warp owns a tile of 32 conformations whose
// coordinates [3N][32] are pulled into shared memory with one TMA bulk copy; the topology "program"
// and the derived parameter table are staged once per CTA and read as warp-uniform broadcasts, so the
// term loops have no divergence, no shuffles and no atomics: forces accumulate in a lane-private
// column of a shared [3N][32] tile.  The O(N^2) pair loop keeps PM_IT atoms (coordinates + force
// accumulators) in registers so that shared-memory traffic per pair stays below the FP64-pipe time.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/paramol_b200.h"
#include "pm_device.cuh"

// =============================================================================================
// Device-visible description of one topology
// =============================================================================================
struct PmProg {
    int n_atoms, n_bonds, n_angles, n_torsions;
    int n_itiles, n_entries, n_pairs;
    int n_int;       // ints in the blob
    int n_derived;   // doubles per parameter vector in the derived table
    // int offsets into the blob
    int off_bond, off_angle, off_tors, off_tper, off_itile, off_entry;
    // double offsets into a derived row
    int doff_bond, doff_angle, doff_tors, doff_pair;
    int n_slots;
    int soff_bond, soff_angle, soff_tors, soff_part, soff_exc;
    double k_coul;
    const int *ints;        // device blob
    const double *metric;   // device [9N]
    const int *pair_map;    // device [n_pairs][3] = (i, j, exception or -1)
};

// =============================================================================================
// K0: slots -> derived table (one block per parameter vector)
// =============================================================================================
__global__ void pm_prepare_kernel(PmProg g, const double *__restrict__ slots, double *__restrict__ derived) {
    const double *s = slots + (size_t)blockIdx.x * g.n_slots;
    double *d = derived + (size_t)blockIdx.x * g.n_derived;
    for (int i = threadIdx.x; i < 2 * g.n_bonds; i += blockDim.x) d[g.doff_bond + i] = s[g.soff_bond + i];
    for (int i = threadIdx.x; i < 2 * g.n_angles; i += blockDim.x) d[g.doff_angle + i] = s[g.soff_angle + i];
    for (int t = threadIdx.x; t < g.n_torsions; t += blockDim.x) {
        const double phase = s[g.soff_tors + 2 * t], k = s[g.soff_tors + 2 * t + 1];
        const int n = g.ints[g.off_tper + t];
        double sn, cs;
        sincos(phase, &sn, &cs);
        double *o = d + g.doff_tors + 4 * t;
        o[0] = k;
        o[1] = k * (double)n;
        o[2] = cs;
        o[3] = sn;
    }
    for (int q = threadIdx.x; q < g.n_pairs; q += blockDim.x) {
        const int i = g.pair_map[3 * q], j = g.pair_map[3 * q + 1], e = g.pair_map[3 * q + 2];
        double qq, sig, eps;
        if (e < 0) {
            const double *pi = s + g.soff_part + 3 * i, *pj = s + g.soff_part + 3 * j;
            qq = pi[0] * pj[0];
            sig = 0.5 * pi[1] + 0.5 * pj[1];
            eps = sqrt(pi[2]) * sqrt(pj[2]);
        } else {
            const double *pe = s + g.soff_exc + 3 * e;
            qq = pe[0];
            sig = pe[1];
            eps = pe[2];
        }
        const double s2 = sig * sig, s6 = s2 * s2 * s2;
        double *o = d + g.doff_pair + 3 * q;
        o[0] = 4.0 * eps * s6 * s6;  // c12
        o[1] = 4.0 * eps * s6;       // c6
        o[2] = g.k_coul * qq;        // Coulomb product
    }
}

// =============================================================================================
// K1: energies (+ forces + force residual) of every conformation, for every parameter vector
// =============================================================================================
template <bool WITH_F>
__device__ __forceinline__ void pm_pair(const pm_d3 xi, const pm_d3 xj, const double *__restrict__ pp, double &e_nb, pm_d3 &fi,
                                        pm_d3 &fj) {
    const pm_d3 d = pm_sub(xi, xj);
    const double r2 = pm_dot(d, d);
    const double rinv = pm_rsqrt(r2);
    const double u = rinv * rinv;
    const double u3 = u * u * u;
    const double c12 = pp[0], c6 = pp[1], cq = pp[2];
    const double a = c12 * u3;
    const double elj = (a - c6) * u3;
    const double ec = cq * rinv;
    e_nb += elj;
    e_nb += ec;
    if (WITH_F) {
        // -(dE/dr)/r = (12 c12 u^6 - 6 c6 u^3 + cq rinv) u = (6 (elj + a u3) + ec) u
        const double g = fma(6.0, fma(a, u3, elj), ec) * u;
        fi.x = fma(g, d.x, fi.x);
        fi.y = fma(g, d.y, fi.y);
        fi.z = fma(g, d.z, fi.z);
        fj.x = fma(g, d.x, fj.x);
        fj.y = fma(g, d.y, fj.y);
        fj.z = fma(g, d.z, fj.z);
    }
}

template <bool WITH_F>
__global__ void __launch_bounds__(PM_MAX_WARPS * 32, 1)
    pm_ef_kernel(PmProg g, const double *__restrict__ derived, int n_vec, const double *__restrict__ xt,
                 const double *__restrict__ fref_t, long n_conf, long n_tiles, double *__restrict__ emm,
                 double *__restrict__ fres, double *__restrict__ fmm_t) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int N = g.n_atoms;
    const int tile_d = 3 * N * PM_TILE;

    // ---- shared memory carve-up -----------------------------------------------------------------
    int *s_int = reinterpret_cast<int *>(smem_raw);
    size_t off = ((size_t)g.n_int * sizeof(int) + 15) & ~(size_t)15;
    double *s_metric = reinterpret_cast<double *>(smem_raw + off);
    off += (size_t)9 * N * sizeof(double);
    double *s_der = reinterpret_cast<double *>(smem_raw + off);
    off += (size_t)g.n_derived * sizeof(double);
    off = (off + 127) & ~(size_t)127;
    const size_t per_warp = 128 + (size_t)tile_d * sizeof(double) * (WITH_F ? 2 : 1);
    unsigned char *wbase = smem_raw + off + per_warp * warp;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase);
    double *xs = reinterpret_cast<double *>(wbase + 128);
    double *fs = xs + tile_d;  // only touched when WITH_F

    for (int i = threadIdx.x; i < g.n_int; i += blockDim.x) s_int[i] = g.ints[i];
    for (int i = threadIdx.x; i < 9 * N; i += blockDim.x) s_metric[i] = g.metric[i];
    if (lane == 0) {
        pm_mbar_init(bar, 1);
        pm_fence_mbar_init();
    }
    uint32_t parity = 0;

    const int *s_bond = s_int + g.off_bond;
    const int *s_angle = s_int + g.off_angle;
    const int *s_tors = s_int + g.off_tors;
    const int *s_tper = s_int + g.off_tper;
    const int *s_itile = s_int + g.off_itile;
    const int2 *s_entry = reinterpret_cast<const int2 *>(s_int + g.off_entry);

    const long warp_stride = (long)gridDim.x * n_warps;
    const long first_tile = (long)blockIdx.x * n_warps + warp;

    for (int p = 0; p < n_vec; ++p) {
        __syncthreads();  // everybody done with the previous derived table (and the int blob is staged)
        {
            const double *dsrc = derived + (size_t)p * g.n_derived;
            for (int i = threadIdx.x; i < g.n_derived; i += blockDim.x) s_der[i] = dsrc[i];
        }
        __syncthreads();
        const double *s_pbond = s_der + g.doff_bond;
        const double *s_pangle = s_der + g.doff_angle;
        const double *s_ptors = s_der + g.doff_tors;
        const double *s_ppair = s_der + g.doff_pair;

        for (long tile = first_tile; tile < n_tiles; tile += warp_stride) {
            // ---- coordinates of 32 conformations: one TMA bulk copy into this warp's tile ---------
            if (lane == 0) {
                pm_fence_proxy_async();
                pm_mbar_expect_tx(bar, (uint32_t)(tile_d * sizeof(double)));
                pm_bulk_g2s(xs, xt + (size_t)tile * tile_d, (uint32_t)(tile_d * sizeof(double)), bar);
            }
            if (WITH_F) {
                for (int k = 0; k < 3 * N; ++k) fs[k * PM_TILE + lane] = 0.0;
            }
            pm_mbar_wait(bar, parity);
            parity ^= 1;

            double e_bond = 0.0, e_angle = 0.0, e_tors = 0.0, e_nb = 0.0;

            // ---- HarmonicBondForce ---------------------------------------------------------------
            for (int b = 0; b < g.n_bonds; ++b) {
                const int ia = s_bond[2 * b], ib = s_bond[2 * b + 1];
                const double r0 = s_pbond[2 * b], k = s_pbond[2 * b + 1];
                const pm_d3 d = pm_sub(pm_ld3(xs, ib, lane), pm_ld3(xs, ia, lane));
                const double r2 = pm_dot(d, d);
                const double rinv = pm_rsqrt(r2);
                const double di = fma(r2, rinv, -r0);
                const double kd = k * di;
                e_bond = fma(0.5 * kd, di, e_bond);
                if (WITH_F) {
                    const double s = kd * rinv;
                    const pm_d3 f{s * d.x, s * d.y, s * d.z};
                    pm_add3(fs, ia, lane, f);
                    pm_sub3(fs, ib, lane, f);
                }
            }

            // ---- HarmonicAngleForce --------------------------------------------------------------
            for (int a = 0; a < g.n_angles; ++a) {
                const int ia = s_angle[3 * a], ib = s_angle[3 * a + 1], ic = s_angle[3 * a + 2];
                const double t0 = s_pangle[2 * a], k = s_pangle[2 * a + 1];
                const pm_d3 xb = pm_ld3(xs, ib, lane);
                const pm_d3 d0 = pm_sub(xb, pm_ld3(xs, ia, lane));
                const pm_d3 d1 = pm_sub(xb, pm_ld3(xs, ic, lane));
                const double r20 = pm_dot(d0, d0), r21 = pm_dot(d1, d1);
                const double dot = pm_dot(d0, d1);
                const double inv01 = pm_rsqrt(r20 * r21);
                double cosine = dot * inv01;
                cosine = fmin(1.0, fmax(-1.0, cosine));
                const double theta = acos(cosine);
                const double di = theta - t0;
                const double kd = k * di;
                e_angle = fma(0.5 * kd, di, e_angle);
                if (WITH_F) {
                    const pm_d3 pv = pm_cross(d0, d1);
                    const double rp2 = pm_dot(pv, pv);
                    // 1/max(|p|, 1e-6)
                    const double rpinv = rp2 > 1.0e-12 ? pm_rsqrt(rp2) : 1.0e6;
                    // termA = kd / (r20 rp) ; termC = -kd / (r21 rp) ; 1/r20 = inv01^2 r21
                    const double w = kd * rpinv * inv01 * inv01;
                    const double termA = w * r21, termC = -w * r20;
                    const pm_d3 c0 = pm_cross(d0, pv), c2 = pm_cross(d1, pv);
                    const pm_d3 fa{termA * c0.x, termA * c0.y, termA * c0.z};
                    const pm_d3 fc{termC * c2.x, termC * c2.y, termC * c2.z};
                    pm_add3(fs, ia, lane, fa);
                    pm_sub3(fs, ib, lane, pm_d3{fa.x + fc.x, fa.y + fc.y, fa.z + fc.z});
                    pm_add3(fs, ic, lane, fc);
                }
            }

            // ---- PeriodicTorsionForce (trig-free: cos/sin(phi) from the cross products) ------------
            for (int t = 0; t < g.n_torsions; ++t) {
                const int i0 = s_tors[4 * t], i1 = s_tors[4 * t + 1], i2 = s_tors[4 * t + 2], i3 = s_tors[4 * t + 3];
                const int n = s_tper[t];
                const double k = s_ptors[4 * t], kn = s_ptors[4 * t + 1], cd = s_ptors[4 * t + 2], sd = s_ptors[4 * t + 3];
                const pm_d3 xb = pm_ld3(xs, i1, lane), xc = pm_ld3(xs, i2, lane);
                const pm_d3 d0 = pm_sub(pm_ld3(xs, i0, lane), xb);
                const pm_d3 d1 = pm_sub(xc, xb);
                const pm_d3 d2 = pm_sub(xc, pm_ld3(xs, i3, lane));
                const pm_d3 c1 = pm_cross(d0, d1), c2 = pm_cross(d1, d2);
                const double nc1 = pm_dot(c1, c1), nc2 = pm_dot(c2, c2), nb2 = pm_dot(d1, d1);
                const double inv12 = pm_rsqrt(nc1 * nc2);
                const double rb_inv = pm_rsqrt(nb2);
                const double rb = nb2 * rb_inv;
                const double cphi = pm_dot(c1, c2) * inv12;
                const double sphi = rb * pm_dot(d0, c2) * inv12;
                double cn = 1.0, sn = 0.0;
                for (int m = 0; m < n; ++m) {
                    const double c_new = fma(cn, cphi, -(sn * sphi));
                    sn = fma(sn, cphi, cn * sphi);
                    cn = c_new;
                }
                // cos(n phi - delta), sin(n phi - delta)
                const double cda = fma(cn, cd, sn * sd);
                const double sda = fma(sn, cd, -(cn * sd));
                e_tors = fma(k, cda, e_tors + k);
                if (WITH_F) {
                    const double dEdA = -kn * sda;
                    const double i2_12 = inv12 * inv12;
                    const double ff0 = -dEdA * rb * (i2_12 * nc2);  // / nc1
                    const double ff3 = dEdA * rb * (i2_12 * nc1);   // / nc2
                    const double ib2 = rb_inv * rb_inv;
                    const double ff1 = pm_dot(d0, d1) * ib2;
                    const double ff2 = pm_dot(d2, d1) * ib2;
                    const pm_d3 f0{ff0 * c1.x, ff0 * c1.y, ff0 * c1.z};
                    const pm_d3 f3{ff3 * c2.x, ff3 * c2.y, ff3 * c2.z};
                    const pm_d3 sv{fma(ff1, f0.x, -(ff2 * f3.x)), fma(ff1, f0.y, -(ff2 * f3.y)), fma(ff1, f0.z, -(ff2 * f3.z))};
                    pm_add3(fs, i0, lane, f0);
                    pm_sub3(fs, i1, lane, pm_d3{f0.x - sv.x, f0.y - sv.y, f0.z - sv.z});
                    pm_sub3(fs, i2, lane, pm_d3{f3.x + sv.x, f3.y + sv.y, f3.z + sv.z});
                    pm_add3(fs, i3, lane, f3);
                }
            }

            // ---- NonbondedForce: plain pairs and scaled 1-4 exceptions, register-tiled over i -------
            for (int it = 0; it < g.n_itiles; ++it) {
                const int *tl = s_itile + it * (PM_IT + 2);
                pm_d3 xi[PM_IT], acc[PM_IT];
#pragma unroll
                for (int k = 0; k < PM_IT; ++k) {
                    const int a = tl[k];
                    xi[k] = pm_ld3(xs, a < 0 ? 0 : a, lane);
                    acc[k] = pm_d3{0.0, 0.0, 0.0};
                }
                const int e_begin = tl[PM_IT], e_end = tl[PM_IT + 1];
                for (int e = e_begin; e < e_end; ++e) {
                    const int2 en = s_entry[e];
                    const int j = en.x & 0xffff;
                    const int mask = en.x >> 16;
                    const double *pp = s_ppair + 3 * en.y;
                    const pm_d3 xj = pm_ld3(xs, j, lane);
                    pm_d3 fj{0.0, 0.0, 0.0};
#pragma unroll
                    for (int k = 0; k < PM_IT; ++k) {
                        if (mask & (1 << k)) {
                            pm_pair<WITH_F>(xi[k], xj, pp, e_nb, acc[k], fj);
                            pp += 3;
                        }
                    }
                    if (WITH_F) pm_sub3(fs, j, lane, fj);
                }
                if (WITH_F) {
#pragma unroll
                    for (int k = 0; k < PM_IT; ++k) {
                        const int a = tl[k];
                        if (a >= 0) pm_add3(fs, a, lane, acc[k]);
                    }
                }
            }

            // ---- epilogue: per-conformation energy and force residual ------------------------------
            const long sconf = tile * PM_TILE + lane;
            const double e_total = ((e_bond + e_angle) + e_tors) + e_nb;
            if (sconf < n_conf) emm[(size_t)p * n_conf + sconf] = e_total;
            if (WITH_F) {
                if (fres != nullptr) {
                    const double *fr = fref_t + (size_t)tile * tile_d + lane;
                    double res = 0.0;
                    for (int a = 0; a < N; ++a) {
                        const double dx = fs[(3 * a) * PM_TILE + lane] - fr[(3 * a) * PM_TILE];
                        const double dy = fs[(3 * a + 1) * PM_TILE + lane] - fr[(3 * a + 1) * PM_TILE];
                        const double dz = fs[(3 * a + 2) * PM_TILE + lane] - fr[(3 * a + 2) * PM_TILE];
                        const double *m = s_metric + 9 * a;
                        // (d^T M) d as in np.matmul(np.matmul(d, M), d), force_property.py:93-94
                        const double t0 = fma(dz, m[6], fma(dy, m[3], dx * m[0]));
                        const double t1 = fma(dz, m[7], fma(dy, m[4], dx * m[1]));
                        const double t2 = fma(dz, m[8], fma(dy, m[5], dx * m[2]));
                        res += fma(t2, dz, fma(t1, dy, t0 * dx));
                    }
                    if (sconf < n_conf) fres[(size_t)p * n_conf + sconf] = res;
                }
                if (fmm_t != nullptr) {
                    double *fo = fmm_t + ((size_t)p * n_tiles + tile) * tile_d + lane;
                    for (int k = 0; k < 3 * N; ++k) fo[k * PM_TILE] = fs[k * PM_TILE + lane];
                }
            }
            __syncwarp();
        }
    }
}

// =============================================================================================
// [S][N][3] <-> tiles
// =============================================================================================
__global__ void pm_pack_kernel(const double *__restrict__ src, long n_conf, int row /* 3N */, double *__restrict__ tiles) {
    extern __shared__ double s_buf[];  // [32][row + 1]
    const long tile = blockIdx.x;
    const long s0 = tile * PM_TILE;
    const int total = PM_TILE * row;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i / row, k = i - c * row;
        long s = s0 + c;
        if (s >= n_conf) s = n_conf - 1;
        s_buf[c * (row + 1) + k] = src[(size_t)s * row + k];
    }
    __syncthreads();
    double *dst = tiles + (size_t)tile * total;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int k = i >> 5, c = i & 31;
        dst[i] = s_buf[c * (row + 1) + k];
    }
}

__global__ void pm_unpack_kernel(const double *__restrict__ tiles, long n_conf, int row, double *__restrict__ dst) {
    extern __shared__ double s_buf[];
    const long tile = blockIdx.x;
    const long s0 = tile * PM_TILE;
    const int total = PM_TILE * row;
    const double *src = tiles + (size_t)tile * total;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int k = i >> 5, c = i & 31;
        s_buf[c * (row + 1) + k] = src[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i / row, k = i - c * row;
        const long s = s0 + c;
        if (s < n_conf) dst[(size_t)s * row + k] = s_buf[c * (row + 1) + k];
    }
}

// =============================================================================================
// R1: weighted residual partial sums (deterministic two-stage tree, no atomics)
// =============================================================================================
#define PM_RED_BLOCKS 256
#define PM_RED_THREADS 256

__device__ __forceinline__ double pm_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void pm_reduce_stage1(const double *__restrict__ emm, const double *__restrict__ fres, const double *__restrict__ eref,
                                 const double *__restrict__ w, double d_shift, long n_conf, double *__restrict__ scratch) {
    const int p = blockIdx.y;
    const double *e = emm + (size_t)p * n_conf;
    const double *fr = fres ? fres + (size_t)p * n_conf : nullptr;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
    // contiguous chunk per block, fixed order inside -> bitwise reproducible for a given S
    const long chunk = (n_conf + gridDim.x - 1) / gridDim.x;
    const long lo = (long)blockIdx.x * chunk;
    long hi = lo + chunk;
    if (hi > n_conf) hi = n_conf;
    for (long s = lo + threadIdx.x; s < hi; s += blockDim.x) {
        const double d = (eref ? eref[s] : 0.0) - e[s] - d_shift;
        const double ws = w ? w[s] : 1.0;
        a0 += d;
        a1 = fma(ws, d, a1);
        a2 = fma(ws * d, d, a2);
        a3 += ws;
        if (fr) a4 = fma(ws, fr[s], a4);
    }
    __shared__ double sh[5][PM_RED_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    a0 = pm_warp_sum(a0);
    a1 = pm_warp_sum(a1);
    a2 = pm_warp_sum(a2);
    a3 = pm_warp_sum(a3);
    a4 = pm_warp_sum(a4);
    if (lane == 0) {
        sh[0][wid] = a0;
        sh[1][wid] = a1;
        sh[2][wid] = a2;
        sh[3][wid] = a3;
        sh[4][wid] = a4;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0;
        for (int i = 0; i < PM_RED_THREADS / 32; ++i) t += sh[threadIdx.x][i];
        scratch[((size_t)p * gridDim.x + blockIdx.x) * 5 + threadIdx.x] = t;
    }
}

__global__ void pm_reduce_stage2(const double *__restrict__ scratch, int n_blocks, long n_conf, double *__restrict__ partials) {
    const int p = blockIdx.x;
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;  // one warp per quantity
    if (q < 5) {
        double t = 0;
        for (int b = lane; b < n_blocks; b += 32) t += scratch[((size_t)p * n_blocks + b) * 5 + q];
        t = pm_warp_sum(t);
        if (lane == 0) partials[(size_t)p * PM_NPARTIAL + q] = t;
    } else if (threadIdx.x == 5 * 32) {
        partials[(size_t)p * PM_NPARTIAL + 5] = (double)n_conf;
        partials[(size_t)p * PM_NPARTIAL + 6] = 0.0;
        partials[(size_t)p * PM_NPARTIAL + 7] = 0.0;
    }
}

// =============================================================================================
// FMA-pipe peak microbenchmark (register-only; gives the roofline denominator for K1)
// =============================================================================================
template <typename T>
__global__ void pm_fma_peak_kernel(T *out, int iters, T b, T c) {
    T a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (T)(threadIdx.x + i) * (T)1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = a[i] * b + c;
        }
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// =============================================================================================
// Host side: handle, topology program builder, C ABI
// =============================================================================================
static thread_local std::string g_err;
static int pm_fail(const std::string &msg) {
    g_err = msg;
    return 1;
}
#define PM_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) return pm_fail(std::string(#call) + ": " + cudaGetErrorString(_e));      \
    } while (0)

struct pm_handle_s {
    int device;
    int n_sm;
    int max_smem;
    PmProg prog;
    std::vector<int> pairs;  // host copy of pair_map
    int *d_ints;
    double *d_metric;
    int *d_pair_map;
    size_t shared_bytes[2];  // CTA-shared part, [with_forces]
    int warps[2];
    size_t smem_total[2];
};

extern "C" const char *pm_last_error(void) { return g_err.c_str(); }
extern "C" int pm_version(void) { return 100; }
extern "C" int pm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
extern "C" long pm_n_tiles(long n_conf) { return (n_conf + PM_TILE - 1) / PM_TILE; }

static void pm_plan_launch(pm_handle_s *h) {
    const PmProg &g = h->prog;
    for (int wf = 0; wf < 2; ++wf) {
        size_t off = ((size_t)g.n_int * sizeof(int) + 15) & ~(size_t)15;
        off += (size_t)9 * g.n_atoms * sizeof(double);
        off += (size_t)g.n_derived * sizeof(double);
        off = (off + 127) & ~(size_t)127;
        const size_t per_warp = 128 + (size_t)3 * g.n_atoms * PM_TILE * sizeof(double) * (wf ? 2 : 1);
        long w = ((long)h->max_smem - (long)off) / (long)per_warp;
        if (w > PM_MAX_WARPS) w = PM_MAX_WARPS;
        h->shared_bytes[wf] = off;
        h->warps[wf] = (int)w;
        h->smem_total[wf] = w > 0 ? off + per_warp * (size_t)w : 0;
    }
}

extern "C" int pm_create(const pm_topology *t, int device, pm_handle *out) {
    if (!t || !out) return pm_fail("pm_create: null argument");
    if (t->n_atoms < 1 || t->n_atoms > 65535) return pm_fail("pm_create: n_atoms out of range");
    if (t->nonbonded_method != 0) return pm_fail("pm_create: only NoCutoff is implemented on the device path");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return pm_fail("pm_create: no CUDA device (this path has no CPU fallback)");
    }
    if (device < 0 || device >= n_dev) return pm_fail("pm_create: bad device index");
    PM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PM_CUDA(cudaGetDeviceProperties(&prop, device));

    const int N = t->n_atoms;
    // ---- classify atom pairs: 0 plain, 1 excluded, 2 + e : active exception e -----------------------
    std::vector<int> cls((size_t)N * N, 0);
    for (int e = 0; e < t->n_exceptions; ++e) {
        const int i = t->exception_idx[2 * e], j = t->exception_idx[2 * e + 1];
        if (i < 0 || j < 0 || i >= N || j >= N || i == j) return pm_fail("pm_create: bad exception indices");
        const int v = t->exception_active[e] ? 2 + e : 1;
        cls[(size_t)i * N + j] = v;
        cls[(size_t)j * N + i] = v;
    }
    auto check = [&](const int *idx, int n, int w, const char *what) -> bool {
        for (int k = 0; k < n * w; ++k)
            if (idx[k] < 0 || idx[k] >= N) {
                g_err = std::string("pm_create: bad atom index in ") + what;
                return false;
            }
        return true;
    };
    if (!check(t->bond_idx, t->n_bonds, 2, "bonds") || !check(t->angle_idx, t->n_angles, 3, "angles") ||
        !check(t->torsion_idx, t->n_torsions, 4, "torsions"))
        return 1;
    for (int k = 0; k < t->n_torsions; ++k)
        if (t->torsion_per[k] < 0 || t->torsion_per[k] > 64) return pm_fail("pm_create: torsion periodicity out of range");

    // ---- i-tiles and their j entries -------------------------------------------------------------
    const int n_itiles = (N + PM_IT - 1) / PM_IT;
    std::vector<int> itile((size_t)n_itiles * (PM_IT + 2));
    std::vector<int> entries;  // int2
    std::vector<int> pair_map; // (i, j, e)
    for (int it = 0; it < n_itiles; ++it) {
        int atoms[PM_IT];
        for (int k = 0; k < PM_IT; ++k) {
            const int a = it * PM_IT + k;
            atoms[k] = a < N ? a : -1;
            itile[(size_t)it * (PM_IT + 2) + k] = atoms[k];
        }
        itile[(size_t)it * (PM_IT + 2) + PM_IT] = (int)(entries.size() / 2);
        for (int j = atoms[0] + 1; j < N; ++j) {
            int mask = 0;
            const int base = (int)(pair_map.size() / 3);
            for (int k = 0; k < PM_IT; ++k) {
                const int a = atoms[k];
                if (a < 0 || a >= j) continue;
                const int c = cls[(size_t)a * N + j];
                if (c == 1) continue;
                mask |= 1 << k;
                pair_map.push_back(a);
                pair_map.push_back(j);
                pair_map.push_back(c == 0 ? -1 : c - 2);
            }
            if (mask) {
                entries.push_back(j | (mask << 16));
                entries.push_back(base);
            }
        }
        itile[(size_t)it * (PM_IT + 2) + PM_IT + 1] = (int)(entries.size() / 2);
    }

    pm_handle_s *h = new pm_handle_s();
    memset(&h->prog, 0, sizeof(PmProg));
    h->device = device;
    h->n_sm = prop.multiProcessorCount;
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    h->d_ints = nullptr;
    h->d_metric = nullptr;
    h->d_pair_map = nullptr;
    PmProg &g = h->prog;
    g.n_atoms = N;
    g.n_bonds = t->n_bonds;
    g.n_angles = t->n_angles;
    g.n_torsions = t->n_torsions;
    g.n_itiles = n_itiles;
    g.n_entries = (int)(entries.size() / 2);
    g.n_pairs = (int)(pair_map.size() / 3);
    g.k_coul = t->one_4pi_eps0;

    std::vector<int> blob;
    auto append = [&](const int *p, size_t n) {
        int o = (int)blob.size();
        blob.insert(blob.end(), p, p + n);
        return o;
    };
    g.off_bond = append(t->bond_idx, (size_t)2 * t->n_bonds);
    g.off_angle = append(t->angle_idx, (size_t)3 * t->n_angles);
    g.off_tors = append(t->torsion_idx, (size_t)4 * t->n_torsions);
    g.off_tper = append(t->torsion_per, (size_t)t->n_torsions);
    g.off_itile = append(itile.data(), itile.size());
    if (blob.size() & 1) blob.push_back(0);  // int2 alignment of the entry table
    g.off_entry = append(entries.data(), entries.size());
    g.n_int = (int)blob.size();

    g.doff_bond = 0;
    g.doff_angle = g.doff_bond + 2 * g.n_bonds;
    g.doff_tors = g.doff_angle + 2 * g.n_angles;
    g.doff_pair = g.doff_tors + 4 * g.n_torsions;
    g.n_derived = g.doff_pair + 3 * g.n_pairs;
    g.n_derived = (g.n_derived + 1) & ~1;

    g.soff_bond = 0;
    g.soff_angle = g.soff_bond + 2 * g.n_bonds;
    g.soff_tors = g.soff_angle + 2 * g.n_angles;
    g.soff_part = g.soff_tors + 2 * g.n_torsions;
    g.soff_exc = g.soff_part + 3 * N;
    g.n_slots = g.soff_exc + 3 * t->n_exceptions;

    h->pairs = pair_map;
    auto cleanup = [&]() {
        if (h->d_ints) cudaFree(h->d_ints);
        if (h->d_metric) cudaFree(h->d_metric);
        if (h->d_pair_map) cudaFree(h->d_pair_map);
        delete h;
    };
#define PM_CUDA_H(call)                                                          \
    do {                                                                         \
        cudaError_t _e = (call);                                                 \
        if (_e != cudaSuccess) {                                                 \
            cleanup();                                                           \
            return pm_fail(std::string(#call) + ": " + cudaGetErrorString(_e));  \
        }                                                                        \
    } while (0)
    PM_CUDA_H(cudaMalloc(&h->d_ints, sizeof(int) * (blob.size() + 2)));
    PM_CUDA_H(cudaMemcpy(h->d_ints, blob.data(), sizeof(int) * blob.size(), cudaMemcpyHostToDevice));
    PM_CUDA_H(cudaMalloc(&h->d_pair_map, sizeof(int) * (pair_map.size() + 3)));
    if (!pair_map.empty())
        PM_CUDA_H(cudaMemcpy(h->d_pair_map, pair_map.data(), sizeof(int) * pair_map.size(), cudaMemcpyHostToDevice));
    std::vector<double> ident((size_t)9 * N, 0.0);
    for (int a = 0; a < N; ++a) ident[9 * a] = ident[9 * a + 4] = ident[9 * a + 8] = 1.0;
    PM_CUDA_H(cudaMalloc(&h->d_metric, sizeof(double) * 9 * N));
    PM_CUDA_H(cudaMemcpy(h->d_metric, ident.data(), sizeof(double) * 9 * N, cudaMemcpyHostToDevice));
    g.ints = h->d_ints;
    g.metric = h->d_metric;
    g.pair_map = h->d_pair_map;

    pm_plan_launch(h);
    if (h->warps[1] < 1 || h->warps[0] < 1) {
        cleanup();
        return pm_fail("pm_create: molecule too large for the shared-memory tile of this kernel generation");
    }
    PM_CUDA_H(cudaFuncSetAttribute(pm_ef_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_total[1]));
    PM_CUDA_H(cudaFuncSetAttribute(pm_ef_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_total[0]));
    *out = h;
    return 0;
}

extern "C" int pm_destroy(pm_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaFree(h->d_ints);
    cudaFree(h->d_metric);
    cudaFree(h->d_pair_map);
    delete h;
    return 0;
}

extern "C" int pm_n_slots(pm_handle h) { return h ? h->prog.n_slots : -1; }
extern "C" int pm_n_derived(pm_handle h) { return h ? h->prog.n_derived : -1; }
extern "C" int pm_n_pairs(pm_handle h) { return h ? h->prog.n_pairs : -1; }
extern "C" long pm_tile_doubles(pm_handle h) { return h ? (long)3 * h->prog.n_atoms * PM_TILE : -1; }
extern "C" int pm_get_pairs(pm_handle h, int *out_host) {
    if (!h || !out_host) return pm_fail("pm_get_pairs: null argument");
    memcpy(out_host, h->pairs.data(), sizeof(int) * h->pairs.size());
    return 0;
}
extern "C" int pm_launch_info(pm_handle h, int with_forces, int *grid, int *block, int *smem_bytes, int *warps_per_cta) {
    if (!h) return pm_fail("pm_launch_info: null handle");
    const int wf = with_forces ? 1 : 0;
    if (grid) *grid = h->n_sm;
    if (block) *block = h->warps[wf] * 32;
    if (smem_bytes) *smem_bytes = (int)h->smem_total[wf];
    if (warps_per_cta) *warps_per_cta = h->warps[wf];
    return 0;
}

extern "C" int pm_set_force_metric(pm_handle h, const double *metric_host) {
    if (!h || !metric_host) return pm_fail("pm_set_force_metric: null argument");
    PM_CUDA(cudaSetDevice(h->device));
    PM_CUDA(cudaMemcpy(h->d_metric, metric_host, sizeof(double) * 9 * h->prog.n_atoms, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int pm_pack_tiles(pm_handle h, const double *src_dev, long n_conf, double *tiles_dev, void *stream) {
    if (!h || !src_dev || !tiles_dev) return pm_fail("pm_pack_tiles: null argument");
    if (n_conf < 1) return pm_fail("pm_pack_tiles: no conformations");
    PM_CUDA(cudaSetDevice(h->device));
    const int row = 3 * h->prog.n_atoms;
    const size_t smem = sizeof(double) * PM_TILE * (row + 1);
    if (smem > 48 * 1024) PM_CUDA(cudaFuncSetAttribute(pm_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pm_pack_kernel<<<(unsigned)pm_n_tiles(n_conf), 256, smem, (cudaStream_t)stream>>>(src_dev, n_conf, row, tiles_dev);
    PM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_unpack_tiles(pm_handle h, const double *tiles_dev, long n_conf, double *dst_dev, void *stream) {
    if (!h || !dst_dev || !tiles_dev) return pm_fail("pm_unpack_tiles: null argument");
    if (n_conf < 1) return pm_fail("pm_unpack_tiles: no conformations");
    PM_CUDA(cudaSetDevice(h->device));
    const int row = 3 * h->prog.n_atoms;
    const size_t smem = sizeof(double) * PM_TILE * (row + 1);
    if (smem > 48 * 1024) PM_CUDA(cudaFuncSetAttribute(pm_unpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pm_unpack_kernel<<<(unsigned)pm_n_tiles(n_conf), 256, smem, (cudaStream_t)stream>>>(tiles_dev, n_conf, row, dst_dev);
    PM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_prepare_params(pm_handle h, const double *slots_dev, int n_vec, double *derived_dev, void *stream) {
    if (!h || !slots_dev || !derived_dev) return pm_fail("pm_prepare_params: null argument");
    if (n_vec < 1) return pm_fail("pm_prepare_params: n_vec < 1");
    PM_CUDA(cudaSetDevice(h->device));
    pm_prepare_kernel<<<n_vec, 128, 0, (cudaStream_t)stream>>>(h->prog, slots_dev, derived_dev);
    PM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_eval(pm_handle h, const double *derived_dev, int n_vec, const double *coord_tiles_dev,
                       const double *fref_tiles_dev, long n_conf, int with_forces, double *emm_dev, double *fres_dev,
                       double *fmm_tiles_dev, void *stream) {
    if (!h || !derived_dev || !coord_tiles_dev || !emm_dev) return pm_fail("pm_eval: null argument");
    if (n_vec < 1 || n_conf < 1) return pm_fail("pm_eval: empty batch");
    if (fres_dev && !fref_tiles_dev) return pm_fail("pm_eval: fres requested without reference forces");
    if (!with_forces && (fres_dev || fmm_tiles_dev)) return pm_fail("pm_eval: force outputs requested with with_forces=0");
    PM_CUDA(cudaSetDevice(h->device));
    const int wf = with_forces ? 1 : 0;
    const long n_tiles = pm_n_tiles(n_conf);
    int warps = h->warps[wf];
    // do not launch more warps than tiles
    long grid = h->n_sm;
    if ((long)grid * warps > n_tiles) {
        grid = (n_tiles + warps - 1) / warps;
        if (grid < 1) grid = 1;
    }
    const size_t smem = h->smem_total[wf];
    if (wf)
        pm_ef_kernel<true><<<(unsigned)grid, warps * 32, smem, (cudaStream_t)stream>>>(
            h->prog, derived_dev, n_vec, coord_tiles_dev, fref_tiles_dev, n_conf, n_tiles, emm_dev, fres_dev, fmm_tiles_dev);
    else
        pm_ef_kernel<false><<<(unsigned)grid, warps * 32, smem, (cudaStream_t)stream>>>(
            h->prog, derived_dev, n_vec, coord_tiles_dev, fref_tiles_dev, n_conf, n_tiles, emm_dev, nullptr, nullptr);
    PM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" long pm_reduce_scratch_doubles(int n_vec) { return (long)n_vec * PM_RED_BLOCKS * 5; }

extern "C" int pm_reduce_objective(const double *emm_dev, const double *fres_dev, const double *eref_dev, const double *weights_dev,
                                   double d_shift, int n_vec, long n_conf, double *partials_dev, double *scratch_dev, void *stream) {
    if (!emm_dev || !partials_dev || !scratch_dev) return pm_fail("pm_reduce_objective: null argument");
    if (n_vec < 1 || n_conf < 1) return pm_fail("pm_reduce_objective: empty batch");
    int nb = PM_RED_BLOCKS;
    const long need = (n_conf + PM_RED_THREADS - 1) / PM_RED_THREADS;
    if (need < nb) nb = (int)need;
    dim3 grid(nb, n_vec);
    pm_reduce_stage1<<<grid, PM_RED_THREADS, 0, (cudaStream_t)stream>>>(emm_dev, fres_dev, eref_dev, weights_dev, d_shift, n_conf,
                                                                       scratch_dev);
    PM_CUDA(cudaGetLastError());
    pm_reduce_stage2<<<n_vec, 6 * 32, 0, (cudaStream_t)stream>>>(scratch_dev, nb, n_conf, partials_dev);
    PM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_measure_fma_peak(int device, int fp64, int iters, double *flops, double *ms_out) {
    if (!flops) return pm_fail("pm_measure_fma_peak: null argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return pm_fail("pm_measure_fma_peak: no CUDA device");
    }
    PM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PM_CUDA(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8;
    void *buf = nullptr;
    PM_CUDA(cudaMalloc(&buf, (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t e0, e1;
    PM_CUDA(cudaEventCreate(&e0));
    PM_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        PM_CUDA(cudaEventRecord(e0));
        if (fp64)
            pm_fma_peak_kernel<double><<<blocks, threads>>>((double *)buf, iters, 1.0000001, 1e-9);
        else
            pm_fma_peak_kernel<float><<<blocks, threads>>>((float *)buf, iters, 1.0000001f, 1e-9f);
        PM_CUDA(cudaEventRecord(e1));
        PM_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        PM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    const double n_fma = (double)threads * blocks * (double)iters * 64.0;
    *flops = 2.0 * n_fma / (best * 1e-3);
    if (ms_out) *ms_out = best;
    return 0;
}
