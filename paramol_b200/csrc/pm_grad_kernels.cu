// pm_grad_kernels.cu -- analytic derivatives of the fitting objective with respect to the force-field parameters (sm_100a).
//
// The reference has no analytic gradient: its optimisers difference ObjectiveFunction.f (scipy jac="2-point",
// ParaMol/Utils/settings.py:64; Optimizers/gradient_descent.py:106-121), i.e. P_opt + 1 objective evaluations per gradient.
// SURVEY.md section 8(f) rank 3 asks for the closed form instead.  For one parameter vector the objective is
//     f = sum_s [ energy part (E_s) + b_s sum_a (F_sa - Fref_sa)^T M_a (F_sa - Fref_sa) ] + regularisation
// (energy_property.py:99-108, force_property.py:85-117), so with per-conformation coefficients a_s = d f / d E_s (computed by
// the caller from the reduced sums of the forward pass) and lambda_sa = b_s (M_a + M_a^T)(F_sa - Fref_sa)
//     d f / d p = sum_s [ a_s dE_s/dp + sum_a lambda_sa . dF_sa/dp ].
// Every term has ONE internal coordinate q (distance, angle, dihedral), E_t = E_t(q; p) and F = -dE/dq grad q, hence
//     sum_a lambda_a . dF_a/dp = -(d^2 E_t / dq dp) (grad q . lambda),
// which needs nothing beyond what the force code already forms.  Parameters are the entries of the DERIVED table of
// pm_prepare_params for pairs (c12, c6, K qq) and the slots themselves for bonded terms (r0, k | theta0, k | phase, k); the
// caller chains the pair part back to (q, sigma, eps) slots.
//
// Mapping: a CTA stages a tile of conformations (coordinates and lambda, [conformation][atom][x y z lx ly lz]) in shared
// memory; every LANE owns one term and loops over the conformations of the tile with its accumulators in registers, so there
// is no cross-lane reduction and no atomic: after the tile the lane adds its sums into its private slot of the CTA's partial
// array (L2-resident), and a second kernel adds the partials of all CTAs in a fixed order.  Deterministic by construction.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "pm_device.cuh"
#include "pm_internal.hpp"

#define PMG_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) return pm_fail_msg(std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define PMG_THREADS 256
#define PMG_MAX_SMEM (100 * 1024)  // at least two CTAs per SM

enum { PMG_BOND = 0, PMG_ANGLE = 1, PMG_TORSION = 2, PMG_PAIR = 3 };

struct PmGradArgs;

struct PmGradState {
    int n_blocks;        // blocks of 32 terms of one kind
    int n_gout;          // doubles in the gradient output: 2 nb + 2 na + 2 nt + 3 np
    int4 *d_atoms;       // [n_blocks * 32]
    int4 *d_meta;        // [n_blocks * 32] = (kind or -1 for padding, offset in the derived table, periodicity, offset in gout)
    int n_items;         // work items = (block, slice of the tile's conformations); = n_blocks unless blocks are split
    int4 *d_items;       // [n_items] = (block, slice index, number of slices of that block, 0), grouped by warp
    int *d_warp_ptr;     // [warps + 1] into d_items
    int cpt;             // conformations per CTA tile
    int threads;         // 32 x warps
    size_t smem;
    int grid_max;
    int minb;
    void (*fn)(const PmGradArgs);
};

static void pm_grad_free(void *p) {
    PmGradState *st = (PmGradState *)p;
    if (!st) return;
    cudaFree(st->d_atoms);
    cudaFree(st->d_meta);
    cudaFree(st->d_items);
    cudaFree(st->d_warp_ptr);
    delete st;
}

struct PmGradArgs {
    int n_atoms, cpw, cpt, n_blocks, n_items;
    const int4 *items;
    const int *warp_ptr;
    long n_conf, n_tiles;  // n_tiles = CTA tiles of cpt conformations
    const int4 *atoms, *meta;
    const double *derived, *x_tiles, *fmm_tiles, *fref_tiles, *metric, *coef_e, *coef_f;
    const double2 *acos_tab;
    double *partial;  // [gridDim.x][n_items][4][32]
};

__device__ __forceinline__ void pmg_load(const double *__restrict__ s, pm_d3 &x, pm_d3 &l) {
    const double2 a = *reinterpret_cast<const double2 *>(s), b = *reinterpret_cast<const double2 *>(s + 2),
                  c = *reinterpret_cast<const double2 *>(s + 4);
    x = pm_d3{a.x, a.y, b.x};
    l = pm_d3{b.y, c.x, c.y};
}

template <int MINB>
__global__ void __launch_bounds__(PMG_THREADS, MINB) pm_grad_kernel(const PmGradArgs g) {
    extern __shared__ __align__(16) double sm[];
    const int N = g.n_atoms, CPT = g.cpt;
    double *tile = sm;                       // [CPT][N][6]
    double *ca = tile + (size_t)CPT * N * 6;  // [CPT] energy coefficients
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long tile_d = (long)3 * N * g.cpw;
    double *my_partial = g.partial + (size_t)blockIdx.x * g.n_items * 128;

    for (long t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
        const long c0 = t * CPT;
        const int nc = (int)min((long)CPT, g.n_conf - c0);
        __syncthreads();  // the previous tile is no longer read
        for (int idx = tid; idx < nc * N; idx += blockDim.x) {
            const int a = idx / nc, c = idx - a * nc;
            const long s = c0 + c;
            const long base = (s / g.cpw) * tile_d + (s % g.cpw);
            const double x = g.x_tiles[base + (size_t)(3 * a) * g.cpw], y = g.x_tiles[base + (size_t)(3 * a + 1) * g.cpw],
                         z = g.x_tiles[base + (size_t)(3 * a + 2) * g.cpw];
            double lx = 0.0, ly = 0.0, lz = 0.0;
            if (g.coef_f) {
                const double b = g.coef_f[s];
                const double dx = g.fmm_tiles[base + (size_t)(3 * a) * g.cpw] - g.fref_tiles[base + (size_t)(3 * a) * g.cpw];
                const double dy = g.fmm_tiles[base + (size_t)(3 * a + 1) * g.cpw] - g.fref_tiles[base + (size_t)(3 * a + 1) * g.cpw];
                const double dz = g.fmm_tiles[base + (size_t)(3 * a + 2) * g.cpw] - g.fref_tiles[base + (size_t)(3 * a + 2) * g.cpw];
                const double *m = g.metric + 9 * a;  // d (d^T M d) = (M + M^T) d
                lx = b * ((m[0] + m[0]) * dx + (m[1] + m[3]) * dy + (m[2] + m[6]) * dz);
                ly = b * ((m[3] + m[1]) * dx + (m[4] + m[4]) * dy + (m[5] + m[7]) * dz);
                lz = b * ((m[6] + m[2]) * dx + (m[7] + m[5]) * dy + (m[8] + m[8]) * dz);
            }
            double *o = tile + ((size_t)c * N + a) * 6;
            *reinterpret_cast<double2 *>(o) = make_double2(x, y);
            *reinterpret_cast<double2 *>(o + 2) = make_double2(z, lx);
            *reinterpret_cast<double2 *>(o + 4) = make_double2(ly, lz);
        }
        for (int c = tid; c < nc; c += blockDim.x) ca[c] = g.coef_e ? g.coef_e[c0 + c] : 0.0;
        __syncthreads();

        for (int item = g.warp_ptr[warp]; item < g.warp_ptr[warp + 1]; ++item) {
            const int4 it4 = g.items[item];  // (block, slice, slices of the block): this warp does conformations [cb, ce) of the tile
            const int blk = it4.x;
            const int cb = (int)((long)nc * it4.y / it4.z), ce = (int)((long)nc * (it4.y + 1) / it4.z);
            const int4 at = g.atoms[blk * 32 + lane], me = g.meta[blk * 32 + lane];
            const int kind = __shfl_sync(0xffffffffu, me.x, 0);  // uniform per block (lane 0 is never padding)
            double *pp = my_partial + (size_t)item * 128 + lane;
            double p0 = pp[0], p1 = pp[32], p2 = pp[64];  // requested now, needed after the loop
            double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
            if (me.x >= 0) {
                const double *dp = g.derived + me.y;
                const int stride = N * 6;
                if (kind == PMG_PAIR) {
                    const double *si = tile + at.x * 6, *sj = tile + at.y * 6;
#pragma unroll 2
                    for (int c = cb; c < ce; ++c) {
                        pm_d3 xi, li, xj, lj;
                        pmg_load(si + c * stride, xi, li);
                        pmg_load(sj + c * stride, xj, lj);
                        const pm_d3 d = pm_sub(xi, xj);
                        const double rinv = pm_rsqrt(pm_dot(d, d));
                        const double u = rinv * rinv, u3 = u * u * u, e12 = u3 * u3;
                        const double us = u * pm_dot(d, pm_sub(li, lj));  // (dr along lambda) / r
                        const double a = ca[c];
                        acc0 += e12 * fma(12.0, us, a);    // d/d c12:  a r^-12 + 12 r^-13 dr
                        acc1 -= u3 * fma(6.0, us, a);      // d/d c6 : -a r^-6  -  6 r^-7  dr
                        acc2 += rinv * (a + us);           // d/d Kqq:  a r^-1  +    r^-2  dr
                    }
                } else if (kind == PMG_BOND) {
                    const double r0 = dp[0], k = dp[1];
                    const double *si = tile + at.x * 6, *sj = tile + at.y * 6;
                    for (int c = cb; c < ce; ++c) {
                        pm_d3 xi, li, xj, lj;
                        pmg_load(si + c * stride, xi, li);
                        pmg_load(sj + c * stride, xj, lj);
                        const pm_d3 d = pm_sub(xj, xi);
                        const double r2 = pm_dot(d, d), rinv = pm_rsqrt(r2);
                        const double dl = r2 * rinv - r0;
                        const double dr = rinv * pm_dot(d, pm_sub(lj, li));
                        const double a = ca[c];
                        acc0 += k * (dr - a * dl);               // d/d r0: -a k (r - r0) + k dr
                        acc1 += dl * fma(0.5 * a, dl, -dr);      // d/d k :  a (r - r0)^2 / 2 - (r - r0) dr
                    }
                } else if (kind == PMG_ANGLE) {
                    const double t0 = dp[0], k = dp[1];
                    const double *sa = tile + at.x * 6, *sb = tile + at.y * 6, *sc = tile + at.z * 6;
                    for (int c = cb; c < ce; ++c) {
                        pm_d3 xa, la, xb, lb, xc, lc;
                        pmg_load(sa + c * stride, xa, la);
                        pmg_load(sb + c * stride, xb, lb);
                        pmg_load(sc + c * stride, xc, lc);
                        const pm_d3 d0 = pm_sub(xb, xa), d1 = pm_sub(xb, xc);
                        const pm_d3 p = pm_cross(d0, d1);
                        const double r0 = pm_dot(d0, d0), r1 = pm_dot(d1, d1), pp2 = pm_dot(p, p);
                        const double i01 = pm_rsqrt(r0 * r1);          // 1 / (|d0| |d1|)
                        const double ip = pm_rsqrt(fmax(pp2, 1.0e-12));  // 1 / max(|p|, 1e-6), as ReferenceAngleBondIxn
                        const double th = pm_acos_cs(pm_dot(d0, d1) * i01, pp2 * ip * i01, g.acos_tab);
                        // grad_a theta = -(d0 x p) / (|d0|^2 |p|),  grad_c theta = (d1 x p) / (|d1|^2 |p|)
                        const double ir0 = r1 * i01 * i01, ir1 = r0 * i01 * i01;  // 1 / |d0|^2, 1 / |d1|^2
                        const double dth = ip * (ir1 * pm_dot(pm_cross(d1, p), pm_sub(lc, lb)) - ir0 * pm_dot(pm_cross(d0, p), pm_sub(la, lb)));
                        const double dl = th - t0;
                        const double a = ca[c];
                        acc0 += k * (dth - a * dl);
                        acc1 += dl * fma(0.5 * a, dl, -dth);
                    }
                } else {  // PMG_TORSION: derived = (k, k n, cos phase, sin phase)
                    const double k = dp[0], cph = dp[2], sph = dp[3];
                    const int per = me.z;
                    const double n = (double)per;
                    const double *s0 = tile + at.x * 6, *s1 = tile + at.y * 6, *s2 = tile + at.z * 6, *s3 = tile + at.w * 6;
                    for (int c = cb; c < ce; ++c) {
                        pm_d3 x0, l0, x1, l1, x2, l2, x3, l3;
                        pmg_load(s0 + c * stride, x0, l0);
                        pmg_load(s1 + c * stride, x1, l1);
                        pmg_load(s2 + c * stride, x2, l2);
                        pmg_load(s3 + c * stride, x3, l3);
                        const pm_d3 d0 = pm_sub(x0, x1), d1 = pm_sub(x2, x1), d2 = pm_sub(x2, x3);
                        const pm_d3 c1 = pm_cross(d0, d1), c2 = pm_cross(d1, d2);
                        const double n1 = pm_dot(c1, c1), n2 = pm_dot(c2, c2), bb = pm_dot(d1, d1);
                        // one reciprocal for 1/n1, 1/n2, 1/bb; phi is the signed angle between c1 and c2 (sign of d0 . c2, as
                        // ReferenceProperDihedralBond): cos phi = c1.c2 / (|c1||c2|), sin phi = |d1| (d0.c2) / (|c1||c2|)
                        const double inv = 1.0 / (n1 * n2 * bb);
                        const double in1 = inv * n2 * bb, in2 = inv * n1 * bb, ibb = inv * n1 * n2;
                        const double i12 = pm_rsqrt(n1 * n2), nb = bb * pm_rsqrt(bb);
                        const double cphi = pm_dot(c1, c2) * i12, sphi = nb * pm_dot(d0, c2) * i12;
                        double cn = 1.0, sn = 0.0;  // cos(n phi), sin(n phi) by n complex multiplications (n is 1..6 in practice)
                        for (int q = 0; q < per; ++q) {
                            const double t = cn * cphi - sn * sphi;
                            sn = fma(sn, cphi, cn * sphi);
                            cn = t;
                        }
                        const double cpsi = cn * cph + sn * sph, spsi = sn * cph - cn * sph;  // psi = n phi - phase
                        // grad_0 phi = |d1| c1 / |c1|^2, grad_3 phi = -|d1| c2 / |c2|^2,
                        // grad_1 = (ff1 - 1) grad_0 - ff2 grad_3, grad_2 = (ff2 - 1) grad_3 - ff1 grad_0
                        const double ff1 = pm_dot(d0, d1) * ibb, ff2 = pm_dot(d2, d1) * ibb;
                        const pm_d3 w0{l0.x + (ff1 - 1.0) * l1.x - ff1 * l2.x, l0.y + (ff1 - 1.0) * l1.y - ff1 * l2.y,
                                       l0.z + (ff1 - 1.0) * l1.z - ff1 * l2.z};
                        const pm_d3 w3{l3.x - ff2 * l1.x + (ff2 - 1.0) * l2.x, l3.y - ff2 * l1.y + (ff2 - 1.0) * l2.y,
                                       l3.z - ff2 * l1.z + (ff2 - 1.0) * l2.z};
                        const double dphi = nb * (pm_dot(c1, w0) * in1 - pm_dot(c2, w3) * in2);
                        const double a = ca[c];
                        acc0 += k * (a * spsi - n * cpsi * dphi);   // d/d phase: a k sin psi - k n cos psi dphi
                        acc1 += fma(a, 1.0 + cpsi, n * spsi * dphi);  // d/d k   : a (1 + cos psi) + n sin psi dphi
                    }
                }
            }
            pp[0] = p0 + acc0;
            pp[32] = p1 + acc1;
            pp[64] = p2 + acc2;
        }
    }
}

// gout[meta.w + q] = sum over CTAs and over the slices of the term's block (fixed order) of the lane's partial sums
__global__ void pm_grad_reduce_kernel(const double *__restrict__ partial, int n_cta, int n_blocks, int n_items,
                                      const int4 *__restrict__ items, const int4 *__restrict__ meta, double *__restrict__ gout) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_blocks * 32) return;
    const int4 me = meta[e];
    if (me.x < 0) return;
    const int blk = e >> 5, lane = e & 31;
    const int nq = me.x == PMG_PAIR ? 3 : 2;
    for (int q = 0; q < nq; ++q) {
        double s = 0.0;
        for (int item = 0; item < n_items; ++item) {
            if (items[item].x != blk) continue;
            for (int c = 0; c < n_cta; ++c) s += partial[((size_t)c * n_items + item) * 128 + q * 32 + lane];
        }
        gout[me.w + q] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static PmGradState *pm_grad_state(pm_handle h) {
    if (h->grad_state) return (PmGradState *)h->grad_state;
    const PmProg &g = h->prog;
    std::vector<int4> atoms, meta;
    auto pad = [&]() {
        while (atoms.size() % 32) {
            atoms.push_back(make_int4(0, 0, 0, 0));
            meta.push_back(make_int4(-1, 0, 0, 0));
        }
    };
    // blocks are dealt to the warps round-robin, so listing the kinds from the most to the least expensive balances them
    const int go_bond = 0, go_angle = 2 * g.n_bonds, go_tors = go_angle + 2 * g.n_angles, go_pair = go_tors + 2 * g.n_torsions;
    for (int t = 0; t < g.n_torsions; ++t) {
        const int *a = h->torsion_idx.data() + 4 * t;
        atoms.push_back(make_int4(a[0], a[1], a[2], a[3]));
        meta.push_back(make_int4(PMG_TORSION, g.doff_tors + 4 * t, h->torsion_per[t], go_tors + 2 * t));
    }
    pad();
    for (int t = 0; t < g.n_angles; ++t) {
        const int *a = h->angle_idx.data() + 3 * t;
        atoms.push_back(make_int4(a[0], a[1], a[2], 0));
        meta.push_back(make_int4(PMG_ANGLE, g.doff_angle + 2 * t, 0, go_angle + 2 * t));
    }
    pad();
    for (int t = 0; t < g.n_bonds; ++t) {
        const int *a = h->bond_idx.data() + 2 * t;
        atoms.push_back(make_int4(a[0], a[1], 0, 0));
        meta.push_back(make_int4(PMG_BOND, g.doff_bond + 2 * t, 0, go_bond + 2 * t));
    }
    pad();
    for (int q = 0; q < g.n_pairs; ++q) {
        const int *pm = h->host.pair_map.data() + 3 * q;
        atoms.push_back(make_int4(pm[0], pm[1], 0, 0));
        meta.push_back(make_int4(PMG_PAIR, g.doff_pair + g.pair_stride * q, 0, go_pair + 3 * q));
    }
    pad();
    PmGradState *st = new PmGradState();
    st->n_blocks = (int)(atoms.size() / 32);
    st->n_gout = go_pair + 3 * g.n_pairs;
    st->d_atoms = nullptr;
    st->d_meta = nullptr;
    // Work items.  A block costs (terms' relative cost) x (conformations of the tile); small molecules have only a handful of
    // blocks of very different cost (aniline: 2 torsion, 1 angle, 1 bond, 2 pair blocks), so blocks are cut into slices of the
    // tile's conformations until no item exceeds the per-warp share, and the items are dealt to the warps longest first.
    const int warps = std::min(PMG_THREADS / 32, std::max(st->n_blocks, 4));
    st->threads = 32 * warps;
    {
        static const double kind_cost[4] = {1.0, 2.0, 3.4, 1.15};  // bond, angle, torsion, pair (instructions per term, relative)
        std::vector<double> bcost(st->n_blocks);
        double total = 0.0;
        for (int b = 0; b < st->n_blocks; ++b) {
            bcost[b] = kind_cost[meta[(size_t)b * 32].x];
            total += bcost[b];
        }
        const double share = total / warps;
        struct Item { int blk, slice, n_slices; double cost; };
        std::vector<Item> items;
        for (int b = 0; b < st->n_blocks; ++b) {
            int ns = st->n_blocks >= 4 * warps ? 1 : (int)ceil(bcost[b] / share - 1e-9);
            ns = std::max(1, std::min(ns, 8));
            for (int j = 0; j < ns; ++j) items.push_back({b, j, ns, bcost[b] / ns});
        }
        std::stable_sort(items.begin(), items.end(), [](const Item &x, const Item &y) { return x.cost > y.cost; });
        std::vector<std::vector<Item>> per_warp(warps);
        std::vector<double> load(warps, 0.0);
        for (const Item &it : items) {
            int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            per_warp[w].push_back(it);
            load[w] += it.cost;
        }
        std::vector<int4> flat;
        std::vector<int> wptr(1, 0);
        for (int w = 0; w < warps; ++w) {
            for (const Item &it : per_warp[w]) flat.push_back(make_int4(it.blk, it.slice, it.n_slices, 0));
            wptr.push_back((int)flat.size());
        }
        st->n_items = (int)flat.size();
        st->d_items = nullptr;
        st->d_warp_ptr = nullptr;
        if (cudaMalloc(&st->d_items, sizeof(int4) * std::max<size_t>(flat.size(), 1)) != cudaSuccess ||
            cudaMalloc(&st->d_warp_ptr, sizeof(int) * wptr.size()) != cudaSuccess) {
            pm_grad_free(st);
            return nullptr;
        }
        cudaMemcpy(st->d_items, flat.data(), sizeof(int4) * flat.size(), cudaMemcpyHostToDevice);
        cudaMemcpy(st->d_warp_ptr, wptr.data(), sizeof(int) * wptr.size(), cudaMemcpyHostToDevice);
    }
    const int ctas_per_sm = std::max(2, std::min(6, (16 + warps - 1) / warps));  // aim at >= 16 resident warps per SM
    const size_t smem_cap = std::min<size_t>((size_t)200 * 1024 / ctas_per_sm, PMG_MAX_SMEM);
    st->cpt = 128;  // long tiles amortise the per-tile staging, barrier and partial-sum update of small molecules
    auto smem_for = [&](int cpt) { return sizeof(double) * ((size_t)cpt * g.n_atoms * 6 + cpt); };
    while (st->cpt > 1 && smem_for(st->cpt) > smem_cap) st->cpt >>= 1;
    st->smem = smem_for(st->cpt);
    if (st->smem > (size_t)h->max_smem) {
        delete st;
        return nullptr;
    }
    const size_t bytes = sizeof(int4) * std::max<size_t>(atoms.size(), 1);
    if (cudaMalloc(&st->d_atoms, bytes) != cudaSuccess || cudaMalloc(&st->d_meta, bytes) != cudaSuccess) {
        pm_grad_free(st);
        return nullptr;
    }
    cudaMemcpy(st->d_atoms, atoms.data(), sizeof(int4) * atoms.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(st->d_meta, meta.data(), sizeof(int4) * meta.size(), cudaMemcpyHostToDevice);
    const char *env = getenv("PM_GRAD_MINB");
    st->minb = env && *env ? atoi(env) : 2;
    st->fn = st->minb >= 3 ? pm_grad_kernel<3> : pm_grad_kernel<2>;
    cudaFuncSetAttribute(st->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem);
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, st->fn, st->threads, st->smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    st->grid_max = h->n_sm * per_sm;
    h->grad_state = st;
    h->grad_free = pm_grad_free;
    return st;
}

extern "C" long pm_grad_out_doubles(pm_handle h) {
    if (!h) return -1;
    const PmProg &g = h->prog;
    return 2L * g.n_bonds + 2L * g.n_angles + 2L * g.n_torsions + 3L * g.n_pairs;
}

extern "C" long pm_grad_scratch_doubles(pm_handle h) {
    if (!h) return -1;
    cudaSetDevice(h->device);
    PmGradState *st = pm_grad_state(h);
    if (!st) return -1;
    return std::max((long)st->grid_max * st->n_items * 128, pm_spec_grad_scratch(h));
}

extern "C" int pm_grad(pm_handle h, const double *derived_dev, const double *coord_tiles_dev, const double *fmm_tiles_dev,
                       const double *fref_tiles_dev, long n_conf, const double *coef_e_dev, const double *coef_f_dev,
                       double *gout_dev, double *scratch_dev, void *stream) {
    if (!h || !derived_dev || !coord_tiles_dev || !gout_dev || !scratch_dev) return pm_fail_msg("pm_grad: null argument");
    if (n_conf < 1) return pm_fail_msg("pm_grad: empty ensemble");
    if (!coef_e_dev && !coef_f_dev) return pm_fail_msg("pm_grad: neither energy nor force coefficients given");
    if (coef_f_dev && (!fmm_tiles_dev || !fref_tiles_dev)) return pm_fail_msg("pm_grad: force coefficients need the MM and reference force tiles");
    if (h->rf) return pm_fail_msg("pm_grad: the cutoff / reaction-field nonbonded variant has no analytic gradient in this kernel generation");
    PMG_CUDA(cudaSetDevice(h->device));
    if (pm_spec_can_grad(h))  // small molecule with a generated gradient kernel (csrc/specialize.py: GradKernelSource)
        return pm_spec_grad(h, derived_dev, coord_tiles_dev, fmm_tiles_dev, fref_tiles_dev, n_conf, coef_e_dev, coef_f_dev, gout_dev,
                            scratch_dev, (cudaStream_t)stream);
    PmGradState *st = pm_grad_state(h);
    if (!st) return pm_fail_msg("pm_grad: molecule too large for the gradient kernel's shared-memory tile");
    cudaStream_t s = (cudaStream_t)stream;
    PmGradArgs a;
    a.n_atoms = h->prog.n_atoms;
    a.cpw = h->prog.cpw;
    a.cpt = st->cpt;
    a.n_blocks = st->n_blocks;
    a.n_items = st->n_items;
    a.items = st->d_items;
    a.warp_ptr = st->d_warp_ptr;
    a.n_conf = n_conf;
    a.n_tiles = (n_conf + st->cpt - 1) / st->cpt;
    a.atoms = st->d_atoms;
    a.meta = st->d_meta;
    a.derived = derived_dev;
    a.x_tiles = coord_tiles_dev;
    a.fmm_tiles = fmm_tiles_dev;
    a.fref_tiles = fref_tiles_dev;
    a.metric = h->d_metric;
    a.acos_tab = h->d_acos;
    a.coef_e = coef_e_dev;
    a.coef_f = coef_f_dev;
    a.partial = scratch_dev;
    const int grid = (int)std::min<long>(st->grid_max, a.n_tiles);
    PMG_CUDA(cudaMemsetAsync(scratch_dev, 0, sizeof(double) * (size_t)grid * st->n_items * 128, s));
    if (st->n_blocks > 0) {
        st->fn<<<grid, st->threads, st->smem, s>>>(a);
        PMG_CUDA(cudaGetLastError());
        pm_grad_reduce_kernel<<<(st->n_blocks * 32 + 255) / 256, 256, 0, s>>>(scratch_dev, grid, st->n_blocks, st->n_items, st->d_items, st->d_meta,
                                                                              gout_dev);
        PMG_CUDA(cudaGetLastError());
    }
    return 0;
}
