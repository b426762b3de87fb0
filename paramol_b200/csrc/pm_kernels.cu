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
// Mapping: every conformation gets a TEAM of L = 32/CPW lanes; a warp owns a tile of CPW conformations
// whose coordinates [3N][CPW] arrive in shared memory by one TMA bulk copy.  The topology "program" (pm_program.hpp) and
// the derived parameter table are staged once per CTA.  All lanes run the same step loops; in a step the L lanes of a
// team execute records that touch disjoint atoms, so forces accumulate in the team's column of a shared [3N][CPW] tile
// without atomics.  With L = 1 (small molecules) every index/parameter read is a warp-uniform broadcast.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/paramol_b200.h"
#include "pm_device.cuh"
#include "pm_program.hpp"
#include "pm_internal.hpp"

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
        double *o = d + g.doff_pair + g.pair_stride * q;
        o[0] = 4.0 * eps * s6 * s6;  // c12
        o[1] = 4.0 * eps * s6;       // c6
        o[2] = g.k_coul * qq;        // Coulomb product
        if (g.pair_stride == 4) o[3] = e < 0 ? 1.0 : 0.0;  // 1: plain pair (cutoff + reaction field apply), 0: exception
    }
}

// =============================================================================================
// K1: energies (+ forces + force residual) of every conformation, for every parameter vector
// =============================================================================================
// CutoffNonPeriodic with reaction field (OpenMM ReferenceLJCoulombIxn with cutoff, the system of the reference's
// Tests/aniline.xml): plain pairs beyond the cutoff vanish, inside it the Coulomb term is K qq (1/r + k_rf r^2 - c_rf);
// exceptions keep the plain 1/r form.  pp = (c12, c6, K qq, plain flag).
template <bool WITH_F>
__device__ __forceinline__ void pm_pair_rf(const pm_d3 xi, const pm_d3 xj, const double *__restrict__ pp, const double krf,
                                           const double crf, const double cut2, double &e_nb, pm_d3 &fi, pm_d3 &fj) {
    const pm_d3 d = pm_sub(xi, xj);
    const double r2 = pm_dot(d, d);
    const double plain = pp[3];
    if (plain != 0.0 && r2 >= cut2) return;
    const double rinv = pm_rsqrt(r2);
    const double u = rinv * rinv;
    const double u3 = u * u * u;
    const double c12 = pp[0], c6 = pp[1], cq = pp[2];
    const double a = c12 * u3;
    const double b = a - c6;
    const double kr2 = plain * krf * r2;
    e_nb = fma(b, u3, e_nb);
    e_nb = fma(cq, rinv + kr2 - plain * crf, e_nb);
    if (WITH_F) {
        const double g = fma(6.0 * u3, a + b, cq * (rinv - 2.0 * kr2)) * u;
        pm_axpy(fi, g, d);
        pm_axpy(fj, g, d);
    }
}

template <bool WITH_F>
__device__ __forceinline__ void pm_pair(const pm_d3 xi, const pm_d3 xj, const double *__restrict__ pp, double &e_nb, pm_d3 &fi,
                                        pm_d3 &fj) {
    const pm_d3 d = pm_sub(xi, xj);
    const double r2 = pm_dot(d, d);
    const double rinv = pm_rsqrt(r2);
    const double u = rinv * rinv;
    const double u3 = u * u * u;
    const double c12 = pp[0], c6 = pp[1], cq = pp[2];
    const double b = fma(c12, u3, -c6);  // c12 u^3 - c6
    const double ec = cq * rinv;
    e_nb = fma(b, u3, e_nb);
    e_nb += ec;
    if (WITH_F) {
        // -(dE/dr)/r = (12 c12 u^6 - 6 c6 u^3 + cq rinv) u = (6 (2 c12 u^3 - c6) u3 + ec) u, with 2 c12 u^3 - c6 = c12 u^3 + b
        const double g = fma(6.0 * u3, fma(c12, u3, b), ec) * u;
        pm_axpy(fi, g, d);
        pm_axpy(fj, g, d);
    }
}

#ifndef PM_BLK_UNROLL
#define PM_BLK_UNROLL 1  // rows of a pair block per loop iteration (1: rolled, 4: fully unrolled)
#endif
constexpr int kBlkUnroll = PM_BLK_UNROLL;

// The pair of a 4 x 4 block: identical to pm_pair except that |d|^2 is kept above 2^-100 by an integer max on its high word (one
// ALU instruction, not an FP64 one).  Block slots that do not interact carry zero parameters and may pair an atom with itself
// (diagonal blocks, padding atoms): with the guard such a slot evaluates to exactly zero energy and force instead of 0 * inf.
template <bool WITH_F>
__device__ __forceinline__ void pm_pair_blk(const pm_d3 xi, const pm_d3 xj, const double *__restrict__ pp, double &e_nb, pm_d3 &fi,
                                            pm_d3 &fj) {
    const pm_d3 d = pm_sub(xi, xj);
    double r2 = pm_dot(d, d);
    r2 = __hiloint2double(max(__double2hiint(r2), 0x39B00000), __double2loint(r2));
    const double rinv = pm_rsqrt(r2);
    const double u = rinv * rinv;
    const double u3 = u * u * u;
    const double c12 = pp[0], c6 = pp[1], cq = pp[2];
    const double b = fma(c12, u3, -c6);
    const double ec = cq * rinv;
    e_nb = fma(b, u3, e_nb);
    e_nb += ec;
    if (WITH_F) {
        const double g = fma(6.0 * u3, fma(c12, u3, b), ec) * u;
        pm_axpy(fi, g, d);
        pm_axpy(fj, g, d);
    }
}

__device__ __forceinline__ double pm_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// All NL cells of one i-substituent of a "simple" axis (one torsion of periodicity 1..4 per cell) as ONE branch-free block:
// NL independent dependency chains.  Returns A = sum_l a_il; B[q] and e_tors are accumulated in place.
template <int NL, bool WITH_F>
__device__ __forceinline__ double pm_axis_cells(const int *__restrict__ cells, const int2 *__restrict__ s_axis_terms,
                                                const double2 *__restrict__ s_ptors, const pm_d3 c1, const pm_d3 d0, const double s1,
                                                const double rb, const pm_d3 (&c2)[3], const double (&s2)[3], double (&B)[3],
                                                double &e_tors) {
    double A = 0.0;
#pragma unroll
    for (int q = 0; q < NL; ++q) {
        const int2 term = s_axis_terms[cells[q]];  // (torsion index, periodicity)
        const double2 kk = s_ptors[2 * term.x], cs = s_ptors[2 * term.x + 1];
        const double inv12 = s1 * s2[q];
        const double cphi = pm_dot(c1, c2[q]) * inv12;
        const double sphi = rb * pm_dot(d0, c2[q]) * inv12;
        double cn = cphi, sn = sphi;
#pragma unroll
        for (int mm = 1; mm < 4; ++mm) {
            const double c_new = fma(cn, cphi, -(sn * sphi));
            const double s_new = fma(sn, cphi, cn * sphi);
            const bool step = mm < term.y;
            cn = step ? c_new : cn;
            sn = step ? s_new : sn;
        }
        const double cda = fma(cn, cs.x, sn * cs.y);
        e_tors = fma(kk.x, cda, e_tors + kk.x);
        if (WITH_F) {
            const double w = -kk.y * fma(sn, cs.x, -(cn * cs.y)) * rb;
            A = fma(-w, s1 * s1, A);
            B[q] = fma(w, s2[q] * s2[q], B[q]);
        }
    }
    return A;
}

// Force write-back of one step.  One lane per conformation (L == 1): plain.  Teams of lanes: the records of a step that share a
// colour touch disjoint atoms (pm_program.hpp), colours are serialised with __syncwarp between them.  Teams of WARPS (WT): the
// same colours, one ROUND of the team's split barrier per colour (wait until the updates of the previous round are complete,
// update if the record has this round's colour, arrive) -- the records of a step are all computed concurrently and only their
// short write-backs are ordered.
#define PM_ORDERED(n_col, colour, cond, ...)           \
    if (L == 1) {                                      \
        if (cond) __VA_ARGS__                          \
    } else if (WT) {                                   \
        for (int _c = 0; _c < (n_col); ++_c) {         \
            round_wait();                              \
            if ((cond) && (colour) == _c) __VA_ARGS__  \
            round_arrive();                            \
        }                                              \
    } else {                                           \
        for (int _c = 0; _c < (n_col); ++_c) {         \
            if ((cond) && (colour) == _c) __VA_ARGS__  \
            __syncwarp();                              \
        }                                              \
    }

// MAXW = warps per CTA the variant is compiled for: 16 (128 registers per thread), 14 (144) or 12 (168 registers, no spills).
// Measured on B200: when shared memory admits at most 12 warps anyway, the 168-register variant is 5-10 % faster.
// TW > 1 ("teams of warps", CPW = 32): a tile of 32 conformations is shared by TW WARPS; lane = conformation, so every index and
// parameter read is warp-uniform and no branch diverges (as with one lane per conformation), while the shared-memory state per
// conformation in flight stays 48 N bytes (as with teams of lanes).  The TW warps of a team take the records of a step like the
// lanes of a lane-team do, with a named barrier where those use __syncwarp.
template <int CPW, bool WITH_F, int MAXW, bool RF = false, int TW = 1>
__global__ void __launch_bounds__(MAXW * 32, 1)
    pm_ef_kernel(PmProg g, const double *__restrict__ derived, int n_vec, const double *__restrict__ xt,
                 const double *__restrict__ fref_t, long n_conf, long n_tiles, double *__restrict__ emm,
                 double *__restrict__ fres, double *__restrict__ fmm_t, PmRed red) {
    constexpr bool WT = TW > 1;
    static_assert(!WT || CPW == 32, "teams of warps work on tiles of 32 conformations");
    constexpr int L = WT ? TW : 32 / CPW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int col = WT ? lane : lane % CPW;       // conformation inside the tile
    const int tl = WT ? warp % TW : lane / CPW;   // member of the team
    const int team = WT ? warp / TW : warp, n_teams = WT ? n_warps / TW : n_warps;
    auto team_sync = [&]() {
        if (WT)
            asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(32 * TW) : "memory");
        else
            __syncwarp();
    };
    const int N = g.n_atoms;
    const int tile_d = 3 * N * CPW;

    // ---- shared memory carve-up -----------------------------------------------------------------
    int *s_int = reinterpret_cast<int *>(smem_raw);
    size_t off = ((size_t)g.n_int * sizeof(int) + 15) & ~(size_t)15;
    double2 *s_acos = reinterpret_cast<double2 *>(smem_raw + off);
    off += (size_t)(PM_ACOS_N + 1) * sizeof(double2);
    double *s_metric = reinterpret_cast<double *>(smem_raw + off);
    off += ((size_t)9 * N * sizeof(double) + 15) & ~(size_t)15;  // keep the derived table 16-byte aligned (double2 reads)
    double *s_der = reinterpret_cast<double *>(smem_raw + off);
    off += (size_t)(g.n_derived + 4) * sizeof(double);  // + the zero pair slot of the 4 x 4 blocks
    off = (off + 127) & ~(size_t)127;
    // a tile is followed by the three rows of the padding atom (pair blocks of teams; zero coordinates, forces never read)
    const int tile_p = tile_d + 3 * CPW;
    const size_t per_warp = 128 + (size_t)tile_p * sizeof(double) * (WITH_F ? 2 : 1) + (WT ? (size_t)TW * 64 * sizeof(double) : 0);
    unsigned char *wbase = smem_raw + off + per_warp * team;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase);
    double *racc = reinterpret_cast<double *>(wbase + 64);  // fused residual reduction: this team's 5 running sums
    double *xs = reinterpret_cast<double *>(wbase + 128);
    double *fs = xs + tile_p;  // only touched when WITH_F
    double *tsc = xs + tile_p * (WITH_F ? 2 : 1);  // teams of warps: [TW][32][2] partial (energy, residual) of the members
    const bool want_res = WITH_F && fref_t != nullptr;

    for (int i = threadIdx.x; i < g.n_int; i += blockDim.x) s_int[i] = g.ints[i];
    for (int i = threadIdx.x; i <= PM_ACOS_N; i += blockDim.x) s_acos[i] = g.acos_tab[i];
    for (int i = threadIdx.x; i < 9 * N; i += blockDim.x) s_metric[i] = g.metric[i];
    uint64_t *tbar = bar + 1;  // teams of warps: "the force updates of a step are complete" (one arrival per member)
    if (lane == 0 && (!WT || tl == 0)) {
        pm_mbar_init(bar, 1);
        if (WT) pm_mbar_init(tbar, TW);
        pm_fence_mbar_init();
    }
    // Teams of warps order the force updates of consecutive steps with a SPLIT barrier: a member computes its record of step r
    // (coordinates only), waits until every member has finished the updates of step r - 1, updates, and arrives for step r -- so
    // the wait overlaps with the member's own arithmetic instead of ending every step (energy-only kernels need no ordering).
    uint32_t n_arr = 0;
    int tile_round = 0;
    auto round_wait = [&]() {
        if (WT && WITH_F && tile_round > 0) pm_mbar_wait(tbar, (n_arr - 1) & 1u);
    };
    auto round_arrive = [&]() {
        if (WT && WITH_F) {
            __syncwarp();
            if (lane == 0) pm_mbar_arrive(tbar);
            ++n_arr;
            ++tile_round;
        }
    };
    if (!WT || tl == 0)
        for (int k = lane; k < 3 * CPW; k += 32) {
            xs[tile_d + k] = 0.0;
            if (WITH_F) fs[tile_d + k] = 0.0;
        }
    uint32_t parity = 0;

    const int2 *s_star_tab = reinterpret_cast<const int2 *>(s_int + g.off_star_tab);
    const int *s_star_hdr = s_int + g.off_star_hdr;
    const int4 *s_star_rec = reinterpret_cast<const int4 *>(s_int + g.off_star_rec);
    const int2 *s_axis_tab = reinterpret_cast<const int2 *>(s_int + g.off_axis_tab);
    const int *s_axis_hdr = s_int + g.off_axis_hdr;
    const int *s_axis_rec = s_int + g.off_axis_rec;
    const int2 *s_axis_terms = reinterpret_cast<const int2 *>(s_int + g.off_axis_terms);
    const int2 *s_seg_hdr = reinterpret_cast<const int2 *>(s_int + g.off_seg_hdr);
    const int4 *s_seg_tile = reinterpret_cast<const int4 *>(s_int + g.off_seg_tile);
    const int2 *s_entries = reinterpret_cast<const int2 *>(s_int + g.off_entries);

    const long warp_stride = (long)gridDim.x * n_teams;
    const long first_tile = (long)blockIdx.x * n_teams + team;

    for (int p = 0; p < n_vec; ++p) {
        __syncthreads();  // everybody done with the previous derived table (and the int blob is staged)
        {
            const double *dsrc = derived + (size_t)p * g.n_derived;
            for (int i = threadIdx.x; i < g.n_derived; i += blockDim.x) s_der[i] = dsrc[i];
            // zero pair slot (index n_pairs; may overlap the alignment pad of the table, hence written after the copy by the
            // thread that copied that entry or by nobody else)
            __syncthreads();
            if (threadIdx.x < 4) s_der[g.doff_pair + g.pair_stride * g.n_pairs + threadIdx.x] = 0.0;
        }
        __syncthreads();
        const double2 *s_pbond = reinterpret_cast<const double2 *>(s_der + g.doff_bond);
        const double2 *s_pangle = reinterpret_cast<const double2 *>(s_der + g.doff_angle);
        const double2 *s_ptors = reinterpret_cast<const double2 *>(s_der + g.doff_tors);
        const double *s_ppair = s_der + g.doff_pair;
        if (red.rows != nullptr && lane < 5 && (!WT || tl == 0)) racc[lane] = 0.0;

        for (long tile = first_tile; tile < n_tiles; tile += warp_stride) {
            // ---- coordinates of CPW conformations: one TMA bulk copy into this warp's tile ---------
            const bool nb_cached = g.nb_e != nullptr;
            if (lane == 0 && (!WT || tl == 0)) {
                pm_fence_proxy_async();
                const uint32_t tile_bytes = (uint32_t)(tile_d * sizeof(double));
                pm_mbar_expect_tx(bar, (WITH_F && nb_cached) ? 2 * tile_bytes : tile_bytes);
                pm_bulk_g2s(xs, xt + (size_t)tile * tile_d, tile_bytes, bar);
                // cached nonbonded forces seed the force tile (instead of zeros) when the pair phase is skipped
                if (WITH_F && nb_cached) pm_bulk_g2s(fs, g.nb_f_tiles + (size_t)tile * tile_d, tile_bytes, bar);
                // hide HBM latency of what comes next: this tile's reference forces (read in the epilogue) and the
                // coordinates of the next tile of this warp
                if (want_res)
                    pm_bulk_prefetch_l2(fref_t + (size_t)tile * tile_d, (uint32_t)(tile_d * sizeof(double)));
                if (tile + warp_stride < n_tiles)
                    pm_bulk_prefetch_l2(xt + (size_t)(tile + warp_stride) * tile_d, (uint32_t)(tile_d * sizeof(double)));
            }
            if (WITH_F && !nb_cached) {
                if (WT)
                    for (int k = tl * 32 + lane; k < tile_d; k += 32 * L) fs[k] = 0.0;
                else
                    for (int k = lane; k < tile_d; k += 32) fs[k] = 0.0;
            }
            pm_mbar_wait(bar, parity);
            parity ^= 1;
            team_sync();
            tile_round = 0;

            double e_bond = 0.0, e_angle = 0.0, e_tors = 0.0, e_nb = 0.0;
            if (nb_cached && tl == 0 && tile * CPW + col < n_conf) e_nb = g.nb_e[tile * CPW + col];
            const double *xc_base = xs + col;  // this lane's column of the coordinate / force tiles
            double *fc_base = fs + col;

            // ================= stars: bonds and angles around a centre atom ==========================
            for (int step = 0; step < g.n_star_steps; ++step) {
                const int2 te = s_star_tab[step * L + tl];  // (record, colour)
                const int rec = te.x;
                const int n_col = L > 1 ? s_star_hdr[step] : 1;
                int center = 0, m = 0;
                int nb[4] = {0, 0, 0, 0};
                pm_d3 f[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) f[k] = pm_d3{0.0, 0.0, 0.0};
                if (L == 1 || rec >= 0) {
                    const int4 h0 = s_star_rec[4 * rec], h1 = s_star_rec[4 * rec + 1], h2 = s_star_rec[4 * rec + 2],
                               h3 = s_star_rec[4 * rec + 3];
                    center = h0.x;
                    m = h0.y;
                    nb[0] = h0.z;
                    nb[1] = h0.w;
                    nb[2] = h1.x;
                    nb[3] = h1.y;
                    const int bp[4] = {h1.z, h1.w, h2.x, h2.y};
                    const int ap[6] = {h2.z, h2.w, h3.x, h3.y, h3.z, h3.w};
                    const pm_d3 xc = pm_ldo<CPW>(xc_base, center);
                    pm_d3 d[4];
                    double r2[4], ri[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        d[k] = pm_d3{1.0, 0.0, 0.0};
                        r2[k] = 1.0;
                        ri[k] = 1.0;
                        if (k < m) {
                            d[k] = pm_sub(xc, pm_ldo<CPW>(xc_base, nb[k]));
                            r2[k] = pm_dot(d[k], d[k]);
                            ri[k] = pm_rsqrt(r2[k]);
                            if (bp[k] >= 0) {
                                const double2 pb = s_pbond[bp[k]];  // (r0, k)
                                const double di = fma(r2[k], ri[k], -pb.x);
                                const double kd = pb.y * di;
                                e_bond = fma(0.5 * kd, di, e_bond);
                                if (WITH_F) pm_axpy(f[k], kd * ri[k], d[k]);
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        constexpr int PA[6] = {0, 0, 0, 1, 1, 2}, PB[6] = {1, 2, 3, 2, 3, 3};
                        if (ap[q] >= 0) {
                            const int a = PA[q], b = PB[q];
                            const double2 pa = s_pangle[ap[q]];  // (theta0, k)
                            const double dot = pm_dot(d[a], d[b]);
                            const double rr = ri[a] * ri[b];
                            double cosine = dot * rr;
                            cosine = fmin(1.0, fmax(-1.0, cosine));
                            // |d_a x d_b|^2 (Lagrange identity)
                            const double rp2 = fmax(fma(-dot, dot, r2[a] * r2[b]), 0.0);
                            const bool tiny = rp2 <= 1.0e-12;  // the reference clamps |p| at 1e-6
                            const double rpinv = tiny ? 1.0e6 : pm_rsqrt(rp2);
                            const double sine = tiny ? sqrt(rp2) * rr : rp2 * rpinv * rr;
                            const double theta = pm_acos_cs(cosine, sine, s_acos);
                            const double di = theta - pa.x;
                            const double kd = pa.y * di;
                            e_angle = fma(0.5 * kd, di, e_angle);
                            if (WITH_F) {
                                // f_a = termA d_a x p, f_b = termC d_b x p with p = d_a x d_b, expanded with the
                                // triple-product rule:  f_a = alpha d_a - w d_b ;  f_b = beta d_b - w d_a
                                const double w = kd * rpinv;
                                const double wd = w * dot;
                                const double alpha = wd * (ri[a] * ri[a]), beta = wd * (ri[b] * ri[b]);
                                pm_axpy(f[a], alpha, d[a]);
                                pm_axpy(f[a], -w, d[b]);
                                pm_axpy(f[b], beta, d[b]);
                                pm_axpy(f[b], -w, d[a]);
                            }
                        }
                    }
                }
                if (WITH_F) {
                    // write-back: records of one colour touch disjoint atoms; colours are serialised
                    PM_ORDERED(n_col, te.y, rec >= 0, {
                        pm_d3 fc{0.0, 0.0, 0.0};
                        _Pragma("unroll") for (int k = 0; k < 4; ++k) {
                            if (k < m) {
                                pm_addo<CPW>(fc_base, nb[k], f[k]);
                                fc.x -= f[k].x;
                                fc.y -= f[k].y;
                                fc.z -= f[k].z;
                            }
                        }
                        pm_addo<CPW>(fc_base, center, fc);
                    })
                }
            }

            // ================= axes: every torsion around a rotation axis j-k (trig-free) ================
            for (int step = 0; step < g.n_axis_steps; ++step) {
                const int2 te = s_axis_tab[step * L + tl];  // (record, colour)
                const int rec = te.x;
                const bool live = L == 1 || rec >= 0;
                const int hdr = L > 1 ? s_axis_hdr[step] : 1;
                const int n_col = hdr & 0xff;
                int aj = 0, ak = 0, nI = 0, nL = 0;
                bool simple = false;
                int La[3] = {0, 0, 0};
                const int *arec = s_axis_rec + 20 * (rec < 0 ? 0 : rec);  // [4..6] i atoms, [10..19] cell starts
                pm_d3 xj{0.0, 0.0, 0.0}, d1{1.0, 0.0, 0.0};
                double rb = 1.0, ib2 = 1.0;
                pm_d3 c2[3];
                double s2[3], g2[3], B[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    c2[q] = pm_d3{1.0, 0.0, 0.0};
                    s2[q] = 1.0;
                    g2[q] = 0.0;
                    B[q] = 0.0;
                }
                if (live) {
                    const int4 *r4 = reinterpret_cast<const int4 *>(s_axis_rec + 20 * rec);
                    const int4 h0 = r4[0], h1 = r4[1], h2 = r4[2];
                    aj = h0.x;
                    ak = h0.y;
                    nI = h0.z & 255;
                    simple = (h0.z & 256) != 0;
                    nL = h0.w;
                    La[0] = h1.w;
                    La[1] = h2.x;
                    La[2] = h2.y;
                    xj = pm_ldo<CPW>(xc_base, aj);
                    const pm_d3 xk = pm_ldo<CPW>(xc_base, ak);
                    d1 = pm_sub(xk, xj);
                    const double nb2 = pm_dot(d1, d1);
                    const double rbi = pm_rsqrt(nb2);
                    rb = nb2 * rbi;
                    ib2 = rbi * rbi;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        if (q < nL) {
                            const pm_d3 d2 = pm_sub(xk, pm_ldo<CPW>(xc_base, La[q]));
                            c2[q] = pm_cross(d1, d2);
                            s2[q] = pm_rsqrt(pm_dot(c2[q], c2[q]));
                            g2[q] = pm_dot(d2, d1) * ib2;
                        }
                    }
                }
                pm_d3 Fj{0.0, 0.0, 0.0}, Fk{0.0, 0.0, 0.0};
                double Ai0 = 0.0, Ai1 = 0.0, Ai2 = 0.0;  // teams of warps: the scalars of F_i = A_i c1_i (re-formed at the write-back)
                const int ni_loop = (L > 1 && !WT) ? (hdr >> 8) : nI;   // (teams of lanes: the longest record of the step)
                for (int pi = 0; pi < ni_loop; ++pi) {
                    const bool act = L == 1 || WT || (live && pi < nI);
                    int ai = 0;
                    pm_d3 fi{0.0, 0.0, 0.0};
                    if (act) {
                        ai = arec[4 + pi];
                        const int *cells = arec + 10 + 3 * pi;
                        const pm_d3 d0 = pm_sub(pm_ldo<CPW>(xc_base, ai), xj);
                        const pm_d3 c1 = pm_cross(d0, d1);
                        const double s1 = pm_rsqrt(pm_dot(c1, c1));
                        double A = 0.0;
                        if ((L == 1 || WT) && simple) {  // (with teams of lanes the lanes of a step would diverge between the two paths)
                            if (nL == 2)
                                A = pm_axis_cells<2, WITH_F>(cells, s_axis_terms, s_ptors, c1, d0, s1, rb, c2, s2, B, e_tors);
                            else if (nL == 3)
                                A = pm_axis_cells<3, WITH_F>(cells, s_axis_terms, s_ptors, c1, d0, s1, rb, c2, s2, B, e_tors);
                            else
                                A = pm_axis_cells<1, WITH_F>(cells, s_axis_terms, s_ptors, c1, d0, s1, rb, c2, s2, B, e_tors);
                        } else
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const int t0 = cells[q], t1 = cells[q + 1];
                            if (t1 > t0) {
                                const double inv12 = s1 * s2[q];
                                const double cphi = pm_dot(c1, c2[q]) * inv12;
                                const double sphi = rb * pm_dot(d0, c2[q]) * inv12;
                                double dE = 0.0;
                                for (int t = t0; t < t1; ++t) {
                                    const int2 term = s_axis_terms[t];  // (torsion index, periodicity)
                                    const double2 kk = s_ptors[2 * term.x], cs = s_ptors[2 * term.x + 1];  // (k, k n), (cos, sin) delta
                                    // (cos, sin)(n phi) by the angle-addition recurrence, started at n = 1
                                    double cn = term.y > 0 ? cphi : 1.0, sn = term.y > 0 ? sphi : 0.0;
                                    for (int mm = 1; mm < term.y; ++mm) {
                                        const double c_new = fma(cn, cphi, -(sn * sphi));
                                        sn = fma(sn, cphi, cn * sphi);
                                        cn = c_new;
                                    }
                                    const double cda = fma(cn, cs.x, sn * cs.y);  // cos(n phi - delta)
                                    e_tors = fma(kk.x, cda, e_tors + kk.x);
                                    if (WITH_F) dE = fma(-kk.y, fma(sn, cs.x, -(cn * cs.y)), dE);  // -k n sin(n phi - delta)
                                }
                                if (WITH_F) {
                                    const double w = dE * rb;
                                    A = fma(-w, s1 * s1, A);
                                    B[q] = fma(w, s2[q] * s2[q], B[q]);
                                }
                            }
                        }
                        if (WITH_F) {
                            const double g1 = pm_dot(d0, d1) * ib2;
                            const double Ag = A * g1;
                            fi = pm_d3{A * c1.x, A * c1.y, A * c1.z};
                            pm_axpy(Fj, Ag - A, c1);  // F_j -= A (1 - g1) c1
                            pm_axpy(Fk, -Ag, c1);     // F_k -= g1 A c1
                            if (WT) {
                                if (pi == 0)
                                    Ai0 = A;
                                else if (pi == 1)
                                    Ai1 = A;
                                else
                                    Ai2 = A;
                            }
                        }
                    }
                    if (WITH_F && !WT) {
                        PM_ORDERED(n_col, te.y, act, { pm_addo<CPW>(fc_base, ai, fi); })
                    }
                }
                if (WITH_F) {
                    if (live) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            if (q < nL) {
                                const double Bg = B[q] * g2[q];
                                pm_axpy(Fj, -Bg, c2[q]);        // F_j -= g2 B c2
                                pm_axpy(Fk, Bg - B[q], c2[q]);  // F_k -= B (1 - g2) c2
                            }
                        }
                    }
                    PM_ORDERED(n_col, te.y, live, {
                        if (WT) {
                            // the i-substituents: F_i = A_i (d0 x d1) with the cross product formed again (9 FP64 operations per
                            // substituent instead of keeping up to three force vectors alive through the write-back rounds)
                            const double Ai[3] = {Ai0, Ai1, Ai2};
                            _Pragma("unroll") for (int q = 0; q < 3; ++q) {
                                if (q < nI) {
                                    const int ai = arec[4 + q];
                                    const pm_d3 c1 = pm_cross(pm_sub(pm_ldo<CPW>(xc_base, ai), xj), d1);
                                    pm_addo<CPW>(fc_base, ai, pm_d3{Ai[q] * c1.x, Ai[q] * c1.y, Ai[q] * c1.z});
                                }
                            }
                        }
                        _Pragma("unroll") for (int q = 0; q < 3; ++q) {
                            if (q < nL) pm_addo<CPW>(fc_base, La[q], pm_d3{B[q] * c2[q].x, B[q] * c2[q].y, B[q] * c2[q].z});
                        }
                        pm_addo<CPW>(fc_base, aj, Fj);
                        pm_addo<CPW>(fc_base, ak, Fk);
                    })
                }
            }

            // ================= nonbonded, teams: 4 x 4 register blocks on disjoint tile pairs ====================
            if (L >= 4 && !RF) {
                for (int pos = 0; pos < g.n_pair_pos; ++pos) {
                    const int *rec = s_int + g.off_blocks + (pos * L + tl) * 16;
                    const int4 ja = *reinterpret_cast<const int4 *>(rec + 4);
                    const int jo[4] = {ja.x, ja.y, ja.z, ja.w};
                    pm_d3 xj[4], aj[4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        xj[m] = pm_ldo<CPW>(xc_base, jo[m]);
                        aj[m] = pm_d3{0.0, 0.0, 0.0};
                    }
                    // rows of the block one at a time (a rolled loop: the body -- four independent pairs -- stays small enough for
                    // the instruction cache; all register indices are static).  Measured and dropped: two rows per iteration (-0.4 %),
                    // a software pipeline that requests the next row's coordinates and forces one row ahead (-0.5 .. -2.5 %)
#pragma unroll kBlkUnroll
                    for (int k = 0; k < 4; ++k) {
                        const int io = rec[k];
                        if (WT && io == tile_d * (int)sizeof(double)) {   // a padding row (tiles of 3 atoms, empty slots): uniform skip
                            if (k == 0) round_wait();
                            continue;
                        }
                        const int2 qq = *reinterpret_cast<const int2 *>(rec + 8 + 2 * k);
                        const unsigned q[4] = {(unsigned)qq.x & 0xffffu, (unsigned)qq.x >> 16, (unsigned)qq.y & 0xffffu, (unsigned)qq.y >> 16};
                        const pm_d3 xi = pm_ldo<CPW>(xc_base, io);
                        pm_d3 ai{0.0, 0.0, 0.0};
#pragma unroll
                        for (int m = 0; m < 4; ++m) pm_pair_blk<WITH_F>(xi, xj[m], s_ppair + 3 * q[m], e_nb, ai, aj[m]);
                        if (WT && k == 0) round_wait();  // the first row's pairs are evaluated while the previous step drains
                        if (WITH_F) pm_addo<CPW>(fc_base, io, ai);
                    }
                    if (WITH_F) {
#pragma unroll
                        for (int m = 0; m < 4; ++m) pm_subo<CPW>(fc_base, jo[m], aj[m]);
                    }
                    if (WT)
                        round_arrive();
                    else
                        __syncwarp();
                }
            }
            // ================= nonbonded: plain pairs and scaled 1-4 exceptions, register-tiled over i ========
            for (int seg = 0; seg < g.n_segs; ++seg) {
                const int2 hdr = s_seg_hdr[seg];
                const int4 ta4 = s_seg_tile[seg * L + tl];
                const int ta[PM_IT] = {ta4.x, ta4.y, ta4.z, ta4.w};
                pm_d3 xi[PM_IT], acc[PM_IT];
#pragma unroll
                for (int k = 0; k < PM_IT; ++k) {
                    xi[k] = pm_ldo<CPW>(xc_base, ta[k] < 0 ? 0 : ta[k]);
                    acc[k] = pm_d3{0.0, 0.0, 0.0};
                }
                const int2 *ent = s_entries + hdr.y + tl;
                for (int e = 0; e < hdr.x; ++e) {
                    const int2 en = ent[e * L];
                    const int mask = (en.x >> 28) & 15;
                    if (mask) {
                        const int j = en.x & 0x0fffffff;  // byte offset of atom j
                        const double *pp = s_ppair + (RF ? 4 : 3) * en.y;
                        const pm_d3 xj = pm_ldo<CPW>(xc_base, j);
                        pm_d3 fj{0.0, 0.0, 0.0};
                        if (RF) {
#pragma unroll
                            for (int k = 0; k < PM_IT; ++k) {
                                if (mask & (1 << k)) {
                                    pm_pair_rf<WITH_F>(xi[k], xj, pp, g.rf_krf, g.rf_crf, g.rf_cut2, e_nb, acc[k], fj);
                                    pp += 4;
                                }
                            }
                        } else if (mask == 15) {
                            // all four pairs interact: one branch-free block, four independent dependency chains
#pragma unroll
                            for (int k = 0; k < PM_IT; ++k) pm_pair<WITH_F>(xi[k], xj, pp + 3 * k, e_nb, acc[k], fj);
                        } else {
#pragma unroll
                            for (int k = 0; k < PM_IT; ++k) {
                                if (mask & (1 << k)) {
                                    pm_pair<WITH_F>(xi[k], xj, pp, e_nb, acc[k], fj);
                                    pp += 3;
                                }
                            }
                        }
                        if (WITH_F) pm_subo<CPW>(fc_base, j, fj);
                    }
                    if (L > 1) __syncwarp();
                }
                if (WITH_F) {
#pragma unroll
                    for (int k = 0; k < PM_IT; ++k) {
                        if (ta[k] >= 0) pm_addo<CPW>(fc_base, ta[k], acc[k]);
                    }
                    if (L > 1) __syncwarp();
                }
            }

            // ================= epilogue: per-conformation energy and force residual ==========================
            round_wait();  // teams of warps: the updates of the last step are complete
            const long sconf = tile * CPW + col;
            double e_total = ((e_bond + e_angle) + e_tors) + e_nb;
            double res = 0.0;
            if (want_res) {
                const double *fr = fref_t + (size_t)tile * tile_d + col;
                for (int a = tl; a < N; a += L) {
                    const double dx = fs[(3 * a) * CPW + col] - fr[(3 * a) * CPW];
                    const double dy = fs[(3 * a + 1) * CPW + col] - fr[(3 * a + 1) * CPW];
                    const double dz = fs[(3 * a + 2) * CPW + col] - fr[(3 * a + 2) * CPW];
                    const double *m = s_metric + 9 * a;
                    // (d^T M) d as in np.matmul(np.matmul(d, M), d), force_property.py:93-94
                    const double t0 = fma(dz, m[6], fma(dy, m[3], dx * m[0]));
                    const double t1 = fma(dz, m[7], fma(dy, m[4], dx * m[1]));
                    const double t2 = fma(dz, m[8], fma(dy, m[5], dx * m[2]));
                    res += fma(t2, dz, fma(t1, dy, t0 * dx));
                }
            }
            if (WT) {
                // the members' partial energies and residuals of each conformation: through shared memory, added in member order
                tsc[(tl * 32 + lane) * 2] = e_total;
                tsc[(tl * 32 + lane) * 2 + 1] = res;
                team_sync();
                if (tl == 0) {
                    e_total = 0.0;
                    res = 0.0;
                    for (int m = 0; m < L; ++m) {
                        e_total += tsc[(m * 32 + lane) * 2];
                        res += tsc[(m * 32 + lane) * 2 + 1];
                    }
                }
            } else if (L > 1) {
#pragma unroll
                for (int o = CPW; o < 32; o <<= 1) {
                    e_total += __shfl_xor_sync(0xffffffffu, e_total, o);
                    res += __shfl_xor_sync(0xffffffffu, res, o);
                }
            }
            if (tl == 0 && sconf < n_conf) {
                if (emm != nullptr) emm[(size_t)p * n_conf + sconf] = e_total;
                if (WITH_F && fres != nullptr) fres[(size_t)p * n_conf + sconf] = res;
            }
            if (red.rows != nullptr && (!WT || tl == 0)) {
                // fused pm_reduce_objective: the five weighted sums of this tile join the team's running sums (fixed order:
                // deterministic for a given launch geometry); no [P][S] arrays are written or re-read
                double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0, q4 = 0.0;
                if (tl == 0 && sconf < n_conf) {
                    const double d = (red.eref != nullptr ? red.eref[sconf] : 0.0) - e_total - red.d_shift;
                    const double ws = red.w != nullptr ? red.w[sconf] : 1.0;
                    q0 = d;
                    q1 = ws * d;
                    q2 = q1 * d;
                    q3 = ws;
                    q4 = ws * res;
                }
                q0 = pm_warp_sum(q0);
                q1 = pm_warp_sum(q1);
                q2 = pm_warp_sum(q2);
                q3 = pm_warp_sum(q3);
                q4 = pm_warp_sum(q4);
                if (lane == 0) {
                    racc[0] += q0;
                    racc[1] += q1;
                    racc[2] += q2;
                    racc[3] += q3;
                    racc[4] += q4;
                }
            }
            if (WITH_F && fmm_t != nullptr) {
                double *fo = fmm_t + ((size_t)p * n_tiles + tile) * tile_d;
                if (WT)
                    for (int k = tl * 32 + lane; k < tile_d; k += 32 * L) fo[k] = fs[k];
                else
                    for (int k = lane; k < tile_d; k += 32) fo[k] = fs[k];
            }
            team_sync();  // the tile (and, for teams of warps, the scratch of the members) is free again
        }
        if (red.rows != nullptr && (!WT || tl == 0)) {
            __syncwarp();
            if (lane < 5) red.rows[(((size_t)p * gridDim.x + blockIdx.x) * n_teams + team) * 5 + lane] = racc[lane];
        }
    }
}

// =============================================================================================
// [S][N][3] <-> tiles of CPW conformations
// =============================================================================================
__global__ void pm_pack_kernel(const double *__restrict__ src, long n_conf, int row /* 3N */, int cpw, double *__restrict__ tiles) {
    extern __shared__ double s_buf[];  // [cpw][row + 1]
    const long tile = blockIdx.x;
    const long s0 = tile * cpw;
    const int total = cpw * row;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i / row, k = i - c * row;
        long s = s0 + c;
        if (s >= n_conf) s = n_conf - 1;
        s_buf[c * (row + 1) + k] = src[(size_t)s * row + k];
    }
    __syncthreads();
    double *dst = tiles + (size_t)tile * total;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int k = i / cpw, c = i - k * cpw;
        dst[i] = s_buf[c * (row + 1) + k];
    }
}

__global__ void pm_unpack_kernel(const double *__restrict__ tiles, long n_conf, int row, int cpw, double *__restrict__ dst) {
    extern __shared__ double s_buf[];
    const long tile = blockIdx.x;
    const long s0 = tile * cpw;
    const int total = cpw * row;
    const double *src = tiles + (size_t)tile * total;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int k = i / cpw, c = i - k * cpw;
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
    // unrolled so that the loads of four iterations are in flight together (the sums keep their order)
#pragma unroll 4
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
__global__ void pm_fma_peak_kernel(T *out, int iters, T b0, T c0) {
    // 16 independent chains per thread; the addend stays a kernel parameter (a UNIFORM register operand).  With three
    // per-thread register operands the same loop tops out 8 % lower on B200 (34.2 vs 37.1 TFLOP/s FP64: register-read
    // bandwidth, not the pipe), so this form is the one that measures the pipe peak the roofline is quoted against.
    T a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (T)(threadIdx.x + i) * (T)1e-3;
    const T b = b0 + (T)threadIdx.x * (T)1e-12, c = c0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = a[i] * b + c;
        }
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// =============================================================================================
// Host side: handle, C ABI
// =============================================================================================
static thread_local std::string g_err;
static int pm_fail(const std::string &msg) {
    g_err = msg;
    return 1;
}
int pm_fail_msg(const std::string &msg) { return pm_fail(msg); }  // for the other translation units
#define PM_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) return pm_fail(std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

typedef void (*pm_ef_fn)(PmProg, const double *, int, const double *, const double *, long, long, double *, double *, double *, PmRed);

template <int MAXW>
static pm_ef_fn pm_pick_kernel_w(int cpw, int wf) {
    switch (cpw) {
        case 32: return wf ? pm_ef_kernel<32, true, MAXW> : pm_ef_kernel<32, false, MAXW>;
        case 16: return wf ? pm_ef_kernel<16, true, MAXW> : pm_ef_kernel<16, false, MAXW>;
        case 8: return wf ? pm_ef_kernel<8, true, MAXW> : pm_ef_kernel<8, false, MAXW>;
        case 4: return wf ? pm_ef_kernel<4, true, MAXW> : pm_ef_kernel<4, false, MAXW>;
        default: return nullptr;
    }
}
#define PM_TEAM_WARPS 8  // members of a team of warps (the one size compiled; teams of 16 warps were measured and dropped: the 4 x 4
                         // pair blocks of a 60-atom molecule fill only 120 of 256 slots -- lig60 68.0 vs 83.8 M conformations/s)
static pm_ef_fn pm_pick_kernel(int cpw, int wf, int maxw, int rf = 0, int team_warps = 0) {
    if (team_warps > 1) {  // teams of PM_TEAM_WARPS warps: 8 warps per CTA (168 registers) or 16 (128 registers)
        if (maxw <= 12) return wf ? pm_ef_kernel<32, true, 8, false, PM_TEAM_WARPS> : pm_ef_kernel<32, false, 8, false, PM_TEAM_WARPS>;
        return wf ? pm_ef_kernel<32, true, 16, false, PM_TEAM_WARPS> : pm_ef_kernel<32, false, 16, false, PM_TEAM_WARPS>;
    }
    if (rf) return wf ? pm_ef_kernel<32, true, 12, true> : pm_ef_kernel<32, false, 12, true>;  // CPW 32 only
    return maxw <= 12 ? pm_pick_kernel_w<12>(cpw, wf) : maxw <= 14 ? pm_pick_kernel_w<14>(cpw, wf) : pm_pick_kernel_w<16>(cpw, wf);
}

extern "C" const char *pm_last_error(void) { return g_err.c_str(); }
extern "C" int pm_version(void) { return 200; }
extern "C" int pm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// shared-memory plan for a given CPW: bytes shared by the CTA, per-warp bytes, warps that fit
static void pm_plan(int n_int, int n_derived, int N, int cpw, int max_smem, int wf, size_t *shared, size_t *per_warp, int *warps,
                    int team_warps = 0) {
    size_t off = ((size_t)n_int * sizeof(int) + 15) & ~(size_t)15;
    off += (size_t)(PM_ACOS_N + 1) * sizeof(double2);
    off += ((size_t)9 * N * sizeof(double) + 15) & ~(size_t)15;
    off += (size_t)(n_derived + 4) * sizeof(double);
    off = (off + 127) & ~(size_t)127;
    size_t pw = 128 + (size_t)3 * (N + 1) * cpw * sizeof(double) * (wf ? 2 : 1);  // tile + the rows of the padding atom
    if (team_warps > 1) pw += (size_t)team_warps * 64 * sizeof(double);           // + the members' partial (energy, residual) per conformation
    long w = ((long)max_smem - (long)off) / (long)pw;                             // tiles (= teams, or warps) that fit
    if (team_warps > 1) {
        if (w > PM_MAX_WARPS / team_warps) w = PM_MAX_WARPS / team_warps;
        w *= team_warps;                                                          // warps
    }
    if (w > PM_MAX_WARPS) w = PM_MAX_WARPS;
    if (w < 0) w = 0;
    *shared = off;
    *per_warp = pw;   // per TEAM for teams of warps
    *warps = (int)w;
}

// pair classes (0 plain, 1 excluded, 2 + e active exception e) + index validation; returns 0 or sets the error
static int pm_classify(const pm_topology *t, std::vector<int> &cls) {
    const int N = t->n_atoms;
    cls.assign((size_t)N * N, 0);
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
    return 0;
}

// Host-only: build the topology program for a given CPW and validate its invariants (no CUDA call).
// stats10 as pm_program_info.  Returns 0 when the program is valid.
extern "C" int pm_program_check(const pm_topology *t, int cpw, int *stats10) {
    if (!t) return pm_fail("pm_program_check: null argument");
    int team_warps = 0;
    if (cpw == -PM_TEAM_WARPS) {  // teams of warps on tiles of 32 conformations
        cpw = 32;
        team_warps = PM_TEAM_WARPS;
    }
    if (cpw != 32 && cpw != 16 && cpw != 8 && cpw != 4) return pm_fail("pm_program_check: cpw must be 32, 16, 8, 4 or -8");
    if (t->n_atoms < 1 || t->n_atoms > 65535) return pm_fail("pm_program_check: n_atoms out of range");
    std::vector<int> cls;
    if (pm_classify(t, cls)) return 1;
    PmHostProgram p;
    if (!pm_build_program(t->n_atoms, cpw, t->n_bonds, t->bond_idx, t->n_angles, t->angle_idx, t->n_torsions, t->torsion_idx,
                          t->torsion_per, cls, p, team_warps))
        return pm_fail("pm_program_check: degenerate bonded term (repeated atom)");
    const int rc = pm_validate_program(p, t->n_atoms, t->n_bonds, t->bond_idx, t->n_angles, t->angle_idx, t->n_torsions,
                                       t->torsion_idx, cls);
    if (stats10) {
        const int v[10] = {cpw, p.L, p.n_stars, p.n_star_steps, p.n_axes, p.n_axis_steps, p.n_segs, p.n_entries,
                           p.n_entry_slots, (int)p.blob.size()};
        memcpy(stats10, v, sizeof(v));
    }
    if (rc) return pm_fail("pm_program_check: invariant " + std::to_string(rc) + " violated");
    return 0;
}

// Host-only: the interacting pair list (i, j, exception or -1) in derived-table order for a given CPW (no CUDA call).
extern "C" int pm_topology_pairs(const pm_topology *t, int cpw, int *n_pairs, int *out_pairs) {
    if (!t || !n_pairs) return pm_fail("pm_topology_pairs: null argument");
    if (cpw != 32 && cpw != 16 && cpw != 8 && cpw != 4) return pm_fail("pm_topology_pairs: cpw must be 32, 16, 8 or 4");
    std::vector<int> cls;
    if (pm_classify(t, cls)) return 1;
    PmHostProgram p;
    if (!pm_build_program(t->n_atoms, cpw, t->n_bonds, t->bond_idx, t->n_angles, t->angle_idx, t->n_torsions, t->torsion_idx,
                          t->torsion_per, cls, p))
        return pm_fail("pm_topology_pairs: degenerate bonded term (repeated atom)");
    *n_pairs = p.n_pairs;
    if (out_pairs) memcpy(out_pairs, p.pair_map.data(), sizeof(int) * p.pair_map.size());
    return 0;
}

extern "C" int pm_create(const pm_topology *t, int device, int conf_per_tile, pm_handle *out) {
    if (!t || !out) return pm_fail("pm_create: null argument");
    if (t->n_atoms < 1 || t->n_atoms > 65535) return pm_fail("pm_create: n_atoms out of range");
    if (t->nonbonded_method != 0 && t->nonbonded_method != 1) return pm_fail("pm_create: unknown nonbonded method");
    const int rf = t->nonbonded_method == 1;
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
    std::vector<int> cls;
    if (pm_classify(t, cls)) return 1;

    // ---- choose CPW (conformations per warp; L = 32 / CPW lanes per conformation) ---------------------------
    // The derived-table size does not depend on L, the program size only weakly: plan with the L = 1 program.
    const int max_smem = (int)prop.sharedMemPerBlockOptin;
    int cpw = conf_per_tile, team_warps = 0;
    if (cpw == -PM_TEAM_WARPS) {  // tiles of 32 conformations shared by teams of PM_TEAM_WARPS warps
        cpw = 32;
        team_warps = PM_TEAM_WARPS;
    }
    if (cpw != 0 && cpw != 32 && cpw != 16 && cpw != 8 && cpw != 4)
        return pm_fail("pm_create: conf_per_tile must be 0 (auto), 32, 16, 8, 4, or -8 (teams of 8 warps on tiles of 32)");
    if (rf) {  // the reaction-field variant is instantiated for one lane per conformation only
        cpw = 32;
        team_warps = 0;
    }
    PmHostProgram host;
    bool built = false;
    if (!cpw) {
        // fewest lanes per conformation that still gives >= 9 warps per SM (measured sweet spot); else the CPW with most
        // warps.  Planned with the real program of every candidate (building one takes milliseconds).
        int best_w = -1;
        for (int c = 32; c >= 4; c >>= 1) {
            PmHostProgram cand;
            if (!pm_build_program(N, c, t->n_bonds, t->bond_idx, t->n_angles, t->angle_idx, t->n_torsions, t->torsion_idx, t->torsion_per,
                                  cls, cand))
                return pm_fail("pm_create: degenerate bonded term (repeated atom)");
            const int n_derived = (2 * t->n_bonds + 2 * t->n_angles + 4 * t->n_torsions + 3 * cand.n_pairs + 1) & ~1;
            size_t sh, pw;
            int w;
            pm_plan((int)cand.blob.size(), n_derived, N, c, max_smem, 1, &sh, &pw, &w);
            if (w >= 9 || w > best_w) {
                best_w = w;
                cpw = c;
                host = cand;
                built = true;
                if (w >= 9) break;
            }
        }
        // Large molecules that need teams of lanes anyway: teams of 8 WARPS on tiles of 32 conformations instead -- every branch is
        // uniform and every index a broadcast, as with one lane per conformation; the records of a step are computed concurrently
        // and their write-backs ordered by colour through the team's split barrier.  Measured on B200 with the coloured rounds
        // (M conformations/s, lanes vs warps, 2^18 conformations): lig28 297 vs 235, lig33 250 vs 201, lig37 171 vs 167,
        // norfloxacin (41) 165 vs 150, lig45 141.5 vs 149.5, lig50 113 vs 128.7, lig55 98.2 vs 96.1 (one team per SM),
        // lig60 75.1 vs 83.8 (one team), lig64 .. vs 87.2.  Rule: from 44 atoms on when two teams (16 warps) fit; with room for one
        // team only, when the lane-team plan has at most 13 warps (about 58 atoms on).
        const char *et = getenv("PM_TEAM_WARPS_AUTO");
        const int lane_warps = best_w;
        if (cpw <= 8 && N >= 44 && !(et && et[0] == '0')) {
            PmHostProgram cand;
            if (pm_build_program(N, 32, t->n_bonds, t->bond_idx, t->n_angles, t->angle_idx, t->n_torsions, t->torsion_idx, t->torsion_per, cls,
                                 cand, PM_TEAM_WARPS)) {
                const int n_derived = (2 * t->n_bonds + 2 * t->n_angles + 4 * t->n_torsions + 3 * cand.n_pairs + 1) & ~1;
                size_t sh, pw;
                int w;
                pm_plan((int)cand.blob.size(), n_derived, N, 32, max_smem, 1, &sh, &pw, &w, PM_TEAM_WARPS);
                if (w >= 2 * PM_TEAM_WARPS || (w >= PM_TEAM_WARPS && lane_warps <= 13)) {
                    cpw = 32;
                    team_warps = PM_TEAM_WARPS;
                    host = cand;
                }
            }
        }
    }
    if ((!built || (team_warps && host.team_warps != team_warps)) &&
        !pm_build_program(N, cpw, t->n_bonds, t->bond_idx, t->n_angles, t->angle_idx, t->n_torsions, t->torsion_idx, t->torsion_per, cls, host,
                          team_warps))
        return pm_fail("pm_create: degenerate bonded term (repeated atom)");

    pm_handle_s *h = new pm_handle_s();
    memset(&h->prog, 0, sizeof(PmProg));
    h->device = device;
    h->n_sm = prop.multiProcessorCount;
    h->max_smem = max_smem;
    h->d_ints = nullptr;
    h->d_metric = nullptr;
    h->d_pair_map = nullptr;
    h->d_acos = nullptr;
    h->host = host;
    h->rf = rf;
    h->team_warps = team_warps;
    h->grad_state = nullptr;
    h->grad_free = nullptr;
    h->spec_state = nullptr;
    h->egemm_state = nullptr;
    h->egemm_free = nullptr;
    h->spec_free = nullptr;
    h->bond_idx.assign(t->bond_idx, t->bond_idx + 2 * (size_t)t->n_bonds);
    h->angle_idx.assign(t->angle_idx, t->angle_idx + 3 * (size_t)t->n_angles);
    h->torsion_idx.assign(t->torsion_idx, t->torsion_idx + 4 * (size_t)t->n_torsions);
    h->torsion_per.assign(t->torsion_per, t->torsion_per + (size_t)t->n_torsions);
    PmProg &g = h->prog;
    g.n_atoms = N;
    g.n_bonds = t->n_bonds;
    g.n_angles = t->n_angles;
    g.n_torsions = t->n_torsions;
    g.n_pairs = host.n_pairs;
    g.cpw = cpw;
    g.L = team_warps ? team_warps : 32 / cpw;
    g.n_star_steps = host.n_star_steps;
    g.n_axis_steps = host.n_axis_steps;
    g.n_segs = host.n_segs;
    g.n_pair_pos = host.n_pair_pos;
    g.off_blocks = host.off_blocks;
    g.n_int = (int)host.blob.size();
    g.off_star_tab = host.off_star_tab;
    g.off_star_hdr = host.off_star_hdr;
    g.off_star_rec = host.off_star_rec;
    g.off_axis_tab = host.off_axis_tab;
    g.off_axis_hdr = host.off_axis_hdr;
    g.off_axis_rec = host.off_axis_rec;
    g.off_axis_terms = host.off_axis_terms;
    g.off_seg_hdr = host.off_seg_hdr;
    g.off_seg_tile = host.off_seg_tile;
    g.off_entries = host.off_entries;
    g.off_tper = host.off_tper;
    g.k_coul = t->one_4pi_eps0;
    g.pair_stride = rf ? 4 : 3;
    if (rf) {
        const double rc = t->cutoff, ed = t->rf_dielectric;
        g.rf_krf = pow(rc, -3.0) * (ed - 1.0) / (2.0 * ed + 1.0);
        g.rf_crf = (1.0 / rc) * (3.0 * ed) / (2.0 * ed + 1.0);
        g.rf_cut2 = rc * rc;
    }

    g.doff_bond = 0;
    g.doff_angle = g.doff_bond + 2 * g.n_bonds;
    g.doff_tors = g.doff_angle + 2 * g.n_angles;
    g.doff_pair = g.doff_tors + 4 * g.n_torsions;
    g.n_derived = g.doff_pair + g.pair_stride * g.n_pairs;
    g.n_derived = (g.n_derived + 1) & ~1;

    g.soff_bond = 0;
    g.soff_angle = g.soff_bond + 2 * g.n_bonds;
    g.soff_tors = g.soff_angle + 2 * g.n_angles;
    g.soff_part = g.soff_tors + 2 * g.n_torsions;
    g.soff_exc = g.soff_part + 3 * N;
    g.n_slots = g.soff_exc + 3 * t->n_exceptions;

    auto cleanup = [&]() {
        if (h->d_ints) cudaFree(h->d_ints);
        if (h->d_metric) cudaFree(h->d_metric);
        if (h->d_pair_map) cudaFree(h->d_pair_map);
        if (h->d_acos) cudaFree(h->d_acos);
        delete h;
    };
#define PM_CUDA_H(call)                                                         \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) {                                                \
            cleanup();                                                          \
            return pm_fail(std::string(#call) + ": " + cudaGetErrorString(_e)); \
        }                                                                       \
    } while (0)
    PM_CUDA_H(cudaMalloc(&h->d_ints, sizeof(int) * (host.blob.size() + 4)));
    PM_CUDA_H(cudaMemcpy(h->d_ints, host.blob.data(), sizeof(int) * host.blob.size(), cudaMemcpyHostToDevice));
    PM_CUDA_H(cudaMalloc(&h->d_pair_map, sizeof(int) * (host.pair_map.size() + 3)));
    if (!host.pair_map.empty())
        PM_CUDA_H(cudaMemcpy(h->d_pair_map, host.pair_map.data(), sizeof(int) * host.pair_map.size(), cudaMemcpyHostToDevice));
    std::vector<double> ident((size_t)9 * N, 0.0);
    for (int a = 0; a < N; ++a) ident[9 * a] = ident[9 * a + 4] = ident[9 * a + 8] = 1.0;
    PM_CUDA_H(cudaMalloc(&h->d_metric, sizeof(double) * 9 * N));
    PM_CUDA_H(cudaMemcpy(h->d_metric, ident.data(), sizeof(double) * 9 * N, cudaMemcpyHostToDevice));
    const std::vector<double> tab = pm_acos_table_host();
    PM_CUDA_H(cudaMalloc(&h->d_acos, sizeof(double) * tab.size()));
    PM_CUDA_H(cudaMemcpy(h->d_acos, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
    g.ints = h->d_ints;
    g.metric = h->d_metric;
    g.pair_map = h->d_pair_map;
    g.acos_tab = h->d_acos;

    for (int wf = 0; wf < 2; ++wf) {
        size_t sh, pw;
        int w;
        pm_plan(g.n_int, g.n_derived, N, cpw, max_smem, wf, &sh, &pw, &w, team_warps);
        const char *envw = getenv("PM_WARPS");
        if (envw && *envw && atoi(envw) > 0 && atoi(envw) < w) w = team_warps ? (atoi(envw) / team_warps) * team_warps : atoi(envw);
        // register budget of the variant: 168 (<= 12 warps), 144 (13-14 warps) or 128 registers per thread
        h->maxw[wf] = (w <= 12 || rf) ? 12 : w <= 14 ? 14 : 16;
        const char *envm = getenv("PM_MAXW");   // sweeps only
        if (envm && *envm && !rf && !team_warps && atoi(envm) >= w && (atoi(envm) == 12 || atoi(envm) == 14 || atoi(envm) == 16)) h->maxw[wf] = atoi(envm);
        if (rf && w > 12) w = 12;
        h->warps[wf] = w;
        h->smem_total[wf] = w > 0 ? sh + pw * (size_t)(team_warps ? w / team_warps : w) : 0;
    }
    if (h->warps[1] < 1 || h->warps[0] < 1) {
        cleanup();
        return pm_fail("pm_create: molecule too large for the shared-memory tiles of this kernel generation");
    }
    PM_CUDA_H(cudaFuncSetAttribute(pm_pick_kernel(cpw, 1, h->maxw[1], rf, team_warps), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_total[1]));
    PM_CUDA_H(cudaFuncSetAttribute(pm_pick_kernel(cpw, 0, h->maxw[0], rf, team_warps), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_total[0]));
    *out = h;
    return 0;
}

extern "C" int pm_destroy(pm_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaFree(h->d_ints);
    cudaFree(h->d_metric);
    cudaFree(h->d_pair_map);
    cudaFree(h->d_acos);
    if (h->grad_state && h->grad_free) h->grad_free(h->grad_state);
    if (h->spec_state && h->spec_free) h->spec_free(h->spec_state);
    if (h->egemm_state && h->egemm_free) h->egemm_free(h->egemm_state);
    delete h;
    return 0;
}

extern "C" int pm_n_slots(pm_handle h) { return h ? h->prog.n_slots : -1; }
extern "C" int pm_n_derived(pm_handle h) { return h ? h->prog.n_derived : -1; }
extern "C" int pm_n_pairs(pm_handle h) { return h ? h->prog.n_pairs : -1; }
extern "C" int pm_conf_per_tile(pm_handle h) { return h ? h->prog.cpw : -1; }
extern "C" long pm_n_tiles(pm_handle h, long n_conf) { return h ? (n_conf + h->prog.cpw - 1) / h->prog.cpw : -1; }
extern "C" long pm_tile_doubles(pm_handle h) { return h ? (long)3 * h->prog.n_atoms * h->prog.cpw : -1; }
extern "C" int pm_get_pairs(pm_handle h, int *out_host) {
    if (!h || !out_host) return pm_fail("pm_get_pairs: null argument");
    memcpy(out_host, h->host.pair_map.data(), sizeof(int) * h->host.pair_map.size());
    return 0;
}
extern "C" int pm_launch_info(pm_handle h, int with_forces, int *grid, int *block, int *smem_bytes, int *warps_per_cta) {
    if (!h) return pm_fail("pm_launch_info: null handle");
    const int wf = with_forces ? 1 : 0;
    if (grid) *grid = h->n_sm;
    if (block) *block = h->warps[wf] * 32;
    if (smem_bytes) *smem_bytes = (int)h->smem_total[wf];
    if (warps_per_cta) *warps_per_cta = h->warps[wf];
    pm_spec_launch_info(h, wf, block, smem_bytes, warps_per_cta);  // full evaluations run the specialised kernels if present
    return 0;
}
extern "C" int pm_program_info(pm_handle h, int *out10) {
    if (!h || !out10) return pm_fail("pm_program_info: null argument");
    const PmHostProgram &p = h->host;
    const int v[10] = {h->prog.cpw, h->prog.L, p.n_stars, p.n_star_steps, p.n_axes, p.n_axis_steps, p.n_segs, p.n_entries,
                       p.n_entry_slots, (int)p.blob.size()};
    memcpy(out10, v, sizeof(v));
    return 0;
}

extern "C" int pm_set_force_metric(pm_handle h, const double *metric_host) {
    if (!h || !metric_host) return pm_fail("pm_set_force_metric: null argument");
    PM_CUDA(cudaSetDevice(h->device));
    PM_CUDA(cudaMemcpy(h->d_metric, metric_host, sizeof(double) * 9 * h->prog.n_atoms, cudaMemcpyHostToDevice));
    if (pm_spec_metric_changed(h)) return 1;
    return 0;
}

extern "C" int pm_pack_tiles(pm_handle h, const double *src_dev, long n_conf, double *tiles_dev, void *stream) {
    if (!h || !src_dev || !tiles_dev) return pm_fail("pm_pack_tiles: null argument");
    if (n_conf < 1) return pm_fail("pm_pack_tiles: no conformations");
    PM_CUDA(cudaSetDevice(h->device));
    const int row = 3 * h->prog.n_atoms, cpw = h->prog.cpw;
    const size_t smem = sizeof(double) * cpw * (row + 1);
    if (smem > 48 * 1024) PM_CUDA(cudaFuncSetAttribute(pm_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pm_pack_kernel<<<(unsigned)pm_n_tiles(h, n_conf), 256, smem, (cudaStream_t)stream>>>(src_dev, n_conf, row, cpw, tiles_dev);
    PM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_unpack_tiles(pm_handle h, const double *tiles_dev, long n_conf, double *dst_dev, void *stream) {
    if (!h || !dst_dev || !tiles_dev) return pm_fail("pm_unpack_tiles: null argument");
    if (n_conf < 1) return pm_fail("pm_unpack_tiles: no conformations");
    PM_CUDA(cudaSetDevice(h->device));
    const int row = 3 * h->prog.n_atoms, cpw = h->prog.cpw;
    const size_t smem = sizeof(double) * cpw * (row + 1);
    if (smem > 48 * 1024) PM_CUDA(cudaFuncSetAttribute(pm_unpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pm_unpack_kernel<<<(unsigned)pm_n_tiles(h, n_conf), 256, smem, (cudaStream_t)stream>>>(tiles_dev, n_conf, row, cpw, dst_dev);
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

// rows of the fused reduction one launch of the interpreting E(+F) kernel writes per parameter vector (one per warp of the grid)
static long pm_red_rows(pm_handle h, int with_forces, long n_conf) {
    const int wf = with_forces ? 1 : 0;
    const int warps = h->team_warps ? h->warps[wf] / h->team_warps : h->warps[wf];  // tiles in flight per CTA (teams, or warps)
    const long n_tiles = pm_n_tiles(h, n_conf);
    long grid = h->n_sm;
    if (grid * warps > n_tiles) grid = (n_tiles + warps - 1) / warps;
    if (grid < 1) grid = 1;
    return grid * warps;
}

// full evaluations of a small molecule go to the generated kernels: one launch (plus a constant-table copy) per parameter vector,
// so batches over small ensembles (launch-bound: < 32768 conformations) stay on the interpreter, which loops over the vectors
// inside one launch
static bool pm_use_spec(pm_handle h, int with_forces, int n_vec, long n_conf, int mode) {
    return mode == 0 && pm_spec_can_eval(h, with_forces ? 1 : 0) && (n_vec == 1 || n_conf >= 32768);
}

static int pm_eval_core(pm_handle h, const double *derived_dev, int n_vec, const double *coord_tiles_dev,
                        const double *fref_tiles_dev, long n_conf, int with_forces, double *emm_dev, double *fres_dev,
                        double *fmm_tiles_dev, int mode, const double *nb_e_dev, const double *nb_f_tiles_dev, PmRed red,
                        void *stream) {
    if (!h || !derived_dev || !coord_tiles_dev || (!emm_dev && !red.rows)) return pm_fail("pm_eval: null argument");
    if (n_vec < 1 || n_conf < 1) return pm_fail("pm_eval: empty batch");
    if (fres_dev && !fref_tiles_dev) return pm_fail("pm_eval: fres requested without reference forces");
    if (!with_forces && (fres_dev || fmm_tiles_dev)) return pm_fail("pm_eval: force outputs requested with with_forces=0");
    if (mode < 0 || mode > 2) return pm_fail("pm_eval: mode must be 0 (full), 1 (nonbonded only) or 2 (bonded + cached nonbonded)");
    if (mode == 2 && (!nb_e_dev || (with_forces && !nb_f_tiles_dev))) return pm_fail("pm_eval: mode 2 needs the nonbonded cache");
    PM_CUDA(cudaSetDevice(h->device));
    if (!red.rows && pm_use_spec(h, with_forces, n_vec, n_conf, mode))
        return pm_spec_eval(h, derived_dev, n_vec, coord_tiles_dev, fref_tiles_dev, n_conf, emm_dev, fres_dev, fmm_tiles_dev,
                            with_forces ? 1 : 0, PmRed{nullptr, nullptr, 0.0, nullptr}, (cudaStream_t)stream);
    const int wf = with_forces ? 1 : 0;
    const long n_tiles = pm_n_tiles(h, n_conf);
    const int warps = h->warps[wf], units = h->team_warps ? warps / h->team_warps : warps;  // tiles in flight per CTA
    long grid = h->n_sm;
    if (grid * units > n_tiles) grid = (n_tiles + units - 1) / units;
    if (grid < 1) grid = 1;
    PmProg prog = h->prog;
    if (mode == 1) {
        prog.n_star_steps = 0;
        prog.n_axis_steps = 0;
    } else if (mode == 2) {
        prog.n_segs = 0;
        prog.n_pair_pos = 0;
        prog.nb_e = nb_e_dev;
        prog.nb_f_tiles = nb_f_tiles_dev;
    }
    pm_ef_fn fn = pm_pick_kernel(h->prog.cpw, wf, h->maxw[wf], h->rf, h->team_warps);
    fn<<<(unsigned)grid, warps * 32, h->smem_total[wf], (cudaStream_t)stream>>>(prog, derived_dev, n_vec, coord_tiles_dev,
                                                                                wf ? fref_tiles_dev : nullptr, n_conf, n_tiles, emm_dev,
                                                                                wf ? fres_dev : nullptr, wf ? fmm_tiles_dev : nullptr, red);
    PM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_eval_ex(pm_handle h, const double *derived_dev, int n_vec, const double *coord_tiles_dev,
                          const double *fref_tiles_dev, long n_conf, int with_forces, double *emm_dev, double *fres_dev,
                          double *fmm_tiles_dev, int mode, const double *nb_e_dev, const double *nb_f_tiles_dev, void *stream) {
    if (!emm_dev) return pm_fail("pm_eval: null argument");
    // the kernels compute the force residual whenever they are given reference forces: only hand them over when it is wanted
    return pm_eval_core(h, derived_dev, n_vec, coord_tiles_dev, fres_dev ? fref_tiles_dev : nullptr, n_conf, with_forces, emm_dev,
                        fres_dev, fmm_tiles_dev, mode, nb_e_dev, nb_f_tiles_dev, PmRed{nullptr, nullptr, 0.0, nullptr}, stream);
}

// ---- fused evaluation + residual reduction (+ cross-GPU sum) ------------------------------------------------------
static int pm_stage1_blocks(long n_conf) {
    const long need = (n_conf + PM_RED_THREADS - 1) / PM_RED_THREADS;
    return (int)(need < PM_RED_BLOCKS ? (need < 1 ? 1 : need) : PM_RED_BLOCKS);
}

extern "C" long pm_eval_reduce_scratch_doubles(pm_handle h, int n_vec, long n_conf) {
    if (!h || n_vec < 1 || n_conf < 1) return -1;
    const long r0 = pm_red_rows(h, 0, n_conf), r1 = pm_red_rows(h, 1, n_conf);
    long rows = r0 > r1 ? r0 : r1;
    if (rows < PM_RED_BLOCKS) rows = PM_RED_BLOCKS;
    const long s0 = pm_spec_rows(h, 0, n_conf), s1 = pm_spec_rows(h, 1, n_conf);
    if (rows < s0) rows = s0;
    if (rows < s1) rows = s1;
    // generated kernels without fused sums: per-conformation energies and residuals of the batch are staged behind the rows
    const long stage = ((pm_spec_can_eval(h, 0) || pm_spec_can_eval(h, 1)) && !pm_spec_fused(h)) ? 2L * n_vec * n_conf : 0L;
    return (long)n_vec * rows * 5 + stage;
}

extern "C" int pm_eval_reduce(pm_handle h, const double *derived_dev, int n_vec, const double *coord_tiles_dev,
                              const double *fref_tiles_dev, long n_conf, int with_forces, const double *eref_dev,
                              const double *weights_dev, double d_shift, double *partials_dev, double *scratch_dev, pm_comm comm,
                              double *emm_dev, double *fres_dev, void *stream) {
    if (!h || !partials_dev || !scratch_dev) return pm_fail("pm_eval_reduce: null argument");
    if (n_vec < 1 || n_conf < 1) return pm_fail("pm_eval_reduce: empty batch");
    if (!with_forces && (fref_tiles_dev || fres_dev)) return pm_fail("pm_eval_reduce: force inputs/outputs given with with_forces=0");
    if (comm && n_vec > pm_comm_max_vec(comm)) return pm_fail("pm_eval_reduce: more parameter vectors than the communicator was created for");
    if (pm_use_spec(h, with_forces, n_vec, n_conf, 0) && pm_spec_fused(h)) {
        // generated kernel with the residual sums in its epilogue: rows of five per warp -> row sum (+ all-reduce)
        PM_CUDA(cudaSetDevice(h->device));
        const int rc = pm_spec_eval(h, derived_dev, n_vec, coord_tiles_dev, fref_tiles_dev, n_conf, emm_dev, fres_dev, nullptr,
                                    with_forces ? 1 : 0, PmRed{eref_dev, weights_dev, d_shift, scratch_dev}, (cudaStream_t)stream);
        if (rc) return rc;
        return pm_reduce_rows(scratch_dev, pm_spec_rows(h, with_forces ? 1 : 0, n_conf), n_vec, n_conf, partials_dev, comm, (cudaStream_t)stream);
    }
    if (pm_use_spec(h, with_forces, n_vec, n_conf, 0)) {
        // generated kernel (round-1 form: energy and residual per conformation) -> stage-1 sums as rows -> row sum (+ all-reduce)
        const long r0 = pm_red_rows(h, 0, n_conf), r1 = pm_red_rows(h, 1, n_conf);
        long rows = r0 > r1 ? r0 : r1;
        if (rows < PM_RED_BLOCKS) rows = PM_RED_BLOCKS;
        const long s0 = pm_spec_rows(h, 0, n_conf), s1 = pm_spec_rows(h, 1, n_conf);
        if (rows < s0) rows = s0;
        if (rows < s1) rows = s1;
        double *stage = scratch_dev + (size_t)n_vec * rows * 5;
        double *e_buf = emm_dev ? emm_dev : stage;
        double *r_buf = fref_tiles_dev ? (fres_dev ? fres_dev : stage + (size_t)n_vec * n_conf) : nullptr;
        int rc = pm_eval_core(h, derived_dev, n_vec, coord_tiles_dev, fref_tiles_dev, n_conf, with_forces, e_buf, r_buf, nullptr, 0, nullptr,
                              nullptr, PmRed{nullptr, nullptr, 0.0, nullptr}, stream);
        if (rc) return rc;
        const int nb = pm_stage1_blocks(n_conf);
        pm_reduce_stage1<<<dim3(nb, n_vec), PM_RED_THREADS, 0, (cudaStream_t)stream>>>(e_buf, r_buf, eref_dev, weights_dev, d_shift, n_conf,
                                                                                     scratch_dev);
        PM_CUDA(cudaGetLastError());
        return pm_reduce_rows(scratch_dev, nb, n_vec, n_conf, partials_dev, comm, (cudaStream_t)stream);
    }
    PmRed red{eref_dev, weights_dev, d_shift, scratch_dev};
    const int rc = pm_eval_core(h, derived_dev, n_vec, coord_tiles_dev, fref_tiles_dev, n_conf, with_forces, emm_dev, fres_dev, nullptr,
                                0, nullptr, nullptr, red, stream);
    if (rc) return rc;
    return pm_reduce_rows(scratch_dev, pm_red_rows(h, with_forces, n_conf), n_vec, n_conf, partials_dev, comm, (cudaStream_t)stream);
}

extern "C" int pm_eval(pm_handle h, const double *derived_dev, int n_vec, const double *coord_tiles_dev,
                       const double *fref_tiles_dev, long n_conf, int with_forces, double *emm_dev, double *fres_dev,
                       double *fmm_tiles_dev, void *stream) {
    return pm_eval_ex(h, derived_dev, n_vec, coord_tiles_dev, fref_tiles_dev, n_conf, with_forces, emm_dev, fres_dev, fmm_tiles_dev,
                      0, nullptr, nullptr, stream);
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
    const int threads = 512, blocks = prop.multiProcessorCount * 2;
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
