// pm_delta_kernels.cu -- delta evaluation of trial parameter vectors that differ from a stored BASE vector in a few terms
// (sm_100a).
//
// The reference's Monte Carlo and simulated-annealing optimisers move ONE parameter per objective evaluation
// (ParaMol/Optimizers/monte_carlo.py:89-100, simulated_annealing.py:90-131) and re-evaluate every term of every conformation
// for it (ParaMol/Objective_function/objective_function.py:231-359 -> System/system.py:324-354).  A single-parameter move
// changes only the terms that read the parameter: for a conformation s
//     E_s(p)  = E_s(base) + sum_{t affected} [E_t(new) - E_t(old)]
//     F_sa(p) = F_sa(base) + Delta_sa,           Delta = sum_{t affected} [F_t(new) - F_t(old)]   (a few atoms)
//     res_s(p) = res_s(base) + sum_{a affected} [Delta^T (M + M^T)(F_base - F_ref) + Delta^T M Delta]_sa
// so the kernel needs the geometry of the affected terms only, and writes the same per-conformation outputs (emm, fres) as
// pm_eval; everything downstream (reduction, weights, allreduce) is unchanged.  commit = 1 folds ONE proposal into the base
// arrays instead (accepted move).
//
// Mapping: one conformation per thread (coalesced along the tile columns), proposals in an inner loop; the force deltas of the
// (at most max_loc) affected atoms accumulate in shared memory, [3 max_loc][threads].  HBM-bound by construction: per proposal
// and conformation it touches the coordinates and base forces of the affected atoms and writes 16 bytes.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>

#include "pm_device.cuh"
#include "pm_internal.hpp"

#define PMD_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) return pm_fail_msg(std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

enum { PMD_BOND = 0, PMD_ANGLE = 1, PMD_TORSION = 2, PMD_PAIR = 3 };

struct PmDeltaArgs {
    int n_atoms, cpw, n_vec, max_loc, commit;
    long n_conf;
    const double *derived_base, *derived;  // [n_derived], [n_vec][n_derived]
    int n_derived;
    const int *prop_ptr;   // [n_vec + 1] into entries
    const int4 *entries;   // 3 int4 per entry: (kind, derived offset, periodicity, 0), atoms, local atom indices
    const int *atom_ptr;   // [n_vec + 1] into atom_list
    const int *atom_list;  // global atom index of every local atom of every proposal
    const double *x_tiles, *fref_tiles, *metric;
    double *fbase_tiles;   // read; written when commit
    double *e_base, *res_base;  // [S]; written when commit
    double *emm, *fres;    // [n_vec][S] outputs (unused when commit)
    const double2 *acos_tab;
};

__device__ __forceinline__ pm_d3 pmd_ldx(const double *__restrict__ t, long base, int a, int cpw) {
    return pm_d3{t[base + (size_t)(3 * a) * cpw], t[base + (size_t)(3 * a + 1) * cpw], t[base + (size_t)(3 * a + 2) * cpw]};
}

__global__ void __launch_bounds__(128) pm_delta_kernel(const PmDeltaArgs g) {
    extern __shared__ double sdf[];  // [3 max_loc][blockDim.x]
    const int T = blockDim.x, tid = threadIdx.x;
    const long s = (long)blockIdx.x * T + tid;
    if (s >= g.n_conf) return;
    const int N = g.n_atoms, cpw = g.cpw;
    const long base = (s / cpw) * ((long)3 * N * cpw) + (s % cpw);
    const bool with_f = g.fbase_tiles != nullptr;
    const double e0 = g.e_base[s], r0 = with_f ? g.res_base[s] : 0.0;
    double *df = sdf + tid;
    auto add = [&](int l, const pm_d3 v) {
        df[(3 * l) * T] += v.x;
        df[(3 * l + 1) * T] += v.y;
        df[(3 * l + 2) * T] += v.z;
    };
    for (int p = 0; p < g.n_vec; ++p) {
        const int n_loc = g.atom_ptr[p + 1] - g.atom_ptr[p];
        if (with_f)
            for (int k = 0; k < 3 * n_loc; ++k) df[k * T] = 0.0;
        const double *dn = g.derived + (size_t)p * g.n_derived;
        double de = 0.0;
        for (int e = g.prop_ptr[p]; e < g.prop_ptr[p + 1]; ++e) {
            const int4 me = g.entries[3 * e], at = g.entries[3 * e + 1], lo = g.entries[3 * e + 2];
            const double *po = g.derived_base + me.y, *pn = dn + me.y;
            if (me.x == PMD_PAIR) {
                const pm_d3 d = pm_sub(pmd_ldx(g.x_tiles, base, at.x, cpw), pmd_ldx(g.x_tiles, base, at.y, cpw));
                const double rinv = pm_rsqrt(pm_dot(d, d)), u = rinv * rinv, u3 = u * u * u;
                const double d12 = pn[0] - po[0], d6 = pn[1] - po[1], dq = pn[2] - po[2];
                de += fma(d12 * u3 - d6, u3, dq * rinv);
                if (with_f) {
                    const double gg = fma(6.0 * u3, 2.0 * d12 * u3 - d6, dq * rinv) * u;
                    add(lo.x, pm_d3{gg * d.x, gg * d.y, gg * d.z});
                    add(lo.y, pm_d3{-gg * d.x, -gg * d.y, -gg * d.z});
                }
            } else if (me.x == PMD_BOND) {
                const pm_d3 d = pm_sub(pmd_ldx(g.x_tiles, base, at.y, cpw), pmd_ldx(g.x_tiles, base, at.x, cpw));
                const double r2 = pm_dot(d, d), rinv = pm_rsqrt(r2), r = r2 * rinv;
                const double dlo = r - po[0], dln = r - pn[0];
                de += 0.5 * (pn[1] * dln * dln - po[1] * dlo * dlo);
                if (with_f) {
                    const double gg = (pn[1] * dln - po[1] * dlo) * rinv;  // (dE/dr new - old) / r
                    add(lo.x, pm_d3{gg * d.x, gg * d.y, gg * d.z});
                    add(lo.y, pm_d3{-gg * d.x, -gg * d.y, -gg * d.z});
                }
            } else if (me.x == PMD_ANGLE) {
                const pm_d3 xa = pmd_ldx(g.x_tiles, base, at.x, cpw), xb = pmd_ldx(g.x_tiles, base, at.y, cpw),
                            xc = pmd_ldx(g.x_tiles, base, at.z, cpw);
                const pm_d3 d0 = pm_sub(xb, xa), d1 = pm_sub(xb, xc), pc = pm_cross(d0, d1);
                const double q0 = pm_dot(d0, d0), q1 = pm_dot(d1, d1), pp2 = pm_dot(pc, pc);
                const double i01 = pm_rsqrt(q0 * q1), ip = pm_rsqrt(fmax(pp2, 1.0e-12));
                const double th = pm_acos_cs(pm_dot(d0, d1) * i01, pp2 * ip * i01, g.acos_tab);
                const double dlo = th - po[0], dln = th - pn[0];
                de += 0.5 * (pn[1] * dln * dln - po[1] * dlo * dlo);
                if (with_f) {
                    const double dd = pn[1] * dln - po[1] * dlo;  // dE/dtheta new - old
                    const double ta = dd * ip * (q1 * i01 * i01), tc = -dd * ip * (q0 * i01 * i01);
                    const pm_d3 c0 = pm_cross(d0, pc), c2 = pm_cross(d1, pc);
                    const pm_d3 fa{ta * c0.x, ta * c0.y, ta * c0.z}, fc{tc * c2.x, tc * c2.y, tc * c2.z};
                    add(lo.x, fa);
                    add(lo.z, fc);
                    add(lo.y, pm_d3{-(fa.x + fc.x), -(fa.y + fc.y), -(fa.z + fc.z)});
                }
            } else {  // torsion: derived = (k, k n, cos phase, sin phase)
                const pm_d3 x0 = pmd_ldx(g.x_tiles, base, at.x, cpw), x1 = pmd_ldx(g.x_tiles, base, at.y, cpw),
                            x2 = pmd_ldx(g.x_tiles, base, at.z, cpw), x3 = pmd_ldx(g.x_tiles, base, at.w, cpw);
                const pm_d3 d0 = pm_sub(x0, x1), d1 = pm_sub(x2, x1), d2 = pm_sub(x2, x3);
                const pm_d3 c1 = pm_cross(d0, d1), c2 = pm_cross(d1, d2);
                const double n1 = pm_dot(c1, c1), n2 = pm_dot(c2, c2), bb = pm_dot(d1, d1);
                const double inv = 1.0 / (n1 * n2 * bb);
                const double in1 = inv * n2 * bb, in2 = inv * n1 * bb, ibb = inv * n1 * n2;
                const double i12 = pm_rsqrt(n1 * n2), nb = bb * pm_rsqrt(bb);
                const double cphi = pm_dot(c1, c2) * i12, sphi = nb * pm_dot(d0, c2) * i12;
                double cn = 1.0, sn = 0.0;
                for (int q = 0; q < me.z; ++q) {
                    const double t = cn * cphi - sn * sphi;
                    sn = fma(sn, cphi, cn * sphi);
                    cn = t;
                }
                const double cpo = cn * po[2] + sn * po[3], spo = sn * po[2] - cn * po[3];
                const double cpn = cn * pn[2] + sn * pn[3], spn = sn * pn[2] - cn * pn[3];
                de += pn[0] * (1.0 + cpn) - po[0] * (1.0 + cpo);
                if (with_f) {
                    const double dd = -(pn[1] * spn - po[1] * spo);  // dE/dphi new - old = -(k n sin psi)
                    const double ff1 = pm_dot(d0, d1) * ibb, ff2 = pm_dot(d2, d1) * ibb;
                    const double s0 = -dd * nb * in1, s3 = dd * nb * in2;  // F_0 = -dd g0, F_3 = -dd g3 (g3 = -|d1| c2 / n2)
                    const pm_d3 f0{s0 * c1.x, s0 * c1.y, s0 * c1.z}, f3{s3 * c2.x, s3 * c2.y, s3 * c2.z};
                    // F_1 = (ff1 - 1) F_0 - ff2 F_3,  F_2 = (ff2 - 1) F_3 - ff1 F_0
                    add(lo.x, f0);
                    add(lo.w, f3);
                    add(lo.y, pm_d3{(ff1 - 1.0) * f0.x - ff2 * f3.x, (ff1 - 1.0) * f0.y - ff2 * f3.y, (ff1 - 1.0) * f0.z - ff2 * f3.z});
                    add(lo.z, pm_d3{(ff2 - 1.0) * f3.x - ff1 * f0.x, (ff2 - 1.0) * f3.y - ff1 * f0.y, (ff2 - 1.0) * f3.z - ff1 * f0.z});
                }
            }
        }
        double dres = 0.0;
        if (with_f) {
            const int *al = g.atom_list + g.atom_ptr[p];
            for (int l = 0; l < n_loc; ++l) {
                const int a = al[l];
                const double dx = df[(3 * l) * T], dy = df[(3 * l + 1) * T], dz = df[(3 * l + 2) * T];
                const pm_d3 fb = pmd_ldx(g.fbase_tiles, base, a, cpw), fr = pmd_ldx(g.fref_tiles, base, a, cpw);
                // D = F_base - F_ref;  (D + d)^T M (D + d) - D^T M D = d^T M (2D + d) + (D^T M d - d^T M D) -> exact form below
                const double ex = fb.x - fr.x, ey = fb.y - fr.y, ez = fb.z - fr.z;
                const double *m = g.metric + 9 * a;
                const double vx = ex + ex + dx, vy = ey + ey + dy, vz = ez + ez + dz;  // 2D + d
                // d^T M (D + d) + D^T M d  =  d^T M (2D + d) + (D^T M d - d^T M D); with w = M^T D the last bracket is
                // d . (w - M D), which vanishes for symmetric M and is kept so that any metric is handled exactly.
                const double mx = fma(m[2], vz, fma(m[1], vy, m[0] * vx)), my = fma(m[5], vz, fma(m[4], vy, m[3] * vx)),
                             mz = fma(m[8], vz, fma(m[7], vy, m[6] * vx));
                const double ax = (m[3] - m[1]) * ey + (m[6] - m[2]) * ez, ay = (m[1] - m[3]) * ex + (m[7] - m[5]) * ez,
                             az = (m[2] - m[6]) * ex + (m[5] - m[7]) * ey;  // (M^T - M) D
                dres += dx * (mx + ax) + dy * (my + ay) + dz * (mz + az);
                if (g.commit) {
                    g.fbase_tiles[base + (size_t)(3 * a) * cpw] = fb.x + dx;
                    g.fbase_tiles[base + (size_t)(3 * a + 1) * cpw] = fb.y + dy;
                    g.fbase_tiles[base + (size_t)(3 * a + 2) * cpw] = fb.z + dz;
                }
            }
        }
        if (g.commit) {
            g.e_base[s] = e0 + de;
            if (with_f) g.res_base[s] = r0 + dres;
        } else {
            g.emm[(size_t)p * g.n_conf + s] = e0 + de;
            if (g.fres) g.fres[(size_t)p * g.n_conf + s] = r0 + dres;
        }
    }
}

extern "C" int pm_delta_max_local_atoms(void) { return 64; }

extern "C" int pm_delta_eval(pm_handle h, const double *derived_base_dev, const double *derived_dev, int n_vec,
                             const int *prop_ptr_dev, const int *entries_dev, const int *atom_ptr_dev, const int *atom_list_dev,
                             int max_loc, const double *coord_tiles_dev, const double *fref_tiles_dev, double *fbase_tiles_dev,
                             double *e_base_dev, double *res_base_dev, long n_conf, double *emm_dev, double *fres_dev, int commit,
                             void *stream) {
    if (!h || !derived_base_dev || !derived_dev || !prop_ptr_dev || !entries_dev || !atom_ptr_dev || !atom_list_dev ||
        !coord_tiles_dev || !e_base_dev)
        return pm_fail_msg("pm_delta_eval: null argument");
    if (n_vec < 1 || n_conf < 1) return pm_fail_msg("pm_delta_eval: empty batch");
    if (commit && n_vec != 1) return pm_fail_msg("pm_delta_eval: commit folds exactly one proposal into the base");
    if (!commit && !emm_dev) return pm_fail_msg("pm_delta_eval: no output array");
    if (fbase_tiles_dev && (!fref_tiles_dev || !res_base_dev)) return pm_fail_msg("pm_delta_eval: force mode needs reference forces and base residuals");
    if (fres_dev && !fbase_tiles_dev) return pm_fail_msg("pm_delta_eval: fres requested without base forces");
    if (max_loc < 1 || max_loc > pm_delta_max_local_atoms()) return pm_fail_msg("pm_delta_eval: too many affected atoms for one proposal");
    if (h->rf) return pm_fail_msg("pm_delta_eval: not available for the cutoff / reaction-field variant");
    PMD_CUDA(cudaSetDevice(h->device));
    const int threads = max_loc <= 32 ? 128 : 64;
    const size_t smem = sizeof(double) * 3 * (size_t)max_loc * threads;
    if (smem > 48 * 1024) PMD_CUDA(cudaFuncSetAttribute(pm_delta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PmDeltaArgs a;
    a.n_atoms = h->prog.n_atoms;
    a.cpw = h->prog.cpw;
    a.n_vec = n_vec;
    a.max_loc = max_loc;
    a.commit = commit;
    a.n_conf = n_conf;
    a.derived_base = derived_base_dev;
    a.derived = derived_dev;
    a.n_derived = h->prog.n_derived;
    a.prop_ptr = prop_ptr_dev;
    a.entries = reinterpret_cast<const int4 *>(entries_dev);
    a.atom_ptr = atom_ptr_dev;
    a.atom_list = atom_list_dev;
    a.x_tiles = coord_tiles_dev;
    a.fref_tiles = fref_tiles_dev;
    a.metric = h->d_metric;
    a.fbase_tiles = fbase_tiles_dev;
    a.e_base = e_base_dev;
    a.res_base = res_base_dev;
    a.emm = emm_dev;
    a.fres = fres_dev;
    a.acos_tab = h->d_acos;
    const long blocks = (n_conf + threads - 1) / threads;
    pm_delta_kernel<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(a);
    PMD_CUDA(cudaGetLastError());
    return 0;
}

/* Offsets a caller needs to build the entry list: where each term kind starts in the derived table and the pair stride. */
extern "C" int pm_derived_layout(pm_handle h, int *out5) {
    if (!h || !out5) return pm_fail_msg("pm_derived_layout: null argument");
    out5[0] = h->prog.doff_bond;
    out5[1] = h->prog.doff_angle;
    out5[2] = h->prog.doff_tors;
    out5[3] = h->prog.doff_pair;
    out5[4] = h->prog.pair_stride;
    return 0;
}
