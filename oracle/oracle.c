/*
 * oracle.c -- CPU restatement (FP64) of the arithmetic behind ParaMol's objective-evaluation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under paramol_b200/ may import, link or call this file; it is
 * the checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * What it restates
 * ----------------
 * ParaMol (reference, /root/reference) evaluates MM energies and forces through the third-party
 * OpenMM library (pinned openmm=7.7.0, .github/workflows/paramol_env.yml:109), Reference platform,
 * called from
 *     ParaMol/MM_engines/openmm.py:448-466   get_potential_energy  (context.setPositions; getState(getEnergy))
 *     ParaMol/MM_engines/openmm.py:484-502   get_forces            (getState(getForces))
 *     ParaMol/System/system.py:324-354       get_forces_ensemble / get_energies_ensemble (loop over conformations)
 *     ParaMol/Objective_function/cpu_objective_function.py:67-108   fused E+F worker loop
 * OpenMM's sources are not part of the reference tree, so the per-term formulas below restate the
 * published algorithms of OpenMM's Reference platform (platforms/reference/src/SimTKReference:
 * ReferenceHarmonicBondIxn, ReferenceAngleBondIxn, ReferenceProperDihedralBond, ReferenceLJCoulombIxn,
 * ReferenceLJCoulomb14) in the same operation order, with ONE_4PI_EPS0 = 138.935456 by default.
 * Parity is PINNED: tests/test_oracle.py checks this file against every golden the reference's own
 * tests hold for the path (energies, forces, objective values, reaction-field variant).
 *
 * Also here: the internal-coordinate conventions of ParaMol/Utils/geometry.py:4-107 (LLS design matrix)
 * and the inverse-distance / normal-equation sums of ParaMol/MM_engines/resp.py:99-134,381-524.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PMO_NOCUTOFF 0
#define PMO_CUTOFF_NONPERIODIC 1

typedef struct {
    int n_atoms;
    int n_bonds;    const int *bond_idx;    const double *bond_par;     /* (r0, k)        */
    int n_angles;   const int *angle_idx;   const double *angle_par;    /* (theta0, k)    */
    int n_torsions; const int *tors_idx;    const int *tors_per; const double *tors_par; /* (phase, k) */
    const double *particle_par;                                         /* (q, sigma, eps) */
    int n_exceptions; const int *exc_idx;   const double *exc_par;      /* (qq, sigma, eps) */
    int method; double cutoff; double rf_dielectric; double k_coul;
} pmo_system;

static inline void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* ReferenceForce::getDeltaR(I, J): J - I, with R2 and R */
static inline void delta_r(const double *xi, const double *xj, double *d /* [5]: dx dy dz r2 r */) {
    d[0] = xj[0] - xi[0];
    d[1] = xj[1] - xi[1];
    d[2] = xj[2] - xi[2];
    d[3] = dot3(d, d);
    d[4] = sqrt(d[3]);
}

/* SimTKOpenMMUtilities / ReferenceBondIxn::getAngleBetweenTwoVectors */
static double angle_between(const double *v1, const double *v2) {
    double dot = dot3(v1, v2);
    double n1 = dot3(v1, v1), n2 = dot3(v2, v2);
    double cosine = dot / sqrt(n1 * n2);
    double angle;
    if (cosine >= 1.0) angle = 0.0;
    else if (cosine <= -1.0) angle = M_PI;
    else angle = acos(cosine);
    if (cosine > 0.99 || cosine < -0.99) {
        /* acos() loses accuracy near 0 / pi: use the cross product (as the Reference platform does) */
        double c[3];
        cross3(v1, v2, c);
        double scale = n1 * n2;
        angle = asin(sqrt(dot3(c, c) / scale));
        if (cosine < 0.0) angle = M_PI - angle;
    }
    return angle;
}

/* One conformation: energy returned, forces accumulated into f (may be NULL). */
static double eval_one(const pmo_system *s, const double *x, double *f) {
    const int N = s->n_atoms;
    double e_bond = 0.0, e_angle = 0.0, e_tors = 0.0, e_nb = 0.0;

    /* ---- HarmonicBondForce (ReferenceHarmonicBondIxn) ---- */
    for (int b = 0; b < s->n_bonds; ++b) {
        int ia = s->bond_idx[2 * b], ib = s->bond_idx[2 * b + 1];
        double r0 = s->bond_par[2 * b], k = s->bond_par[2 * b + 1];
        double d[5];
        delta_r(x + 3 * ia, x + 3 * ib, d);
        double di = d[4] - r0;
        double dEdR = k * di;
        dEdR = d[4] > 0.0 ? dEdR / d[4] : 0.0;
        if (f) for (int c = 0; c < 3; ++c) { f[3 * ia + c] += dEdR * d[c]; f[3 * ib + c] -= dEdR * d[c]; }
        e_bond += 0.5 * k * di * di;
    }

    /* ---- HarmonicAngleForce (ReferenceAngleBondIxn) ---- */
    for (int a = 0; a < s->n_angles; ++a) {
        int ia = s->angle_idx[3 * a], ib = s->angle_idx[3 * a + 1], ic = s->angle_idx[3 * a + 2];
        double t0 = s->angle_par[2 * a], k = s->angle_par[2 * a + 1];
        double d0[5], d1[5], p[3];
        delta_r(x + 3 * ia, x + 3 * ib, d0); /* B - A */
        delta_r(x + 3 * ic, x + 3 * ib, d1); /* B - C */
        cross3(d0, d1, p);
        double rp = sqrt(dot3(p, p));
        if (rp < 1.0e-06) rp = 1.0e-06;
        double dot = dot3(d0, d1);
        double cosine = dot / sqrt(d0[3] * d1[3]);
        double angle;
        if (cosine >= 1.0) angle = 0.0;
        else if (cosine <= -1.0) angle = M_PI;
        else angle = acos(cosine);
        double di = angle - t0;
        double dEdR = k * di;
        e_angle += 0.5 * k * di * di;
        if (f) {
            double termA = dEdR / (d0[3] * rp);
            double termC = -dEdR / (d1[3] * rp);
            double c0[3], c2[3];
            cross3(d0, p, c0);
            cross3(d1, p, c2);
            for (int c = 0; c < 3; ++c) {
                double fa = termA * c0[c], fc = termC * c2[c];
                f[3 * ia + c] += fa;
                f[3 * ib + c] += -1.0 * (fa + fc);
                f[3 * ic + c] += fc;
            }
        }
    }

    /* ---- PeriodicTorsionForce (ReferenceProperDihedralBond) ---- */
    for (int t = 0; t < s->n_torsions; ++t) {
        const int *ix = s->tors_idx + 4 * t;
        double phase = s->tors_par[2 * t], k = s->tors_par[2 * t + 1];
        double n = (double)s->tors_per[t];
        double d0[5], d1[5], d2[5], c1[3], c2[3];
        delta_r(x + 3 * ix[1], x + 3 * ix[0], d0); /* A - B */
        delta_r(x + 3 * ix[1], x + 3 * ix[2], d1); /* C - B */
        delta_r(x + 3 * ix[3], x + 3 * ix[2], d2); /* C - D */
        cross3(d0, d1, c1);
        cross3(d1, d2, c2);
        double phi = angle_between(c1, c2);
        double sgn = dot3(d0, c2);
        if (sgn < 0.0) phi = -phi;
        double da = n * phi - phase;
        double dEdA = -k * n * sin(da);
        e_tors += k * (1.0 + cos(da));
        if (f) {
            double nc1 = dot3(c1, c1), nc2 = dot3(c2, c2);
            double nbc = d1[4];
            double ff0 = (-dEdA * nbc) / nc1;
            double ff3 = (dEdA * nbc) / nc2;
            double ff1 = dot3(d0, d1) / d1[3];
            double ff2 = dot3(d2, d1) / d1[3];
            for (int c = 0; c < 3; ++c) {
                double f0 = ff0 * c1[c];
                double f3 = ff3 * c2[c];
                double sv = ff1 * f0 - ff2 * f3;
                double f1 = f0 - sv;
                double f2 = f3 + sv;
                f[3 * ix[0] + c] += f0;
                f[3 * ix[1] + c] -= f1;
                f[3 * ix[2] + c] -= f2;
                f[3 * ix[3] + c] += f3;
            }
        }
    }

    /* ---- NonbondedForce: particle pairs (ReferenceLJCoulombIxn::calculateOneIxn) ---- */
    {
        /* exclusion table from the exception list (every exception removes the plain pair) */
        unsigned char *excl = (unsigned char *)calloc((size_t)N * N, 1);
        for (int e = 0; e < s->n_exceptions; ++e) {
            int i = s->exc_idx[2 * e], j = s->exc_idx[2 * e + 1];
            excl[i * N + j] = excl[j * N + i] = 1;
        }
        const int cutoff = (s->method == PMO_CUTOFF_NONPERIODIC);
        double krf = 0.0, crf = 0.0;
        if (cutoff) {
            double rc = s->cutoff, ed = s->rf_dielectric;
            krf = pow(rc, -3.0) * (ed - 1.0) / (2.0 * ed + 1.0);
            crf = (1.0 / rc) * (3.0 * ed) / (2.0 * ed + 1.0);
        }
        for (int i = 0; i < N; ++i) {
            double qi = s->particle_par[3 * i], hsi = 0.5 * s->particle_par[3 * i + 1];
            double ei = 2.0 * sqrt(s->particle_par[3 * i + 2]);
            for (int j = i + 1; j < N; ++j) {
                if (excl[i * N + j]) continue;
                double d[5];
                delta_r(x + 3 * j, x + 3 * i, d); /* i - j */
                if (cutoff && d[4] >= s->cutoff) continue;
                double qj = s->particle_par[3 * j], hsj = 0.5 * s->particle_par[3 * j + 1];
                double ej = 2.0 * sqrt(s->particle_par[3 * j + 2]);
                double inverseR = 1.0 / d[4];
                double sig = hsi + hsj;
                double sig2 = inverseR * sig;
                sig2 *= sig2;
                double sig6 = sig2 * sig2 * sig2;
                double eps = ei * ej;
                double dEdR = eps * (12.0 * sig6 - 6.0) * sig6;
                double energy = eps * (sig6 - 1.0) * sig6;
                if (cutoff) {
                    dEdR += s->k_coul * qi * qj * (inverseR - 2.0 * krf * d[3]);
                    energy += s->k_coul * qi * qj * (inverseR + krf * d[3] - crf);
                } else {
                    dEdR += s->k_coul * qi * qj * inverseR;
                    energy += s->k_coul * qi * qj * inverseR;
                }
                dEdR *= inverseR * inverseR;
                if (f) for (int c = 0; c < 3; ++c) { double fc = dEdR * d[c]; f[3 * i + c] += fc; f[3 * j + c] -= fc; }
                e_nb += energy;
            }
        }
        free(excl);
    }

    /* ---- NonbondedForce: exceptions (ReferenceLJCoulomb14) ---- */
    for (int e = 0; e < s->n_exceptions; ++e) {
        int i = s->exc_idx[2 * e], j = s->exc_idx[2 * e + 1];
        double qq = s->exc_par[3 * e], sigma = s->exc_par[3 * e + 1], eps4 = 4.0 * s->exc_par[3 * e + 2];
        if (qq == 0.0 && eps4 == 0.0) continue; /* plain exclusion */
        double d[5];
        delta_r(x + 3 * j, x + 3 * i, d); /* i - j */
        double inverseR = 1.0 / d[4];
        double sig2 = inverseR * sigma;
        sig2 *= sig2;
        double sig6 = sig2 * sig2 * sig2;
        double dEdR = eps4 * (12.0 * sig6 - 6.0) * sig6;
        dEdR += s->k_coul * qq * inverseR;
        dEdR *= inverseR * inverseR;
        double energy = eps4 * (sig6 - 1.0) * sig6;
        energy += s->k_coul * qq * inverseR;
        if (f) for (int c = 0; c < 3; ++c) { double fc = dEdR * d[c]; f[3 * i + c] += fc; f[3 * j + c] -= fc; }
        e_nb += energy;
    }

    return ((e_bond + e_angle) + e_tors) + e_nb;
}

/*
 * Energies [S] and (optionally) forces [S][N][3] for S conformations (coords [S][N][3], nm).
 * n_threads > 1 uses OpenMP over conformations (the analogue of the reference's fork pool,
 * ParaMol/Objective_function/objective_function.py:268-311).  Returns 0.
 */
int pmo_energy_forces(int n_atoms,
                      int n_bonds, const int *bond_idx, const double *bond_par,
                      int n_angles, const int *angle_idx, const double *angle_par,
                      int n_torsions, const int *tors_idx, const int *tors_per, const double *tors_par,
                      const double *particle_par,
                      int n_exceptions, const int *exc_idx, const double *exc_par,
                      int method, double cutoff, double rf_dielectric, double k_coul,
                      long n_conf, const double *coords, double *energies, double *forces, int n_threads) {
    pmo_system s;
    s.n_atoms = n_atoms;
    s.n_bonds = n_bonds; s.bond_idx = bond_idx; s.bond_par = bond_par;
    s.n_angles = n_angles; s.angle_idx = angle_idx; s.angle_par = angle_par;
    s.n_torsions = n_torsions; s.tors_idx = tors_idx; s.tors_per = tors_per; s.tors_par = tors_par;
    s.particle_par = particle_par;
    s.n_exceptions = n_exceptions; s.exc_idx = exc_idx; s.exc_par = exc_par;
    s.method = method; s.cutoff = cutoff; s.rf_dielectric = rf_dielectric; s.k_coul = k_coul;
    const size_t stride = (size_t)3 * n_atoms;
    if (forces) memset(forces, 0, sizeof(double) * stride * (size_t)n_conf);
#ifdef _OPENMP
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(static) num_threads(n_threads)
#endif
    for (long c = 0; c < n_conf; ++c) {
        double e = eval_one(&s, coords + stride * c, forces ? forces + stride * c : NULL);
        if (energies) energies[c] = e;
    }
    (void)n_threads;
    return 0;
}

int pmo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Internal coordinates, ParaMol/Utils/geometry.py:4-107 (used by LinearLeastSquare._calculate_a,
 * ParaMol/MM_engines/least_square.py:492-645).
 * ---------------------------------------------------------------------------------------------- */
double pmo_distance(const double *p0, const double *p1) {
    double d[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    return sqrt(dot3(d, d));
}

/* calculate_angle(v1, v2): arccos(clip(unit(v1).unit(v2), -1, 1)) */
double pmo_angle(const double *v1, const double *v2) {
    double n1 = sqrt(dot3(v1, v1)), n2 = sqrt(dot3(v2, v2));
    double u1[3] = {v1[0] / n1, v1[1] / n1, v1[2] / n1};
    double u2[3] = {v2[0] / n2, v2[1] / n2, v2[2] / n2};
    double c = dot3(u1, u2);
    if (c > 1.0) c = 1.0;
    if (c < -1.0) c = -1.0;
    return acos(c);
}

/* calculate_dihedral(p0..p3): "praxeolitic" atan2 formula */
double pmo_dihedral(const double *p0, const double *p1, const double *p2, const double *p3) {
    double b0[3], b1[3], b2[3], v[3], w[3], bxv[3];
    for (int c = 0; c < 3; ++c) { b0[c] = -1.0 * (p1[c] - p0[c]); b1[c] = p2[c] - p1[c]; b2[c] = p3[c] - p2[c]; }
    double n1 = sqrt(dot3(b1, b1));
    for (int c = 0; c < 3; ++c) b1[c] /= n1;
    double d0 = dot3(b0, b1), d2 = dot3(b2, b1);
    for (int c = 0; c < 3; ++c) { v[c] = b0[c] - d0 * b1[c]; w[c] = b2[c] - d2 * b1[c]; }
    double xx = dot3(v, w);
    cross3(b1, v, bxv);
    double yy = dot3(bxv, w);
    return atan2(yy, xx);
}

/*
 * LLS design-matrix internal coordinates: for every conformation and every listed term the raw internal
 * coordinate (distance / angle / dihedral).  kind: 0 bond, 1 angle, 2 torsion.  out [S][n_terms].
 */
int pmo_internal_coordinates(int n_atoms, long n_conf, const double *coords,
                             int n_terms, const int *kind, const int *atoms /* [n_terms][4] */, double *out) {
    const size_t stride = (size_t)3 * n_atoms;
    for (long c = 0; c < n_conf; ++c) {
        const double *x = coords + stride * c;
        for (int t = 0; t < n_terms; ++t) {
            const int *a = atoms + 4 * t;
            double v;
            if (kind[t] == 0) v = pmo_distance(x + 3 * a[0], x + 3 * a[1]);
            else if (kind[t] == 1) {
                double v1[3], v2[3];
                for (int k = 0; k < 3; ++k) { v1[k] = x[3 * a[0] + k] - x[3 * a[1] + k]; v2[k] = x[3 * a[2] + k] - x[3 * a[1] + k]; }
                v = pmo_angle(v1, v2);
            } else v = pmo_dihedral(x + 3 * a[0], x + 3 * a[1], x + 3 * a[2], x + 3 * a[3]);
            out[(size_t)c * n_terms + t] = v;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * RESP sums, ParaMol/MM_engines/resp.py:99-134 (inverse distances), :412-421 (A_aux), :490-496 (B_aux),
 * for ONE conformation with G grid points:  inv_r[j][i] = 1/|x_j - g_i|,
 * A[j][k] = sum_i inv_r[j][i] inv_r[k][i],  B[j] = sum_i V[i] inv_r[j][i]  (summed in grid order).
 * ---------------------------------------------------------------------------------------------- */
int pmo_resp_sums(int n_atoms, const double *x /*[N][3]*/, long n_grid, const double *grid /*[G][3]*/,
                  const double *v /*[G]*/, double *inv_r /*[N][G] or NULL*/, double *A /*[N][N]*/, double *B /*[N]*/) {
    double *r = inv_r ? inv_r : (double *)malloc(sizeof(double) * (size_t)n_atoms * n_grid);
    for (int j = 0; j < n_atoms; ++j)
        for (long i = 0; i < n_grid; ++i) {
            double d[3] = {x[3 * j] - grid[3 * i], x[3 * j + 1] - grid[3 * i + 1], x[3 * j + 2] - grid[3 * i + 2]};
            r[(size_t)j * n_grid + i] = 1.0 / sqrt(dot3(d, d));
        }
    for (int j = 0; j < n_atoms; ++j) {
        for (int k = j; k < n_atoms; ++k) {
            double acc = 0.0;
            for (long i = 0; i < n_grid; ++i) acc = acc + r[(size_t)j * n_grid + i] * r[(size_t)k * n_grid + i];
            A[j * n_atoms + k] = acc;
            A[k * n_atoms + j] = acc;
        }
        double b = 0.0;
        for (long i = 0; i < n_grid; ++i) b = b + v[i] * r[(size_t)j * n_grid + i];
        B[j] = b;
    }
    if (!inv_r) free(r);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Per-conformation force residual  r_s = sum_j dF_sj^T M_j dF_sj ,  dF = F_mm - F_ref
 * (inner loops of ForceProperty.calculate_property, ParaMol/Objective_function/Properties/force_property.py:85-117;
 *  M_j = inverse covariance of the QM forces for "components", I / var_j for "norm").  OpenMP over conformations.
 * ---------------------------------------------------------------------------------------------- */
int pmo_force_residuals(int n_atoms, long n_conf, const double *fmm, const double *fref, const double *metric /*[N][9]*/,
                        double *out, int n_threads) {
    const size_t stride = (size_t)3 * n_atoms;
#ifdef _OPENMP
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(static) num_threads(n_threads)
#endif
    for (long s = 0; s < n_conf; ++s) {
        const double *a = fmm + stride * s, *b = fref + stride * s;
        double acc = 0.0;
        for (int j = 0; j < n_atoms; ++j) {
            const double d[3] = {a[3 * j] - b[3 * j], a[3 * j + 1] - b[3 * j + 1], a[3 * j + 2] - b[3 * j + 2]};
            const double *m = metric + 9 * j;
            double t[3];
            for (int c = 0; c < 3; ++c) t[c] = d[0] * m[c] + d[1] * m[3 + c] + d[2] * m[6 + c];
            acc += t[0] * d[0] + t[1] * d[1] + t[2] * d[2];
        }
        out[s] = acc;
    }
    (void)n_threads;
    return 0;
}
