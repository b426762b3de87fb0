# -*- coding: utf-8 -*-
"""
oracle.py -- CPU oracle for the objective-evaluation path (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package ``paramol_b200`` never does.

* E / F for conformations: ``oracle.c`` (FP64 restatement of the OpenMM Reference-platform formulas that
  the reference calls through ``ParaMol/MM_engines/openmm.py:448-502``), loaded with ctypes.
* Everything ParaMol itself computes in Python/numpy is restated here in numpy, function by function,
  with the reference file:line each one follows.

Parity status: PINNED.  ``tests/test_oracle.py`` checks this module against every golden value of the
reference's own test-suite for this path (``tests/golden/reference_goldens.json``).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# CODATA-2018 values as used by simtk.unit in openmm>=7.5 (BOLTZMANN_CONSTANT_kB, AVOGADRO_CONSTANT_NA)
KB_J_PER_K = 1.380649e-23
NA_PER_MOL = 6.02214076e23
ONE_4PI_EPS0 = 138.935456


def build(force=False):
    """Compile oracle.c -> liboracle.so next to it (gcc, OpenMP if available)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if not force and os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(src):
        return so
    base = ["-O2", "-fPIC", "-ffp-contract=off", "-std=gnu11", "-shared", "-o", so, src, "-lm"]
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + omp + base, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                return so
            except Exception:
                continue
    raise RuntimeError("could not compile oracle/oracle.c")


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = ctypes.CDLL(so)
        ip = ctypes.POINTER(ctypes.c_int)
        dp = ctypes.POINTER(ctypes.c_double)
        L.pmo_energy_forces.restype = ctypes.c_int
        L.pmo_energy_forces.argtypes = [ctypes.c_int,
                                        ctypes.c_int, ip, dp,
                                        ctypes.c_int, ip, dp,
                                        ctypes.c_int, ip, ip, dp,
                                        dp,
                                        ctypes.c_int, ip, dp,
                                        ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_long, dp, dp, dp, ctypes.c_int]
        L.pmo_max_threads.restype = ctypes.c_int
        L.pmo_internal_coordinates.restype = ctypes.c_int
        L.pmo_internal_coordinates.argtypes = [ctypes.c_int, ctypes.c_long, dp, ctypes.c_int, ip, ip, dp]
        L.pmo_force_residuals.restype = ctypes.c_int
        L.pmo_force_residuals.argtypes = [ctypes.c_int, ctypes.c_long, dp, dp, dp, dp, ctypes.c_int]
        L.pmo_resp_sums.restype = ctypes.c_int
        L.pmo_resp_sums.argtypes = [ctypes.c_int, dp, ctypes.c_long, dp, dp, dp, dp, dp]
        _LIB = L
    return _LIB


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# --------------------------------------------------------------------------------------------------
# E / F  (system.py:324-354 loops over openmm.py:448-502)
# --------------------------------------------------------------------------------------------------
def energies_forces(system, coords, want_forces=True, n_threads=1, k_coul=None):
    """
    MM energies [S] and forces [S,N,3] of ``coords`` [S,N,3] (nm) for an MMSystem-like object
    (attributes bond_idx/bond_par/angle_idx/angle_par/torsion_idx/torsion_per/torsion_par/particle_par/
    exception_idx/exception_par/nonbonded_method/cutoff/rf_dielectric/one_4pi_eps0).
    """
    L = lib()
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    if coords.ndim == 2:
        coords = coords[None]
    S, N, _ = coords.shape
    assert N == system.n_atoms
    keep = []

    def I(a):
        a, p = _i(a)
        keep.append(a)
        return p

    def D(a):
        a, p = _d(a)
        keep.append(a)
        return p

    energies = np.zeros(S)
    forces = np.zeros((S, N, 3)) if want_forces else None
    method = 0 if system.nonbonded_method == "NoCutoff" else 1
    kc = float(system.one_4pi_eps0 if k_coul is None else k_coul)
    rc = L.pmo_energy_forces(N,
                             len(system.bond_idx), I(system.bond_idx), D(system.bond_par),
                             len(system.angle_idx), I(system.angle_idx), D(system.angle_par),
                             len(system.torsion_idx), I(system.torsion_idx), I(system.torsion_per), D(system.torsion_par),
                             D(system.particle_par),
                             len(system.exception_idx), I(system.exception_idx), D(system.exception_par),
                             method, float(system.cutoff), float(system.rf_dielectric), kc,
                             S, coords.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                             energies.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                             forces.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if want_forces else None,
                             int(n_threads))
    assert rc == 0
    return energies, forces


def max_threads():
    return int(lib().pmo_max_threads())


def energy_single_python(system, x, k_coul=ONE_4PI_EPS0):
    """Pure-Python/numpy energy of ONE conformation (slow; cross-checks oracle.c on small cases)."""
    e = 0.0
    for (a, b), (r0, k) in zip(system.bond_idx, system.bond_par):
        r = np.linalg.norm(x[b] - x[a])
        e += 0.5 * k * (r - r0) ** 2
    ea = 0.0
    for (a, b, c), (t0, k) in zip(system.angle_idx, system.angle_par):
        v1 = x[b] - x[a]
        v2 = x[b] - x[c]
        cs = np.dot(v1, v2) / np.sqrt(np.dot(v1, v1) * np.dot(v2, v2))
        th = np.arccos(np.clip(cs, -1, 1))
        ea += 0.5 * k * (th - t0) ** 2
    et = 0.0
    for (a, b, c, d), n, (ph, k) in zip(system.torsion_idx, system.torsion_per, system.torsion_par):
        d0 = x[a] - x[b]
        d1 = x[c] - x[b]
        d2 = x[c] - x[d]
        c1 = np.cross(d0, d1)
        c2 = np.cross(d1, d2)
        cs = np.dot(c1, c2) / np.sqrt(np.dot(c1, c1) * np.dot(c2, c2))
        phi = np.arccos(np.clip(cs, -1, 1))
        if np.dot(d0, c2) < 0:
            phi = -phi
        et += k * (1.0 + np.cos(n * phi - ph))
    enb = 0.0
    excl = {(min(i, j), max(i, j)) for i, j in system.exception_idx}
    q, sig, eps = system.particle_par.T
    N = system.n_atoms
    for i in range(N):
        for j in range(i + 1, N):
            if (i, j) in excl:
                continue
            r = np.linalg.norm(x[i] - x[j])
            s6 = (0.5 * (sig[i] + sig[j]) / r) ** 6
            enb += 4 * np.sqrt(eps[i] * eps[j]) * (s6 * s6 - s6) + k_coul * q[i] * q[j] / r
    for (i, j), (qq, s, ee) in zip(system.exception_idx, system.exception_par):
        r = np.linalg.norm(x[i] - x[j])
        s6 = (s / r) ** 6
        enb += 4 * ee * (s6 * s6 - s6) + k_coul * qq / r
    return ((e + ea) + et) + enb


# --------------------------------------------------------------------------------------------------
# Conformation weights (system.py:165-248)
# --------------------------------------------------------------------------------------------------
def beta_kj_mol(temperature_kelvin):
    """1/(kB T NA) in mol/kJ (system.py:211-212, 229-230)."""
    return 1.0 / (KB_J_PER_K * float(temperature_kelvin) * NA_PER_MOL / 1000.0)


def conformation_weights(method, n_structures, ref_energies=None, emm=None, temperature=300.0, manual=None):
    method = method.upper()
    if method == "UNIFORM":
        w = np.ones(n_structures)
        return w / np.sum(w)
    if method == "NON_BOLTZMANN":
        diff = np.asarray(emm) - np.asarray(ref_energies)
        diff = diff - np.mean(diff)
        beta = beta_kj_mol(temperature)
        with np.errstate(all="raise"):
            try:
                w = np.exp(-beta * diff)
                return w / np.sum(w)
            except FloatingPointError:
                return np.ones(n_structures) / n_structures
    if method == "BOLTZMANN":
        beta = beta_kj_mol(temperature)
        ref = np.asarray(ref_energies)
        with np.errstate(all="raise"):
            try:
                w = np.exp(-np.asarray(ref - np.mean(ref)) * beta)
                return w / np.sum(w)
            except FloatingPointError:
                return np.ones(n_structures) / n_structures
    if method == "MANUAL":
        return np.asarray(manual, dtype=np.float64)
    raise NotImplementedError(method)


# --------------------------------------------------------------------------------------------------
# Properties (Objective_function/Properties/*.py)
# --------------------------------------------------------------------------------------------------
def energy_property(ref_energies, emm, weights, wham_weights=1.0):
    """EnergyProperty.calculate_property for one system, energy_property.py:99-108 (variance :131)."""
    ref = np.asarray(ref_energies)
    var = np.var(ref)
    diff = ref - np.asarray(emm)
    avg = np.mean(diff)
    obj = diff - avg
    obj = obj * obj
    obj = np.sum(weights * wham_weights * obj)
    return obj / var


def force_variance(ref_forces):
    """ForceProperty.calculate_variance, force_property.py:192-197."""
    return np.var(np.linalg.norm(np.asarray(ref_forces), axis=2), axis=0)


def force_inverse_covariance(ref_forces):
    """ForceProperty.calculate_inverse_covariance_qm_forces, force_property.py:126-173."""
    ref = np.asarray(ref_forces)
    S, N, _ = ref.shape
    out = np.zeros((N, 3, 3))
    for i in range(N):
        avg = np.zeros((3, 3))
        for j in range(S):
            avg = avg + np.outer(ref[j, i, :], ref[j, i, :])
        avg = avg / S
        if abs(np.linalg.det(avg)) < 1e-8:
            avg = np.identity(3)
        else:
            avg = np.linalg.inv(avg)
        out[i] = avg
    return out


def force_property(ref_forces, fmm, weights, term_type="components", wham_weights=1.0):
    """ForceProperty.calculate_property for one system, force_property.py:85-117."""
    ref = np.asarray(ref_forces)
    fmm = np.asarray(fmm)
    S, N, _ = ref.shape
    diff = fmm - ref
    if term_type.lower() == "components":
        inv_cov = force_inverse_covariance(ref)
        # sum_j diff_ij^T C_j^-1 diff_ij, accumulated over j in atom order like the reference loop (:91-95)
        err = np.zeros(S)
        for j in range(N):
            err = err + np.einsum("sa,ab,sb->s", diff[:, j, :], inv_cov[j], diff[:, j, :])
        return np.sum(weights * wham_weights * err) / (3 * N)
    elif term_type.lower() == "norm":
        var = force_variance(ref)
        obj = np.sum(np.power(diff, 2), axis=2)
        obj = obj / var
        obj = np.sum(obj, 1)
        obj = np.sum(weights * wham_weights * obj)
        return obj / (3 * N)
    raise NotImplementedError(term_type)


def force_residuals(fmm, ref_forces, metric, n_threads=1):
    """r_s = sum_j dF^T M_j dF per conformation (C, OpenMP); metric [N,3,3] or [N,9]."""
    fmm_a, fp = _d(fmm)
    ref_a, rp = _d(ref_forces)
    m_a, mp = _d(np.asarray(metric).reshape(-1, 9))
    S, N, _ = fmm_a.shape
    out = np.zeros(S)
    lib().pmo_force_residuals(N, S, fp, rp, mp, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(n_threads))
    return out


def regularization(x, x0, prior_widths, method="L2", a=1.0, b=0.1):
    """Regularization.calculate_property, regularization.py:137-228."""
    x = np.asarray(x, dtype=np.float64)
    m = method.upper()
    if m == "L2":
        diff = (x - x0) / prior_widths
        return a * np.sum(np.power(diff, 2))
    if m == "L1":
        diff = (x - x0) / prior_widths
        return a * np.sum(np.abs(diff))
    if m == "HYPERBOLIC":
        return a * np.sum((x ** 2 + b ** 2) ** (1 / 2.) - b)
    raise NotImplementedError(method)


def esp_ensemble(charges, inv_rij):
    """ParaMolSystem.get_esp_ensemble, system.py:356-382: V[m][i] = sum_j q_j / r_ij."""
    q = np.asarray(charges, dtype=np.float64)
    return [np.asarray(r).T @ q for r in inv_rij]


def esp_property(ref_esp, esp, weights):
    """ESPProperty.calculate_property for one system, esp_property.py:89-104 (variance not applied)."""
    tmp = np.array([np.sum((np.asarray(ref_esp[m]) - np.asarray(esp[m])) ** 2) for m in range(len(ref_esp))])
    return np.sum(tmp * weights, axis=0)


def inverse_distances(coords, grids):
    """RESP.calculate_inverse_distances, resp.py:99-134: list over conformations of [N, G_m]."""
    out = []
    for x, g in zip(coords, grids):
        x = np.asarray(x)
        g = np.asarray(g)
        out.append(1.0 / np.linalg.norm(x[:, None, :] - g[None, :, :], axis=2))
    return out


# --------------------------------------------------------------------------------------------------
# Whole objective for one system (objective_function.py:162-226), given MM system parameters.
# --------------------------------------------------------------------------------------------------
def objective(system, coords, ref_energies=None, ref_forces=None, include_energy=True, include_force=True,
              force_term_type="components", weight_energy=1.0, weight_force=1.0, weighting_method="UNIFORM",
              temperature=300.0, manual_weights=None, reg=None, n_threads=1):
    """
    Returns (objective, dict of parts).  ``reg`` = None or dict(x=, x0=, prior_widths=, method=, a=, b=, weight=).
    """
    S = len(coords)
    emm, fmm = energies_forces(system, coords, want_forces=include_force, n_threads=n_threads)
    w = conformation_weights(weighting_method, S, ref_energies=ref_energies, emm=emm, temperature=temperature,
                             manual=manual_weights)
    parts = {"emm": emm, "fmm": fmm, "weights": w}
    total = 0.0
    if include_energy:
        parts["ENERGY"] = energy_property(ref_energies, emm, w)
        total += weight_energy * parts["ENERGY"]
    if include_force:
        parts["FORCE"] = force_property(ref_forces, fmm, w, force_term_type)
        total += weight_force * parts["FORCE"]
    if reg is not None:
        parts["REGULARIZATION"] = regularization(reg["x"], reg["x0"], reg["prior_widths"], reg.get("method", "L2"),
                                                 reg.get("a", 1.0), reg.get("b", 0.1))
        total += reg.get("weight", 1.0) * parts["REGULARIZATION"]
    if np.isnan(total):
        total = 1e16
    return total, parts


# --------------------------------------------------------------------------------------------------
# Internal coordinates (Utils/geometry.py) through oracle.c
# --------------------------------------------------------------------------------------------------
def internal_coordinates(coords, kinds, atoms):
    """out[S, n_terms]; kinds 0/1/2 = bond/angle/torsion; atoms int [n_terms,4] (unused slots ignored)."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    S, N, _ = coords.shape
    kinds_a, kp = _i(kinds)
    atoms_a, ap = _i(atoms)
    out = np.zeros((S, len(kinds_a)))
    lib().pmo_internal_coordinates(N, S, coords.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(kinds_a), kp, ap,
                                   out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def resp_sums(x, grid, v):
    """(inv_r [N,G], A [N,N], B [N]) for one conformation: resp.py:122-130, 412-421, 490-496."""
    x_a, xp = _d(x)
    g_a, gp = _d(grid)
    v_a, vp = _d(v)
    N = x_a.shape[0]
    G = g_a.shape[0]
    inv_r = np.zeros((N, G))
    A = np.zeros((N, N))
    B = np.zeros(N)
    dp = ctypes.POINTER(ctypes.c_double)
    lib().pmo_resp_sums(N, xp, G, gp, vp, inv_r.ctypes.data_as(dp), A.ctypes.data_as(dp), B.ctypes.data_as(dp))
    return inv_r, A, B


# --------------------------------------------------------------------------------------------------
# LLS (MM_engines/least_square.py) -- numpy restatement, PINNED by tests/golden/lls_reference_fixtures.npz, which
# tests/golden/make_lls_fixtures.py produced by running the reference's own least_square.py (tests/test_lls_reference_fixtures.py)
# --------------------------------------------------------------------------------------------------
def lls_design_matrix(coords, parameter_space, alpha_bond=0.05):
    """LinearLeastSquare._calculate_a, least_square.py:469-649 (vectorised over conformations).
    Returns (A centred [S, M], p0 [M], keys, init [M], symmetry tags)."""
    X = np.asarray(coords, dtype=np.float64)
    cols, p0, keys, init, symm = [], [], [], [], []
    spec = {"bond_k": ("bond_eq", 0), "angle_k": ("angle_eq", 1), "torsion_k": ("torsion_phase", 2)}
    for parameter in parameter_space.optimizable_parameters:
        t = parameter.ff_term
        if parameter.param_key in spec:
            partner_key, kind = spec[parameter.param_key]
            atoms = (list(t.atoms) + [0, 0, 0])[:4]
            q = internal_coordinates(X, [kind], [atoms])[:, 0]
            partner = t.parameters[partner_key]
            if partner.optimize:
                if kind == 0:      # :517-518
                    centres = [partner.value * (1 - alpha_bond), partner.value * (1 + alpha_bond)]
                elif kind == 1:    # :566-567
                    centres = [q.min(), q.max()]
                else:              # :612-613
                    centres = [-np.pi / 4, np.pi / 4]
                tags = ["_x", "_y"]
            else:
                centres, tags = [partner.value], [""]
            for c, tag in zip(centres, tags):
                if kind == 2:
                    cols.append(1 + np.cos(t.parameters["torsion_periodicity"].value * q - c))
                else:
                    cols.append(0.5 * (q - c) * (q - c))
                p0.append(c)
                keys.append(parameter.param_key)
                init.append(parameter.value)
                symm.append(parameter.symmetry_group + tag)
        elif parameter.param_key not in ("torsion_phase", "bond_eq", "angle_eq"):
            raise NotImplementedError(parameter.param_key)
    A = np.stack(cols, axis=1)
    A = A - A.mean(axis=0)   # :640-642
    return A, np.asarray(p0), keys, np.asarray(init), symm


def lls_symmetry_rows(symm, n_parameters):
    """LinearLeastSquare._add_symmetries, least_square.py:293-337."""
    rows, covered = [], []
    for i, si in enumerate(symm):
        if si in covered or si in ("X_x", "X_y", "X"):
            continue
        for j in range(i + 1, len(symm)):
            if symm[j] == si:
                r = np.zeros(n_parameters)
                r[i], r[j] = 1.0, -1.0
                rows.append(r)
        covered.append(si)
    return np.asarray(rows).reshape(-1, n_parameters)


def lls_fit(A, B, weights, ref_energies, scaling_constants, prior_widths, init, symm, include_regularization=True,
            scaling_factor=1.0):
    """fit_parameters_lls, least_square.py:99-140: row weighting, preconditioning, regularisation rows (:253-291), symmetry rows,
    np.linalg.lstsq on the stacked system, scaling reverted.  Returns the parameter vector (``_parameters``)."""
    sc = np.asarray(scaling_constants, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float64) * np.ones(A.shape[0])
    var = np.var(ref_energies)
    Aw = A * np.sqrt(w)[:, None] / np.sqrt(var) / sc[None, :]
    Bw = B * np.sqrt(w) / np.sqrt(var)
    rows, rhs = [Aw], [Bw]
    if include_regularization:
        alpha = scaling_factor / sc
        rows.append(np.identity(len(sc)) / np.asarray(prior_widths)[None, :] * alpha)
        rhs.append(alpha * np.asarray(init))
    srows = lls_symmetry_rows(symm, len(sc))
    if len(srows):
        rows.append(srows)
        rhs.append(np.zeros(len(srows)))
    return np.linalg.lstsq(np.vstack(rows), np.concatenate(rhs), rcond=None)[0] / sc
