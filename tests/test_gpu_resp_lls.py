"""
GPU parity of the two linear fitting paths (run with ``-m gpu``).

RESP: the reference's own test body (ParaMol/Tests/Tasks/test_resp.py:78-151) -- explicit solver without restraint,
with L2 (a = 0.5) and with hyperbolic (a = 0.001, b = 0.1) restraints -- against the three golden charge vectors
(tolerance 1e-4 as in the reference), plus the assembled sums against the CPU oracle on seeded multi-conformation input.
LLS: the reference has no test for this path, so the device path is compared (a) with fixtures produced by running the reference's
own ParaMol/MM_engines/least_square.py in the build container (tests/golden/make_lls_fixtures.py) and (b) at larger sizes with the
oracle's numpy restatement of that file, which the same fixtures pin (tests/test_lls_reference_fixtures.py).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from paramol_b200.MM_engines.b200_engine import B200Engine
from paramol_b200.MM_engines.resp import RESP
from paramol_b200.System.system import ParaMolSystem
from paramol_b200.Parameter_space.parameter_space import ParameterSpace


@pytest.fixture()
def kwargs_dict(golden_dir):
    return {"topology_format": "AMBER", "crd_format": "AMBER",
            "top_file": os.path.join(golden_dir, "aniline.prmtop"), "crd_file": os.path.join(golden_dir, "aniline.inpcrd")}


def _aniline_esp(kwargs_dict, golden_dir):
    engine = B200Engine(True, **kwargs_dict)
    aniline = ParaMolSystem(name="aniline", engine=engine, n_atoms=14)
    d = np.load(os.path.join(golden_dir, "aniline_esp.npz"))
    aniline.ref_coordinates, aniline.ref_esp_grid, aniline.ref_esp = [d["conformation"]], [d["grid"]], [d["esp"]]
    aniline.ref_energies = [float(d["energy"])]
    assert aniline.ref_esp_grid[0].shape == (39946, 3) and len(aniline.ref_esp[0]) == 39946
    aniline.n_structures = len(aniline.ref_coordinates)
    return aniline


def _fit(aniline, include_reg, method="L2", scaling_factor=1.0, hyperbolic_beta=0.01):
    # RESPFitting.run_task(solver="explicit"), ParaMol/Tasks/resp_fitting.py:112-139
    aniline.engine.set_non_bonded_parameters_to_zero()
    aniline.force_field.create_force_field(opt_charges=True)
    parameter_space = ParameterSpace()
    parameter_space.get_optimizable_parameters([aniline], symmetry_constrained=False)
    aniline.resp_engine = RESP(0, include_reg, method=method, scaling_factor=scaling_factor, hyperbolic_beta=hyperbolic_beta,
                               weighting_method="uniform", weighting_temperature=300.0)
    aniline.resp_engine.set_initial_charges(aniline.force_field.force_field)
    aniline.resp_engine.set_charges(aniline.force_field.force_field)
    aniline.resp_engine.calculate_inverse_distances(aniline)
    aniline.resp_engine.set_symmetry_constraints(aniline, True)
    aniline.resp_engine.set_not_optimize_atoms(aniline)
    charges = aniline.resp_engine.fit_resp_charges_explicitly(aniline, 1e-8, 1000)
    charges_to_update = [charges[i] for i in range(len(charges)) if i not in aniline.resp_engine._not_optimized.keys()]
    parameter_space.update_systems([aniline], charges_to_update, symmetry_constrained=False)
    return np.asarray(parameter_space.optimizable_parameters_values)[:14]


def test_resp_explicit_goldens(kwargs_dict, golden_dir, goldens):
    aniline = _aniline_esp(kwargs_dict, golden_dir)
    charges = _fit(aniline, False)
    assert np.abs(charges - np.asarray(goldens["test_resp.py::test_resp::charges_to_compare"])).max() < 1e-4
    assert abs(charges.sum()) < 1e-10
    charges = _fit(aniline, True, "L2", 0.5)
    assert np.abs(charges - np.asarray(goldens["test_resp.py::test_resp::charges_to_compare#1"])).max() < 1e-4
    charges = _fit(aniline, True, "hyperbolic", 0.001, 0.1)
    assert np.abs(charges - np.asarray(goldens["test_resp.py::test_resp::charges_to_compare#2"])).max() < 1e-4


def test_resp_sums_potential_and_objective_vs_oracle(golden_dir):
    from oracle import oracle as O
    from paramol_b200.Utils.synthetic import make_ligand
    from paramol_b200.Objective_function.Properties.esp_property import ESPProperty
    rng = np.random.default_rng(11)
    mm, x0 = make_ligand(33)
    n_conf, sizes = 7, [120, 97, 128, 33, 128, 64, 1]
    X = [(x0 + rng.normal(0, 0.01, x0.shape)) * 18.897 for _ in range(n_conf)]
    grids, vrefs = [], []
    q_true = mm.particle_par[:, 0]
    for m in range(n_conf):
        g = rng.normal(0, 1.0, (sizes[m], 3))
        g = X[m][rng.integers(0, 33, sizes[m])] + 4.0 * g / np.linalg.norm(g, axis=1)[:, None]
        grids.append(g)
        vrefs.append((1.0 / np.linalg.norm(X[m][:, None, :] - g[None], axis=2)).T @ q_true + rng.normal(0, 1e-3, sizes[m]))
    engine = B200Engine(system=mm)
    system = ParaMolSystem("lig33", engine, 33, ref_coordinates=X, ref_esp=vrefs, ref_esp_grid=grids)
    system.force_field.create_force_field(opt_charges=True)
    resp = RESP(0, False, "L2", 1.0, 0.01, "uniform", 300.0)
    system.resp_engine = resp
    resp.set_charges(system.force_field.force_field)
    resp.calculate_inverse_distances(system)
    w = rng.uniform(0.1, 1.0, n_conf)
    A, B, C = resp._assemble(w)
    A_want, B_want, C_want = np.zeros((33, 33)), np.zeros(33), 0.0
    for m in range(n_conf):
        inv_r, Am, Bm = O.resp_sums(X[m], grids[m], vrefs[m])
        A_want += w[m] * Am
        B_want += w[m] * Bm
        C_want += w[m] * float(vrefs[m] @ vrefs[m])
        assert np.abs(resp.inv_rij[m] - inv_r).max() < 1e-13
    assert np.abs(A - A_want).max() <= 1e-12 * np.abs(A_want).max()
    assert np.abs(B - B_want).max() <= 1e-12 * np.abs(B_want).max()
    assert abs(C - C_want) <= 1e-12 * abs(C_want)
    # potential and ESP objective, against the oracle's restatement of system.py:356-382 / esp_property.py:89-104
    esp = system.get_esp_ensemble()
    want = O.esp_ensemble(resp.charges, O.inverse_distances(X, grids))
    for a, b in zip(esp, want):
        assert np.abs(a - b).max() < 1e-12
    system.weights = w
    val = resp.esp_objective(np.asarray(resp.charges), w)
    assert abs(val - O.esp_property(vrefs, want, w)) <= 1e-8 * abs(val)
    prop = ESPProperty([system], weight=1.0)
    prop.calculate_variance()
    assert abs(prop.calculate_property([esp]) - val) <= 1e-8 * abs(val)


def _numpy_lls(system, parameter_space, weights, include_reg, a_reg, alpha_bond=0.05):
    """The oracle's restatement of LinearLeastSquare._calculate_a (least_square.py:469-649), which tests/test_lls_reference_fixtures.py
    pins to fixtures produced by the reference's own code."""
    from oracle import oracle as O
    return O.lls_design_matrix(system.ref_coordinates, parameter_space, alpha_bond)


def test_lls_design_matrix_and_fit_vs_numpy(kwargs_dict, golden_dir):
    from oracle import oracle as O
    from paramol_b200.MM_engines.least_square import LinearLeastSquare
    from paramol_b200.Utils.synthetic import perturbed_conformations
    from paramol_b200.Utils.netcdf_lite import read_paramol_data
    X10 = read_paramol_data(os.path.join(golden_dir, "aniline_10_struct.nc"))[0]
    # the fit runs on the fused normal-equation kernels with iterative refinement, and on the orthogonal route (materialised
    # matrix, QR) when that is refused (forced in the third case): the result must not depend on which one ran
    for eq_opt, force_householder in ((False, False), (True, False), (True, True)):
        engine = B200Engine(True, **kwargs_dict)
        system = ParaMolSystem("aniline", engine, 14)
        system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True)
        if not eq_opt:
            for force in ("HarmonicBondForce", "HarmonicAngleForce", "PeriodicTorsionForce"):
                for term in system.force_field.force_field[force][0]:
                    for key, p in term.parameters.items():
                        if key in ("bond_eq", "angle_eq", "torsion_phase"):
                            p.optimize = 0
            system.force_field.create_force_field_optimizable()
        X = perturbed_conformations(X10, 600, sigma=0.004, seed=5)
        E, _ = O.energies_forces(engine.system, X)
        rng = np.random.default_rng(1)
        system.ref_coordinates, system.ref_energies = X, E + rng.normal(0, 2.0, len(E)) - 43000.0
        system.n_structures = len(X)
        ps = ParameterSpace()
        ps.get_optimizable_parameters([system], symmetry_constrained=False)
        lls = LinearLeastSquare(ps, True, method="L2", scaling_factor=1.0, hyperbolic_beta=0.01, weighting_method="uniform",
                                weighting_temperature=300.0)
        A_want, p0, keys, init, symm = _numpy_lls(system, ps, None, True, 1.0)
        A_dev = lls._calculate_a(system, 0.05, 0.05)
        assert A_dev.shape == A_want.shape
        assert np.abs(A_dev - A_want).max() <= 1e-10 * max(1.0, np.abs(A_want).max())
        assert np.allclose(lls._p0, p0, rtol=0, atol=1e-12) and lls._param_keys_list == keys and lls._param_symmetries_list == symm
        # full fit against the reference-structured numpy pipeline (lstsq on the full stacked matrix)
        x_before = np.asarray(ps.optimizable_parameters_values).copy()
        B = np.asarray(system.ref_energies) - lls._energies_with_zeroed_parameters(system)
        B = B - B.mean()
        w = np.ones(len(X)) / len(X)
        var = np.var(system.ref_energies)
        sc_dict, _ = ps.calculate_scaling_constants("default")
        sc = np.asarray([sc_dict[k] for k in keys])
        pw_dict, _ = ps.calculate_prior_widths("default")
        pw = np.asarray([pw_dict[k] for k in keys])
        Aw = A_want * np.sqrt(w)[:, None] / np.sqrt(var) / sc[None, :]
        Bw = B * np.sqrt(w) / np.sqrt(var)
        alpha = 1.0 / sc
        Afull = np.vstack([Aw, np.identity(len(sc)) / pw[None, :] * alpha[None, :]])
        Bfull = np.concatenate([Bw, alpha * init])
        want = np.linalg.lstsq(Afull, Bfull, rcond=None)[0] / sc
        assert np.allclose(np.asarray(ps.optimizable_parameters_values), x_before)
        if force_householder:     # refuse the refined normal-equation solve AND CholeskyQR2: the orthogonal fallback must agree
            lls.REFINE_MAX_STEPS = 0
            lls.CHOLESKY_QR_MAX_DEFECT = 0.0
        fitted = lls.fit_parameters_lls([system])
        got = lls._parameters
        assert lls.reduction == ("householder" if force_householder else "normal_equations_refined"), lls.reduction
        assert np.abs(got - want).max() <= 1e-7 * np.abs(want).max(), (lls.reduction, np.abs(got - want).max())
        assert len(fitted) == len(x_before)


@pytest.mark.parametrize("name", ["a10_eq_uniform", "a600_keq_uniform", "a600_eq_uniform", "a600_eq_boltzmann"])
def test_lls_matches_reference_fixtures(golden_dir, name):
    """The device LLS against fixtures produced by RUNNING the reference's own ``ParaMol/MM_engines/least_square.py``
    (``tests/golden/make_lls_fixtures.py``): design matrix, right-hand side, column bookkeeping, fitted parameter vector and
    the reconstructed optimizable values.  This pins SURVEY row a19 to the reference itself."""
    from lls_fixture_util import build_system, check_design, load_case
    from paramol_b200.MM_engines.least_square import LinearLeastSquare
    fx = load_case(golden_dir, name)
    system, ps = build_system(golden_dir, fx)
    lls = LinearLeastSquare(ps, True, method="L2", scaling_factor=1.0, hyperbolic_beta=0.01, weighting_method=str(fx["weighting"]),
                            weighting_temperature=300.0)
    A = lls._calculate_a(system, 0.05, 0.05)
    B = lls._calculate_b(system)
    assert np.array_equal(np.array(lls._param_keys_list), fx["keys"]) and np.array_equal(np.array(lls._param_symmetries_list), fx["symm"])
    assert np.abs(np.asarray(lls._p0) - fx["p0"]).max() <= 1e-12
    assert np.abs(lls.mm_energies_zero - fx["mm_energies_zero"]).max() <= 1e-9 * max(1.0, np.abs(fx["mm_energies_zero"]).max())
    check_design(fx, A, B, rtol_a=1e-10)
    fitted = np.asarray(lls.fit_parameters_lls([system]), dtype=np.float64)
    want = fx["parameters"]
    assert np.abs(lls._parameters - want).max() <= 1e-6 * np.abs(want).max(), (lls.reduction, np.abs(lls._parameters - want).max())
    assert np.abs(fitted - fx["fitted_values"]).max() <= 1e-6 * np.abs(fx["fitted_values"]).max()
    assert np.abs(np.asarray(system.weights) * np.ones(len(B)) - fx["weights"]).max() <= 1e-14


def test_lls_config4_shape_lig60_vs_numpy():
    """BASELINE.json config 4 in shape: the 60-atom ligand with every bonded force constant, equilibrium value and phase
    optimizable (742 design-matrix columns), 20 000 conformations: device design matrix and fitted parameters against the
    reference-structured numpy pipeline (lstsq on the full stacked matrix).  This is where the conditioning of the reduction
    (nearly collinear bracketing columns) matters."""
    from oracle import oracle as O
    from paramol_b200.MM_engines.least_square import LinearLeastSquare
    from paramol_b200.Utils.synthetic import make_ligand, perturbed_conformations
    mm, x0 = make_ligand(60)
    engine = B200Engine(system=mm)
    system = ParaMolSystem("lig60", engine, 60)
    system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True)
    n_conf = 20000
    X = perturbed_conformations(x0, n_conf, sigma=0.004, seed=17)
    E, _ = O.energies_forces(engine.system, X, n_threads=4)
    rng = np.random.default_rng(2)
    system.ref_coordinates, system.ref_energies = X, E + rng.normal(0, 2.0, len(E))
    system.n_structures = n_conf
    ps = ParameterSpace()
    ps.get_optimizable_parameters([system], symmetry_constrained=False)
    lls = LinearLeastSquare(ps, True, method="L2", scaling_factor=1.0, hyperbolic_beta=0.01, weighting_method="uniform",
                            weighting_temperature=300.0)
    A_want, p0, keys, init, symm = _numpy_lls(system, ps, None, True, 1.0)
    assert A_want.shape == (n_conf, 742)
    A_dev = lls._calculate_a(system, 0.05, 0.05)
    assert A_dev.shape == A_want.shape
    assert np.abs(A_dev - A_want).max() <= 1e-10 * max(1.0, np.abs(A_want).max())
    assert np.allclose(lls._p0, p0, rtol=0, atol=1e-12) and lls._param_keys_list == keys and lls._param_symmetries_list == symm
    B = np.asarray(system.ref_energies) - lls._energies_with_zeroed_parameters(system)
    B = B - B.mean()
    E_zero = O.energies_forces(_zeroed(engine.system), X, want_forces=False, n_threads=4)[0]
    assert np.abs(lls.mm_energies_zero - E_zero).max() <= 1e-9 * max(1.0, np.abs(E_zero).max())
    w = np.ones(n_conf) / n_conf
    var = np.var(system.ref_energies)
    sc_dict, _ = ps.calculate_scaling_constants("default")
    sc = np.asarray([sc_dict[k] for k in keys])
    pw_dict, _ = ps.calculate_prior_widths("default")
    pw = np.asarray([pw_dict[k] for k in keys])
    Aw = A_want * np.sqrt(w)[:, None] / np.sqrt(var) / sc[None, :]
    Bw = B * np.sqrt(w) / np.sqrt(var)
    alpha = 1.0 / sc
    rows = [Aw, np.identity(len(sc)) / pw[None, :] * alpha[None, :]]
    rhs = [Bw, alpha * init]
    sym_rows = []
    covered = []
    for i, si in enumerate(symm):
        if si in covered or si in ("X_x", "X_y", "X"):
            continue
        for j in range(i + 1, len(symm)):
            if symm[j] == si:
                r = np.zeros(len(sc))
                r[i], r[j] = 1.0, -1.0
                sym_rows.append(r)
        covered.append(si)
    if sym_rows:
        rows.append(np.asarray(sym_rows))
        rhs.append(np.zeros(len(sym_rows)))
    want = np.linalg.lstsq(np.vstack(rows), np.concatenate(rhs), rcond=None)[0] / sc
    # the fused kernels against the same numpy matrices: normal equations (DMMA), residual and gradient passes
    import torch
    plan = lls._device_plan(system, 0.05, 0.05)
    assert plan.n_cols == 742 and np.abs(np.asarray(lls._p0) - p0).max() <= 1e-12
    rs = torch.from_numpy(np.sqrt(w) / np.sqrt(var) * rng.uniform(0.5, 1.5, n_conf)).cuda()
    rs_h = rs.cpu().numpy()
    plan.set_columns(lls._p0, plan.means, 1.0 / sc)
    G = plan.normal(rs, torch.from_numpy(B).cuda()).cpu().numpy()
    Ab = np.hstack([A_want / sc[None, :], B[:, None]]) * rs_h[:, None]
    G_want = Ab.T @ Ab
    assert np.array_equal(G, G.T)
    assert np.abs(G - G_want).max() <= 1e-11 * np.abs(G_want).max(), np.abs(G - G_want).max() / np.abs(G_want).max()
    z = rng.normal(0, 1.0, 742) * sc
    e = plan.residual(torch.from_numpy(z).cuda(), torch.from_numpy(B).cuda())
    e_want = B - A_want @ z
    assert np.abs(e.cpu().numpy() - e_want).max() <= 1e-11 * np.abs(e_want).max()
    y = plan.colstats(e, rs * rs)[0].cpu().numpy()
    y_want = A_want.T @ (rs_h * rs_h * e_want)
    assert np.abs(y - y_want).max() <= 1e-10 * np.abs(y_want).max()
    lls.fit_parameters_lls([system])
    got = lls._parameters
    assert lls.reduction == "normal_equations_refined", lls.reduction
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max(), (lls.reduction, np.abs(got - want).max(), np.abs(want).max())


def _zeroed(mm):
    """Copy of an MMSystem with the bonded force constants, equilibrium values and phases zeroed (what update_systems(zeros) does
    for a bonded-only parameter space; abs() and the phase wrap leave zeros alone)."""
    z = mm.copy()
    z.bond_par[...] = 0.0
    z.angle_par[...] = 0.0
    z.torsion_par[...] = 0.0
    return z


def test_lls_reduction_cholesky_qr2_and_householder_fallback():
    """``LinearLeastSquare._reduce_tall``: (R, Q^T b) must reproduce the least-squares problem, || A x - b ||^2 =
    || R x - Q^T b ||^2 + const, i.e. R^T R = A^T A and R^T (Q^T b) = A^T b; a rank-deficient matrix (duplicated
    column) must fall back to the Householder factorisation."""
    import torch
    from paramol_b200.MM_engines.least_square import LinearLeastSquare
    lls = LinearLeastSquare.__new__(LinearLeastSquare)
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn((5000, 40), dtype=torch.float64, device="cuda", generator=g)
    A[:, 7] = A[:, 6] * (1.0 + 1e-4) + 1e-4 * A[:, 7]          # a nearly collinear pair, as the eq-bracketing columns are
    b = torch.randn(5000, dtype=torch.float64, device="cuda", generator=g)
    for case in ("well", "deficient"):
        if case == "deficient":
            A = torch.cat([A, A[:, 3:4]], dim=1)
        R, qtb = lls._reduce_tall(A, b)
        assert lls.reduction == ("cholesky_qr2" if case == "well" else "householder"), (case, lls.reduction)
        G, rhs = (A.T @ A).cpu().numpy(), (A.T @ b).cpu().numpy()
        R, qtb = R.cpu().numpy(), qtb.cpu().numpy()
        assert np.abs(R.T @ R - G).max() <= 1e-11 * np.abs(G).max()
        assert np.abs(R.T @ qtb - rhs).max() <= 1e-10 * np.abs(rhs).max()
        assert np.abs(np.tril(R, -1)).max() <= 1e-12 * np.abs(R).max()


@pytest.mark.parametrize("n_atoms,n_grid,n_conf", [(1, 5, 3), (3, 33, 4), (14, 600, 3), (31, 64, 5), (59, 77, 4), (60, 513, 3),
                                                   (60, 500, 256), (63, 40, 3), (64, 100, 2), (97, 130, 3), (131, 70, 2), (160, 45, 2)])
def test_resp_assemble_sizes_vs_numpy(n_atoms, n_grid, n_conf):
    """pm_resp_assemble through the C ABI for atom counts on both sides of every tiling boundary (micro-tile padding, number of
    warps, micro-tiles per thread) and grid sizes that leave partial tiles and partial chunks, against numpy."""
    import ctypes
    import torch
    from paramol_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(100 * n_atoms + n_grid)
    X = rng.normal(0, 3.0, (n_conf, n_atoms, 3))
    Gd = rng.normal(0, 3.0, (n_conf, n_grid, 3)) + 12.0 * np.sign(rng.normal(size=(n_conf, n_grid, 3)))
    V = rng.normal(0, 0.1, (n_conf, n_grid))
    w = rng.uniform(0.2, 1.0, n_conf)
    R = 1.0 / np.linalg.norm(X[:, :, None, :] - Gd[:, None, :, :], axis=3)      # [m][j][i]
    A_want = np.einsum("m,mji,mki->jk", w, R, R)
    B_want = np.einsum("m,mji,mi->j", w, R, V)
    C_want = float(np.einsum("m,mi,mi->", w, V, V))
    dev = torch.device("cuda:0")
    xd, gd, vd, wd = (torch.from_numpy(a).to(dev) for a in (X, Gd, V, w))
    out = torch.empty(n_atoms * n_atoms + n_atoms + 1, dtype=torch.float64, device=dev)
    scratch = torch.empty(int(L.pm_resp_scratch_doubles(n_atoms, n_conf, n_grid)), dtype=torch.float64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for wt, scale in ((wd, 1.0), (None, None)):
        _lib.check(L.pm_resp_assemble(n_atoms, xd.data_ptr(), gd.data_ptr(), vd.data_ptr(), wt.data_ptr() if wt is not None else None,
                                      n_conf, n_grid, out.data_ptr(), scratch.data_ptr(), stream), "pm_resp_assemble")
        host = out.cpu().numpy()
        A, B, C = host[:n_atoms * n_atoms].reshape(n_atoms, n_atoms), host[n_atoms * n_atoms:-1], host[-1]
        if wt is None:
            A_want = np.einsum("mji,mki->jk", R, R)
            B_want = np.einsum("mji,mi->j", R, V)
            C_want = float(np.einsum("mi,mi->", V, V))
        assert np.array_equal(A, A.T)
        assert np.abs(A - A_want).max() <= 1e-12 * np.abs(A_want).max()
        assert np.abs(B - B_want).max() <= 1e-12 * max(np.abs(B_want).max(), np.abs(A_want).max() * 0.1)
        assert abs(C - C_want) <= 1e-12 * abs(C_want)


def test_resp_assemble_rejects_oversize():
    import torch
    from paramol_b200 import _lib
    L = _lib.lib()
    d = torch.zeros(8, dtype=torch.float64, device="cuda:0")
    assert L.pm_resp_assemble(400, d.data_ptr(), d.data_ptr(), d.data_ptr(), None, 1, 1, d.data_ptr(), d.data_ptr(), None) != 0
    assert b"too many atoms" in L.pm_last_error()
