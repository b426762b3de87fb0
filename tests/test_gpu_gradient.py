"""
GPU tests of the analytic parameter gradient (``pm_grad`` / ``ObjectiveFunction.f_and_grad``; run with ``-m gpu``).

The reference has no analytic gradient (it differences ``ObjectiveFunction.f``: scipy ``jac="2-point"``,
ParaMol/Utils/settings.py:64; ParaMol/Optimizers/gradient_descent.py:106-121), so the oracle here is a central finite
difference: of the kernel's own energy / force-residual sums in slot space, of ``f_batch`` through the reference-facing
API, and of the CPU oracle's objective.  Tolerance: 2e-6 of the gradient's largest component (central differences with a
relative step of 1e-5 on O(1) objectives are good to ~1e-9 absolute).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from paramol_b200.System.system import ParaMolSystem
from paramol_b200.MM_engines.b200_engine import B200Engine
from paramol_b200.MM_engines.device import DeviceEnsemble
from paramol_b200.Objective_function.objective_function import ObjectiveFunction
from paramol_b200.Objective_function.Properties.regularization import Regularization
from paramol_b200.Objective_function.Properties.force_property import ForceProperty
from paramol_b200.Objective_function.Properties.energy_property import EnergyProperty
from paramol_b200.Parameter_space.parameter_space import ParameterSpace
from paramol_b200.Utils.synthetic import make_ligand

SETTINGS = {"parallel": False, "platform_name": "Reference", "weighting_temperature": 300.0, "checkpoint_freq": 10 ** 9}


@pytest.fixture()
def kwargs_dict(golden_dir):
    return {"topology_format": "AMBER", "crd_format": "AMBER",
            "top_file": os.path.join(golden_dir, "aniline.prmtop"), "crd_file": os.path.join(golden_dir, "aniline.inpcrd")}


@pytest.mark.parametrize("n_atoms,n_conf", [(12, 37), (24, 200), (41, 65), (60, 129)])
def test_term_gradient_vs_central_differences_in_slot_space(n_atoms, n_conf):
    """pm_grad + the adjoint of pm_prepare_params against central differences of sum_s a_s E_s + b_s fres_s evaluated by the
    forward kernel, for every slot kind (bond, angle, torsion, particle, exception), with a non-symmetric force metric and
    a conformation count that leaves partial tiles."""
    import torch
    mm, x0 = make_ligand(n_atoms)
    engine = B200Engine(system=mm)
    topo = engine.device_topology()
    rng = np.random.default_rng(3)
    X = x0[None] + rng.normal(0, 0.01, (n_conf,) + x0.shape)
    ens = DeviceEnsemble(topo, X, rng.normal(0, 50.0, X.shape), rng.normal(0, 5.0, n_conf))
    metric = np.tile(np.eye(3).ravel(), (n_atoms, 1)) * rng.uniform(0.5, 2.0, (n_atoms, 1))
    metric[:, 1] += 0.1
    metric[:, 3] += 0.05
    topo.set_force_metric(metric)
    base = engine.current_slots().copy()
    a = torch.from_numpy(rng.normal(0, 1.0, n_conf)).cuda()
    b = torch.from_numpy(rng.uniform(0.1, 1.0, n_conf)).cuda()

    def forward(rows):
        der = topo.prepare(torch.from_numpy(np.ascontiguousarray(rows)).cuda())
        emm, fres, _ = ens.evaluate(der, with_forces=True, want_fres=True)
        return ((emm * a[None]).sum(1) + (fres * b[None]).sum(1)).cpu().numpy()

    der = topo.prepare(torch.from_numpy(base[None].copy()).cuda())
    _, _, fmm = ens.evaluate(der, with_forces=True, want_fres=True, want_fmm=True)
    g = topo.term_gradient_to_slots(ens.term_gradient(der, a, b, fmm[0]).cpu().numpy(), base)
    g2 = topo.term_gradient_to_slots(ens.term_gradient(der, a, b, fmm[0]).cpu().numpy(), base)
    assert np.array_equal(g, g2)                       # deterministic: no atomics anywhere
    bounds = [0, topo.off_angle, topo.off_torsion, topo.off_particle, topo.off_exception, topo.n_slots]
    for lo, hi in zip(bounds[:-1], bounds[1:]):        # every kind is sampled
        idx = rng.choice(np.arange(lo, hi), size=min(hi - lo, 40), replace=False)
        steps = 1e-6 * np.maximum(1.0, np.abs(base[idx]))
        rows = np.repeat(base[None], 2 * len(idx), 0)
        rows[2 * np.arange(len(idx)), idx] += steps
        rows[2 * np.arange(len(idx)) + 1, idx] -= steps
        vals = forward(rows)
        fd = (vals[0::2] - vals[1::2]) / (2 * steps)
        # sigma derivatives are ~1e13 while bonded ones are ~1e5: compare within each kind
        assert np.abs(fd - g[idx]).max() <= 2e-5 * max(np.abs(g[lo:hi]).max(), 1e-30), (lo, hi)
    # energy-only and force-only calls add up to the combined one
    ge = ens.term_gradient(der, a, None, None).cpu().numpy().copy()
    gf = ens.term_gradient(der, None, b, fmm[0]).cpu().numpy().copy()
    both = ens.term_gradient(der, a, b, fmm[0]).cpu().numpy()
    assert np.abs(ge + gf - both).max() <= 1e-12 * np.abs(both).max()


def _aniline(kwargs_dict, golden_dir, **flags):
    engine = B200Engine(True, **kwargs_dict)
    system = ParaMolSystem(name="aniline", engine=engine, n_atoms=14)
    system.force_field.create_force_field(ff_file=None, **flags)
    system.read_data(os.path.join(golden_dir, "aniline_10_struct.nc"))
    return system


def _central_differences(objective_function, x, rel=1e-5):
    steps = rel * np.maximum(1.0, np.abs(x))
    rows = np.repeat(x[None], 2 * len(x), 0)
    rows[2 * np.arange(len(x)), np.arange(len(x))] += steps
    rows[2 * np.arange(len(x)) + 1, np.arange(len(x))] -= steps
    vals = objective_function.f_batch(rows)
    return (vals[0::2] - vals[1::2]) / (2 * steps)


CASES = [
    dict(flags=dict(opt_bonds=True, opt_angles=True, opt_torsions=True), weighting="uniform", term="components", reg="L2", pre=False),
    dict(flags=dict(opt_bonds=True, opt_angles=True, opt_torsions=True, opt_charges=True, opt_lj=True), weighting="boltzmann",
         term="norm", reg="hyperbolic", pre=True),
    dict(flags=dict(opt_sc=True, opt_torsions=True), weighting="uniform", term="components", reg=None, pre=False),
    dict(flags=dict(opt_charges=True, opt_lj=True), weighting="manual", term="norm", reg="L1", pre=False),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("symmetry", [True, False])
def test_objective_gradient_vs_central_differences(kwargs_dict, golden_dir, case, symmetry, tmp_path, monkeypatch):
    """f_and_grad through ParameterSpace / ForceField / B200Engine (symmetry groups, abs() conventions, the three nonbonded
    branches, preconditioning, regularisation) against central differences of f_batch, away from the initial point."""
    monkeypatch.chdir(tmp_path)
    system = _aniline(kwargs_dict, golden_dir, **case["flags"])
    rng = np.random.default_rng(5)
    if case["weighting"] == "manual":
        system.compute_conformations_weights(weighting_method="manual", manual_weights_array=rng.uniform(0.02, 0.2, system.n_structures))
    parameter_space = ParameterSpace()
    _, x0 = parameter_space.get_optimizable_parameters([system], symmetry_constrained=symmetry)
    x0 = np.asarray(x0, dtype=np.float64)
    parameter_space.calculate_prior_widths("arithmetic")
    props = []
    e_prop = EnergyProperty([system], weight=0.7)
    e_prop.calculate_variance()
    f_prop = ForceProperty([system], term_type=case["term"], weight=1.3)
    f_prop.calculate_variance()
    props += [e_prop, f_prop]
    if case["pre"]:
        parameter_space.calculate_scaling_constants("arithmetic")
        parameter_space.jacobi_preconditioning()
        x0 = np.asarray(parameter_space.optimizable_parameters_values_scaled, dtype=np.float64)
    if case["reg"]:
        props.append(Regularization(initial_parameters_values=x0.copy(), prior_widths=np.ones(len(x0)) if case["pre"]
                                    else parameter_space.prior_widths, method=case["reg"], weight=0.5, scaling_factor=0.3))
    objective_function = ObjectiveFunction(restart_settings=None, parameter_space=parameter_space, properties=props,
                                           systems=[system], weighting_method=case["weighting"], **SETTINGS)
    x = x0 * (1.0 + 0.01 * rng.uniform(-1, 1, len(x0))) + 1e-4 * rng.uniform(-1, 1, len(x0))
    value, grad = objective_function.f_and_grad(x)
    assert abs(value - objective_function.f_batch(x[None])[0]) <= 1e-10 * abs(value)
    assert abs(value - objective_function.f(x)) <= 1e-10 * abs(value)
    fd = _central_differences(objective_function, x)
    assert grad.shape == x.shape and np.all(np.isfinite(grad))
    assert np.abs(grad - fd).max() <= 2e-6 * np.abs(fd).max()
    assert np.array_equal(objective_function.grad(x), grad)


def test_gradient_two_systems_and_scipy_jac(kwargs_dict, golden_dir, tmp_path, monkeypatch):
    """Two systems sharing symmetry groups; scipy's BFGS driven by the analytic gradient lowers the objective."""
    import scipy.optimize
    monkeypatch.chdir(tmp_path)
    flags = dict(opt_bonds=True, opt_angles=True, opt_torsions=True)
    s1, s2 = _aniline(kwargs_dict, golden_dir, **flags), _aniline(kwargs_dict, golden_dir, **flags)
    s2.ref_energies = np.asarray(s2.ref_energies) + np.linspace(-3, 3, s2.n_structures)
    parameter_space = ParameterSpace()
    _, x0 = parameter_space.get_optimizable_parameters([s1, s2])
    x0 = np.asarray(x0, dtype=np.float64)
    e_prop = EnergyProperty([s1, s2], weight=1.0)
    e_prop.calculate_variance()
    f_prop = ForceProperty([s1], term_type="components", weight=1.0)
    f_prop.calculate_variance()
    objective_function = ObjectiveFunction(restart_settings=None, parameter_space=parameter_space, properties=[e_prop, f_prop],
                                           systems=[s1, s2], weighting_method="uniform", **SETTINGS)
    value, grad = objective_function.f_and_grad(x0)
    fd = _central_differences(objective_function, x0)
    assert np.abs(grad - fd).max() <= 2e-6 * np.abs(fd).max()
    res = scipy.optimize.minimize(objective_function.f_and_grad, x0, jac=True, method="L-BFGS-B", options=dict(maxiter=15))
    assert res.fun < 0.7 * value
    assert res.nfev <= 40                      # a 2-point gradient would need ~ len(x0) + 1 evaluations per step


def test_non_boltzmann_is_refused(kwargs_dict, golden_dir):
    system = _aniline(kwargs_dict, golden_dir, opt_bonds=True)
    parameter_space = ParameterSpace()
    _, x0 = parameter_space.get_optimizable_parameters([system])
    e_prop = EnergyProperty([system], weight=1.0)
    e_prop.calculate_variance()
    objective_function = ObjectiveFunction(restart_settings=None, parameter_space=parameter_space, properties=[e_prop],
                                           systems=[system], weighting_method="non_boltzmann", **SETTINGS)
    with pytest.raises(NotImplementedError):
        objective_function.f_and_grad(np.asarray(x0))


def test_gradient_with_esp_property(kwargs_dict, golden_dir, tmp_path, monkeypatch):
    """ENERGY + ESP: the charge derivatives of the ESP quadratic form (2 (A q - B)) reach the gradient through the particle
    slots, together with the energy part, and agree with central differences."""
    from paramol_b200.MM_engines.resp import RESP
    from paramol_b200.Objective_function.Properties.esp_property import ESPProperty
    monkeypatch.chdir(tmp_path)
    engine = B200Engine(True, **kwargs_dict)
    system = ParaMolSystem(name="aniline", engine=engine, n_atoms=14)
    d = np.load(os.path.join(golden_dir, "aniline_esp.npz"))
    rng = np.random.default_rng(2)
    sel = rng.choice(len(d["esp"]), 2000, replace=False)
    system.ref_coordinates = [d["conformation"], d["conformation"] + rng.normal(0, 0.01, (14, 3))]
    system.ref_esp_grid = [d["grid"][sel], d["grid"][sel[:1500]]]
    system.ref_esp = [d["esp"][sel], d["esp"][sel[:1500]]]
    system.n_structures = 2
    # MM-scale reference energies (the QM energy of the fixture is on another scale and would make the objective ~1e10, where
    # central differences drown in round-off)
    system.ref_energies = np.asarray(system.get_energies_ensemble()) + np.array([1.0, -2.0])
    system.force_field.create_force_field(opt_charges=True, opt_bonds=True)
    ps = ParameterSpace()
    _, x0 = ps.get_optimizable_parameters([system])
    x0 = np.asarray(x0, dtype=np.float64)
    system.resp_engine = RESP(0, False, "L2", 1.0, 0.01, "uniform", 300.0)
    system.resp_engine.set_initial_charges(system.force_field.force_field)
    system.resp_engine.set_charges(system.force_field.force_field)
    system.resp_engine.calculate_inverse_distances(system)
    esp_prop = ESPProperty([system], weight=1.5)
    esp_prop.calculate_variance()
    e_prop = EnergyProperty([system], weight=0.8)
    e_prop.calculate_variance()
    of = ObjectiveFunction(None, ps, [e_prop, esp_prop], [system], weighting_method="uniform", **SETTINGS)
    x = x0 * (1.0 + 0.01 * rng.uniform(-1, 1, len(x0))) + 1e-3 * rng.uniform(-1, 1, len(x0))
    value, grad = of.f_and_grad(x)
    assert abs(value - of.f_batch(x[None])[0]) <= 1e-10 * abs(value)
    fd = _central_differences(of, x)
    assert np.abs(grad - fd).max() <= 2e-6 * np.abs(fd).max()
    charge = np.asarray([p.param_key == "charge" for p in ps.optimizable_parameters])
    assert np.abs(grad[charge]).max() > 0 and np.abs(grad[~charge]).max() > 0
