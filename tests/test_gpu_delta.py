"""
GPU tests of delta evaluation (``pm_delta_eval``, ``ObjectiveFunction(delta_evaluation=True)``; run with ``-m gpu``): trial
vectors that differ from the base in one parameter -- the moves of the reference's Monte Carlo and simulated annealing
(ParaMol/Optimizers/monte_carlo.py:89-100) -- must give the per-conformation energies / residuals and the objective of a full
evaluation (1e-11 relative per conformation, 1e-10 on the objective), through commits of accepted moves as well.
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
from paramol_b200.Optimizers.monte_carlo import MonteCarlo
from paramol_b200.Utils.synthetic import make_ligand

SETTINGS = {"parallel": False, "platform_name": "Reference", "weighting_temperature": 300.0, "checkpoint_freq": 10 ** 9}


@pytest.mark.parametrize("n_atoms,n_conf", [(12, 70), (41, 130), (60, 257)])
def test_delta_rows_equal_full_evaluation_in_slot_space(n_atoms, n_conf):
    """Every kind of slot perturbed one at a time (and a few at once): emm and fres of pm_delta_eval vs pm_eval; then a commit
    and a second round relative to the new base."""
    import torch
    mm, x0 = make_ligand(n_atoms)
    engine = B200Engine(system=mm)
    topo = engine.device_topology()
    rng = np.random.default_rng(8)
    X = x0[None] + rng.normal(0, 0.01, (n_conf,) + x0.shape)
    ens = DeviceEnsemble(topo, X, rng.normal(0, 50.0, X.shape), rng.normal(0, 5.0, n_conf))
    metric = np.tile(np.eye(3).ravel(), (n_atoms, 1)) * rng.uniform(0.5, 2.0, (n_atoms, 1))
    metric[:, 1] += 0.1
    topo.set_force_metric(metric)
    base = engine.current_slots().copy()

    def full(rows):
        der = topo.prepare(torch.from_numpy(np.ascontiguousarray(rows)).cuda())
        emm, fres, fmm = ens.evaluate(der, with_forces=True, want_fres=True, want_fmm=True)
        return der, emm.clone(), fres.clone(), fmm.clone()

    def check(base_row, rows):
        plan = topo.delta_plan(base_row, rows)
        assert plan is not None
        der, emm_f, fres_f, _ = full(rows)
        emm_d, fres_d = ens.delta_evaluate(plan, der)
        assert (emm_d - emm_f).abs().max().item() <= 1e-11 * emm_f.abs().max().item()
        assert (fres_d - fres_f).abs().max().item() <= 1e-11 * fres_f.abs().max().item()
        return plan

    der0, emm0, fres0, fmm0 = full(base[None])
    ens.set_delta_base(base, der0[0], emm0[0], fres0[0], fmm0[0])
    bounds = [0, topo.off_angle, topo.off_torsion, topo.off_particle, topo.off_exception, topo.n_slots]
    rows = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        for idx in rng.choice(np.arange(lo, hi), size=min(hi - lo, 6), replace=False):
            if lo == topo.off_particle and n_atoms > 30 and (idx - lo) % 3 != 0:
                pass   # sigma / eps of one atom touch all its pairs, like the charge: all are exercised
            row = base.copy()
            row[idx] = row[idx] * (1.0 + 0.03 * rng.uniform(0.5, 1.0)) + 1e-3
            rows.append(row)
    multi = base.copy()
    multi[[0, topo.off_angle + 1, topo.off_torsion + 2, topo.off_torsion + 3]] *= 1.02
    rows.append(multi)
    rows.append(base.copy())                       # an untouched row returns the base itself
    rows = np.stack(rows)
    plan = check(base, rows)
    assert plan["n_entries"] > 0
    # commit the multi-slot row, then evaluate moves relative to it
    one = topo.delta_plan(base, multi[None])
    ens.delta_commit(one, multi, topo.prepare(torch.from_numpy(multi[None].copy()).cuda()))
    _, emm_m, fres_m, fmm_m = full(multi[None])
    b = ens.delta_base
    assert (b["emm"] - emm_m[0]).abs().max().item() <= 1e-11 * emm_m.abs().max().item()
    assert (b["fres"] - fres_m[0]).abs().max().item() <= 1e-11 * fres_m.abs().max().item()
    f_commit, f_full = topo.unpack(b["fmm"], n_conf), topo.unpack(fmm_m[0], n_conf)   # (tile padding is not maintained)
    assert (f_commit - f_full).abs().max().item() <= 1e-11 * f_full.abs().max().item()
    rows2 = np.repeat(multi[None], 5, 0)
    rows2[np.arange(5), rng.choice(topo.off_particle, 5, replace=False)] *= 0.97
    check(multi, rows2)
    # energies only
    ens_e = DeviceEnsemble(topo, X, None, rng.normal(0, 5.0, n_conf))
    der, emm_f, _, _ = (lambda r: (topo.prepare(torch.from_numpy(np.ascontiguousarray(r)).cuda()),) + (None,) * 3)(rows)
    e_base, _, _ = ens_e.evaluate(der0, with_forces=False, want_fres=False)
    ens_e.set_delta_base(base, der0[0], e_base[0], None, None)
    emm_full = ens_e.evaluate(der, with_forces=False, want_fres=False)[0].clone()
    emm_d, fres_d = ens_e.delta_evaluate(topo.delta_plan(base, rows), der)
    assert fres_d is None and (emm_d - emm_full).abs().max().item() <= 1e-11 * emm_full.abs().max().item()
    # a row that rewrites most of the force field is not a delta
    assert topo.delta_plan(base, (base * 1.01)[None]) is None


def _objective(golden_dir, delta, weighting="uniform", flags=None):
    engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(golden_dir, "aniline.prmtop"),
                        crd_file=os.path.join(golden_dir, "aniline.inpcrd"))
    system = ParaMolSystem(name="aniline", engine=engine, n_atoms=14)
    system.force_field.create_force_field(ff_file=None, **(flags or dict(opt_bonds=True, opt_angles=True, opt_torsions=True,
                                                                          opt_charges=True, opt_lj=True)))
    system.read_data(os.path.join(golden_dir, "aniline_10_struct.nc"))
    ps = ParameterSpace()
    _, x0 = ps.get_optimizable_parameters([system])
    x0 = np.asarray(x0, dtype=np.float64)
    ps.calculate_prior_widths("arithmetic")
    e_prop = EnergyProperty([system], weight=1.0)
    e_prop.calculate_variance()
    f_prop = ForceProperty([system], term_type="components", weight=1.0)
    f_prop.calculate_variance()
    r_prop = Regularization(x0.copy(), ps.prior_widths, "L2", weight=1.0, scaling_factor=1.0)
    of = ObjectiveFunction(None, ps, [e_prop, f_prop, r_prop], [system], weighting_method=weighting, delta_evaluation=delta, **SETTINGS)
    return of, x0


@pytest.mark.parametrize("weighting", ["uniform", "boltzmann", "non_boltzmann"])
def test_objective_single_parameter_moves(golden_dir, weighting, tmp_path, monkeypatch):
    """f_batch rows = current vector + one displaced parameter (bonded, charge, sigma, epsilon), through an accepted move."""
    monkeypatch.chdir(tmp_path)
    of_d, x0 = _objective(golden_dir, True, weighting)
    of_f, _ = _objective(golden_dir, False, weighting)
    rng = np.random.default_rng(2)
    x = x0.copy()
    assert abs(of_d.f(x) - of_f.f(x)) <= 1e-12
    for step in range(3):
        trial = np.repeat(x[None], 24, 0)
        which = rng.choice(len(x), 24, replace=False)
        trial[np.arange(24), which] *= 1.0 + 0.02 * rng.uniform(-1, 1, 24)
        got, want = of_d.f_batch(trial), of_f.f_batch(trial)
        assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
        x = trial[int(np.argmin(want))].copy()      # "accept" the best one: the next batch is built on it
    assert of_d.n_delta_rows == 72                  # every row went through pm_delta_eval
    assert abs(of_d.f(x) - of_f.f(x)) <= 1e-12     # a single vector is always a full evaluation


def test_monte_carlo_trajectory_with_delta_evaluation(golden_dir, tmp_path, monkeypatch):
    """The batched Monte Carlo driver walks the same trajectory with and without delta evaluation."""
    monkeypatch.chdir(tmp_path)
    results = []
    for delta in (False, True):
        of, x0 = _objective(golden_dir, delta, flags=dict(opt_bonds=True, opt_angles=True, opt_torsions=True))
        np.random.seed(11)
        mc = MonteCarlo(n_blocks=5, max_iter=3, f_tol=1e-12, avg_acceptance_rate=0.3, batch_size=16)
        x = mc.run_optimization(of.f, list(x0))
        results.append((np.asarray(x, dtype=np.float64), of.f(np.asarray(x, dtype=np.float64)), of.n_delta_rows))
    assert np.abs(results[0][0] - results[1][0]).max() <= 1e-12 * np.abs(results[0][0]).max()
    assert abs(results[0][1] - results[1][1]) <= 1e-10 * abs(results[0][1])
    assert results[0][2] == 0 and results[1][2] > 0


def test_delta_evaluation_two_systems(golden_dir, tmp_path, monkeypatch):
    """Two systems with their own bases: single-parameter rows of f_batch through pm_delta_eval for both, values equal to full
    evaluations."""
    monkeypatch.chdir(tmp_path)
    built = []
    for delta in (True, False):
        systems = []
        for k in range(2):
            engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(golden_dir, "aniline.prmtop"),
                                crd_file=os.path.join(golden_dir, "aniline.inpcrd"))
            system = ParaMolSystem(name="aniline%d" % k, engine=engine, n_atoms=14)
            system.force_field.create_force_field(ff_file=None, opt_bonds=True, opt_angles=True, opt_torsions=True)
            system.read_data(os.path.join(golden_dir, "aniline_10_struct.nc"))
            if k == 1:
                system.ref_energies = np.asarray(system.ref_energies) + np.linspace(-2, 2, system.n_structures)
            systems.append(system)
        ps = ParameterSpace()
        _, x0 = ps.get_optimizable_parameters(systems)
        e_prop = EnergyProperty(systems, weight=1.0)
        e_prop.calculate_variance()
        f_prop = ForceProperty(systems, term_type="norm", weight=1.0)
        f_prop.calculate_variance()
        built.append((ObjectiveFunction(None, ps, [e_prop, f_prop], systems, weighting_method="uniform", delta_evaluation=delta, **SETTINGS),
                      np.asarray(x0, dtype=np.float64)))
    (of_d, x0), (of_f, _) = built
    rng = np.random.default_rng(4)
    assert abs(of_d.f(x0) - of_f.f(x0)) <= 1e-12
    trial = np.repeat(x0[None], 16, 0)
    which = rng.choice(len(x0), 16, replace=False)
    trial[np.arange(16), which] *= 1.0 + 0.02 * rng.uniform(-1, 1, 16)
    got, want = of_d.f_batch(trial), of_f.f_batch(trial)
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    assert of_d.n_delta_rows == 32      # 16 rows x 2 systems
