"""
GPU parity of the objective function through the reference-facing API (run with ``-m gpu``).  The first tests
are the reference's own test bodies (ParaMol/Tests/Objective_function/test_objective_function.py:26-145,
Tests/System/test_system.py:94-169, Tests/MM_engines/test_openmm_wrapper.py:43-72) with the imports pointed at
this package; the rest compare the fused device path with the CPU oracle on seeded inputs.
Tolerances: objective 1e-8 relative (the reference asserts 1e-8 absolute on O(0.1-0.5) values), E/F 1e-6 rel, 1e-4 abs floor.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from paramol_b200.System.system import ParaMolSystem
from paramol_b200.MM_engines.b200_engine import B200Engine
from paramol_b200.Force_field.force_field import ForceField
from paramol_b200.Objective_function.objective_function import ObjectiveFunction
from paramol_b200.Objective_function.Properties.regularization import Regularization
from paramol_b200.Objective_function.Properties.force_property import ForceProperty
from paramol_b200.Objective_function.Properties.energy_property import EnergyProperty
from paramol_b200.Parameter_space.parameter_space import ParameterSpace

objective_function_settings = {"parallel": False,
                               "platform_name": "Reference",
                               "weighting_method": "uniform",
                               "weighting_temperature": 300.0,
                               "checkpoint_freq": 100}


@pytest.fixture()
def kwargs_dict(golden_dir):
    return {"topology_format": "AMBER", "crd_format": "AMBER",
            "top_file": os.path.join(golden_dir, "aniline.prmtop"), "crd_file": os.path.join(golden_dir, "aniline.inpcrd")}


def _aniline(kwargs_dict, golden_dir, **opt):
    openmm_engine = B200Engine(True, **kwargs_dict)
    system = ParaMolSystem(name="aniline", engine=openmm_engine, n_atoms=14)
    assert type(system.force_field) is ForceField
    flags = dict(opt_bonds=True, opt_angles=True, opt_torsions=True, opt_charges=True, opt_lj=True, opt_sc=True)
    flags.update(opt)
    system.force_field.create_force_field(ff_file=None, **flags)
    system.read_data(os.path.join(golden_dir, "aniline_10_struct.nc"))
    return system


@pytest.mark.parametrize("fused", [True, False])
def test_objective_function_and_properties(kwargs_dict, golden_dir, goldens, fused, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)  # the checkpoint XML of f() lands in the cwd, as in the reference
    system = _aniline(kwargs_dict, golden_dir)
    parameter_space = ParameterSpace()
    optimizable_parameters, optimizable_parameters_values = parameter_space.get_optimizable_parameters([system])
    gold = goldens["test_objective_function.py::test_objective_function_and_properties::optimizable_parameters_values_to_compare"]
    np.testing.assert_almost_equal(gold, optimizable_parameters_values, decimal=4)

    prop_energy = EnergyProperty([system], **{"weight": 1.0})
    prop_energy.calculate_variance()
    objective_function = ObjectiveFunction(restart_settings=None, parameter_space=parameter_space, properties=[prop_energy],
                                           systems=[system], fused=fused, **objective_function_settings)
    f_val = objective_function.f(optimizable_parameters_values, opt_mode=False)
    assert abs(f_val - 0.16587870225911971) < 1e-8
    assert abs(f_val - 0.16587870225911971) <= 1e-8 * 0.16587870225911971  # north-star bar: 1e-8 relative
    assert abs(prop_energy.value - f_val) < 1e-15

    prop_forces_components = ForceProperty([system], **{"term_type": "components", "weight": 1.0})
    prop_forces_components.calculate_variance()
    objective_function = ObjectiveFunction(restart_settings=None, parameter_space=parameter_space,
                                           properties=[prop_forces_components], systems=[system], fused=fused,
                                           **objective_function_settings)
    f_val = objective_function.f(optimizable_parameters_values, opt_mode=False)
    assert abs(f_val - 0.3022524295062905) <= 1e-8 * 0.3022524295062905

    prop_forces_norm = ForceProperty([system], **{"term_type": "norm", "weight": 1.0})
    prop_forces_norm.calculate_variance()
    objective_function = ObjectiveFunction(restart_settings=None, parameter_space=parameter_space, properties=[prop_forces_norm],
                                           systems=[system], fused=fused, **objective_function_settings)
    f_val = objective_function.f(optimizable_parameters_values, opt_mode=False)
    assert abs(f_val - 0.498237598004599) <= 1e-8 * 0.498237598004599

    parameter_space.calculate_prior_widths("arithmetic")
    prop_reg = Regularization(initial_parameters_values=parameter_space.optimizable_parameters_values,
                              prior_widths=parameter_space.prior_widths, **{"method": "L2", "weight": 1.0, "scaling_factor": 1.0})
    objective_function = ObjectiveFunction(restart_settings=None, parameter_space=parameter_space, properties=[prop_reg],
                                           systems=[system], fused=fused, **objective_function_settings)
    f_val = objective_function.f(optimizable_parameters_values, opt_mode=False)
    assert abs(f_val) < 1e-8

    # all four together, counting calls
    objective_function = ObjectiveFunction(restart_settings=None, parameter_space=parameter_space,
                                           properties=[prop_energy, prop_forces_norm, prop_reg], systems=[system], fused=fused,
                                           **objective_function_settings)
    f_val = objective_function.f(optimizable_parameters_values)
    assert abs(f_val - (0.16587870225911971 + 0.498237598004599)) < 1e-8
    assert objective_function.get_f_count() == 1 and objective_function.reset_f_count() == 0


def test_system_ensembles_and_weights(kwargs_dict, golden_dir, goldens):
    # Tests/System/test_system.py:94-169
    system = _aniline(kwargs_dict, golden_dir)
    emm = system.get_energies_ensemble()
    np.testing.assert_almost_equal(emm, goldens["test_system.py::test_get_energies_ensemble::energies_to_compare"], decimal=4)
    assert np.abs(emm - np.asarray(goldens["test_system.py::test_get_energies_ensemble::energies_to_compare"])).max() < 1e-7
    fmm = system.get_forces_ensemble()
    assert fmm.shape == (10, 14, 3)
    np.testing.assert_almost_equal(fmm[0], goldens["test_system.py::test_get_forces_ensemble::forces_to_compare"], decimal=4)
    k = "test_system.py::test_weighting_methods::"
    w = system.compute_conformations_weights(temperature=300.0, emm=emm, weighting_method="non_boltzmann")
    np.testing.assert_almost_equal(w, goldens[k + "non_boltzmann_to_compare"], decimal=6)
    # single-conformation engine calls
    e = system.engine.get_potential_energy(system.ref_coordinates[3])
    assert abs(e - emm[3]) < 1e-9
    f = system.engine.get_forces(system.ref_coordinates[3])
    assert np.abs(f - fmm[3]).max() < 1e-9


def _oracle_objective(system, x_slots_system, weighting, temperature=300.0, term_type="norm"):
    from oracle import oracle as O
    total, parts = O.objective(x_slots_system, np.asarray(system.ref_coordinates), np.asarray(system.ref_energies),
                               np.asarray(system.ref_forces), force_term_type=term_type, weighting_method=weighting,
                               temperature=temperature)
    return total, parts


@pytest.mark.parametrize("weighting", ["uniform", "boltzmann", "non_boltzmann"])
@pytest.mark.parametrize("term_type", ["norm", "components"])
def test_fused_objective_vs_oracle_with_perturbed_parameters(kwargs_dict, golden_dir, weighting, term_type):
    system = _aniline(kwargs_dict, golden_dir, opt_sc=False)  # branch A: charges/LJ reach the 1-4 exceptions
    ps = ParameterSpace()
    _, x0 = ps.get_optimizable_parameters([system])
    x0 = np.asarray(x0)
    e_prop, f_prop = EnergyProperty([system], weight=0.7), ForceProperty([system], term_type=term_type, weight=1.3)
    e_prop.calculate_variance()
    f_prop.calculate_variance()
    ps.calculate_prior_widths("arithmetic")
    r_prop = Regularization(x0.copy(), ps.prior_widths, "L2", weight=0.5, scaling_factor=2.0)
    settings = dict(objective_function_settings, weighting_method=weighting, checkpoint_freq=10 ** 9)
    of = ObjectiveFunction(None, ps, [e_prop, f_prop, r_prop], [system], **settings)
    of_serial = ObjectiveFunction(None, ps, [e_prop, f_prop, r_prop], [system], fused=False, **settings)
    rng = np.random.default_rng(5)
    X = np.stack([x0 * (1.0 + 0.01 * rng.uniform(-1, 1, x0.shape)) for _ in range(4)])
    X[0] = x0
    # oracle values: push each vector through the (CPU-tested) parameter plumbing, evaluate with the oracle
    want = []
    for x in X:
        ps.update_systems([system], x.copy())
        total, parts = _oracle_objective(system, system.engine.system, weighting, term_type=term_type)
        val = 0.7 * parts["ENERGY"] + 1.3 * parts["FORCE"] + 0.5 * 2.0 * np.sum(((x - x0) / ps.prior_widths) ** 2)
        want.append(val)
    want = np.asarray(want)
    ps.update_systems([system], x0.copy())
    got_batch = of.f_batch(X)
    assert np.all(np.abs(got_batch - want) <= 1e-8 * np.abs(want)), (got_batch, want)
    # the batch left the parameter space untouched
    assert np.array_equal(np.asarray(ps.optimizable_parameters_values), x0)
    for x, w in zip(X, want):
        got = of.f(x.copy())
        assert abs(got - w) <= 1e-8 * abs(w)
        got_serial = of_serial.f(x.copy())
        assert abs(got_serial - w) <= 1e-8 * abs(w)


@pytest.mark.parametrize("spread", [30.0, 400.0, 2500.0, -1850.0])
def test_non_boltzmann_fused_equals_seam_for_wide_energy_spreads(kwargs_dict, golden_dir, spread):
    """NON_BOLTZMANN weights fall back to uniform on ANY floating-point error of the reference's np.seterr(all='raise') block
    (system.py:193, 215-221) -- including UNDERFLOW of exp(-beta d), which sets in at a spread of ~1.7e3 kJ/mol at 300 K.  The
    fused device weights must take the same branch as the host path (``fused=False``), whatever the spread."""
    system = _aniline(kwargs_dict, golden_dir, opt_charges=False, opt_lj=False, opt_sc=False)
    if spread > 0:
        system.ref_energies = np.asarray(system.ref_energies) + spread * np.arange(10)
    else:   # one outlier: nothing overflows, ONE exponential underflows
        system.ref_energies = np.asarray(system.ref_energies) + spread * (np.arange(10) == 7)
    ps = ParameterSpace()
    _, x0 = ps.get_optimizable_parameters([system])
    e_prop, f_prop = EnergyProperty([system]), ForceProperty([system], term_type="norm")
    e_prop.calculate_variance()
    f_prop.calculate_variance()
    settings = dict(objective_function_settings, weighting_method="non_boltzmann", checkpoint_freq=10 ** 9)
    fused = ObjectiveFunction(None, ps, [e_prop, f_prop], [system], **settings)
    seam = ObjectiveFunction(None, ps, [e_prop, f_prop], [system], fused=False, **settings)
    a = fused.f(np.asarray(x0).copy())
    w_fused = np.array(system.weights, copy=True)
    b = seam.f(np.asarray(x0).copy())
    w_seam = np.array(system.weights, copy=True)
    assert np.abs(w_fused - w_seam).max() <= 1e-12 * np.abs(w_seam).max(), (w_fused, w_seam)
    assert abs(a - b) <= 1e-10 * abs(b)
    if spread >= 2500.0 or spread < 0:
        assert np.allclose(w_seam, 0.1)    # the underflow branch was taken


def test_nan_objective_becomes_1e16(kwargs_dict, golden_dir):
    system = _aniline(kwargs_dict, golden_dir, opt_charges=False, opt_lj=False, opt_sc=False)
    ps = ParameterSpace()
    _, x0 = ps.get_optimizable_parameters([system])
    e_prop = EnergyProperty([system])
    e_prop.calculate_variance()
    of = ObjectiveFunction(None, ps, [e_prop], [system], **dict(objective_function_settings, checkpoint_freq=10 ** 9))
    x = np.asarray(x0).copy()
    x[1] = np.nan
    assert of.f(x) == 1e16
    assert of.f_batch(np.stack([x, np.asarray(x0)]))[0] == 1e16


def test_optimizers_drive_the_device_objective(kwargs_dict, golden_dir):
    """Monte Carlo and finite-difference gradient descent on the real objective: f_batch launches, objective decreases."""
    from paramol_b200.Optimizers.monte_carlo import MonteCarlo
    from paramol_b200.Optimizers.gradient_descent import GradientDescent
    system = _aniline(kwargs_dict, golden_dir, opt_charges=False, opt_lj=False, opt_sc=False, opt_bonds=False, opt_angles=False)
    ps = ParameterSpace()
    _, x0 = ps.get_optimizable_parameters([system])
    ps.calculate_scaling_constants("default")
    ps.jacobi_preconditioning()
    xs0 = np.asarray(ps.optimizable_parameters_values_scaled).copy()
    e_prop = EnergyProperty([system])
    e_prop.calculate_variance()
    of = ObjectiveFunction(None, ps, [e_prop], [system], **dict(objective_function_settings, checkpoint_freq=10 ** 9))
    f0 = of.f(xs0.copy())
    np.random.seed(3)
    mc = MonteCarlo(n_blocks=2, max_iter=3, f_tol=1e-14, avg_acceptance_rate=0.25, batch_size=16)
    x_mc = mc.run_optimization(of.f, list(xs0))
    f_mc = of.f(np.asarray(x_mc))
    assert f_mc <= f0 and mc.n_launches < 2 * len(xs0)
    gd = GradientDescent(max_iter=3, derivative_calculation="always", derivative_type="1-point", g_tol=1e9, f_tol=1e-16,
                         dx=1e-4, derivative_h=1e-3)
    x_gd = gd.run_optimization(of.f, xs0.copy())
    assert of.f(np.asarray(x_gd)) <= f0 and gd.n_launches == 3
    # f_batch on the gradient stencil equals row-by-row f (1e-12)
    X = np.repeat(xs0[None], 5, axis=0)
    X[np.arange(5), np.arange(5)] += 1e-3
    assert np.allclose(of.f_batch(X), [of.f(x.copy()) for x in X], rtol=1e-12, atol=0)


def test_openmm_wrapper_goldens_reaction_field(kwargs_dict, golden_dir, goldens):
    """
    Tests/MM_engines/test_openmm_wrapper.py:43-72: because an earlier test of that class mutates the shared kwargs, the
    reference's single-structure energy/force goldens are evaluated on the serialized aniline.xml system
    (CutoffNonPeriodic, reaction field eps = 78.3, r_c = 1.2 nm).  Same here, on the device.
    """
    from paramol_b200.Utils.amber import read_inpcrd
    kw = dict(kwargs_dict, xml_file=os.path.join(golden_dir, "aniline.xml"))
    engine = B200Engine(True, **kw)
    assert engine.system.nonbonded_method == "CutoffNonPeriodic"
    x = read_inpcrd(kwargs_dict["crd_file"])
    e = engine.get_potential_energy(x)
    assert abs(e - goldens["test_openmm_wrapper.py::test_OpenMMEngine_get_potential_energy::energy_to_compare"]) < 1e-8
    f = engine.get_forces(x)
    gold = np.asarray(goldens["test_openmm_wrapper.py::test_OpenMMEngine_get_forces::forces_to_compare"])
    assert np.abs(f - gold).max() < 1e-4  # the golden table is printed to 8 significant digits
    from oracle import oracle as O
    E, F = O.energies_forces(engine.system, x)
    assert abs(e - E[0]) < 1e-10 and np.abs(f - F[0]).max() < 1e-9
    assert engine.get_atom_list() == goldens["test_openmm_wrapper.py::test_OpenMMEngine_get_atom_list::atom_list_to_compare"]
    assert engine.get_atomic_numbers() == goldens["test_openmm_wrapper.py::test_OpenMMEngine_get_atomic_numbers::atomic_numbers_to_compare"]


@pytest.mark.parametrize("n_gen", [3, 20])
def test_wham_reweighing_matches_sequential_restatement(kwargs_dict, golden_dir, n_gen):
    """ParaMolSystem.wham_reweighing (System/system.py:250-322): batched energy passes == one pass per generation.  With 16 or
    more generations the energies of all generations come from the batch contraction (``pm_energy_batch_reduce``)."""
    system = _aniline(kwargs_dict, golden_dir, opt_charges=False, opt_lj=False, opt_sc=False)
    _, x0 = system.force_field.get_optimizable_parameters()
    x0 = np.asarray(x0)
    rng = np.random.default_rng(4)
    gens = [list(x0 * (1.0 + 0.002 * rng.uniform(-1, 1, x0.shape))) for _ in range(n_gen)]
    got = system.wham_reweighing(gens)
    # restatement with one ensemble pass per generation
    weights = []
    for g in gens:
        system.force_field.update_force_field(g)
        system.engine.set_bonded_parameters(system.force_field.force_field_optimizable)
        system.engine.set_nonbonded_parameters(system.force_field.force_field_optimizable)
        e = np.exp(-0.5 * system.get_energies_ensemble())
        weights.append(e / e.sum())
    weights = np.asarray(weights)
    n_gen, S = weights.shape
    A = np.ones(n_gen) / n_gen
    prev, rmsd = weights[-1], 1e9
    while rmsd > 1e-6:
        final = np.zeros_like(weights)
        for G in range(n_gen):
            for j in range(G + 1):
                final[G] += A[j] * weights[G] / weights[j]
            A[G] = final[G].sum()
            A = A / A.sum()
        rmsd = np.sqrt(np.sum((final[-1] - prev) ** 2) / S)
        prev = final[-1]
    assert got.shape == (10,) and np.allclose(got, final[-1], rtol=1e-10 if n_gen < 16 else 1e-8, atol=0)
    assert np.ndim(system.wham_weights) == 1


def test_two_systems_manual_weights_fused_equals_seam_path(kwargs_dict, golden_dir):
    """Two systems (aniline twice, different data), MANUAL weights: fused device reduction == reference-structured path."""
    from oracle import oracle as O
    from paramol_b200.Utils.synthetic import perturbed_conformations
    rng = np.random.default_rng(21)
    systems = []
    for k, n_conf in enumerate((37, 64)):
        system = _aniline(kwargs_dict, golden_dir, opt_charges=False, opt_lj=False, opt_sc=False)
        system.name = "aniline%d" % k
        X = perturbed_conformations(np.asarray(system.ref_coordinates), n_conf, sigma=0.004, seed=30 + k)
        E, F = O.energies_forces(system.engine.system, X)
        system.ref_coordinates, system.n_structures = X, n_conf
        system.ref_energies = E + rng.normal(0, 4.0, n_conf) - 43000.0
        system.ref_forces = F + rng.normal(0, 60.0, F.shape)
        w = rng.uniform(0.2, 1.0, n_conf)
        system.compute_conformations_weights(weighting_method="manual", manual_weights_array=w / w.sum())
        systems.append(system)
    ps = ParameterSpace()
    _, x0 = ps.get_optimizable_parameters(systems)
    x0 = np.asarray(x0)
    assert len(x0) == 2 * (28 + 42 + 70)
    e_prop, f_prop = EnergyProperty(systems, weight=0.4), ForceProperty(systems, term_type="components", weight=2.0)
    e_prop.calculate_variance()
    f_prop.calculate_variance()
    settings = dict(objective_function_settings, weighting_method="manual", checkpoint_freq=10 ** 9)
    of = ObjectiveFunction(None, ps, [e_prop, f_prop], systems, **settings)
    of_seam = ObjectiveFunction(None, ps, [e_prop, f_prop], systems, fused=False, **settings)
    x = x0 * (1.0 + 0.01 * rng.uniform(-1, 1, x0.shape))
    a = of.f(x.copy())
    fused_values = (e_prop.value, f_prop.value)
    b = of_seam.f(x.copy())
    assert abs(a - b) <= 1e-10 * abs(b)
    assert abs(fused_values[0] - e_prop.value) <= 1e-10 * abs(e_prop.value)
    assert abs(fused_values[1] - f_prop.value) <= 1e-10 * abs(f_prop.value)
    # oracle cross-check of the first system's energy term
    ps.update_systems(systems, x.copy())
    E0, _ = O.energies_forces(systems[0].engine.system, systems[0].ref_coordinates)
    want0 = O.energy_property(systems[0].ref_energies, E0, systems[0].weights)
    got = e_prop.calculate_property([s.get_energies_ensemble() for s in systems])
    assert abs(got[0] - want0) <= 1e-8 * abs(want0)


def test_esp_property_in_the_objective(kwargs_dict, golden_dir):
    """ESP property through ObjectiveFunction.f: fused quadratic form == seam path (system.py:356-382 + esp_property.py:89-104)."""
    from paramol_b200.MM_engines.resp import RESP
    from paramol_b200.Objective_function.Properties.esp_property import ESPProperty
    engine = B200Engine(True, **kwargs_dict)
    system = ParaMolSystem(name="aniline", engine=engine, n_atoms=14)
    d = np.load(os.path.join(golden_dir, "aniline_esp.npz"))
    rng = np.random.default_rng(2)
    sel = rng.choice(len(d["esp"]), 3000, replace=False)
    system.ref_coordinates = [d["conformation"], d["conformation"] + rng.normal(0, 0.01, (14, 3))]
    system.ref_esp_grid = [d["grid"][sel], d["grid"][sel[:2000]]]
    system.ref_esp = [d["esp"][sel], d["esp"][sel[:2000]]]
    system.n_structures = 2
    system.force_field.create_force_field(opt_charges=True)
    ps = ParameterSpace()
    _, q0 = ps.get_optimizable_parameters([system])
    system.resp_engine = RESP(0, False, "L2", 1.0, 0.01, "uniform", 300.0)
    system.resp_engine.set_initial_charges(system.force_field.force_field)
    system.resp_engine.set_charges(system.force_field.force_field)
    system.resp_engine.calculate_inverse_distances(system)
    prop = ESPProperty([system], weight=1.5)
    prop.calculate_variance()
    settings = dict(objective_function_settings, checkpoint_freq=10 ** 9)
    of = ObjectiveFunction(None, ps, [prop], [system], **settings)
    of_seam = ObjectiveFunction(None, ps, [prop], [system], fused=False, **settings)
    q = np.asarray(q0) + rng.normal(0, 0.05, len(q0))
    a, b = of.f(q.copy()), of_seam.f(q.copy())
    assert abs(a - b) <= 1e-8 * abs(b) and a > 0
    vals = of.f_batch(np.stack([q, np.asarray(q0)]))
    assert abs(vals[0] - b) <= 1e-8 * abs(b)


def test_nonbonded_cache_gives_the_same_objective(kwargs_dict, golden_dir):
    """cache_nonbonded=True: pair terms evaluated once per set of nonbonded parameters, bonded terms every call."""
    rng = np.random.default_rng(8)
    for opt, expect_refresh in ((dict(opt_charges=False, opt_lj=False, opt_sc=False), False), (dict(opt_sc=False), True)):
        systems = [_aniline(kwargs_dict, golden_dir, **opt) for _ in range(2)]
        objs = []
        for system, cached in zip(systems, (False, True)):
            ps = ParameterSpace()
            _, x0 = ps.get_optimizable_parameters([system])
            e_prop, f_prop = EnergyProperty([system]), ForceProperty([system], term_type="components")
            e_prop.calculate_variance()
            f_prop.calculate_variance()
            objs.append(ObjectiveFunction(None, ps, [e_prop, f_prop], [system], cache_nonbonded=cached,
                                          **dict(objective_function_settings, checkpoint_freq=10 ** 9)))
        x0 = np.asarray(x0)
        keys = set()
        for _ in range(4):
            x = x0 * (1.0 + 0.01 * rng.uniform(-1, 1, x0.shape))
            a, b = objs[0].f(x.copy()), objs[1].f(x.copy())
            assert abs(a - b) <= 1e-12 * abs(a), (a, b)
            keys.add(objs[1]._state[0]["nb_key"])
        X = np.stack([x0 * (1.0 + 0.01 * rng.uniform(-1, 1, x0.shape)) for _ in range(3)])
        va, vb = objs[0].f_batch(X), objs[1].f_batch(X)
        assert np.all(np.abs(va - vb) <= 1e-12 * np.abs(va))
        # bonded-only fit: one cache for all calls; with optimizable charges/LJ every vector brings its own
        assert (len(keys) > 1) == expect_refresh


@pytest.mark.parametrize("mode", ["zero_copy", "copies", "eager"])
def test_single_vector_graph_paths_agree_and_keep_the_host_state(kwargs_dict, golden_dir, mode, monkeypatch):
    """``f(x)`` through the captured graph -- launched and waited for by ``pm_graph_launch`` / ``pm_graph_wait``, with zero-copy ends
    (default) or with the copy nodes (``PARAMOL_B200_ZERO_COPY=0``) -- and without a graph (``PARAMOL_B200_GRAPHS=0``) gives the same
    values over a sequence of trial vectors (the graph is captured at the third call, so the sequence crosses the switch), and after
    every call the host state the deferred half of ``update_systems`` maintains (MM-system arrays, value lists) matches the slots
    row that was evaluated."""
    if mode == "copies":
        monkeypatch.setenv("PARAMOL_B200_ZERO_COPY", "0")
    if mode == "eager":
        monkeypatch.setenv("PARAMOL_B200_GRAPHS", "0")

    def run():
        system = _aniline(kwargs_dict, golden_dir)
        ps = ParameterSpace()
        _, x0 = ps.get_optimizable_parameters([system])
        x0 = np.asarray(x0, dtype=np.float64)
        ps.calculate_prior_widths("arithmetic")
        e_prop, f_prop = EnergyProperty([system]), ForceProperty([system], term_type="components")
        e_prop.calculate_variance()
        f_prop.calculate_variance()
        of = ObjectiveFunction(None, ps, [e_prop, f_prop, Regularization(x0.copy(), ps.prior_widths, "L2")], [system],
                               checkpoint_freq=10 ** 9)
        values = []
        for k in range(7):
            x = x0 * (1.0 + 0.02 * np.random.default_rng(100 + k).uniform(-1, 1, len(x0)))
            values.append(of.f(x))
            assert np.array_equal(system.engine.current_slots(), ps.last_slots[0][0])
            s = system.engine.system
            flat = np.concatenate([np.ravel(a) for a in (s.bond_par, s.angle_par, s.torsion_par, s.particle_par, s.exception_par)])
            assert np.array_equal(flat, ps.last_slots[0][0])
            assert len(ps.optimizable_parameters_values_by_system[0]) == len(x0)
        return np.asarray(values)

    got = run()
    monkeypatch.setenv("PARAMOL_B200_GRAPHS", "0")
    want = run()
    assert np.all(np.isfinite(got)) and len(set(np.round(got, 12))) == len(got)
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-12
