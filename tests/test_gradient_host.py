"""
CPU tests of the host half of the analytic gradient: the adjoints that carry d f / d slots back through
``B200Engine.slots_from_store`` (branches A/B/C, phase wrap, exceptions rebuilt from particle parameters) and through the
scatter of ``ParameterSpace.update_systems`` (symmetry groups, abs() conventions, preconditioning), checked against central
differences of ``update_systems_batch`` -- which is the forward map itself and runs without a GPU.  Plus the closed-form
regularisation gradients.  (The device half is covered by tests/test_gpu_gradient.py.)
"""
import os

import numpy as np
import pytest

from paramol_b200.System.system import ParaMolSystem
from paramol_b200.MM_engines.b200_engine import B200Engine
from paramol_b200.MM_engines.mm_system import MMSystem
from paramol_b200.Parameter_space.parameter_space import ParameterSpace
from paramol_b200.Objective_function.Properties.regularization import Regularization


def _system(golden_dir, name, flags):
    engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(golden_dir, name + ".prmtop"),
                        crd_file=os.path.join(golden_dir, name + ".inpcrd"))
    system = ParaMolSystem(name=name, engine=engine, n_atoms=engine.get_number_of_atoms())
    system.force_field.create_force_field(ff_file=None, **flags)
    return system


@pytest.mark.parametrize("flags,pre", [
    (dict(opt_bonds=True, opt_angles=True, opt_torsions=True), False),
    (dict(opt_bonds=True, opt_angles=True, opt_torsions=True, opt_charges=True, opt_lj=True), True),   # branch A
    (dict(opt_sc=True, opt_torsions=True), False),                                                      # branch B
    (dict(opt_charges=True, opt_lj=True, opt_sc=True), False),                                          # branch C: nothing pushed
    (dict(opt_charges=True, opt_lj=True), False),
])
@pytest.mark.parametrize("symmetry", [True, False])
def test_slot_gradient_adjoint_matches_forward_map(golden_dir, flags, pre, symmetry):
    system = _system(golden_dir, "aniline", flags)
    ps = ParameterSpace()
    _, x0 = ps.get_optimizable_parameters([system], symmetry_constrained=symmetry)
    x0 = np.asarray(x0, dtype=np.float64)
    if pre:
        ps.calculate_scaling_constants("arithmetic")
        ps.jacobi_preconditioning()
        x0 = np.asarray(ps.optimizable_parameters_values_scaled, dtype=np.float64)
    rng = np.random.default_rng(1)
    x = x0 * (1 + 0.01 * rng.uniform(-1, 1, len(x0))) + 1e-4 * rng.uniform(-1, 1, len(x0))
    # flip the sign of one lj parameter: the forward pass stores |value|, so its derivative carries sign(x)
    names = [p.param_key for p in ps.optimizable_parameters]
    if "lj_sigma" in names and not pre:
        x[names.index("lj_sigma")] *= -1.0
    ps.update_systems([system], x, symmetry_constrained=symmetry)
    g_slots = rng.normal(size=len(system.engine.current_slots()))
    g_store = system.engine.slots_gradient_to_store(system.force_field, g_slots, system.force_field.store)
    g_x = ps.gradient_from_store([system], x, [g_store], symmetry_constrained=symmetry)
    h = 1e-6 * np.maximum(1, np.abs(x))
    rows = np.repeat(x[None], 2 * len(x), 0)
    rows[2 * np.arange(len(x)), np.arange(len(x))] += h
    rows[2 * np.arange(len(x)) + 1, np.arange(len(x))] -= h
    slots = ps.update_systems_batch([system], rows, symmetry_constrained=symmetry)[0]
    fd = ((slots[0::2] - slots[1::2]) @ g_slots) / (2 * h)
    assert np.abs(fd - g_x).max() <= 1e-7 * max(np.abs(fd).max(), 1.0)
    if flags.get("opt_sc") and flags.get("opt_charges"):
        assert np.all(g_x == 0.0)     # reference quirk kept: with both optimizable nothing reaches the engine (openmm.py:752-754)


@pytest.mark.parametrize("method", ["L2", "L1", "hyperbolic"])
def test_regularization_gradient(method):
    rng = np.random.default_rng(2)
    x0, widths = rng.normal(size=9), rng.uniform(0.5, 2.0, 9)
    reg = Regularization(x0, widths, method, weight=1.0, scaling_factor=0.7, hyperbolic_beta=0.05)
    x = x0 + rng.normal(0, 0.3, 9)
    g = reg.calculate_gradient(x)
    h = 1e-6
    fd = np.array([(reg.calculate_property(x + h * e) - reg.calculate_property(x - h * e)) / (2 * h) for e in np.eye(9)])
    assert np.abs(g - fd).max() < 1e-7
