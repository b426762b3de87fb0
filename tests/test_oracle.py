"""
Pins the CPU oracle (oracle/) against every golden the reference's own tests hold for the path.
Reference tests: ParaMol/Tests/System/test_system.py, Tests/Objective_function/test_objective_function.py,
Tests/MM_engines/test_openmm_wrapper.py (values extracted into tests/golden/reference_goldens.json).
"""
import os
import numpy as np

from oracle import oracle as O
from paramol_b200.Utils.amber import system_from_prmtop, read_inpcrd
from paramol_b200.Utils.openmm_xml import system_from_xml

OBJ = "test_objective_function.py::test_objective_function_and_properties::"


def test_energies_ensemble_golden(aniline_system, aniline_data, goldens):
    X, Eq, Fq = aniline_data
    E, _ = O.energies_forces(aniline_system, X, want_forces=False)
    # goldens are printed to 8 decimals; the reference asserts decimal=4 (test_system.py:109-110)
    assert np.abs(E - np.asarray(goldens["test_system.py::test_get_energies_ensemble::energies_to_compare"])).max() < 1e-8


def test_forces_golden(aniline_system, aniline_data, goldens):
    X, Eq, Fq = aniline_data
    _, F = O.energies_forces(aniline_system, X)
    assert np.abs(F[0] - np.asarray(goldens["test_system.py::test_get_forces_ensemble::forces_to_compare"])).max() < 1e-8


def test_reaction_field_goldens(golden_dir, goldens):
    # Tests/MM_engines/test_openmm_wrapper.py:43-58,71 are evaluated on aniline.xml (CutoffNonPeriodic + RF)
    s = system_from_xml(os.path.join(golden_dir, "aniline.xml"))
    x = read_inpcrd(os.path.join(golden_dir, "aniline.inpcrd"))
    E, F = O.energies_forces(s, x)
    assert abs(E[0] - goldens["test_openmm_wrapper.py::test_OpenMMEngine_get_potential_energy::energy_to_compare"]) < 1e-10
    assert np.abs(F[0] - np.asarray(goldens["test_openmm_wrapper.py::test_OpenMMEngine_get_forces::forces_to_compare"])).max() < 1e-8


def test_objective_goldens_bit_exact(aniline_system, aniline_data, goldens):
    X, Eq, Fq = aniline_data
    E, F = O.energies_forces(aniline_system, X)
    w = O.conformation_weights("uniform", 10)
    # reference tolerance is 1e-8 (test_objective_function.py:92,108,125); the restatement is bit-exact
    assert O.energy_property(Eq, E, w) == goldens[OBJ + "f_val"] == 0.16587870225911971
    assert O.force_property(Fq, F, w, "components") == goldens[OBJ + "f_val#1"] == 0.3022524295062905
    assert O.force_property(Fq, F, w, "norm") == goldens[OBJ + "f_val#2"] == 0.498237598004599
    total, parts = O.objective(aniline_system, X, Eq, Fq, force_term_type="norm")
    assert abs(total - (0.16587870225911971 + 0.498237598004599)) < 1e-15


def test_weights_goldens(aniline_system, aniline_data, goldens):
    X, Eq, Fq = aniline_data
    E, _ = O.energies_forces(aniline_system, X, want_forces=False)
    k = "test_system.py::test_weighting_methods::"
    np.testing.assert_almost_equal(O.conformation_weights("uniform", 10), goldens[k + "uniform_to_compare"])
    np.testing.assert_almost_equal(O.conformation_weights("boltzmann", 10, ref_energies=Eq), goldens[k + "boltzmann_to_compare"], decimal=6)
    np.testing.assert_almost_equal(O.conformation_weights("non_boltzmann", 10, ref_energies=Eq, emm=E),
                                   goldens[k + "non_boltzmann_to_compare"], decimal=6)
    # overflow guard -> uniform (system.py:215-221)
    w = O.conformation_weights("boltzmann", 3, ref_energies=np.array([0.0, 1e6, -1e6]))
    np.testing.assert_almost_equal(w, np.ones(3) / 3)


def test_c_oracle_matches_python_restatement_and_finite_differences(golden_dir):
    rng = np.random.default_rng(5)
    for name in ("aniline", "caffeine", "norfloxacin"):
        s = system_from_prmtop(os.path.join(golden_dir, name + ".prmtop"))
        x0 = read_inpcrd(os.path.join(golden_dir, name + ".inpcrd"))
        x = x0 + rng.normal(0, 0.005, x0.shape)
        E, F = O.energies_forces(s, x)
        assert abs(E[0] - O.energy_single_python(s, x)) < 1e-9 * max(1.0, abs(E[0]))
        h = 1e-6
        for (i, c) in [(0, 0), (s.n_atoms // 2, 1), (s.n_atoms - 1, 2)]:
            xp = x.copy(); xp[i, c] += h
            xm = x.copy(); xm[i, c] -= h
            fd = -(O.energies_forces(s, xp, False)[0][0] - O.energies_forces(s, xm, False)[0][0]) / (2 * h)
            assert abs(fd - F[0, i, c]) < 2e-5 * max(1.0, abs(fd))
        assert np.abs(F[0].sum(axis=0)).max() < 1e-8  # Newton's third law


def test_threads_do_not_change_results(aniline_system, aniline_data):
    X = np.tile(aniline_data[0], (13, 1, 1))
    E1, F1 = O.energies_forces(aniline_system, X, n_threads=1)
    E4, F4 = O.energies_forces(aniline_system, X, n_threads=4)
    assert np.array_equal(E1, E4) and np.array_equal(F1, F4)


def test_regularization_forms():
    x = np.array([1.0, -2.0, 0.5])
    x0 = np.array([0.5, -1.0, 0.5])
    w = np.array([0.5, 2.0, 1.0])
    assert abs(O.regularization(x, x0, w, "L2", a=2.0) - 2.0 * (1.0 + 0.25)) < 1e-15
    assert abs(O.regularization(x, x0, w, "L1", a=2.0) - 2.0 * (1.0 + 0.5)) < 1e-15
    assert abs(O.regularization(x, x0, w, "hyperbolic", a=1.0, b=0.1) - np.sum(np.sqrt(x * x + 0.01) - 0.1)) < 1e-15
    assert O.regularization(x0, x0, w, "L2") == 0.0  # test_objective_function.py:145
