"""Shared helpers of the LLS fixture tests (CPU: test_lls_reference_fixtures.py; GPU: test_gpu_resp_lls.py)."""
import os

import numpy as np

CASES = ["a10_eq_uniform", "a600_keq_uniform", "a600_eq_uniform", "a600_eq_boltzmann"]


def load_case(golden_dir, name):
    d = np.load(os.path.join(golden_dir, "lls_reference_fixtures.npz"))
    return {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith(name + "/")}


def build_system(golden_dir, fx):
    """The fixture's inputs as a ParaMolSystem + ParameterSpace (same construction as the generating script)."""
    from paramol_b200.MM_engines.b200_engine import B200Engine
    from paramol_b200.Parameter_space.parameter_space import ParameterSpace
    from paramol_b200.System.system import ParaMolSystem
    engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(golden_dir, "aniline.prmtop"),
                        crd_file=os.path.join(golden_dir, "aniline.inpcrd"))
    system = ParaMolSystem("aniline", engine, 14)
    system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True)
    if not int(fx["eq_opt"]):
        for force in ("HarmonicBondForce", "HarmonicAngleForce", "PeriodicTorsionForce"):
            for term in system.force_field.force_field[force][0]:
                for key, p in term.parameters.items():
                    if key in ("bond_eq", "angle_eq", "torsion_phase"):
                        p.optimize = 0
        system.force_field.create_force_field_optimizable()
    system.ref_coordinates = np.array(fx["coords"])
    system.ref_energies = np.array(fx["ref_energies"])
    system.n_structures = len(system.ref_coordinates)
    ps = ParameterSpace()
    ps.get_optimizable_parameters([system], symmetry_constrained=False)
    assert np.array_equal(np.asarray(ps.optimizable_parameters_values, dtype=np.float64), fx["x_before"])
    return system, ps


def check_design(fx, A, B=None, rtol_a=1e-12):
    """A (centred design matrix, numpy [S, M]) and B against what the reference built."""
    if "A" in fx:
        assert A.shape == fx["A"].shape
        assert np.abs(A - fx["A"]).max() <= rtol_a * max(1.0, np.abs(fx["A"]).max())
    else:
        assert np.abs(A[::12] - fx["A_rows"]).max() <= rtol_a * max(1.0, np.abs(fx["A_rows"]).max())
        assert np.abs(A.T @ A - fx["AtA"]).max() <= 10 * rtol_a * np.abs(fx["AtA"]).max()
        if B is not None:
            assert np.abs(A.T @ B - fx["AtB"]).max() <= 1e-9 * np.abs(fx["AtB"]).max()
    if B is not None:
        assert np.abs(B - fx["B"]).max() <= 1e-9 * max(1.0, np.abs(fx["B"]).max())
