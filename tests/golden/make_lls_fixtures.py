#!/usr/bin/env python
"""
Generates ``tests/golden/lls_reference_fixtures.npz`` by RUNNING THE REFERENCE'S OWN LLS CODE in the build container.

``/root/reference/ParaMol/MM_engines/least_square.py`` imports only numpy and ``ParaMol/Utils/geometry.py`` (reference
``least_square.py:1-4``), so -- unlike the OpenMM-backed modules -- it can be executed here: both files are loaded BY PATH
into a throw-away package (the reference's ``ParaMol/__init__.py``, which pulls in OpenMM, is never imported; nothing is copied
into the repository) and ``LinearLeastSquare.fit_parameters_lls`` (reference ``:68-146``, with ``_calculate_a`` ``:469-649`` and
``_calculate_b`` ``:651-692``) is driven with

  * this repository's ``ParameterSpace`` / ``ForceField`` (the duck-typed ``parameter_space`` the reference class asks for:
    ``optimizable_parameters`` with ``ff_term`` / ``param_key`` / ``value`` / ``symmetry_group``, ``update_systems``,
    ``calculate_prior_widths``, ``calculate_scaling_constants``, ``get_optimizable_parameters``; their own parity is pinned by
    the 232-entry golden vector of ``Tests/Parameter_space/test_parameter_space.py``), and
  * a system object whose ``get_energies_ensemble`` is the pinned CPU oracle (``oracle/oracle.c``), i.e. the OpenMM
    Reference-platform arithmetic the reference would call (``ParaMol/System/system.py:340-354``).

Stored per case: the input conformations and reference energies, the centred design matrix A the reference built (whole for
the 10-conformation case, every 12th row plus A^T A and A^T B for the 600-conformation cases), B, the column bookkeeping
(p0, keys, symmetry tags) and the fitted parameter vector (``_parameters``) and the reconstructed optimizable values.
``tests/test_lls_reference_fixtures.py`` checks the numpy restatement against them on the CPU and
``tests/test_gpu_resp_lls.py::test_lls_matches_reference_fixtures`` checks the device LLS against them on the GPU.

Run from the repository root:  python tests/golden/make_lls_fixtures.py
"""
import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
REFERENCE = os.environ.get("PARAMOL_REFERENCE", "/root/reference")


def load_reference_lls():
    """The reference's LinearLeastSquare class, loaded by file path into a scratch package (no OpenMM import)."""
    pkg = types.ModuleType("_paramol_ref")
    pkg.__path__ = []
    sys.modules["_paramol_ref"] = pkg
    for sub in ("Utils", "MM_engines"):
        m = types.ModuleType("_paramol_ref." + sub)
        m.__path__ = []
        sys.modules["_paramol_ref." + sub] = m
        setattr(pkg, sub, m)

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE, "ParaMol", rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    load("_paramol_ref.Utils.geometry", os.path.join("Utils", "geometry.py"))
    return load("_paramol_ref.MM_engines.least_square", os.path.join("MM_engines", "least_square.py")).LinearLeastSquare


class OracleSystem:
    """Duck-typed ParaMolSystem for the reference LLS: data attributes + weights + oracle-backed energies."""

    def __init__(self, paramol_system, coords, ref_energies):
        from oracle import oracle as O
        self._O = O
        self._ps = paramol_system
        self.engine = paramol_system.engine
        self.force_field = paramol_system.force_field
        self.name = paramol_system.name
        self.n_atoms = paramol_system.n_atoms
        self.ref_coordinates = np.asarray(coords, dtype=np.float64)
        self.ref_energies = np.asarray(ref_energies, dtype=np.float64)
        self.n_structures = len(self.ref_coordinates)
        self.weights = 1.0
        self.wham_weights = 1.0
        self.resp_engine = None

    def get_energies_ensemble(self):
        # system.py:340-354 -> OpenMM Reference platform == the pinned oracle
        return self._O.energies_forces(self.engine.system, self.ref_coordinates, want_forces=False, n_threads=4)[0]

    def compute_conformations_weights(self, temperature=None, emm=None, weighting_method="UNIFORM"):
        # system.py:165-248 through the oracle's restatement (pinned by the reference's three weight goldens)
        self.weights = self._O.conformation_weights(weighting_method.lower(), self.n_structures, ref_energies=self.ref_energies,
                                                    emm=emm, temperature=float(getattr(temperature, "_value", temperature or 300.0)))
        return self.weights


def build_case(n_conf, eq_opt, weighting, seed):
    from oracle import oracle as O
    from paramol_b200.MM_engines.b200_engine import B200Engine
    from paramol_b200.Parameter_space.parameter_space import ParameterSpace
    from paramol_b200.System.system import ParaMolSystem
    from paramol_b200.Utils.netcdf_lite import read_paramol_data
    from paramol_b200.Utils.synthetic import perturbed_conformations
    engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(HERE, "aniline.prmtop"),
                        crd_file=os.path.join(HERE, "aniline.inpcrd"))
    system = ParaMolSystem("aniline", engine, 14)
    system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True)
    if not eq_opt:
        for force in ("HarmonicBondForce", "HarmonicAngleForce", "PeriodicTorsionForce"):
            for term in system.force_field.force_field[force][0]:
                for key, p in term.parameters.items():
                    if key in ("bond_eq", "angle_eq", "torsion_phase"):
                        p.optimize = 0
        system.force_field.create_force_field_optimizable()
    X10, E10, _ = read_paramol_data(os.path.join(HERE, "aniline_10_struct.nc"))
    if n_conf == 10:
        X, Eref = X10, E10
    else:
        X = perturbed_conformations(X10, n_conf, sigma=0.004, seed=seed)
        E = O.energies_forces(engine.system, X, want_forces=False, n_threads=4)[0]
        Eref = E + np.random.default_rng(seed + 1).normal(0, 2.0, n_conf) - 43000.0
    duck = OracleSystem(system, X, Eref)
    ps = ParameterSpace()
    ps.get_optimizable_parameters([duck], symmetry_constrained=False)
    return duck, ps


def main():
    RefLLS = load_reference_lls()
    out = {}
    cases = [("a10_eq_uniform", 10, True, "uniform", 5), ("a600_keq_uniform", 600, False, "uniform", 5),
             ("a600_eq_uniform", 600, True, "uniform", 5), ("a600_eq_boltzmann", 600, True, "boltzmann", 7)]
    for name, n_conf, eq_opt, weighting, seed in cases:
        duck, ps = build_case(n_conf, eq_opt, weighting, seed)
        x_before = np.asarray(ps.optimizable_parameters_values, dtype=np.float64).copy()
        lls = RefLLS(ps, True, "L2", 1.0, 0.01, weighting, 300.0)
        with contextlib.redirect_stdout(io.StringIO()):
            # the pieces first (on copies of the state), then the whole fit exactly as LLSFitting.run_task calls it
            A = np.array(lls._calculate_a(duck, 0.05, 0.05), copy=True)
            B = np.array(lls._calculate_b(duck), copy=True)
            p0, keys, symm = np.asarray(lls._p0, dtype=np.float64), list(lls._param_keys_list), list(lls._param_symmetries_list)
            init = np.asarray(lls._initial_param_regularization, dtype=np.float64)
            e_zero = np.asarray(lls.mm_energies_zero, dtype=np.float64)
            fitted_values = np.asarray(lls.fit_parameters_lls([duck], 0.05, 0.05), dtype=np.float64)
        out[name + "/coords"] = duck.ref_coordinates
        out[name + "/ref_energies"] = duck.ref_energies
        out[name + "/eq_opt"] = np.array(int(eq_opt))
        out[name + "/weighting"] = np.array(weighting)
        out[name + "/x_before"] = x_before
        if n_conf <= 10:
            out[name + "/A"] = A
        else:
            out[name + "/A_rows"] = A[::12]
            out[name + "/AtA"] = A.T @ A
            out[name + "/AtB"] = A.T @ B
        out[name + "/B"] = B
        out[name + "/mm_energies_zero"] = e_zero
        out[name + "/p0"] = p0
        out[name + "/keys"] = np.array(keys)
        out[name + "/symm"] = np.array(symm)
        out[name + "/init"] = init
        out[name + "/weights"] = np.asarray(duck.weights, dtype=np.float64) * np.ones(n_conf)
        out[name + "/parameters"] = np.asarray(lls._parameters, dtype=np.float64)
        out[name + "/fitted_values"] = fitted_values
        print("%-20s S=%4d  M=%3d  |params|max=%.4g" % (name, n_conf, A.shape[1], np.abs(lls._parameters).max()))
    path = os.path.join(HERE, "lls_reference_fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
