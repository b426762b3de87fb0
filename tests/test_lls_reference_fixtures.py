"""
CPU (``-m "not gpu"``): the oracle's LLS restatement (``oracle.lls_design_matrix`` / ``oracle.lls_fit``) and the host-side
column bookkeeping of ``paramol_b200.MM_engines.least_square`` against fixtures produced by RUNNING the reference's own
``ParaMol/MM_engines/least_square.py`` (``tests/golden/make_lls_fixtures.py`` -> ``tests/golden/lls_reference_fixtures.npz``).
This is what pins the LLS oracle; the GPU test ``tests/test_gpu_resp_lls.py::test_lls_matches_reference_fixtures`` compares the
device path with the same fixtures.
"""
import os

import numpy as np
import pytest

from lls_fixture_util import CASES, build_system, check_design, load_case


@pytest.mark.parametrize("name", CASES)
def test_oracle_lls_matches_reference_run(golden_dir, name):
    from oracle import oracle as O
    fx = load_case(golden_dir, name)
    system, ps = build_system(golden_dir, fx)
    A, p0, keys, init, symm = O.lls_design_matrix(system.ref_coordinates, ps)
    assert np.array_equal(np.array(keys), fx["keys"]) and np.array_equal(np.array(symm), fx["symm"])
    assert np.abs(p0 - fx["p0"]).max() <= 1e-13 and np.array_equal(init, fx["init"])
    zero = system.engine.system.copy()
    zero.bond_par[...] = 0.0
    zero.angle_par[...] = 0.0
    zero.torsion_par[...] = 0.0
    e_zero = O.energies_forces(zero, system.ref_coordinates, want_forces=False)[0]
    assert np.abs(e_zero - fx["mm_energies_zero"]).max() <= 1e-10 * max(1.0, np.abs(e_zero).max())
    B = system.ref_energies - e_zero
    B = B - B.mean()
    check_design(fx, A, B)
    w = O.conformation_weights(str(fx["weighting"]), len(B), ref_energies=system.ref_energies, temperature=300.0)
    assert np.abs(w * np.ones(len(B)) - fx["weights"]).max() <= 1e-15
    sc_dict, _ = ps.calculate_scaling_constants(method=None)
    pw_dict, _ = ps.calculate_prior_widths(method=None)
    sc = np.asarray([sc_dict[k] for k in keys])
    pw = np.asarray([pw_dict[k] for k in keys])
    got = O.lls_fit(A, B, w, system.ref_energies, sc, pw, init, symm, True, 1.0)
    want = fx["parameters"]
    # the stacked systems are ill conditioned (nearly collinear bracketing columns: kappa ~ 1e5-1e7), so rounding-level differences in
    # A move the solution by ~1e-8 relative; 1e-6 is far inside the reference's own reproducibility
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max(), np.abs(got - want).max() / np.abs(want).max()
    print(name, "relative parameter difference", np.abs(got - want).max() / np.abs(want).max())
