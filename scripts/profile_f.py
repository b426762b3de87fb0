#!/usr/bin/env python
"""cProfile of ObjectiveFunction.f on a small ensemble (host-side overhead per call)."""
import cProfile, pstats, os, sys, io
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paramol_b200.MM_engines.b200_engine import B200Engine
from paramol_b200.System.system import ParaMolSystem
from paramol_b200.Parameter_space.parameter_space import ParameterSpace
from paramol_b200.Objective_function.objective_function import ObjectiveFunction
from paramol_b200.Objective_function.Properties.energy_property import EnergyProperty
from paramol_b200.Objective_function.Properties.force_property import ForceProperty
from paramol_b200.Objective_function.Properties.regularization import Regularization
from paramol_b200.Utils.netcdf_lite import read_paramol_data
from paramol_b200.Utils.synthetic import perturbed_conformations
G = "tests/golden/"
e = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=G + "aniline.prmtop", crd_file=G + "aniline.inpcrd")
s = ParaMolSystem("aniline", e, 14)
s.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True, opt_charges=True, opt_lj=True, opt_sc=True)
X10 = read_paramol_data(G + "aniline_10_struct.nc")[0]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
s.ref_coordinates = perturbed_conformations(X10, n); s.n_structures = n
rng = np.random.default_rng(0)
s.ref_energies = rng.normal(0, 5, n); s.ref_forces = rng.normal(0, 100, (n, 14, 3))
ps = ParameterSpace(); _, x0 = ps.get_optimizable_parameters([s]); x0 = np.asarray(x0)
ep, fp = EnergyProperty([s]), ForceProperty([s]); ep.calculate_variance(); fp.calculate_variance()
ps.calculate_prior_widths("default")
rp = Regularization(x0.copy(), ps.prior_widths, "L2")
of = ObjectiveFunction(None, ps, [ep, fp, rp], [s], checkpoint_freq=10 ** 12)
for _ in range(20): of.f(x0.copy())
import time
t0 = time.perf_counter()
for _ in range(500): of.f(x0.copy())
print("us per f() call (S=%d): %.1f" % (n, 1e6 * (time.perf_counter() - t0) / 500))
pr = cProfile.Profile(); pr.enable()
for _ in range(500): of.f(x0.copy())
pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(22); print(st.getvalue()[:4500])
