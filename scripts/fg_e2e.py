#!/usr/bin/env python
"""End-to-end time of ObjectiveFunction.f and f_and_grad on the headline workload (aniline, S conformations, host vectors)."""
import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
from paramol_b200.MM_engines.b200_engine import B200Engine
from paramol_b200.System.system import ParaMolSystem
from paramol_b200.Parameter_space.parameter_space import ParameterSpace
from paramol_b200.Objective_function.objective_function import ObjectiveFunction
from paramol_b200.Objective_function.Properties.energy_property import EnergyProperty
from paramol_b200.Objective_function.Properties.force_property import ForceProperty
from paramol_b200.Objective_function.Properties.regularization import Regularization
from paramol_b200.Utils.netcdf_lite import read_paramol_data
from paramol_b200.Utils.synthetic import perturbed_conformations
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
G = os.path.join(REPO, "tests", "golden")
X10, _, _ = read_paramol_data(os.path.join(G, "aniline_10_struct.nc"))
X = perturbed_conformations(X10, S, sigma=0.005, seed=1234)
rng = np.random.default_rng(5)
engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(G, "aniline.prmtop"), crd_file=os.path.join(G, "aniline.inpcrd"))
system = ParaMolSystem("aniline", engine, 14, ref_coordinates=X, ref_energies=rng.normal(-43000.0, 5.0, S), ref_forces=rng.normal(0, 50.0, (S, 14, 3)))
system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True, opt_charges=True, opt_lj=True)
ps = ParameterSpace(); _, x0 = ps.get_optimizable_parameters([system]); x0 = np.asarray(x0, float)
ps.calculate_prior_widths("arithmetic")
e_prop, f_prop = EnergyProperty([system]), ForceProperty([system], term_type="components")
e_prop.calculate_variance(); f_prop.calculate_variance()
of = ObjectiveFunction(None, ps, [e_prop, f_prop, Regularization(x0.copy(), ps.prior_widths, "L2")], [system], checkpoint_freq=10**9)
trial = [x0 * (1 + 0.01 * rng.uniform(-1, 1, len(x0))) for _ in range(60)]
for x in trial[:10]: of.f(x); of.f_and_grad(x)
torch.cuda.synchronize(); t0 = time.perf_counter()
for x in trial: of.f(x)
torch.cuda.synchronize(); tf = (time.perf_counter() - t0) / len(trial)
t0 = time.perf_counter()
for x in trial: of.f_and_grad(x)
torch.cuda.synchronize(); tg = (time.perf_counter() - t0) / len(trial)
print("S=%d P_opt=%d  f: %.3f ms   f_and_grad: %.3f ms  (2-point gradient through f_batch would cost about %.0f ms)" % (S, len(x0), tf * 1e3, tg * 1e3, (len(x0) + 1) * 0.66))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for x in trial[:30]: of.f_and_grad(x)
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(14)
