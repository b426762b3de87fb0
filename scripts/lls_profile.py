#!/usr/bin/env python
"""Where the time of one LLS fit goes (config 4: lig60, 200k conformations, 742 columns): python scripts/lls_profile.py [S]"""
import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
from paramol_b200.MM_engines.b200_engine import B200Engine
from paramol_b200.MM_engines.least_square import LinearLeastSquare
from paramol_b200.Parameter_space.parameter_space import ParameterSpace
from paramol_b200.System.system import ParaMolSystem
from paramol_b200.Utils.synthetic import make_ligand
S = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
mm, x0 = make_ligand(60)
engine = B200Engine(system=mm)
system = ParaMolSystem("lig60", engine, 60)
system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True)
rng = np.random.default_rng(2)
system.ref_coordinates = x0[None] + rng.normal(0, 0.005, (S,) + x0.shape)
system.n_structures = S
system.ref_energies = system.get_energies_ensemble() + rng.normal(0, 3.0, S)
ps = ParameterSpace()
ps.get_optimizable_parameters([system], symmetry_constrained=False)
lls = LinearLeastSquare(ps, True, method="L2", scaling_factor=1.0, hyperbolic_beta=0.01, weighting_method="uniform", weighting_temperature=300.0)
x_start = np.asarray(ps.optimizable_parameters_values, dtype=np.float64).copy()
lls.fit_parameters_lls([system])
ps.update_systems([system], x_start)
import cProfile, pstats
pr = cProfile.Profile()
torch.cuda.synchronize()
t0 = time.perf_counter()
pr.enable()
lls.fit_parameters_lls([system])
torch.cuda.synchronize()
pr.disable()
print("fit %.1f ms, refinement steps %d, reduction %s, timings %s" % (1e3 * (time.perf_counter() - t0), lls.refinement_steps, lls.reduction,
      {k: round(1e3 * v, 2) for k, v in lls.timings.items()}))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
plan = lls._plan
ev = lambda: torch.cuda.Event(enable_timing=True)
def timed(fn, n=5):
    fn(); torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
rs = torch.full((S,), 1e-3, dtype=torch.float64, device="cuda"); B = torch.randn(S, dtype=torch.float64, device="cuda")
z = torch.randn(plan.n_cols, dtype=torch.float64, device="cuda")
print("kernels: normal %.3f ms   colstats %.3f ms   residual %.3f ms" % (timed(lambda: plan.normal(rs, B)), timed(lambda: plan.colstats(B, rs)), timed(lambda: plan.residual(z, B))))
