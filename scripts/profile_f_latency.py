#!/usr/bin/env python
"""cProfile of `ObjectiveFunction.f` on BASELINE config 1 (aniline, the 10 stored conformations): where the ~130 us per call
go on the host.  Prints the top entries by cumulative and by own time."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from paramol_b200.MM_engines.b200_engine import B200Engine  # noqa: E402
from paramol_b200.System.system import ParaMolSystem  # noqa: E402
from paramol_b200.Parameter_space.parameter_space import ParameterSpace  # noqa: E402
from paramol_b200.Objective_function.objective_function import ObjectiveFunction  # noqa: E402
from paramol_b200.Objective_function.Properties.energy_property import EnergyProperty  # noqa: E402
from paramol_b200.Objective_function.Properties.force_property import ForceProperty  # noqa: E402
from paramol_b200.Objective_function.Properties.regularization import Regularization  # noqa: E402

golden = os.path.join(REPO, "tests", "golden")
engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(golden, "aniline.prmtop"),
                    crd_file=os.path.join(golden, "aniline.inpcrd"))
system = ParaMolSystem("aniline", engine, 14)
system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True, opt_charges=True, opt_lj=True, opt_sc=True)
system.read_data(os.path.join(golden, "aniline_10_struct.nc"))
ps = ParameterSpace()
_, x0 = ps.get_optimizable_parameters([system])
x0 = np.asarray(x0, dtype=np.float64)
ps.calculate_prior_widths("arithmetic")
e_prop, f_prop = EnergyProperty([system]), ForceProperty([system])
e_prop.calculate_variance()
f_prop.calculate_variance()
of = ObjectiveFunction(None, ps, [e_prop, f_prop, Regularization(x0.copy(), ps.prior_widths, "L2")], [system], checkpoint_freq=10 ** 9)
rng = np.random.default_rng(0)
n = 2000
trial = [x0 * (1.0 + 0.01 * rng.uniform(-1, 1, len(x0))) for _ in range(n)]
for x in trial[:50]:
    of.f(x)
t0 = time.perf_counter()
for x in trial:
    of.f(x)
print("us per f (no profiler): %.1f" % ((time.perf_counter() - t0) / n * 1e6))
pr = cProfile.Profile()
pr.enable()
for x in trial:
    of.f(x)
pr.disable()
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
    print(s.getvalue()[:6000])
