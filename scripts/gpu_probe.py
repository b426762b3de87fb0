#!/usr/bin/env python
"""Quick device probe: FMA-pipe peaks, rsqrt accuracy, and raw K1 timings for a few molecules (not the bench)."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from paramol_b200 import _lib  # noqa: E402
from paramol_b200.MM_engines.device import DeviceTopology, DeviceEnsemble  # noqa: E402
from paramol_b200.Utils.amber import system_from_prmtop, read_inpcrd  # noqa: E402

G = os.path.join(REPO, "tests", "golden")


def peak(fp64):
    L = _lib.lib()
    fl, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(L.pm_measure_fma_peak(0, fp64, 4000, ctypes.byref(fl), ctypes.byref(ms)))
    return fl.value, ms.value


def time_kernel(name, n_conf, n_vec=1, with_forces=True, reps=5):
    s = system_from_prmtop(os.path.join(G, name + ".prmtop"))
    x0 = read_inpcrd(os.path.join(G, name + ".inpcrd"))
    topo = DeviceTopology(s)
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.from_numpy(x0).cuda()[None] + 0.005 * torch.randn((n_conf,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    F = 50.0 * torch.randn((n_conf,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    E = torch.randn(n_conf, dtype=torch.float64, device="cuda", generator=g)
    ens = DeviceEnsemble(topo, X, F, E)
    slots = torch.from_numpy(np.tile(topo.slots_from_system(s), (n_vec, 1))).cuda()
    derived = topo.prepare(slots)
    for _ in range(2):
        ens.evaluate(derived, with_forces=with_forces)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record()
        emm, fres, _ = ens.evaluate(derived, with_forces=with_forces)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    info = topo.launch_info(with_forces)
    return dict(molecule=name, n_atoms=s.n_atoms, n_pairs=topo.n_pairs, n_conf=n_conf, n_vec=n_vec, with_forces=with_forces,
                ms=best, conf_evals_per_s=n_conf * n_vec / (best * 1e-3), **info)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs")
    for fp64 in (1, 0):
        fl, ms = peak(fp64)
        print("fma_peak fp64=%d: %.2f TFLOP/s (%.3f ms)" % (fp64, fl / 1e12, ms))
    runs = [("aniline", 1 << 20, 1, True), ("aniline", 1 << 20, 1, False), ("caffeine", 1 << 18, 1, True),
            ("norfloxacin", 1 << 18, 1, True), ("norfloxacin", 1 << 18, 1, False), ("aniline", 1 << 16, 16, True)]
    for r in runs:
        print(json.dumps(time_kernel(*r)))
