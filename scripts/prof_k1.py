#!/usr/bin/env python
"""Minimal K1 driver for ncu: `python scripts/prof_k1.py <molecule> <n_conf> [n_vec] [with_forces]` (3 launches)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from paramol_b200.MM_engines.device import DeviceTopology, DeviceEnsemble  # noqa: E402
from paramol_b200.Utils.amber import system_from_prmtop, read_inpcrd  # noqa: E402
from paramol_b200.Utils.synthetic import make_ligand  # noqa: E402

G = os.path.join(REPO, "tests", "golden")

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "aniline"
    n_conf = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
    n_vec = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    wf = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
    if name.startswith("lig"):
        s, x0 = make_ligand(int(name[3:]))
    else:
        s = system_from_prmtop(os.path.join(G, name + ".prmtop"))
        x0 = read_inpcrd(os.path.join(G, name + ".inpcrd"))
    topo = DeviceTopology(s)
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.from_numpy(x0).cuda()[None] + 0.005 * torch.randn((n_conf,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    F = 50.0 * torch.randn((n_conf,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    E = torch.randn(n_conf, dtype=torch.float64, device="cuda", generator=g)
    ens = DeviceEnsemble(topo, X, F, E)
    derived = topo.prepare(torch.from_numpy(np.tile(topo.slots_from_system(s), (n_vec, 1))).cuda())
    for _ in range(3):
        ens.evaluate(derived, with_forces=wf)
    torch.cuda.synchronize()
    print("ok", name, n_conf, topo.launch_info(wf))
