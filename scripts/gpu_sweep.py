#!/usr/bin/env python
"""Sweep lanes-per-conformation (PM_CPW) and warps (PM_WARPS) for the E+F kernel on several molecules; prints JSON lines."""
import json
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


def load(name):
    if name.startswith("lig"):
        return make_ligand(int(name[3:]))
    return system_from_prmtop(os.path.join(G, name + ".prmtop")), read_inpcrd(os.path.join(G, name + ".inpcrd"))


def run(name, n_conf, cpw, warps=None, with_forces=True, reps=5):
    s, x0 = load(name)
    os.environ["PM_CPW"] = str(cpw) if cpw else ""
    os.environ["PM_WARPS"] = str(warps) if warps else ""
    topo = DeviceTopology(s)
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.from_numpy(x0).cuda()[None] + 0.005 * torch.randn((n_conf,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    F = 50.0 * torch.randn((n_conf,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    E = torch.randn(n_conf, dtype=torch.float64, device="cuda", generator=g)
    ens = DeviceEnsemble(topo, X, F, E)
    derived = topo.prepare(torch.from_numpy(topo.slots_from_system(s)[None]).cuda())
    for _ in range(2):
        ens.evaluate(derived, with_forces=with_forces)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record()
        ens.evaluate(derived, with_forces=with_forces)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    info = topo.launch_info(with_forces)
    pi = topo.program_info()
    return dict(mol=name, N=s.n_atoms, S=n_conf, wf=with_forces, cpw=pi["conf_per_tile"], L=pi["lanes_per_conformation"],
                warps=info["warps_per_cta"], ms=round(best, 4), Mconf_s=round(n_conf / best / 1e3, 1),
                star_steps=pi["star_steps"], axis_steps=pi["axis_steps"], slots=pi["pair_entry_slots"], entries=pi["pair_entries"])


if __name__ == "__main__":
    mols = sys.argv[1].split(",") if len(sys.argv) > 1 else ["aniline", "caffeine", "norfloxacin", "lig50", "lig60"]
    sizes = {"aniline": 1 << 20, "caffeine": 1 << 19, "norfloxacin": 1 << 18, "lig50": 1 << 18, "lig60": 1 << 17}
    for m in mols:
        for cpw in (32, 16, 8, 4):
            try:
                print(json.dumps(run(m, sizes.get(m, 1 << 17), cpw)), flush=True)
            except Exception as exc:  # noqa: BLE001
                print(json.dumps(dict(mol=m, cpw=cpw, error=str(exc)[:200])), flush=True)
