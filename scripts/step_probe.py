#!/usr/bin/env python
"""Where the device step goes beyond the E+F kernel: CUDA-event timings of launch sequences on aniline x S conformations.
`python scripts/step_probe.py [S]`"""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
from paramol_b200.MM_engines.device import DeviceTopology, DeviceEnsemble
from paramol_b200.Utils.amber import system_from_prmtop, read_inpcrd
G = os.path.join(REPO, "tests", "golden")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
s = system_from_prmtop(os.path.join(G, "aniline.prmtop")); x0 = np.asarray(read_inpcrd(os.path.join(G, "aniline.inpcrd"))).reshape(-1, 3)
topo = DeviceTopology(s)
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.from_numpy(x0).cuda()[None] + 0.005 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
F = 50.0 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
ens = DeviceEnsemble(topo, X, F, torch.zeros(S, dtype=torch.float64, device="cuda"))
slots = torch.from_numpy(topo.slots_from_system(s)[None].copy()).cuda()
der = topo.prepare(slots)

def timed(fn, n=100, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def ef(): ens.evaluate(der, with_forces=True, want_fres=True)
def prep(): topo.prepare(slots, der)
def red():
    emm, fres, _ = ens.evaluate(der, with_forces=True, want_fres=True)
def evalred(): ens.evaluate_reduce(der, with_forces=True, d_shift=0.1)
def step(): topo.prepare(slots, der); ens.evaluate_reduce(der, with_forces=True, d_shift=0.1)
emm, fres, _ = ens.evaluate(der, with_forces=True, want_fres=True)
def redonly(): ens.reduce(emm, fres, d_shift=0.1)
for rep in range(2):
    print("S=%d  EF (copy+kernel) %.1f us | prepare %.1f | reduce (stage1+stage2) %.1f | eval_reduce %.1f | prepare+eval_reduce %.1f"
          % (S, timed(ef), timed(prep), timed(redonly), timed(evalred), timed(step)), flush=True)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    step()
print("graph of prepare+eval_reduce: %.1f us per replay" % timed(gr.replay))
