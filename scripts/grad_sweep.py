#!/usr/bin/env python
"""Times pm_grad on a small molecule for the current PM_SPEC_GRAD_* settings: `python scripts/grad_sweep.py [molecule] [n_conf]`."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
from paramol_b200.MM_engines.device import DeviceTopology, DeviceEnsemble
from paramol_b200.Utils.amber import system_from_prmtop, read_inpcrd
from paramol_b200.Utils.synthetic import make_ligand
G = os.path.join(REPO, "tests", "golden")
name = sys.argv[1] if len(sys.argv) > 1 else "aniline"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
if name.startswith("lig"):
    s, x0 = make_ligand(int(name[3:]))
else:
    s = system_from_prmtop(os.path.join(G, name + ".prmtop")); x0 = read_inpcrd(os.path.join(G, name + ".inpcrd"))
topo = DeviceTopology(s)
g = torch.Generator(device="cuda").manual_seed(1)
x0 = np.asarray(x0, dtype=np.float64).reshape(-1, 3)
X = torch.from_numpy(x0).cuda()[None] + 0.005 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
F = 50.0 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
ens = DeviceEnsemble(topo, X, F, torch.zeros(S, dtype=torch.float64, device="cuda"))
der = topo.prepare(torch.from_numpy(topo.slots_from_system(s)[None].copy()).cuda())
_, _, fmm = ens.evaluate(der, with_forces=True, want_fres=True, want_fmm=True)
ca = torch.randn(S, dtype=torch.float64, device="cuda", generator=g); cb = torch.rand(S, dtype=torch.float64, device="cuda", generator=g)
for _ in range(3): ens.term_gradient(der, ca, cb, fmm[0])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): ens.term_gradient(der, ca, cb, fmm[0])
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print("%s S=%d specialized=%s grad warps=%s fence=%s : %.4f ms  %.1f Mconf/s" % (name, S, topo.specialized, os.environ.get("PM_SPEC_GRAD_WARPS", "-"),
      os.environ.get("PM_SPEC_GRAD_FENCE", "-"), ms, S / ms / 1e3))
