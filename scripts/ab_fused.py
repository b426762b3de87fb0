#!/usr/bin/env python
"""A/B on one box: pm_eval + pm_reduce_objective against the fused pm_eval_reduce (interleaved rounds, CUDA events).
   python scripts/ab_fused.py [molecule] [n_conf] [n_vec]"""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
from paramol_b200.MM_engines.device import DeviceTopology, DeviceEnsemble
from paramol_b200.Utils.amber import system_from_prmtop, read_inpcrd
from paramol_b200.Utils.synthetic import make_ligand
G = os.path.join(REPO, "tests", "golden")
name = sys.argv[1] if len(sys.argv) > 1 else "aniline"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
P = int(sys.argv[3]) if len(sys.argv) > 3 else 1
if name.startswith("lig"):
    s, x0 = make_ligand(int(name[3:]))
else:
    s = system_from_prmtop(os.path.join(G, name + ".prmtop")); x0 = read_inpcrd(os.path.join(G, name + ".inpcrd"))
if os.environ.get("AB_OLD_IMAGE"):   # A/B against an image built by an earlier generator (E+F kernel only, forces in registers, 8 warps)
    from paramol_b200.csrc import specialize
    blob = open(os.environ["AB_OLD_IMAGE"], "rb").read()
    specialize.image_for = lambda *a, **k: (blob, 8, 0, 0, 0)
topo = DeviceTopology(s)
g = torch.Generator(device="cuda").manual_seed(1)
x0 = np.asarray(x0, dtype=np.float64).reshape(-1, 3)
X = torch.from_numpy(x0).cuda()[None] + 0.005 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
F = 50.0 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
ens = DeviceEnsemble(topo, X, F, torch.randn(S, dtype=torch.float64, device="cuda", generator=g))
ens.set_weights(np.full(S, 1.0 / S))
base = topo.slots_from_system(s)
rows = base[None] * (1.0 + 0.01 * np.random.default_rng(0).uniform(-1, 1, (P, len(base))))
der = topo.prepare(torch.from_numpy(rows).cuda())


def split():
    emm, fres, _ = ens.evaluate(der, with_forces=True, want_fres=True)
    return ens.reduce(emm, fres)


def fused():
    return ens.evaluate_reduce(der, with_forces=True)[0]


def timeit(fn, reps=30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for fn in (split, fused):
    for _ in range(3):
        fn()
for rnd in range(3):
    print("%s S=%d P=%d spec=%s round %d: split %.4f ms   fused %.4f ms" % (name, S, P, topo.specialized, rnd, timeit(split), timeit(fused)), flush=True)
a = split().cpu().numpy().copy(); b = fused().cpu().numpy()
print("max rel diff of the sums:", np.abs(a - b).max() / np.abs(a).max(), topo.launch_info(True))
