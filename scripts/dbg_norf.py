import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from paramol_b200.Utils.amber import system_from_prmtop, read_inpcrd
from paramol_b200.MM_engines.device import DeviceTopology, DeviceEnsemble
G = "tests/golden/"
name = sys.argv[1] if len(sys.argv) > 1 else "norfloxacin"
s = system_from_prmtop(G + name + ".prmtop"); x0 = read_inpcrd(G + name + ".inpcrd")
rng = np.random.default_rng(1234)
X = x0[None] + rng.normal(0, 0.005, (8,) + x0.shape)
def run(sys_):
    topo = DeviceTopology(sys_)
    ens = DeviceEnsemble(topo, X)
    der = topo.prepare(torch.from_numpy(topo.slots_from_system(sys_)[None]).cuda())
    emm, _, fmm = ens.evaluate(der, True, False, True)
    return emm[0].cpu().numpy(), ens.forces_from_tiles(fmm)[0].cpu().numpy(), topo
for what in ("bond", "angle", "torsion", "nb", "all"):
    t = s.copy()
    if what != "all":
        if what != "bond": t.bond_par[:, 1] = 0
        if what != "angle": t.angle_par[:, 1] = 0
        if what != "torsion": t.torsion_par[:, 1] = 0
        if what != "nb":
            t.particle_par[:, 0] = 0; t.particle_par[:, 2] = 0; t.exception_par[:, 0] = 0; t.exception_par[:, 2] = 0
    # keep exception activity pattern identical to the full system
    E, F = O.energies_forces(t, X)
    topo = DeviceTopology(s)
    ens = DeviceEnsemble(topo, X)
    der = topo.prepare(torch.from_numpy(topo.slots_from_system(t)[None]).cuda())
    emm, _, fmm = ens.evaluate(der, True, False, True)
    e = emm[0].cpu().numpy(); f = ens.forces_from_tiles(fmm)[0].cpu().numpy()
    print(what, "dE max", np.abs(e - E).max(), "dF max", np.abs(f - F).max(), "|F|max", np.abs(F).max())
    if what == "torsion" or what == "angle" or what=="bond":
        bad = np.argwhere(np.abs(f - F).max(axis=2) > 1e-6)
        print("  bad (conf, atom):", bad[:20].tolist())
print(topo.program_info())
