"""Quick GPU check of pm_grad against central differences of the kernel's own E and fres sums (slot space)."""
import numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
from paramol_b200.MM_engines.b200_engine import B200Engine
from paramol_b200.MM_engines.device import DeviceEnsemble
from paramol_b200.Utils.synthetic import make_ligand

def run(N, S):
    mm, x0 = make_ligand(N)
    engine = B200Engine(system=mm)
    topo = engine.device_topology()
    rng = np.random.default_rng(3)
    X = x0[None] + rng.normal(0, 0.01, (S,) + x0.shape)
    Fref = rng.normal(0, 50.0, (S,) + x0.shape)
    Eref = rng.normal(0, 5.0, S)
    ens = DeviceEnsemble(topo, X, Fref, Eref)
    metric = np.tile(np.eye(3).ravel(), (N, 1)) * rng.uniform(0.5, 2.0, (N, 1))
    metric[:, 1] += 0.1; metric[:, 3] += 0.05   # non-symmetric on purpose
    topo.set_force_metric(metric)
    base = engine.current_slots().copy()
    a = torch.from_numpy(rng.normal(0, 1.0, S)).cuda()
    b = torch.from_numpy(rng.uniform(0.1, 1.0, S)).cuda()
    def F(slots_rows):
        sd = torch.from_numpy(np.ascontiguousarray(slots_rows)).cuda()
        der = topo.prepare(sd)
        emm, fres, _ = ens.evaluate(der, with_forces=True, want_fres=True)
        return ((emm * a[None]).sum(1) + (fres * b[None]).sum(1)).cpu().numpy()
    sd = torch.from_numpy(base[None].copy()).cuda()
    der = topo.prepare(sd)
    emm, fres, fmm = ens.evaluate(der, with_forces=True, want_fres=True, want_fmm=True)
    gout = ens.term_gradient(der, a, b, fmm[0]).cpu().numpy()
    g = topo.term_gradient_to_slots(gout, base)
    # central differences on a subset of slots
    idx = rng.choice(len(base), size=min(len(base), 160), replace=False)
    h = 1e-6
    rows = np.repeat(base[None], 2 * len(idx), 0)
    for k, i in enumerate(idx):
        step = h * max(1.0, abs(base[i]))
        rows[2 * k, i] += step; rows[2 * k + 1, i] -= step
    vals = F(rows)
    worst = 0
    for k, i in enumerate(idx):
        step = h * max(1.0, abs(base[i]))
        fd = (vals[2 * k] - vals[2 * k + 1]) / (2 * step)
        err = abs(fd - g[i]) / max(1e-6 * np.abs(g).max(), abs(fd))
        if err > 1e-4:
            kind = "bond" if i < topo.off_angle else "angle" if i < topo.off_torsion else "tors" if i < topo.off_particle else "part" if i < topo.off_exception else "exc"
            print("  MISMATCH slot", i, kind, (i - [0, topo.off_angle, topo.off_torsion, topo.off_particle, topo.off_exception][["bond","angle","tors","part","exc"].index(kind)]) % (2 if kind in ("bond","angle","tors") else 3), "fd", fd, "an", g[i])
        worst = max(worst, err)
    print("N=%d S=%d n_slots=%d checked=%d worst rel err %.3e |g|max %.3e" % (N, S, len(base), len(idx), worst, np.abs(g).max()))

for N, S in ((12, 37), (24, 200), (60, 129)):
    run(N, S)
