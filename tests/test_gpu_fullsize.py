"""
Full-size checks (``-m gpu``): BASELINE.json config 2 (aniline, 2^20 perturbed conformations) through size-independent
properties, plus the oracle on every conformation:
  * Newton's third law: the MM forces of every conformation sum to zero;
  * rigid-body invariance: rotating + translating every conformation leaves the energies unchanged and rotates the forces;
  * shard additivity: the partial rows of three uneven shards add up to the row of the whole set, and the assembled
    objective is the same (what the multi-GPU allreduce relies on);
  * determinism: two evaluations are bitwise identical;
  * ALL 2^20 conformations against the CPU oracle (energies and forces 1e-9, residual sums 1e-8).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

S_FULL = 1 << 20


def test_config2_full_size_properties(golden_dir):
    import torch
    from oracle import oracle as O
    from paramol_b200.MM_engines.device import DeviceTopology, DeviceEnsemble
    from paramol_b200.Utils.amber import system_from_prmtop
    from paramol_b200.Utils.netcdf_lite import read_paramol_data
    mm = system_from_prmtop(os.path.join(golden_dir, "aniline.prmtop"))
    X10 = torch.from_numpy(read_paramol_data(os.path.join(golden_dir, "aniline_10_struct.nc"))[0]).cuda()
    g = torch.Generator(device="cuda").manual_seed(1234)
    X = X10[torch.arange(S_FULL, device="cuda") % 10] + 0.005 * torch.randn((S_FULL, 14, 3), dtype=torch.float64, device="cuda",
                                                                            generator=g)
    topo = DeviceTopology(mm)
    slots = torch.from_numpy(topo.slots_from_system(mm)[None]).cuda()
    derived = topo.prepare(slots)
    Fref = 50.0 * torch.randn((S_FULL, 14, 3), dtype=torch.float64, device="cuda", generator=g)
    Eref = 5.0 * torch.randn(S_FULL, dtype=torch.float64, device="cuda", generator=g) - 43000.0
    w = torch.rand(S_FULL, dtype=torch.float64, device="cuda", generator=g) + 0.1
    w /= w.sum()
    ens = DeviceEnsemble(topo, X, Fref, Eref)
    ens.set_weights(w)
    emm, fres, fmm = ens.evaluate(derived, want_fmm=True)
    emm, fres = emm.clone(), fres.clone()
    F = ens.forces_from_tiles(fmm)[0]
    # Newton's third law on every conformation
    assert float(F.sum(dim=1).abs().max()) < 1e-8 * float(F.abs().max())
    # determinism
    emm2, fres2, _ = ens.evaluate(derived)
    assert torch.equal(emm, emm2) and torch.equal(fres, fres2)
    # EVERY conformation against the CPU oracle (the C port does millions of aniline conformations per second), in slabs so
    # that the host copies stay small; the residual sums against the oracle's forces as well
    slab = 1 << 17
    n_thr = max(1, min(O.max_threads(), 32))
    Fref_h_metric = np.tile(np.eye(3).ravel(), (14, 1))
    for lo in range(0, S_FULL, slab):
        hi = min(S_FULL, lo + slab)
        E_o, F_o = O.energies_forces(mm, X[lo:hi].cpu().numpy(), n_threads=n_thr)
        assert np.abs(emm[0, lo:hi].cpu().numpy() - E_o).max() < 1e-9 * np.abs(E_o).max(), lo
        assert np.abs(F[lo:hi].cpu().numpy() - F_o).max() < 1e-9 * np.abs(F_o).max(), lo
        res_o = O.force_residuals(F_o, Fref[lo:hi].cpu().numpy(), Fref_h_metric.reshape(14, 3, 3), n_threads=n_thr)
        assert np.abs(fres[0, lo:hi].cpu().numpy() - res_o).max() <= 1e-8 * np.abs(res_o).max(), lo
    # shard additivity of the partial rows
    d_shift = float(Eref.mean())
    whole = ens.reduce(emm, fres, d_shift=d_shift).cpu().numpy()[0]
    cuts = [0, 300_001, 777_777, S_FULL]
    acc = np.zeros(8)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        shard = DeviceEnsemble(topo, X[lo:hi].contiguous(), Fref[lo:hi].contiguous(), Eref[lo:hi].contiguous())
        shard.set_weights(w[lo:hi].contiguous())
        e_s, r_s, _ = shard.evaluate(derived)
        assert torch.equal(e_s[0], emm[0, lo:hi])  # the same conformation gives the same bits wherever it sits
        acc += shard.reduce(e_s, r_s, d_shift=d_shift).cpu().numpy()[0]
        del shard
    assert np.all(np.abs(acc - whole) <= 1e-11 * np.maximum(1.0, np.abs(whole))), (acc, whole)

    def objective(p):
        m = p[0] / p[5]
        return (p[2] - 2 * m * p[1] + m * m * p[3]) + p[4] / 42.0
    assert abs(objective(acc) - objective(whole)) <= 1e-11 * abs(objective(whole))
    # rigid-body invariance
    del ens, fmm
    th = 0.7
    R = torch.tensor([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64,
                     device="cuda")
    Rx = torch.tensor([[1.0, 0.0, 0.0], [0.0, np.cos(0.3), -np.sin(0.3)], [0.0, np.sin(0.3), np.cos(0.3)]], dtype=torch.float64,
                      device="cuda")
    R = R @ Rx
    X2 = X @ R.T + torch.tensor([0.3, -1.1, 2.0], dtype=torch.float64, device="cuda")
    ens2 = DeviceEnsemble(topo, X2)
    emm_r, _, fmm_r = ens2.evaluate(derived, want_fres=False, want_fmm=True)
    assert float((emm_r[0] - emm[0]).abs().max()) < 1e-9 * float(emm.abs().max())
    F_r = ens2.forces_from_tiles(fmm_r)[0]
    assert float((F_r - F @ R.T).abs().max()) < 1e-9 * float(F.abs().max())
