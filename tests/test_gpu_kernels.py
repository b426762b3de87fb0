"""
GPU parity tests proper (run with ``-m gpu`` on the B200 box): the CUDA path, called through the C ABI
(``include/paramol_b200.h`` via ``paramol_b200._lib``), against the CPU oracle on the same seeded inputs.

Tolerances are the north-star ones: energies/forces 1e-6 relative with a 1e-4 kJ/mol(/nm) absolute floor,
residual sums / objective 1e-8 relative, pair and exclusion lists bit-exact.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

E_RTOL, E_ATOL = 1e-6, 1e-4
OBJ_RTOL = 1e-8


def _close_ef(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return np.all(np.abs(a - b) <= np.maximum(E_RTOL * np.abs(b), E_ATOL))


def _system(golden_dir, name):
    from paramol_b200.Utils.amber import system_from_prmtop, read_inpcrd
    if name.startswith("lig"):   # synthetic drug-like ligands of BASELINE.json configs 3-5 (8 lanes per conformation at 50-60 atoms)
        from paramol_b200.Utils.synthetic import make_ligand
        return make_ligand(int(name[3:]))
    return (system_from_prmtop(os.path.join(golden_dir, name + ".prmtop")),
            read_inpcrd(os.path.join(golden_dir, name + ".inpcrd")))


def _run(sys_, X, Fref=None, Eref=None, metric=None, slots=None, conf_per_tile=None):
    import torch
    from paramol_b200.MM_engines.device import DeviceTopology, DeviceEnsemble
    topo = DeviceTopology(sys_, conf_per_tile=conf_per_tile)
    if metric is not None:
        topo.set_force_metric(metric)
    ens = DeviceEnsemble(topo, X, Fref, Eref)
    if slots is None:
        slots = topo.slots_from_system(sys_)[None, :]
    slots_dev = torch.from_numpy(np.ascontiguousarray(slots)).cuda()
    derived = topo.prepare(slots_dev)
    emm, fres, fmm = ens.evaluate(derived, with_forces=True, want_fres=Fref is not None, want_fmm=True)
    F = ens.forces_from_tiles(fmm).cpu().numpy()
    out = dict(emm=emm.cpu().numpy(), fmm=F, fres=None if fres is None else fres.cpu().numpy(), topo=topo, ens=ens,
               derived=derived)
    emm_only, _, _ = ens.evaluate(derived, with_forces=False)
    out["emm_energy_only_kernel"] = emm_only.cpu().numpy()
    torch.cuda.synchronize()
    return out


def test_pair_list_bit_exact(golden_dir):
    from paramol_b200.MM_engines.device import DeviceTopology
    for name in ("aniline", "caffeine", "norfloxacin"):
        s, _ = _system(golden_dir, name)
        topo = DeviceTopology(s)
        pairs = topo.pairs()
        plain = {tuple(p) for p in s.nonexcluded_pairs().tolist()}
        active = {(int(min(a, b)), int(max(a, b))): e for e, ((a, b), par) in
                  enumerate(zip(s.exception_idx, s.exception_par)) if par[0] != 0.0 or par[2] != 0.0}
        got_plain = {(int(i), int(j)) for i, j, e in pairs if e < 0}
        got_exc = {(int(i), int(j)): int(e) for i, j, e in pairs if e >= 0}
        assert got_plain == plain
        assert got_exc == active
        assert len(pairs) == len(plain) + len(active)  # no duplicates


def test_aniline_goldens_through_cuda(aniline_system, aniline_data, goldens):
    X, Eq, Fq = aniline_data
    r = _run(aniline_system, X, Fq, Eq)
    gold_e = np.asarray(goldens["test_system.py::test_get_energies_ensemble::energies_to_compare"])
    gold_f = np.asarray(goldens["test_system.py::test_get_forces_ensemble::forces_to_compare"])
    # the reference's own tolerance is decimal=4 (Tests/System/test_system.py:109-142)
    assert np.abs(r["emm"][0] - gold_e).max() < 1e-7
    assert np.abs(r["fmm"][0, 0] - gold_f).max() < 1e-7
    assert np.abs(r["emm_energy_only_kernel"][0] - gold_e).max() < 1e-7


@pytest.mark.parametrize("name,n_conf", [("aniline", 1000), ("caffeine", 257), ("norfloxacin", 96), ("ethane", 33),
                                         ("lig50", 203), ("lig60", 131), ("lig33", 77), ("lig45", 64)])
def test_ef_parity_vs_oracle(golden_dir, name, n_conf):
    from oracle import oracle as O
    s, x0 = _system(golden_dir, name)
    rng = np.random.default_rng(1234)
    X = x0[None] + rng.normal(0, 0.005, (n_conf,) + x0.shape)
    E, F = O.energies_forces(s, X, n_threads=4)
    Fref = F + rng.normal(0, 50.0, F.shape)
    metric = np.tile(np.eye(3).ravel(), (s.n_atoms, 1)) * rng.uniform(0.5, 2.0, (s.n_atoms, 1))
    r = _run(s, X, Fref, E + rng.normal(0, 5, E.shape), metric)
    assert _close_ef(r["emm"][0], E), np.abs(r["emm"][0] - E).max()
    assert _close_ef(r["fmm"][0], F), np.abs(r["fmm"][0] - F).max()
    assert _close_ef(r["emm_energy_only_kernel"][0], E)
    # far tighter than the bar in practice: FP64 evaluation end to end
    assert np.abs(r["emm"][0] - E).max() < 1e-9 * max(1.0, np.abs(E).max())
    assert np.abs(r["fmm"][0] - F).max() < 1e-9 * np.abs(F).max()
    d = F - Fref
    res = np.einsum("sna,nab,snb->s", d, metric.reshape(-1, 3, 3), d)
    assert np.abs(r["fres"][0] - res).max() <= OBJ_RTOL * np.abs(res).max()


def _system_with_slots(s, row):
    sp = s.copy()
    o = 0
    for arr in (sp.bond_par, sp.angle_par, sp.torsion_par, sp.particle_par, sp.exception_par):
        arr[...] = row[o:o + arr.size].reshape(arr.shape)
        o += arr.size
    return sp


def test_config3_shape_batched_vectors_vs_oracle():
    """BASELINE.json config 3 in shape (60-atom ligand, a batch of perturbed parameter vectors per launch; here 9 vectors x 4100
    conformations, a ragged last tile): per-vector energies, forces and residual sums against the oracle evaluated with that
    vector's parameters.  This is the 8-lanes-per-conformation path with the parameter-vector loop inside the launch."""
    import torch
    from oracle import oracle as O
    from paramol_b200.MM_engines.device import DeviceTopology, DeviceEnsemble
    from paramol_b200.Utils.synthetic import make_ligand
    s, x0 = make_ligand(60)
    rng = np.random.default_rng(60)
    n_conf, n_vec = 4100, 9
    X = x0[None] + rng.normal(0, 0.005, (n_conf,) + x0.shape)
    E0, F0 = O.energies_forces(s, X, n_threads=4)
    Eref = E0 + rng.normal(0, 5, n_conf)
    Fref = F0 + rng.normal(0, 50, F0.shape)
    metric = np.tile(np.eye(3).ravel(), (s.n_atoms, 1)) * rng.uniform(0.5, 2.0, (s.n_atoms, 1))
    topo = DeviceTopology(s)
    topo.set_force_metric(metric)
    base = topo.slots_from_system(s)
    slots = base[None] * (1.0 + 0.02 * rng.uniform(-1, 1, (n_vec, len(base))))
    slots[0] = base
    ens = DeviceEnsemble(topo, X, Fref, Eref)
    derived = topo.prepare(torch.from_numpy(slots).cuda())
    emm, fres, fmm = ens.evaluate(derived, want_fmm=True)
    emm_h, fres_h = emm.cpu().numpy(), fres.cpu().numpy()
    Fmm = ens.forces_from_tiles(fmm).cpu().numpy()
    d_shift = float(np.mean(Eref))
    partials = ens.reduce(emm, fres, d_shift=d_shift).cpu().numpy()
    e_only = ens.evaluate(derived, with_forces=False)[0].cpu().numpy()
    for p in range(n_vec):
        E, F = O.energies_forces(_system_with_slots(s, slots[p]), X, n_threads=4)
        assert np.abs(emm_h[p] - E).max() < 1e-9 * max(1.0, np.abs(E).max()), p
        assert np.abs(e_only[p] - E).max() < 1e-9 * max(1.0, np.abs(E).max()), p
        assert np.abs(Fmm[p] - F).max() < 1e-9 * np.abs(F).max(), p
        d = F - Fref
        res = np.einsum("sna,nab,snb->s", d, metric.reshape(-1, 3, 3), d)
        assert np.abs(fres_h[p] - res).max() <= OBJ_RTOL * np.abs(res).max(), p
        dd = (Eref - E) - d_shift
        want = np.array([dd.sum(), dd.sum(), (dd * dd).sum(), n_conf, res.sum(), n_conf, 0, 0])
        assert np.all(np.abs(partials[p] - want) <= 1e-9 * np.maximum(1.0, np.abs(want))), (p, partials[p], want)


def test_batched_parameter_vectors_and_reduction(golden_dir):
    import torch
    from oracle import oracle as O
    s, x0 = _system(golden_dir, "aniline")
    rng = np.random.default_rng(7)
    n_conf, n_vec = 333, 5
    X = x0[None] + rng.normal(0, 0.004, (n_conf,) + x0.shape)
    E0, F0 = O.energies_forces(s, X)
    Eref = E0 + rng.normal(0, 5, n_conf) - 43000.0
    Fref = F0 + rng.normal(0, 50, F0.shape)
    w = rng.uniform(0.1, 1.0, n_conf)
    w /= w.sum()
    from paramol_b200.MM_engines.device import DeviceTopology, DeviceEnsemble
    topo = DeviceTopology(s)
    base = topo.slots_from_system(s)
    slots = np.stack([base * (1.0 + 0.02 * rng.uniform(-1, 1, base.shape)) for _ in range(n_vec)])
    slots[0] = base
    ens = DeviceEnsemble(topo, X, Fref, Eref)
    ens.set_weights(w)
    derived = topo.prepare(torch.from_numpy(slots).cuda())
    emm, fres, _ = ens.evaluate(derived)
    d_shift = float(np.mean(Eref))
    partials = ens.reduce(emm, fres, d_shift=d_shift).cpu().numpy()
    for p in range(n_vec):
        sp = s.copy()
        o = 0
        for arr in (sp.bond_par, sp.angle_par, sp.torsion_par, sp.particle_par, sp.exception_par):
            arr[...] = slots[p, o:o + arr.size].reshape(arr.shape)
            o += arr.size
        E, F = O.energies_forces(sp, X)
        assert _close_ef(emm[p].cpu().numpy(), E)
        d = (Eref - E) - d_shift
        res = np.sum((F - Fref) ** 2, axis=(1, 2))
        want = np.array([d.sum(), (w * d).sum(), (w * d * d).sum(), w.sum(), (w * res).sum(), n_conf, 0, 0])
        assert np.all(np.abs(partials[p] - want) <= 1e-10 * np.maximum(1.0, np.abs(want))), (partials[p], want)
        # objective assembled from the partial sums == oracle property values (1e-8 relative)
        m = partials[p, 0] / partials[p, 5]
        energy_obj = (partials[p, 2] - 2 * m * partials[p, 1] + m * m * partials[p, 3]) / np.var(Eref)
        assert abs(energy_obj - O.energy_property(Eref, E, w)) <= OBJ_RTOL * abs(energy_obj)
        force_obj = partials[p, 4] / (3 * s.n_atoms)
        var1 = np.ones(s.n_atoms)
        want_f = np.sum(w * np.sum(np.sum((F - Fref) ** 2, axis=2) / var1, axis=1)) / (3 * s.n_atoms)
        assert abs(force_obj - want_f) <= OBJ_RTOL * abs(want_f)


def test_ragged_tail_and_single_conformation(golden_dir):
    from oracle import oracle as O
    s, x0 = _system(golden_dir, "aniline")
    rng = np.random.default_rng(3)
    for n_conf in (1, 31, 32, 33, 65):
        X = x0[None] + rng.normal(0, 0.005, (n_conf,) + x0.shape)
        E, F = O.energies_forces(s, X)
        r = _run(s, X, F, E)
        assert r["emm"].shape == (1, n_conf) and r["fmm"].shape == (1, n_conf, s.n_atoms, 3)
        assert np.abs(r["emm"][0] - E).max() < 1e-9 * np.abs(E).max()
        assert np.abs(r["fres"][0]).max() < 1e-12 * np.abs(F).max() ** 2


def test_fma_peak_probe():
    import ctypes
    from paramol_b200 import _lib
    L = _lib.lib()
    fl, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(L.pm_measure_fma_peak(0, 1, 2000, ctypes.byref(fl), ctypes.byref(ms)))
    assert fl.value > 1e12  # a B200 sustains tens of TFLOP/s of DFMA


@pytest.mark.parametrize("n_atoms,shape", [(5, "auto"), (12, "auto"), (12, "0"), (12, "top6"), (15, "auto"), (16, "auto"), (16, "0"),
                                           (20, "auto")])
def test_specialized_kernel_equals_generic(n_atoms, shape, monkeypatch):
    """The topology-specialised kernel (csrc/specialize.py, installed automatically for small molecules) against the generic
    program-interpreting kernel on the same tiles: energies, residuals and forces to 1e-11 of the largest entry (a few of the 32805 random conformations are near-singular), several parameter vectors per
    call, a partial last tile, non-trivial metric; energy-only calls keep using the generic kernel."""
    import torch
    from paramol_b200.MM_engines.b200_engine import B200Engine
    from paramol_b200.MM_engines.device import DeviceEnsemble, DeviceTopology
    from paramol_b200.Utils.synthetic import make_ligand
    mm, x0 = make_ligand(n_atoms)
    monkeypatch.setenv("PM_CPW", "32")
    monkeypatch.setenv("PARAMOL_B200_SPECIALIZE", "0")
    generic = DeviceTopology(mm)
    monkeypatch.setenv("PARAMOL_B200_SPECIALIZE", "auto")
    # shapes of the generated E+F kernel: forces in registers ("auto" picks it when it fits), in the shared force tile ("0"),
    # or the busiest atoms in registers and the rest in a packed tile ("top6")
    monkeypatch.setenv("PM_SPEC_FREG", shape)
    spec = DeviceTopology(mm)
    assert not generic.specialized and spec.specialized
    rng = np.random.default_rng(n_atoms)
    n_conf = 32768 + 37     # batches over fewer than 32768 conformations stay on the generic kernel (launch-bound)
    X = x0[None] + rng.normal(0, 0.01, (n_conf,) + x0.shape)
    F = rng.normal(0, 50.0, X.shape)
    metric = np.tile(np.eye(3).ravel(), (n_atoms, 1)) * rng.uniform(0.5, 2.0, (n_atoms, 1))
    metric[:, 1] += 0.1
    base = B200Engine(system=mm).current_slots() if False else generic.slots_from_system(mm)
    rows = base[None] * (1.0 + 0.02 * rng.uniform(-1, 1, (3, len(base))))
    out = []
    for topo in (generic, spec):
        topo.set_force_metric(metric)
        ens = DeviceEnsemble(topo, X, F, np.zeros(n_conf))
        der = topo.prepare(torch.from_numpy(rows).cuda())
        emm, fres, fmm = ens.evaluate(der, with_forces=True, want_fres=True, want_fmm=True)
        forces = ens.forces_from_tiles(fmm)
        e_only = ens.evaluate(der, with_forces=False, want_fres=False)[0].clone()
        out.append((emm.clone(), fres.clone(), forces.clone(), e_only))
    for a, b in zip(out[0], out[1]):
        assert (a - b).abs().max().item() <= 1e-11 * a.abs().max().item()
    # single vectors use the specialised kernel whatever the ensemble size (partial last tile included)
    small = DeviceEnsemble(spec, X[:229], F[:229], np.zeros(229))
    one = small.evaluate(spec.prepare(torch.from_numpy(rows[1:2].copy()).cuda()), with_forces=True, want_fres=True)
    assert (one[0][0] - out[0][0][1, :229]).abs().max().item() <= 1e-12 * out[0][0].abs().max().item()
    assert (one[1][0] - out[0][1][1, :229]).abs().max().item() <= 1e-12 * out[0][1].abs().max().item()
    # the metric can change after the kernel was installed
    spec.set_force_metric(2.0 * metric)
    ens = DeviceEnsemble(spec, X, F, np.zeros(n_conf))
    fres2 = ens.evaluate(spec.prepare(torch.from_numpy(rows).cuda()), with_forces=True, want_fres=True)[1]
    assert (fres2 - 2.0 * out[1][1]).abs().max().item() <= 1e-12 * out[1][1].abs().max().item()


def test_empty_batches_are_refused_with_a_message(golden_dir):
    """The reference never evaluates an empty ensemble (``system.py:324-354`` loops over ``ref_coordinates``); the C
    ABI refuses zero-sized batches instead of launching an empty grid, and says why through ``pm_last_error``."""
    import torch
    from paramol_b200 import _lib
    from paramol_b200.MM_engines.device import DeviceTopology
    s, x0 = _system(golden_dir, "aniline")
    topo = DeviceTopology(s)
    L = _lib.lib()
    slots = torch.from_numpy(topo.slots_from_system(s)[None, :]).cuda()
    derived = topo.prepare(slots)
    buf = torch.zeros(4096, dtype=torch.float64, device="cuda")
    h, p = topo._h, buf.data_ptr()
    for what, rc in (("pm_eval", L.pm_eval(h, derived.data_ptr(), 1, p, p, 0, 1, p, p, None, None)),
                     ("pm_eval", L.pm_eval(h, derived.data_ptr(), 0, p, p, 8, 1, p, p, None, None)),
                     ("pm_pack_tiles", L.pm_pack_tiles(h, p, 0, p, None)),
                     ("pm_reduce_objective", L.pm_reduce_objective(p, p, p, None, 0.0, 1, 0, p, p, None))):
        assert rc != 0, what
        with pytest.raises(_lib.B200Error, match="empty|no conformations"):
            _lib.check(rc, what)
    torch.cuda.synchronize()


@pytest.mark.parametrize("name,n_conf,n_vec,weighted", [("aniline", 40000, 1, True), ("aniline", 1003, 3, False), ("aniline", 10, 1, True),
                                                        ("lig50", 5000, 2, True), ("caffeine", 777, 1, False)])
def test_fused_eval_reduce_equals_eval_then_reduce(golden_dir, name, n_conf, n_vec, weighted):
    """``pm_eval_reduce`` (the E(+F) kernel with the residual reduction in its epilogue + the row-sum kernel) against
    ``pm_eval`` followed by ``pm_reduce_objective`` on the same inputs: the eight sums per vector to 1e-12 relative (the two paths
    add the same terms in a different order), with and without weights, generated and interpreting kernels, E+F and energy-only;
    the optional per-conformation outputs of the fused call equal those of ``pm_eval``."""
    import torch
    from paramol_b200.MM_engines.device import DeviceTopology, DeviceEnsemble
    s, x0 = _system(golden_dir, name)
    rng = np.random.default_rng(99)
    X = x0[None] + rng.normal(0, 0.005, (n_conf,) + x0.shape)
    Fref = rng.normal(0, 80.0, X.shape)
    Eref = rng.normal(-100.0, 5.0, n_conf)
    topo = DeviceTopology(s)
    base = topo.slots_from_system(s)
    slots = base[None] * (1.0 + 0.02 * rng.uniform(-1, 1, (n_vec, len(base))))
    ens = DeviceEnsemble(topo, X, Fref, Eref)
    if weighted:
        w = rng.uniform(0.1, 1.0, n_conf)
        ens.set_weights(w / w.sum())
    derived = topo.prepare(torch.from_numpy(slots).cuda())
    for with_forces in (True, False):
        emm, fres, _ = ens.evaluate(derived, with_forces=with_forces, want_fres=with_forces)
        want = ens.reduce(emm, fres if with_forces else None, d_shift=-100.0).cpu().numpy().copy()
        emm_h = emm.cpu().numpy().copy()
        fres_h = fres.cpu().numpy().copy() if with_forces else None
        got, emm2, fres2 = ens.evaluate_reduce(derived, with_forces=with_forces, d_shift=-100.0, want_emm=True, want_fres=with_forces)
        got = got.cpu().numpy()
        scale = np.maximum(np.abs(want), 1e-300)
        if not with_forces:
            want[:, 4] = got[:, 4] = 0.0
        assert np.all(np.abs(got - want) <= 1e-12 * np.maximum(scale, np.abs(want).max(axis=1, keepdims=True) * 1e-4)), (got, want)
        # (batches over small ensembles: pm_eval interprets the topology, the fused call runs the generated kernel -> 1e-13, not bits)
        assert np.abs(emm2.cpu().numpy() - emm_h).max() <= 1e-13 * np.abs(emm_h).max()
        if with_forces:
            assert np.abs(fres2.cpu().numpy() - fres_h).max() <= 1e-13 * np.abs(fres_h).max()
        bare = ens.evaluate_reduce(derived, with_forces=with_forces, d_shift=-100.0)[0].cpu().numpy()
        if not with_forces:
            bare[:, 4] = 0.0
        assert np.array_equal(bare, got)   # deterministic: same launch geometry, same order


@pytest.mark.parametrize("name,n_conf,n_vec,weighted", [("aniline", 1000, 5, True), ("lig60", 4100, 130, False), ("caffeine", 333, 16, True),
                                                        ("lig33", 129, 257, False), ("ethane", 64, 3, True)])
def test_energy_batch_contraction_vs_energy_kernel_and_oracle(golden_dir, name, n_conf, n_vec, weighted):
    """``pm_energy_batch_reduce`` (geometry features once per conformation, E = C G on DMMA) against the energy-only E kernel for
    every vector of the batch (1e-11 of the largest energy: the expansion of the harmonic terms cancels three digits) and, for a
    few vectors, against the oracle evaluated with that vector's parameters; the energy sums against ``pm_eval_reduce``.  Batch
    sizes on both sides of the block of 128 vectors, ragged conformation tiles."""
    import torch
    from oracle import oracle as O
    from paramol_b200.MM_engines.device import DeviceTopology, DeviceEnsemble
    s, x0 = _system(golden_dir, name)
    rng = np.random.default_rng(5)
    X = x0[None] + rng.normal(0, 0.005, (n_conf,) + x0.shape)
    Eref = rng.normal(50.0, 5.0, n_conf)
    topo = DeviceTopology(s)
    base = topo.slots_from_system(s)
    slots = base[None] * (1.0 + 0.02 * rng.uniform(-1, 1, (n_vec, len(base))))
    slots[0] = base
    ens = DeviceEnsemble(topo, X, None, Eref)
    if weighted:
        w = rng.uniform(0.1, 1.0, n_conf)
        ens.set_weights(w / w.sum())
    assert ens.energy_batch_supported()
    derived = topo.prepare(torch.from_numpy(slots).cuda())
    want_p, emm_want, _ = ens.evaluate_reduce(derived, with_forces=False, d_shift=50.0, want_emm=True)
    want_p, emm_want = want_p.cpu().numpy().copy(), emm_want.cpu().numpy().copy()
    got_p, emm = ens.energy_batch_reduce(derived, d_shift=50.0, want_emm=True)
    got_p, emm = got_p.cpu().numpy(), emm.cpu().numpy()
    scale = np.abs(emm_want).max()
    assert np.abs(emm - emm_want).max() <= 1e-11 * max(scale, 1.0), np.abs(emm - emm_want).max() / scale
    want_p[:, 4] = 0.0
    tol = 1e-10 * np.maximum(np.abs(want_p), np.abs(want_p).max(axis=1, keepdims=True) * 1e-3) + 1e-9
    assert np.all(np.abs(got_p - want_p) <= tol), (got_p[0], want_p[0])
    bare = ens.energy_batch_reduce(derived, d_shift=50.0)[0].cpu().numpy()
    assert np.array_equal(bare, got_p)
    for p in (0, n_vec - 1):
        E, _ = O.energies_forces(_system_with_slots(s, slots[p]), X, n_threads=4)
        assert _close_ef(emm[p], E), np.abs(emm[p] - E).max()


@pytest.mark.parametrize("name,n_conf", [("lig50", 203), ("lig60", 131), ("norfloxacin", 96), ("lig33", 77), ("caffeine", 65), ("aniline", 40)])
def test_warp_team_mode_parity_vs_oracle(golden_dir, name, n_conf, monkeypatch):
    """Teams of 8 WARPS on tiles of 32 conformations (``conf_per_tile = -8``: lane = conformation, the members of a team are warps
    that share the tile, a named barrier per step): energies, forces, the energy-only kernel and the residual epilogue against the
    oracle, plus the fused reduction against ``pm_eval`` + ``pm_reduce_objective``; ragged last tile."""
    import torch
    from oracle import oracle as O
    monkeypatch.setenv("PARAMOL_B200_SPECIALIZE", "0")
    s, x0 = _system(golden_dir, name)
    rng = np.random.default_rng(4321)
    X = x0[None] + rng.normal(0, 0.005, (n_conf,) + x0.shape)
    E, F = O.energies_forces(s, X, n_threads=4)
    Fref = F + rng.normal(0, 50.0, F.shape)
    metric = np.tile(np.eye(3).ravel(), (s.n_atoms, 1)) * rng.uniform(0.5, 2.0, (s.n_atoms, 1))
    r = _run(s, X, Fref, E + rng.normal(0, 5, E.shape), metric, conf_per_tile=-8)
    assert r["topo"].program_info()["lanes_per_conformation"] == 8 and r["topo"].conf_per_tile == 32
    assert np.abs(r["emm"][0] - E).max() < 1e-9 * max(1.0, np.abs(E).max())
    assert np.abs(r["fmm"][0] - F).max() < 1e-9 * np.abs(F).max()
    assert np.abs(r["emm_energy_only_kernel"][0] - E).max() < 1e-9 * max(1.0, np.abs(E).max())
    d = F - Fref
    res = np.einsum("sna,nab,snb->s", d, metric.reshape(-1, 3, 3), d)
    assert np.abs(r["fres"][0] - res).max() <= OBJ_RTOL * np.abs(res).max()
    ens, derived = r["ens"], r["derived"]
    emm, fres, _ = ens.evaluate(derived, with_forces=True, want_fres=True)
    want = ens.reduce(emm, fres, d_shift=1.5).cpu().numpy().copy()
    got = ens.evaluate_reduce(derived, with_forces=True, d_shift=1.5)[0].cpu().numpy()
    assert np.all(np.abs(got - want) <= 1e-12 * np.maximum(np.abs(want), np.abs(want).max() * 1e-4)), (got, want)


@pytest.mark.parametrize("name,n_conf,cpw", [("caffeine", 257, 16), ("caffeine", 130, 8), ("aniline", 99, 16), ("lig26", 70, 8), ("lig26", 41, 4)])
def test_lane_team_mode_parity_vs_oracle(golden_dir, name, n_conf, cpw, monkeypatch):
    """Teams of 32 / conf_per_tile LANES of one warp per conformation (the interpreter's shapes for molecules that have no generated
    kernel), forced here on small molecules with the specialisation off: E, F, the energy-only kernel and the residual epilogue
    against the oracle."""
    from oracle import oracle as O
    monkeypatch.setenv("PARAMOL_B200_SPECIALIZE", "0")
    s, x0 = _system(golden_dir, name)
    rng = np.random.default_rng(97)
    X = x0[None] + rng.normal(0, 0.005, (n_conf,) + x0.shape)
    E, F = O.energies_forces(s, X, n_threads=4)
    Fref = F + rng.normal(0, 50.0, F.shape)
    metric = np.tile(np.eye(3).ravel(), (s.n_atoms, 1)) * rng.uniform(0.5, 2.0, (s.n_atoms, 1))
    r = _run(s, X, Fref, E + rng.normal(0, 5, E.shape), metric, conf_per_tile=cpw)
    assert r["topo"].conf_per_tile == cpw and not r["topo"].specialized
    assert np.abs(r["emm"][0] - E).max() < 1e-9 * max(1.0, np.abs(E).max())
    assert np.abs(r["fmm"][0] - F).max() < 1e-9 * np.abs(F).max()
    assert np.abs(r["emm_energy_only_kernel"][0] - E).max() < 1e-9 * max(1.0, np.abs(E).max())
    d = F - Fref
    res = np.einsum("sna,nab,snb->s", d, metric.reshape(-1, 3, 3), d)
    assert np.abs(r["fres"][0] - res).max() <= OBJ_RTOL * np.abs(res).max()


@pytest.mark.parametrize("name", ["caffeine", "lig26", "lig22", "lig12"])
def test_generated_kernel_parity_vs_oracle_at_launch_size(golden_dir, name):
    """The GENERATED E+F and energy-only kernels (installed automatically up to 26 atoms; they serve batches of >= 32768
    conformations) against the oracle on 32768 + 45 conformations: energies, forces, residuals; ragged last tile."""
    from oracle import oracle as O
    s, x0 = _system(golden_dir, name)
    rng = np.random.default_rng(7)
    n_conf = 32768 + 45
    X = x0[None] + rng.normal(0, 0.005, (n_conf,) + x0.shape)
    E, F = O.energies_forces(s, X, n_threads=8)
    Fref = F + rng.normal(0, 50.0, F.shape)
    metric = np.tile(np.eye(3).ravel(), (s.n_atoms, 1)) * rng.uniform(0.5, 2.0, (s.n_atoms, 1))
    metric[:, 1] += 0.1
    r = _run(s, X, Fref, E + rng.normal(0, 5, E.shape), metric)
    assert r["topo"].specialized and r["topo"].conf_per_tile == 32, r["topo"].specialize_status
    assert np.abs(r["emm"][0] - E).max() < 1e-9 * max(1.0, np.abs(E).max())
    assert np.abs(r["fmm"][0] - F).max() < 1e-9 * np.abs(F).max()
    assert np.abs(r["emm_energy_only_kernel"][0] - E).max() < 1e-9 * max(1.0, np.abs(E).max())
    d = F - Fref
    res = np.einsum("sna,nab,snb->s", d, metric.reshape(-1, 3, 3), d)
    assert np.abs(r["fres"][0] - res).max() <= OBJ_RTOL * np.abs(res).max()


def test_generated_kernel_fused_sums_variant(monkeypatch):
    """The measured-and-not-shipped variant of the generated kernels (``PM_SPEC_FUSED=1``: the kernel accumulates the five residual
    sums itself and writes one row of five per warp; no per-conformation energy / residual, no ``pm_reduce_stage1``): its
    ``pm_eval_reduce`` equals ``pm_eval`` + ``pm_reduce_objective`` of the same image, with and without forces, weighted, ragged."""
    import torch
    from paramol_b200.MM_engines.device import DeviceEnsemble, DeviceTopology
    from paramol_b200.Utils.synthetic import make_ligand
    mm, x0 = make_ligand(12)
    monkeypatch.setenv("PM_SPEC_FUSED", "1")
    topo = DeviceTopology(mm)
    assert topo.specialized
    rng = np.random.default_rng(5)
    n_conf = 32768 + 77
    X = x0[None] + rng.normal(0, 0.01, (n_conf,) + x0.shape)
    F = rng.normal(0, 50.0, X.shape)
    topo.set_force_metric(np.tile(np.eye(3).ravel(), (12, 1)) * rng.uniform(0.5, 2.0, (12, 1)))
    ens = DeviceEnsemble(topo, X, F, rng.normal(0, 5.0, n_conf))
    ens.set_weights(rng.uniform(0.5, 1.5, n_conf))
    base = topo.slots_from_system(mm)
    rows = base[None] * (1.0 + 0.02 * rng.uniform(-1, 1, (3, len(base))))
    der = topo.prepare(torch.from_numpy(rows).cuda())
    for wf in (True, False):
        emm, fres, _ = ens.evaluate(der, with_forces=wf, want_fres=wf)
        want = ens.reduce(emm, fres, d_shift=0.7).cpu().numpy().copy()
        got = ens.evaluate_reduce(der, with_forces=wf, d_shift=0.7)[0].cpu().numpy()
        assert np.all(np.abs(got - want) <= 1e-11 * np.maximum(np.abs(want), np.abs(want).max() * 1e-4)), (wf, got, want)
