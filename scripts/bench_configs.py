#!/usr/bin/env python
"""
Measurements for BASELINE.json configs 3-5 (the headline config 2 is bench.py): prints one JSON line per config.
  config 3  lig60, 256 parameter vectors x 100k conformations per launch (Monte Carlo / SA / FD-gradient batches)
  config 4  LLS on lig60 bonded terms over 200k conformations (design matrix + QR reduction)
  config 5  RESP: 500 grid points x 50k conformations, N = 60 (normal-equation assembly + ESP objective)
"""
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
from paramol_b200.MM_engines.b200_engine import B200Engine  # noqa: E402
from paramol_b200.MM_engines.device import DeviceEnsemble  # noqa: E402
from paramol_b200.MM_engines.resp import RESP  # noqa: E402
from paramol_b200.MM_engines.least_square import LinearLeastSquare  # noqa: E402
from paramol_b200.System.system import ParaMolSystem  # noqa: E402
from paramol_b200.Parameter_space.parameter_space import ParameterSpace  # noqa: E402
from paramol_b200.Utils.synthetic import make_ligand  # noqa: E402


def ev():
    return torch.cuda.Event(enable_timing=True)


def config1(n_calls=300):
    """BASELINE config 1: the repository's own small case -- aniline, the 10 stored conformations, ENERGY + FORCE + L2 -- where
    one objective evaluation is pure latency; reported as microseconds per `ObjectiveFunction.f` call with host vectors."""
    from paramol_b200.Objective_function.objective_function import ObjectiveFunction
    from paramol_b200.Objective_function.Properties.energy_property import EnergyProperty
    from paramol_b200.Objective_function.Properties.force_property import ForceProperty
    from paramol_b200.Objective_function.Properties.regularization import Regularization
    golden = os.path.join(REPO, "tests", "golden")
    engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(golden, "aniline.prmtop"),
                        crd_file=os.path.join(golden, "aniline.inpcrd"))
    system = ParaMolSystem("aniline", engine, 14)
    system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True, opt_charges=True, opt_lj=True, opt_sc=True)
    system.read_data(os.path.join(golden, "aniline_10_struct.nc"))
    ps = ParameterSpace()
    _, x0 = ps.get_optimizable_parameters([system])
    x0 = np.asarray(x0, dtype=np.float64)
    ps.calculate_prior_widths("arithmetic")
    out = {}
    for term in ("components", "norm"):
        e_prop, f_prop = EnergyProperty([system]), ForceProperty([system], term_type=term)
        e_prop.calculate_variance()
        f_prop.calculate_variance()
        of = ObjectiveFunction(None, ps, [e_prop, f_prop, Regularization(x0.copy(), ps.prior_widths, "L2")], [system], checkpoint_freq=10 ** 9)
        rng = np.random.default_rng(0)
        trial = [x0 * (1.0 + 0.01 * rng.uniform(-1, 1, len(x0))) for _ in range(n_calls)]
        for x in trial[:20]:
            of.f(x)
        t0 = time.perf_counter()
        vals = [of.f(x) for x in trial]
        dt = (time.perf_counter() - t0) / n_calls
        t0 = time.perf_counter()
        of.f_batch(np.stack(trial))
        db = (time.perf_counter() - t0) / n_calls
        vg = of.f_and_grad(trial[0])
        t0 = time.perf_counter()
        for x in trial[:100]:
            of.f_and_grad(x)
        dg = (time.perf_counter() - t0) / 100
        out[term] = dict(us_per_f=dt * 1e6, us_per_vector_in_f_batch=db * 1e6, us_per_f_and_grad=dg * 1e6, f_first=vals[0])
    return dict(config=1, workload="aniline, the 10 stored conformations, ENERGY + FORCE + L2, 232 parameters, host vectors in / value out",
                **{k + "_" + kk: vv for k, v in out.items() for kk, vv in v.items()})


def config3(P=256, S=100_000):
    mm, x0 = make_ligand(60)
    engine = B200Engine(system=mm)
    topo = engine.device_topology()
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.from_numpy(x0).cuda()[None] + 0.005 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    F = 50.0 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    E = torch.randn(S, dtype=torch.float64, device="cuda", generator=g)
    ens = DeviceEnsemble(topo, X, F, E)
    base = engine.current_slots()
    rng = np.random.default_rng(0)
    slots = torch.from_numpy(base[None, :] * (1.0 + 0.02 * rng.uniform(-1, 1, (P, len(base))))).cuda()
    derived = topo.prepare(slots)
    emm, fres, _ = ens.evaluate(derived)
    ens.reduce(emm, fres)
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    derived = topo.prepare(slots)
    emm, fres, _ = ens.evaluate(derived)
    part = ens.reduce(emm, fres)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    flops = 36 * topo.n_pairs + 30 * topo.n_bonds + 90 * topo.n_angles + 160 * topo.n_torsions + 21 * 60 + 4
    return dict(config=3, workload="lig60 x %d vectors x %d conformations, E+F+residual" % (P, S), ms=ms,
                conf_evals_per_s=P * S / (ms * 1e-3), objective_evals_per_s=P / (ms * 1e-3),
                tflops_algorithmic=flops * P * S / (ms * 1e-3) / 1e12, launch=topo.launch_info(True), program=topo.program_info())


def config4(S=200_000):
    mm, x0 = make_ligand(60)
    engine = B200Engine(system=mm)
    system = ParaMolSystem("lig60", engine, 60)
    system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True)
    rng = np.random.default_rng(2)
    system.ref_coordinates = x0[None] + rng.normal(0, 0.005, (S,) + x0.shape)
    system.n_structures = S
    system.ref_energies = system.get_energies_ensemble() + rng.normal(0, 3.0, S)
    ps = ParameterSpace()
    ps.get_optimizable_parameters([system], symmetry_constrained=False)
    lls = LinearLeastSquare(ps, True, method="L2", scaling_factor=1.0, hyperbolic_beta=0.01, weighting_method="uniform",
                            weighting_temperature=300.0)
    lls._calculate_a_device(system, 0.05, 0.05)  # warm-up (allocations, cusolver handles)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    A = lls._calculate_a_device(system, 0.05, 0.05)
    torch.cuda.synchronize()
    t_design = time.perf_counter() - t0
    M = A.shape[1]
    t0 = time.perf_counter()
    lls.fit_parameters_lls([system])
    torch.cuda.synchronize()
    t_fit_first = time.perf_counter() - t0          # includes the one-time cuSOLVER / cuBLAS handle and workspace set-up
    t0 = time.perf_counter()
    lls.fit_parameters_lls([system])
    torch.cuda.synchronize()
    fit_reduction = lls.reduction
    t_fit = time.perf_counter() - t0
    # the tall reduction alone, both ways, on the matrix the fit just used
    red = {}
    B = torch.randn(S, dtype=torch.float64, device="cuda")
    for name, defect in (("cholesky_qr2", lls.CHOLESKY_QR_MAX_DEFECT), ("householder", 0.0)):
        lls.CHOLESKY_QR_MAX_DEFECT = defect
        lls._reduce_tall(A, B)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lls._reduce_tall(A, B)
        torch.cuda.synchronize()
        red[name + "_s"] = time.perf_counter() - t0
        red[name + "_ran"] = lls.reduction
    return dict(config=4, reduction_in_fit=fit_reduction, full_fit_first_call_s=t_fit_first, **red, workload="LLS lig60: S=%d conformations x M=%d columns" % (S, M), design_matrix_s=t_design,
                design_GB_s=(S * 60 * 24 + S * M * 8) / t_design / 1e9, full_fit_s=t_fit,
                note="full fit = design matrix + zero-parameter energy pass + weighting + device reduction to (R, Q^T b) (CholeskyQR2 or Householder, library calls) + host lstsq on the "
                     "stacked (M + reg) x M system")


def config5(S=50_000, G=500, N=60):
    mm, x0 = make_ligand(N)
    rng = np.random.default_rng(11)
    X = (x0[None] + rng.normal(0, 0.005, (S,) + x0.shape)) * 18.897
    dirs = rng.normal(size=(S, G, 3))
    dirs /= np.linalg.norm(dirs, axis=2, keepdims=True)
    centres = X[np.arange(S)[:, None], rng.integers(0, N, (S, G))]
    grid = centres + dirs * rng.uniform(1.4, 2.0, (S, G, 1)) * 3.2
    q_true = mm.particle_par[:, 0]
    L = _lib.lib()
    dev = torch.device("cuda")
    xd, gd = torch.from_numpy(X).to(dev), torch.from_numpy(grid).to(dev)
    qd = torch.from_numpy(np.ascontiguousarray(q_true)).to(dev)
    vref = torch.empty((S, G), dtype=torch.float64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    a, b = ev(), ev()
    _lib.check(L.pm_esp_potential(N, xd.data_ptr(), gd.data_ptr(), qd.data_ptr(), S, G, vref.data_ptr(), st))
    torch.cuda.synchronize()
    a.record()
    _lib.check(L.pm_esp_potential(N, xd.data_ptr(), gd.data_ptr(), qd.data_ptr(), S, G, vref.data_ptr(), st))
    b.record()
    torch.cuda.synchronize()
    ms_pot = a.elapsed_time(b)
    vref += 1e-3 * torch.randn_like(vref)
    out = torch.empty(N * N + N + 1, dtype=torch.float64, device=dev)
    scratch = torch.empty(int(L.pm_resp_scratch_doubles(N, S, G)), dtype=torch.float64, device=dev)
    for _ in range(2):
        _lib.check(L.pm_resp_assemble(N, xd.data_ptr(), gd.data_ptr(), vref.data_ptr(), None, S, G, out.data_ptr(), scratch.data_ptr(), st))
    torch.cuda.synchronize()
    a.record()
    _lib.check(L.pm_resp_assemble(N, xd.data_ptr(), gd.data_ptr(), vref.data_ptr(), None, S, G, out.data_ptr(), scratch.data_ptr(), st))
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    flops = 2.0 * S * G * N * N + 2.0 * S * G * N + 12.0 * S * G * N  # SURVEY count: full N x N product, B, inverse distances
    # what the kernel executes: 8x8 DMMA tiles of the upper triangle of the (N+1)-row matrix over grid tiles of 32, + 11 FP64
    # operations (22 flops) per generated inverse distance
    nb8, gpad = (N + 1 + 7) // 8, (G + 31) // 32 * 32
    executed = S * gpad * (2.0 * 64 * nb8 * (nb8 + 1) / 2 + 22.0 * N)
    return dict(config=5, workload="RESP N=%d, G=%d grid points x S=%d conformations" % (N, G, S), assemble_ms=ms,
                assemble_tflops=flops / (ms * 1e-3) / 1e12, executed_tflops=executed / (ms * 1e-3) / 1e12,
                fp64_peak_tflops=37.1, executed_frac_of_fp64_peak=executed / (ms * 1e-3) / 1e12 / 37.1,
                conformations_per_s=S / (ms * 1e-3), esp_potential_ms=ms_pot,
                bytes_per_conformation=24 * N + 32 * G, hbm_GB_s=S * (24 * N + 32 * G) / (ms * 1e-3) / 1e9)


def config6(which="lig60", S=100_000):
    """Analytic gradient: forward pass (E, F, residuals, MM force tiles) + pm_grad, against the evaluations a 2-point
    finite-difference gradient would need (n_slots + 1 forward passes of the same kernel, batched)."""
    if which == "aniline":
        golden = os.path.join(REPO, "tests", "golden")
        engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(golden, "aniline.prmtop"),
                            crd_file=os.path.join(golden, "aniline.inpcrd"))
        x0 = np.asarray(engine.get_positions(), dtype=np.float64).reshape(-1, 3)
    else:
        mm, x0 = make_ligand(60)
        engine = B200Engine(system=mm)
    topo = engine.device_topology()
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.from_numpy(x0).cuda()[None] + 0.005 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    F = 50.0 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    E = torch.randn(S, dtype=torch.float64, device="cuda", generator=g)
    ens = DeviceEnsemble(topo, X, F, E)
    base = engine.current_slots()
    derived = topo.prepare(torch.from_numpy(base[None].copy()).cuda())
    ca = torch.randn(S, dtype=torch.float64, device="cuda", generator=g)
    cb = torch.rand(S, dtype=torch.float64, device="cuda", generator=g)
    times = {}
    for name, fn in (("forward", lambda: ens.evaluate(derived, with_forces=True, want_fres=True)),
                     ("forward_with_force_tiles", lambda: ens.evaluate(derived, with_forces=True, want_fres=True, want_fmm=True))):
        fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(5):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        times[name] = a.elapsed_time(b) / 5
    fmm = out[2]
    ens.term_gradient(derived, ca, cb, fmm[0])
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(5):
        ens.term_gradient(derived, ca, cb, fmm[0])
    b.record()
    torch.cuda.synchronize()
    times["pm_grad"] = a.elapsed_time(b) / 5
    n_terms = topo.n_bonds + topo.n_angles + topo.n_torsions + topo.n_pairs
    return dict(config=6, workload="analytic gradient, %s, S=%d conformations, %d slots, %d terms" % (which, S, topo.n_slots, n_terms),
                forward_ms=times["forward"], forward_with_force_tiles_ms=times["forward_with_force_tiles"], grad_ms=times["pm_grad"],
                gradient_conformations_per_s=S / (times["pm_grad"] * 1e-3),
                analytic_total_ms=times["forward_with_force_tiles"] + times["pm_grad"],
                two_point_fd_ms=times["forward"] * (topo.n_slots + 1),
                speedup_vs_fd=times["forward"] * (topo.n_slots + 1) / (times["forward_with_force_tiles"] + times["pm_grad"]))


def config7(S=1 << 20, batch=64, rounds=6):
    """Monte Carlo-style batches through the public API: `batch` proposals per f_batch call, each the current vector with ONE
    displaced parameter (ParaMol/Optimizers/monte_carlo.py:89-100), aniline, S conformations, ENERGY + FORCE + L2; every
    round "accepts" one proposal so that the next batch is built on a new vector.  Full evaluation vs delta evaluation."""
    from paramol_b200.Objective_function.objective_function import ObjectiveFunction
    from paramol_b200.Objective_function.Properties.energy_property import EnergyProperty
    from paramol_b200.Objective_function.Properties.force_property import ForceProperty
    from paramol_b200.Objective_function.Properties.regularization import Regularization
    from paramol_b200.Utils.netcdf_lite import read_paramol_data
    from paramol_b200.Utils.synthetic import perturbed_conformations
    golden = os.path.join(REPO, "tests", "golden")
    X10, _, _ = read_paramol_data(os.path.join(golden, "aniline_10_struct.nc"))
    X = perturbed_conformations(X10, S, sigma=0.005, seed=1234)
    rng = np.random.default_rng(5)
    E_ref, F_ref = rng.normal(-43000.0, 5.0, S), rng.normal(0, 50.0, (S, 14, 3))
    out = {}
    for delta in (False, True):
        engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(golden, "aniline.prmtop"),
                            crd_file=os.path.join(golden, "aniline.inpcrd"))
        system = ParaMolSystem("aniline", engine, 14, ref_coordinates=X, ref_energies=E_ref, ref_forces=F_ref)
        system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True)
        ps = ParameterSpace()
        _, x0 = ps.get_optimizable_parameters([system])
        x0 = np.asarray(x0, dtype=np.float64)
        ps.calculate_prior_widths("arithmetic")
        e_prop, f_prop = EnergyProperty([system]), ForceProperty([system], term_type="components")
        e_prop.calculate_variance()
        f_prop.calculate_variance()
        of = ObjectiveFunction(None, ps, [e_prop, f_prop, Regularization(x0.copy(), ps.prior_widths, "L2")], [system],
                               checkpoint_freq=10 ** 9, delta_evaluation=delta)
        r2 = np.random.default_rng(9)
        x = x0.copy()
        of.f(x)
        vals = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(rounds):
            trial = np.repeat(x[None], batch, 0)
            which = r2.choice(len(x), batch, replace=False)
            trial[np.arange(batch), which] *= 1.0 + 0.01 * r2.uniform(-1, 1, batch)
            v = of.f_batch(trial)
            vals.append(v)
            x = trial[int(np.argmin(v))].copy()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["delta" if delta else "full"] = dict(seconds=dt, objective_evals_per_s=rounds * batch / dt, values=np.concatenate(vals),
                                                 delta_rows=of.n_delta_rows)
        del of, system, engine
        torch.cuda.empty_cache()
    err = float(np.abs(out["delta"]["values"] - out["full"]["values"]).max() / np.abs(out["full"]["values"]).max())
    return dict(config=7, workload="Monte Carlo batches: aniline, S=%d conformations, %d single-parameter proposals per f_batch, %d batches, "
                "through ObjectiveFunction.f_batch (host vectors in, values out)" % (S, batch, rounds),
                full_objective_evals_per_s=out["full"]["objective_evals_per_s"],
                delta_objective_evals_per_s=out["delta"]["objective_evals_per_s"],
                speedup=out["delta"]["objective_evals_per_s"] / out["full"]["objective_evals_per_s"],
                delta_rows=out["delta"]["delta_rows"], max_rel_difference=err)


if __name__ == "__main__":
    which = [int(a) for a in sys.argv[1:]] or [3, 4, 5]
    if 6 in which:
        which = [c for c in which if c != 6]
        print(json.dumps(config6("aniline", 1 << 20)), flush=True)
        print(json.dumps(config6("lig60", 100_000)), flush=True)
    for c in which:
        print(json.dumps({1: config1, 3: config3, 4: config4, 5: config5, 7: config7}[c]()), flush=True)
