#!/usr/bin/env python
"""
Short, CUDA-event timed measurements of the BASELINE.json configurations other than the headline one; ``bench.py`` runs them in
the same process after the headline measurement (N = 1 only) and puts the results under ``configs`` of its JSON line, so that the
driver-recorded line carries every configuration the north star names:

  lig50    ~50-atom ligand, one parameter vector x 2^18 conformations: the size the north-star target sentence is about
  config1  the repository test-suite case: aniline, the 10 stored conformations, microseconds per ``ObjectiveFunction.f`` call,
           with the CPU port of the reference path beside it
  config3  lig60, 256 parameter vectors x 100 000 conformations per launch (E+F batch and energy-only batch)
  config4  LLS fit of the bonded terms of lig60 over 200 000 conformations
  config5  RESP normal equations, N = 60, 500 grid points x 50 000 conformations

Stand-alone:  python bench_configs.py [lig50 config1 config3 config4 config5 grad delta]
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def _ev():
    import torch
    return torch.cuda.Event(enable_timing=True)


def flops_per_conformation(n, nb, na, nt, npair, term_type="components", with_forces=True):
    """36/pair + 30/bond + 90/angle + 160/torsion + epilogue (9N+4 norm | 21N+4 components)."""
    if not with_forces:
        return 18 * npair + 15 * nb + 45 * na + 80 * nt
    return 36 * npair + 30 * nb + 90 * na + 160 * nt + ((21 * n + 4) if term_type == "components" else (9 * n + 4))


def _timed(fn, reps, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = _ev(), _ev()
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def _ligand_ensemble(n_atoms, S, seed=1):
    import torch
    from paramol_b200.MM_engines.b200_engine import B200Engine
    from paramol_b200.MM_engines.device import DeviceEnsemble
    from paramol_b200.Utils.synthetic import make_ligand
    mm, x0 = make_ligand(n_atoms)
    engine = B200Engine(system=mm)
    topo = engine.device_topology()
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.from_numpy(x0).cuda()[None] + 0.005 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    F = 50.0 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    E = torch.randn(S, dtype=torch.float64, device="cuda", generator=g)
    return mm, engine, topo, DeviceEnsemble(topo, X, F, E)


def lig50(fp64_peak, S=1 << 18, n_atoms=50, reps=10):
    """One parameter vector over S conformations of the ~50-atom ligand: prepare + fused evaluate/reduce, device resident."""
    import torch
    mm, engine, topo, ens = _ligand_ensemble(n_atoms, S)
    slots = torch.from_numpy(engine.current_slots()[None].copy()).cuda()

    def step():
        return ens.evaluate_reduce(topo.prepare(slots), with_forces=True)[0]
    ms, _ = _timed(step, reps)
    fl = flops_per_conformation(n_atoms, topo.n_bonds, topo.n_angles, topo.n_torsions, topo.n_pairs)
    tf = fl * S / (ms * 1e-3) / 1e12
    return dict(workload="lig%d (synthetic drug-like ligand, %d atoms) x %d conformations, 1 parameter vector, E+F+residual+reduction"
                % (n_atoms, n_atoms, S), ms_per_step=ms, conformations_per_s=S / (ms * 1e-3),
                roofline=dict(bound="fp64", kernel="pm_ef_spec_kernel" if topo.specialized else
                              ("pm_ef_kernel<teams of %d warps on tiles of 32 conformations>" if (topo.conf_per_tile == 32 and topo.program_info()["lanes_per_conformation"] > 1)
                               else "pm_ef_kernel<%d lanes per conformation>") % topo.program_info()["lanes_per_conformation"], achieved=tf, peak=fp64_peak, unit="TFLOP/s",
                              frac=tf / fp64_peak, flops_per_conformation=fl, note="fraction of the whole step, launches included"),
                launch=topo.launch_info(True), program=topo.program_info())


def config1(n_calls=200):
    """Aniline, the 10 stored conformations, ENERGY + FORCE(components) + L2: latency of one ``ObjectiveFunction.f(x)`` call with a
    host vector, and the same objective through the CPU port of the reference path (C oracle for E/F + numpy properties)."""
    from oracle import oracle as O
    from paramol_b200.MM_engines.b200_engine import B200Engine
    from paramol_b200.Objective_function.objective_function import ObjectiveFunction
    from paramol_b200.Objective_function.Properties.energy_property import EnergyProperty
    from paramol_b200.Objective_function.Properties.force_property import ForceProperty
    from paramol_b200.Objective_function.Properties.regularization import Regularization
    from paramol_b200.Parameter_space.parameter_space import ParameterSpace
    from paramol_b200.System.system import ParaMolSystem
    engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(GOLDEN, "aniline.prmtop"),
                        crd_file=os.path.join(GOLDEN, "aniline.inpcrd"))
    system = ParaMolSystem("aniline", engine, 14)
    system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True, opt_charges=True, opt_lj=True, opt_sc=True)
    system.read_data(os.path.join(GOLDEN, "aniline_10_struct.nc"))
    ps = ParameterSpace()
    _, x0 = ps.get_optimizable_parameters([system])
    x0 = np.asarray(x0, dtype=np.float64)
    ps.calculate_prior_widths("arithmetic")
    e_prop, f_prop = EnergyProperty([system]), ForceProperty([system], term_type="components")
    e_prop.calculate_variance()
    f_prop.calculate_variance()
    of = ObjectiveFunction(None, ps, [e_prop, f_prop, Regularization(x0.copy(), ps.prior_widths, "L2")], [system], checkpoint_freq=10 ** 9)
    rng = np.random.default_rng(0)
    trial = [x0 * (1.0 + 0.01 * rng.uniform(-1, 1, len(x0))) for _ in range(n_calls)]
    for x in trial[:20]:
        of.f(x)
    t0 = time.perf_counter()
    vals = [of.f(x) for x in trial]
    us_f = (time.perf_counter() - t0) / n_calls * 1e6
    of.f_batch(np.stack(trial))   # warm: the batch buffers are allocated on the first call
    t0 = time.perf_counter()
    of.f_batch(np.stack(trial))
    us_batch = (time.perf_counter() - t0) / n_calls * 1e6
    # the CPU port of the same call: parameters already applied (the host-side update is common to both), E/F of the 10
    # conformations by the C oracle on one thread, the properties in numpy as the reference does
    X, Eq, Fq = (np.asarray(a) for a in (system.ref_coordinates, system.ref_energies, system.ref_forces))
    O.objective(engine.system, X, Eq, Fq, force_term_type="components")
    t0 = time.perf_counter()
    for _ in range(n_calls):
        O.objective(engine.system, X, Eq, Fq, force_term_type="components")
    us_cpu = (time.perf_counter() - t0) / n_calls * 1e6
    t0 = time.perf_counter()
    for x in trial:
        ps.update_systems([system], x)
    us_update = (time.perf_counter() - t0) / n_calls * 1e6
    return dict(workload="aniline, the 10 stored conformations, ENERGY + FORCE(components) + L2, 232 parameters, host vector in / value out",
                us_per_f=us_f, us_per_vector_in_f_batch=us_batch, f_first=vals[0],
                host_update_systems_us=us_update,
                cpu_port=dict(us_per_objective=us_cpu, us_with_the_same_host_update=us_cpu + us_update, cores=1, kind="port",
                              note="oracle.objective: C E/F of the 10 conformations + numpy properties, parameters already applied"))


def config3(fp64_peak, P=256, S=100_000, n_atoms=60):
    import torch
    mm, engine, topo, ens = _ligand_ensemble(n_atoms, S)
    base = engine.current_slots()
    rng = np.random.default_rng(0)
    slots = torch.from_numpy(base[None, :] * (1.0 + 0.02 * rng.uniform(-1, 1, (P, len(base))))).cuda()

    def step():
        return ens.evaluate_reduce(topo.prepare(slots), with_forces=True)[0]
    ms, _ = _timed(step, 2, warm=1)

    def step_e():
        return ens.evaluate_reduce(topo.prepare(slots), with_forces=False)[0]
    ms_e, _ = _timed(step_e, 2, warm=1)

    def step_g():    # energy-only batch with geometry reuse: features once per conformation, E = C G on DMMA
        return ens.energy_batch_reduce(topo.prepare(slots))[0]
    ms_g, _ = _timed(step_g, 3, warm=1)
    n_feat = 2 * topo.n_bonds + 2 * topo.n_angles + 2 * topo.n_torsions + 3 * topo.n_pairs + 1
    tf_g = 2.0 * P * S * n_feat / (ms_g * 1e-3) / 1e12
    fl = flops_per_conformation(n_atoms, topo.n_bonds, topo.n_angles, topo.n_torsions, topo.n_pairs)
    tf = fl * P * S / (ms * 1e-3) / 1e12
    return dict(workload="lig%d x %d parameter vectors x %d conformations per launch" % (n_atoms, P, S), ms=ms,
                conf_evals_per_s=P * S / (ms * 1e-3), objective_evals_per_s=P / (ms * 1e-3),
                energy_only_ms=ms_e, energy_only_conf_evals_per_s=P * S / (ms_e * 1e-3),
                energy_batch_contraction=dict(ms=ms_g, conf_evals_per_s=P * S / (ms_g * 1e-3), features=n_feat,
                                              kernel="pm_egemm_burst_kernel (DMMA.8x8x4): geometry features generated once per conformation in bursts, coefficients streamed by TMA",
                                              roofline=dict(bound="fp64", achieved=tf_g, peak=fp64_peak, unit="TFLOP/s", frac=tf_g / fp64_peak,
                                                            note="2 P S T flop of the contraction / time; the term-pass kernel above does the same job in energy_only_ms")),
                roofline=dict(bound="fp64", achieved=tf, peak=fp64_peak, unit="TFLOP/s", frac=tf / fp64_peak, flops_per_conformation=fl),
                launch=topo.launch_info(True), program=topo.program_info())


def config4(S=200_000, n_atoms=60):
    import torch
    from paramol_b200.MM_engines.b200_engine import B200Engine
    from paramol_b200.MM_engines.least_square import LinearLeastSquare
    from paramol_b200.Parameter_space.parameter_space import ParameterSpace
    from paramol_b200.System.system import ParaMolSystem
    from paramol_b200.Utils.synthetic import make_ligand
    mm, x0 = make_ligand(n_atoms)
    engine = B200Engine(system=mm)
    system = ParaMolSystem("lig%d" % n_atoms, engine, n_atoms)
    system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True)
    rng = np.random.default_rng(2)
    system.ref_coordinates = x0[None] + rng.normal(0, 0.005, (S,) + x0.shape)
    system.n_structures = S
    system.ref_energies = system.get_energies_ensemble() + rng.normal(0, 3.0, S)
    ps = ParameterSpace()
    ps.get_optimizable_parameters([system], symmetry_constrained=False)
    lls = LinearLeastSquare(ps, True, method="L2", scaling_factor=1.0, hyperbolic_beta=0.01, weighting_method="uniform",
                            weighting_temperature=300.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lls.fit_parameters_lls([system])
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    x_fit = np.array(lls._parameters, copy=True)
    ps.update_systems([system], np.asarray(ps.optimizable_parameters_values, dtype=np.float64))
    t0 = time.perf_counter()
    lls.fit_parameters_lls([system])
    torch.cuda.synchronize()
    t_fit = time.perf_counter() - t0
    return dict(workload="LLS lig%d: S=%d conformations x M=%d columns" % (n_atoms, S, len(x_fit)), full_fit_s=t_fit,
                full_fit_first_call_s=t_first, reduction=lls.reduction, timings=dict(getattr(lls, "timings", {})),
                note="full fit = term table, column means (pm_lls_colstats), zero-parameter energy pass, normal equations (pm_lls_slab + pm_lls_syrk on DMMA), Cholesky solve + iterative refinement (pm_lls_residual, pm_lls_colstats)")


def config5(fp64_peak, S=50_000, G=500, N=60):
    import torch
    from paramol_b200 import _lib
    from paramol_b200.Utils.synthetic import make_ligand
    mm, x0 = make_ligand(N)
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(11)
    xd = (torch.from_numpy(x0).to(dev)[None] + 0.005 * torch.randn((S, N, 3), dtype=torch.float64, device=dev, generator=g)) * 18.897
    dirs = torch.randn((S, G, 3), dtype=torch.float64, device=dev, generator=g)
    dirs /= dirs.norm(dim=2, keepdim=True)
    centres = xd[torch.arange(S, device=dev)[:, None], torch.randint(0, N, (S, G), device=dev, generator=g)]
    gd = (centres + dirs * (1.4 + 0.6 * torch.rand((S, G, 1), dtype=torch.float64, device=dev, generator=g)) * 3.2).contiguous()
    del dirs, centres
    qd = torch.from_numpy(np.ascontiguousarray(mm.particle_par[:, 0])).to(dev)
    L = _lib.lib()
    vref = torch.empty((S, G), dtype=torch.float64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ms_pot, _ = _timed(lambda: _lib.check(L.pm_esp_potential(N, xd.data_ptr(), gd.data_ptr(), qd.data_ptr(), S, G, vref.data_ptr(), st)), 2, warm=1)
    vref += 1e-3 * torch.randn(vref.shape, dtype=torch.float64, device=dev, generator=g)
    out = torch.empty(N * N + N + 1, dtype=torch.float64, device=dev)
    scratch = torch.empty(int(L.pm_resp_scratch_doubles(N, S, G)), dtype=torch.float64, device=dev)
    ms, _ = _timed(lambda: _lib.check(L.pm_resp_assemble(N, xd.data_ptr(), gd.data_ptr(), vref.data_ptr(), None, S, G, out.data_ptr(),
                                                         scratch.data_ptr(), st)), 3, warm=2)
    flops = 2.0 * S * G * N * N + 2.0 * S * G * N + 12.0 * S * G * N  # SURVEY count: full N x N product, B, inverse distances
    nb8, gpad = (N + 1 + 7) // 8, (G + 31) // 32 * 32
    executed = S * gpad * (2.0 * 64 * nb8 * (nb8 + 1) / 2 + 22.0 * N)
    return dict(workload="RESP N=%d, G=%d grid points x S=%d conformations" % (N, G, S), assemble_ms=ms,
                conformations_per_s=S / (ms * 1e-3), esp_potential_ms=ms_pot,
                roofline=dict(bound="fp64", kernel="pm_resp_kernel (DMMA.8x8x4)", achieved=flops / (ms * 1e-3) / 1e12, peak=fp64_peak,
                              unit="TFLOP/s", frac=flops / (ms * 1e-3) / 1e12 / fp64_peak,
                              executed=executed / (ms * 1e-3) / 1e12, executed_frac=executed / (ms * 1e-3) / 1e12 / fp64_peak,
                              note="achieved = SURVEY flop count (full N x N product); executed = upper-triangle DMMA tiles + generation"),
                bytes_per_conformation=24 * N + 32 * G, hbm_GB_s=S * (24 * N + 32 * G) / (ms * 1e-3) / 1e9)


def grad(which="lig60", S=100_000):
    """Analytic gradient: forward pass (E, F, residuals, MM force tiles) + pm_grad, against the evaluations a 2-point
    finite-difference gradient would need (n_slots + 1 forward passes of the same kernel, batched)."""
    import torch
    from paramol_b200.MM_engines.b200_engine import B200Engine
    from paramol_b200.MM_engines.device import DeviceEnsemble
    if which == "aniline":
        engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(GOLDEN, "aniline.prmtop"),
                            crd_file=os.path.join(GOLDEN, "aniline.inpcrd"))
        x0 = np.asarray(engine.get_positions(), dtype=np.float64).reshape(-1, 3)
        topo = engine.device_topology()
        g = torch.Generator(device="cuda").manual_seed(1)
        X = torch.from_numpy(x0).cuda()[None] + 0.005 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
        F = 50.0 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
        ens = DeviceEnsemble(topo, X, F, torch.randn(S, dtype=torch.float64, device="cuda", generator=g))
    else:
        _, engine, topo, ens = _ligand_ensemble(60, S)
    g = torch.Generator(device="cuda").manual_seed(3)
    derived = topo.prepare(torch.from_numpy(engine.current_slots()[None].copy()).cuda())
    ca = torch.randn(S, dtype=torch.float64, device="cuda", generator=g)
    cb = torch.rand(S, dtype=torch.float64, device="cuda", generator=g)
    t_fwd, _ = _timed(lambda: ens.evaluate(derived, with_forces=True, want_fres=True), 5)
    t_fwd_f, out = _timed(lambda: ens.evaluate(derived, with_forces=True, want_fres=True, want_fmm=True), 5)
    fmm = out[2]
    t_grad, _ = _timed(lambda: ens.term_gradient(derived, ca, cb, fmm[0]), 5)
    return dict(workload="analytic gradient, %s, S=%d conformations, %d slots" % (which, S, topo.n_slots), forward_ms=t_fwd,
                forward_with_force_tiles_ms=t_fwd_f, grad_ms=t_grad, analytic_total_ms=t_fwd_f + t_grad,
                two_point_fd_ms=t_fwd * (topo.n_slots + 1), speedup_vs_fd=t_fwd * (topo.n_slots + 1) / (t_fwd_f + t_grad))


def run_all(fp64_peak, which=("lig50", "config1", "config3", "config4", "config5")):
    import torch
    table = {"lig50": lambda: lig50(fp64_peak), "config1": config1, "config3": lambda: config3(fp64_peak), "config4": config4,
             "config5": lambda: config5(fp64_peak), "grad": grad}
    out = {}
    for name in which:
        t0 = time.perf_counter()
        try:
            out[name] = table[name]()
        except Exception as exc:  # a failing side configuration must not lose the headline line; it is reported, not hidden
            out[name] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
        out[name]["wall_s"] = time.perf_counter() - t0
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import torch
    from paramol_b200 import _lib
    fl, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(_lib.lib().pm_measure_fma_peak(0, 1, 4000, ctypes.byref(fl), ctypes.byref(ms)))
    names = sys.argv[1:] or ["lig50", "config1", "config3", "config4", "config5"]
    for k, v in run_all(fl.value / 1e12, names).items():
        print(json.dumps({k: v}), flush=True)
