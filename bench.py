#!/usr/bin/env python
"""
bench.py -- headline benchmark of the B200-native ParaMol objective-evaluation path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload aniline|lig50|lig60|norfloxacin]
                    [--conformations S] [--no-configs]

Workload (BASELINE.json configs[1]): the reference test-suite molecule (aniline, 14 atoms), S = 2^20 synthetic perturbed
conformations IN TOTAL, sharded over the N GPUs by conformation (STRONG scaling: the reference splits a fixed ensemble over its
workers, ParaMol/Objective_function/objective_function.py:147-155); objective = ENERGY + FORCE("components") + L2
REGULARIZATION with uniform weights.  A "step" is ONE objective evaluation over all conformations.  At N > 1 the same
measurement is repeated with 2^20 conformations PER GPU and reported under ``weak``.

Printed JSON line (rank 0):
  value      conformation E+F evaluations / s, whole job, parameters and conformations resident in HBM
             (pm_prepare_params + pm_eval_reduce per step: the E+F kernel, the residual sums, then the row-sum kernel that is also
             the all-reduce over NVLink peer memory; CUDA-event timed, max over ranks)
  e2e        the same metric through the public API ``ObjectiveFunction.f(x)`` with a HOST parameter vector: numpy
             plumbing, pinned H2D of the slot row, one CUDA-graph replay, D2H of the 8 global sums, every step
  roofline   dominant kernel (pm_ef_spec_kernel for small molecules, else pm_ef_kernel), timed ALONE with CUDA events in a loop of
             the same length right after the step loop: algorithmic FP64 flops / that time against the DFMA peak measured live on this GPU (the kernel is FP64-pipe bound; the nominal pipe
             peak 148 SM x 64 DFMA x 2 x f_SM is printed beside it); ``frac_of_step`` divides by the whole step instead of the
             kernel; ``hbm`` holds the same kernel against the measured copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the FP64 CPU port of the reference path (oracle/oracle.c, OpenMP over conformations, all host cores) on exactly
             the bench workload (same S)
  configs    (N = 1) the other BASELINE.json configurations and the ~50-atom ligand of the target sentence, measured in this run
  multi_gpu_parity  (N > 1) the N-rank objective against the same ensemble evaluated unsharded on rank 0 alone
``--impl reference`` times that CPU port alone (OpenMM, where ParaMol's arithmetic lives, is not installable here).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")
METRIC = "conformation E+F evals/sec"
BLOCK = 4096  # conformations per data-generation block (global index -> seed), so that any shard of the global set can be made alone


def load_molecule(name):
    from paramol_b200.Utils.amber import system_from_prmtop, read_inpcrd
    from paramol_b200.Utils.synthetic import make_ligand
    if name.startswith("lig"):
        return make_ligand(int(name[3:]))
    return (system_from_prmtop(os.path.join(GOLDEN, name + ".prmtop")), read_inpcrd(os.path.join(GOLDEN, name + ".inpcrd")))


def base_conformations(name, x0):
    if name == "aniline":
        from paramol_b200.Utils.netcdf_lite import read_paramol_data
        return read_paramol_data(os.path.join(GOLDEN, "aniline_10_struct.nc"))[0]
    return x0[None]


def global_slice(x_base, lo, hi, n_atoms):
    """Conformations [lo, hi) of the GLOBAL synthetic set and their label noise: coords_s = X[s mod len(X)] + N(0, 0.005 nm), noise
    N(0, 5 kJ/mol) on energies and N(0, 50 kJ/mol/nm) on forces; block b of 4096 conformations is seeded by b, so every rank makes
    exactly its shard and an unsharded run makes the same numbers."""
    xs, es, fs = [], [], []
    for b in range(lo // BLOCK, (hi + BLOCK - 1) // BLOCK):
        rng = np.random.default_rng([1234, b])
        idx = (np.arange(BLOCK) + b * BLOCK) % x_base.shape[0]
        x = x_base[idx] + rng.normal(0.0, 0.005, (BLOCK, n_atoms, 3))
        e = rng.normal(0.0, 5.0, BLOCK)
        f = rng.normal(0.0, 50.0, (BLOCK, n_atoms, 3))
        a, z = max(lo, b * BLOCK) - b * BLOCK, min(hi, (b + 1) * BLOCK) - b * BLOCK
        xs.append(x[a:z])
        es.append(e[a:z])
        fs.append(f[a:z])
    return np.concatenate(xs), np.concatenate(es), np.concatenate(fs)


class ClockSampler:
    """Streams nvidia-smi clocks / throttle reasons of one GPU (`-lms 20`) while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, universal_newlines=True)
            self._thread = threading.Thread(target=self._read, daemon=True)
            self._thread.start()
            time.sleep(0.15)  # let the first samples arrive before the timed region starts
        except Exception:
            self.proc = None
        return self

    def _read(self):
        try:
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def mark(self):
        """Only samples taken after this call count (start of the timed region)."""
        self.t_start = time.perf_counter()

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        t0 = getattr(self, "t_start", 0.0)
        for t, r in self.rows:
            if t < t0 or len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
class CpuOracleWorkload:
    """
    The CPU arm: the FP64 port of the reference path (oracle/oracle.c: E+F of every conformation, OpenMP over conformations =
    the analogue of the reference's fork pool; per-conformation force residual in C; the scalar reductions of EnergyProperty /
    ForceProperty in numpy) on ``n_conf`` conformations of the bench workload -- the same S as the GPU arm unless that would
    take longer than ``max_seconds_per_step``.
    """

    def __init__(self, mm_system, x_base, term_type, n_conf, max_seconds_per_step=20.0, threads=None):
        from oracle import oracle as O
        self.O = O
        self.mm = mm_system
        # every host core this process may run on (torchrun exports OMP_NUM_THREADS=1; the C port takes the count explicitly)
        self.threads = threads or max(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1), 1)
        probe = global_slice(x_base, 0, BLOCK, mm_system.n_atoms)[0]
        O.energies_forces(mm_system, probe[:64], n_threads=self.threads)  # spin up the OpenMP team
        t0 = time.perf_counter()
        O.energies_forces(mm_system, probe, n_threads=self.threads)
        dt = time.perf_counter() - t0
        self.requested = int(n_conf)
        self.n = int(min(n_conf, max(BLOCK, BLOCK * max_seconds_per_step / max(dt, 1e-6))))
        self.X, e_noise, f_noise = global_slice(x_base, 0, self.n, mm_system.n_atoms)
        E, F = O.energies_forces(mm_system, self.X, n_threads=self.threads)
        self.Eref = E + e_noise
        self.Fref = F + f_noise
        self.w = O.conformation_weights("uniform", self.n)
        if term_type == "norm":
            self.metric = np.eye(3)[None] / O.force_variance(self.Fref)[:, None, None]
        else:
            self.metric = O.force_inverse_covariance(self.Fref[:min(self.n, 2048)])
        self.var_e = np.var(self.Eref)

    def step(self):
        """One objective evaluation over the sample; returns seconds."""
        O = self.O
        t0 = time.perf_counter()
        E, F = O.energies_forces(self.mm, self.X, n_threads=self.threads)
        res = O.force_residuals(F, self.Fref, self.metric, n_threads=self.threads)
        d = self.Eref - E
        value = np.sum(self.w * (d - np.mean(d)) ** 2) / self.var_e + np.sum(self.w * res) / (3 * self.mm.n_atoms)
        self.last_value = float(value)
        return time.perf_counter() - t0

    def describe(self):
        same = "the whole bench workload" if self.n == self.requested else "a bounded sample of the %d of the bench workload" % self.requested
        return ("%d perturbed conformations per step (%s); FP64 C port of the reference path, OpenMP over conformations, %d threads"
                % (self.n, same, self.threads))


def cpu_oracle_rate(mm_system, x_base, term_type, n_conf, target_seconds):
    wl = CpuOracleWorkload(mm_system, x_base, term_type, n_conf, max(1.0, target_seconds / 3.0))
    wl.step()
    times = [wl.step() for _ in range(2)]
    return dict(value=wl.n / float(np.mean(times)), unit="conformations/s", cores=int(wl.threads), kind="port", sample=wl.describe(),
                conformations_per_step=wl.n)


def run_reference(args):
    """--impl reference: the CPU port of the reference's OpenMM Reference-platform path, all host threads, same S as ours."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mm, x0 = load_molecule(args.workload)
    xb = base_conformations(args.workload, x0)
    n_steps = max(1, args.steps + args.warmup)
    wl = CpuOracleWorkload(mm, xb, args.term_type, args.conformations, min(20.0, max(0.05, 150.0 / n_steps)))
    for _ in range(args.warmup):
        wl.step()
    times = [wl.step() for _ in range(args.steps)]
    value = wl.n * len(times) / float(np.sum(times))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "conformations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(args, mm, 1, {"conformations_per_cpu_step": wl.n,
                                                 "note": "CPU arm: every step is one objective evaluation over the bench workload on the host cores"}),
            "cpu_baseline": {"value": value, "unit": "conformations/s", "cores": int(wl.threads), "kind": "port",
                             "sample": wl.describe()},
            "e2e": {"value": value, "unit": "conformations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(args):
    return "%s x %d perturbed conformations in total (sharded over the GPUs), ENERGY+FORCE(%s)+L2 objective, uniform weights" % (
        args.workload, args.conformations, args.term_type)


def config_block(args, mm, world, extra=None):
    cfg = {"workload": workload_name(args), "molecule": args.workload, "n_atoms": mm.n_atoms,
           "conformations_total": args.conformations, "parameter_vectors_per_step": 1, "term_type": args.term_type,
           "l2": "inputs larger than L2 (%.0f MB of conformation tiles in total vs 126 MB per GPU)" % (2 * args.conformations * mm.n_atoms * 24 / 1e6),
           "parallelism": ("conformation shards of %d, 1 all-reduce of 8 doubles per step inside the reduction kernel" % (args.conformations // world))
           if world > 1 else "single GPU"}
    cfg.update(extra or {})
    return cfg


# ---------------------------------------------------------------------------------------------------------------------
class Arm:
    """One sharded workload on this rank: system, objective function, device ensemble."""

    def __init__(self, args, mm, xb, lo, hi, local_rank, pinned=True):
        from paramol_b200.MM_engines.b200_engine import B200Engine
        from paramol_b200.System.system import ParaMolSystem
        from paramol_b200.Parameter_space.parameter_space import ParameterSpace
        from paramol_b200.Objective_function.objective_function import ObjectiveFunction
        from paramol_b200.Objective_function.Properties.energy_property import EnergyProperty
        from paramol_b200.Objective_function.Properties.force_property import ForceProperty
        from paramol_b200.Objective_function.Properties.regularization import Regularization
        import torch
        self.engine = B200Engine(system=mm.copy(), device=local_rank)
        self.system = system = ParaMolSystem(args.workload, self.engine, mm.n_atoms)
        system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True, opt_charges=True, opt_lj=True,
                                              opt_sc=True)
        # ---- synthetic data on the HOST (this rank's shard of the global set), labels = MM(x) + noise -----------------------
        X, e_noise, f_noise = global_slice(xb, lo, hi, mm.n_atoms)
        self.S = S = hi - lo
        system.ref_coordinates = X
        system.n_structures = S
        emm0 = system.get_energies_ensemble()
        fmm0 = system.get_forces_ensemble()
        system.ref_energies = emm0 + e_noise
        system.ref_forces = fmm0 + f_noise
        del fmm0, f_noise
        torch.cuda.synchronize()
        # ---- cold start: pinned host -> HBM tiles (reported, not part of the steady-state metric) ----------------------------
        t0 = time.perf_counter()
        self.ens = system.device_ensemble(pinned=pinned)
        torch.cuda.synchronize()
        self.t_upload = time.perf_counter() - t0
        self.upload_bytes = self.ens.h2d_bytes
        self.ps = ps = ParameterSpace()
        _, x_init = ps.get_optimizable_parameters([system])
        self.x_init = x_init = np.asarray(x_init, dtype=np.float64)
        self.e_prop = EnergyProperty([system], weight=1.0)
        self.f_prop = ForceProperty([system], term_type=args.term_type, weight=1.0)
        self.e_prop.calculate_variance()
        self.f_prop.calculate_variance()
        ps.calculate_prior_widths("default")
        self.r_prop = Regularization(x_init.copy(), ps.prior_widths, "L2", weight=1.0, scaling_factor=1.0)
        self.of = ObjectiveFunction(None, ps, [self.e_prop, self.f_prop, self.r_prop], [system], platform_name="B200",
                                    weighting_method="uniform", checkpoint_freq=10 ** 12)
        self.topo = self.engine.device_topology()


def measure(arm, args, trial, world, dev, local_rank, barrier, max_over_ranks, sample_clocks):
    """(t_e2e, t_dev, t_kernel, objective values, clocks) of one arm: K timed steps after W warm-up steps, both ways."""
    import torch
    of, ens, topo = arm.of, arm.ens, arm.topo
    barrier()   # the ranks finish building their shards at different times; the in-kernel collective waits at most a few seconds
    # ======================= (1) e2e through the public API, host parameter vectors ===========================================
    for x in trial[:args.warmup]:
        of.f(x)
    barrier()
    t0 = time.perf_counter()
    f_vals = [of.f(x) for x in trial[args.warmup:]]
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    # ======================= (2) device-resident: prepare + fused evaluate / reduce / all-reduce ===============================
    slots_dev = torch.from_numpy(np.stack([arm.engine.current_slots()])).to(dev)
    st = of._system_state(0, arm.system)
    comm = st.get("comm")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step(k=None):
        derived = topo.prepare(slots_dev)
        if k is not None:
            ev[k][0].record()
        partials, _, _ = ens.evaluate_reduce(derived, with_forces=True, d_shift=st["d_shift"], comm=comm)
        if k is not None:
            ev[k][1].record()
        if world > 1 and comm is None:
            import torch.distributed as dist
            dist.all_reduce(partials)
        return partials

    for _ in range(args.warmup):
        step()
    barrier()
    clocks = None
    if sample_clocks:
        with ClockSampler(local_rank) as clocks:
            clocks.mark()
            e_all0.record()
            for k in range(args.steps):
                step(k)
            e_all1.record()
            barrier()
    else:
        e_all0.record()
        for k in range(args.steps):
            step(k)
        e_all1.record()
        barrier()
    t_dev = max_over_ranks(e_all0.elapsed_time(e_all1) * 1e-3)
    t_reduce = max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in ev])) * 1e-3)   # pm_eval_reduce as a whole
    # ======================= (3) the dominant kernel alone (roofline): pm_eval = constant-table copy + E+F kernel =================
    derived = topo.prepare(slots_dev)
    for _ in range(3):
        ens.evaluate(derived, with_forces=True, want_fres=True)
    torch.cuda.synchronize()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in kev:
        a.record()
        ens.evaluate(derived, with_forces=True, want_fres=True)
        b.record()
    torch.cuda.synchronize()
    t_kernel = max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in kev])) * 1e-3)
    return t_e2e, t_dev, (t_kernel, t_reduce), f_vals, clocks


def run_ours(args):
    import torch
    import torch.distributed as dist
    from paramol_b200 import _lib
    from paramol_b200.Utils import dist as pdist
    from bench_configs import flops_per_conformation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    os.environ.setdefault("PARAMOL_B200_SPECIALIZE", "require")   # a missing generated kernel is an error here, not a slower run
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    mm, x0 = load_molecule(args.workload)
    xb = base_conformations(args.workload, x0)
    S_total = args.conformations

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- strong scaling: the global set of S_total conformations, sharded ----------------------------------------------------
    lo, hi = pdist.shard_bounds(S_total, world, rank)
    arm = Arm(args, mm, xb, lo, hi, local_rank)
    rng = np.random.default_rng(0)
    trial = [arm.x_init * (1.0 + 0.01 * rng.uniform(-1, 1, arm.x_init.shape)) for _ in range(args.warmup + args.steps)]
    t_e2e, t_dev, (t_kernel, t_reduce), f_vals, clocks = measure(arm, args, trial, world, dev, local_rank, barrier, max_over_ranks, True)
    topo = arm.topo
    comm_kind = pdist.comm_kind(local_rank)
    comm_timeouts = arm.of._state[0]["comm"].timeouts() if arm.of._state[0].get("comm") is not None else 0

    # ---- multi-GPU parity: rank 0 evaluates the SAME global ensemble unsharded and compares the objective ---------------------
    parity = None
    if world > 1:
        if rank == 0:
            with pdist.local_only():
                ref = Arm(args, mm, xb, 0, S_total, local_rank, pinned=False)
                want = [ref.of.f(x) for x in trial[-2:]]
                del ref
            torch.cuda.empty_cache()
            got = f_vals[-2:]
            rel = max(abs(a - b) / abs(b) for a, b in zip(got, want))
            parity = {"sharded": got[-1], "unsharded_on_rank0": want[-1], "max_rel_diff": rel, "tolerance": 1e-10, "ok": bool(rel <= 1e-10)}
        barrier()

    # ---- weak scaling beside it (N > 1): S_total conformations PER GPU ----------------------------------------------------------
    weak = None
    if world > 1 and not args.no_weak:
        del arm.of, arm.ens
        arm.system._ensemble = None
        torch.cuda.empty_cache()
        warm = Arm(args, mm, xb, rank * S_total, (rank + 1) * S_total, local_rank)
        w_e2e, w_dev, (w_kernel, _), _, _ = measure(warm, args, trial, world, dev, local_rank, barrier, max_over_ranks, False)
        weak = {"scaling": "weak", "conformations_per_gpu": S_total, "value": world * S_total * args.steps / w_dev,
                "ms_per_step": 1e3 * w_dev / args.steps, "e2e_value": world * S_total * args.steps / w_e2e,
                "e2e_ms_per_step": 1e3 * w_e2e / args.steps, "kernel_ms": 1e3 * w_kernel}
        del warm
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ======================= roofline denominators ================================================================================
    L = _lib.lib()
    fl, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(L.pm_measure_fma_peak(local_rank, 1, 4000, ctypes.byref(fl), ctypes.byref(ms)))
    fp64_peak = fl.value / 1e12
    clk = clocks.summary() if clocks is not None else {}
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
    f_sm = (clk.get("sm_mhz") or 1965.0) * 1e6
    fp64_nominal = n_sm * 64 * 2 * f_sm / 1e12
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
            hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(REPO, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        traffic = tj.get("%s:%d" % (args.workload, arm.S), {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    counts = (mm.n_atoms, topo.n_bonds, topo.n_angles, topo.n_torsions, topo.n_pairs)
    flops_conf = flops_per_conformation(*counts, term_type=args.term_type)
    bytes_conf = 48 * mm.n_atoms + 16  # f64 coords + f64 reference forces + reference energy in (the sums stay on chip)
    S_rank = arm.S
    achieved_tf = flops_conf * S_rank / t_kernel / 1e12
    achieved_gbs = bytes_conf * S_rank / t_kernel / 1e9
    step_tf = flops_conf * S_rank * args.steps / t_dev / 1e12
    info = topo.launch_info(True)

    # the CPU port is timed beside the single-GPU run only (at N > 1 the driver's reference arm provides it)
    cpu = cpu_oracle_rate(mm, xb, args.term_type, S_total, args.cpu_seconds) if world == 1 else None

    configs = None
    if world == 1 and not args.no_configs and args.workload == "aniline":
        import bench_configs
        del arm.of, arm.ens
        arm.system._ensemble = None
        torch.cuda.empty_cache()
        configs = bench_configs.run_all(fp64_peak)

    value = S_total * args.steps / t_dev
    line = {
        "metric": METRIC, "value": value, "unit": "conformations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_block(args, mm, world, {"conformations_per_gpu": S_rank, "collective": comm_kind}),
        "e2e": {"value": S_total * args.steps / t_e2e, "unit": "conformations/s", "h2d_bytes_per_step": topo.n_slots * 8,
                "d2h_bytes_per_step": 8 * 8, "ms_per_step": 1e3 * t_e2e / args.steps,
                "objective_evals_per_s": args.steps / t_e2e, "api": "ObjectiveFunction.f(x) with host x (one CUDA-graph launch per call through pm_graph_launch / pm_graph_wait; the slots row is read from pinned host memory by pm_prepare_kernel and the 8 sums are written to pinned host memory by the row-sum kernel: those are the h2d / d2h bytes)",
                "cold_upload": {"seconds": arm.t_upload, "bytes": arm.upload_bytes,
                                "note": "one-time pinned-host -> HBM staging of the conformation tiles, outside the step"}},
        "gpu_launches": (4 if getattr(topo, "specialized", False) else 3) * args.steps,
        "gpu_launches_note": ("per step: pm_prepare_kernel, pm_ef_spec_kernel, pm_reduce_stage1, pm_reduce_rows_kernel (row sum + all-reduce over NVLink peer memory)"
                              if getattr(topo, "specialized", False) else
                              "per step: pm_prepare_kernel, pm_ef_kernel (residual reduction in its epilogue), pm_reduce_rows_kernel (row sum + all-reduce)"),
        "objective_evals_per_s": args.steps / t_dev,
        "roofline": {"bound": "fp64", "kernel": ("pm_ef_spec_kernel (topology-specialised code generated for this molecule, csrc/specialize.py)"
                                                 if getattr(topo, "specialized", False) else "pm_ef_kernel<true> (topology interpreter)"),
                     "specialize_status": getattr(topo, "specialize_status", None),
                     "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved_tf / fp64_peak, "frac_of_step": step_tf / fp64_peak, "traffic": traffic,
                     "peak_source": "FP64 pipe peak measured live by pm_measure_fma_peak on this GPU (DFMA with a uniform addend = DMMA rate; 3-register-operand DFMA streams reach 0.92 of it)",
                     "peak_nominal": fp64_nominal, "peak_nominal_note": "%d SMs x 64 DFMA/clk x 2 flop x %.0f MHz (sampled under load)" % (n_sm, f_sm / 1e6),
                     "frac_of_nominal": achieved_tf / fp64_nominal,
                     "precision_note": "FP64 storage and arithmetic (objective tolerance 1e-8, DESIGN.md section 3): against the FP32 CUDA-core peak SURVEY 8(d) named the same kernel is half this fraction",
                     "flops_per_conformation": flops_conf, "kernel_ms": 1e3 * t_kernel,
                     "kernel_ms_note": "the E+F kernel alone (pm_eval: constant-table copy + kernel), CUDA events on the launch stream, timed in a separate loop of the same length right after the step loop",
                     "eval_reduce_ms": 1e3 * t_reduce,
                     "eval_reduce_ms_note": "pm_eval_reduce inside the step loop: E+F kernel + residual sums + row-sum kernel (with the all-reduce at N > 1)",
                     "kernel_share_of_step": t_kernel * args.steps / t_dev, "launch": info,
                     "hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": achieved_gbs / hbm_peak, "bytes_per_conformation": bytes_conf, "peak_source": hbm_src}},
        "cpu_baseline": cpu,
        "clocks": clk,
        "objective_value_last_step": f_vals[-1] if f_vals else None,
    }
    if world > 1:
        line["multi_gpu_parity"] = parity
        line["collective"] = {"kind": comm_kind, "timeouts": comm_timeouts,
                              "note": "p2p = one-shot all-reduce inside pm_reduce_rows_kernel over NVLink peer memory (csrc/pm_comm.cu)"}
        if weak is not None:
            line["weak"] = weak
    if configs is not None:
        line["configs"] = configs
    print(json.dumps(line))
    if parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: the sharded objective differs from the unsharded one: %r" % (parity,))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="aniline")
    ap.add_argument("--conformations", type=int, default=1 << 20, help="conformations in total (sharded over the GPUs)")
    ap.add_argument("--term-type", dest="term_type", default="components", choices=["components", "norm"])
    ap.add_argument("--cpu-seconds", dest="cpu_seconds", type=float, default=10.0)
    ap.add_argument("--no-configs", dest="no_configs", action="store_true", help="skip the side configurations (N = 1)")
    ap.add_argument("--no-weak", dest="no_weak", action="store_true", help="skip the weak-scaling repetition (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
