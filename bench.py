#!/usr/bin/env python
"""
bench.py -- headline benchmark of the B200-native ParaMol objective-evaluation path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload aniline|lig50|lig60|norfloxacin]
                    [--conformations S]

Workload (BASELINE.json configs[1]): the reference test-suite molecule (aniline, 14 atoms), S = 2^20 synthetic perturbed
conformations PER GPU (weak scaling: conformations are sharded by rank, the only collective is one allreduce of 8
doubles per objective call), objective = ENERGY + FORCE("components") + L2 REGULARIZATION with uniform weights.
A "step" is ONE objective evaluation over all conformations.

Printed JSON line (rank 0):
  value      conformation E+F evaluations / s, whole job, parameters and conformations resident in HBM
             (pm_prepare_params + pm_eval + pm_reduce_objective + allreduce per step, CUDA-event timed)
  e2e        the same metric through the public API ``ObjectiveFunction.f(x)`` with a HOST parameter vector: numpy
             plumbing, pinned H2D of the slot row, kernels, allreduce, D2H of the partial sums, every step
  roofline   dominant kernel (pm_ef_spec_kernel for small molecules, else pm_ef_kernel<true>): algorithmic FP64 flops / CUDA-event kernel time against the DFMA peak
             measured live on this GPU (the kernel is FP64-pipe bound); ``hbm`` holds the same kernel
             against the measured copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the FP64 CPU oracle (oracle/oracle.c, OpenMP over conformations, all host cores) on a bounded sample
``--impl reference`` times that CPU oracle alone (OpenMM, where ParaMol's arithmetic lives, is not installable here).
"""
import argparse
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


def flops_per_conformation(topo_counts, term_type):
    """36/pair + 30/bond + 90/angle + 160/torsion + epilogue (9N+4 norm | 21N+4 components)."""
    n, nb, na, nt, npair = topo_counts
    epi = (21 * n + 4) if term_type == "components" else (9 * n + 4)
    return 36 * npair + 30 * nb + 90 * na + 160 * nt + epi


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
    The CPU arm: the FP64 oracle port of the reference path (oracle/oracle.c: E+F of every conformation, OpenMP over
    conformations = the analogue of the reference's fork pool; per-conformation force residual in C; the scalar
    reductions of EnergyProperty / ForceProperty in numpy) on a bounded sample of the bench workload.
    """

    def __init__(self, mm_system, x_base, term_type, seconds_per_step, threads=None):
        from oracle import oracle as O
        from paramol_b200.Utils.synthetic import perturbed_conformations
        self.O = O
        self.mm = mm_system
        # every host core this process may run on (torchrun exports OMP_NUM_THREADS=1; the C port takes the count explicitly)
        self.threads = threads or max(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1), 1)
        probe = perturbed_conformations(x_base, 4096, seed=99)
        O.energies_forces(mm_system, probe[:64], n_threads=self.threads)  # spin up the OpenMP team
        t0 = time.perf_counter()
        O.energies_forces(mm_system, probe, n_threads=self.threads)
        dt = time.perf_counter() - t0
        self.n = int(min(max(4096, 4096 * seconds_per_step / max(dt, 1e-6)), 2_000_000))
        self.X = perturbed_conformations(x_base, self.n, seed=100)
        rng = np.random.default_rng(3)
        E, F = O.energies_forces(mm_system, self.X, n_threads=self.threads)
        self.Eref = E + rng.normal(0, 5, E.shape)
        self.Fref = F + rng.normal(0, 50, F.shape)
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
        return ("%d perturbed conformations of the bench molecule per step; FP64 C oracle, OpenMP over conformations, %d threads"
                % (self.n, self.threads))


def cpu_oracle_rate(mm_system, x_base, term_type, target_seconds):
    wl = CpuOracleWorkload(mm_system, x_base, term_type, max(1.0, target_seconds / 3.0))
    wl.step()
    times = [wl.step() for _ in range(2)]
    return dict(value=wl.n / float(np.mean(times)), unit="conformations/s", cores=int(wl.threads), kind="port", sample=wl.describe())


def run_reference(args):
    """--impl reference: the CPU oracle (port of the reference's OpenMM Reference-platform path), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mm, x0 = load_molecule(args.workload)
    xb = base_conformations(args.workload, x0)
    n_steps = max(1, args.steps + args.warmup)
    wl = CpuOracleWorkload(mm, xb, args.term_type, min(10.0, max(0.05, 100.0 / n_steps)))
    for _ in range(args.warmup):
        wl.step()
    times = [wl.step() for _ in range(args.steps)]
    value = wl.n * len(times) / float(np.sum(times))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "conformations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "conformations_per_gpu": args.conformations, "term_type": args.term_type,
                       "conformations_per_cpu_step": wl.n,
                       "note": "CPU arm: each step is one objective evaluation over a bounded sample of the workload; the "
                               "reference is O(S), so conformations/s is size independent"},
            "cpu_baseline": {"value": value, "unit": "conformations/s", "cores": int(wl.threads), "kind": "port",
                             "sample": wl.describe()},
            "e2e": {"value": value, "unit": "conformations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(args):
    return "%s x %d perturbed conformations per GPU, ENERGY+FORCE(%s)+L2 objective, uniform weights" % (
        args.workload, args.conformations, args.term_type)


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from paramol_b200 import _lib
    from paramol_b200.MM_engines.b200_engine import B200Engine
    from paramol_b200.System.system import ParaMolSystem
    from paramol_b200.Parameter_space.parameter_space import ParameterSpace
    from paramol_b200.Objective_function.objective_function import ObjectiveFunction
    from paramol_b200.Objective_function.Properties.energy_property import EnergyProperty
    from paramol_b200.Objective_function.Properties.force_property import ForceProperty
    from paramol_b200.Objective_function.Properties.regularization import Regularization
    from paramol_b200.Utils.synthetic import perturbed_conformations

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    mm, x0 = load_molecule(args.workload)
    xb = base_conformations(args.workload, x0)
    S = args.conformations
    engine = B200Engine(system=mm, device=local_rank)
    system = ParaMolSystem(args.workload, engine, mm.n_atoms)
    system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True, opt_charges=True, opt_lj=True,
                                          opt_sc=True)
    # ---- synthetic data on the HOST (per-rank shard, seeded by rank), labels = MM(x_ref) + noise --------------------------
    X = perturbed_conformations(xb, S, sigma=0.005, seed=1234 + rank)
    system.ref_coordinates = X
    system.n_structures = S
    rng = np.random.default_rng(4321 + rank)
    t0 = time.perf_counter()
    emm0 = system.get_energies_ensemble()
    fmm0 = system.get_forces_ensemble()
    system.ref_energies = emm0 + rng.normal(0.0, 5.0, S)
    system.ref_forces = fmm0 + rng.normal(0.0, 50.0, fmm0.shape)
    del fmm0
    torch.cuda.synchronize()
    # ---- cold start: pinned host -> HBM tiles (reported, not part of the steady-state metric) --------------------------------
    t0 = time.perf_counter()
    ens = system.device_ensemble(pinned=True)
    torch.cuda.synchronize()
    t_upload = time.perf_counter() - t0
    upload_bytes = ens.h2d_bytes

    ps = ParameterSpace()
    _, x_init = ps.get_optimizable_parameters([system])
    x_init = np.asarray(x_init, dtype=np.float64)
    e_prop = EnergyProperty([system], weight=1.0)
    f_prop = ForceProperty([system], term_type=args.term_type, weight=1.0)
    e_prop.calculate_variance()
    f_prop.calculate_variance()
    ps.calculate_prior_widths("default")
    r_prop = Regularization(x_init.copy(), ps.prior_widths, "L2", weight=1.0, scaling_factor=1.0)
    of = ObjectiveFunction(None, ps, [e_prop, f_prop, r_prop], [system], platform_name="B200", weighting_method="uniform",
                           checkpoint_freq=10 ** 12)
    rng = np.random.default_rng(0)
    trial = [x_init * (1.0 + 0.01 * rng.uniform(-1, 1, x_init.shape)) for _ in range(args.warmup + args.steps)]

    topo = engine.device_topology()
    counts = (mm.n_atoms, topo.n_bonds, topo.n_angles, topo.n_torsions, topo.n_pairs)
    flops_conf = flops_per_conformation(counts, args.term_type)
    bytes_conf = 48 * mm.n_atoms + 16  # f64 coords + f64 reference forces + reference energy in, energy out

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

    # ======================= (1) e2e through the public API, host parameter vectors ===========================================
    for x in trial[:args.warmup]:
        of.f(x)
    barrier()
    t0 = time.perf_counter()
    f_vals = [of.f(x) for x in trial[args.warmup:]]
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    h2d_step = topo.n_slots * 8
    d2h_step = 8 * 8

    # ======================= (2) device-resident: prepare + eval + reduce (+ allreduce) ========================================
    slots_dev = torch.from_numpy(np.stack([engine.current_slots()])).to(dev)
    st = of._system_state(0, system)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step(k=None):
        derived = topo.prepare(slots_dev)
        if k is not None:
            ev[k][0].record()
        emm, fres, _ = ens.evaluate(derived, with_forces=True, want_fres=True)
        if k is not None:
            ev[k][1].record()
        partials = ens.reduce(emm, fres, d_shift=st["d_shift"])
        if world > 1:
            dist.all_reduce(partials)
        return partials

    for _ in range(args.warmup):
        step()
    barrier()
    with ClockSampler(local_rank) as clocks:
        clocks.mark()
        e_all0.record()
        for k in range(args.steps):
            step(k)
        e_all1.record()
        barrier()
    t_dev = max_over_ranks(e_all0.elapsed_time(e_all1) * 1e-3)
    t_kernel = max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in ev])) * 1e-3)

    # ======================= (3) optional mode, reported separately: nonbonded terms cached across calls =====================
    # (legitimate here because the reference's all-optimizable setup never pushes nonbonded parameters, openmm.py:752-754;
    #  NOT part of `value` / `e2e`, which time full evaluations)
    of_c = ObjectiveFunction(None, ps, [e_prop, f_prop, r_prop], [system], platform_name="B200", weighting_method="uniform",
                             checkpoint_freq=10 ** 12, cache_nonbonded=True)
    for x in trial[:args.warmup]:
        of_c.f(x)
    barrier()
    t0 = time.perf_counter()
    f_vals_c = [of_c.f(x) for x in trial[args.warmup:]]
    barrier()
    t_e2e_cached = max_over_ranks(time.perf_counter() - t0)
    cache_ok = bool(np.allclose(f_vals_c, f_vals, rtol=1e-12, atol=0))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ======================= roofline denominators ================================================================================
    L = _lib.lib()
    import ctypes
    fl, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(L.pm_measure_fma_peak(local_rank, 1, 4000, ctypes.byref(fl), ctypes.byref(ms)))
    fp64_peak = fl.value / 1e12
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
        key = "%s:%d" % (args.workload, S)
        traffic = tj.get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    achieved_tf = flops_conf * S / t_kernel / 1e12
    achieved_gbs = bytes_conf * S / t_kernel / 1e9
    info = topo.launch_info(True)

    # the CPU port is timed beside the single-GPU run only (at N > 1 the driver's reference arm provides it)
    cpu = cpu_oracle_rate(mm, xb, args.term_type, args.cpu_seconds) if world == 1 else None

    value = world * S * args.steps / t_dev
    line = {
        "metric": METRIC, "value": value, "unit": "conformations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "molecule": args.workload, "n_atoms": mm.n_atoms,
                   "conformations_per_gpu": S, "parameter_vectors_per_step": 1, "term_type": args.term_type,
                   "l2": "inputs larger than L2 (%.0f MB of tiles per GPU vs 126 MB)" % (2 * S * mm.n_atoms * 24 / 1e6),
                   "parallelism": "conformation shards, 1 allreduce of 8 doubles per step" if world > 1 else "single GPU"},
        "e2e": {"value": world * S * args.steps / t_e2e, "unit": "conformations/s", "h2d_bytes_per_step": h2d_step,
                "d2h_bytes_per_step": d2h_step, "ms_per_step": 1e3 * t_e2e / args.steps,
                "objective_evals_per_s": args.steps / t_e2e, "api": "ObjectiveFunction.f(x) with host x",
                "cold_upload": {"seconds": t_upload, "bytes": upload_bytes,
                                "note": "one-time pinned-host -> HBM staging of the conformation tiles, outside the step"}},
        "gpu_launches": 4 * args.steps,
        "objective_evals_per_s": args.steps / t_dev,
        "roofline": {"bound": "fp64", "kernel": ("pm_ef_spec_kernel (topology-specialised code generated for this molecule, csrc/specialize.py)"
                                                 if getattr(topo, "specialized", False) else "pm_ef_kernel<true>"),
                     "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved_tf / fp64_peak, "traffic": traffic,
                     "peak_source": "FP64 pipe peak measured live by pm_measure_fma_peak on this GPU (DFMA with a uniform addend = DMMA rate; 3-register-operand DFMA streams reach 0.92 of it)",
                     "flops_per_conformation": flops_conf, "kernel_ms": 1e3 * t_kernel,
                     "kernel_share_of_step": t_kernel * args.steps / t_dev, "launch": info,
                     "hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": achieved_gbs / hbm_peak, "bytes_per_conformation": bytes_conf, "peak_source": hbm_src}},
        "nonbonded_cache_mode": {"e2e_value": world * S * args.steps / t_e2e_cached, "unit": "conformations/s",
                                 "ms_per_step": 1e3 * t_e2e_cached / args.steps, "same_objective_values": cache_ok,
                                 "note": "optional ObjectiveFunction(cache_nonbonded=True): pair terms evaluated once while the "
                                         "nonbonded parameters stay unchanged, bonded terms every call; not the headline"},
        "cpu_baseline": cpu,
        "clocks": clocks.summary(),
        "objective_value_last_step": f_vals[-1] if f_vals else None,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="aniline")
    ap.add_argument("--conformations", type=int, default=1 << 20)
    ap.add_argument("--term-type", dest="term_type", default="components", choices=["components", "norm"])
    ap.add_argument("--cpu-seconds", dest="cpu_seconds", type=float, default=10.0)
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
