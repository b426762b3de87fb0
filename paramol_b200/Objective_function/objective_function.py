# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`ObjectiveFunction` -- the callable handed to the optimisers; same constructor keywords, attributes and
methods as the reference class (``ParaMol/Objective_function/objective_function.py:26-361``):
``f(parameters_values, opt_mode=True)``, ``calculate_objective_function``, ``reset_f_count``, ``get_f_count``,
``init_parallel``.

What changes underneath ``f``:

* reference: ``update_systems`` (O(#terms) SWIG calls) -> per-conformation OpenMM loops for energies, then
  again for forces -> numpy/Python property loops;
* here: ``update_systems`` (numpy gathers) -> ONE host->device copy of the slot rows -> ``pm_prepare_params`` ->
  ``pm_eval`` (energies, forces and the per-conformation force residual in one pass) -> ``pm_reduce_objective``
  -> ONE allreduce of 8 doubles per (system, vector) when ``torch.distributed`` is initialised -> ONE
  device->host copy of those doubles -> scalar assembly.

``f_batch(X)`` evaluates P trial vectors per launch (Monte Carlo / simulated annealing / finite differences,
``ParaMol/Optimizers/monte_carlo.py:101``, ``gradient_descent.py:106-121``, ``simulated_annealing.py:104``)
without mutating the parameter space.  ``fused=False`` keeps the reference's structure (ensemble arrays brought
to the host through ``ParaMolSystem.get_*_ensemble`` and reduced by the numpy property formulas); it exists for
parity tests of the seam, not for speed.
"""
import ctypes
import logging
import os
import time

import numpy as np

from .. import _lib
from ..Utils import dist as _dist

NPARTIAL = 8


class ObjectiveFunction:
    """
    Parameters
    ----------
    restart_settings : dict or None
    parameter_space : :obj:`paramol_b200.Parameter_space.parameter_space.ParameterSpace`
    properties : list of properties
    systems : list of :obj:`paramol_b200.System.system.ParaMolSystem`
    platform_name : str
        Kept from the reference ("Reference", "CPU", "OpenCL", "CUDA") plus "B200"; evaluation always runs on the
        B200 path (there is no CPU fallback).
    parallel : bool
        Kept for signature compatibility.  Data parallelism here is one process per GPU (``torch.distributed``),
        each rank holding a shard of the conformations; see ``ParaMolSystem.shard_for_rank``.
    weighting_method : str
        "uniform", "boltzmann", "non_boltzmann" or "manual".
    weighting_temperature : float (K) or quantity with ``_value``
    checkpoint_freq : int
    fused : bool
        True (default): device-side residual reduction.  False: reference-structured seam path.
    cache_nonbonded : bool
        When the trial vectors of a call share their nonbonded parameters with the previous call (bonded-only fits; the
        reference's all-optimizable setup, where ``set_nonbonded_parameters`` pushes nothing), evaluate the pair terms
        once, keep E_nb and F_nb per conformation on the device and re-evaluate only the bonded terms.  Off by default
        (and off in ``bench.py``'s headline numbers, which time full evaluations).
    delta_evaluation : bool
        ``f`` keeps the per-conformation results of its vector as a BASE; ``f_batch`` rows that differ from it in a few terms
        (the single-parameter moves of Monte Carlo / simulated annealing, ``monte_carlo.py:89-100``) are evaluated through
        ``pm_delta_eval``: only the affected terms are recomputed.  Values agree with full evaluations to rounding
        (~1e-13 relative); off by default.
    """

    #: energy-only batches of at least this many vectors go through the contraction kernel (``pm_energy_batch_reduce``); below it the
    #: per-vector term pass is cheaper than building the coefficient matrix and padding the block of 128 vectors
    ENERGY_BATCH_MIN_VECTORS = 16

    def __init__(self, restart_settings, parameter_space, properties, systems, platform_name="B200", parallel=False,
                 weighting_method="uniform", weighting_temperature=300.0, checkpoint_freq=1000, fused=True,
                 cache_nonbonded=False, delta_evaluation=False):
        self.restart_settings = restart_settings
        self.parameter_space = parameter_space
        self.properties = properties
        self.systems = systems
        self._platform = platform_name
        self._parallel = parallel
        self._checkpoint_freq = checkpoint_freq
        self._weighting_method = weighting_method
        self._weighting_temperature = weighting_temperature
        self._f_count = 0
        self._fused = bool(fused)
        self._cache_nonbonded = bool(cache_nonbonded)
        self._delta = bool(delta_evaluation)
        self._graphs = os.environ.get("PARAMOL_B200_GRAPHS", "1") != "0"
        self._sync_stream = None
        self.n_delta_rows = 0
        self._total_n_batches = None
        self._batch_lims = None
        self._parallel_function = None
        self._state = {}
        self._state_version = 0
        self.last_timings = {}
        if self._parallel:
            self.init_parallel()

    # ------------------------------------------------------------------------------------------
    #                                       PUBLIC METHODS
    # ------------------------------------------------------------------------------------------
    def reset_f_count(self):
        self._f_count = 0
        return self._f_count

    def get_f_count(self):
        return self._f_count

    def invalidate(self):
        """Drop every cached per-system device state (force metric, energy shift, static weights, captured graphs).  Needed only
        after IN-PLACE edits of property variances or manual weights (replacing the arrays is detected by itself); see
        :meth:`paramol_b200.System.system.ParaMolSystem.invalidate_device_data` for in-place edits of the reference data."""
        self._state_version += 1
        self._state = {}

    def init_parallel(self):
        """
        Batch limits with the reference's rule ``batch_size = int(S / n_cpus) + 1`` (reference ``:112-160``).  No worker
        pool is created: on the B200 path conformations are already evaluated in parallel on the device.
        """
        logging.info("Initializing parallel objective function.")
        self._total_n_batches = 0
        self._batch_lims = []
        for system in self.systems:
            n_structures = len(system.ref_coordinates)
            self._total_n_batches = self._total_n_batches + system.n_cpus
            batch_size = int(n_structures / float(system.n_cpus)) + 1
            self._batch_lims.append([[n * batch_size, (n + 1) * batch_size] for n in range(system.n_cpus)])
        return self._total_n_batches

    def calculate_objective_function(self, parameters_values, fmm, emm, esp):
        """Reference-structured assembly from host arrays (reference ``:162-226``)."""
        objective_function = 0.0
        for system in self.systems:
            system.compute_conformations_weights(temperature=self._weighting_temperature, emm=emm,
                                                 weighting_method=self._weighting_method)
        for prop in self.properties:
            if prop.name == "ENERGY":
                objective_function += prop.weight * np.sum(prop.calculate_property(emm))
            elif prop.name == "FORCE":
                objective_function += prop.weight * np.sum(prop.calculate_property(fmm))
            elif prop.name == "ESP":
                objective_function += prop.weight * np.sum(prop.calculate_property(esp))
            elif prop.name == "REGULARIZATION":
                objective_function += prop.weight * prop.calculate_property(parameters_values)
        return objective_function

    # ------------------------------------------------------------------------------------------
    #                                    Objective function
    # ------------------------------------------------------------------------------------------
    def f(self, parameters_values, opt_mode=True):
        """
        Objective function value for ``parameters_values`` (scaled if the parameter space is preconditioned).
        Side effects as in the reference: parameters, force fields and engines are updated, ``system.weights`` and
        ``property.value`` are set, ``_f_count`` is incremented in ``opt_mode``; NaN becomes 1e16.
        """
        start_time = time.time()
        if opt_mode:
            self._f_count += 1
        if self._fused:
            # the slot rows first; the rest of the update (value lists, MM-system arrays, RESP charges) runs while the device
            # evaluates (``between``), and in any case before this call returns
            deferred = []
            try:
                self.parameter_space.update_systems(self.systems, parameters_values, defer=deferred)
                slots = self.parameter_space.last_slots
                objective_function = float(self._evaluate_fused(slots, np.asarray(parameters_values, dtype=np.float64)[None, :],
                                                               set_values=True, between=deferred)[0])
            finally:
                while deferred:
                    deferred.pop()()
        else:
            self.parameter_space.update_systems(self.systems, parameters_values)
            fmm, emm, esp = self._run_serial()
            objective_function = self.calculate_objective_function(parameters_values, fmm, emm, esp)
        if np.isnan(objective_function):
            objective_function = 1e16
        if self._f_count % self._checkpoint_freq == 1:
            for system in self.systems:
                system.engine.write_system_xml("{}_checkpoint.xml".format(system.name))
        if not opt_mode:
            self._print_table(time.time() - start_time)
        return objective_function

    def f_and_grad(self, parameters_values, opt_mode=True):
        """
        (f, d f / d parameters_values) with the ANALYTIC gradient: one forward pass (energies, forces, residual sums) and
        one ``pm_grad`` pass per system instead of the P_opt + 1 objective evaluations of a finite-difference gradient
        (the only gradient the reference has: scipy ``jac="2-point"``, ``ParaMol/Utils/settings.py:64``, and
        ``GradientDescent``, ``ParaMol/Optimizers/gradient_descent.py:106-121``).  Side effects as :meth:`f`.

        Supported: ENERGY, FORCE, ESP and REGULARIZATION properties; uniform, Boltzmann and manual weights (which do not
        depend on the parameters).  NON_BOLTZMANN weights depend on the MM energies and raise ``NotImplementedError``.
        Derivatives with respect to a vanishing lj_eps are reported as 0 (sqrt is not differentiable there).
        """
        if not self._fused:
            raise NotImplementedError("the analytic gradient needs the fused device path")
        if self._weighting_method.upper() == "NON_BOLTZMANN":
            raise NotImplementedError("analytic gradient: NON_BOLTZMANN weights depend on the parameters")
        for system in self.systems:
            if getattr(system.engine.system, "nonbonded_method", "NoCutoff") != "NoCutoff":
                raise NotImplementedError("analytic gradient: the cutoff / reaction-field variant ({}) is not implemented; "
                                          "use finite differences through f_batch".format(system.name))
        start_time = time.time()
        x = np.asarray(parameters_values, dtype=np.float64)
        self.parameter_space.update_systems(self.systems, parameters_values)
        if opt_mode:
            self._f_count += 1
        slots = [system.engine.current_slots()[None, :] for system in self.systems]
        value, grad = self._evaluate_fused_grad(slots, x)
        if np.isnan(value):
            value = 1e16
        if self._f_count % self._checkpoint_freq == 1:
            for system in self.systems:
                system.engine.write_system_xml("{}_checkpoint.xml".format(system.name))
        if not opt_mode:
            self._print_table(time.time() - start_time)
        return value, grad

    def grad(self, parameters_values):
        """Analytic gradient only (see :meth:`f_and_grad`); usable as ``jac`` of ``scipy.optimize.minimize``."""
        return self.f_and_grad(parameters_values)[1]

    def f_batch(self, parameters_matrix):
        """
        Objective values, numpy [P], for the rows of ``parameters_matrix`` [P, P_opt] in one set of launches.  Nothing
        is mutated except ``_f_count`` (+P).  NaN rows become 1e16.
        """
        X = np.atleast_2d(np.asarray(parameters_matrix, dtype=np.float64))
        slots = self.parameter_space.update_systems_batch(self.systems, X)
        values = self._evaluate_fused(slots, X, set_values=False)
        self._f_count += X.shape[0]
        values = np.where(np.isnan(values), 1e16, values)
        return values

    # ------------------------------------------------------------------------------------------
    #                                       PRIVATE METHODS
    # ------------------------------------------------------------------------------------------
    def _run_serial(self):
        fmm, emm, esp = [], [], []
        for prop in self.properties:
            if prop.systems is not None:
                for system in prop.systems:
                    if prop.name == "ENERGY":
                        emm.append(system.get_energies_ensemble())
                    elif prop.name == "FORCE":
                        fmm.append(system.get_forces_ensemble())
                    elif prop.name == "ESP":
                        esp.append(system.get_esp_ensemble())
        return fmm, emm, esp

    def _print_table(self, elapsed):
        print("!=================================================================================!")
        print("!                           Objective Function Analysis                           !")
        print("! {:<25s}{:^17s}{:<1s}{:^17s}{:<2s}{:^17s} !".format("PROPERTY", "Value", "x", "Weight", "=", "Contribution"))
        print("!---------------------------------------------------------------------------------!")
        obj_fun_total = 0.0
        for prop in self.properties:
            print("! {:<25s}{:^17.8f} {:^17.8f} {:^17.8f}  !".format(prop.name, prop.value, prop.weight, prop.value * prop.weight))
            obj_fun_total = obj_fun_total + prop.value * prop.weight
        print("! {:<61s}{:^17.8f}  !".format("TOTAL", obj_fun_total))
        print("!---------------------------------------------------------------------------------!")
        print("! {:<25s}{:<10}{:<25s}{:<8.3f}{:11s} !".format("Function Evaluations:", self._f_count, "Time (s):", elapsed, " "))
        print("!=================================================================================!\n")

    def _prop(self, name):
        for prop in self.properties:
            if prop.name == name:
                return prop
        return None

    def _system_state(self, s_idx, system):
        """Per-system device state: ensemble, force metric, static weights, energy shift."""
        import torch
        e_prop, f_prop = self._prop("ENERGY"), self._prop("FORCE")
        need_e = e_prop is not None and system in e_prop.systems
        need_f = f_prop is not None and system in f_prop.systems
        ens = system.device_ensemble()
        key = (id(ens), need_e, need_f, None if f_prop is None else f_prop.term_type,
               None if f_prop is None else id(f_prop.variance), self._weighting_method.upper(), id(system.wham_weights),
               getattr(system, "_data_version", 0), self._state_version)
        st = self._state.get(s_idx)
        if st is not None and st["key"] == key:
            return st
        st = dict(key=key, ens=ens, need_e=need_e, need_f=need_f)
        if need_f:
            assert f_prop.variance is not None, "ForceProperty.calculate_variance() was not called."
            ens.topology.set_force_metric(f_prop.metric(f_prop.systems.index(system)))
        st["n_global"] = system.n_structures_global()
        # multi-GPU: the peer-memory communicator whose all-reduce runs inside the reduction kernel (None: single rank, gloo, or
        # IPC unavailable -> NCCL all-reduce of the partial sums); creating it is a collective, done here on every rank
        st["comm"] = _dist.device_comm(ens.device.index)
        st["d_shift"] = 0.0
        if need_e:
            assert e_prop.variance is not None, "EnergyProperty.calculate_variance() was not called."
            k = e_prop.systems.index(system)
            st["d_shift"] = float(e_prop.mean_ref[k])
        # static weights (everything except NON_BOLTZMANN, which depends on the MM energies of the call)
        method = self._weighting_method.upper()
        st["w_scale"] = 1.0
        wham = system.wham_weights
        wham_is_array = np.ndim(wham) > 0
        if method == "UNIFORM" and not wham_is_array:
            system.compute_conformations_weights(weighting_method="UNIFORM")
            ens.set_weights(None)
            st["w_scale"] = float(wham) / st["n_global"]
        elif method in ("UNIFORM", "BOLTZMANN", "MANUAL"):
            system.compute_conformations_weights(temperature=self._weighting_temperature, weighting_method=method)
            ens.set_weights(np.asarray(system.weights, dtype=np.float64) * np.asarray(wham, dtype=np.float64))
        elif method == "NON_BOLTZMANN":
            st["beta"] = system.beta(self._weighting_temperature)
            st["wham_dev"] = torch.from_numpy(np.asarray(wham, dtype=np.float64) * np.ones(system.n_structures)).to(ens.device) \
                if wham_is_array else None
            st["wham_scalar"] = 1.0 if wham_is_array else float(wham)
        else:
            raise NotImplementedError("Weighting method {} is not implemented.".format(self._weighting_method))
        self._state[s_idx] = st
        return st

    def _non_boltzmann_weights(self, st, system, emm):
        """w = softmax(-beta (d - mean d)), d = E_mm - E_ref, per parameter vector; [P,S] on the device (global over ranks)."""
        import torch
        ens = st["ens"]
        d = emm - ens.eref[None, :]
        sums = torch.stack([d.sum(dim=1)], dim=1)
        _dist.allreduce_sum_tensor_(sums)
        mean = sums[:, 0] / st["n_global"]
        w = torch.exp(-st["beta"] * (d - mean[:, None]))
        z = w.sum(dim=1, keepdim=True)
        _dist.allreduce_sum_tensor_(z)
        # The reference (and ParaMolSystem.compute_conformations_weights) evaluates this under np.seterr(all='raise') and falls
        # back to uniform weights on ANY floating-point error (system.py:193, 215-221): overflow of the exponentials or of their
        # sum, but also UNDERFLOW of an exponential or of w / z (a result below the smallest normal double).  Same conditions
        # here, decided globally over ranks.
        tiny = torch.finfo(torch.float64).tiny
        w_min = w.min(dim=1, keepdim=True).values
        _dist.allreduce_min_tensor_(w_min)
        bad = ~torch.isfinite(z) | (z <= 0) | (w_min < tiny) | (w_min / z < tiny)
        w = torch.where(bad, torch.full_like(w, 1.0 / st["n_global"]), w / z)
        if emm.shape[0] == 1:
            system.weights = w[0].cpu().numpy()
        if st["wham_dev"] is not None:
            w = w * st["wham_dev"][None, :]
        else:
            w = w * st["wham_scalar"]
        return w.contiguous()

    def assemble_from_partials(self, per_system, w_scales, set_values=False):
        """
        ENERGY and FORCE contributions, numpy [P], from the ALLREDUCED partial sums of every system
        (``per_system[s]`` = float64 [P, 8] as written by ``pm_reduce_objective``, or None):
            ENERGY_s = ([2] - 2 m [1] + m^2 [3]) / var_s(E_ref),  m = [0] / [5]      (energy_property.py:99-108)
            FORCE_s  = [4] / (3 N_s)                                                (force_property.py:85-117)
        ``w_scales[s]`` folds the normalisation of uniform weights (device weights are 1 in that case).
        """
        e_prop, f_prop = self._prop("ENERGY"), self._prop("FORCE")
        n_vec = next(hp.shape[0] for hp in per_system if hp is not None)
        values = np.zeros(n_vec)
        for hp, scale in zip(per_system, w_scales):
            if hp is not None and scale != 1.0:
                hp[:, 1:5] *= scale
        if n_vec == 1:
            # the optimiser's f(x): the same expressions on Python floats (identical IEEE operations in the same order,
            # without a dozen one-element numpy calls)
            rows = [None if hp is None else hp[0].tolist() for hp in per_system]
            value = 0.0
            if e_prop is not None:
                val = 0.0
                for k, system in enumerate(e_prop.systems):
                    r = rows[self.systems.index(system)]
                    m = r[0] / r[5]
                    val += (r[2] - 2.0 * m * r[1] + m * m * r[3]) / float(e_prop.variance[k])
                if set_values:
                    e_prop.value = val
                value += e_prop.weight * val
            if f_prop is not None:
                val = 0.0
                for system in f_prop.systems:
                    val += rows[self.systems.index(system)][4] / (3 * system.n_atoms)
                if set_values:
                    f_prop.value = val
                value += f_prop.weight * val
            return np.array([value])
        if e_prop is not None:
            vals = np.zeros(n_vec)
            for system in e_prop.systems:
                hp = per_system[self.systems.index(system)]
                var = e_prop.variance[e_prop.systems.index(system)]
                m = hp[:, 0] / hp[:, 5]
                vals += (hp[:, 2] - 2.0 * m * hp[:, 1] + m * m * hp[:, 3]) / var
            if set_values:
                e_prop.value = float(vals[0])
            values += e_prop.weight * vals
        if f_prop is not None:
            vals = np.zeros(n_vec)
            for system in f_prop.systems:
                hp = per_system[self.systems.index(system)]
                vals += hp[:, 4] / (3 * system.n_atoms)
            if set_values:
                f_prop.value = float(vals[0])
            values += f_prop.weight * vals
        return values

    def _evaluate_fused_grad(self, slots_per_system, x):
        """Forward pass + ``pm_grad`` for ONE parameter vector; returns (value, gradient [P_opt])."""
        import torch
        e_prop, f_prop, r_prop = self._prop("ENERGY"), self._prop("FORCE"), self._prop("REGULARIZATION")
        X = x[None, :]
        work, partial_list, states = [], [], []
        for s_idx, (system, slots) in enumerate(zip(self.systems, slots_per_system)):
            st = self._system_state(s_idx, system)
            states.append(st)
            if not (st["need_e"] or st["need_f"]):
                partial_list.append(None)
                work.append(None)
                continue
            ens = st["ens"]
            slots_dev = torch.from_numpy(np.ascontiguousarray(slots, dtype=np.float64)).to(ens.device)
            derived = ens.topology.prepare(slots_dev)
            emm, fres, fmm = ens.evaluate(derived, with_forces=st["need_f"], want_fres=st["need_f"], want_fmm=st["need_f"])
            partial_list.append(ens.reduce(emm, fres, d_shift=st["d_shift"]).clone())
            work.append((derived, emm, fmm))
        live = [p for p in partial_list if p is not None]
        value = 0.0
        per_system = [None] * len(self.systems)
        if live:
            stacked = torch.stack(live, dim=0)
            _dist.allreduce_sum_tensor_(stacked)
            host_partials = stacked.cpu().numpy()
            k = 0
            for i, p in enumerate(partial_list):
                if p is not None:
                    per_system[i] = host_partials[k] * 1.0
                    k += 1
            raw = [None if hp is None else hp.copy() for hp in per_system]
            value = float(self.assemble_from_partials(per_system, [st["w_scale"] for st in states], True)[0])
        # per-conformation coefficients and the term-gradient pass
        slot_grads = [None] * len(self.systems)
        gouts = []
        for s_idx, (system, st) in enumerate(zip(self.systems, states)):
            if work[s_idx] is None:
                continue
            derived, emm, fmm = work[s_idx]
            ens = st["ens"]
            hp = raw[s_idx][0]
            scale = st["w_scale"]
            w = ens.weights if ens.weights is not None else None
            coef_e = coef_f = None
            if st["need_e"]:
                var = float(e_prop.variance[e_prop.systems.index(system)])
                n = hp[5]
                m = hp[0] / n
                sum_w, sum_wd = scale * hp[3], scale * hp[1]
                d = ens.eref - emm[0] - st["d_shift"]      # d' of pm_reduce_objective: (E_ref - E_mm) - shift, so dd'/dE_mm = -1
                wv = w * scale if w is not None else scale
                coef_e = ((2.0 * wv) * (d - m) + (2.0 / n) * (m * sum_w - sum_wd)) * (-e_prop.weight / var)
                coef_e = coef_e.contiguous()
            if st["need_f"]:
                c = f_prop.weight * scale / (3 * system.n_atoms)
                coef_f = (w * c).contiguous() if w is not None else torch.full((ens.n_conf,), c, dtype=torch.float64, device=ens.device)
            gouts.append((s_idx, ens.term_gradient(derived, coef_e, coef_f, None if fmm is None else fmm[0]).clone()))
        for s_idx, gout in gouts:
            _dist.allreduce_sum_tensor_(gout)
            topo = states[s_idx]["ens"].topology
            slot_grads[s_idx] = topo.term_gradient_to_slots(gout.cpu().numpy(), slots_per_system[s_idx][0])
        esp_prop = self._prop("ESP")
        if esp_prop is not None:
            vals = 0.0
            for system in esp_prop.systems:
                s_idx = self.systems.index(system)
                if np.ndim(system.weights) == 0 or len(np.atleast_1d(system.weights)) != system.n_structures:
                    system.compute_conformations_weights(temperature=self._weighting_temperature,
                                                         weighting_method=self._weighting_method)
                topo_n = system.n_atoms
                off = 2 * len(system.engine.system.bond_idx) + 2 * len(system.engine.system.angle_idx) + \
                    2 * len(system.engine.system.torsion_idx)
                charges = np.asarray(slots_per_system[s_idx])[0, off:off + 3 * topo_n:3]
                vals = float(system.resp_engine.esp_objective(charges, system.weights))
                last = (s_idx, off, topo_n, charges)
            esp_prop.value = vals
            value += esp_prop.weight * vals
            # as in the forward pass only the LAST system's ESP term survives (esp_property.py:104)
            s_idx, off, topo_n, charges = last
            system = self.systems[s_idx]
            if slot_grads[s_idx] is None:
                slot_grads[s_idx] = np.zeros(slots_per_system[s_idx].shape[1])
            slot_grads[s_idx][off:off + 3 * topo_n:3] += esp_prop.weight * system.resp_engine.esp_objective_gradient(charges, system.weights)
        store_grads = []
        for system, gs in zip(self.systems, slot_grads):
            store_grads.append(None if gs is None else
                               system.engine.slots_gradient_to_store(system.force_field, gs, system.force_field.store))
        grad = self.parameter_space.gradient_from_store(self.systems, x, store_grads)
        if r_prop is not None:
            reg = float(np.atleast_1d(r_prop.calculate_property(x))[0])
            value += r_prop.weight * reg
            grad = grad + r_prop.weight * r_prop.calculate_gradient(x)
        return value, grad

    # refresh the delta base with a full evaluation after this many committed moves (bounds the rounding drift)
    DELTA_COMMITS_BEFORE_REFRESH = 32

    def _forward_delta(self, st, ens, slots, derived):
        """
        ``delta_evaluation=True``: per-conformation energies / residuals of the rows of ``slots`` [P, n_slots].  A single row
        (the ``f`` path) is evaluated in full and becomes the BASE; a batch whose rows differ from the base in a few terms
        each (single-parameter moves of Monte Carlo / simulated annealing, finite-difference stencils) goes through
        ``pm_delta_eval``.  Slots that ALL rows changed in the same way are an accepted move of the caller: they are
        committed into the base first.  Anything else falls back to the full kernel.
        """
        import torch
        topo = ens.topology
        need_f = st["need_f"]

        def full(rows_derived):
            return ens.evaluate(rows_derived, with_forces=need_f, want_fres=need_f, want_fmm=need_f)

        def rebase(slots_row, derived_row):
            emm, fres, fmm = full(derived_row)
            ens.set_delta_base(slots_row, derived_row, emm[0], None if fres is None else fres[0], None if fmm is None else fmm[0])
            return emm, fres

        if slots.shape[0] == 1 or ens.delta_base is None:
            if slots.shape[0] == 1:
                return rebase(slots[0], derived)
            emm, fres, _ = full(derived)
            return emm, fres
        base = ens.delta_base
        diff = slots != base["slots"][None, :]
        common = diff.all(axis=0) & (slots == slots[0:1]).all(axis=0)
        if common.any():
            new_base = base["slots"].copy()
            new_base[common] = slots[0, common]
            new_derived = topo.prepare(torch.from_numpy(new_base[None, :].copy()).to(ens.device))
            plan = None if base["n_commits"] >= self.DELTA_COMMITS_BEFORE_REFRESH else topo.delta_plan(base["slots"], new_base[None, :])
            if plan is None:
                rebase(new_base, new_derived)
            else:
                ens.delta_commit(plan, new_base, new_derived)
            base = ens.delta_base
        plan = topo.delta_plan(base["slots"], slots)
        if plan is None:
            emm, fres, _ = full(derived)
            return emm, fres
        self.n_delta_rows += slots.shape[0]
        return ens.delta_evaluate(plan, derived)

    def _capture_graph(self, st, ens):
        """CUDA graph of one single-vector evaluation of ``ens``: pinned slots -> device, ``pm_prepare_params``,
        ``pm_eval_reduce`` (the E(+F) kernel with the fused residual reduction, then the row sum that is also the all-reduce over
        NVLink peer memory in multi-GPU runs), global sums -> pinned host.  The graph owns its buffers."""
        import torch
        topo = ens.topology
        g = dict(pin_in=torch.empty((1, topo.n_slots), dtype=torch.float64).pin_memory(),
                 slots_dev=torch.empty((1, topo.n_slots), dtype=torch.float64, device=ens.device),
                 derived=torch.empty((1, topo.n_derived), dtype=torch.float64, device=ens.device),
                 pin_out=torch.empty((1, NPARTIAL), dtype=torch.float64).pin_memory(), bufs={})
        saved_bufs, saved_scratch = ens._buffers, ens._scratch
        ens._buffers, ens._scratch = g["bufs"], None
        try:
            # allocate the graph's private buffers eagerly (same shapes as the capture will ask for), then capture.  The eager
            # call communicates too (every rank runs it: the sequence numbers of the communicator stay in step).
            comm = st.get("comm")
            g["global"] = comm is not None or _dist.world_size() == 1    # the sums the graph delivers are already global
            # zero-copy ends (no copy-engine nodes in the graph): pm_prepare_kernel reads the pinned slots row over PCIe and the
            # row-sum kernel writes its 8 sums straight into pinned host memory.  Only the NCCL fallback keeps the sums on the
            # device (they still have to be all-reduced).
            zero_copy = g["global"] and os.environ.get("PARAMOL_B200_ZERO_COPY", "1") != "0"
            g["slots_dev"].copy_(g["pin_in"])
            topo.prepare(g["slots_dev"], g["derived"])
            ens.evaluate_reduce(g["derived"], with_forces=st["need_f"], d_shift=st["d_shift"], comm=comm)
            g["scratch"] = ens._scratch
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                if zero_copy:
                    topo.prepare(g["pin_in"], g["derived"])
                    part, _, _ = ens.evaluate_reduce(g["derived"], with_forces=st["need_f"], d_shift=st["d_shift"], comm=comm,
                                                     partials_out=g["pin_out"])
                else:
                    g["slots_dev"].copy_(g["pin_in"], non_blocking=True)
                    topo.prepare(g["slots_dev"], g["derived"])
                    part, _, _ = ens.evaluate_reduce(g["derived"], with_forces=st["need_f"], d_shift=st["d_shift"], comm=comm)
                    g["pin_out"].copy_(part, non_blocking=True)
            g["graph"] = graph
            g["part_dev"] = part
            # raw handles for pm_graph_call (one C call per objective evaluation: stage, launch, wait, read back)
            g["exec"] = ctypes.c_void_p(graph.raw_cuda_graph_exec())
            g["dev_index"] = ens.device.index or 0
            g["pin_in_ptr"] = ctypes.c_void_p(g["pin_in"].data_ptr())
            g["pin_out_ptr"] = ctypes.c_void_p(g["pin_out"].data_ptr())
            g["out"] = np.zeros((1, NPARTIAL))
        finally:
            ens._buffers, ens._scratch = saved_bufs, saved_scratch
        return g

    def _evaluate_fused(self, slots_per_system, X, set_values, between=None):
        """
        slots_per_system: list over ``self.systems`` of float64 [P, n_slots]; X [P, P_opt] for the regularisation.
        Returns numpy [P].
        """
        import torch
        n_vec = X.shape[0]
        e_prop, f_prop, r_prop = self._prop("ENERGY"), self._prop("FORCE"), self._prop("REGULARIZATION")
        partial_list = []
        states = []
        use_graph = (n_vec == 1 and self._graphs and not self._delta and not self._cache_nonbonded
                     and self._weighting_method.upper() != "NON_BOLTZMANN")
        multi_rank = _dist.world_size() > 1
        graph_rows = {}
        global_parts = {}
        for s_idx, (system, slots) in enumerate(zip(self.systems, slots_per_system)):
            st = self._system_state(s_idx, system)
            states.append(st)
            if not (st["need_e"] or st["need_f"]):
                partial_list.append(None)
                continue
            ens = st["ens"]
            if use_graph:
                # single-vector calls (the optimiser's f(x)): upload, prepare, E+F kernel, reduction and download replayed as
                # ONE CUDA graph; the first calls run eagerly (they allocate and warm up what the capture then pins)
                st["eager_calls"] = st.get("eager_calls", 0) + 1
                if "graph" not in st and st["eager_calls"] > 2:
                    st["graph"] = self._capture_graph(st, ens)
                if "graph" in st:
                    g = st["graph"]
                    row = np.ascontiguousarray(slots, dtype=np.float64)
                    if not g["global"]:   # NCCL fallback: the partial sums stay on the device for the allreduce
                        g["pin_in"].copy_(torch.from_numpy(row))
                        g["graph"].replay()
                        partial_list.append(g["part_dev"])
                    else:
                        # one C call: slots -> pinned staging, graph launch on torch's current stream (ordered after whatever
                        # was queued there); the wait and the read-back follow the host-only terms below
                        g["stream_ptr"] = ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(g["dev_index"]))
                        _lib.check(_lib.lib().pm_graph_launch(g["exec"], g["stream_ptr"], row.ctypes.data, g["pin_in_ptr"], row.size),
                                   "pm_graph_launch")
                        graph_rows[s_idx] = g
                        partial_list.append("graph")
                    continue
            host = torch.from_numpy(np.ascontiguousarray(slots, dtype=np.float64))
            pin = st.get("pin")
            if pin is None or tuple(pin.shape) != tuple(host.shape):
                pin = torch.empty(tuple(host.shape), dtype=torch.float64).pin_memory()
                st["pin"] = pin
            pin.copy_(host)
            slots_dev = pin.to(ens.device, non_blocking=True)
            derived = ens.topology.prepare(slots_dev)
            nb_cache = None
            if self._cache_nonbonded:
                block = np.ascontiguousarray(slots[:, ens.topology.off_particle:])
                if np.all(block == block[0]):
                    key = block[0].tobytes()
                    if st.get("nb_key") != key:
                        st["nb_cache"] = ens.build_nonbonded_cache(derived[0:1])
                        st["nb_key"] = key
                    nb_cache = st["nb_cache"]
            if not self._delta and nb_cache is None and self._weighting_method.upper() != "NON_BOLTZMANN":
                # the common case: evaluation, residual reduction and (multi-GPU) the sum over ranks in one pass
                comm = st.get("comm")
                if comm is not None and n_vec > comm.max_vec:
                    comm = None
                if not st["need_f"] and n_vec >= self.ENERGY_BATCH_MIN_VECTORS and ens.energy_batch_supported():
                    # energy-only batch: the geometry of a conformation is evaluated once and reused by every vector (K2)
                    part, _ = ens.energy_batch_reduce(derived, d_shift=st["d_shift"], comm=comm)
                else:
                    part, _, _ = ens.evaluate_reduce(derived, with_forces=st["need_f"], d_shift=st["d_shift"], comm=comm)
                if comm is not None or not multi_rank:
                    global_parts[s_idx] = part
                    partial_list.append("global")
                else:
                    partial_list.append(part)
                continue
            if self._delta and nb_cache is None:
                emm, fres = self._forward_delta(st, ens, np.asarray(slots, dtype=np.float64), derived)
            else:
                emm, fres, _ = ens.evaluate(derived, with_forces=st["need_f"], want_fres=st["need_f"], nb_cache=nb_cache)
            if self._weighting_method.upper() == "NON_BOLTZMANN":
                w = self._non_boltzmann_weights(st, system, emm)
                rows = []
                for p in range(n_vec):
                    rows.append(ens.reduce(emm[p:p + 1], None if fres is None else fres[p:p + 1], d_shift=st["d_shift"],
                                           weights=w[p]).clone())
                partial_list.append(torch.cat(rows, dim=0))
            else:
                partial_list.append(ens.reduce(emm, fres, d_shift=st["d_shift"]))
        live = [p for p in partial_list if p is not None and not isinstance(p, str)]
        values = np.zeros(n_vec)
        # host-only work first: it overlaps the kernels that are still running
        while between:
            between.pop()()
        reg = None
        if r_prop is not None:
            if set_values:
                reg = np.atleast_1d(r_prop.calculate_property(X[0]))
            else:
                keep = r_prop.value
                reg = np.atleast_1d(r_prop.calculate_property(X))
                r_prop.value = keep
        if live or graph_rows or global_parts:
            host_partials = None
            if live:
                stacked = live[0].unsqueeze(0) if len(live) == 1 else torch.stack(live, dim=0)  # [n_live, P, 8]
                _dist.allreduce_sum_tensor_(stacked)         # the single data-path collective (NCCL / gloo fallback)
                host_partials = stacked.cpu().numpy()
            global_host = {i: t.cpu().numpy() for i, t in global_parts.items()}
            for g in graph_rows.values():   # one stream wait + read-back per replayed graph, on one GPU as on eight
                _lib.check(_lib.lib().pm_graph_wait(g["stream_ptr"], g["pin_out_ptr"], NPARTIAL, g["out"].ctypes.data), "pm_graph_wait")
            k = 0
            per_system = []
            for i, p in enumerate(partial_list):
                if p is None:
                    per_system.append(None)
                elif isinstance(p, str) and p == "global":
                    per_system.append(global_host[i] * 1.0)
                elif isinstance(p, str):
                    per_system.append(graph_rows[i]["out"].copy())
                else:
                    per_system.append(host_partials[k] * 1.0)
                    k += 1
            values = self.assemble_from_partials(per_system, [st["w_scale"] for st in states], set_values)
        esp_prop = self._prop("ESP")
        if esp_prop is not None:
            # quadratic form in the charges assembled once on the device (RESP.esp_objective); as in the reference the
            # property value is the LAST system's (esp_property.py:104 overwrites instead of summing)
            vals = np.zeros(n_vec)
            for system in esp_prop.systems:
                s_idx = self.systems.index(system)
                if np.ndim(system.weights) == 0 or len(np.atleast_1d(system.weights)) != system.n_structures:
                    system.compute_conformations_weights(temperature=self._weighting_temperature,
                                                         weighting_method=self._weighting_method)
                topo_n = system.n_atoms
                off = 2 * len(system.engine.system.bond_idx) + 2 * len(system.engine.system.angle_idx) + \
                    2 * len(system.engine.system.torsion_idx)
                charges = np.asarray(slots_per_system[s_idx])[:, off:off + 3 * topo_n:3]
                vals = np.atleast_1d(system.resp_engine.esp_objective(charges, system.weights))
            if set_values:
                esp_prop.value = float(vals[0])
            values = values + esp_prop.weight * vals
        if reg is not None:
            values = values + r_prop.weight * reg
        return values
