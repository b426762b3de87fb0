# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`ParameterSpace` maps the flat optimisation vector x onto the :obj:`Parameter` objects of every system and
pushes them into the engines; same public surface as the reference class
(``ParaMol/Parameter_space/parameter_space.py:25-399``).

``update_systems`` is the per-objective-call entry (reference ``:348-399``).  Here it is a handful of numpy
gathers: the scatter to tied parameters is one fancy-indexed assignment into each force field's value store and
the engine update is the compiled plan of :meth:`B200Engine.slots_from_store`.  ``update_systems_batch`` does the
same for P trial vectors at once WITHOUT touching the Parameter objects (Monte Carlo / simulated annealing /
finite-difference batches).
"""
import logging

import numpy as np


class ParameterSpace:
    """
    Parameters
    ----------
    parameters_magnitudes : dict
        Default magnitude per ``param_key`` (used for prior widths / scaling constants with method "default").
    prior_widths_method : str
        "default", "arithmetic" or "geometric".
    scaling_constants_method : str
        "default", "arithmetic" or "geometric".
    """
    symmetry_group_default = "X"

    def __init__(self, parameters_magnitudes={'charge': 0.5, 'lj_sigma': 0.30, 'lj_eps': 0.20,
                                              'torsion_phase': np.pi, 'torsion_k': 4 * 4.184,
                                              'bond_eq': 0.05, 'bond_k': 100000, 'angle_eq': np.pi / 16.0,
                                              'angle_k': 100.0, 'scee': 1.0, 'scnb': 1.0},
                 prior_widths_method="default", scaling_constants_method="default"):
        self.optimizable_parameters_values_scaled = None
        self.optimizable_parameters_values = None
        self.optimizable_parameters = None
        self.optimizable_parameters_by_system = None
        self.optimizable_parameters_values_by_system = None
        self.optimizable_parameters_by_symmetry = None
        self.parameters_magnitudes = parameters_magnitudes
        self.prior_widths = None
        self.prior_widths_dict = None
        self.scaling_constants = None
        self.scaling_constants_dict = None
        self.initial_optimizable_parameters_values = None
        self.initial_optimizable_parameters_values_scaled = None
        self.scaling_constants_method = scaling_constants_method
        self.prior_widths_method = prior_widths_method
        self.preconditioned = False
        self._scatter = None

    # ------------------------------------------------------------------------------------------
    def _print_table(self, title, values):
        print("!=================================================================================!")
        print("!{:^81s}!".format(title))
        print("! {:<30s}{:<30s}{:<19s} !".format("Term type", "Value", " "))
        print("!---------------------------------------------------------------------------------!")
        for term_type in values:
            print('! {:<30s}{:<30.8f}{:<19s} !'.format(term_type, values[term_type], " "))
        print("!=================================================================================!")

    def calculate_scaling_constants(self, method=None):
        if method is None:
            method = self.scaling_constants_method
        self.scaling_constants_dict, self.scaling_constants = self.calculate_parameters_magnitudes(method=method)
        self._print_table("Scaling Constants", self.scaling_constants_dict)
        return self.scaling_constants_dict, self.scaling_constants

    def calculate_prior_widths(self, method=None):
        if method is None:
            method = self.prior_widths_method
        logging.info("Calculating prior widths using the method {}".format(method))
        self.prior_widths_dict, self.prior_widths = self.calculate_parameters_magnitudes(method=method)
        self._print_table("Prior Widths", self.prior_widths_dict)
        return self.prior_widths_dict, self.prior_widths

    def calculate_parameters_magnitudes(self, method=None):
        """
        Magnitude per ``param_key`` over the optimizable parameters (reference ``:174-246``): the default table, the
        arithmetic mean of |v|, or the geometric mean of the |v| > 1e-8 (running product, as the reference computes it).
        """
        assert method is not None, "No method was chosen to calculate the parameters' magnitudes."
        param_keys = {}
        for parameter in self.optimizable_parameters:
            param_keys.setdefault(parameter.param_key, []).append(parameter.value)
        magnitudes = {}
        m = method.lower()
        if m == "geometric":
            for param_key, values in param_keys.items():
                geometric_mean = 1.0
                n = 0.0
                for value in values:
                    if abs(value) > 1e-8:
                        geometric_mean = geometric_mean * np.abs(value)
                        n = n + 1
                if abs(geometric_mean) > 1e-8 and n > 0:
                    magnitudes[param_key] = geometric_mean ** (1.0 / n)
                else:
                    magnitudes[param_key] = self.parameters_magnitudes[param_key]
        elif m == "arithmetic":
            for param_key, values in param_keys.items():
                arithmetic_mean = 0.0
                n = 0.0
                for value in values:
                    arithmetic_mean = arithmetic_mean + np.abs(value)
                    n = n + 1
                if abs(arithmetic_mean) > 1e-8 and n > 0:
                    magnitudes[param_key] = arithmetic_mean / n
                else:
                    magnitudes[param_key] = self.parameters_magnitudes[param_key]
        elif m == "default":
            for param_key in param_keys:
                magnitudes[param_key] = self.parameters_magnitudes[param_key]
        else:
            raise NotImplementedError("\t * Mean type {} not available to guess the prior widths.".format(method))
        per_parameter = np.asarray([magnitudes[parameter.param_key] for parameter in self.optimizable_parameters])
        return magnitudes, per_parameter

    def jacobi_preconditioning(self):
        """x_scaled = x / scaling_constants (reference ``:248-269``)."""
        assert self.scaling_constants is not None
        assert self.optimizable_parameters_values is not None
        self.optimizable_parameters_values_scaled = self.optimizable_parameters_values / self.scaling_constants
        self.preconditioned = True
        return self.optimizable_parameters_values_scaled

    def get_optimizable_parameters(self, systems, symmetry_constrained=True):
        """
        Builds the optimisation vector over all systems (reference ``:276-346``).  With symmetry constraints the
        first parameter met for a (symmetry_group, param_key) represents the group ACROSS systems; the others are
        recorded in ``optimizable_parameters_by_symmetry`` and follow it in ``update_systems``.
        """
        self.optimizable_parameters_by_system = []
        self.optimizable_parameters_values_by_system = []
        self.optimizable_parameters_by_symmetry = []
        self.optimizable_parameters = []
        self.optimizable_parameters_values = []
        for system in systems:
            parameters, values = system.force_field.get_optimizable_parameters(symmetry_constrained=symmetry_constrained)
            self.optimizable_parameters_by_system.append(parameters)
            self.optimizable_parameters_values_by_system.append(values)
        symm_groups = {}
        for system_parameters in self.optimizable_parameters_by_system:
            for parameter in system_parameters:
                if symmetry_constrained and parameter.symmetry_group != self.symmetry_group_default:
                    group = symm_groups.setdefault(parameter.symmetry_group, {})
                    if parameter.param_key in group:
                        self.optimizable_parameters_by_symmetry[group[parameter.param_key]].append(parameter)
                        continue
                    group[parameter.param_key] = len(self.optimizable_parameters_by_symmetry)
                self.optimizable_parameters.append(parameter)
                self.optimizable_parameters_values.append(parameter.value)
                self.optimizable_parameters_by_symmetry.append([parameter])
        self._scatter = None
        return self.optimizable_parameters, self.optimizable_parameters_values

    # ------------------------------------------------------------------------------------------
    # update_systems: object path semantics, array implementation
    # ------------------------------------------------------------------------------------------
    def _compile_scatter(self, systems, symmetry_constrained):
        """
        Per system: (store indices, x indices) such that ``store[dst] = x_phys[src]`` reproduces the two scatters of the
        reference -- over ``optimizable_parameters_by_symmetry`` (:378-380) and then, inside
        ``ForceField.update_force_field`` (:102-155), over the system's own symmetry groups.
        """
        key = (tuple(id(s) for s in systems), tuple(s.force_field.version for s in systems),
               tuple(id(s.force_field.store) for s in systems), id(self.optimizable_parameters_by_symmetry),
               bool(symmetry_constrained))
        if self._scatter is not None and self._scatter["key"] == key:
            return self._scatter
        owner = {}
        for s_idx, system in enumerate(systems):
            owner[id(system.force_field.store)] = s_idx
        per_system = [dict(dst=[], src=[]) for _ in systems]
        x_index_of = {}
        for i, group in enumerate(self.optimizable_parameters_by_symmetry):
            for parameter in group:
                x_index_of[id(parameter)] = i
                s_idx = owner.get(id(parameter._store))
                if s_idx is None:
                    raise ValueError("A parameter of the parameter space does not belong to any of the systems.")
                per_system[s_idx]["dst"].append(parameter._idx)
                per_system[s_idx]["src"].append(i)
        for s_idx, system in enumerate(systems):
            ff = system.force_field
            if symmetry_constrained:
                # update_force_field: every optimizable parameter of a non-default symmetry group takes the value of
                # the group's first entry in the system's optimizable_parameters list
                symm_groups = {}
                for parameter in ff.optimizable_parameters:
                    if parameter.symmetry_group != ff.symmetry_group_default:
                        group = symm_groups.setdefault(parameter.symmetry_group, {})
                        if parameter.param_key not in group:
                            group[parameter.param_key] = x_index_of[id(parameter)]
                for force in ff.force_field_optimizable:
                    for sub_force in ff.force_field_optimizable[force]:
                        for term in sub_force:
                            for parameter in term.parameters.values():
                                if parameter.optimize and parameter.symmetry_group != ff.symmetry_group_default:
                                    per_system[s_idx]["dst"].append(parameter._idx)
                                    per_system[s_idx]["src"].append(symm_groups[parameter.symmetry_group][parameter.param_key])
            per_system[s_idx]["dst"] = np.asarray(per_system[s_idx]["dst"], dtype=np.int64)
            per_system[s_idx]["src"] = np.asarray(per_system[s_idx]["src"], dtype=np.int64)
        by_system_idx = [np.asarray([parameter._idx for parameter in parameters], dtype=np.int64)
                         for parameters in self.optimizable_parameters_by_system]
        self._scatter = dict(key=key, per_system=per_system, by_system_idx=by_system_idx)
        return self._scatter

    def _physical(self, parameters_values):
        x = np.asarray(parameters_values, dtype=np.float64)
        if self.preconditioned:
            return x * self.scaling_constants
        return x

    def update_systems(self, systems, parameters_values, symmetry_constrained=True, defer=None):
        """
        Pushes ``parameters_values`` (scaled if ``preconditioned``) into every Parameter, force field and engine
        (reference ``:348-399``).  Returns the physical parameter vector.

        ``defer``: a list, or None.  With a list, the part of the update nobody needs before the evaluation starts (the
        per-system value lists, the MM-system parameter arrays, the RESP engine's charges) is appended to it as ONE callable
        instead of being executed: ``ObjectiveFunction.f`` runs it while the device evaluates.  ``last_slots`` holds the slot row
        of every system either way.
        """
        if self.preconditioned:
            self.optimizable_parameters_values_scaled = parameters_values
            self.optimizable_parameters_values = parameters_values * self.scaling_constants
        else:
            self.optimizable_parameters_values = parameters_values
        x = np.asarray(self.optimizable_parameters_values, dtype=np.float64)
        plan = self._compile_scatter(systems, symmetry_constrained)
        pending = []
        last_slots = []
        for system, sc, idx in zip(systems, plan["per_system"], plan["by_system_idx"]):
            ff = system.force_field
            if len(sc["dst"]):
                ff.store[sc["dst"]] = x[sc["src"]]
            # per-system value list as the reference builds it right after the scatter (parameter_space.py:383-385),
            # gathered from the value store instead of reading 232 Python properties per call
            values = ff.store[idx]
            engine_plan = system.engine.compile_update_plan(ff)
            if len(engine_plan["abs_idx"]):
                ff.store[engine_plan["abs_idx"]] = np.abs(ff.store[engine_plan["abs_idx"]])
            slots = system.engine.slots_from_store(ff, ff.store[None, :])
            last_slots.append(slots)
            pending.append((system, values, slots))
        self.last_slots = last_slots

        def finish():
            by_system = []
            system = None
            for system, values, slots in pending:
                by_system.append(values.tolist())
                system.engine.load_slots(slots[0])
            self.optimizable_parameters_values_by_system = by_system
            # as in the reference, only the LAST system's RESP engine is refreshed (:394-395 uses the loop variable)
            if system is not None and system.resp_engine is not None:
                if hasattr(system.resp_engine, "set_charges"):
                    system.resp_engine.set_charges(system.force_field.force_field)
                else:
                    raise NotImplementedError("MM RESP Engine {} is not implemented.".format(type(system.resp_engine)))

        if defer is None:
            finish()
        else:
            defer.append(finish)
        return self.optimizable_parameters_values

    def gradient_from_store(self, systems, parameters_values, store_gradients, symmetry_constrained=True):
        """
        Adjoint of the scatter of :meth:`update_systems`: ``store_gradients[s]`` = d f / d (value store of system s, after
        the abs() sign conventions) -> d f / d parameters_values (scaled coordinates if ``preconditioned``).  Where the
        scatter writes a store entry twice the last write wins, as in the forward pass; abs() contributes sign(x).
        """
        x = self._physical(parameters_values)
        plan = self._compile_scatter(systems, symmetry_constrained)
        g = np.zeros(len(x))
        for system, sc, gs in zip(systems, plan["per_system"], store_gradients):
            if gs is None or not len(sc["dst"]):
                continue
            gs = np.asarray(gs, dtype=np.float64).copy()
            engine_plan = system.engine.compile_update_plan(system.force_field)
            last = {}
            for d, src in zip(sc["dst"].tolist(), sc["src"].tolist()):
                last[d] = src
            dst = np.fromiter(last.keys(), dtype=np.int64, count=len(last))
            src = np.fromiter(last.values(), dtype=np.int64, count=len(last))
            sign = np.ones(len(gs))
            if len(engine_plan["abs_idx"]):
                pre = np.zeros(len(gs))
                pre[dst] = x[src]
                idx = engine_plan["abs_idx"]
                in_scatter = np.zeros(len(gs), dtype=bool)
                in_scatter[dst] = True
                sign[idx] = np.where(in_scatter[idx], np.sign(pre[idx]), 1.0)
            np.add.at(g, src, gs[dst] * sign[dst])
        if self.preconditioned:
            g = g * self.scaling_constants
        return g

    def update_systems_batch(self, systems, parameters_matrix, symmetry_constrained=True):
        """
        Slot rows for P trial vectors: list over systems of float64 [P, n_slots].  Parameter objects, force fields
        and engines are left untouched (trial vectors are hypothetical until one is accepted through
        ``update_systems``).
        """
        X = np.atleast_2d(np.asarray(parameters_matrix, dtype=np.float64))
        if self.preconditioned:
            X = X * self.scaling_constants[None, :]
        plan = self._compile_scatter(systems, symmetry_constrained)
        out = []
        for system, sc in zip(systems, plan["per_system"]):
            ff = system.force_field
            rows = np.repeat(ff.store[None, :], X.shape[0], axis=0)
            if len(sc["dst"]):
                rows[:, sc["dst"]] = X[:, sc["src"]]
            engine_plan = system.engine.compile_update_plan(ff)
            if len(engine_plan["abs_idx"]):
                rows[:, engine_plan["abs_idx"]] = np.abs(rows[:, engine_plan["abs_idx"]])
            out.append(system.engine.slots_from_store(ff, rows))
        return out
