# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`EnergyProperty`: sum over conformations of w_s (d_s - mean(d))^2 / var(E_ref) with d = E_ref - E_mm
(reference ``ParaMol/Objective_function/Properties/energy_property.py:72-134``).

``calculate_property(emm_data)`` is the reference's numpy formula (used on the S3 seam and by the tests);
``from_partials`` assembles the same value from the partial sums the reduction kernel produced
(``pm_reduce_objective``, include/paramol_b200.h) after the allreduce.
"""
import numpy as np

from .property import Property
from ...Utils import dist as _dist


class EnergyProperty(Property):
    """
    Parameters
    ----------
    systems : list of :obj:`paramol_b200.System.system.ParaMolSystem`
    weight : float
    """

    def __init__(self, systems=[], weight=1.0):
        self.name = 'ENERGY'
        self.value = None
        self.weight = weight
        self.systems = systems
        self.variance = None
        self.units = "kilojoule_per_mole"
        self.mean_ref = None

    def add_system(self, system):
        self.systems.append(system)
        return self.systems

    def calculate_property(self, emm_data):
        obj_fun_energies = []
        # (the reference converts emm_data with np.asarray, which only works when every system has the same number of
        # conformations; the per-system arrays are used as they are here)
        for system, emm, var in zip(self.systems, emm_data, self.variance):
            diff = np.asarray(system.ref_energies) - np.asarray(emm)
            n_global = system.n_structures_global()
            avg_diff = float(_dist.allreduce_sum_numpy(np.array([np.sum(diff)]))[0]) / n_global \
                if _dist.world_size() > 1 else np.mean(diff)
            obj_fun = diff - avg_diff
            obj_fun = obj_fun * obj_fun
            obj_fun = np.sum(system.weights * system.wham_weights * obj_fun)
            if _dist.world_size() > 1:
                obj_fun = float(_dist.allreduce_sum_numpy(np.array([obj_fun]))[0])
            obj_fun_energies.append(obj_fun / var)
        self.value = np.sum(obj_fun_energies)
        return obj_fun_energies

    def from_partials(self, partials_per_system):
        """``partials_per_system``: list over systems of the (allreduced) 8-vector of one parameter vector."""
        out = []
        for p, var in zip(partials_per_system, self.variance):
            m = p[0] / p[5]
            out.append((p[2] - 2.0 * m * p[1] + m * m * p[3]) / var)
        return out

    def calculate_variance(self):
        """Population variance of the reference energies of each system (global over ranks)."""
        self.variance = []
        self.mean_ref = []
        for system in self.systems:
            assert system.ref_energies is not None, "ERROR: It is not possible to calculate the variance, data was not set."
            e = np.asarray(system.ref_energies, dtype=np.float64)
            if _dist.world_size() > 1:
                n, s1 = _dist.allreduce_sum_numpy(np.array([float(len(e)), np.sum(e)]))
                mean = s1 / n
                var = float(_dist.allreduce_sum_numpy(np.array([np.sum((e - mean) ** 2)]))[0]) / n
            else:
                mean = float(np.mean(e))
                var = np.var(e)
            self.mean_ref.append(mean)
            self.variance.append(var)
        self.variance = np.asarray(self.variance)
        return self.variance
