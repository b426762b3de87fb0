# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`ForceProperty`: sum over conformations of w_s sum_j dF_sj^T M_j dF_sj / (3N), dF = F_mm - F_ref, with
M_j the inverse covariance of the QM forces on atom j ("components") or I / var_s(|F_ref,sj|) ("norm")
(reference ``ParaMol/Objective_function/Properties/force_property.py:58-202``).

On the device both forms are the same kernel epilogue with a per-atom 3x3 metric (:meth:`metric`).
"""
import logging

import numpy as np

from .property import Property
from ...Utils import dist as _dist


class ForceProperty(Property):
    """
    Parameters
    ----------
    systems : list of :obj:`paramol_b200.System.system.ParaMolSystem`
    term_type : str
        "components" or "norm".
    weight : float
    """

    def __init__(self, systems=[], term_type="components", weight=1.0):
        self.name = "FORCE"
        self.value = None
        self.weight = weight
        self.systems = systems
        self.units = "kilojoule_per_mole/nanometer"
        self.term_type = term_type
        self.variance = None
        self.inv_covariance = None

    def add_system(self, system):
        self.systems.append(system)
        return self.systems

    def calculate_property(self, fmm_data, term_type=None):
        if term_type is None:
            term_type = self.term_type
        obj_fun_forces = []
        if term_type.lower() == "components":
            for system, fmm, inv_cov in zip(self.systems, fmm_data, self.inv_covariance):
                diff = np.asarray(fmm) - system.ref_forces
                # sum_j (d^T C_j^-1) d, accumulated atom by atom like the reference loop (:91-95)
                force_error = np.zeros(diff.shape[0])
                for j in range(system.n_atoms):
                    force_error = force_error + np.einsum("sa,ab,sb->s", diff[:, j, :], inv_cov[j], diff[:, j, :])
                force_error = np.sum(system.weights * system.wham_weights * force_error) / (3 * system.n_atoms)
                obj_fun_forces.append(force_error)
        elif term_type.lower() == "norm":
            for system, fmm, var in zip(self.systems, fmm_data, self.variance):
                diff = np.asarray(fmm) - system.ref_forces
                obj_fun = np.sum(np.power(diff, 2), axis=2)
                obj_fun = obj_fun / var
                obj_fun = np.sum(obj_fun, 1)
                obj_fun = np.sum(system.weights * system.wham_weights * obj_fun)
                obj_fun = obj_fun / (3 * system.n_atoms)
                obj_fun_forces.append(obj_fun)
        else:
            raise NotImplementedError("Force property term of type {} is not implemented.".format(term_type))
        if _dist.world_size() > 1:
            obj_fun_forces = list(_dist.allreduce_sum_numpy(np.asarray(obj_fun_forces)))
        self.value = np.sum(obj_fun_forces)
        return obj_fun_forces

    def metric(self, system_index, term_type=None):
        """Per-atom 3x3 metric [N,9] the kernel epilogue applies for system ``system_index``."""
        if term_type is None:
            term_type = self.term_type
        if term_type.lower() == "components":
            return np.asarray(self.inv_covariance[system_index]).reshape(-1, 9)
        if term_type.lower() == "norm":
            var = np.asarray(self.variance[system_index])
            return (np.eye(3)[None, :, :] / var[:, None, None]).reshape(-1, 9)
        raise NotImplementedError("Force property term of type {} is not implemented.".format(term_type))

    def from_partials(self, partials_per_system):
        return [p[4] / (3 * system.n_atoms) for p, system in zip(partials_per_system, self.systems)]

    def calculate_inverse_covariance_qm_forces(self):
        """C_j^-1 with C_j = <F_ref,j F_ref,j^T> over conformations; identity when |det C_j| < 1e-8 (reference :126-173)."""
        self.inv_covariance = []
        for system in self.systems:
            assert system.ref_forces is not None, \
                "\t * Impossible to calculate the covariance of the QM forces since these were not set yet."
            ref = np.asarray(system.ref_forces, dtype=np.float64)
            outer = np.einsum("sja,sjb->jab", ref, ref)
            n = float(len(ref))
            if _dist.world_size() > 1:
                outer = _dist.allreduce_sum_numpy(outer)
                n = float(_dist.allreduce_sum_numpy(np.array([n]))[0])
            outer = outer / n
            inv_covariance_forces = np.zeros((system.n_atoms, 3, 3))
            for i in range(system.n_atoms):
                if abs(np.linalg.det(outer[i])) < 1e-8:
                    logging.info("Determinant is zero and therefore de covariance matrix is not invertible.")
                    inv_covariance_forces[i] = np.identity(3)
                else:
                    inv_covariance_forces[i] = np.linalg.inv(outer[i])
            self.inv_covariance.append(inv_covariance_forces)
        return self.inv_covariance

    def calculate_variance(self):
        """Variance over conformations of |F_ref,j| per atom, and the inverse covariances (reference :175-202)."""
        self.variance = []
        for system in self.systems:
            assert system.ref_forces is not None, "ERROR: It is not possible to calculate the variance, data was not set."
            qm_forces_norm = np.linalg.norm(np.asarray(system.ref_forces, dtype=np.float64), axis=2)
            if _dist.world_size() > 1:
                n = float(_dist.allreduce_sum_numpy(np.array([float(len(qm_forces_norm))]))[0])
                mean = _dist.allreduce_sum_numpy(np.sum(qm_forces_norm, axis=0)) / n
                var = _dist.allreduce_sum_numpy(np.sum((qm_forces_norm - mean[None, :]) ** 2, axis=0)) / n
            else:
                var = np.var(qm_forces_norm, axis=0)
            self.variance.append(var)
        self.calculate_inverse_covariance_qm_forces()
        return self.variance
