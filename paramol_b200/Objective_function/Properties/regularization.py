# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`Regularization`: L2 / L1 / hyperbolic penalty on the optimisation vector (reference
``ParaMol/Objective_function/Properties/regularization.py:19-228``).  Host arithmetic on P_opt doubles; it is
added once after the allreduce, never per rank.  ``calculate_property`` also accepts a [P, P_opt] matrix.
"""
import numpy as np

from .property import Property


class Regularization(Property):
    """
    Parameters
    ----------
    initial_parameters_values : array
    prior_widths : array
    method : str
        "L2", "L1" or "hyperbolic".
    weight : float
    scaling_factor : float
        alpha.
    hyperbolic_beta : float
    """

    def __init__(self, initial_parameters_values, prior_widths, method, weight=1.0, scaling_factor=1.0, hyperbolic_beta=0.01):
        self.name = "REGULARIZATION"
        self.systems = None
        self._regularization_type = method
        self._scaling_factor = scaling_factor
        self._hyperbolic_beta = hyperbolic_beta
        self._initial_parameters_values = initial_parameters_values
        self._prior_widths = prior_widths
        self.units = 'ADIMENSIONAL'
        self.value = None
        self.weight = weight

    def set_initial_parameters_values(self, initial_parameters_values):
        self._initial_parameters_values = initial_parameters_values
        return self._initial_parameters_values

    def set_prior_widths(self, prior_widths):
        self._prior_widths = prior_widths
        return self._prior_widths

    def calculate_property(self, current_parameters, a=None, b=None):
        kind = self._regularization_type.upper()
        if kind == "L2":
            return self._l2_regularization(current_parameters, a)
        elif kind == "L1":
            return self._l1_regularization(current_parameters, a)
        elif kind == "HYPERBOLIC":
            return self._hyperbolic_regularization(current_parameters, a, b)
        raise NotImplementedError("Regularization {} scheme not implement.".format(self._regularization_type))

    def _store(self, value):
        self.value = value if np.ndim(value) else float(value)
        return self.value

    def _l2_regularization(self, current_parameters, a=None):
        if a is None:
            a = self._scaling_factor
        diff = (np.asarray(current_parameters) - self._initial_parameters_values) / self._prior_widths
        return self._store(a * np.sum(np.power(diff, 2), axis=-1))

    def _l1_regularization(self, current_parameters, a=None):
        if a is None:
            a = self._scaling_factor
        diff = (np.asarray(current_parameters) - self._initial_parameters_values) / self._prior_widths
        return self._store(a * np.sum(np.abs(diff), axis=-1))

    def _hyperbolic_regularization(self, current_parameters, a=None, b=None):
        if a is None:
            a = self._scaling_factor
        if b is None:
            b = self._hyperbolic_beta
        reg = np.sum(((np.asarray(current_parameters)) ** 2 + b ** 2) ** (1 / 2.) - b, axis=-1)
        return self._store(a * reg)

    def calculate_gradient(self, current_parameters, a=None, b=None):
        """d(penalty)/d(current_parameters), numpy [P_opt] (no reference counterpart: the reference differentiates numerically)."""
        if a is None:
            a = self._scaling_factor
        x = np.asarray(current_parameters, dtype=np.float64)
        kind = self._regularization_type.upper()
        if kind == "L2":
            return 2.0 * a * (x - self._initial_parameters_values) / np.power(self._prior_widths, 2)
        elif kind == "L1":
            return a * np.sign(x - self._initial_parameters_values) / np.asarray(self._prior_widths)
        elif kind == "HYPERBOLIC":
            if b is None:
                b = self._hyperbolic_beta
            return a * x / np.sqrt(x ** 2 + b ** 2)
        raise NotImplementedError("Regularization {} scheme not implement.".format(self._regularization_type))
