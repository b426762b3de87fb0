# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`GradientDescent` -- the reference's steepest descent with finite-difference gradients
(``ParaMol/Optimizers/gradient_descent.py:11-151``).  The P_opt (1-point) or 2 P_opt (2-point) displaced vectors of a
gradient are evaluated in ONE ``f_batch`` call when the objective supports it -- the case the reference calls
"prohibitive" for >= 1000 conformations (``:54-57``).  Formulas, including the 2-point quotient (f+ - f-)/dx, are the
reference's.  Additive: ``derivative_type="analytic"`` takes the closed-form gradient of ``ObjectiveFunction.grad``.
"""
import copy

import numpy as np

from ._batching import batch_evaluator


class GradientDescent:
    def __init__(self, max_iter, derivative_calculation, derivative_type, g_tol, f_tol, dx, derivative_h):
        self._max_iter = max_iter
        self._derivative_calculation = derivative_calculation
        self._derivative_type = derivative_type
        self._g_tol = g_tol
        self._f_tol = f_tol
        self._dx = dx
        self._derivative_h = derivative_h
        self.n_launches = 0
        self._analytic = None

    def _gradient(self, evaluate, parameters, dp, f_new):
        n = len(parameters)
        if self._derivative_type == "analytic":
            if self._analytic is None:
                raise ValueError('derivative_type="analytic" needs an objective function that offers grad')
            self.n_launches += 1
            return np.asarray(self._analytic(parameters), dtype=np.float64)
        if self._derivative_type == "1-point":
            X = np.repeat(parameters[None, :], n, axis=0)
            X[np.arange(n), np.arange(n)] += dp
            f_plus = evaluate(X)
            self.n_launches += 1
            return (f_plus - f_new) / dp
        elif self._derivative_type == "2-point":
            X = np.repeat(parameters[None, :], 2 * n, axis=0)
            X[np.arange(n), np.arange(n)] += dp
            X[n + np.arange(n), np.arange(n)] -= dp
            values = evaluate(X)
            self.n_launches += 1
            return (values[:n] - values[n:]) / dp
        raise NotImplementedError("derivate_type {} is not implemented".format(self._derivative_type))

    def run_optimization(self, f, parameters, constraints=None):
        assert constraints is None, "Gradient Descent cannot handle restraints."
        print("!=================================================================================!")
        print("!                        STARTING GRADIENT DESCENT OPTIMIZER                      !")
        print("!=================================================================================!")
        evaluate, _ = batch_evaluator(f)
        self._analytic = getattr(getattr(f, "__self__", None), "grad", None)
        parameters = np.asarray(parameters, dtype=np.float64)
        grad = np.ones(len(parameters))
        dp = np.ones(len(parameters)) * self._dx
        alpha = np.ones(len(parameters)) * self._derivative_h
        param_min = copy.deepcopy(parameters)
        f_best = f(parameters)
        f_new = f_best
        f_old = 0.0
        iteration = 0
        while np.linalg.norm(grad) > self._g_tol or iteration < self._max_iter:
            iteration = iteration + 1
            parameters_dummy = copy.deepcopy(parameters)
            if self._derivative_calculation == "always" or (f_new > f_old and self._derivative_calculation == "f_increase"):
                grad = self._gradient(evaluate, parameters, dp, f_new)
            parameters = parameters - alpha * grad
            f_old = f_new
            f_new = f(parameters)
            if f_new < f_best:
                if (f_best - f_new) < self._f_tol:
                    break
                print("Iter: {} F: {:<15.8f} Gradient norm: {:<15.8f} deltaF: {:<15.8f}".format(
                    iteration, f_new, np.linalg.norm(grad), -(f_new - f_best)))
                param_min = copy.deepcopy(parameters)
                f_old = f_new
                f_best = f_new
            else:
                parameters = parameters_dummy
        print("!=================================================================================!")
        print("!                GRADIENT DESCENT OPTIMIZER TERMINATED SUCCESSFULLY! :)           !")
        print("!=================================================================================!")
        return param_min
