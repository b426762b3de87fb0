# -*- coding: utf-8 -*-
"""SciPy wrapper (``ParaMol/Optimizers/scipy_optimizers.py:13-77``): ``scipy.optimize.minimize(fun=f, x0, **settings)``."""


class ScipyOptimizer:
    def __init__(self, **minimizer_params):
        self.__dict__.update(**minimizer_params)

    def run_optimization(self, f, parameters, constraints=None):
        from scipy.optimize import minimize
        print("!=================================================================================!")
        print("!                           STARTING SCIPY OPTIMIZER                              !")
        print("!=================================================================================!")
        if constraints is None:
            optimization = minimize(fun=f, x0=parameters, **self.__dict__)
        else:
            optimization = minimize(fun=f, x0=parameters, constraints=constraints, **self.__dict__)
        print("!=================================================================================!")
        print("!                      SCIPY OPTIMIZER TERMINATED SUCCESSFULLY! :)                !")
        print("!=================================================================================!")
        return optimization.x
