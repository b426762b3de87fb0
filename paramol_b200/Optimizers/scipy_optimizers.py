# -*- coding: utf-8 -*-
"""
SciPy wrapper (``ParaMol/Optimizers/scipy_optimizers.py:13-77``): ``scipy.optimize.minimize(fun=f, x0, **settings)``.

Additive: ``jac="analytic"`` in the minimizer settings hands scipy the closed-form gradient of
``ObjectiveFunction.f_and_grad`` (one extra device pass per evaluation) instead of letting it difference ``f``
(``jac="2-point"``, the reference default, ``ParaMol/Utils/settings.py:64``: P_opt + 1 evaluations per gradient).
"""


class ScipyOptimizer:
    def __init__(self, **minimizer_params):
        self.__dict__.update(**minimizer_params)

    def run_optimization(self, f, parameters, constraints=None):
        from scipy.optimize import minimize
        print("!=================================================================================!")
        print("!                           STARTING SCIPY OPTIMIZER                              !")
        print("!=================================================================================!")
        settings = dict(self.__dict__)
        fun = f
        if isinstance(settings.get("jac"), str) and settings["jac"].lower() == "analytic":
            owner = getattr(f, "__self__", None)
            if not callable(getattr(owner, "f_and_grad", None)):
                raise ValueError('jac="analytic" needs the f method of an objective function that offers f_and_grad')
            fun, settings["jac"] = owner.f_and_grad, True
        if constraints is None:
            optimization = minimize(fun=fun, x0=parameters, **settings)
        else:
            optimization = minimize(fun=fun, x0=parameters, constraints=constraints, **settings)
        print("!=================================================================================!")
        print("!                      SCIPY OPTIMIZER TERMINATED SUCCESSFULLY! :)                !")
        print("!=================================================================================!")
        return optimization.x
