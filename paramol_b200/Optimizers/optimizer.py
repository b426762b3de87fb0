# -*- coding: utf-8 -*-
"""
:obj:`Optimizer` -- front end that instantiates one of the optimisers by name (``ParaMol/Optimizers/optimizer.py:13-129``).
"""


class Optimizer:
    optimizers = ["scipy", "monte_carlo", "simulated_annealing", "gradient_descent"]

    def __init__(self, method, settings, create_optimizer=True):
        self.settings = settings
        self.method_name = None
        self._optimizer = None
        if create_optimizer:
            self._create_optimizer(method, settings)

    def __str__(self):
        return "Instance of optimizer {}.".format(self.method_name)

    def run_optimization(self, f, parameters_values, constraints=None):
        return self._optimizer.run_optimization(f, parameters_values, constraints)

    def _create_optimizer(self, method, settings):
        assert method.lower() in self.optimizers
        if method.lower() == "monte_carlo":
            from .monte_carlo import MonteCarlo
            self._optimizer = MonteCarlo(**settings)
        elif method.lower() == "simulated_annealing":
            from .simulated_annealing import SimulatedAnnealing
            self._optimizer = SimulatedAnnealing(**settings)
        elif method.lower() == "scipy":
            from .scipy_optimizers import ScipyOptimizer
            self._optimizer = ScipyOptimizer(**settings)
        elif method.lower() == "gradient_descent":
            from .gradient_descent import GradientDescent
            self._optimizer = GradientDescent(**settings)
        self.method_name = method.lower()
        return self._optimizer
