# -*- coding: utf-8 -*-
"""
Callers of the objective function (mirrors ``ParaMol/Optimizers``): scipy wrapper, Monte Carlo, simulated annealing and
finite-difference gradient descent.  Each keeps the reference's algorithm and random-number stream; when the callable
they receive is ``ObjectiveFunction.f`` they evaluate trial vectors in batches through ``ObjectiveFunction.f_batch``
(P vectors per kernel launch) and replay the accept/reject decisions sequentially, so the result is the one the
sequential algorithm would produce.
"""
__all__ = ['optimizer', 'monte_carlo', 'simulated_annealing', 'gradient_descent', 'scipy_optimizers']
