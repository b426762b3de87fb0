# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`MonteCarlo` -- the reference's greedy Monte Carlo optimiser (``ParaMol/Optimizers/monte_carlo.py:11-129``): single
parameter moves of adaptive amplitude, a move is kept when it lowers the objective.

Batched driver: within a sweep the proposal stream (which parameter, which displacement) does not depend on the
outcomes, so the next ``batch_size`` proposals are drawn ahead, applied one by one to the CURRENT accepted vector and
evaluated in one ``f_batch`` call; the decisions are then replayed in order up to and including the first acceptance,
after which the unconsumed proposals are re-applied to the new vector.  Random numbers are consumed exactly as in the
sequential algorithm, so the trajectory is the sequential one.
"""
import copy

import numpy as np

from ._batching import batch_evaluator


class MonteCarlo:
    """
    Parameters
    ----------
    n_blocks : int
        Sweeps between two adaptations of the maximum displacements.
    max_iter : int
    f_tol : float
    avg_acceptance_rate : float
        Target acceptance rate (``Settings.optimizer["monte_carlo"]`` calls it ``prob``; both names are accepted).
    batch_size : int
        Proposals evaluated per launch when the objective supports batches.
    """

    def __init__(self, n_blocks, max_iter, f_tol, avg_acceptance_rate=None, prob=None, batch_size=64):
        self._n_blocks = n_blocks
        self._max_iter = max_iter
        self._f_tol = f_tol
        self._avg_acceptance_rate = avg_acceptance_rate if avg_acceptance_rate is not None else prob
        assert self._avg_acceptance_rate is not None, "avg_acceptance_rate was not set."
        self._batch_size = max(1, int(batch_size))
        self.n_launches = 0

    def run_optimization(self, f, parameters, constraints=None):
        assert constraints is None, "Monte Carlo optimizer cannot handle restraints."
        print("!=================================================================================!")
        print("!                           STARTING MONTE CARLO OPTIMIZER                        !")
        print("!=================================================================================!")
        evaluate, batched = batch_evaluator(f)
        np.random.seed(np.random.randint(2 ** 32 - 1))
        n_param = len(parameters)
        acc = [0.0 for _ in range(n_param)]
        p_max = [copy.deepcopy(parameters[p]) * 0.5 if abs(parameters[p]) > 1e-3 else 1 for p in range(n_param)]
        old_f = f(parameters)
        block_counter = 0
        sweep = 0
        for sweep in range(1, self._max_iter):
            if block_counter == self._n_blocks:
                p_max = [p_max[p] * ((acc[p] / float(sweep) - self._avg_acceptance_rate) + 1) for p in range(n_param)]
                block_counter = 0
            block_counter += 1
            # the whole sweep's proposal stream, in the reference's draw order (randint, then uniform)
            draws = []
            for _ in range(n_param):
                p = np.random.randint(0, n_param)
                draws.append((p, np.random.uniform(-p_max[p], p_max[p])))
            n = 0
            while n < n_param:
                width = min(self._batch_size if batched else 1, n_param - n)
                trial = np.repeat(np.asarray(parameters, dtype=np.float64)[None, :], width, axis=0)
                for k in range(width):
                    p, step = draws[n + k]
                    trial[k, p] += step
                values = evaluate(trial)
                self.n_launches += 1
                consumed = width
                for k in range(width):
                    new_f = float(values[k])
                    if new_f < old_f:
                        p, step = draws[n + k]
                        parameters[p] += step
                        acc[p] += 1.0
                        print("MC move accepted. Objective function value: {}".format(new_f))
                        if abs(new_f - old_f) < self._f_tol:
                            print("\nFinal Acceptance rate: " + str(sum(acc) / ((sweep) * n_param)))
                            print("Convergence was achieved after {} MC sweeps.".format(sweep))
                            print("Last objective function value is {} .".format(new_f))
                            self._banner()
                            return parameters
                        old_f = new_f
                        consumed = k + 1  # later proposals of this batch were built on the old vector: redo them
                        break
                n += consumed
        print("Final Acceptance rate: " + str(sum(acc) / (max(sweep, 1) * n_param)))
        print("Maximum number of iterations was reached.")
        self._banner()
        return parameters

    @staticmethod
    def _banner():
        print("!=================================================================================!")
        print("!                MONTE CARLO OPTIMIZER TERMINATED SUCCESSFULLY! :)                !")
        print("!=================================================================================!")
