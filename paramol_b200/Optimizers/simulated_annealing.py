# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`SimulatedAnnealing` -- Metropolis annealing over single-parameter moves with the reference's schedule
(``ParaMol/Optimizers/simulated_annealing.py:11-142``): geometric cooling from T(p_init) to T(p_final), 99 sweeps per
temperature, displacement amplitudes adapted every sweep.

The reference class cannot run as shipped (it unpacks ``error, old_f = f(parameters)`` while ``ObjectiveFunction.f``
returns a float, ``:86``, and ``Optimizer`` looks it up under a misspelt name, ``optimizer.py:111``); this class is the
same algorithm with ``f`` returning the objective value.

Batched driver: proposals are drawn ahead under the assumption "rejected after a Metropolis draw"; the random-number
state is saved before every proposal, so at the first proposal whose outcome breaks the assumption (a downhill move draws
no Metropolis number; an accepted move changes the vector) the state is rewound and the stream continues exactly as the
sequential algorithm would.
"""
import copy

import numpy as np

from ._batching import batch_evaluator


class SimulatedAnnealing:
    def __init__(self, n_iter, p_init, p_final, avg_acceptance_rate, n_sweeps=99, batch_size=32):
        self._n_iter = n_iter
        self._p_init = p_init
        self._p_final = p_final
        self._avg_acceptance_rate = avg_acceptance_rate
        self._n_sweeps = n_sweeps
        self._batch_size = max(1, int(batch_size))
        self.n_launches = 0

    def run_optimization(self, f, parameters, constraints=None):
        assert constraints is None, "Simulated Annealing optimizer cannot handle restraints."
        print("!=================================================================================!")
        print("!                      STARTING SIMULATED ANNEALING OPTIMIZER                     !")
        print("!=================================================================================!")
        evaluate, batched = batch_evaluator(f)
        t_init = - 1.0 / np.log(self._p_init)
        t_final = - 1.0 / np.log(self._p_final)
        frac = (t_final / t_init) ** (1.0 / (self._n_iter - 1.0)) if self._n_iter > 1 else 1.0
        temp = t_init
        np.random.seed(np.random.randint(2 ** 32 - 1))
        n_param = len(parameters)
        old_f = f(parameters)
        best_f = old_f
        best_parameters = copy.deepcopy(parameters)
        new_f = old_f
        acc = [0.0] * n_param
        sweep = 1
        for _ntemp in range(self._n_iter):
            acc = [0.0 for _ in range(n_param)]
            p_max = [copy.deepcopy(parameters[p]) * 0.5 for p in range(n_param)]
            for sweep in range(1, self._n_sweeps + 1):
                p_max = [p_max[p] * ((acc[p] / float(sweep) - self._avg_acceptance_rate) + 1) for p in range(n_param)]
                n = 0
                while n < n_param:
                    width = min(self._batch_size if batched else 1, n_param - n)
                    states, moves = [], []
                    trial = np.repeat(np.asarray(parameters, dtype=np.float64)[None, :], width, axis=0)
                    for k in range(width):
                        states.append(np.random.get_state())
                        p = np.random.randint(0, n_param)
                        step = np.random.uniform(-p_max[p], p_max[p])
                        np.random.random()  # speculative Metropolis draw
                        moves.append((p, step))
                        trial[k, p] += step
                    values = evaluate(trial)
                    self.n_launches += 1
                    consumed = width
                    for k in range(width):
                        p, step = moves[k]
                        new_f = float(values[k])
                        if new_f - old_f < 0:
                            # downhill: accepted, and the sequential algorithm draws no Metropolis number here
                            np.random.set_state(states[k])
                            np.random.randint(0, n_param)
                            np.random.uniform(-p_max[p], p_max[p])
                            parameters[p] += step
                            if new_f < best_f:
                                best_f = new_f
                                best_parameters = copy.deepcopy(parameters)
                            old_f = new_f
                            acc[p] += 1.0
                            consumed = k + 1
                            break
                        np.random.set_state(states[k])
                        np.random.randint(0, n_param)
                        np.random.uniform(-p_max[p], p_max[p])
                        prob = np.exp(- (new_f - old_f) / temp)
                        if prob > np.random.random():
                            parameters[p] += step
                            old_f = new_f
                            acc[p] += 1.0
                            consumed = k + 1
                            break
                    # if every proposal was rejected the state is the speculative one; otherwise it was rewound to just
                    # after proposal `consumed - 1`
                    n += consumed
            temp = temp * frac
        print("Acceptance rate: " + str(sum(acc) / ((sweep) * n_param)))
        print("Last objective function value is {} .".format(new_f))
        print("!=================================================================================!")
        print("!               SIMULATED ANNEALING OPTIMIZER TERMINATED SUCCESSFULLY! :)         !")
        print("!=================================================================================!")
        return best_parameters
