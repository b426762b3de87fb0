"""
Batched optimiser drivers (CPU): with an objective that offers f_batch the Monte Carlo, simulated-annealing and
finite-difference gradient-descent drivers must follow exactly the trajectory of the sequential algorithm (same random
stream, same decisions), using fewer launches.
"""
import numpy as np

from paramol_b200.Optimizers.optimizer import Optimizer
from paramol_b200.Optimizers.monte_carlo import MonteCarlo
from paramol_b200.Optimizers.simulated_annealing import SimulatedAnnealing
from paramol_b200.Optimizers.gradient_descent import GradientDescent


class Objective:
    """Stand-in for ObjectiveFunction: f(x) and f_batch(X) of a smooth non-convex function."""

    def __init__(self):
        self.calls = 0
        self.batch_calls = 0
        self.rows = 0
        self.c = np.linspace(0.5, 2.0, 6)

    def f(self, x):
        self.calls += 1
        x = np.asarray(x, dtype=np.float64)
        return float(np.sum(self.c * (x - 1.0) ** 2) + 0.3 * np.sum(np.sin(3.0 * x)))

    def f_batch(self, X):
        self.batch_calls += 1
        self.rows += len(X)
        X = np.asarray(X, dtype=np.float64)
        return np.sum(self.c[None, :] * (X - 1.0) ** 2, axis=1) + 0.3 * np.sum(np.sin(3.0 * X), axis=1)


def _sequential(obj):
    return lambda x: obj.f(x)  # a plain function: no f_batch is discoverable -> row-by-row evaluation


def test_monte_carlo_batched_equals_sequential():
    x0 = [0.2, -0.4, 1.7, 0.9, 2.5, -1.0]
    a, b = Objective(), Objective()
    np.random.seed(123)
    xs = MonteCarlo(n_blocks=5, max_iter=40, f_tol=1e-12, avg_acceptance_rate=0.25).run_optimization(_sequential(a), list(x0))
    np.random.seed(123)
    mc = MonteCarlo(n_blocks=5, max_iter=40, f_tol=1e-12, prob=0.25, batch_size=4)
    xb = mc.run_optimization(b.f, list(x0))
    assert np.array_equal(np.asarray(xs), np.asarray(xb))
    assert b.batch_calls > 0 and b.batch_calls < a.calls - 1
    assert a.f(xs) < a.f(x0)


def test_simulated_annealing_batched_equals_sequential():
    x0 = [0.2, -0.4, 1.7, 0.9, 2.5, -1.0]
    a, b = Objective(), Objective()
    kw = dict(n_iter=4, p_init=0.2, p_final=0.001, avg_acceptance_rate=0.25, n_sweeps=6)
    np.random.seed(7)
    xs = SimulatedAnnealing(**kw).run_optimization(_sequential(a), list(x0))
    np.random.seed(7)
    xb = SimulatedAnnealing(batch_size=5, **kw).run_optimization(b.f, list(x0))
    assert np.array_equal(np.asarray(xs), np.asarray(xb))
    assert b.batch_calls > 0 and b.batch_calls < a.calls


def test_gradient_descent_batches_the_finite_differences():
    x0 = np.array([0.2, -0.4, 1.7, 0.9, 2.5, -1.0])
    for kind in ("1-point", "2-point"):
        a, b = Objective(), Objective()
        kw = dict(max_iter=25, derivative_calculation="always", derivative_type=kind, g_tol=1e3, f_tol=1e-14, dx=1e-4,
                  derivative_h=5e-2)
        xs = GradientDescent(**kw).run_optimization(_sequential(a), x0.copy())
        gd = GradientDescent(**kw)
        xb = gd.run_optimization(b.f, x0.copy())
        assert np.allclose(xs, xb, rtol=0, atol=1e-12)
        assert b.batch_calls == gd.n_launches and b.rows == gd.n_launches * (6 if kind == "1-point" else 12)
        assert a.f(xs) < a.f(x0)


def test_optimizer_front_end_and_scipy():
    obj = Objective()
    opt = Optimizer("scipy", {"method": "SLSQP", "options": {"maxiter": 200, "ftol": 1e-12}})
    x = opt.run_optimization(obj.f, np.zeros(6))
    assert np.linalg.norm(np.asarray(x) - 1.0) < 1.0 and obj.f(x) < obj.f(np.zeros(6))
    assert str(opt) == "Instance of optimizer scipy."
    assert isinstance(Optimizer("monte_carlo", dict(n_blocks=1, max_iter=2, f_tol=1e-8, prob=0.25))._optimizer, MonteCarlo)
