"""
Batched optimiser drivers (CPU): with an objective that offers f_batch the Monte Carlo, simulated-annealing and
finite-difference gradient-descent drivers must follow exactly the trajectory of the sequential algorithm (same random
stream, same decisions), using fewer launches.
"""
import numpy as np
import pytest

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


class _Quadratic:
    """Stand-in objective with the f / f_batch / f_and_grad / grad surface of ObjectiveFunction."""

    def __init__(self, n=6):
        rng = np.random.default_rng(4)
        a = rng.normal(size=(n, n))
        self.h = a @ a.T + n * np.eye(n)
        self.c = rng.normal(size=n)
        self.n_f = self.n_g = 0

    def f(self, x):
        self.n_f += 1
        x = np.asarray(x, dtype=np.float64)
        return float(0.5 * x @ self.h @ x - self.c @ x)

    def f_batch(self, X):
        return np.asarray([self.f(x) for x in np.atleast_2d(X)])

    def f_and_grad(self, x):
        self.n_g += 1
        x = np.asarray(x, dtype=np.float64)
        return float(0.5 * x @ self.h @ x - self.c @ x), self.h @ x - self.c

    def grad(self, x):
        return self.f_and_grad(x)[1]


def test_scipy_optimizer_analytic_jacobian():
    from paramol_b200.Optimizers.scipy_optimizers import ScipyOptimizer
    obj = _Quadratic()
    x = ScipyOptimizer(method="BFGS", jac="analytic", options=dict(gtol=1e-10)).run_optimization(obj.f, np.zeros(6))
    assert np.abs(x - np.linalg.solve(obj.h, obj.c)).max() < 1e-8
    assert obj.n_f == 0 and obj.n_g > 0                     # scipy never differenced f
    with pytest.raises(ValueError):
        ScipyOptimizer(method="BFGS", jac="analytic").run_optimization(lambda v: float(v @ v), np.zeros(3))
    # the reference's default settings still work unchanged
    x2 = ScipyOptimizer(method="BFGS", jac="2-point", options=dict(gtol=1e-8)).run_optimization(obj.f, np.zeros(6))
    assert np.abs(x2 - x).max() < 1e-5


def test_gradient_descent_analytic_matches_two_point():
    from paramol_b200.Optimizers.gradient_descent import GradientDescent
    obj = _Quadratic()
    kw = dict(max_iter=30, derivative_calculation="always", g_tol=1e-12, f_tol=1e-14, dx=1e-6, derivative_h=0.02)
    xa = GradientDescent(derivative_type="analytic", **kw).run_optimization(obj.f, np.zeros(6))
    xf = GradientDescent(derivative_type="2-point", **kw).run_optimization(obj.f, np.zeros(6))
    # the reference's 2-point quotient is (f+ - f-)/dx, i.e. twice the derivative: same path with half the step length
    xh = GradientDescent(derivative_type="analytic", **dict(kw, derivative_h=0.04)).run_optimization(obj.f, np.zeros(6))
    assert np.abs(xh - xf).max() < 1e-5
    assert obj.f(xa) < obj.f(np.zeros(6))
