# -*- coding: utf-8 -*-
"""Helper shared by the batched optimiser drivers."""
import numpy as np


def batch_evaluator(f):
    """
    Returns ``g(X) -> np.ndarray`` evaluating the rows of X.  If ``f`` is a bound method of an object that offers
    ``f_batch`` (``ObjectiveFunction``), one batched call is made; otherwise ``f`` is called row by row.
    The second return value tells whether batching is available.
    """
    owner = getattr(f, "__self__", None)
    fb = getattr(owner, "f_batch", None)
    if callable(fb) and getattr(f, "__name__", "") == "f":
        return (lambda X: np.asarray(fb(np.asarray(X, dtype=np.float64)), dtype=np.float64)), True
    return (lambda X: np.asarray([f(np.array(x, dtype=np.float64)) for x in X], dtype=np.float64)), False
