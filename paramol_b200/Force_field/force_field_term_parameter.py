# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`Parameter` is one force-field parameter (a force constant, an equilibrium value, a charge ...), with the
same attributes as the reference class (``ParaMol/Force_field/force_field_term_parameter.py:11-49``):
``symmetry_group``, ``optimize``, ``param_key``, ``value``, ``multiplicity``, ``ff_term``.

Difference from the reference: ``value`` is a property backed by a slot of a shared float64 array (the
:obj:`ForceField` "store").  Scattering a trial parameter vector to every tied parameter of every system
(``ParameterSpace.update_systems``, ``ParaMol/Parameter_space/parameter_space.py:378-380``) is then one
vectorised numpy assignment instead of a Python loop over objects, while ``parameter.value`` still reads and
writes like a plain attribute.
"""
import numpy as np


class Parameter:
    """
    Parameters
    ----------
    symmetry_group : str
        Symmetry group; parameters sharing (symmetry_group, param_key) are tied in symmetry-constrained fits.
    optimize : bool or int
        Whether this parameter is optimizable.
    param_key : str
        Role of the parameter inside its term (e.g. "bond_k").
    value : float or int
    ff_term : :obj:`paramol_b200.Force_field.force_field_term.FFTerm`, optional
    """
    __slots__ = ("symmetry_group", "optimize", "param_key", "multiplicity", "ff_term", "_store", "_idx", "_is_int")

    def __init__(self, symmetry_group, optimize, param_key, value, ff_term=None):
        self.symmetry_group = symmetry_group
        self.optimize = optimize
        self.param_key = param_key
        self.multiplicity = 1.0
        self.ff_term = ff_term
        self._is_int = isinstance(value, (int, np.integer)) and not isinstance(value, bool)
        self._store = np.array([float(value)], dtype=np.float64)
        self._idx = 0

    # -- array-backed value --------------------------------------------------------------------
    @property
    def value(self):
        v = self._store[self._idx]
        return int(round(v)) if self._is_int else float(v)

    @value.setter
    def value(self, v):
        self._store[self._idx] = v

    def _bind(self, store, idx):
        """Move this parameter's storage into ``store[idx]`` (keeps the current value)."""
        store[idx] = self._store[self._idx]
        self._store = store
        self._idx = int(idx)

    def __str__(self):
        description = "\n"
        description += "Parameter key: {} \n".format(self.param_key)
        description += "Value: {} \n".format(self.value)
        description += "Optimize: {} \n".format(self.optimize)
        description += "Symmetry group:  {} \n".format(self.symmetry_group)
        return description

    def __getstate__(self):
        return dict(symmetry_group=self.symmetry_group, optimize=self.optimize, param_key=self.param_key,
                    multiplicity=self.multiplicity, ff_term=self.ff_term, value=self.value, is_int=self._is_int)

    def __setstate__(self, st):
        self.symmetry_group = st["symmetry_group"]
        self.optimize = st["optimize"]
        self.param_key = st["param_key"]
        self.multiplicity = st["multiplicity"]
        self.ff_term = st["ff_term"]
        self._is_int = st["is_int"]
        self._store = np.array([float(st["value"])], dtype=np.float64)
        self._idx = 0
