# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`FFTerm` is one force-field term (a bond, an angle, a torsion, a nonbonded particle or a 1-4 scaling pair)
holding its :obj:`Parameter` objects by ``param_key``; same surface as the reference class
(``ParaMol/Force_field/force_field_term.py:11-81``).
"""
from .force_field_term_parameter import Parameter


class FFTerm:
    """
    Parameters
    ----------
    force_group : int or list of int
        Force group(s) of the force this term belongs to.
    idx : int
        Index of the term inside its force (the order the topology reader emitted).
    atoms : list of int
    symmetry_group : str
    """

    def __init__(self, force_group, idx, atoms, symmetry_group="X"):
        self.force_group = force_group
        self.idx = idx
        self.atoms = atoms
        self.parameters = {}
        self.symmetry_group = symmetry_group

    def __repr__(self):
        return "Force Field term with id {} belonging to force group {}. Contains {} parameters.".format(
            self.idx, self.force_group, len(self.parameters))

    __str__ = __repr__

    def add_parameter(self, symmetry_group, optimize, param_key, value):
        """Adds a :obj:`Parameter` under ``param_key`` (insertion order defines the order of the parameter vector)."""
        self.parameters[param_key] = Parameter(symmetry_group, optimize, param_key, value, self)
        return self.parameters
