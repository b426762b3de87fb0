# -*- coding: utf-8 -*-
"""
Force-field data model of the objective-evaluation path (mirrors ``ParaMol/Force_field``):
:obj:`Parameter`, :obj:`FFTerm` and :obj:`ForceField`.

Parameter keys per force group: HarmonicBondForce (bond_eq, bond_k), HarmonicAngleForce (angle_eq, angle_k),
PeriodicTorsionForce (torsion_periodicity, torsion_phase, torsion_k), NonbondedForce (charge, lj_sigma, lj_eps),
Scaling14 (scee, scnb).
"""
__all__ = ['force_field', 'force_field_term', 'force_field_term_parameter']
