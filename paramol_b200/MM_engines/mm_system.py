# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`MMSystem` is the plain-array molecular-mechanics system description that the B200 engine
consumes.  It plays the role that ``simtk.openmm.System`` plays in the reference
(``ParaMol/MM_engines/openmm.py:161-254``): the flat list of harmonic bonds, harmonic angles,
periodic torsions, nonbonded particles and nonbonded exceptions *in the order the topology reader
emitted them*.  That order is part of the parity contract: ``ForceField.create_force_field``
walks it to define the order of the optimisable-parameter vector
(``ParaMol/Force_field/force_field.py:351-563``).

All quantities are in the OpenMM unit system (nm, kJ/mol, rad, elementary charge).
"""
import copy
import numpy as np


class MMSystem:
    """
    Parameters / attributes
    -----------------------
    n_atoms : int
    masses : np.ndarray [N]
    bond_idx : int32 [nb,2];  bond_par : float64 [nb,2]  = (length r0, k)
    angle_idx : int32 [na,3]; angle_par : float64 [na,2] = (angle theta0, k)
    torsion_idx : int32 [nt,4]; torsion_per : int32 [nt]; torsion_par : float64 [nt,2] = (phase, k)
    particle_par : float64 [N,3] = (charge, sigma, epsilon)
    exception_idx : int32 [ne,2]; exception_par : float64 [ne,3] = (chargeProd, sigma, epsilon)
    nonbonded_method : "NoCutoff" | "CutoffNonPeriodic"
    cutoff, rf_dielectric : float (only used by CutoffNonPeriodic)
    one_4pi_eps0 : float, Coulomb constant (kJ nm / mol / e^2); OpenMM 7.x value 138.935456.
    """
    ONE_4PI_EPS0 = 138.935456

    def __init__(self, n_atoms):
        self.n_atoms = int(n_atoms)
        self.masses = np.zeros(self.n_atoms)
        self.atomic_numbers = np.zeros(self.n_atoms, dtype=np.int32)
        self.atom_names = [""] * self.n_atoms
        self.bond_idx = np.zeros((0, 2), dtype=np.int32)
        self.bond_par = np.zeros((0, 2))
        self.angle_idx = np.zeros((0, 3), dtype=np.int32)
        self.angle_par = np.zeros((0, 2))
        self.torsion_idx = np.zeros((0, 4), dtype=np.int32)
        self.torsion_per = np.zeros((0,), dtype=np.int32)
        self.torsion_par = np.zeros((0, 2))
        self.particle_par = np.zeros((self.n_atoms, 3))
        self.exception_idx = np.zeros((0, 2), dtype=np.int32)
        self.exception_par = np.zeros((0, 3))
        self.nonbonded_method = "NoCutoff"
        self.cutoff = 1.2
        self.rf_dielectric = 78.3
        self.one_4pi_eps0 = MMSystem.ONE_4PI_EPS0
        self.box = np.zeros((3, 3))
        # Order of the forces as an OpenMM System built from AMBER files would list them
        # (reference: OpenMMEngine._set_force_groups, openmm.py:849-878).
        self.force_names = ["HarmonicBondForce", "HarmonicAngleForce", "PeriodicTorsionForce",
                            "NonbondedForce", "CMMotionRemover"]

    # -- OpenMM-System-like accessors used by Force_field.create_force_field_from_engine --------
    def getNumParticles(self):
        return self.n_atoms

    def copy(self):
        return copy.deepcopy(self)

    def add_torsions(self, idx, per, par):
        self.torsion_idx = np.concatenate([self.torsion_idx, np.asarray(idx, dtype=np.int32).reshape(-1, 4)])
        self.torsion_per = np.concatenate([self.torsion_per, np.asarray(per, dtype=np.int32).reshape(-1)])
        self.torsion_par = np.concatenate([self.torsion_par, np.asarray(par, dtype=np.float64).reshape(-1, 2)])

    def excluded_pair_set(self):
        """Set of (min(i,j), max(i,j)) for every exception (they replace the plain pair interaction)."""
        return {(int(min(a, b)), int(max(a, b))) for a, b in self.exception_idx}

    def nonexcluded_pairs(self):
        """
        Pairs i<j that interact through the particle parameters, i.e. every pair that is not the
        subject of an exception, in row-major (i, then j) order.  int32 [n_pairs, 2].
        Indexing contract: bit-exact with the oracle's ``pm_oracle`` pair enumeration.
        """
        excl = self.excluded_pair_set()
        n = self.n_atoms
        pairs = [(i, j) for i in range(n) for j in range(i + 1, n) if (i, j) not in excl]
        return np.asarray(pairs, dtype=np.int32).reshape(-1, 2)

    def summary(self):
        return dict(n_atoms=self.n_atoms, n_bonds=len(self.bond_idx), n_angles=len(self.angle_idx),
                    n_torsions=len(self.torsion_idx), n_exceptions=len(self.exception_idx),
                    n_pairs=len(self.nonexcluded_pairs()))
