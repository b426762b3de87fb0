# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`ForceField` is the in-memory force field of one system,
``force_field[force_name][occurrence] = [FFTerm, ...]``, with the same public surface and the same ordering
contract as the reference class (``ParaMol/Force_field/force_field.py:15-929``).  The order in which forces,
terms and parameters are walked defines the order of the optimisation vector x
(``get_optimizable_parameters``, reference ``:273-349``), so it is part of the parity contract.

Where the reference reads the terms out of a ``simtk.openmm.System`` (``:351-563``) this class reads them out of
the engine's :obj:`paramol_b200.MM_engines.mm_system.MMSystem`, which lists them in the same order.

All :obj:`Parameter` values of a force field live in one float64 array (``self.store``): see
``force_field_term_parameter.py``.  ``self.version`` is bumped whenever the set of optimizable terms may have
changed, which invalidates the compiled update plans of the engine.
"""
import copy
import logging
import os

import numpy as np

from .force_field_term import FFTerm


class ForceField:
    """
    Parameters
    ----------
    openmm_engine : :obj:`paramol_b200.MM_engines.b200_engine.B200Engine`
        Engine that owns the MM system (argument name kept from the reference).
    """
    symmetry_group_default = "X"

    def __init__(self, openmm_engine):
        self._openmm = openmm_engine
        self.force_field = None
        self.force_field_optimizable = None
        self.force_groups = None
        self.optimizable_parameters = None
        self.optimizable_parameters_values = None
        self.store = None
        self.version = 0

    # ------------------------------------------------------------------------------------------
    #                                       PUBLIC METHODS
    # ------------------------------------------------------------------------------------------
    def create_force_field(self, opt_bonds=False, opt_angles=False, opt_torsions=False, opt_charges=False, opt_lj=False,
                           opt_sc=False, ff_file=None):
        """Builds ``force_field`` from the engine's system (or a ``.ff`` file) and ``force_field_optimizable``."""
        self.force_field = {}
        self.force_groups = {}
        if ff_file is None:
            logging.info("Creating force field directly from the MM system.")
            assert self._openmm is not None, "OpenMM was not set."
            self.create_force_field_from_openmm(opt_bonds, opt_angles, opt_torsions, opt_charges, opt_lj, opt_sc)
        else:
            logging.info("Creating force from .ff file named '{}'.".format(ff_file))
            assert os.path.exists(ff_file), "\t * ERROR: .param file provided - {} - does not exist."
            self.read_ff_file(ff_file)
        self._bind_store()
        self.create_force_field_optimizable()
        return self.force_field

    def _bind_store(self):
        """Moves every Parameter of ``force_field`` into one shared float64 array."""
        params = [p for force in self.force_field for sub in self.force_field[force] for term in sub
                  for p in term.parameters.values()]
        # Scaling14 terms of pure exclusions are not listed in force_field (reference :532-538) -- nothing to bind
        self.store = np.zeros(len(params), dtype=np.float64)
        for i, p in enumerate(params):
            p._bind(self.store, i)
        return self.store

    def update_force_field(self, optimizable_parameters_values, symmetry_constrained=True):
        """
        Writes ``optimizable_parameters_values`` into the Parameter objects (reference ``:102-169``), tying
        symmetry groups, then forces ``scee``/``scnb`` and ``lj_eps``/``lj_sigma`` to be non-negative when their
        force group is optimizable.
        """
        if symmetry_constrained:
            symm_groups = {}
            for i in range(len(self.optimizable_parameters)):
                parameter = self.optimizable_parameters[i]
                if parameter.symmetry_group == self.symmetry_group_default:
                    parameter.value = optimizable_parameters_values[i]
                else:
                    group = symm_groups.setdefault(parameter.symmetry_group, {})
                    if parameter.param_key not in group:
                        group[parameter.param_key] = optimizable_parameters_values[i]
            for force in self.force_field_optimizable:
                for sub_force in self.force_field_optimizable[force]:
                    for force_field_term in sub_force:
                        for parameter in force_field_term.parameters.values():
                            if parameter.optimize and parameter.symmetry_group != self.symmetry_group_default:
                                parameter.value = symm_groups[parameter.symmetry_group][parameter.param_key]
        else:
            for i in range(len(self.optimizable_parameters)):
                self.optimizable_parameters[i].value = optimizable_parameters_values[i]
        self.apply_sign_conventions()
        return self.optimizable_parameters

    def apply_sign_conventions(self):
        """abs() on scee/scnb and lj_eps/lj_sigma of optimizable groups (reference ``:157-167``)."""
        if "Scaling14" in self.force_field_optimizable:
            for sub_force in self.force_field_optimizable["Scaling14"]:
                for ff_term in sub_force:
                    ff_term.parameters["scee"].value = abs(ff_term.parameters["scee"].value)
                    ff_term.parameters["scnb"].value = abs(ff_term.parameters["scnb"].value)
        if "NonbondedForce" in self.force_field_optimizable:
            for sub_force in self.force_field_optimizable["NonbondedForce"]:
                for ff_term in sub_force:
                    ff_term.parameters["lj_eps"].value = abs(ff_term.parameters["lj_eps"].value)
                    ff_term.parameters["lj_sigma"].value = abs(ff_term.parameters["lj_sigma"].value)

    def create_force_field_from_openmm(self, opt_bonds, opt_angles, opt_torsions, opt_charges, opt_lj, opt_sc):
        """Walks the engine's system force by force (reference ``:171-229``)."""
        assert self._openmm.system is not None, "System was not set"
        self.force_groups = copy.deepcopy(self._openmm.forces_indexes)
        force_key = "Scaling14"
        assert force_key not in self.force_groups, "\t * ERROR: Force {} already in the dictionary.".format(force_key)
        self.force_groups[force_key] = self.force_groups["NonbondedForce"]
        self.create_harmonic_bond_force_field(opt_bonds)
        self.create_harmonic_angle_force_field(opt_angles)
        self.create_periodic_torsion_force_field(opt_torsions)
        self.create_nonbonded_force_field(opt_charges, opt_lj, opt_sc)
        return self.force_field

    def create_force_field_optimizable(self):
        """Subset of ``force_field`` holding the terms with at least one optimizable parameter (reference ``:231-271``)."""
        assert self.force_field is not None, "\t * force_field dictionary was not created yet. Run create_force_field " \
                                             "method before"
        self.force_field_optimizable = {}
        for force in self.force_field:
            for sub_force in self.force_field[force]:
                sub_force_field_optimizable = []
                for force_field_term in sub_force:
                    for parameter in force_field_term.parameters.values():
                        if parameter.optimize:
                            if force not in self.force_field_optimizable:
                                self.force_field_optimizable[force] = []
                            sub_force_field_optimizable.append(force_field_term)
                            break
                if force in self.force_field_optimizable:
                    self.force_field_optimizable[force].append(sub_force_field_optimizable)
        self.version += 1
        return self.force_field_optimizable

    def get_optimizable_parameters(self, symmetry_constrained=True):
        """
        Lists of optimizable Parameter instances and of their values, in vector order (reference ``:273-349``).
        With symmetry constraints the first parameter met for a (symmetry_group, param_key) represents the group
        and later ones bump its ``multiplicity``.
        """
        assert self.force_field_optimizable is not None, "\t * force_field_optimizable dictionary was not created yet." \
                                                         " First run create_force_field_optimizable method."
        self.optimizable_parameters = []
        self.optimizable_parameters_values = []
        ref_parameters = {}
        for force in self.force_field_optimizable:
            for sub_force in self.force_field_optimizable[force]:
                for force_field_term in sub_force:
                    for parameter in force_field_term.parameters.values():
                        if not parameter.optimize:
                            continue
                        if not symmetry_constrained or parameter.symmetry_group == self.symmetry_group_default:
                            self.optimizable_parameters.append(parameter)
                            self.optimizable_parameters_values.append(parameter.value)
                            continue
                        group = ref_parameters.setdefault(parameter.symmetry_group, {})
                        if parameter.param_key not in group:
                            group[parameter.param_key] = parameter
                            parameter.multiplicity = 1
                            self.optimizable_parameters.append(parameter)
                            self.optimizable_parameters_values.append(parameter.value)
                        else:
                            group[parameter.param_key].multiplicity += 1
        return self.optimizable_parameters, self.optimizable_parameters_values

    # ---- per-force builders ------------------------------------------------------------------
    def _new_force(self, force_key):
        assert force_key not in self.force_field, "\t * ERROR: Force group {} already exists.".format(force_key)
        self.force_field[force_key] = []

    def create_harmonic_bond_force_field(self, opt_bonds):
        force_key = "HarmonicBondForce"
        self._new_force(force_key)
        mm = self._openmm.system
        for _force_idx in self.force_groups[force_key]:
            sub_force_field = []
            for i in range(len(mm.bond_idx)):
                term = FFTerm(self.force_groups[force_key], i, [int(mm.bond_idx[i, 0]), int(mm.bond_idx[i, 1])])
                term.add_parameter(self.symmetry_group_default, int(opt_bonds), "bond_eq", float(mm.bond_par[i, 0]))
                term.add_parameter(self.symmetry_group_default, int(opt_bonds), "bond_k", float(mm.bond_par[i, 1]))
                sub_force_field.append(term)
            self.force_field[force_key].append(sub_force_field)
        return self.force_field

    def create_harmonic_angle_force_field(self, opt_angles):
        force_key = "HarmonicAngleForce"
        self._new_force(force_key)
        mm = self._openmm.system
        for _force_idx in self.force_groups[force_key]:
            sub_force_field = []
            for i in range(len(mm.angle_idx)):
                term = FFTerm(self.force_groups[force_key], i, [int(a) for a in mm.angle_idx[i]])
                term.add_parameter(self.symmetry_group_default, int(opt_angles), "angle_eq", float(mm.angle_par[i, 0]))
                term.add_parameter(self.symmetry_group_default, int(opt_angles), "angle_k", float(mm.angle_par[i, 1]))
                sub_force_field.append(term)
            self.force_field[force_key].append(sub_force_field)
        return self.force_field

    def create_periodic_torsion_force_field(self, opt_torsions):
        force_key = "PeriodicTorsionForce"
        self._new_force(force_key)
        mm = self._openmm.system
        sub_force_field = []
        for _force_idx in self.force_groups[force_key]:
            sub_force_field = []
            for i in range(len(mm.torsion_idx)):
                term = FFTerm(self.force_groups[force_key], i, [int(a) for a in mm.torsion_idx[i]])
                term.add_parameter(self.symmetry_group_default, 0, "torsion_periodicity", int(mm.torsion_per[i]))
                term.add_parameter(self.symmetry_group_default, int(opt_torsions), "torsion_phase", float(mm.torsion_par[i, 0]))
                term.add_parameter(self.symmetry_group_default, int(opt_torsions), "torsion_k", float(mm.torsion_par[i, 1]))
                sub_force_field.append(term)
        # the reference appends once, after the loop over force occurrences (:472-473)
        self.force_field[force_key].append(sub_force_field)
        return self.force_field

    def create_nonbonded_force_field(self, opt_charges, opt_lj, opt_sc):
        force_key = "NonbondedForce"
        self._new_force(force_key)
        mm = self._openmm.system
        for _force_idx in self.force_groups[force_key]:
            sub_force_field = []
            for i in range(mm.n_atoms):
                charge, sigma, eps = (float(v) for v in mm.particle_par[i])
                term = FFTerm(self.force_groups[force_key], i, [i])
                term.add_parameter(self.symmetry_group_default, int(opt_charges), "charge", charge)
                term.add_parameter(self.symmetry_group_default, int(opt_lj), "lj_sigma", sigma)
                term.add_parameter(self.symmetry_group_default, int(opt_lj), "lj_eps", eps)
                sub_force_field.append(term)
            self.force_field[force_key].append(sub_force_field)

            force_key = "Scaling14"
            self._new_force(force_key)
            sub_force_field = []
            for i in range(len(mm.exception_idx)):
                at1, at2 = int(mm.exception_idx[i, 0]), int(mm.exception_idx[i, 1])
                charge_prod, _sigma, eps = (float(v) for v in mm.exception_par[i])
                if abs(charge_prod) < 1e-8 and abs(eps) < 1e-8:
                    # pure exclusion: the reference builds the term and then skips it (:532-538)
                    continue
                charge_at1, _s1, eps_at1 = (float(v) for v in mm.particle_par[at1])
                charge_at2, _s2, eps_at2 = (float(v) for v in mm.particle_par[at2])
                try:
                    scee = charge_prod / (charge_at1 * charge_at2)
                except ZeroDivisionError:
                    scee = 1 / 1.2
                try:
                    scnb = eps / np.sqrt(eps_at1 * eps_at2)
                    if not np.isfinite(scnb):
                        raise ZeroDivisionError
                except (ZeroDivisionError, FloatingPointError):
                    scnb = 1 / 2.0
                term = FFTerm(self.force_groups[force_key], i, [at1, at2])
                term.add_parameter(self.symmetry_group_default, int(opt_sc), "scee", float(scee))
                term.add_parameter(self.symmetry_group_default, int(opt_sc), "scnb", float(scnb))
                sub_force_field.append(term)
            self.force_field[force_key].append(sub_force_field)
        return self.force_field

    # ---- .ff text format (reference :565-703; docs/source/Files_specification.rst:4-134) ------------
    def write_ff_file(self, file_name):
        logging.info("Writing force field to .ff file named '{}'.".format(file_name))
        with open(file_name, 'w') as ff_file:
            for force in self.force_field:
                for k, sub_force in enumerate(self.force_field[force]):
                    ff_file.write("{} {:3d} \n".format(force, self.force_groups[force][k]))
                    for force_field_term in sub_force:
                        line = ("{:3d} " + "{:3d} " * len(force_field_term.atoms)).format(force_field_term.idx,
                                                                                         *force_field_term.atoms)
                        flags = ""
                        parameter = None
                        for parameter in force_field_term.parameters.values():
                            line += "{:16.8f} ".format(parameter.value)
                            flags += "{:3d} ".format(int(parameter.optimize))
                        line += flags
                        line += "  " + str(parameter.symmetry_group) + " \n"
                        ff_file.write(line)
            ff_file.write("END \n")

    _FF_LAYOUT = {
        'HarmonicBondForce': (2, [("bond_eq", float), ("bond_k", float)]),
        'HarmonicAngleForce': (3, [("angle_eq", float), ("angle_k", float)]),
        'PeriodicTorsionForce': (4, [("torsion_periodicity", lambda s: int(float(s))), ("torsion_phase", float),
                                     ("torsion_k", float)]),
        'NonbondedForce': (1, [("charge", float), ("lj_sigma", float), ("lj_eps", float)]),
        'Scaling14': (2, [("scee", float), ("scnb", float)]),
    }

    def read_ff_file(self, file_name):
        force_key = None
        current_occurrence = None
        with open(file_name, 'r') as ff_file:
            for line in ff_file:
                line_split = line.split()
                if 'END' in line_split:
                    break
                if len(line_split) == 2:
                    force_key = line_split[0]
                    force_index = int(line_split[1])
                    self.force_groups.setdefault(force_key, []).append(force_index)
                    self.force_field.setdefault(force_key, []).append([])
                    current_occurrence = len(self.force_field[force_key]) - 1
                    continue
                if not line_split:
                    continue
                if force_key not in self._FF_LAYOUT:
                    raise NotImplementedError("Force {} of the .ff file is not known.".format(force_key))
                n_at, keys = self._FF_LAYOUT[force_key]
                n_par = len(keys)
                assert len(line_split) == 1 + n_at + 2 * n_par + 1, "Malformed .ff line: {}".format(line)
                idx = int(line_split[0])
                atoms = [int(a) for a in line_split[1:1 + n_at]]
                values = line_split[1 + n_at:1 + n_at + n_par]
                flags = [int(v) for v in line_split[1 + n_at + n_par:1 + n_at + 2 * n_par]]
                symm_group = line_split[-1]
                term = FFTerm(self.force_groups[force_key], idx, atoms)
                for (key, conv), val, flag in zip(keys, values, flags):
                    if key == "torsion_periodicity":
                        assert flag == 0, "Flag to parameterize torsions was set to {} but this is not possible.".format(flag)
                    term.add_parameter(symm_group, flag, key, conv(val))
                self.force_field[force_key][current_occurrence].append(term)

    # ---- selection helpers (reference :705-929) ------------------------------------------------
    def _touch(self):
        self.version += 1

    def optimize_selection(self, lower_idx, upper_idx, change_other=False):
        assert len(lower_idx) == len(upper_idx)
        for force in self.force_field:
            for sub_force in self.force_field[force]:
                for force_field_term in sub_force:
                    for at in force_field_term.atoms:
                        for low_limit, upper_limit in zip(lower_idx, upper_idx):
                            if (at >= low_limit) and (at <= upper_limit):
                                for parameter in force_field_term.parameters.values():
                                    parameter.optimize = 1
                            elif (at < low_limit) or (at > upper_limit) and change_other:
                                for parameter in force_field_term.parameters.values():
                                    parameter.optimize = 0
        self._touch()
        return self.force_field

    def _clear_force(self, force):
        for sub_force in self.force_field[force]:
            for force_field_term in sub_force:
                for parameter in force_field_term.parameters.values():
                    parameter.optimize = 0

    def optimize_torsions(self, torsions, change_other_torsions=False, change_other_parameters=False):
        for force in self.force_field:
            if force == 'PeriodicTorsionForce':
                for sub_force in self.force_field[force]:
                    for force_field_term in sub_force:
                        for parameter in force_field_term.parameters.values():
                            if parameter.param_key != "torsion_periodicity":
                                if force_field_term.atoms in torsions:
                                    parameter.optimize = 1
                                elif change_other_torsions:
                                    parameter.optimize = 0
            elif change_other_parameters:
                self._clear_force(force)
        self._touch()
        return self.force_field

    def optimize_scaling_constants(self, atom_pairs, change_other_sc=False, change_other_parameters=False):
        for force in self.force_field:
            if force == 'Scaling14':
                for sub_force in self.force_field[force]:
                    for force_field_term in sub_force:
                        for parameter in force_field_term.parameters.values():
                            if force_field_term.atoms in atom_pairs:
                                parameter.optimize = 1
                            elif change_other_sc:
                                parameter.optimize = 0
            elif change_other_parameters:
                self._clear_force(force)
        self._touch()
        return self.force_field

    def optimize_torsions_by_symmetry(self, torsions, change_other_torsions=False, change_other_parameters=False,
                                      set_zero=False):
        dihedral_types = []
        for sub_force in self.force_field.get('PeriodicTorsionForce', []):
            for force_field_term in sub_force:
                if force_field_term.atoms in torsions:
                    dihedral_types.append(force_field_term.parameters["torsion_periodicity"].symmetry_group)
        for force in self.force_field:
            if force == 'PeriodicTorsionForce':
                for sub_force in self.force_field[force]:
                    for force_field_term in sub_force:
                        for parameter in force_field_term.parameters.values():
                            if parameter.param_key != "torsion_periodicity":
                                if parameter.symmetry_group in dihedral_types:
                                    parameter.optimize = 1
                                    if parameter.param_key == "torsion_k" and set_zero:
                                        parameter.value = 0.0
                                elif change_other_torsions:
                                    parameter.optimize = 0
            elif change_other_parameters:
                self._clear_force(force)
        self._touch()
        return self.force_field

    def set_parameter_optimization(self, force_key, sub_force, idx, param_key, optimize):
        self.force_field[force_key][sub_force][idx].parameters[param_key].optimize = optimize
        self._touch()
        return self.force_field
