# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`B200Engine` is the drop-in for ``ParaMol.MM_engines.openmm.OpenMMEngine`` on the objective-evaluation
path (seam S2 of SURVEY.md section 8b).  It keeps the constructor keywords and the methods the hot path uses

    get_potential_energy / get_forces                (ParaMol/MM_engines/openmm.py:448-502)
    set_harmonic_bond_force_parameters / ..angle.. / ..torsion.. / set_bonded_parameters   (:507-613)
    set_non_bonded_parameters_to_zero / set_nonbonded_parameters                           (:618-756)
    write_system_xml, get_atom_list, get_atomic_numbers, get_masses, get_cell, get_number_of_atoms

but owns a plain-array :obj:`MMSystem` instead of a ``simtk.openmm.System`` and evaluates on the GPU through the
C ABI (``include/paramol_b200.h``).  There is no CPU fallback: evaluation without a CUDA device raises.

Two parameter-update paths exist and are tested against each other:

* the object path: the reference's ``set_*_parameters(force_field_optimizable)`` methods, term by term;
* the compiled path (:meth:`compile_update_plan` / :meth:`slots_from_store`): index arrays built once per
  force field that turn the force field's value store into slot rows for any number of trial vectors with a
  few numpy gathers -- this is what ``ObjectiveFunction.f`` uses.
"""
import logging

import numpy as np

from .mm_system import MMSystem
from ..Utils.amber import system_from_prmtop, read_inpcrd
from ..Utils.openmm_xml import system_from_xml, write_system_xml as _write_system_xml

_SYMBOLS = {1: "H", 2: "He", 3: "Li", 4: "Be", 5: "B", 6: "C", 7: "N", 8: "O", 9: "F", 10: "Ne", 11: "Na", 12: "Mg",
            13: "Al", 14: "Si", 15: "P", 16: "S", 17: "Cl", 18: "Ar", 19: "K", 20: "Ca", 26: "Fe", 29: "Cu", 30: "Zn",
            34: "Se", 35: "Br", 53: "I"}


class B200Engine:
    """
    Parameters
    ----------
    init_openmm : bool, optional, default=False
        Whether to build the MM system from the files upon construction (keyword name kept from the reference).
    topology_format : str
        "AMBER" (prmtop) or "XML" is accepted through ``xml_file``; GROMACS/CHARMM readers are out of scope.
    top_file, crd_format, crd_file, xml_file : str
    platform_name : str
        Accepted for signature compatibility ("Reference", "CPU", "OpenCL", "CUDA", "B200"); evaluation always
        runs on the B200 path.
    system : :obj:`MMSystem`, optional
        Pre-built system.
    create_system_params : dict
        ``nonbondedMethod`` ("NoCutoff" / "CutoffNonPeriodic"), ``nonbondedCutoff`` (nm), ``rfDielectric``.
    device : int, optional
        CUDA device index.
    """
    force_groups_dict = {'HarmonicBondForce': 0,
                         'HarmonicAngleForce': 1,
                         'PeriodicTorsionForce': 2,
                         'NonbondedForce': 11,
                         'CMMotionRemover': 3,
                         'CustomBondForce': 5,
                         'CustomAngleForce': 6,
                         'CustomTorsionForce': 7,
                         "CMAPTorsionForce": 12}

    def __init__(self, init_openmm=False, topology_format=None, top_file=None, crd_format=None, crd_file=None,
                 charmm_param_file=None, xml_file=None, platform_name='B200', system=None, integrator=None, platform=None,
                 context=None, topology=None,
                 integrator_params={'temperature': 300.0, 'stepSize': 0.001, 'frictionCoeff': 2.0},
                 create_system_params={'nonbondedMethod': "NoCutoff", 'nonbondedCutoff': 1.2, 'constraints': None,
                                       'rigidWater': True},
                 device=None):
        self.topology_format = topology_format
        self.crd_format = crd_format
        self.xml_file = xml_file
        self.top_file = top_file
        self.crd_file = crd_file
        self.charmm_param_file = charmm_param_file

        self.system = system
        self.integrator = integrator
        self.platform = platform
        self.context = context
        self.topology = topology
        self.platform_name = platform_name if platform_name is not None else 'B200'

        self.force_groups = []
        self.forces_indexes = {}
        self.atom_list = None
        self.atomic_number_list = None
        self.masses_list = None
        self.n_atoms = None
        self.cell = None
        self.positions = None

        # 1-4 scaling constants used by branch A of set_nonbonded_parameters; the reference swaps the
        # attribute names but the numbers land on the right terms (openmm.py:143-144, 701-716)
        self.scnb = 0.833333
        self.scee = 0.5

        self._create_system_params = create_system_params
        self._integrator_params = integrator_params
        self._device = device
        self._topo = None            # DeviceTopology, created lazily
        self._plan = None
        self._slots_cache = None
        self._single = {}            # cached 1-conformation ensembles for get_potential_energy/get_forces

        if init_openmm:
            self.init_openmm(self._integrator_params, self._create_system_params)
        if self.system is not None:
            self._set_force_groups()
            self.n_atoms = self.system.n_atoms

    # ------------------------------------------------------------------------------------------
    #                                       PUBLIC METHODS
    # ------------------------------------------------------------------------------------------
    def init_openmm(self, integrator_params=None, create_system_params=None):
        """Builds the MM system from ``top_file`` / ``xml_file`` and reads ``crd_file`` (reference :161-254)."""
        assert self.topology_format is not None, "No topology format was provided."
        assert self.crd_format is not None, "No coordinate format was provided."
        if self.topology_format.upper() in ["AMBER", "GROMACS", "CHARMM"]:
            assert self.top_file is not None, "Topology format is {} but no topology file was provided.".format(self.topology_format)
        else:
            raise NotImplementedError("Topology format {} is not known.".format(self.topology_format))
        assert self.crd_file is not None, "create_system flag is True but no crd_file was provided."
        if self.topology_format.upper() != "AMBER":
            raise NotImplementedError("Topology format {} is not currently supported by the B200 engine.".format(self.topology_format))
        if self.crd_format.upper() != "AMBER":
            raise NotImplementedError("Coordinate format {} is not currently supported by the B200 engine.".format(self.crd_format))

        params = dict(create_system_params or {})
        method = params.get('nonbondedMethod', "NoCutoff")
        method = getattr(method, "__name__", method)
        method = str(method)
        if self.system is None:
            if self.xml_file is None:
                logging.info("Creating MM system from {} file.".format(self.topology_format))
                self.system = system_from_prmtop(self.top_file, nonbonded_method=method,
                                                 cutoff=float(getattr(params.get('nonbondedCutoff', 1.2), "_value",
                                                                      params.get('nonbondedCutoff', 1.2))),
                                                 rf_dielectric=float(params.get('rfDielectric', 78.3)))
            else:
                logging.info("Creating MM system from XML file.")
                self.system = system_from_xml(self.xml_file)
                if not self.system.atom_names or not any(self.system.atom_names):
                    meta = system_from_prmtop(self.top_file)
                    self.system.atom_names = meta.atom_names
                    self.system.atomic_numbers = meta.atomic_numbers
        self.positions = read_inpcrd(self.crd_file)
        self.n_atoms = self.system.n_atoms
        self._set_force_groups()
        return self.system

    def get_atom_list(self):
        self.atom_list = [_SYMBOLS.get(int(z), "X") for z in self.system.atomic_numbers]
        return self.atom_list

    def get_atomic_numbers(self):
        self.atomic_number_list = [int(z) for z in self.system.atomic_numbers]
        return self.atomic_number_list

    def get_number_of_atoms(self):
        assert self.system is not None, "MM system is not set."
        self.n_atoms = self.system.getNumParticles()
        return self.n_atoms

    def get_masses(self):
        self.masses_list = [float(m) for m in self.system.masses]
        return self.masses_list

    def get_cell(self):
        """(3,3) cell vectors in angstrom (reference :318-339)."""
        assert self.system is not None, "MM system is not set."
        self.cell = np.zeros((3, 3))
        for i in range(3):
            self.cell[i, i] = self.system.box[i, i]
        self.cell = self.cell * 10.0
        return self.cell

    def write_system_xml(self, file_name):
        """Serialises the current MM system (reference :341-360)."""
        logging.info("Writing serialized system to XML file {}.".format(file_name))
        _write_system_xml(self.system, file_name)

    def set_positions(self, positions):
        self.positions = np.asarray(getattr(positions, "_value", positions), dtype=np.float64).reshape(-1, 3)
        return self.positions

    def get_positions(self):
        return self.positions

    # ---- device plumbing ---------------------------------------------------------------------
    def device_topology(self):
        """The :obj:`DeviceTopology` of this engine (created on first use; raises without a GPU)."""
        if self._topo is None:
            from .device import DeviceTopology
            self._topo = DeviceTopology(self.system, self._device)
        return self._topo

    def current_slots(self):
        """Physical parameter table of the current MM system as one slots row (a fresh array)."""
        if self._slots_cache is not None:   # left by load_slots; every other writer of the parameter arrays clears it
            return self._slots_cache.copy()
        s = self.system
        return np.concatenate([s.bond_par.ravel(), s.angle_par.ravel(), s.torsion_par.ravel(), s.particle_par.ravel(),
                               s.exception_par.ravel()]).astype(np.float64)

    def _evaluate_positions(self, positions, with_forces):
        import torch
        from .device import DeviceEnsemble
        topo = self.device_topology()
        x = np.asarray(getattr(positions, "_value", positions), dtype=np.float64).reshape(1, self.system.n_atoms, 3)
        ens = DeviceEnsemble(topo, x)
        derived = topo.prepare(torch.from_numpy(self.current_slots()[None, :]).to(ens.device))
        emm, _, fmm = ens.evaluate(derived, with_forces=with_forces, want_fres=False, want_fmm=with_forces)
        e = float(emm[0, 0].item())
        f = ens.forces_from_tiles(fmm)[0, 0].cpu().numpy() if with_forces else None
        return e, f

    # ---- energies and forces (reference :448-502) ------------------------------------------------
    def get_potential_energy(self, positions):
        """Potential energy (kJ/mol) of one conformation (positions in nm)."""
        return self._evaluate_positions(positions, False)[0]

    def get_forces(self, positions):
        """(N,3) forces (kJ/mol/nm) of one conformation."""
        return self._evaluate_positions(positions, True)[1]

    # ---- bonded parameters (reference :507-613) ---------------------------------------------------
    def set_harmonic_bond_force_parameters(self, ff_bond_terms):
        self._slots_cache = None
        for k, _ in enumerate(self.forces_indexes["HarmonicBondForce"]):
            for bond_term in ff_bond_terms[k]:
                self.system.bond_par[bond_term.idx, 0] = bond_term.parameters["bond_eq"].value
                self.system.bond_par[bond_term.idx, 1] = bond_term.parameters["bond_k"].value
        return self.context

    def set_harmonic_angle_force_parameters(self, ff_angle_terms):
        self._slots_cache = None
        for k, _ in enumerate(self.forces_indexes["HarmonicAngleForce"]):
            for angle_term in ff_angle_terms[k]:
                self.system.angle_par[angle_term.idx, 0] = angle_term.parameters["angle_eq"].value
                self.system.angle_par[angle_term.idx, 1] = angle_term.parameters["angle_k"].value
        return self.context

    @staticmethod
    def wrap_phase(phase):
        """phase % (sign(phase) * 2 pi), with 2 pi for phase == 0 (reference :576-584); works on arrays."""
        # a modulus that carries the sign of the dividend is the C fmod: bit-identical to np.mod(phase, sign(phase) * 2 pi)
        # (np.mod is fmod plus a correction that only fires when the signs of divisor and remainder differ), in one pass
        return np.fmod(np.asarray(phase, dtype=np.float64), 2.0 * np.pi)

    def set_periodic_torsion_force_parameters(self, ff_torsion_terms):
        self._slots_cache = None
        for k, _ in enumerate(self.forces_indexes["PeriodicTorsionForce"]):
            for torsion_term in ff_torsion_terms[k]:
                self.system.torsion_per[torsion_term.idx] = torsion_term.parameters["torsion_periodicity"].value
                self.system.torsion_par[torsion_term.idx, 0] = float(self.wrap_phase(torsion_term.parameters["torsion_phase"].value))
                self.system.torsion_par[torsion_term.idx, 1] = torsion_term.parameters["torsion_k"].value
        return self.context

    def set_bonded_parameters(self, force_field_optimizable):
        if "HarmonicBondForce" in force_field_optimizable:
            self.set_harmonic_bond_force_parameters(force_field_optimizable["HarmonicBondForce"])
        if "HarmonicAngleForce" in force_field_optimizable:
            self.set_harmonic_angle_force_parameters(force_field_optimizable["HarmonicAngleForce"])
        if "PeriodicTorsionForce" in force_field_optimizable:
            self.set_periodic_torsion_force_parameters(force_field_optimizable["PeriodicTorsionForce"])
        return self.context

    # ---- nonbonded parameters (reference :618-756) ------------------------------------------------
    def set_non_bonded_parameters_to_zero(self):
        """
        Zeroes charges, sigmas and epsilons of all particles; interacting exceptions are set to 1e-16 as the
        reference does on the CPU/Reference platforms (:640-650) so that they stay "interacting".
        """
        self._slots_cache = None
        s = self.system
        s.particle_par[:, :] = 0.0
        sel = (np.abs(s.exception_par[:, 0]) > 1e-16) & (s.exception_par[:, 2] > 1e-16)
        s.exception_par[sel, :] = 1e-16
        return self.context

    def set_nonbonded_parameters(self, force_field_optimizable):
        """The three-way branch of the reference (:661-756), including its quirks."""
        self._slots_cache = None
        s = self.system
        if "NonbondedForce" in force_field_optimizable and "Scaling14" not in force_field_optimizable:
            nonbonded_force_terms = force_field_optimizable["NonbondedForce"]
            for k, _ in enumerate(self.forces_indexes["NonbondedForce"]):
                for term in nonbonded_force_terms[k]:
                    s.particle_par[term.idx, 0] = term.parameters["charge"].value
                    s.particle_par[term.idx, 1] = term.parameters["lj_sigma"].value
                    s.particle_par[term.idx, 2] = term.parameters["lj_eps"].value
                for i in range(len(s.exception_idx)):
                    at1, at2 = int(s.exception_idx[i, 0]), int(s.exception_idx[i, 1])
                    if abs(s.exception_par[i, 0]) < 1e-8:
                        continue
                    # the reference indexes the *optimizable* term list by atom index
                    p1, p2 = nonbonded_force_terms[k][at1].parameters, nonbonded_force_terms[k][at2].parameters
                    epsilon = self.scee * np.sqrt(p1["lj_eps"].value * p2["lj_eps"].value)
                    sigma = 0.5 * (p1["lj_sigma"].value + p2["lj_sigma"].value)
                    charge_prod = self.scnb * p1["charge"].value * p2["charge"].value
                    s.exception_par[i, :] = (charge_prod, sigma, epsilon)
        elif "NonbondedForce" not in force_field_optimizable and "Scaling14" in force_field_optimizable:
            scaling_constants = force_field_optimizable["Scaling14"]
            for k, _ in enumerate(self.forces_indexes["NonbondedForce"]):
                for i in range(len(scaling_constants[k])):
                    # exception i, i.e. the POSITION in the Scaling14 list, not term.idx (reference :731-748)
                    at1, at2 = int(s.exception_idx[i, 0]), int(s.exception_idx[i, 1])
                    chg1, _, eps1 = s.particle_par[at1]
                    chg2, _, eps2 = s.particle_par[at2]
                    scee_local = abs(scaling_constants[k][i].parameters['scee'].value)
                    scnb_local = abs(scaling_constants[k][i].parameters['scnb'].value)
                    s.exception_par[i, 0] = scee_local * chg1 * chg2
                    s.exception_par[i, 2] = scnb_local * np.sqrt(eps1 * eps2)
        else:
            pass
        return self.context

    # ---- torsion terms (reference :362-446) --------------------------------------------------------
    def add_torsion_terms(self, periodicities=[1, 2, 3, 4], phase_default=0.0, v_default=0.0):
        """
        Adds, for every distinct torsion quadruple, the periodicities of ``periodicities`` that it does not
        have yet (with ``phase_default`` and ``v_default``).  The device topology is rebuilt on next use.
        """
        self._slots_cache = None
        s = self.system
        new_idx, new_per, new_par = [], [], []
        prev_dihedral = None
        for idx, per in zip(s.torsion_idx, s.torsion_per):
            curr_dihedral = [int(a) for a in idx]
            if curr_dihedral == prev_dihedral:
                # same rule as the reference: a quadruple repeated on consecutive entries is expanded once (:398-401)
                continue
            prev_dihedral = curr_dihedral
            for n in periodicities:
                if n != int(per):
                    new_idx.append(curr_dihedral)
                    new_per.append(n)
                    new_par.append((phase_default, v_default))
        if new_idx:
            s.add_torsions(new_idx, new_per, new_par)
            self._topo = None
            self._plan = None
        return self.system

    # ------------------------------------------------------------------------------------------
    #                        compiled update plan (fast path of update_systems)
    # ------------------------------------------------------------------------------------------
    def compile_update_plan(self, force_field):
        """
        Index arrays mapping the force field's value store to slot rows, equivalent to
        ``set_bonded_parameters`` + ``set_nonbonded_parameters`` on ``force_field.force_field_optimizable``.
        Rebuilt when ``force_field.version`` changes.
        """
        key = (id(force_field), force_field.version, id(force_field.store))
        if self._plan is not None and self._plan["key"] == key:
            return self._plan
        s = self.system
        ffo = force_field.force_field_optimizable
        nb, na, nt, N = len(s.bond_idx), len(s.angle_idx), len(s.torsion_idx), s.n_atoms
        off_bond, off_angle = 0, 2 * nb
        off_tors = off_angle + 2 * na
        off_part = off_tors + 2 * nt
        off_exc = off_part + 3 * N
        dst, src = [], []
        ph_dst, ph_src = [], []

        def direct(terms_by_occurrence, base, width, keys):
            for sub in terms_by_occurrence:
                for term in sub:
                    for c, key_ in enumerate(keys):
                        dst.append(base + width * term.idx + c)
                        src.append(term.parameters[key_]._idx)

        if "HarmonicBondForce" in ffo:
            direct(ffo["HarmonicBondForce"], off_bond, 2, ("bond_eq", "bond_k"))
        if "HarmonicAngleForce" in ffo:
            direct(ffo["HarmonicAngleForce"], off_angle, 2, ("angle_eq", "angle_k"))
        if "PeriodicTorsionForce" in ffo:
            for sub in ffo["PeriodicTorsionForce"]:
                for term in sub:
                    ph_dst.append(off_tors + 2 * term.idx)
                    ph_src.append(term.parameters["torsion_phase"]._idx)
                    dst.append(off_tors + 2 * term.idx + 1)
                    src.append(term.parameters["torsion_k"]._idx)
        plan = dict(key=key, branch="C", dst=np.asarray(dst, dtype=np.int64), src=np.asarray(src, dtype=np.int64),
                    ph_dst=np.asarray(ph_dst, dtype=np.int64), ph_src=np.asarray(ph_src, dtype=np.int64),
                    off_part=off_part, off_exc=off_exc)
        # sign conventions of ForceField.update_force_field (:157-167), applied to the store columns
        abs_idx = []
        if "Scaling14" in ffo:
            for sub in ffo["Scaling14"]:
                for term in sub:
                    abs_idx += [term.parameters["scee"]._idx, term.parameters["scnb"]._idx]
        if "NonbondedForce" in ffo:
            for sub in ffo["NonbondedForce"]:
                for term in sub:
                    abs_idx += [term.parameters["lj_eps"]._idx, term.parameters["lj_sigma"]._idx]
        plan["abs_idx"] = np.asarray(abs_idx, dtype=np.int64)

        if "NonbondedForce" in ffo and "Scaling14" not in ffo:
            plan["branch"] = "A"
            terms = ffo["NonbondedForce"][0]
            plan["a_part_dst"] = np.asarray([off_part + 3 * t.idx + c for t in terms for c in range(3)], dtype=np.int64)
            plan["a_part_src"] = np.asarray([t.parameters[k_]._idx for t in terms for k_ in ("charge", "lj_sigma", "lj_eps")],
                                            dtype=np.int64)
            # exceptions read the optimizable list by POSITION = atom index (reference quirk)
            n_exc = len(s.exception_idx)
            ok = np.asarray([max(int(a), int(b)) < len(terms) for a, b in s.exception_idx], dtype=bool) if n_exc else np.zeros(0, bool)

            def col(at, key_):
                return np.asarray([terms[int(a)].parameters[key_]._idx if int(a) < len(terms) else 0 for a in at], dtype=np.int64)
            at1 = s.exception_idx[:, 0] if n_exc else np.zeros(0, int)
            at2 = s.exception_idx[:, 1] if n_exc else np.zeros(0, int)
            plan["a_ok"] = ok
            for name, key_ in (("q", "charge"), ("s", "lj_sigma"), ("e", "lj_eps")):
                plan["a_%s1" % name] = col(at1, key_)
                plan["a_%s2" % name] = col(at2, key_)
        elif "NonbondedForce" not in ffo and "Scaling14" in ffo:
            plan["branch"] = "B"
            terms = ffo["Scaling14"][0]
            n = len(terms)
            plan["b_n"] = n
            plan["b_scee"] = np.asarray([t.parameters["scee"]._idx for t in terms], dtype=np.int64)
            plan["b_scnb"] = np.asarray([t.parameters["scnb"]._idx for t in terms], dtype=np.int64)
            plan["b_at1"] = np.asarray(s.exception_idx[:n, 0], dtype=np.int64)
            plan["b_at2"] = np.asarray(s.exception_idx[:n, 1], dtype=np.int64)
        self._plan = plan
        return plan

    def slots_from_store(self, force_field, store_rows, base_slots=None):
        """
        Slot rows [P, n_slots] for ``store_rows`` [P, len(store)] (trial copies of the force field's value store,
        sign conventions already applied).  ``base_slots`` defaults to the current MM system.
        Equivalent to running the object path on each row; for branch A the "skip exceptions with
        |chargeProd| < 1e-8" test (reference :697-698) is evaluated on ``base_slots``.
        """
        plan = self.compile_update_plan(force_field)
        store_rows = np.atleast_2d(np.asarray(store_rows, dtype=np.float64))
        n_vec = store_rows.shape[0]
        base = self.current_slots() if base_slots is None else np.asarray(base_slots, dtype=np.float64)
        if n_vec == 1 and plan["branch"] not in ("A", "B"):
            # the optimiser's f(x): one row and no nonbonded rebuild -- 1-D scatters (same values, half the indexing cost)
            row = store_rows[0]
            slots = np.array(base, dtype=np.float64).reshape(-1)
            if len(plan["dst"]):
                slots[plan["dst"]] = row[plan["src"]]
            if len(plan["ph_dst"]):
                slots[plan["ph_dst"]] = self.wrap_phase(row[plan["ph_src"]])
            return slots[None, :]
        slots = np.repeat(base[None, :], n_vec, axis=0)
        if len(plan["dst"]):
            slots[:, plan["dst"]] = store_rows[:, plan["src"]]
        if len(plan["ph_dst"]):
            slots[:, plan["ph_dst"]] = self.wrap_phase(store_rows[:, plan["ph_src"]])
        off_part, off_exc = plan["off_part"], plan["off_exc"]
        if plan["branch"] == "A":
            slots[:, plan["a_part_dst"]] = store_rows[:, plan["a_part_src"]]
            exc = slots[:, off_exc:].reshape(n_vec, -1, 3)
            if exc.shape[1]:
                if not np.all(plan["a_ok"]):
                    raise IndexError("set_nonbonded_parameters: every particle must be optimizable (reference indexes the "
                                     "optimizable list by atom index)")
                upd = np.abs(base[off_exc:].reshape(-1, 3)[:, 0]) >= 1e-8
                eps = self.scee * np.sqrt(store_rows[:, plan["a_e1"]] * store_rows[:, plan["a_e2"]])
                sig = 0.5 * (store_rows[:, plan["a_s1"]] + store_rows[:, plan["a_s2"]])
                qq = self.scnb * store_rows[:, plan["a_q1"]] * store_rows[:, plan["a_q2"]]
                new = np.stack([qq, sig, eps], axis=2)
                exc[:, upd, :] = new[:, upd, :]
        elif plan["branch"] == "B":
            n = plan["b_n"]
            part = slots[:, off_part:off_exc].reshape(n_vec, -1, 3)
            exc = slots[:, off_exc:].reshape(n_vec, -1, 3)
            chg1, chg2 = part[:, plan["b_at1"], 0], part[:, plan["b_at2"], 0]
            eps1, eps2 = part[:, plan["b_at1"], 2], part[:, plan["b_at2"], 2]
            exc[:, :n, 0] = np.abs(store_rows[:, plan["b_scee"]]) * chg1 * chg2
            exc[:, :n, 2] = np.abs(store_rows[:, plan["b_scnb"]]) * np.sqrt(eps1 * eps2)
        return slots

    def slots_gradient_to_store(self, force_field, g_slots, store_row, base_slots=None):
        """
        Adjoint of :meth:`slots_from_store` at ``store_row`` (sign conventions already applied): maps d f / d slots
        [n_slots] to d f / d store [len(store)], following the same branches (direct bonded entries, wrapped phases,
        branch A exceptions rebuilt from the particle parameters, branch B exceptions scaled by scee / scnb).
        """
        plan = self.compile_update_plan(force_field)
        g_slots = np.asarray(g_slots, dtype=np.float64)
        store_row = np.asarray(store_row, dtype=np.float64).reshape(-1)
        base = self.current_slots() if base_slots is None else np.asarray(base_slots, dtype=np.float64)
        g = np.zeros_like(store_row)
        if len(plan["dst"]):
            np.add.at(g, plan["src"], g_slots[plan["dst"]])
        if len(plan["ph_dst"]):
            np.add.at(g, plan["ph_src"], g_slots[plan["ph_dst"]])   # wrap_phase has unit slope
        off_part, off_exc = plan["off_part"], plan["off_exc"]
        if plan["branch"] == "A":
            np.add.at(g, plan["a_part_src"], g_slots[plan["a_part_dst"]])
            gexc = g_slots[off_exc:].reshape(-1, 3)
            if gexc.shape[0]:
                upd = np.abs(base[off_exc:].reshape(-1, 3)[:, 0]) >= 1e-8
                q1, q2 = store_row[plan["a_q1"]], store_row[plan["a_q2"]]
                e1, e2 = store_row[plan["a_e1"]], store_row[plan["a_e2"]]
                g_qq, g_sig, g_eps = (np.where(upd, gexc[:, c], 0.0) for c in range(3))
                np.add.at(g, plan["a_q1"], g_qq * self.scnb * q2)
                np.add.at(g, plan["a_q2"], g_qq * self.scnb * q1)
                np.add.at(g, plan["a_s1"], 0.5 * g_sig)
                np.add.at(g, plan["a_s2"], 0.5 * g_sig)
                with np.errstate(divide="ignore", invalid="ignore"):
                    d1 = np.where(e1 > 0, 0.5 * self.scee * np.sqrt(e2) / np.sqrt(e1), 0.0)
                    d2 = np.where(e2 > 0, 0.5 * self.scee * np.sqrt(e1) / np.sqrt(e2), 0.0)
                np.add.at(g, plan["a_e1"], g_eps * d1)
                np.add.at(g, plan["a_e2"], g_eps * d2)
        elif plan["branch"] == "B":
            n = plan["b_n"]
            part = base[off_part:off_exc].reshape(-1, 3)
            gexc = g_slots[off_exc:].reshape(-1, 3)
            chg1, chg2 = part[plan["b_at1"], 0], part[plan["b_at2"], 0]
            eps1, eps2 = part[plan["b_at1"], 2], part[plan["b_at2"], 2]
            np.add.at(g, plan["b_scee"], gexc[:n, 0] * chg1 * chg2)
            np.add.at(g, plan["b_scnb"], gexc[:n, 2] * np.sqrt(eps1 * eps2))
        return g

    def load_slots(self, slots_row):
        """Writes one slots row back into the MM system arrays (keeps ``self.system`` in step with the fast path)."""
        s = self.system
        row = np.array(slots_row, dtype=np.float64).reshape(-1)   # one private copy: source of the writes and the new cache
        o = 0
        for arr in (s.bond_par, s.angle_par, s.torsion_par, s.particle_par, s.exception_par):
            n = arr.size
            arr[...] = row[o:o + n].reshape(arr.shape)
            o += n
        self._slots_cache = row

    # ------------------------------------------------------------------------------------------
    #                                       PRIVATE METHODS
    # ------------------------------------------------------------------------------------------
    def _set_force_groups(self):
        """force name -> list of force indices, in system order (reference :849-878)."""
        self.force_groups = []
        self.forces_indexes = {}
        for force_idx, force_key in enumerate(self.system.force_names):
            self.force_groups.append(self.force_groups_dict[force_key])
            self.forces_indexes.setdefault(force_key, []).append(force_idx)
        return self.force_groups
