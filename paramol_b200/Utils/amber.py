# -*- coding: utf-8 -*-
"""
Description
-----------
AMBER ``prmtop`` / ``inpcrd`` ingest producing an :obj:`paramol_b200.MM_engines.mm_system.MMSystem`.

In the reference this step is the third-party call
``app.AmberPrmtopFile(top_file).createSystem(nonbondedMethod=NoCutoff, constraints=None, ...)``
(``ParaMol/MM_engines/openmm.py:198,232``; pinned ``openmm=7.7.0``,
``.github/workflows/paramol_env.yml:109``).  OpenMM is not vendored in the reference tree, so this
module restates the published conventions of OpenMM's AMBER reader
(``openmm/app/internal/amber_file_parser.py``: ``PrmtopLoader`` + ``readAmberSystem``):

* charge  q = CHARGE / 18.2223
* per-atom LJ from the diagonal A/B coefficients of the atom's type:
  ``rMin = (2A/B)**(1/6)``, ``eps = 0.25*B*B/A``; ``rVdw = rMin/2 * 0.1`` nm, ``eps *= 4.184``;
  ``sigma = rVdw * (2**(-1/6) * 2)``
* bonds: INC_HYDROGEN then WITHOUT_HYDROGEN, atom = pointer // 3, ``k = 2 * (K * 418.4)``,
  ``r0 = R * 0.1``; angles likewise with ``k = 2 * (K * 4.184)``
* torsions: INC_HYDROGEN then WITHOUT_HYDROGEN, atoms = |pointer| // 3,
  ``periodicity = int(0.5 + PERIODICITY)``, ``phase``, ``k = PK * 4.184``
* 1-4 exceptions: every dihedral whose 3rd and 4th pointers are positive contributes the pair (i,l)
  with ``chargeProd = q_i q_l / SCEE``, ``sigma = (rVdw_i + rVdw_l) * 2**(-1/6)``,
  ``eps = sqrt(eps_i eps_l) / SCNB``
* then, per atom, every excluded atom not already a 1-4 pair becomes the exception (0, 0.1, 0).

The operation order above reproduces the 17-digit parameter values printed in the reference's own
golden vector (``ParaMol/Tests/Objective_function/test_objective_function.py:49-76``), see
``tests/test_ingest.py``.
"""
import math
import re
import numpy as np

from ..MM_engines.mm_system import MMSystem

_FORMAT_RE = re.compile(r"%FORMAT\(\s*(\d*)\s*([aAiIeEfF])\s*(\d+)(?:\.(\d+))?\s*\)")


def read_prmtop_raw(file_name):
    """Parse a prmtop into {FLAG: list of str/int/float}."""
    data = {}
    flag = None
    kind = None
    width = None
    with open(file_name, "r") as f:
        for line in f:
            if line.startswith("%VERSION") or line.startswith("%COMMENT"):
                continue
            if line.startswith("%FLAG"):
                flag = line.split()[1]
                data[flag] = []
                continue
            if line.startswith("%FORMAT"):
                m = _FORMAT_RE.match(line.strip())
                if m is None:
                    raise ValueError("unparsable prmtop format line: %r" % line)
                kind = m.group(2).lower()
                width = int(m.group(3))
                continue
            if flag is None:
                continue
            line = line.rstrip("\n")
            for i in range(0, len(line), width):
                field = line[i:i + width]
                if kind == "a":
                    if flag == "TITLE":
                        data[flag].append(field)
                    elif field.strip() or i + width <= len(line):
                        data[flag].append(field.strip())
                else:
                    if not field.strip():
                        continue
                    data[flag].append(int(field) if kind == "i" else float(field))
    return data


def _grouped(values, n):
    return [values[i:i + n] for i in range(0, len(values) - n + 1, n)]


def system_from_prmtop(file_name, nonbonded_method="NoCutoff", cutoff=1.2, rf_dielectric=78.3):
    """
    Build the MMSystem an ``AmberPrmtopFile.createSystem(nonbondedMethod=NoCutoff, constraints=None)``
    call would describe.
    """
    raw = read_prmtop_raw(file_name)
    ptr = raw["POINTERS"]
    n_atoms = int(ptr[0])
    n_types = int(ptr[1])

    length_conv = 0.1                 # angstrom -> nm
    energy_conv = 4.184               # kcal/mol -> kJ/mol
    bond_k_conv = 4.184 * 100         # kcal/mol/A^2 -> kJ/mol/nm^2 (= 418.40000000000003)

    system = MMSystem(n_atoms)
    system.masses = np.asarray(raw["MASS"], dtype=np.float64)
    if "ATOMIC_NUMBER" in raw:
        system.atomic_numbers = np.asarray(raw["ATOMIC_NUMBER"], dtype=np.int32)
    system.atom_names = list(raw.get("ATOM_NAME", [""] * n_atoms))[:n_atoms]

    charges = [float(x) / 18.2223 for x in raw["CHARGE"]]

    # --- per-atom Lennard-Jones terms (PrmtopLoader.getNonbondTerms) ---------------------------
    type_idx = raw["ATOM_TYPE_INDEX"]
    nb_parm_index = raw["NONBONDED_PARM_INDEX"]
    acoef = raw["LENNARD_JONES_ACOEF"]
    bcoef = raw["LENNARD_JONES_BCOEF"]
    rvdw = []
    eps = []
    for i in range(n_atoms):
        index = (n_types + 1) * (int(type_idx[i]) - 1)
        nb_index = int(nb_parm_index[index]) - 1
        if nb_index < 0:
            raise ValueError("10-12 interactions are not supported")
        a = float(acoef[nb_index])
        b = float(bcoef[nb_index])
        try:
            r_min = (2 * a / b) ** (1 / 6.0)
            e = 0.25 * b * b / a
        except ZeroDivisionError:
            r_min = 1.0
            e = 0.0
        rvdw.append(r_min / 2.0 * length_conv)
        eps.append(e * energy_conv)

    sigma_scale = 2 ** (-1. / 6.) * 2.0
    system.particle_par = np.asarray([[charges[i], rvdw[i] * sigma_scale, eps[i]] for i in range(n_atoms)],
                                     dtype=np.float64).reshape(n_atoms, 3)

    # --- bonds --------------------------------------------------------------------------------
    bk = raw["BOND_FORCE_CONSTANT"]
    br = raw["BOND_EQUIL_VALUE"]
    bonds = _grouped(raw.get("BONDS_INC_HYDROGEN", []), 3) + _grouped(raw.get("BONDS_WITHOUT_HYDROGEN", []), 3)
    system.bond_idx = np.asarray([[a // 3, b // 3] for a, b, t in bonds], dtype=np.int32).reshape(-1, 2)
    system.bond_par = np.asarray([[float(br[t - 1]) * length_conv, 2 * (float(bk[t - 1]) * bond_k_conv)]
                                  for a, b, t in bonds], dtype=np.float64).reshape(-1, 2)

    # --- angles -------------------------------------------------------------------------------
    ak = raw["ANGLE_FORCE_CONSTANT"]
    at = raw["ANGLE_EQUIL_VALUE"]
    angles = _grouped(raw.get("ANGLES_INC_HYDROGEN", []), 4) + _grouped(raw.get("ANGLES_WITHOUT_HYDROGEN", []), 4)
    system.angle_idx = np.asarray([[a // 3, b // 3, c // 3] for a, b, c, t in angles], dtype=np.int32).reshape(-1, 3)
    system.angle_par = np.asarray([[float(at[t - 1]), 2 * (float(ak[t - 1]) * energy_conv)]
                                   for a, b, c, t in angles], dtype=np.float64).reshape(-1, 2)

    # --- torsions -----------------------------------------------------------------------------
    dk = raw["DIHEDRAL_FORCE_CONSTANT"]
    dph = raw["DIHEDRAL_PHASE"]
    dper = raw["DIHEDRAL_PERIODICITY"]
    dihedrals = _grouped(raw.get("DIHEDRALS_INC_HYDROGEN", []), 5) + _grouped(raw.get("DIHEDRALS_WITHOUT_HYDROGEN", []), 5)
    system.torsion_idx = np.asarray([[a // 3, b // 3, abs(c) // 3, abs(d) // 3] for a, b, c, d, t in dihedrals],
                                    dtype=np.int32).reshape(-1, 4)
    system.torsion_per = np.asarray([int(0.5 + float(dper[t - 1])) for a, b, c, d, t in dihedrals],
                                    dtype=np.int32).reshape(-1)
    system.torsion_par = np.asarray([[float(dph[t - 1]), float(dk[t - 1]) * energy_conv]
                                     for a, b, c, d, t in dihedrals], dtype=np.float64).reshape(-1, 2)

    # --- 1-4 exceptions, then plain exclusions -------------------------------------------------
    scee = raw.get("SCEE_SCALE_FACTOR")
    scnb = raw.get("SCNB_SCALE_FACTOR")
    sigma_scale14 = 2 ** (-1. / 6.)
    exc_idx = []
    exc_par = []
    excluded_pairs = set()
    for a, b, c, d, t in dihedrals:
        if int(c) > 0 and int(d) > 0:
            i = int(a) // 3
            l = int(d) // 3
            charge_prod = charges[i] * charges[l]
            r_min = rvdw[i] + rvdw[l]
            epsilon = math.sqrt(eps[i] * eps[l])
            i_scee = float(scee[t - 1]) if scee is not None else 1.2
            i_scnb = float(scnb[t - 1]) if scnb is not None else 2.0
            charge_prod /= i_scee
            epsilon /= i_scnb
            exc_idx.append([i, l])
            exc_par.append([charge_prod, r_min * sigma_scale14, epsilon])
            excluded_pairs.add((min(i, l), max(i, l)))

    n_excl = raw["NUMBER_EXCLUDED_ATOMS"]
    excl_list = raw["EXCLUDED_ATOMS_LIST"]
    total = 0
    for i in range(n_atoms):
        n = int(n_excl[i])
        for j_raw in excl_list[total:total + n]:
            j = int(j_raw)
            if j > 0:
                j -= 1
                key = (min(i, j), max(i, j))
                if key in excluded_pairs:
                    continue
                exc_idx.append([i, j])
                exc_par.append([0.0, 0.1, 0.0])
        total += n

    system.exception_idx = np.asarray(exc_idx, dtype=np.int32).reshape(-1, 2)
    system.exception_par = np.asarray(exc_par, dtype=np.float64).reshape(-1, 3)

    system.nonbonded_method = nonbonded_method
    system.cutoff = float(cutoff)
    system.rf_dielectric = float(rf_dielectric)
    return system


def read_inpcrd(file_name):
    """
    Read AMBER ASCII restart/inpcrd coordinates.  Returns positions in nm, float64 [N,3]
    (``app.AmberInpcrdFile(...).positions`` in the reference, ``openmm.py:211-212``).
    """
    with open(file_name, "r") as f:
        lines = f.read().split("\n")
    n_atoms = int(lines[1].split()[0])
    values = []
    for line in lines[2:]:
        for i in range(0, len(line.rstrip()), 12):
            field = line[i:i + 12]
            if field.strip():
                values.append(float(field))
        if len(values) >= 3 * n_atoms:
            break
    return np.asarray(values[:3 * n_atoms], dtype=np.float64).reshape(n_atoms, 3) * 0.1
