# -*- coding: utf-8 -*-
"""
Description
-----------
Deterministic synthetic workloads for the benchmark configurations of ``BASELINE.json`` that have no
in-tree molecule: a drug-like ligand of N atoms ("lig50", "lig60"; the largest molecule shipped with the
reference is norfloxacin, N=41, ``Examples/Example_5/norfloxacin.prmtop``) and perturbed-conformation
ensembles.

The ligand is a branched heavy-atom skeleton (about N/2 heavy atoms, bond length 0.14-0.153 nm) saturated
with hydrogens, with every proper torsion present (some bonds carry two periodicities, as GAFF does) and
AMBER-style exceptions: 1-4 pairs first (Coulomb / 1.2, LJ / 2), then the 1-2 / 1-3 exclusions -- the
order OpenMM's AMBER reader emits and ``OpenMMEngine.set_nonbonded_parameters`` relies on
(``ParaMol/MM_engines/openmm.py:731-748``).  Equilibrium lengths and angles are the generated geometry, so
the reference structure is close to a minimum and no minimiser is needed; parameter magnitudes follow the
ranges seen in the in-tree GAFF topologies.
"""
import numpy as np

from ..MM_engines.mm_system import MMSystem


def _unit(v):
    return v / np.linalg.norm(v)


def make_ligand(n_atoms=60, seed=7):
    """Returns (MMSystem, x0 [N,3] nm)."""
    rng = np.random.default_rng(seed)
    n_heavy = max(2, n_atoms // 2)
    pos = [np.zeros(3)]
    nbrs = [[]]
    max_val = [4]

    def ok(p, parent, min_d, min_angle=100.0):
        for k, q in enumerate(pos):
            if k == parent:
                continue
            if np.linalg.norm(p - q) < min_d:
                return False
        for k in nbrs[parent]:
            a = _unit(pos[k] - pos[parent])
            b = _unit(p - pos[parent])
            if np.dot(a, b) > np.cos(np.deg2rad(min_angle)):
                return False
        return True

    def add_atom(parent, p, val):
        pos.append(p)
        nbrs.append([parent])
        nbrs[parent].append(len(pos) - 1)
        max_val.append(val)

    def try_ring(parent):
        """Attach a planar six-membered ring (side 0.14 nm) to ``parent``; returns True on success."""
        for _ in range(200):
            d = _unit(rng.normal(size=3))
            r0 = pos[parent] + 0.148 * d
            if not ok(r0, parent, 0.235):
                continue
            u = rng.normal(size=3)
            u = _unit(u - np.dot(u, d) * d)
            c = r0 + 0.14 * d
            ring = [c + 0.14 * (np.cos(np.pi + k * np.pi / 3) * d + np.sin(np.pi + k * np.pi / 3) * u) for k in range(6)]
            if any(np.linalg.norm(r - q) < 0.235 for r in ring[1:] for q in pos):
                continue
            first = len(pos)
            add_atom(parent, ring[0], 3)
            for k in range(1, 6):
                add_atom(first + k - 1, ring[k], 3)
            nbrs[first].append(first + 5)
            nbrs[first + 5].append(first)
            return True
        return False

    def grow(n_target, bond_len, min_d, heavy):
        stall = 0
        while len(pos) < n_target:
            stall += 1
            if stall > 400000:
                raise RuntimeError("synthetic ligand generator stalled")
            relax = 0.97 ** (stall // 2000)  # crowded sites: accept slightly closer contacts rather than stall
            if heavy:
                # favour recently added atoms: chain-like growth with branches and rings
                cand = [k for k in range(len(pos)) if len(nbrs[k]) < max_val[k] - 1]
                if not cand:
                    cand = [k for k in range(len(pos)) if len(nbrs[k]) < max_val[k]]
                w = np.array([1.0 + 3.0 * (k >= len(pos) - 4) for k in cand])
                parent = int(rng.choice(cand, p=w / w.sum()))
                if n_target - len(pos) >= 6 and rng.uniform() < 0.3:
                    if try_ring(parent):
                        stall = 0
                    continue
            else:
                cand = [k for k in range(n_heavy) if len(nbrs[k]) < max_val[k]]
                if not cand:
                    cand = [k for k in range(n_heavy) if max_val[k] == 3 and len(nbrs[k]) < 4]
                parent = int(rng.choice(cand))
            d = _unit(rng.normal(size=3))
            p = pos[parent] + bond_len * (1.0 + 0.03 * rng.uniform(-1, 1)) * d
            if not ok(p, parent, max(0.12, min_d * relax), max(75.0, 100.0 * relax)):
                continue
            add_atom(parent, p, 4 if heavy else 1)
            stall = 0

    max_val[0] = 4
    grow(n_heavy, 0.148, 0.235, True)
    grow(n_atoms, 0.109, 0.165, False)
    pos = np.asarray(pos)
    is_h = np.arange(n_atoms) >= n_heavy

    # AMBER-like order: each heavy atom followed by its hydrogens
    order = []
    for k in range(n_heavy):
        order.append(k)
        order.extend(sorted(j for j in nbrs[k] if is_h[j]))
    order = np.asarray(order)
    inv = np.empty(n_atoms, dtype=int)
    inv[order] = np.arange(n_atoms)
    pos = pos[order]
    is_h = is_h[order]
    adj = [sorted(int(inv[j]) for j in nbrs[int(order[i])]) for i in range(n_atoms)]

    system = MMSystem(n_atoms)
    system.masses = np.where(is_h, 1.008, 12.01)
    system.atomic_numbers = np.where(is_h, 1, 6).astype(np.int32)
    system.atom_names = ["H%d" % i if h else "C%d" % i for i, h in enumerate(is_h)]

    bonds = [(i, j) for i in range(n_atoms) for j in adj[i] if i < j]
    # hydrogen-containing terms first, like %FLAG BONDS_INC_HYDROGEN / _WITHOUT_HYDROGEN
    bonds.sort(key=lambda b: (not (is_h[b[0]] or is_h[b[1]]), b))
    bond_par = []
    for i, j in bonds:
        r = np.linalg.norm(pos[i] - pos[j])
        k = rng.uniform(2.8e5, 3.1e5) if (is_h[i] or is_h[j]) else rng.uniform(2.2e5, 4.0e5)
        bond_par.append((r * (1.0 + 0.004 * rng.uniform(-1, 1)), k))
    angles = []
    for j in range(n_atoms):
        for a in range(len(adj[j])):
            for b in range(a + 1, len(adj[j])):
                angles.append((adj[j][a], j, adj[j][b]))
    angles.sort(key=lambda t: (not (is_h[t[0]] or is_h[t[2]]), t))
    angle_par = []
    for i, j, k in angles:
        v1, v2 = pos[i] - pos[j], pos[k] - pos[j]
        th = np.arccos(np.clip(np.dot(_unit(v1), _unit(v2)), -1, 1))
        angle_par.append((th + 0.02 * rng.uniform(-1, 1), rng.uniform(300.0, 600.0)))
    torsions, tper, tpar = [], [], []
    one_four = []
    for (j, k) in bonds:
        for i in adj[j]:
            if i == k:
                continue
            for l in adj[k]:
                if l == j or l == i:
                    continue
                n_terms = 2 if rng.uniform() < 0.25 else 1
                pers = rng.choice([1, 2, 3], size=n_terms, replace=False)
                for n in pers:
                    torsions.append((i, j, k, l))
                    tper.append(int(n))
                    tpar.append((float(rng.choice([0.0, np.pi])), float(rng.uniform(0.0, 8.0)) / (1 + (n_terms > 1))))
                pair = (min(i, l), max(i, l))
                if pair not in one_four:
                    one_four.append(pair)
    q = rng.normal(0.0, 0.25, n_atoms)
    q -= q.mean()
    sigma = np.where(is_h, rng.uniform(0.19, 0.27, n_atoms), rng.uniform(0.30, 0.34, n_atoms))
    eps = np.where(is_h, rng.uniform(0.06, 0.07, n_atoms), rng.uniform(0.36, 0.46, n_atoms))
    system.particle_par = np.stack([q, sigma, eps], axis=1)

    excluded = set()
    for i in range(n_atoms):
        for j in adj[i]:
            excluded.add((min(i, j), max(i, j)))
            for k in adj[j]:
                if k != i:
                    excluded.add((min(i, k), max(i, k)))
    one_four = [p for p in one_four if p not in excluded]
    exc_idx, exc_par = [], []
    for i, l in one_four:
        exc_idx.append((i, l))
        exc_par.append((q[i] * q[l] / 1.2, 0.5 * (sigma[i] + sigma[l]), np.sqrt(eps[i] * eps[l]) / 2.0))
    for i, j in sorted(excluded):
        exc_idx.append((i, j))
        exc_par.append((0.0, 0.1, 0.0))

    system.bond_idx = np.asarray(bonds, dtype=np.int32).reshape(-1, 2)
    system.bond_par = np.asarray(bond_par, dtype=np.float64).reshape(-1, 2)
    system.angle_idx = np.asarray(angles, dtype=np.int32).reshape(-1, 3)
    system.angle_par = np.asarray(angle_par, dtype=np.float64).reshape(-1, 2)
    system.torsion_idx = np.asarray(torsions, dtype=np.int32).reshape(-1, 4)
    system.torsion_per = np.asarray(tper, dtype=np.int32)
    system.torsion_par = np.asarray(tpar, dtype=np.float64).reshape(-1, 2)
    system.exception_idx = np.asarray(exc_idx, dtype=np.int32).reshape(-1, 2)
    system.exception_par = np.asarray(exc_par, dtype=np.float64).reshape(-1, 3)
    return system, pos - pos.mean(axis=0)


def perturbed_conformations(x_base, n_conf, sigma=0.005, seed=1234):
    """coords_s = X[s mod len(X)] + N(0, sigma) per coordinate; float64 [S,N,3]."""
    x_base = np.asarray(x_base, dtype=np.float64)
    if x_base.ndim == 2:
        x_base = x_base[None]
    rng = np.random.default_rng(seed)
    idx = np.arange(n_conf) % x_base.shape[0]
    return x_base[idx] + rng.normal(0.0, sigma, (n_conf,) + x_base.shape[1:])
