# -*- coding: utf-8 -*-
"""
Description
-----------
Thin Python layer over the C ABI (``include/paramol_b200.h``): :obj:`DeviceTopology` owns one ``pm_handle``
and :obj:`DeviceEnsemble` keeps the conformations of one ParaMol system resident in HBM (tile layout) and
launches the evaluation / reduction kernels on them.  PyTorch is used only for device memory and streams.

This is the layer that replaces the ``context.setPositions -> context.getState`` loops of
``ParaMol/System/system.py:324-354`` and ``ParaMol/Objective_function/cpu_objective_function.py:67-108``.
"""
import ctypes
import os

import numpy as np

from .. import _lib
from .._lib import B200Error, PmTopology, c_int_p, c_double_p

NPARTIAL = 8


def _torch():
    import torch
    return torch


def _iptr(a):
    return a.ctypes.data_as(c_int_p)


def _stream_ptr(device):
    """cudaStream_t (as void*) of torch's current stream on ``device``."""
    torch = _torch()
    index = device.index if hasattr(device, "index") else int(device)
    if index is None:
        index = torch.cuda.current_device()
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return ctypes.c_void_p(raw(index))
    return ctypes.c_void_p(torch.cuda.current_stream(index).cuda_stream)


def _pm_topology(mm_system):
    """(PmTopology, arrays it points into) for an :obj:`MMSystem`.  Exceptions whose initial chargeProd and epsilon are both
    zero are pure exclusions; every other exception is a scaled 1-4 interaction."""
    s = mm_system
    keep = dict(
        bond=np.ascontiguousarray(s.bond_idx, dtype=np.int32).reshape(-1, 2),
        angle=np.ascontiguousarray(s.angle_idx, dtype=np.int32).reshape(-1, 3),
        tors=np.ascontiguousarray(s.torsion_idx, dtype=np.int32).reshape(-1, 4),
        tper=np.ascontiguousarray(s.torsion_per, dtype=np.int32).reshape(-1),
        exc=np.ascontiguousarray(s.exception_idx, dtype=np.int32).reshape(-1, 2),
    )
    par = np.asarray(s.exception_par, dtype=np.float64).reshape(-1, 3)
    keep["act"] = np.ascontiguousarray(((par[:, 0] != 0.0) | (par[:, 2] != 0.0)).astype(np.int32))
    k = keep
    topo = PmTopology(int(s.n_atoms),
                      len(k["bond"]), _iptr(k["bond"]),
                      len(k["angle"]), _iptr(k["angle"]),
                      len(k["tors"]), _iptr(k["tors"]), _iptr(k["tper"]),
                      len(k["exc"]), _iptr(k["exc"]), _iptr(k["act"]),
                      0 if s.nonbonded_method == "NoCutoff" else 1,
                      float(s.cutoff), float(s.rf_dielectric), float(s.one_4pi_eps0))
    return topo, keep


def prebuild_specialized(mm_system):
    """Generate and compile (into the in-tree cache) the topology-specialised kernel of ``mm_system`` WITHOUT a GPU: the pair
    list comes from the host-only ``pm_topology_pairs``.  Returns the image size in bytes, or 0 when the molecule is not
    eligible.  ``DeviceTopology`` finds the cached image later by the hash of the same source."""
    topo, _keep = _pm_topology(mm_system)
    built = DeviceTopology._build_specialized(mm_system, topo)
    return 0 if built is None else len(built[0])


class DeviceTopology:
    """
    One ``pm_handle``: the topology "program" of a molecule on one GPU.

    Parameters
    ----------
    mm_system : :obj:`paramol_b200.MM_engines.mm_system.MMSystem`
        Term lists in OpenMM-System order.  Exceptions whose initial chargeProd and epsilon are both zero are
        pure exclusions (OpenMM decides this once, when the Context is created; see the note in
        ``ParaMol/MM_engines/openmm.py:622-626``); every other exception is a scaled 1-4 interaction.
    device : int or None
        CUDA device index (default: current device).
    conf_per_tile : int or None
        Conformations per tile (32, 16, 8, 4: teams of 32 / conf_per_tile lanes per conformation), -8 (tiles of 32 conformations
        shared by teams of 8 warps), or None to let the library choose (``pm_create(..., conf_per_tile=0, ...)``).

    ``specialize_status`` says which E+F kernel serves full evaluations: "installed" (generated for this topology), "not eligible"
    (molecule too large / reaction field / switched off) or the reason the generated kernel could not be built or installed.
    ``PARAMOL_B200_SPECIALIZE=require`` turns the last case into a :obj:`B200Error` (``bench.py`` sets it).
    """

    def __init__(self, mm_system, device=None, conf_per_tile=None):
        torch = _torch()
        L = _lib.lib()
        if not torch.cuda.is_available() or L.pm_device_count() < 1:
            raise B200Error("no CUDA device is visible: the B200 path has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        s = mm_system
        self.n_atoms = int(s.n_atoms)
        self.k_coul = float(s.one_4pi_eps0)
        #: CutoffNonPeriodic + reaction field (the system of the reference's Tests/aniline.xml): served by the generic kernel
        #: only; pm_delta_eval and pm_grad do not implement it
        self.rf = s.nonbonded_method != "NoCutoff"
        topo, self._keep = _pm_topology(s)
        self.exception_active = self._keep["act"]
        k = self._keep
        # Small molecules: a kernel generated for this topology (csrc/specialize.py).  It works on tiles of 32 conformations, so
        # the image is built first (host-only pair list) and, if there is one, the handle is created with that tile shape.
        self.specialize_status = "not eligible"
        built = self._build_specialized(s, topo, status=self)
        h = ctypes.c_void_p()
        # conformations per tile: 32 for the generated kernels, else the caller's choice (``conf_per_tile`` argument or, for
        # sweeps only, the PM_CPW environment variable), else 0 = chosen by the library per topology
        cpw = 32 if built is not None else int(conf_per_tile if conf_per_tile is not None else (os.environ.get("PM_CPW") or 0))
        _lib.check(L.pm_create(ctypes.byref(topo), self.device, cpw, ctypes.byref(h)), "pm_create")
        self._h = h
        self._L = L
        self.n_bonds, self.n_angles, self.n_torsions, self.n_exceptions = (len(k["bond"]), len(k["angle"]),
                                                                           len(k["tors"]), len(k["exc"]))
        self.n_slots = int(L.pm_n_slots(h))
        self.n_derived = int(L.pm_n_derived(h))
        self.n_pairs = int(L.pm_n_pairs(h))
        self.tile_doubles = int(L.pm_tile_doubles(h))
        self.conf_per_tile = int(L.pm_conf_per_tile(h))
        # slot offsets (include/paramol_b200.h)
        self.off_bond = 0
        self.off_angle = self.off_bond + 2 * self.n_bonds
        self.off_torsion = self.off_angle + 2 * self.n_angles
        self.off_particle = self.off_torsion + 2 * self.n_torsions
        self.off_exception = self.off_particle + 3 * self.n_atoms
        assert self.off_exception + 3 * self.n_exceptions == self.n_slots
        self.specialized = False
        if built is not None and self.conf_per_tile == 32:
            image, warps, tiles, warps_energy, warps_grad = built
            buf = ctypes.create_string_buffer(image, len(image))
            if self._L.pm_set_specialized(self._h, buf, len(image), warps, tiles, warps_energy, warps_grad) == 0:
                self.specialized = True
                self.specialize_status = "installed"
            else:   # the generic kernel of the same handle keeps working
                self.specialize_status = "not installed: " + self._L.pm_last_error().decode()
                if os.environ.get("PARAMOL_B200_SPECIALIZE", "auto") == "require":
                    raise B200Error("topology-specialised kernel " + self.specialize_status)
                import logging
                logging.warning("paramol_b200: specialised kernel %s", self.specialize_status)
        # every rank must run the same kernel (same tile shape, bitwise reproducible partial sums): a rank whose compile failed
        # while the others succeeded is an error, not a silent per-rank difference
        from ..Utils import dist as _dist
        if _dist.world_size() > 1:
            flags = _dist.allreduce_sum_numpy(np.array([1.0 if self.specialized else 0.0, float(self.conf_per_tile)]))
            if flags[0] not in (0.0, float(_dist.world_size())) or flags[1] != self.conf_per_tile * _dist.world_size():
                raise B200Error("ranks disagree on the E+F kernel of this topology (specialised on %d of %d ranks; this rank: %s)"
                                % (int(flags[0]), _dist.world_size(), self.specialize_status))

    @staticmethod
    def _build_specialized(mm_system, topo, status=None):
        """
        (image, warps, atoms in the force tile, energy warps, gradient warps) of the topology-specialised kernels of ``mm_system`` or None.  Controlled by
        ``PARAMOL_B200_SPECIALIZE`` ("0" off, anything else / unset: automatic).  If the compiler is not available the generic
        kernel stays in place -- both are the CUDA path; nothing falls back to the CPU.
        """
        if os.environ.get("PARAMOL_B200_SPECIALIZE", "auto") == "0" or mm_system.nonbonded_method != "NoCutoff":
            return None
        from ..csrc import specialize
        n_atoms = int(mm_system.n_atoms)
        if not specialize.eligible(n_atoms) and not os.environ.get("PM_SPEC_FREG", "").startswith("top"):
            return None
        L = _lib.lib()
        n = ctypes.c_int()
        _lib.check(L.pm_topology_pairs(ctypes.byref(topo), 32, ctypes.byref(n), None), "pm_topology_pairs")
        pairs = np.zeros((max(n.value, 1), 3), dtype=np.int32)
        _lib.check(L.pm_topology_pairs(ctypes.byref(topo), 32, ctypes.byref(n), _iptr(pairs)), "pm_topology_pairs")
        k = dict(bond=np.ascontiguousarray(mm_system.bond_idx, dtype=np.int32).reshape(-1, 2),
                 angle=np.ascontiguousarray(mm_system.angle_idx, dtype=np.int32).reshape(-1, 3),
                 tors=np.ascontiguousarray(mm_system.torsion_idx, dtype=np.int32).reshape(-1, 4),
                 tper=np.ascontiguousarray(mm_system.torsion_per, dtype=np.int32).reshape(-1))
        try:
            return specialize.image_for(n_atoms, k["bond"], k["angle"], k["tors"], k["tper"], pairs[:n.value])
        except (RuntimeError, OSError) as exc:   # no compiler, compile error, cache directory not writable
            if status is not None:
                status.specialize_status = "not built: %s" % (str(exc).splitlines()[0] if str(exc) else type(exc).__name__)
            if os.environ.get("PARAMOL_B200_SPECIALIZE", "auto") == "require":
                raise B200Error("topology-specialised kernel could not be built: %s" % exc)
            import logging
            logging.warning("paramol_b200: topology-specialised kernel not built (%s); using the generic kernel", exc)
            return None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.pm_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def slots_from_system(self, mm_system):
        """Flatten the physical parameter table of ``mm_system`` into one slots row (float64 [n_slots])."""
        s = mm_system
        return np.concatenate([np.asarray(s.bond_par, dtype=np.float64).ravel(),
                               np.asarray(s.angle_par, dtype=np.float64).ravel(),
                               np.asarray(s.torsion_par, dtype=np.float64).ravel(),
                               np.asarray(s.particle_par, dtype=np.float64).ravel(),
                               np.asarray(s.exception_par, dtype=np.float64).ravel()])

    def pairs(self):
        """Interacting pair list as the kernel enumerates it: int32 [n_pairs, 3] = (i, j, exception or -1)."""
        out = np.zeros((self.n_pairs, 3), dtype=np.int32)
        _lib.check(self._L.pm_get_pairs(self._h, _iptr(out)), "pm_get_pairs")
        return out

    def launch_info(self, with_forces=True):
        g, b, s, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(self._L.pm_launch_info(self._h, int(with_forces), ctypes.byref(g), ctypes.byref(b), ctypes.byref(s),
                                          ctypes.byref(w)), "pm_launch_info")
        return dict(grid=g.value, block=b.value, smem_bytes=s.value, warps_per_cta=w.value)

    def set_force_metric(self, metric):
        m = np.ascontiguousarray(metric, dtype=np.float64).reshape(self.n_atoms, 9)
        _lib.check(self._L.pm_set_force_metric(self._h, m.ctypes.data_as(c_double_p)), "pm_set_force_metric")

    def n_tiles(self, n_conf):
        return (int(n_conf) + self.conf_per_tile - 1) // self.conf_per_tile

    def program_info(self):
        out = np.zeros(10, dtype=np.int32)
        _lib.check(self._L.pm_program_info(self._h, _iptr(out)), "pm_program_info")
        keys = ("conf_per_tile", "lanes_per_conformation", "stars", "star_steps", "axes", "axis_steps", "pair_segments",
                "pair_entries", "pair_entry_slots", "program_ints")
        return dict(zip(keys, (int(v) for v in out)))

    # ---- raw launches on torch tensors ---------------------------------------------------------
    def pack(self, src, n_conf):
        """[S,N,3] float64 CUDA tensor -> tile tensor."""
        torch = _torch()
        assert src.is_cuda and src.dtype == torch.float64 and src.is_contiguous()
        tiles = torch.empty(self.n_tiles(n_conf) * self.tile_doubles, dtype=torch.float64, device=src.device)
        _lib.check(self._L.pm_pack_tiles(self._h, src.data_ptr(), n_conf, tiles.data_ptr(), _stream_ptr(src.device)),
                   "pm_pack_tiles")
        return tiles

    def unpack(self, tiles, n_conf):
        torch = _torch()
        dst = torch.empty((n_conf, self.n_atoms, 3), dtype=torch.float64, device=tiles.device)
        _lib.check(self._L.pm_unpack_tiles(self._h, tiles.data_ptr(), n_conf, dst.data_ptr(), _stream_ptr(tiles.device)),
                   "pm_unpack_tiles")
        return dst

    def term_gradient_to_slots(self, gout, slots_row):
        """
        Chain the kernel's term gradient ``gout`` (numpy [pm_grad_out_doubles]: bonded entries are slot derivatives already,
        pair entries are d/d(c12, c6, K qq)) back to the slot layout: the adjoint of ``pm_prepare_params`` with
        c12 = 4 eps sigma^12, c6 = 4 eps sigma^6, plain pairs combined as qq = q_i q_j, sigma = (s_i + s_j)/2,
        eps = sqrt(e_i e_j) and exceptions carrying their own (qq, sigma, eps).  Returns numpy [n_slots].
        """
        gout = np.asarray(gout, dtype=np.float64)
        slots_row = np.asarray(slots_row, dtype=np.float64)
        g = np.zeros(self.n_slots)
        nbond = 2 * (self.n_bonds + self.n_angles + self.n_torsions)
        g[:nbond] = gout[:nbond]
        if self.n_pairs == 0:
            return g
        if getattr(self, "_pairs_cache", None) is None:
            self._pairs_cache = self.pairs()
        pm = self._pairs_cache
        gp = gout[nbond:nbond + 3 * self.n_pairs].reshape(-1, 3)
        part = slots_row[self.off_particle:self.off_exception].reshape(-1, 3)
        exc = slots_row[self.off_exception:].reshape(-1, 3)
        i, j, e = pm[:, 0], pm[:, 1], pm[:, 2]
        plain = e < 0
        sig = np.where(plain, 0.5 * part[i, 1] + 0.5 * part[j, 1], exc[np.maximum(e, 0), 1])
        eps = np.where(plain, np.sqrt(part[i, 2]) * np.sqrt(part[j, 2]), exc[np.maximum(e, 0), 2])
        s6 = sig ** 6
        g_eps = gp[:, 0] * 4.0 * s6 * s6 + gp[:, 1] * 4.0 * s6
        g_sig = gp[:, 0] * 48.0 * eps * sig ** 11 + gp[:, 1] * 24.0 * eps * sig ** 5
        g_qq = gp[:, 2] * self.k_coul
        gpart = np.zeros_like(part)
        gexc = np.zeros_like(exc)
        ip, jp = i[plain], j[plain]
        np.add.at(gpart[:, 0], ip, g_qq[plain] * part[jp, 0])
        np.add.at(gpart[:, 0], jp, g_qq[plain] * part[ip, 0])
        np.add.at(gpart[:, 1], ip, 0.5 * g_sig[plain])
        np.add.at(gpart[:, 1], jp, 0.5 * g_sig[plain])
        with np.errstate(divide="ignore", invalid="ignore"):
            # d sqrt(e_i e_j) / d e_i = sqrt(e_j / e_i) / 2; a vanishing epsilon has no finite derivative: report 0
            di = np.where(part[ip, 2] > 0, 0.5 * np.sqrt(part[jp, 2]) / np.sqrt(part[ip, 2]), 0.0)
            dj = np.where(part[jp, 2] > 0, 0.5 * np.sqrt(part[ip, 2]) / np.sqrt(part[jp, 2]), 0.0)
        np.add.at(gpart[:, 2], ip, g_eps[plain] * di)
        np.add.at(gpart[:, 2], jp, g_eps[plain] * dj)
        ee = e[~plain]
        np.add.at(gexc[:, 0], ee, g_qq[~plain])
        np.add.at(gexc[:, 1], ee, g_sig[~plain])
        np.add.at(gexc[:, 2], ee, g_eps[~plain])
        g[self.off_particle:self.off_exception] = gpart.ravel()
        g[self.off_exception:] = gexc.ravel()
        return g

    # ---- delta evaluation (pm_delta_eval) --------------------------------------------------------------
    def _delta_maps(self):
        if getattr(self, "_dmaps", None) is None:
            if getattr(self, "_pairs_cache", None) is None:
                self._pairs_cache = self.pairs()
            pm = self._pairs_cache
            lay = (ctypes.c_int * 5)()
            _lib.check(self._L.pm_derived_layout(self._h, lay), "pm_derived_layout")
            plain = [[] for _ in range(self.n_atoms)]
            of_exc = {}
            for q, (i, j, e) in enumerate(pm.tolist()):
                if e < 0:
                    plain[i].append(q)
                    plain[j].append(q)
                else:
                    of_exc[e] = q
            self._dmaps = dict(pairs=pm, plain=plain, of_exc=of_exc, doff=list(lay)[:4], pair_stride=int(lay[4]),
                               max_loc=int(self._L.pm_delta_max_local_atoms()),
                               n_terms=self.n_bonds + self.n_angles + self.n_torsions + self.n_pairs)
        return self._dmaps

    def delta_plan(self, base_slots, slots_rows, max_fraction=0.5):
        """
        Entry lists for ``pm_delta_eval``: which terms each row of ``slots_rows`` [P, n_slots] changes with respect to
        ``base_slots`` (a bonded slot changes its term, a particle slot every plain pair of the atom, an exception slot its
        pair).  Returns None when a row touches more than ``max_fraction`` of the terms or more atoms than the kernel's
        per-proposal limit -- the caller then evaluates the batch in full.
        """
        if self.rf:      # pm_delta_eval has no reaction-field variant: the caller evaluates in full
            return None
        m = self._delta_maps()
        k = self._keep
        rows = np.atleast_2d(np.asarray(slots_rows, dtype=np.float64))
        base = np.asarray(base_slots, dtype=np.float64)
        prop_ptr, atom_ptr, entries, atom_list = [0], [0], [], []
        max_loc = 1
        for row in rows:
            changed = np.nonzero(row != base)[0]
            terms = set()
            for idx in changed.tolist():
                if idx < self.off_angle:
                    terms.add((0, (idx - self.off_bond) // 2))
                elif idx < self.off_torsion:
                    terms.add((1, (idx - self.off_angle) // 2))
                elif idx < self.off_particle:
                    terms.add((2, (idx - self.off_torsion) // 2))
                elif idx < self.off_exception:
                    terms.update((3, q) for q in m["plain"][(idx - self.off_particle) // 3])
                else:
                    q = m["of_exc"].get((idx - self.off_exception) // 3)
                    if q is not None:
                        terms.add((3, q))
            if len(terms) > max_fraction * m["n_terms"]:
                return None
            terms = sorted(terms)
            atoms = set()
            for kind, t in terms:
                if kind == 0:
                    atoms.update(k["bond"][t].tolist())
                elif kind == 1:
                    atoms.update(k["angle"][t].tolist())
                elif kind == 2:
                    atoms.update(k["tors"][t].tolist())
                else:
                    atoms.update(m["pairs"][t, :2].tolist())
            atoms = sorted(atoms)
            if len(atoms) > m["max_loc"]:
                return None
            loc = {a: l for l, a in enumerate(atoms)}
            for kind, t in terms:
                if kind == 0:
                    at, pofs, per = k["bond"][t].tolist() + [0, 0], m["doff"][0] + 2 * t, 0
                elif kind == 1:
                    at, pofs, per = k["angle"][t].tolist() + [0], m["doff"][1] + 2 * t, 0
                elif kind == 2:
                    at, pofs, per = k["tors"][t].tolist(), m["doff"][2] + 4 * t, int(k["tper"][t])
                else:
                    at, pofs, per = m["pairs"][t, :2].tolist() + [0, 0], m["doff"][3] + m["pair_stride"] * t, 0
                n_at = (2, 3, 4, 2)[kind]
                lo = [loc[a] for a in at[:n_at]] + [0] * (4 - n_at)
                entries.append([kind, pofs, per, 0] + at + lo)
            atom_list += atoms
            prop_ptr.append(len(entries))
            atom_ptr.append(len(atom_list))
            max_loc = max(max_loc, len(atoms))
        return dict(prop_ptr=np.asarray(prop_ptr, dtype=np.int32),
                    entries=np.asarray(entries, dtype=np.int32).reshape(-1, 12) if entries else np.zeros((1, 12), dtype=np.int32),
                    atom_ptr=np.asarray(atom_ptr, dtype=np.int32),
                    atom_list=np.asarray(atom_list if atom_list else [0], dtype=np.int32), max_loc=int(max_loc),
                    n_entries=len(entries))

    def prepare(self, slots_dev, derived_dev=None):
        """slots [P, n_slots] float64, a CUDA tensor or a PINNED host tensor (read by the kernel over PCIe: the zero-copy input of
        the single-vector graph; then ``derived_dev`` must be given) -> derived [P, n_derived]."""
        torch = _torch()
        assert slots_dev.dtype == torch.float64 and slots_dev.is_contiguous()
        assert slots_dev.is_cuda or (slots_dev.is_pinned() and derived_dev is not None)
        assert slots_dev.dim() == 2 and slots_dev.shape[1] == self.n_slots
        n_vec = slots_dev.shape[0]
        if derived_dev is None:
            derived_dev = torch.empty((n_vec, self.n_derived), dtype=torch.float64, device=slots_dev.device)
        _lib.check(self._L.pm_prepare_params(self._h, slots_dev.data_ptr(), n_vec, derived_dev.data_ptr(),
                                             _stream_ptr(derived_dev.device)), "pm_prepare_params")
        return derived_dev


class DeviceEnsemble:
    """
    The conformations of one system, resident on one GPU in tile layout, plus their QM reference data.

    Parameters
    ----------
    topology : :obj:`DeviceTopology`
    coordinates : array [S,N,3] (nm), host numpy or CUDA tensor
    ref_forces : array [S,N,3] (kJ/mol/nm) or None
    ref_energies : array [S] (kJ/mol) or None
    """

    def __init__(self, topology, coordinates, ref_forces=None, ref_energies=None, pinned=False):
        torch = _torch()
        self.topology = topology
        self.device = torch.device("cuda", topology.device)
        self.h2d_bytes = 0
        x = self._to_device(coordinates, pinned)
        if x.dim() != 3 or x.shape[1] != topology.n_atoms or x.shape[2] != 3:
            raise ValueError("coordinates must have shape [S, %d, 3]" % topology.n_atoms)
        self.n_conf = int(x.shape[0])
        if self.n_conf < 1:
            raise ValueError("no conformations")
        self.coord_tiles = topology.pack(x, self.n_conf)
        del x
        self.fref_tiles = None
        self.eref = None
        if ref_forces is not None:
            f = self._to_device(ref_forces, pinned)
            if tuple(f.shape) != (self.n_conf, topology.n_atoms, 3):
                raise ValueError("ref_forces shape mismatch")
            self.fref_tiles = topology.pack(f, self.n_conf)
            del f
        if ref_energies is not None:
            self.eref = self._to_device(ref_energies, pinned).reshape(-1)
            if self.eref.shape[0] != self.n_conf:
                raise ValueError("ref_energies shape mismatch")
        self.weights = None            # CUDA [S] or None (= 1)
        self._scratch = None
        self._buffers = {}
        self.delta_base = None

    def _to_device(self, a, pinned=False):
        torch = _torch()
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=torch.float64).contiguous()
            if not a.is_cuda:
                self.h2d_bytes += t.numel() * 8
            return t
        a = np.ascontiguousarray(a, dtype=np.float64)
        t = torch.from_numpy(a)
        if pinned:
            t = t.pin_memory()
        self.h2d_bytes += t.numel() * 8
        return t.to(self.device, non_blocking=pinned)

    def set_weights(self, weights):
        self.weights = None if weights is None else self._to_device(weights).reshape(-1)

    def _buf(self, name, shape):
        torch = _torch()
        b = self._buffers.get(name)
        if b is None or tuple(b.shape) != tuple(shape):
            b = torch.empty(shape, dtype=torch.float64, device=self.device)
            self._buffers[name] = b
        return b

    # ------------------------------------------------------------------------------------------
    def evaluate(self, derived, with_forces=True, want_fres=True, want_fmm=False, nb_cache=None):
        """
        Launch the E(+F) kernel for ``derived`` [P, n_derived] (CUDA).  Returns (emm [P,S], fres [P,S] or None,
        fmm_tiles [P, tiles*tile_doubles] or None); outputs are views of reusable device buffers.
        ``nb_cache`` = (E_nb [S], F_nb tiles) from :meth:`build_nonbonded_cache`: bonded terms are evaluated, nonbonded
        contributions come from the cache.
        """
        topo = self.topology
        n_vec = int(derived.shape[0])
        emm = self._buf("emm", (n_vec, self.n_conf))
        fres = None
        fmm = None
        if with_forces and want_fres:
            if self.fref_tiles is None:
                raise B200Error("force residual requested but the ensemble holds no reference forces")
            fres = self._buf("fres", (n_vec, self.n_conf))
        if with_forces and want_fmm:
            fmm = self._buf("fmm", (n_vec, topo.n_tiles(self.n_conf) * topo.tile_doubles))
        mode, nb_e, nb_f = 0, None, None
        if nb_cache is not None:
            mode, nb_e, nb_f = 2, nb_cache[0].data_ptr(), nb_cache[1].data_ptr()
        _lib.check(topo._L.pm_eval_ex(topo._h, derived.data_ptr(), n_vec, self.coord_tiles.data_ptr(),
                                      self.fref_tiles.data_ptr() if (fres is not None) else None, self.n_conf,
                                      int(bool(with_forces)), emm.data_ptr(),
                                      fres.data_ptr() if fres is not None else None,
                                      fmm.data_ptr() if fmm is not None else None, mode, nb_e, nb_f,
                                      _stream_ptr(self.device)), "pm_eval_ex")
        return emm, fres, fmm

    def build_nonbonded_cache(self, derived_row):
        """(E_nb [S], F_nb tiles) of ONE parameter vector ``derived_row`` [1, n_derived]: the nonbonded terms only."""
        torch = _torch()
        topo = self.topology
        e_nb = torch.empty((1, self.n_conf), dtype=torch.float64, device=self.device)
        f_nb = torch.empty(topo.n_tiles(self.n_conf) * topo.tile_doubles, dtype=torch.float64, device=self.device)
        _lib.check(topo._L.pm_eval_ex(topo._h, derived_row.data_ptr(), 1, self.coord_tiles.data_ptr(), None, self.n_conf, 1,
                                      e_nb.data_ptr(), None, f_nb.data_ptr(), 1, None, None, _stream_ptr(self.device)),
                   "pm_eval_ex")
        return e_nb.reshape(-1), f_nb

    def reduce(self, emm, fres, d_shift=0.0, weights="own"):
        """Partial sums [P, 8] (CUDA) of this ensemble; see ``pm_reduce_objective`` in include/paramol_b200.h."""
        torch = _torch()
        topo = self.topology
        n_vec = int(emm.shape[0])
        need = int(topo._L.pm_reduce_scratch_doubles(n_vec))
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.float64, device=self.device)
        partials = self._buf("partials", (n_vec, NPARTIAL))
        w = self.weights if isinstance(weights, str) else weights
        _lib.check(topo._L.pm_reduce_objective(emm.data_ptr(), fres.data_ptr() if fres is not None else None,
                                               self.eref.data_ptr() if self.eref is not None else None,
                                               w.data_ptr() if w is not None else None, float(d_shift), n_vec, self.n_conf,
                                               partials.data_ptr(), self._scratch.data_ptr(), _stream_ptr(self.device)),
                   "pm_reduce_objective")
        return partials

    def evaluate_reduce(self, derived, with_forces=True, d_shift=0.0, weights="own", comm=None, want_emm=False, want_fres=False,
                        partials_out=None):
        """
        ``pm_eval_reduce``: evaluation, residual reduction and -- with ``comm`` (:obj:`paramol_b200.Utils.dist.DeviceComm`) --
        the sum over ranks as one pass; no [P, S] arrays are written unless asked for.  Returns (partials [P, 8] CUDA: global
        sums when ``comm`` is given, emm or None, fres or None).  The force residual is included when ``with_forces`` and the
        ensemble holds reference forces.
        """
        torch = _torch()
        topo = self.topology
        n_vec = int(derived.shape[0])
        need = int(topo._L.pm_eval_reduce_scratch_doubles(topo._h, n_vec, self.n_conf))
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.float64, device=self.device)
        # partials_out: a pinned host tensor [P, 8] the row-sum kernel writes directly (zero-copy output of the single-vector graph)
        partials = self._buf("partials", (n_vec, NPARTIAL)) if partials_out is None else partials_out
        assert tuple(partials.shape) == (n_vec, NPARTIAL) and partials.dtype == torch.float64 and (partials.is_cuda or partials.is_pinned())
        w = self.weights if isinstance(weights, str) else weights
        use_f = bool(with_forces) and self.fref_tiles is not None
        emm = self._buf("emm", (n_vec, self.n_conf)) if want_emm else None
        fres = self._buf("fres", (n_vec, self.n_conf)) if (want_fres and use_f) else None
        _lib.check(topo._L.pm_eval_reduce(topo._h, derived.data_ptr(), n_vec, self.coord_tiles.data_ptr(),
                                          self.fref_tiles.data_ptr() if use_f else None, self.n_conf, int(bool(with_forces)),
                                          self.eref.data_ptr() if self.eref is not None else None,
                                          w.data_ptr() if w is not None else None, float(d_shift), partials.data_ptr(),
                                          self._scratch.data_ptr(), comm.handle if comm is not None else None,
                                          emm.data_ptr() if emm is not None else None,
                                          fres.data_ptr() if fres is not None else None, _stream_ptr(self.device)),
                   "pm_eval_reduce")
        return partials, emm, fres

    def energy_batch_supported(self):
        """True when ``pm_energy_batch_reduce`` serves this topology (NoCutoff, coordinate tile fits the shared memory)."""
        return bool(self.topology._L.pm_energy_batch_supported(self.topology._h))

    def energy_batch_reduce(self, derived, d_shift=0.0, weights="own", comm=None, want_emm=False):
        """
        ``pm_energy_batch_reduce``: energies of a BATCH of parameter vectors with the geometry evaluated once per conformation
        (E = C G on the FP64 tensor-core path) and the energy sums of ``pm_reduce_objective`` ([4] = 0); for energy-only
        objectives, WHAM passes and finite-difference gradients of the energy term.  Returns (partials [P, 8] CUDA -- global sums
        when ``comm`` is given --, emm [P, S] or None).
        """
        torch = _torch()
        topo = self.topology
        n_vec = int(derived.shape[0])
        need = int(topo._L.pm_energy_batch_scratch_doubles(topo._h, n_vec, self.n_conf))
        if getattr(self, "_eg_scratch", None) is None or self._eg_scratch.numel() < need:
            self._eg_scratch = torch.empty(need, dtype=torch.float64, device=self.device)
        partials = self._buf("partials", (n_vec, NPARTIAL))
        w = self.weights if isinstance(weights, str) else weights
        emm = self._buf("emm", (n_vec, self.n_conf)) if want_emm else None
        _lib.check(topo._L.pm_energy_batch_reduce(topo._h, derived.data_ptr(), n_vec, self.coord_tiles.data_ptr(), self.n_conf,
                                                  self.eref.data_ptr() if self.eref is not None else None,
                                                  w.data_ptr() if w is not None else None, float(d_shift), partials.data_ptr(),
                                                  self._eg_scratch.data_ptr(), comm.handle if comm is not None else None,
                                                  emm.data_ptr() if emm is not None else None, _stream_ptr(self.device)),
                   "pm_energy_batch_reduce")
        return partials, emm

    # ---- delta evaluation ---------------------------------------------------------------------------------
    def set_delta_base(self, slots_row, derived_row, emm_row, fres_row, fmm_tiles_row):
        """Keep copies of the per-conformation results of ONE fully evaluated vector as the base of later delta evaluations."""
        self.delta_base = dict(slots=np.array(slots_row, dtype=np.float64, copy=True), derived=derived_row.reshape(-1).clone(),
                               emm=emm_row.reshape(-1).clone(),
                               fres=None if fres_row is None else fres_row.reshape(-1).clone(),
                               fmm=None if fmm_tiles_row is None else fmm_tiles_row.reshape(-1).clone(), n_commits=0)
        return self.delta_base

    def _delta_call(self, plan, derived, emm, fres, commit):
        torch = _torch()
        topo, b = self.topology, self.delta_base
        dev = [torch.from_numpy(plan[key]).to(self.device) for key in ("prop_ptr", "entries", "atom_ptr", "atom_list")]
        with_f = b["fmm"] is not None
        _lib.check(topo._L.pm_delta_eval(topo._h, b["derived"].data_ptr(), derived.data_ptr(), int(derived.shape[0]),
                                         dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(),
                                         plan["max_loc"], self.coord_tiles.data_ptr(),
                                         self.fref_tiles.data_ptr() if with_f else None,
                                         b["fmm"].data_ptr() if with_f else None, b["emm"].data_ptr(),
                                         b["fres"].data_ptr() if with_f else None, self.n_conf,
                                         emm.data_ptr() if emm is not None else None,
                                         fres.data_ptr() if fres is not None else None, int(commit),
                                         _stream_ptr(self.device)), "pm_delta_eval")

    def delta_evaluate(self, plan, derived):
        """(emm [P,S], fres [P,S] or None) of the rows described by ``plan`` (:meth:`DeviceTopology.delta_plan`) relative to
        the stored base; same meaning as the outputs of :meth:`evaluate`."""
        n_vec = int(derived.shape[0])
        emm = self._buf("emm", (n_vec, self.n_conf))
        fres = self._buf("fres", (n_vec, self.n_conf)) if self.delta_base["fmm"] is not None else None
        self._delta_call(plan, derived, emm, fres, 0)
        return emm, fres

    def delta_commit(self, plan, slots_row, derived_row):
        """Fold ONE proposal (an accepted move) into the base arrays."""
        self._delta_call(plan, derived_row, None, None, 1)
        b = self.delta_base
        b["slots"] = np.array(slots_row, dtype=np.float64, copy=True)
        b["derived"] = derived_row.reshape(-1).clone()
        b["n_commits"] += 1

    def term_gradient(self, derived_row, coef_e=None, coef_f=None, fmm_tiles=None):
        """
        ``pm_grad``: sum_s [coef_e[s] dE_s/dp + coef_f[s] d/dp sum_a (F_sa - Fref_sa)^T M_a (F_sa - Fref_sa)] for ONE parameter
        vector ``derived_row`` [1, n_derived]; ``fmm_tiles`` = the MM force tiles of that vector (``evaluate(want_fmm=True)``).
        Returns a CUDA tensor [pm_grad_out_doubles] (see include/paramol_b200.h for its layout).
        """
        torch = _torch()
        topo = self.topology
        L = topo._L
        n_out = int(L.pm_grad_out_doubles(topo._h))
        need = int(L.pm_grad_scratch_doubles(topo._h))
        if need < 0:
            raise B200Error("pm_grad_scratch_doubles: " + L.pm_last_error().decode())
        scratch = self._buffers.get("grad_scratch")
        if scratch is None or scratch.numel() < max(need, 1):
            scratch = torch.empty(max(need, 1), dtype=torch.float64, device=self.device)
            self._buffers["grad_scratch"] = scratch
        gout = self._buf("gout", (max(n_out, 1),))
        gout.zero_()
        if coef_f is not None and (fmm_tiles is None or self.fref_tiles is None):
            raise B200Error("force coefficients need MM force tiles and reference forces")
        _lib.check(L.pm_grad(topo._h, derived_row.data_ptr(), self.coord_tiles.data_ptr(),
                             fmm_tiles.data_ptr() if coef_f is not None else None,
                             self.fref_tiles.data_ptr() if coef_f is not None else None, self.n_conf,
                             coef_e.data_ptr() if coef_e is not None else None,
                             coef_f.data_ptr() if coef_f is not None else None, gout.data_ptr(), scratch.data_ptr(),
                             _stream_ptr(self.device)), "pm_grad")
        return gout[:n_out]

    def forces_from_tiles(self, fmm_tiles):
        """[P, tiles] -> torch [P, S, N, 3]."""
        torch = _torch()
        return torch.stack([self.topology.unpack(fmm_tiles[p], self.n_conf) for p in range(fmm_tiles.shape[0])])
