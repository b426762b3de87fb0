"""
CPU-only checks of the C-ABI library: it loads, exports every symbol include/paramol_b200.h declares, fails loudly
without a GPU (no CPU fallback), and the host-side topology-program builder (paramol_b200/csrc/pm_program.hpp) keeps
its invariants for every molecule and every lanes-per-conformation setting: each bond / angle / torsion / interacting
pair executed exactly once, records that share a step atom-disjoint, no two lanes updating the same j in one step.
"""
import ctypes
import os
import re

import numpy as np
import pytest

from paramol_b200 import _lib
from paramol_b200._lib import PmTopology, c_int_p
from paramol_b200.Utils.amber import system_from_prmtop
from paramol_b200.Utils.synthetic import make_ligand

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "paramol_b200.h")).read()
    declared = set(re.findall(r"\b(pm_[a-z0-9_]+)\s*\(", header))
    declared -= {"pm_handle_s"}
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "symbol %s declared in the header is not exported" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.pm_version() >= 200


def _topology(s):
    keep = dict(bond=np.ascontiguousarray(s.bond_idx, dtype=np.int32), angle=np.ascontiguousarray(s.angle_idx, dtype=np.int32),
                tors=np.ascontiguousarray(s.torsion_idx, dtype=np.int32), tper=np.ascontiguousarray(s.torsion_per, dtype=np.int32),
                exc=np.ascontiguousarray(s.exception_idx, dtype=np.int32))
    par = np.asarray(s.exception_par).reshape(-1, 3)
    keep["act"] = np.ascontiguousarray(((par[:, 0] != 0) | (par[:, 2] != 0)).astype(np.int32))
    p = lambda a: a.ctypes.data_as(c_int_p)
    topo = PmTopology(s.n_atoms, len(keep["bond"]), p(keep["bond"]), len(keep["angle"]), p(keep["angle"]), len(keep["tors"]),
                      p(keep["tors"]), p(keep["tper"]), len(keep["exc"]), p(keep["exc"]), p(keep["act"]), 0, 1.2, 78.3, 138.935456)
    return topo, keep


def _molecules(golden_dir):
    for name in ("aniline", "caffeine", "aspirin", "norfloxacin", "ethane"):
        yield name, system_from_prmtop(os.path.join(golden_dir, name + ".prmtop"))
    for n in (7, 33, 50, 60):
        yield "lig%d" % n, make_ligand(n)[0]


def test_topology_program_invariants(golden_dir):
    L = _lib.lib()
    for name, s in _molecules(golden_dir):
        for cpw in (32, 16, 8, 4, -8):   # -8: teams of 8 warps on tiles of 32 conformations
            topo, keep = _topology(s)
            stats = np.zeros(10, dtype=np.int32)
            rc = L.pm_program_check(ctypes.byref(topo), cpw, stats.ctypes.data_as(c_int_p))
            assert rc == 0, (name, cpw, L.pm_last_error())
            assert (stats[0], stats[1]) == ((32, 8) if cpw < 0 else (cpw, 32 // cpw))
            # padding of the pair schedule stays moderate
            if stats[7] > 0:
                assert stats[8] <= 2 * stats[7] + 8 * stats[1], (name, cpw, stats)


def test_program_with_extra_torsions_and_repeated_terms(golden_dir):
    from paramol_b200.MM_engines.b200_engine import B200Engine
    e = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(golden_dir, "aniline.prmtop"),
                   crd_file=os.path.join(golden_dir, "aniline.inpcrd"))
    e.add_torsion_terms()  # 4 periodicities per quadruple: multi-term cells
    s = e.system
    # a duplicated bond and a duplicated angle must still be executed once each
    s.bond_idx = np.concatenate([s.bond_idx, s.bond_idx[:1]])
    s.bond_par = np.concatenate([s.bond_par, s.bond_par[:1]])
    s.angle_idx = np.concatenate([s.angle_idx, s.angle_idx[:1]])
    s.angle_par = np.concatenate([s.angle_par, s.angle_par[:1]])
    L = _lib.lib()
    for cpw in (32, 16, 8, 4):
        topo, keep = _topology(s)
        assert L.pm_program_check(ctypes.byref(topo), cpw, None) == 0, L.pm_last_error()


def test_degenerate_terms_are_rejected(golden_dir):
    s = system_from_prmtop(os.path.join(golden_dir, "ethane.prmtop")).copy()
    s.angle_idx[0] = (1, 1, 2)
    topo, keep = _topology(s)
    L = _lib.lib()
    assert L.pm_program_check(ctypes.byref(topo), 32, None) != 0
    assert b"degenerate" in L.pm_last_error()


def test_compute_calls_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from paramol_b200.MM_engines.device import DeviceTopology
    s = make_ligand(7)[0]
    with pytest.raises(_lib.B200Error):
        DeviceTopology(s)
    L = _lib.lib()
    topo, keep = _topology(s)
    h = ctypes.c_void_p()
    assert L.pm_create(ctypes.byref(topo), 0, 0, ctypes.byref(h)) != 0
    assert b"no CUDA device" in L.pm_last_error()


def test_specialized_kernel_source_builds_without_gpu(tmp_path, monkeypatch):
    """Host-only half of the topology-specialised kernel: pm_topology_pairs (no CUDA call), the generator and the nvcc cross
    compile.  The image must hold both kernels and the constant table, and ptxas must not spill beyond the accepted budget."""
    from paramol_b200.MM_engines.device import prebuild_specialized, _pm_topology
    from paramol_b200.csrc import specialize
    from paramol_b200.Utils.synthetic import make_ligand
    import ctypes
    monkeypatch.setenv("PARAMOL_B200_SPEC_CACHE", str(tmp_path))
    mm, _ = make_ligand(9)
    L = _lib.lib()
    topo, keep = _pm_topology(mm)
    n = ctypes.c_int()
    assert L.pm_topology_pairs(ctypes.byref(topo), 32, ctypes.byref(n), None) == 0
    exc = {tuple(sorted(map(int, e))) for e, a in zip(keep["exc"], keep["act"]) if not a}
    assert n.value == 9 * 8 // 2 - len(exc)             # every pair except the pure exclusions
    size = prebuild_specialized(mm)
    assert size > 0
    cubins = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(cubins) == 1
    blob = open(os.path.join(tmp_path, cubins[0]), "rb").read()
    assert b"pm_ef_spec_kernel" in blob and b"pm_e_spec_kernel" in blob and b"spec_tab" in blob
    report = open(os.path.join(tmp_path, cubins[0] + ".txt")).read()
    assert 0 <= specialize._spills(report, "pm_ef_spec_kernel") <= specialize.F_REG_MAX_SPILL
    assert specialize._spills(report, "pm_e_spec_kernel") == 0
    # molecules whose tiles leave too few warps are left to the generic kernel
    big, _ = make_ligand(30)
    assert prebuild_specialized(big) == 0
