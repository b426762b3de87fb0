# -*- coding: utf-8 -*-
"""
ctypes binding of ``libparamol_b200.so`` (declared in ``include/paramol_b200.h``).

There is deliberately no fallback: if the library cannot be loaded, or a compute entry point reports an
error (for instance "no CUDA device"), a :obj:`B200Error` is raised.  The CPU oracle under ``oracle/`` is
test infrastructure and is never imported from here.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PARAMOL_B200_LIB") or os.path.join(_HERE, "libparamol_b200.so")  # env: A/B builds only
_LIB = None

c_int_p = ctypes.POINTER(ctypes.c_int)
c_double_p = ctypes.POINTER(ctypes.c_double)


class B200Error(RuntimeError):
    pass


class PmTopology(ctypes.Structure):
    _fields_ = [("n_atoms", ctypes.c_int),
                ("n_bonds", ctypes.c_int), ("bond_idx", c_int_p),
                ("n_angles", ctypes.c_int), ("angle_idx", c_int_p),
                ("n_torsions", ctypes.c_int), ("torsion_idx", c_int_p), ("torsion_per", c_int_p),
                ("n_exceptions", ctypes.c_int), ("exception_idx", c_int_p), ("exception_active", c_int_p),
                ("nonbonded_method", ctypes.c_int),
                ("cutoff", ctypes.c_double),
                ("rf_dielectric", ctypes.c_double),
                ("one_4pi_eps0", ctypes.c_double)]


# name -> (restype, argtypes); the single source of truth for tests/test_abi.py as well
SIGNATURES = {
    "pm_last_error": (ctypes.c_char_p, []),
    "pm_version": (ctypes.c_int, []),
    "pm_device_count": (ctypes.c_int, []),
    "pm_create": (ctypes.c_int, [ctypes.POINTER(PmTopology), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "pm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pm_n_slots": (ctypes.c_int, [ctypes.c_void_p]),
    "pm_n_derived": (ctypes.c_int, [ctypes.c_void_p]),
    "pm_n_pairs": (ctypes.c_int, [ctypes.c_void_p]),
    "pm_conf_per_tile": (ctypes.c_int, [ctypes.c_void_p]),
    "pm_n_tiles": (ctypes.c_long, [ctypes.c_void_p, ctypes.c_long]),
    "pm_program_info": (ctypes.c_int, [ctypes.c_void_p, c_int_p]),
    "pm_program_check": (ctypes.c_int, [ctypes.POINTER(PmTopology), ctypes.c_int, c_int_p]),
    "pm_tile_doubles": (ctypes.c_long, [ctypes.c_void_p]),
    "pm_get_pairs": (ctypes.c_int, [ctypes.c_void_p, c_int_p]),
    "pm_launch_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_int_p, c_int_p, c_int_p, c_int_p]),
    "pm_set_force_metric": (ctypes.c_int, [ctypes.c_void_p, c_double_p]),
    "pm_pack_tiles": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p]),
    "pm_unpack_tiles": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p]),
    "pm_prepare_params": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pm_eval": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_long, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_void_p]),
    "pm_eval_ex": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_long, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pm_reduce_scratch_doubles": (ctypes.c_long, [ctypes.c_int]),
    "pm_reduce_objective": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_double, ctypes.c_int, ctypes.c_long, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p]),
    "pm_eval_reduce_scratch_doubles": (ctypes.c_long, [ctypes.c_void_p, ctypes.c_int, ctypes.c_long]),
    "pm_eval_reduce": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long,
                                      ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pm_energy_batch_supported": (ctypes.c_int, [ctypes.c_void_p]),
    "pm_energy_batch_scratch_doubles": (ctypes.c_long, [ctypes.c_void_p, ctypes.c_int, ctypes.c_long]),
    "pm_energy_batch_reduce": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p]),
    "pm_comm_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.POINTER(ctypes.c_void_p)]),
    "pm_comm_connect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "pm_comm_allreduce": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "pm_comm_timeouts": (ctypes.c_long, [ctypes.c_void_p]),
    "pm_comm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pm_lls_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "pm_lls_n_cols": (ctypes.c_int, [ctypes.c_void_p]),
    "pm_lls_set_columns": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pm_lls_scratch_doubles": (ctypes.c_long, [ctypes.c_void_p, ctypes.c_long]),
    "pm_lls_colstats": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pm_lls_residual": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "pm_lls_normal": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p]),
    "pm_lls_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pm_resp_scratch_doubles": (ctypes.c_long, [ctypes.c_int, ctypes.c_long, ctypes.c_int]),
    "pm_resp_assemble": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_long, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pm_esp_potential": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pm_lls_internal": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pm_lls_columns": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pm_grad_out_doubles": (ctypes.c_long, [ctypes.c_void_p]),
    "pm_grad_scratch_doubles": (ctypes.c_long, [ctypes.c_void_p]),
    "pm_grad": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long,
                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pm_delta_max_local_atoms": (ctypes.c_int, []),
    "pm_derived_layout": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]),
    "pm_delta_eval": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                 ctypes.c_void_p]),
    "pm_set_specialized": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int]),
    "pm_has_specialized": (ctypes.c_int, [ctypes.c_void_p]),
    "pm_topology_pairs": (ctypes.c_int, [ctypes.POINTER(PmTopology), ctypes.c_int, c_int_p, c_int_p]),
    "pm_graph_launch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]),
    "pm_graph_wait": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]),
    "pm_measure_fma_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p]),
}


def lib():
    """Load (once) and return the C-ABI library; raises B200Error when it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error("%s is missing: build it with `python -m paramol_b200.csrc.build` "
                            "(the B200 path has no CPU fallback)" % LIB_PATH)
        try:
            L = ctypes.CDLL(LIB_PATH)
        except OSError as exc:
            raise B200Error("cannot load %s: %s" % (LIB_PATH, exc))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc, what=""):
    if rc != 0:
        msg = lib().pm_last_error()
        raise B200Error("%s failed: %s" % (what or "libparamol_b200 call", msg.decode() if msg else "unknown error"))
