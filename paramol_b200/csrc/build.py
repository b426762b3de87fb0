# -*- coding: utf-8 -*-
"""
Builds ``paramol_b200/libparamol_b200.so`` (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

Usage:  python -m paramol_b200.csrc.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libparamol_b200.so")
SOURCES = ["pm_kernels.cu", "pm_fit_kernels.cu", "pm_grad_kernels.cu", "pm_delta_kernels.cu", "pm_spec.cu", "pm_comm.cu", "pm_lls_kernels.cu", "pm_egemm_kernels.cu"]
HEADERS = ["pm_device.cuh", "pm_program.hpp", "pm_internal.hpp", os.path.join("..", "..", "include", "paramol_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 path cannot be built (there is no CPU fallback)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a (-lineinfo so ncu's source page maps to our code)."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB] + [os.path.join(HERE, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
