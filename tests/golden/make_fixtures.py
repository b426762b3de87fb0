#!/usr/bin/env python
"""
Generate ``tests/golden/`` from the reference checkout (run in the BUILD container only;
``/root/reference`` does not exist on the GPU box, so nothing at test/bench time reads it).

What it produces
----------------
1. Input data files copied verbatim (they are data, not code): AMBER topologies/coordinates, the
   10-structure NetCDF file and the serialized ``aniline.xml`` system used by the reference's own tests,
   plus the largest in-tree molecules used for parity sweeps.
2. ``reference_goldens.json``: every numeric golden value asserted by the reference's pytest suite for the
   objective-evaluation path, extracted *from the test sources with ``ast``* (no re-typing), keyed by
   ``<test file>::<variable>``.
3. ``aniline_esp.npz``: the Gaussian ESP data of ``Tests/Tasks/aniline_opt.log`` parsed with this repo's
   ``GaussianESP`` reader (the 3.5 MB log itself is not committed), and ``aniline_opt_excerpt.log``: a
   short excerpt of the log for reader tests.

Usage:  python tests/golden/make_fixtures.py [/root/reference]
"""
import ast
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

COPY = [
    ("ParaMol/Tests/aniline.prmtop", "aniline.prmtop"),
    ("ParaMol/Tests/aniline.inpcrd", "aniline.inpcrd"),
    ("ParaMol/Tests/aniline.xml", "aniline.xml"),
    ("ParaMol/Tests/aniline_10_struct.nc", "aniline_10_struct.nc"),
    ("Examples/Example_5/norfloxacin.prmtop", "norfloxacin.prmtop"),
    ("Examples/Example_5/norfloxacin.inpcrd", "norfloxacin.inpcrd"),
    ("Examples/Example_4/aspirin.prmtop", "aspirin.prmtop"),
    ("Examples/Example_4/aspirin.inpcrd", "aspirin.inpcrd"),
    ("Examples/Example_3/caffeine.prmtop", "caffeine.prmtop"),
    ("Examples/Example_3/caffeine.inpcrd", "caffeine.inpcrd"),
    ("Examples/Example_7/ethane.prmtop", "ethane.prmtop"),
    ("Examples/Example_7/ethane.inpcrd", "ethane.inpcrd"),
]

TEST_FILES = [
    "ParaMol/Tests/System/test_system.py",
    "ParaMol/Tests/Objective_function/test_objective_function.py",
    "ParaMol/Tests/Parameter_space/test_parameter_space.py",
    "ParaMol/Tests/Force_field/test_force_field.py",
    "ParaMol/Tests/MM_engines/test_openmm_wrapper.py",
    "ParaMol/Tests/Tasks/test_resp.py",
]


def _literal(node):
    """Evaluate list/number/dict literals, unwrapping np.asarray(...)/np.array(...)."""
    if isinstance(node, ast.Call) and getattr(node.func, "attr", "") in ("asarray", "array") and node.args:
        return _literal(node.args[0])
    return ast.literal_eval(node)


def extract_goldens(ref_root):
    out = {}
    for rel in TEST_FILES:
        path = os.path.join(ref_root, rel)
        tree = ast.parse(open(path).read())
        short = os.path.basename(rel)
        for func in ast.walk(tree):
            if not isinstance(func, ast.FunctionDef):
                continue
            counter = {}
            for node in ast.walk(func):
                # <name> = literal
                if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
                    name = node.targets[0].id
                    try:
                        val = _literal(node.value)
                    except Exception:
                        continue
                    if isinstance(val, (int, float)) and not name.endswith("to_compare"):
                        continue
                    if isinstance(val, str) or val is None:
                        continue
                    if isinstance(val, dict) and not all(isinstance(v, (int, float)) for v in val.values()):
                        continue
                    k = counter.get(name, 0)
                    counter[name] = k + 1
                    key = "{}::{}::{}".format(short, func.name, name) + ("" if k == 0 else "#%d" % k)
                    out[key] = val
                # assert abs(f_val - <const>) < tol
                if isinstance(node, ast.Assert) and isinstance(node.test, ast.Compare):
                    left = node.test.left
                    if isinstance(left, ast.Call) and getattr(left.func, "id", "") == "abs" and left.args:
                        arg = left.args[0]
                        if isinstance(arg, ast.BinOp) and isinstance(arg.op, ast.Sub) and isinstance(arg.right, ast.Constant) \
                                and isinstance(arg.left, ast.Name):
                            name = arg.left.id
                            k = counter.get(name, 0)
                            counter[name] = k + 1
                            key = "{}::{}::{}".format(short, func.name, name) + ("" if k == 0 else "#%d" % k)
                            out[key] = arg.right.value
    return out


def main():
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    for src, dst in COPY:
        shutil.copyfile(os.path.join(ref_root, src), os.path.join(HERE, dst))
        os.chmod(os.path.join(HERE, dst), 0o644)

    goldens = extract_goldens(ref_root)
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as f:
        json.dump(goldens, f, indent=1, sort_keys=True)
    print("goldens:", len(goldens))
    for k in sorted(goldens):
        v = goldens[k]
        print("  ", k, np.shape(v) if not isinstance(v, dict) else "dict[%d]" % len(v))

    from paramol_b200.Utils.gaussian_esp import GaussianESP
    log = os.path.join(ref_root, "ParaMol/Tests/Tasks/aniline_opt.log")
    conf, grid, esp, energy = GaussianESP().read_log_files([log])
    np.savez_compressed(os.path.join(HERE, "aniline_esp.npz"), conformation=conf[0], grid=grid[0], esp=esp[0],
                        energy=np.asarray(energy[0]))
    # excerpt: keep every relevant line kind but only the first 200 fit centres / fit values
    keep = []
    n_fc = n_fit = 0
    for line in open(log):
        if "Fit Center" in line:
            n_fc += 1
            if n_fc <= 200:
                keep.append(line)
        elif "Fit    " in line:
            n_fit += 1
            if n_fit <= 200:
                keep.append(line)
        elif "Atomic Center" in line or "SCF Done:" in line:
            keep.append(line)
    with open(os.path.join(HERE, "aniline_opt_excerpt.log"), "w") as f:
        f.writelines(keep)
    print("esp grid:", grid[0].shape, "excerpt lines:", len(keep))


if __name__ == "__main__":
    main()
