#!/usr/bin/env python
"""Where the generated E+F kernel beats the interpreter as molecules grow: `python scripts/spec_size_sweep.py build|run`.
`build` (no GPU needed) compiles every variant into the in-tree cache; `run` times them on the GPU (2^18 conformations)."""
import os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOLS = ["lig20", "lig22", "caffeine", "lig26", "lig28", "lig32"]
VARIANTS = {
    "generic": {"PARAMOL_B200_SPECIALIZE": "0"},
    "freg_f1": {"PM_SPEC_FREG": "1", "PM_SPEC_FREG_MAX": "64", "PM_SPEC_FENCE": "1"},
    "freg_f2": {"PM_SPEC_FREG": "1", "PM_SPEC_FREG_MAX": "64", "PM_SPEC_FENCE": "2"},
    "tile": {"PM_SPEC_FREG": "0", "PM_SPEC_MIN_WARPS": "4"},
}
mode = sys.argv[1]
for mol in MOLS:
    for name, env in VARIANTS.items():
        e = dict(os.environ); e.update(env)
        if mode == "build":
            if name == "generic":
                continue
            code = ("import sys; sys.path.insert(0, %r)\n"
                    "from paramol_b200.MM_engines.device import prebuild_specialized\n"
                    "from paramol_b200.Utils.synthetic import make_ligand\n"
                    "from paramol_b200.Utils.amber import system_from_prmtop\n"
                    "m = %r\n"
                    "s = make_ligand(int(m[3:]))[0] if m.startswith('lig') else system_from_prmtop(%r + '/tests/golden/' + m + '.prmtop')\n"
                    "print(m, %r, prebuild_specialized(s))\n") % (REPO, mol, REPO, name)
            subprocess.run([sys.executable, "-c", code], env=e)
        else:
            r = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "spec_sweep.py"), mol, str(1 << 18)], env=e,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
            lines = [l for l in r.stdout.splitlines() if "specialized=" in l]
            print("%-9s %-8s %s" % (mol, name, lines[0] if lines else r.stdout[-300:]), flush=True)
