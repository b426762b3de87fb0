#!/usr/bin/env python
"""SASS opcode histogram per kernel of libparamol_b200.so (and of the generated image of a molecule, if given):
   python scripts/sass_histogram.py [extra.cubin ...] > profiles/r2_sass_histogram.txt"""
import collections, os, re, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
files = [os.path.join(REPO, "paramol_b200", "libparamol_b200.so")] + sys.argv[1:]
KEY = ("DMMA", "DFMA", "DMUL", "DADD", "MUFU", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "ATOMS", "LDC", "LDCU", "UTCMMA", "LDTM")
for f in files:
    out = subprocess.run(["cuobjdump", "-sass", f], stdout=subprocess.PIPE, universal_newlines=True).stdout
    name, hist = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            hist[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            hist[name][m.group(1).split(".")[0]] += 1
    print("== %s" % os.path.relpath(f, REPO))
    for k in sorted(hist, key=lambda n: -sum(hist[n].values())):
        h = hist[k]
        tot = sum(h.values())
        if tot < 200:
            continue
        short = subprocess.run(["c++filt", k], stdout=subprocess.PIPE, universal_newlines=True).stdout.strip().split("(")[0][:90]
        print("%-92s %6d instr | %s" % (short, tot, "  ".join("%s %d" % (op, h[op]) for op in KEY if h.get(op))))
