#!/usr/bin/env python
"""Executed-instruction mix by SASS opcode from an .ncu-rep: `python scripts/ncu_opmix.py rep.ncu-rep`."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, universal_newlines=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
ismp = hdr.index("# Samples")
mix = collections.Counter(); smp = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) != len(hdr): continue
    src = r[isrc].strip()
    toks = src.split()
    if not toks: continue
    op = toks[0]
    if op.startswith("@"): op = toks[1]
    op = op.split(".")[0]
    try: n = int(r[iex]); s = int(r[ismp])
    except ValueError: continue
    mix[op] += n; smp[op] += s; tot += n
print("total executed warp instructions", tot)
fp64 = sum(v for k, v in mix.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("FP64-pipe %.1f%%" % (100.0 * fp64 / tot))
for op, n in mix.most_common(28):
    print("%-10s %12d %5.1f%%  samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * smp[op] / max(1, sum(smp.values()))))
