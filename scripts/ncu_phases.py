#!/usr/bin/env python
"""Executed instructions and stall samples of pm_ef_kernel per PHASE (prologue / stars / axes / pair blocks / pair segments /
epilogue) from an .ncu-rep: the SASS listing is cut at the first instruction attributed to each phase's marker comment in
pm_kernels.cu (inlined helpers stay with the phase that calls them).  `python scripts/ncu_phases.py rep.ncu-rep`"""
import csv, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, universal_newlines=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
sass = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        sass.append(dict(zip(hdr, r)))
if not sass:
    print(rows[:5]); sys.exit("no SASS rows")
print("columns:", [h for h in hdr][:12])
import collections
hist = collections.Counter(); samp = collections.Counter(); n = collections.Counter(); thr = collections.Counter(); fp64 = collections.Counter()
for d in sass:
    c = int(d["Instructions Executed"] or 0)
    hist[c] += c
    samp[c] += int(d["# Samples"] or 0)
    n[c] += 1
    thr[c] += int(d["Thread Instructions Executed"] or 0)
    op = d["Source"].split()[0] if d["Source"].split() else ""
    if op.startswith("@"):
        op = d["Source"].split()[1]
    if op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "MUFU.RSQ64", "DMNMX")):
        fp64[c] += c
tot = sum(hist.values()); ts = sum(samp.values())
print("total warp instructions %d, samples %d" % (tot, ts))
print("%12s %6s %14s %7s %9s %7s %6s" % ("exec count", "#sass", "warp instr", "share", "samples", "share", "fp64"))
for c, v in sorted(hist.items(), key=lambda kv: -kv[1])[:25]:
    print("%12d %6d %14d %6.1f%% %9d %6.1f%% %5.0f%%  thr/inst %.1f" % (c, n[c], v, 100.0 * v / tot, samp[c], 100.0 * samp[c] / max(ts, 1), 100.0 * fp64[c] / max(v, 1), thr[c] / max(v, 1)))
