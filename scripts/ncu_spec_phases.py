#!/usr/bin/env python
"""Executed instructions per phase of a GENERATED kernel from an .ncu-rep (source imported or the .cu beside the cubin):
`python scripts/ncu_spec_phases.py rep.ncu-rep generated.cu` -> per phase (stars / axes / pairs / epilogue / other):
executed warp instructions, the FP64 share, stall samples; per tile iteration of 32 conformations."""
import csv
import re
import subprocess
import sys

rep, cu = sys.argv[1], sys.argv[2]
src = open(cu).read().splitlines()
phase_of = {}
cur = "other"
for n, line in enumerate(src, 1):
    if "__global__" in line:
        cur = "prologue"
    m = re.search(r"\{  // (star|axis|pairs)", line)
    if m:
        cur = m.group(1)
    if "const long sconf" in line:
        cur = "epilogue"
    phase_of[n] = cur
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, universal_newlines=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
line_no = None
n_files = 0
insts = {}   # address -> (phase or None, executed, samples, opcode)
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        n_files += 1   # first section: the generated source; later sections: inlined helpers (phase taken from the neighbours)
        continue
    if len(r) > 3 and r[0] == "Line No":
        hdr = r
        i_addr, i_src = 2, 3
        i_exec = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] != "":
        line_no = int(r[0])
        continue
    try:
        n = int(r[i_exec])
    except ValueError:
        continue
    toks = r[i_src].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    ph = phase_of.get(line_no, "other") if n_files == 1 else None
    addr = int(r[i_addr], 16)
    if addr not in insts or (insts[addr][0] is None and ph is not None):
        insts[addr] = (ph, n, int(r[i_samp]), op.split(".")[0])
agg = {}
cur = "prologue"
for addr in sorted(insts):
    ph, n, smp, base = insts[addr]
    if ph is not None:
        cur = ph
    if n == 0:
        continue
    a = agg.setdefault(cur, dict(inst=0, fp64=0, samples=0, static=0, ops={}))
    a["inst"] += n
    a["static"] += 1
    a["samples"] += smp
    if base in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
        a["fp64"] += n
    a["ops"][base] = a["ops"].get(base, 0) + n
tot = sum(a["inst"] for a in agg.values())
tots = sum(a["samples"] for a in agg.values())
iters = max(a["ops"].get("DFMA", 0) for a in agg.values())
print("total executed warp instructions %d, samples %d" % (tot, tots))
print("%-9s %12s %6s %6s %8s %7s  top opcodes" % ("phase", "warp inst", "share", "fp64", "samples", "static"))
for ph, a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"]):
    top = sorted(a["ops"].items(), key=lambda kv: -kv[1])[:6]
    print("%-9s %12d %5.1f%% %5.1f%% %7.1f%% %7d  %s" % (ph, a["inst"], 100.0 * a["inst"] / tot, 100.0 * a["fp64"] / max(a["inst"], 1),
                                                    100.0 * a["samples"] / max(tots, 1), a["static"],
                                                    " ".join("%s=%.1f%%" % (k, 100.0 * v / a["inst"]) for k, v in top)))
