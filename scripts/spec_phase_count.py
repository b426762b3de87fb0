#!/usr/bin/env python
"""Static instruction count per phase of a generated kernel: `python scripts/spec_phase_count.py <hash>.cubin [kernel]`.
Attributes every SASS instruction of the kernel to the generated-source line nvdisasm reports (inlined pm_device.cuh code is
attributed to the line it was inlined at) and sums per phase (star / axis / pairs blocks, epilogue, prologue)."""
import re
import subprocess
import sys

cubin = sys.argv[1]
kernel = sys.argv[2] if len(sys.argv) > 2 else "pm_ef_spec_kernel"
src = open(cubin.replace(".cubin", ".cu")).read().splitlines()
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, universal_newlines=True).stdout.splitlines()

# phase of every source line
phase_of = {}
cur = "other"
in_kernel = False
for n, line in enumerate(src, 1):
    if "__global__" in line or line.startswith("extern \"C\" __global__"):
        cur = "prologue"
    m = re.search(r"\{  // (star|axis|pairs)", line)
    if m:
        cur = m.group(1)
    if "const long sconf" in line:
        cur = "epilogue"
    phase_of[n] = cur

counts = {}
fp = {}
active = False
line_no = None
for t in txt:
    if t.startswith("\t.section") or ".text." in t and t.strip().startswith(".section"):
        active = (".text." + kernel) in t
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', t)
    if m:
        if m.group(1).endswith(".cu"):
            line_no = int(m.group(2))
        else:
            m2 = re.search(r'inlined at "[^"]+\.cu", line (\d+)', t)
            if m2:
                line_no = int(m2.group(1))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", t)
    if m:
        op = m.group(2)
        ph = phase_of.get(line_no, "other")
        counts[ph] = counts.get(ph, 0) + 1
        if op.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX", "MUFU"):
            k = op.split(".")[0] if not op.startswith("MUFU") else op
            fp.setdefault(ph, {})
            fp[ph][k] = fp[ph].get(k, 0) + 1
tot = sum(counts.values())
print("kernel %s: %d SASS instructions" % (kernel, tot))
for ph, c in sorted(counts.items(), key=lambda kv: -kv[1]):
    d = fp.get(ph, {})
    n64 = sum(v for k, v in d.items() if k[0] == "D")
    print("%-9s %6d (%4.1f%%)  fp64 %5d  %s" % (ph, c, 100.0 * c / tot, n64, " ".join("%s=%d" % kv for kv in sorted(d.items()))))
