#!/usr/bin/env python
"""Per-source-line stall samples from an .ncu-rep: `python scripts/ncu_lines.py rep.ncu-rep [top]`."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, universal_newlines=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r; continue
    if cur and hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = {}
        for k, v in zip(hdr, r):
            if k not in d: d[k] = v
        try: s = int(d.get("# Samples", "0"))
        except ValueError: s = 0
        out.append((s, cur, d))
tot = sum(s for s, _, _ in out)
out.sort(key=lambda t: -t[0])
print("total samples", tot)
keys = ["stall_wait", "stall_short_sb", "stall_long_sb", "stall_math", "stall_not_selected", "stall_branch_resolving", "stall_barrier", "stall_dispatch", "stall_mio", "stall_no_inst"]
for s, f, d in out[:top]:
    st = " ".join("%s=%s" % (k.replace("stall_", ""), d.get(k, "")) for k in keys if d.get(k, "0") not in ("0", ""))
    print("%5.1f%% %s:%s inst=%s | %s\n        %s" % (100.0 * s / max(tot, 1), f, d["Line No"], d.get("Instructions Executed", ""), d["Source"].strip()[:100], st))
