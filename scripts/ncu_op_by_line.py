#!/usr/bin/env python
"""Executed count of one SASS opcode by CUDA source line: `python scripts/ncu_op_by_line.py rep.ncu-rep IMAD [top]`."""
import csv, subprocess, sys, collections
rep, op = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, universal_newlines=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = None; hdr = None; cur_line = None; agg = collections.Counter(); ex = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if not hdr or len(r) != len(hdr): continue
    if r[0] not in ("", "Line No"):
        cur_line = (cur_file, r[0], r[1].strip()[:90]); continue
    sass = r[3].strip().split()
    if not sass: continue
    o = sass[1] if sass[0].startswith("@") and len(sass) > 1 else sass[0]
    if o.split(".")[0] == op:
        try: n = int(r[hdr.index("Instructions Executed")])
        except ValueError: n = 0
        agg[cur_line] += n
        ex.setdefault(cur_line, r[3].strip()[:70])
tot = sum(agg.values()); print(op, "total", tot)
for k, n in agg.most_common(top):
    print("%5.1f%% %s:%s | %s   e.g. %s" % (100.0 * n / max(tot, 1), k[0], k[1], k[2], ex[k]))
