#!/bin/bash
# lane-teams vs warp-teams (PM_CPW=-8) on one box
for spec in "lig60 4" "lig60 -8" "lig50 8" "lig50 -8" "norfloxacin 8" "norfloxacin -8" "lig33 8" "lig33 -8" "caffeine 16" "caffeine -8"; do
  set -- $spec
  PARAMOL_B200_SPECIALIZE=0 python - "$1" "$2" <<'PY'
import sys, os
sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
from gpu_sweep import run
mol, cpw = sys.argv[1], int(sys.argv[2])
sizes = {"lig60": 1 << 17, "lig50": 1 << 18, "norfloxacin": 1 << 18, "lig33": 1 << 18, "caffeine": 1 << 19}
r = run(mol, sizes[mol], cpw)
print("%-12s cpw=%-3d L=%d warps=%-2d ms=%.4f Mconf/s=%.1f" % (mol, cpw, r["L"], r["warps"], r["ms"], r["Mconf_s"]), flush=True)
PY
done
