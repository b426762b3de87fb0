#!/bin/bash
# A/B of team-mode kernel variants on one box: scripts/ab_team.sh  (prints Mconf/s for lig60 / lig50 / norfloxacin)
run() {  # label, env...
  label=$1; shift
  for spec in "lig60 4" "lig60 8" "lig50 8" "lig50 4" "norfloxacin 8"; do
    set -- $spec
    env "${EXTRA[@]}" python - "$1" "$2" "$label" <<'PY'
import sys, os
sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
from gpu_sweep import run
mol, cpw, label = sys.argv[1], int(sys.argv[2]), sys.argv[3]
sizes = {"lig60": 1 << 17, "lig50": 1 << 18, "norfloxacin": 1 << 18}
r = run(mol, sizes[mol], cpw)
print("%-28s %-12s cpw=%-2d warps=%-2d ms=%.4f Mconf/s=%.1f" % (label, mol, r["cpw"], r["warps"], r["ms"], r["Mconf_s"]), flush=True)
PY
  done
}
EXTRA=(); run "roll1 maxw-auto"
EXTRA=(PM_MAXW=16); run "roll1 maxw16"
EXTRA=(PARAMOL_B200_LIB=build/libu4.so); run "unroll4 maxw-auto"
EXTRA=(PARAMOL_B200_LIB=build/libu4.so PM_MAXW=16); run "unroll4 maxw16"
