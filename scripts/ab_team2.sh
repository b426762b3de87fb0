#!/bin/bash
# A/B of team-mode kernel variants given as "label:ENV=.. ENV=.." arguments: scripts/ab_team2.sh "base:" "u2:PARAMOL_B200_LIB=build/libu2.so"
for arg in "$@"; do
  label=${arg%%:*}; envs=${arg#*:}
  for spec in "lig60 4" "lig50 8" "norfloxacin 8"; do
    set -- $spec
    env $envs python - "$1" "$2" "$label" <<'PY'
import sys, os
sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
from gpu_sweep import run
mol, cpw, label = sys.argv[1], int(sys.argv[2]), sys.argv[3]
sizes = {"lig60": 1 << 17, "lig50": 1 << 18, "norfloxacin": 1 << 18}
r = run(mol, sizes[mol], cpw)
print("%-28s %-12s cpw=%-2d warps=%-2d ms=%.4f Mconf/s=%.1f" % (label, mol, r["cpw"], r["warps"], r["ms"], r["Mconf_s"]), flush=True)
PY
  done
done
