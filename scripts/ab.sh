#!/bin/bash
# A/B helper: scripts/ab.sh "<lib> <cpw> <molecule>" ...   (run on the GPU box)
for spec in "$@"; do
  set -- $spec
  lib=$1; cpw=$2; mol=$3
  if [ "$lib" = "default" ]; then unset PARAMOL_B200_LIB; else export PARAMOL_B200_LIB=$lib; fi
  PM_CPW=$cpw python - "$mol" "$cpw" "$lib" <<'PY'
import sys, json, os
sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
from gpu_sweep import run
mol, cpw, lib = sys.argv[1], int(sys.argv[2]), sys.argv[3]
sizes = {"aniline": 1 << 20, "caffeine": 1 << 19, "norfloxacin": 1 << 18, "lig50": 1 << 18, "lig60": 1 << 17}
r = run(mol, sizes.get(mol, 1 << 17), cpw)
print("%-22s %-12s cpw=%-2d warps=%-2d ms=%.4f Mconf/s=%.1f" % (lib.split("/")[-1], mol, r["cpw"], r["warps"], r["ms"], r["Mconf_s"]))
PY
done
