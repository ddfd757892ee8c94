#!/bin/bash
# K2 configuration sweep (exploration): per-config total K1/K2 ms from tools/iter_profile.py
for spec in "T 1" "T 2" "T 4" "C3 1" "C3 2" "C2 8" "C2 16" ; do
  set -- $spec
  echo "== $1 ICNN_PC_WPS=$2"; ICNN_PC_WPS=$2 python tools/iter_profile.py $1 2>&1 | grep -E "total|t= 0|t= 9"
done
echo "== legacy"; for W in T C3 C2; do ICNN_K2_PC=legacy python tools/iter_profile.py $W 2>&1 | grep -E "total"; done
