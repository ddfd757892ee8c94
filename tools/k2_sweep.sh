#!/bin/bash
for spec in "T 1 2" "T 1 3" "T 2 3" "T 2 2" "T 4 3" "C2 8 3" "C2 8 2" "C2 4 3"; do
  set -- $spec
  echo "== $1 ICNN_PC_GV=$2 MINB=$3"; ICNN_PC_MINB=$3 ICNN_PC_GV=$2 python tools/iter_profile.py $1 2>&1 | grep -E "total"
done
echo "== C5/1024 GV=8 MINB=3 (hoisted loads)"; ICNN_PC_MINB=3 ICNN_PC_GV=8 python tools/iter_profile.py C5 pc 1024 2>&1 | grep -E "total"
