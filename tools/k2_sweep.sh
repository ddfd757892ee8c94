#!/bin/bash
# K2 configuration sweep (exploration): per-config total K1/K2 ms from tools/iter_profile.py
for spec in "T 1 2" "T 2 2" "C3 1 2" ; do
  set -- $spec
  echo "== $1 ICNN_PC_WPS=$2 MINB=$3"; ICNN_PC_MINB=$3 ICNN_PC_WPS=$2 python tools/iter_profile.py $1 2>&1 | grep -E "total|t= 0|t= 9"
done
echo "== default T"; python tools/iter_profile.py T 2>&1 | grep -E "total"
echo "== default C3"; python tools/iter_profile.py C3 2>&1 | grep -E "total"
echo "== default C2"; python tools/iter_profile.py C2 2>&1 | grep -E "total"
