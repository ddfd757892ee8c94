#!/bin/bash
# Round-2 (second session) GPU call 5: the whole GPU suite on the final build, then the C5 A/B of the two exploration flags.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( timeout 780 python -m pytest tests -q -m gpu --durations=12 ) > $O/g5_pytest_all.txt 2>&1
echo "rc=$?" >> $O/g5_pytest_all.txt
for fl in 0 1 2 3; do
  ( ICNN_PC_FLAGS=$fl timeout 100 python tools/iter_profile.py C5 pc 1024 ) > $O/g5_c5_flags$fl.txt 2>&1
done
tail -25 $O/g5_pytest_all.txt
grep -h "total" $O/g5_c5_flags*.txt
