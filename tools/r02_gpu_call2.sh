#!/bin/bash
# Round-2 (second session) GPU call 2: K1 chunk-length A/B (accuracy + time), ncu source-level captures of K1 (T) and of
# the three-vector K2 (C5).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for ch in 1 2 4; do
  ( ICNN_TC_CH=$ch timeout 120 python -m pytest tests/test_gpu_picnn.py -q -m gpu -s -k "long_reductions or (fg_matches and C2)" ) > $O/g2_acc_ch$ch.txt 2>&1
done
for ch in 2 4; do
  ( ICNN_TC_CH=$ch timeout 100 python tools/iter_profile.py T ) > $O/g2_t_ch$ch.txt 2>&1
  ( ICNN_TC_CH=$ch timeout 150 python tools/iter_profile.py C5 pc 2048 ) > $O/g2_c5_ch$ch.txt 2>&1
done
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 24 -c 4 -o $O/g2_k1_T python tools/iter_profile.py T ) > $O/g2_ncu_k1.log 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:bundle_pc -s 30 -c 1 -o $O/g2_k2_C5 python tools/iter_profile.py C5 pc 2048 ) > $O/g2_ncu_k2.log 2>&1
grep -h "C5 f:\|passed\|failed" $O/g2_acc_ch*.txt
grep -h "total" $O/g2_t_ch*.txt $O/g2_c5_ch*.txt
ls -la $O/*.ncu-rep
tail -2 $O/g2_ncu_k1.log $O/g2_ncu_k2.log
