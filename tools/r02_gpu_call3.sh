#!/bin/bash
# Round-2 (second session) GPU call 3: K1 after the MMA-issue fix (uniform TMEM base + elect.sync): parity, A/B of the
# tile variants, ncu re-capture.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( timeout 400 python -m pytest tests/test_gpu_picnn.py tests/test_gpu_gd_grad.py -x -q -m gpu -s ) > $O/g3_pytest_k1.txt 2>&1
echo "rc=$?" >> $O/g3_pytest_k1.txt
( timeout 100 python tools/iter_profile.py T ) > $O/g3_t_def.txt 2>&1
( ICNN_TC_CFG=128 timeout 100 python tools/iter_profile.py T ) > $O/g3_t_tc128.txt 2>&1
( ICNN_TC_CFG=64x4 timeout 100 python tools/iter_profile.py T ) > $O/g3_t_tc644.txt 2>&1
( timeout 150 python tools/iter_profile.py C5 pc 2048 ) > $O/g3_c5_def.txt 2>&1
( ICNN_TC_CFG=128 timeout 150 python tools/iter_profile.py C5 pc 2048 ) > $O/g3_c5_tc128.txt 2>&1
( timeout 100 python tools/iter_profile.py C2 ) > $O/g3_c2_def.txt 2>&1
( timeout 100 python tools/iter_profile.py C3 ) > $O/g3_c3_def.txt 2>&1
( timeout 100 python tools/iter_profile.py C4 newton ) > $O/g3_c4_def.txt 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 24 -c 4 -o $O/g3_k1_T python tools/iter_profile.py T ) > $O/g3_ncu_k1.log 2>&1
tail -4 $O/g3_pytest_k1.txt
grep -h "C5 f:" $O/g3_pytest_k1.txt
grep -h "total" $O/g3_t_def.txt $O/g3_t_tc128.txt $O/g3_t_tc644.txt $O/g3_c5_def.txt $O/g3_c5_tc128.txt $O/g3_c2_def.txt $O/g3_c3_def.txt $O/g3_c4_def.txt
ls -la $O/g3*.ncu-rep
