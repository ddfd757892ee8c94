#!/bin/bash
# Round-2 (second session) GPU call 1: parity of the new K2 build / loop graph, then A/B timings.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/g1_smi.txt 2>&1
timeout 420 python -m pytest tests/test_gpu_bundle.py -x -q -m gpu -k "three_vector or loop_graph or k2_pc_matches or state_reuse" > $O/g1_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/g1_pytest.txt
run() { # name, env..., -- args
  local tag=$1; shift
  ( env "$@" ) > /dev/null 2>&1
}
( timeout 150 python tools/iter_profile.py C5 pc 2048 ) > $O/g1_c5_v3.txt 2>&1
( ICNN_PC_V3=0 timeout 150 python tools/iter_profile.py C5 pc 2048 ) > $O/g1_c5_v4.txt 2>&1
( ICNN_TC_CFG=128 ICNN_PC_V3=0 timeout 150 python tools/iter_profile.py C5 pc 2048 ) > $O/g1_c5_tc128.txt 2>&1
( timeout 100 python tools/iter_profile.py C2 ) > $O/g1_c2_def.txt 2>&1
( ICNN_PC_V3=1 timeout 100 python tools/iter_profile.py C2 ) > $O/g1_c2_v3.txt 2>&1
( timeout 100 python tools/iter_profile.py T ) > $O/g1_t_def.txt 2>&1
( ICNN_TC_CFG=128 timeout 100 python tools/iter_profile.py T ) > $O/g1_t_tc128.txt 2>&1
( timeout 200 python bench.py --workload T --no-sub --steps 5 --warmup 3 --cpu-seconds 2 ) > $O/g1_bench_T.json 2> $O/g1_bench_T.err
tail -3 $O/g1_pytest.txt
grep -h "total" $O/g1_c5_v3.txt $O/g1_c5_v4.txt $O/g1_c5_tc128.txt $O/g1_c2_def.txt $O/g1_c2_v3.txt $O/g1_t_def.txt $O/g1_t_tc128.txt
tail -c 600 $O/g1_bench_T.err
