#!/bin/bash
# Round-2 (second session) GPU call 4: lean k x k stage (k <= 32 instantiation) -- parity of every K2 path, then timings
# incl. the two-sweep kernel on the shapes that so far took the five-sweep one.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( timeout 500 python -m pytest tests/test_gpu_bundle.py -x -q -m gpu -k "k2_ or three_vector or golden or resident or thread_per or fused_vs_oracle or loop_graph" ) > $O/g4_pytest_k2.txt 2>&1
echo "rc=$?" >> $O/g4_pytest_k2.txt
( timeout 100 python tools/iter_profile.py T ) > $O/g4_t_def.txt 2>&1
( ICNN_PC_WPS=2 timeout 100 python tools/iter_profile.py T ) > $O/g4_t_pc2.txt 2>&1
( ICNN_PC_WPS=1 timeout 100 python tools/iter_profile.py T ) > $O/g4_t_pc1.txt 2>&1
( ICNN_PC_WPS=4 timeout 100 python tools/iter_profile.py T ) > $O/g4_t_pc4.txt 2>&1
( timeout 100 python tools/iter_profile.py C3 ) > $O/g4_c3_def.txt 2>&1
( timeout 100 python tools/iter_profile.py C2 ) > $O/g4_c2_def.txt 2>&1
( ICNN_PC_WPS=8 timeout 100 python tools/iter_profile.py C2 ) > $O/g4_c2_pc8.txt 2>&1
( ICNN_PC_V3=1 timeout 100 python tools/iter_profile.py C2 ) > $O/g4_c2_v3.txt 2>&1
( timeout 150 python tools/iter_profile.py C5 pc 2048 ) > $O/g4_c5_def.txt 2>&1
tail -3 $O/g4_pytest_k2.txt
grep -h "total" $O/g4_t_def.txt $O/g4_t_pc2.txt $O/g4_t_pc1.txt $O/g4_t_pc4.txt $O/g4_c3_def.txt $O/g4_c2_def.txt $O/g4_c2_pc8.txt $O/g4_c2_v3.txt $O/g4_c5_def.txt
