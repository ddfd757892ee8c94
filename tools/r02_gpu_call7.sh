#!/bin/bash
# Round-2 (third session) GPU call 7 (the round's last 3 GPU-minutes): the bench.py instrumented pass with iteration statistics
# off in the timed passes (C3, where the atomics distorted K2) and the three GPU tests against the reference-graph goldens.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( timeout 100 python bench.py --workload C3 --steps 3 --warmup 3 --no-sub --no-cpu-baseline ) > $O/g7_bench_c3.json 2> $O/g7_bench_c3.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g7_bench_c3.json').read().strip().splitlines()[-1])
k=d['kernels']['K2_bundle_step']
print('C3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'K2 ms/launch', k['ms_per_launch'], 'with stats', k['ms_per_launch_with_statistics'], 'frac', k['frac'])
PY
( timeout 80 python -m pytest tests/test_gpu_picnn.py -q -x -m gpu -k "reference_graph_golden" ) > $O/g7_pytest.txt 2>&1
echo "pytest rc=$?"
tail -3 $O/g7_pytest.txt
tail -c 300 $O/g7_bench_c3.err
