#!/bin/bash
# Round-2 (second session) GPU call 6: the bench line of the final build + ncu launch lists of the same command.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( timeout 330 python bench.py --steps 10 --warmup 3 ) > $O/g6_bench_c5.json 2> $O/g6_bench_c5.err
echo "bench rc=$?"
( ICNN_BENCH_GRAPH=0 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1200 --csv --log-file $O/g6_launches_C5.csv python bench.py --workload C5 --steps 1 --warmup 1 --no-sub --no-cpu-baseline ) > $O/g6_ncu_c5.log 2>&1
( ICNN_BENCH_GRAPH=0 timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 800 --csv --log-file $O/g6_launches_T.csv python bench.py --workload T --steps 2 --warmup 1 --no-sub --no-cpu-baseline ) > $O/g6_ncu_t.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g6_bench_c5.json').read().strip().splitlines()[-1])
print('C5', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], d['loop_graph'])
for k,v in d['configs'].items(): print(k, v['value'], v['ms_per_step'], 'e2e', v['e2e']['ms_per_step'], v.get('loop_graph'))
PY
tail -c 400 $O/g6_bench_c5.err
wc -l $O/g6_launches_C5.csv $O/g6_launches_T.csv
