#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q --no-header -p no:cacheprovider -m gpu 2>&1 | tail -4 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2/bench_final_n1.json 2> gpurun_out/r2/bench_final_n1.err
echo bench rc=$?; tail -c 300 gpurun_out/r2/bench_final_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_final_n1.json').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('traffic'), d['clocks'])
for w in d.get('workloads',[]):
    print(w['name'], round(w['ms_per_step'],4), '%.3g'%w['value'], '%.3g'%w['e2e']['value'], round(w['e2e']['iters_per_sec'],1), round(w['roofline']['frac'],3), '%.3g'%w['cpu_baseline']['value'])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2/bench_final_ref.json 2> gpurun_out/r2/bench_final_ref.err
tail -c 400 gpurun_out/r2/bench_final_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r2/launches_bench_marsrover.csv python bench.py --workload marsrover --headline-only --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu_bench_marsrover.log 2>&1
python tools/launch_summary.py gpurun_out/r2/launches_bench_marsrover.csv | head -14
