#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
export RK_TRAJECTORY_LOG=gpurun_out/r2/trajectory.jsonl; rm -f $RK_TRAJECTORY_LOG
timeout 1800 python -m pytest tests -q --no-header -p no:cacheprovider -m gpu 2>&1 | grep -E "AssertionError: |passed|failed|FAILED" | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_n1_23.json 2> gpurun_out/r2/bench_n1_23.err
tail -c 300 gpurun_out/r2/bench_n1_23.err; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_n1_23.json').read().strip().splitlines()[-1])
for w in d['workloads']: print(w['name'], round(w['ms_per_step'],4), {k: round(v,4) for k,v in w['kernels_ms'].items()}, round(w['roofline']['frac'],4))
print('headline', d['value'], d['ms_per_step'], d['e2e'])"
