#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q --no-header -p no:cacheprovider -m gpu 2>&1 | tail -15 | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2/bench_n1_29.json 2> gpurun_out/r2/bench_n1_29.err
echo bench rc=$?; tail -c 600 gpurun_out/r2/bench_n1_29.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_n1_29.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])
for w in d.get('workloads',[]):
    print(w['config']['workload'][:30], w['ms_per_step'], w['roofline']['frac'], w.get('kernels_ms'))
PY
