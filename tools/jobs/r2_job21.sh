#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
export RK_TRAJECTORY_LOG=gpurun_out/r2/trajectory.jsonl; rm -f $RK_TRAJECTORY_LOG
timeout 1500 python -m pytest tests -q --no-header -p no:cacheprovider -m gpu -x --deselect tests/test_drp.py 2>&1 | tail -n 15 | cut -c1-300
for w in wildfire; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench_${w}21.json 2> gpurun_out/r2/bench_${w}21.err
tail -c 300 gpurun_out/r2/bench_${w}21.err; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_${w}21.json').read().strip().splitlines()[-1]); print('$w', d['ms_per_step'], d.get('kernels_ms'), d['roofline']['frac'], d['program'])"
done
