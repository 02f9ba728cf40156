#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
export RK_TRAJECTORY_LOG=gpurun_out/r2/trajectory.jsonl; rm -f $RK_TRAJECTORY_LOG
timeout 1500 python -m pytest tests/test_optimizer_gpu.py tests/test_drp.py tests/test_golden_drp.py tests/test_trajectory_gpu.py tests/test_planner_api.py -q --no-header -p no:cacheprovider -m gpu 2>&1 | grep -E "AssertionError|passed|failed|FAILED|Error" | cut -c1-300
for w in wildfire powergen_drp; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench_${w}26.json 2> gpurun_out/r2/bench_${w}26.err
tail -c 300 gpurun_out/r2/bench_${w}26.err; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_${w}26.json').read().strip().splitlines()[-1]); print('$w', d['ms_per_step'], d['e2e']['iters_per_sec'], d.get('kernels_ms'))"
done
