#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_optimizer_gpu.py tests/test_planner_api.py tests/test_trajectory_gpu.py -q --no-header -p no:cacheprovider -m gpu 2>&1 | grep -E "Error|passed|failed|FAILED|^E  " | cut -c1-300 | head -30
for w in reservoir quadcopter wildfire marsrover; do
timeout 600 python bench.py --workload $w --headline-only --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2/bench_${w}46.json 2> gpurun_out/r2/bench_${w}46.err
tail -c 300 gpurun_out/r2/bench_${w}46.err | grep -i "error"; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_${w}46.json').read().strip().splitlines()[-1]); print('$w', round(d['ms_per_step'],4), round(d['e2e']['iters_per_sec'],1))"
done
