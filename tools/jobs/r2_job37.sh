#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_optimizer_gpu.py tests/test_planner_api.py tests/test_trajectory_gpu.py -q --no-header -p no:cacheprovider -m gpu 2>&1 | grep -E "Error|passed|failed|FAILED|^E  " | cut -c1-300 | head -30
timeout 600 python bench.py --workload reservoir --headline-only --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2/bench_reservoir37.json 2> gpurun_out/r2/bench_reservoir37.err
tail -c 300 gpurun_out/r2/bench_reservoir37.err; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_reservoir37.json').read().strip().splitlines()[-1]); print('reservoir', d['ms_per_step'], d['e2e']['iters_per_sec'], d.get('kernels_ms'))"
