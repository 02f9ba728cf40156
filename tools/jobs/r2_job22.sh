#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
export RK_TRAJECTORY_LOG=gpurun_out/r2/trajectory.jsonl; rm -f $RK_TRAJECTORY_LOG
timeout 900 python -m pytest tests/test_trajectory_gpu.py tests/test_kernels_gpu.py tests/test_golden.py -q --no-header -p no:cacheprovider -m gpu 2>&1 | grep -E "AssertionError: |passed|failed|FAILED" | cut -c1-400
timeout 600 python bench.py --workload wildfire --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench_wildfire22.json 2> gpurun_out/r2/bench_wildfire22.err
tail -c 300 gpurun_out/r2/bench_wildfire22.err; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_wildfire22.json').read().strip().splitlines()[-1]); print('wildfire', d['ms_per_step'], d.get('kernels_ms'), d['roofline']['frac'])"
timeout 300 python tools/profile_workload.py --workload wildfire --reps 3 > gpurun_out/r2/plain_wildfire.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rk_spec -s 3 -c 3 -o gpurun_out/r2/ncu_wildfire_ell -f python tools/profile_workload.py --workload wildfire --reps 3 > gpurun_out/r2/ncu_wildfire_ell.log 2>&1
