#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
export RK_TRAJECTORY_LOG=gpurun_out/r2/trajectory.jsonl; rm -f $RK_TRAJECTORY_LOG
timeout 900 python -m pytest tests/test_drp.py tests/test_golden_drp.py tests/test_trajectory_gpu.py -q --no-header -p no:cacheprovider -m gpu 2>&1 | grep -E "AssertionError|passed|failed|FAILED|^E  " | cut -c1-300
timeout 600 python bench.py --workload powergen_drp --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_drp20.json 2> gpurun_out/r2/bench_drp20.err
tail -c 600 gpurun_out/r2/bench_drp20.err; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_drp20.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('kernels_ms'), d['roofline']['frac'])"
timeout 300 python tools/profile_workload.py --workload powergen_drp --horizon 100 --reps 1 > gpurun_out/r2/plain_drp.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/r2/launches_drp20.csv python tools/profile_workload.py --workload powergen_drp --horizon 100 --reps 1 > gpurun_out/r2/ncu_drp20.log 2>&1
