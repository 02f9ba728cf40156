#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 600 python tools/drp_grad_debug.py 2>&1 | grep -E "train|W0 |bh |head product"
export RK_TRAJECTORY_LOG=gpurun_out/r2/trajectory.jsonl; rm -f $RK_TRAJECTORY_LOG
timeout 900 python -m pytest tests/test_drp.py tests/test_golden_drp.py tests/test_trajectory_gpu.py -q --no-header -p no:cacheprovider -m gpu 2>&1 | grep -E "AssertionError|passed|failed|FAILED|^E  " | cut -c1-300
timeout 300 python tools/mlp2_phases.py > gpurun_out/r2/mlp2_phases.log 2>&1; cat gpurun_out/r2/mlp2_phases.log
timeout 600 python bench.py --workload powergen_drp --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_drp19.json 2> gpurun_out/r2/bench_drp19.err
tail -c 600 gpurun_out/r2/bench_drp19.err; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_drp19.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('kernels_ms'))"
