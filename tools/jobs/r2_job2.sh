#!/bin/bash
# round 2, job 2: new parity tests (ops/samplers fixtures, optimiser, trajectory, element-wise tolerances)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
export RK_TRAJECTORY_LOG=gpurun_out/r2/trajectory.jsonl
rm -f $RK_TRAJECTORY_LOG
timeout 900 python -m pytest tests/test_optimizer_gpu.py tests/test_trajectory_gpu.py -q -x --no-header -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2/t_new.log
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_golden.py tests/test_zz_gamma_fluent_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r2/t_kernels.log
tail -5 gpurun_out/r2/t_new.log gpurun_out/r2/t_kernels.log
