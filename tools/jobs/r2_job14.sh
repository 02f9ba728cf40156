#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python tools/drp_grad_debug.py 2>&1 | tail -n 8
export RK_TRAJECTORY_LOG=gpurun_out/r2/trajectory.jsonl; rm -f $RK_TRAJECTORY_LOG
timeout 900 python -m pytest tests/test_drp.py tests/test_golden_drp.py tests/test_trajectory_gpu.py -q --no-header -p no:cacheprovider -m gpu 2>&1 | grep -E "AssertionError|passed|failed|FAILED|^E  " | cut -c1-300
