#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_optimizer_gpu.py tests/test_planner_api.py "tests/test_drp.py::test_drp_utility_update_matches_oracle" -q --no-header -p no:cacheprovider -m gpu -x 2>&1 | grep -E "Error|passed|failed|FAILED|^E  " | cut -c1-400 | head -40
