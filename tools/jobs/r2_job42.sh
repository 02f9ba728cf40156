#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests/test_drp.py tests/test_golden_drp.py tests/test_trajectory_gpu.py -q --no-header -p no:cacheprovider -m gpu 2>&1 | grep -E "Error|passed|failed|FAILED|^E  " | cut -c1-300 | head -30
