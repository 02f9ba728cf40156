#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 200 python -m pytest tests/test_simulator_gpu.py tests/test_planner_api.py -x -q --no-header -p no:cacheprovider -m gpu 2>&1 | tail -3 | cut -c1-300
