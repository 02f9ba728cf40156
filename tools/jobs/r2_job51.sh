#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_planner_api.py -q --no-header -p no:cacheprovider -m gpu -k "pgpe" 2>&1 | grep -E "Error|passed|failed|FAILED|^E  " | cut -c1-400 | head -20
