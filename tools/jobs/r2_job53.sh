#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 200 python -m pytest tests/test_simulator_gpu.py -q --no-header -p no:cacheprovider -m gpu -k "tuning" 2>&1 | grep -E "Error|passed|failed|FAILED|^E  " | cut -c1-400 | head -20
