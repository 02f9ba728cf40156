#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_drp.py -q --no-header -p no:cacheprovider -m gpu -k "activation" 2>&1 | grep -E "Error|passed|failed|FAILED|^E  " | cut -c1-400 | head -30
