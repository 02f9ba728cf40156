#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests -x -q --no-header -p no:cacheprovider -m gpu 2>&1 | tail -4 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --workload quadcopter --headline-only --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('quadcopter', d['ms_per_step'])"
