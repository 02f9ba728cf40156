#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 120 python bench.py --workload powergen_drp --headline-only --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('drp', d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'])"
