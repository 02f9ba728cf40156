#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
for cfg in "" "RK_DRP_FUSED_BWD=0" "RK_DRP_FUSED_FWD=0"; do
  echo "=== $cfg"
  env $cfg timeout 600 python -m pytest tests/test_trajectory_gpu.py -q --no-header -p no:cacheprovider -k "powergen and False" 2>&1 | grep -E "AssertionError: \{|passed|failed" | cut -c1-400
done
