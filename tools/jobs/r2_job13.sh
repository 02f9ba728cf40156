#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for cfg in "" "RK_DRP_FUSED_BWD=0" "RK_DRP_FUSED_FWD=0"; do
  echo "=== $cfg"
  env $cfg timeout 600 python tools/drp_grad_debug.py 2>&1 | tail -n 8
done
