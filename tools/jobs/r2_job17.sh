#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
for cfg in "RK_DEBUG_PATCH=heads" "RK_DEBUG_PATCH=h1" "RK_DEBUG_PATCH=h1 RK_DRP_FUSED_FWD=0" "RK_TCGEMM=0"; do
  echo "=== $cfg"
  env $cfg timeout 600 python tools/drp_grad_debug.py 2>&1 | grep -E "train|W0 |bh "
done
