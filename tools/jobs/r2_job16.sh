#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
RK_DRP_FUSED_FWD=0 timeout 600 python tools/drp_grad_debug.py 2>&1 | tail -n 5
