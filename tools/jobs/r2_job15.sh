#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 600 python tools/drp_grad_debug.py 2>&1 | tail -n 12
