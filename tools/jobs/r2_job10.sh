#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 300 python tools/mlp2_phases.py > gpurun_out/r2/mlp2_phases.log 2>&1; cat gpurun_out/r2/mlp2_phases.log
timeout 600 python tools/wgrad_numerics.py > gpurun_out/r2/wgrad_numerics.log 2>&1; cat gpurun_out/r2/wgrad_numerics.log
