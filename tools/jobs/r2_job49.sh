#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/check_multi_gpu_utility.py > gpurun_out/r2/mgu.log 2>&1
echo rc=$?; grep -v "^\*\|OMP_NUM" gpurun_out/r2/mgu.log | grep "rank\|ok\|Error" | tail -30
