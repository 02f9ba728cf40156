#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_optimizer_gpu.py -q --no-header -p no:cacheprovider -m gpu -k "chain_step" 2>&1 | grep -E "Error|passed|failed|FAILED|^E  " | cut -c1-400 | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/check_multi_gpu_utility.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -15
