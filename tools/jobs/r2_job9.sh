#!/bin/bash
# round 2, job 9: ncu --set full of the fused DRP kernels v2 + weight-gradient kernels; DRP tests
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 300 python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > gpurun_out/r2/plain_drp.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rk_mlp2_kernel|rk_wgrad_kernel" -s 18 -c 12 -o gpurun_out/r2/ncu_mlp2_v3 -f python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > gpurun_out/r2/ncu_mlp2_v3.log 2>&1
tail -3 gpurun_out/r2/ncu_mlp2_v3.log
timeout 900 python -m pytest tests/test_drp.py tests/test_golden_drp.py tests/test_trajectory_gpu.py -q --no-header -p no:cacheprovider -m gpu > gpurun_out/r2/t_drp.log 2>&1
tail -n 15 gpurun_out/r2/t_drp.log
