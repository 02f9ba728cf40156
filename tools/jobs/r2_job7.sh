#!/bin/bash
# round 2, job 7: ncu --set full of the fused DRP forward / backward kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 300 python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > gpurun_out/r2/plain_drp.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rk_mlp2 -s 10 -c 3 -o gpurun_out/r2/ncu_mlp2_v2 -f python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > gpurun_out/r2/ncu_mlp2_v2.log 2>&1
tail -3 gpurun_out/r2/ncu_mlp2_v2.log
