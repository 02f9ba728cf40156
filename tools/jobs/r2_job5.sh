#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
export RK_TRAJECTORY_LOG=gpurun_out/r2/trajectory.jsonl
rm -f $RK_TRAJECTORY_LOG
timeout 600 python -m pytest tests/test_trajectory_gpu.py -q --no-header -p no:cacheprovider -k "wildfire or powergen" 2>&1 | grep -E "AssertionError|passed|failed" | head
timeout 300 python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > gpurun_out/r2/plain_drp.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r2/launches_drp2.csv python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > gpurun_out/r2/ncu_drp2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk_mlp2 -s 6 -c 2 -o gpurun_out/r2/ncu_mlp2 -f python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > gpurun_out/r2/ncu_mlp2.log 2>&1
timeout 300 python tools/profile_workload.py --workload marsrover --reps 3 > gpurun_out/r2/plain_mr.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rk_spec -s 3 -c 3 -o gpurun_out/r2/ncu_marsrover_tpr -f python tools/profile_workload.py --workload marsrover --reps 3 > gpurun_out/r2/ncu_mr.log 2>&1
