#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
for w in wildfire quadcopter marsrover reservoir; do
timeout 300 python tools/profile_workload.py --workload $w --reps 3 > gpurun_out/r2/plain_$w.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2/launches_${w}39.csv python tools/profile_workload.py --workload $w --reps 3 > gpurun_out/r2/ncu_${w}39.log 2>&1
echo "== $w"; python tools/launch_summary.py gpurun_out/r2/launches_${w}39.csv | head -16
done
