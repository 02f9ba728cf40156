#!/bin/bash
# round 2, job 1: baseline numbers of the five workloads + ncu captures of Wildfire / MarsRover / DRP
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
O=gpurun_out/r2
for w in quadcopter wildfire marsrover powergen_drp reservoir; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/base_$w.json 2> $O/base_$w.err
done
for w in wildfire marsrover; do
  timeout 300 python tools/profile_workload.py --workload $w --reps 3 > $O/plain_$w.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rk_spec -s 3 -c 3 \
      -o $O/ncu_$w -f python tools/profile_workload.py --workload $w --reps 3 > $O/ncu_$w.log 2>&1
done
timeout 300 python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > $O/plain_drp.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file $O/launches_drp.csv python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > $O/ncu_drp.log 2>&1
ls -la $O
