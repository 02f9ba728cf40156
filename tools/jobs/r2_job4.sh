#!/bin/bash
# round 2, job 4: fused DRP forward (kernel test, planner-level DRP tests, bench), fixed optimiser / trajectory tests
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
export RK_TRAJECTORY_LOG=gpurun_out/r2/trajectory.jsonl
rm -f $RK_TRAJECTORY_LOG
timeout 300 python -m pytest tests/test_drp.py -q --no-header -p no:cacheprovider -x -k "fused_mlp2" > gpurun_out/r2/t_mlp2.log 2>&1
tail -n 30 gpurun_out/r2/t_mlp2.log
timeout 900 python -m pytest tests/test_optimizer_gpu.py tests/test_trajectory_gpu.py -q --no-header -p no:cacheprovider > gpurun_out/r2/t_new.log 2>&1
tail -n 30 gpurun_out/r2/t_new.log
timeout 900 python -m pytest tests/test_drp.py tests/test_golden_drp.py -q --no-header -p no:cacheprovider -m gpu > gpurun_out/r2/t_drp.log 2>&1
tail -n 30 gpurun_out/r2/t_drp.log
timeout 600 python bench.py --workload powergen_drp --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_drp.json 2> gpurun_out/r2/bench_drp.err
tail -c 1000 gpurun_out/r2/bench_drp.err; cat gpurun_out/r2/bench_drp.json | head -c 1500
timeout 600 python bench.py --workload wildfire --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_wf.json 2> gpurun_out/r2/bench_wf.err
