#!/bin/bash
# round 2, job 3: optimiser / trajectory / thread-per-rollout tests, the new bench line at N=1, whole GPU suite
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
export RK_TRAJECTORY_LOG=gpurun_out/r2/trajectory.jsonl
rm -f $RK_TRAJECTORY_LOG
timeout 900 python -m pytest tests/test_optimizer_gpu.py tests/test_trajectory_gpu.py -q --no-header -p no:cacheprovider > gpurun_out/r2/t_new.log 2>&1
tail -n 30 gpurun_out/r2/t_new.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --no-header -p no:cacheprovider -k "thread_per_rollout or large_batch" > gpurun_out/r2/t_tpr.log 2>&1
tail -n 30 gpurun_out/r2/t_tpr.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_n1.json 2> gpurun_out/r2/bench_n1.err
tail -c 1500 gpurun_out/r2/bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2/bench_ref.json 2> gpurun_out/r2/bench_ref.err
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2/t_all.log 2>&1
tail -n 25 gpurun_out/r2/t_all.log
