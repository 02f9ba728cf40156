#!/bin/bash
# round 2, job 6: fused DRP forward / backward kernels (rk_mlp2.cu)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_drp.py -q --no-header -p no:cacheprovider -x -k "fused_mlp2" > gpurun_out/r2/t_mlp2.log 2>&1
tail -n 40 gpurun_out/r2/t_mlp2.log
timeout 900 python -m pytest tests/test_drp.py tests/test_golden_drp.py -q --no-header -p no:cacheprovider -m gpu > gpurun_out/r2/t_drp.log 2>&1
tail -n 15 gpurun_out/r2/t_drp.log
timeout 600 python bench.py --workload powergen_drp --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_drp6.json 2> gpurun_out/r2/bench_drp6.err
tail -c 600 gpurun_out/r2/bench_drp6.err; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_drp6.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('kernels_ms'))"
timeout 300 python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > gpurun_out/r2/plain_drp.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r2/launches_drp6.csv python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > gpurun_out/r2/ncu_drp6.log 2>&1
