#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_golden.py tests/test_trajectory_gpu.py -q --no-header -p no:cacheprovider -m gpu 2>&1 | grep -E "AssertionError: |passed|failed|FAILED" | cut -c1-300
timeout 300 python bench.py --workload marsrover --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench_mr25.json 2> gpurun_out/r2/bench_mr25.err
python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_mr25.json').read().strip().splitlines()[-1]); print('marsrover', d['ms_per_step'], d.get('kernels_ms'))"
timeout 300 python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > gpurun_out/r2/plain_drp.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rk_mlp2_kernel|rk_wgrad_kernel" -s 18 -c 12 -o gpurun_out/r2/ncu_mlp2_v5 -f python tools/profile_workload.py --workload powergen_drp --horizon 4 --reps 2 > gpurun_out/r2/ncu_mlp2_v5.log 2>&1
tail -2 gpurun_out/r2/ncu_mlp2_v5.log
