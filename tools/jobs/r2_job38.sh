#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
for r in "" "backward=3" "backward=4" "3" "forward=3,backward=4"; do
export RK_TPR_MINBLOCKS="$r"
timeout 300 python bench.py --workload marsrover --headline-only --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench_mr38.json 2> gpurun_out/r2/bench_mr38.err
tail -c 200 gpurun_out/r2/bench_mr38.err; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_mr38.json').read().strip().splitlines()[-1]); print('minblocks=[$r]', round(d['ms_per_step'],4), d.get('kernels_ms'))"
done
