#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
for B in 32768 131072; do
for L in 1 4 8 16 32; do
export RK_LANES=$L
timeout 300 python bench.py --workload marsrover --batch $B --headline-only --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench_mr41.json 2> gpurun_out/r2/bench_mr41.err
tail -c 200 gpurun_out/r2/bench_mr41.err | grep -i error; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_mr41.json').read().strip().splitlines()[-1]); print('B=$B L=$L', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d.get('kernels_ms').items()}, d['program']['lanes'], d['program']['rpw'])"
done
done
