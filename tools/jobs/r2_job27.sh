#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_drp.py -q --no-header -p no:cacheprovider -k "fused_mlp2 or wgrad" 2>&1 | tail -n 3
timeout 300 python tools/mlp2_phases.py > gpurun_out/r2/mlp2_phases.log 2>&1; cat gpurun_out/r2/mlp2_phases.log
timeout 600 python bench.py --workload powergen_drp --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_drp27.json 2> gpurun_out/r2/bench_drp27.err
tail -c 300 gpurun_out/r2/bench_drp27.err; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_drp27.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('kernels_ms'))"
