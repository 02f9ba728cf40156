#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2/bench_final_n${N}.json 2> gpurun_out/r2/bench_final_n${N}.err
tail -c 400 gpurun_out/r2/bench_final_n${N}.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2/bench_final_n${N}.json').read().strip().splitlines() if l.startswith('{')][-1])
print('headline', d['n_gpus'], d['value'], d['ms_per_step'], d['scaling'], d.get('setup'))
for w in d.get('workloads', []): print(w['name'], round(w['ms_per_step'],4), w.get('scaling'), '%.4g' % w['value'], w['setup']['exchange'])"
