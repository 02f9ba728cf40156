#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r2
N=2
for w in quadcopter marsrover reservoir; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 10 --workload $w --headline-only --no-cpu-baseline > gpurun_out/r2/bench_${w}_n${N}_36.json 2> gpurun_out/r2/bench_${w}_n${N}_36.err
tail -c 300 gpurun_out/r2/bench_${w}_n${N}_36.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2/bench_${w}_n${N}_36.json').read().strip().splitlines() if l.startswith('{')][-1])
print('$w', d['n_gpus'], d['ms_per_step'], d['e2e']['iters_per_sec'], d.get('setup',{}).get('exchange'))"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/check_peer_allreduce.py 2>&1 | tail -8
