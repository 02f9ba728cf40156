"""torchrun check of the one-shot NVLink peer all-reduce (csrc/rk_comm.cu) against NCCL:
eager calls, calls replayed from a CUDA graph, identical bits on every rank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pyrddlgym_jax_b200.core import parallel

os.environ.setdefault("NCCL_DEBUG", "WARN")
rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 1602
ar = parallel.PeerAllReduce.create(n, dev)
assert ar is not None, "symmetric memory unavailable"
g = torch.Generator(device=dev).manual_seed(100 + rank)
ok = True
for it in range(20):
    v = torch.randn(n, device=dev, generator=g)
    ref = v.clone()
    dist.all_reduce(ref)
    ar(v)
    torch.cuda.synchronize()
    ok &= torch.allclose(v, ref, rtol=1e-6, atol=1e-6)
    ok &= parallel.replicas_identical(v)
# graph replay
buf = torch.zeros(n, device=dev)
src = torch.randn(n, device=dev, generator=g)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    buf.copy_(src); ar(buf)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    buf.copy_(src); ar(buf)
ref = src.clone(); dist.all_reduce(ref)
for _ in range(50):
    graph.replay()
torch.cuda.synchronize()
ok &= torch.allclose(buf, ref, rtol=1e-6, atol=1e-6) and int(ar.status.item()) == 0
# latency
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    graph.replay()
e1.record(); torch.cuda.synchronize()
t_peer = e0.elapsed_time(e1) * 1e3 / 200
graph2 = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    buf.copy_(src); dist.all_reduce(buf)
torch.cuda.synchronize()
with torch.cuda.graph(graph2):
    buf.copy_(src); dist.all_reduce(buf)
e0.record()
for _ in range(200):
    graph2.replay()
e1.record(); torch.cuda.synchronize()
t_nccl = e0.elapsed_time(e1) * 1e3 / 200
if rank == 0:
    print(f"peer all-reduce ok={bool(ok)} world={dist.get_world_size()} n={n}: "
          f"{t_peer:.1f} us per copy+allreduce vs NCCL {t_nccl:.1f} us", flush=True)
dist.barrier()
os._exit(0 if ok else 1)
