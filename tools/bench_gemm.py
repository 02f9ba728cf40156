"""Times the hidden-layer GEMM variants (SIMT fp32, tf32 split, tcgen05 3xTF32) with CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyrddlgym_jax_b200.core import engine

dev = torch.device("cuda:0")
M, N, K = int(os.environ.get("M", 32768)), 256, 256
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(M, K, device=dev, generator=g) / 16
W = torch.randn(K, N, device=dev, generator=g)
Wt = W.t().contiguous()
b = torch.randn(N, device=dev, generator=g)
C = torch.empty(M, N, device=dev)
Ah, Al, Bh, Bl = torch.empty_like(A), torch.empty_like(A), torch.empty_like(Wt), torch.empty_like(Wt)
engine.tf32_split(Wt, Bh, Bl)
flush = torch.empty(64 * 1024 * 1024, device=dev)


def t(fn, reps=20):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


print("sgemm      us", t(lambda: engine.sgemm(engine.EPI_BIAS_TANH, M, N, K, A, K, W, N, C, N, bias=b)))
print("tf32_split us", t(lambda: engine.tf32_split(A, Ah, Al)))
print("tcgemm     us", t(lambda: engine.tcgemm(2, M, N, K, A, K, Bh, Bl, K, C, N, bias=b)))
print("copy 33MB  us", t(lambda: C.copy_(A)))
