"""Times the DRP streaming kernels alone (CUDA events), cold (rotating buffers larger than L2)
and warm (same buffer): GB/s against the bytes each must move."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyrddlgym_jax_b200.core import engine

B, H, NH, D = 32768, 256, 10, 6
dev = torch.device("cuda:0")
NBUF = 8          # 8 x 33.5 MB = 268 MB > L2
Hs = [torch.tanh(torch.randn(B, H, device=dev)) for _ in range(NBUF)]
Zs = [torch.zeros(B, H, device=dev) for _ in range(NBUF)]
G = torch.randn(B, NH, device=dev)
Wh = torch.randn(H, NH, device=dev) * 0.1
heads = torch.zeros(B, NH, device=dev)
segs = [(0, 5), (5, 1)]
x = torch.randn(D * B, device=dev)
W0 = torch.randn(D, H, device=dev) * 0.3
b0 = torch.zeros(H, device=dev)
gs = torch.zeros(D * B, device=dev)
n_tail = H + H * NH + NH
tail = torch.zeros(n_tail, device=dev)
gw = torch.zeros((D + 1) * H, device=dev)
ws = torch.zeros(engine.drp_head_backward_ctas(B) * max(n_tail, (D + 1) * H), device=dev)


def timeit(name, fn, bytes_moved, reps=40):
    """The calls are captured into one CUDA graph so the host's launch rate is not what is timed."""
    for mode in ("cold", "warm"):
        pick = (lambda i: i % NBUF) if mode == "cold" else (lambda i: 0)
        for i in range(5):
            fn(pick(i))
        torch.cuda.synchronize()
        st = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(st):
            with torch.cuda.graph(graph, stream=st):
                for i in range(reps):
                    fn(pick(i))
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print(f"{name:28s} {mode}: {us:7.1f} us  {bytes_moved / us / 1e3:7.1f} GB/s", flush=True)


MB = B * H * 4
timeit("rowdot heads", lambda i: engine.rowdot(False, B, NH, H, Hs[i], H, Wh, NH, heads, NH), MB)
timeit("layer0 forward", lambda i: engine.drp_layer0(B, D, H, segs, x, W0, b0, Zs[i], None), MB)
timeit("head backward", lambda i: engine.drp_head_backward(B, H, NH, Hs[i], G, Wh, Zs[i], tail, ws), 2 * MB)
W0T = W0.t().contiguous()
timeit("layer0 backward", lambda i: engine.drp_layer0_backward(B, D, H, segs, x, Hs[i], W0T, gs, gw, ws), 2 * MB)
timeit("torch copy", lambda i: Zs[i].copy_(Hs[i]), 2 * MB)
timeit("torch sum", lambda i: Hs[i].sum(), MB)
