"""Phase timing of the fused policy kernels (csrc/rk_mlp2.cu) at the PowerGen bench shape: every CTA
stamps clock64 at its phase boundaries; prints the median / max duration of each phase over the CTAs.

    python tools/mlp2_phases.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pyrddlgym_jax_b200.core import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D, h0, h1, NH = 6, 256, 256, 10
segs = [(0, 5), (5, 1)]
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g).to(dev)     # noqa: E731
x, W0, b0, W1, b1, Wh, bh = r(D * B), r(D, h0) * 0.4, r(h0) * 0.1, r(h0, h1) * 0.06, r(h1) * 0.1, r(h1, NH) * 0.06, r(NH)
W1T = W1.t().contiguous()
hiT, loT, hi, lo = (torch.empty_like(W1) for _ in range(4))
engine.tf32_split(W1T.reshape(-1), hiT.reshape(-1), loT.reshape(-1), W1.numel())
engine.tf32_split(W1.reshape(-1), hi.reshape(-1), lo.reshape(-1), W1.numel())
img_f = torch.zeros(engine.drp_mlp2_image_floats(D, h0, h1, NH, False), device=dev)
img_b = torch.zeros(engine.drp_mlp2_image_floats(D, h0, h1, NH, True), device=dev)
engine.drp_mlp2_prepare(D, h0, h1, NH, W0.reshape(-1), Wh.reshape(-1), img_f, img_b)
H0, H1, heads = torch.zeros(B, h0, device=dev), torch.zeros(B, h1, device=dev), torch.zeros(B, NH, device=dev)
G, gz0, gz1, gs = r(B, NH), torch.zeros(B, h0, device=dev), torch.zeros(B, h1, device=dev), torch.zeros(D * B, device=dev)
stamps = torch.zeros(256 * 8, dtype=torch.int64, device=dev)
names = ["prologue", "produce + hidden MMAs (to last tile written)", "wait hidden MMAs", "epilogue chunks",
         "wait second product", "tail"]


def fwd(keep):
    engine.drp_mlp2_forward("tanh", B, D, h0, h1, NH, segs, x, img_f, b0, hiT.reshape(-1), loT.reshape(-1),
                            b1, bh, H0 if keep else None, H1 if keep else None, heads)


def bwd():
    engine.drp_mlp2_backward("tanh", B, D, h0, h1, NH, segs, G, H0, H1, img_b, hi.reshape(-1), lo.reshape(-1),
                             gz0, gz1, gs)


flush = torch.empty(64 * 1024 * 1024, device=dev)
for label, fn in (("forward (test sweep: nothing kept)", lambda: fwd(False)), ("forward (train sweep)", lambda: fwd(True)),
                  ("backward", bwd)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    engine.drp_mlp2_set_timing(stamps)
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    engine.drp_mlp2_set_timing(None)
    ncta = -(-B // 222) if B >= 148 * 128 else -(-B // 128)
    s = stamps.view(256, 8)[:ncta, :7].cpu().double()
    d = s[:, 1:] - s[:, :-1]
    print(f"{label}: {e0.elapsed_time(e1) * 1000:.1f} us; per-CTA phases in cycles (median / max over {ncta} CTAs), "
          f"total {float((s[:, 6] - s[:, 0]).median()):.0f}")
    for i, nm in enumerate(names):
        print(f"    {nm:48s} {float(d[:, i].median()):9.0f} {float(d[:, i].max()):9.0f}")
