"""Accuracy probe of the one-shot weight-gradient product (rk_drp_wgrad_hidden) against fp64 at the
bench's contraction length K = T * B, and of the fused forward's head outputs: maximum error relative
to the largest entry and the mean SIGNED error along the result's sign (a truncating accumulator shows
up as a negative bias)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pyrddlgym_jax_b200.core import engine  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
M = N = 256
for K in (32768 * 4, 32768 * 100):
    A = torch.tanh(torch.randn(K, M, generator=g, device=dev))
    Bm = torch.randn(K, N, generator=g, device=dev) * 3e-5
    dW = torch.zeros(M, N, device=dev)
    db = torch.zeros(N, device=dev)
    ws = torch.zeros(engine.drp_wgrad_ctas() * (M * N + N), device=dev)
    engine.drp_wgrad_hidden(False, M, N, K, A, Bm, dW, db, ws)
    torch.cuda.synchronize()
    want = torch.zeros(M, N, dtype=torch.float64, device=dev)
    step = 1 << 18
    for k0 in range(0, K, step):
        want += A[k0:k0 + step].double().t() @ Bm[k0:k0 + step].double()
    err = dW.double() - want
    scale = float(want.abs().max())
    print(f"K={K}: max|err|/max|want| = {float(err.abs().max()) / scale:.3e}; rms err / rms want = "
          f"{float(err.pow(2).mean().sqrt() / want.pow(2).mean().sqrt()):.3e}; mean signed err along sign(want) "
          f"/ mean|want| = {float((err * want.sign()).mean() / want.abs().mean()):.3e}")
    wb = Bm.double().sum(0)
    print(f"        bias sums: max|err|/max|want| = {float((db.double() - wb).abs().max() / wb.abs().max()):.3e}")
    del A, Bm
