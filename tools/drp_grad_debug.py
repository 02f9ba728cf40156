"""Per-segment comparison of the first DRP gradient (PowerGen bench workload, reduced batch) with the
CPU oracle's: max |diff| relative to the segment's and to the whole gradient's largest entry."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from tests.test_trajectory_gpu import _planner  # noqa: E402
from oracle.iteration import OracleIteration  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 160
model, wl, pl = _planner("powergen_drp", B, False)
T = wl["T"]
fw, te = pl.train_fwd.program, pl.test_fwd.program
relax = pl.compiled.relax_config()
flat0 = pl.params[0].cpu().clone()
init = {k: {kk: vv.clone() for kk, vv in v.items()} for k, v in pl.plan.params_to_pytree(flat0)["params"].items()}
it = OracleIteration(model, relax, (fw.noise_layout, te.noise_layout), B, T, wl["lr"], plan="drp",
                     plan_sigmoid_weight=float(wl["plan_kwargs"].get("sigmoid_weight", 1.0)),
                     topology=wl["plan_kwargs"].get("topology", ()), ent_coeff=float(pl.start_entropy_coeff),
                     policy_layout=pl.plan.layout, init=init)
# experiment switches: RK_DEBUG_PATCH=heads -> the head outputs of the train sweep are recomputed in fp32
# from the stored H1; =h1 -> H1 and the heads are recomputed in fp32 from the stored H0
patch = os.environ.get("RK_DEBUG_PATCH", "")
if patch:
    from pyrddlgym_jax_b200.core.drp import DrpPipeline
    orig = DrpPipeline.mlp_forward

    def patched(self, params, Bq, x_fm, H, heads, xp=None, film=None, keep_hidden=True):
        orig(self, params, Bq, x_fm, H, heads, xp, film, keep_hidden)
        if keep_hidden:
            h0, h1 = self.hid
            if patch == "h1":
                H[1].copy_(torch.tanh(torch.addmm(self._W(params, "b1"), H[0], self._W(params, "W1").view(h0, h1))))
            heads.copy_(torch.addmm(self._W(params, "bh"), H[1], self._W(params, "Wh").view(h1, self.NH)))
    DrpPipeline.mlp_forward = patched
want_train, want_test = it()
z = it.last_noise
if z.get("train") is not None:
    pl.train_noise.copy_(z["train"].reshape(-1))
if z.get("test") is not None:
    pl.test_noise.copy_(z["test"].reshape(-1))
pl.drp.pol_noise.copy_(z["policy"])
pl.iterate()
train, test, _ = pl.read_stats()
g_gpu = pl.grad_used[0].cpu().double()
g_ref = pl.plan.pytree_to_params(it.last_grad).double()
gmax = float(g_ref.abs().max())
print(f"train {float(train[0])} vs {want_train}; test {float(test[0])} vs {want_test}; max|g| {gmax:.4e}")
for name, off, r, c in pl.plan.segments:
    a, b = g_gpu[off:off + r * c], g_ref[off:off + r * c]
    d = (a - b).abs()
    i = int(d.argmax())
    print(f"  {name:6s} [{r}x{c}] max|g| {float(b.abs().max()):.3e}  max|diff| {float(d.max()):.3e} "
          f"({float(d.max()) / max(float(b.abs().max()), 1e-30):.2e} of segment max, {float(d.max()) / gmax:.2e} of global max)"
          f"  at {i}: gpu {float(a[i]):.6e} ref {float(b[i]):.6e}")

# ---- stored activations of a few epochs against a torch evaluation of the same network
drp, plan = pl.drp, pl.plan
S = drp.S
params = flat0.double()
def W(name):
    off, r, c = plan.seg[name]
    return params[off:off + r * c].view(r, c)
for t in (0, 1, 50, 99):
    xt = drp.tape[t * S * B:(t + 1) * S * B].cpu().double()
    x = torch.cat([xt[off * B:(off + n) * B].view(B, n) for off, n in plan.state_segs], dim=1)
    H0r = torch.tanh(x @ W("W0") + W("b0"))
    H1r = torch.tanh(H0r @ W("W1") + W("b1"))
    hr = H1r @ W("Wh") + W("bh")
    d0 = (drp.H[0][t].cpu().double() - H0r).abs()
    d1 = (drp.H[1][t].cpu().double() - H1r).abs()
    dh = (drp.heads[t].cpu().double() - hr).abs()
    print(f"t={t}: max|dH0| {float(d0.max()):.2e} (row {int(d0.max(1).values.argmax())}), max|dH1| {float(d1.max()):.2e} "
          f"(row {int(d1.max(1).values.argmax())}), max|dheads| {float(dh.max()):.2e} (row {int(dh.max(1).values.argmax())}); "
          f"rows with |dH0| > 1e-5: {int((d0.max(1).values > 1e-5).sum())}, |dH1| > 1e-4: {int((d1.max(1).values > 1e-4).sum())}")

# ---- the head product alone: stored heads against an fp64 product of the STORED H1
for t in (0, 50):
    H1s = drp.H[1][t].cpu().double()
    hr = H1s @ W("Wh") + W("bh")
    got = drp.heads[t].cpu().double()
    d = got - hr
    print(f"t={t}: head product error: max {float(d.abs().max()):.3e}, mean signed per column "
          + " ".join(f"{float(d[:, n].mean()):+.2e}" for n in range(d.shape[1])))
    print("   row 0 heads:", " ".join(f"{float(v):+.6f}" for v in got[0]))
    print("   row 0 ref  :", " ".join(f"{float(v):+.6f}" for v in hr[0]))
print("sigma range (log):", drp.ls_lo, drp.ls_hi, " bh:", W("bh").flatten().tolist())
