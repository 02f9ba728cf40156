"""GPU parity of rk_optimizer_step (through the C ABI) against the oracle's restatement of the optax
chain (oracle/planner.py: clip_by_global_norm, rmsprop_step, adam_step, sgd_step, sgd_momentum_step;
reference planner.py:2349-2387, 2885-2899) and of the box projection (planner.py:896-904): five
consecutive steps from the same gradients, parameters and optimiser state compared element-wise
after every step."""
import numpy as np
import pytest
import torch

from oracle import planner as OP

from .parity import assert_close

pytestmark = pytest.mark.gpu

CASES = [  # kind, kwargs, clip, box
    ("rmsprop", {}, None, False),
    ("rmsprop", {"decay": 0.95, "eps": 1e-6}, 0.5, True),
    ("adam", {}, None, False),
    ("adam", {"b1": 0.8, "b2": 0.99, "eps": 1e-6}, 2.0, True),
    ("sgd", {}, None, True),
    ("sgd", {"momentum": 0.9}, 1.0, False),
]


def _hyper(kind, lr, kw, clip, step):
    if kind == "rmsprop":
        h = [lr, kw.get("decay", 0.9), kw.get("eps", 1e-8), 0., 0., 0., step]
    elif kind == "adam":
        h = [lr, kw.get("b1", 0.9), kw.get("eps", 1e-8), kw.get("b2", 0.999), 0., 0., step]
    else:
        h = [lr, 0., 0., 0., 0., kw.get("momentum", 0.) or 0., step]
    h[4] = float(clip) if clip is not None else 0.
    return h + [1.0 - h[1], 1.0 - h[3]]


@pytest.mark.parametrize("n", [37, 1600, 70154])          # single-CTA path, SLP size, DRP size
@pytest.mark.parametrize("kind,kw,clip,box", CASES)
def test_optimizer_step_matches_oracle(cuda_device, n, kind, kw, clip, box):
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(1234 + n)
    lr = 0.03
    p = torch.randn(n, generator=g)
    lo = (p - torch.rand(n, generator=g) * 0.05) if box else None
    hi = (p + torch.rand(n, generator=g) * 0.05) if box else None
    if box:                                              # one-sided and open boxes too
        lo[::7] = -float("inf")
        hi[::5] = float("inf")
    n_state = 2 if kind == "adam" else 1
    state_ref = [torch.zeros(n) for _ in range(n_state)]
    p_ref = p.clone()
    p_dev = p.to(cuda_device)
    st_dev = torch.zeros(n_state * n, device=cuda_device)
    flag = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    lo_d = None if lo is None else lo.to(cuda_device)
    hi_d = None if hi is None else hi.to(cuda_device)
    for step in range(1, 6):
        grad = torch.randn(n, generator=g) * (10.0 ** float(torch.randint(-3, 2, (1,), generator=g)))
        if step == 4:
            grad = torch.zeros(n)                        # zero-gradient flag (planner.py:2861-2864)
        gr = grad
        if clip is not None:
            (gr,) = OP.clip_by_global_norm([grad], clip)
        if kind == "rmsprop":
            p_ref, state_ref[0] = OP.rmsprop_step(p_ref, gr, state_ref[0], lr, kw.get("decay", 0.9),
                                                  kw.get("eps", 1e-8))
        elif kind == "adam":
            p_ref, state_ref[0], state_ref[1] = OP.adam_step(
                p_ref, gr, state_ref[0], state_ref[1], step, lr, kw.get("b1", 0.9),
                kw.get("b2", 0.999), kw.get("eps", 1e-8))
        elif kw.get("momentum"):
            p_ref, state_ref[0] = OP.sgd_momentum_step(p_ref, gr, state_ref[0], lr, kw["momentum"])
        else:
            p_ref = OP.sgd_step(p_ref, gr, lr)
        if box:
            p_ref = torch.minimum(torch.maximum(p_ref, lo), hi)
        hyper = torch.tensor(_hyper(kind, lr, kw, clip, float(step)), device=cuda_device)
        engine.optimizer_step(kind, hyper, n, p_dev, st_dev, grad.to(cuda_device), lo_d, hi_d, flag)
        torch.cuda.synchronize()
        assert_close(p_dev.cpu().numpy(), p_ref.numpy(), 2e-6, f"{kind} params after step {step}")
        if kind != "sgd" or kw.get("momentum"):
            for k in range(n_state):
                assert_close(st_dev[k * n:(k + 1) * n].cpu().numpy(), state_ref[k].numpy(), 2e-6,
                             f"{kind} state {k} after step {step}")
        assert int(flag.item()) == (1 if step == 4 else 0)


def test_reduce_partials_and_mean_loss_match_numpy(cuda_device):
    """rk_reduce_partials (fixed-order sum of the per-warp gradient rows) and rk_mean_loss."""
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(7)
    rows, n = 513, 1600
    gp = torch.randn(rows, n, generator=g)
    out = torch.zeros(n, device=cuda_device)
    engine.reduce_partials(gp.to(cuda_device).reshape(-1), rows, n, out)
    ret = torch.randn(4097, generator=g) * 50
    loss = torch.zeros(1, device=cuda_device)
    gret = torch.zeros(4097, device=cuda_device)
    engine.mean_loss(ret.to(cuda_device), 4097, loss, gret)
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy(), gp.double().sum(0).numpy(), 1e-5, "reduce_partials", floor_frac=1.0)
    assert abs(float(loss) + float(ret.double().mean())) <= 1e-5 * float(ret.abs().mean())
    np.testing.assert_array_equal(gret.cpu().numpy(), np.full(4097, np.float32(-1.0 / 4097)))


def test_philox_noise_laws_and_counter(cuda_device):
    """rk_draw_noise: every law has the moments / quantiles of its distribution, sites land in their own
    fluent-major blocks, the same (seed, draw number) reproduces the tensor and rk_noise_advance changes it."""
    import math
    import torch
    from pyrddlgym_jax_b200.core import engine
    T, B = 8, 20000
    sites = [("uniform", 0, 1), ("normal", 1, 2), ("exponential", 3, 1), ("gumbel", 4, 1),
             ("logistic", 5, 3), ("laplace", 8, 1), ("cauchy", 9, 1)]
    N = 10
    out = torch.zeros(T, N * B, device=cuda_device)
    draws = torch.zeros(1, dtype=torch.int64, device=cuda_device)
    engine.draw_noise(T, B, sites, 1234, draws, out)
    again = torch.zeros_like(out)
    engine.draw_noise(T, B, sites, 1234, draws, again)
    assert torch.equal(out, again)
    engine.noise_advance(draws)
    engine.draw_noise(T, B, sites, 1234, draws, again)
    assert int(draws.item()) == 1 and not torch.equal(out, again)
    other = torch.zeros_like(out)
    engine.draw_noise(T, B, sites, 1235, None, other)
    assert not torch.equal(out, other)
    blk = lambda off, n: out[:, off * B:(off + n) * B].double().reshape(-1).cpu()     # noqa: E731
    n = T * B
    se = 6.0 / math.sqrt(n)                                     # six standard errors of a unit-variance mean
    u = blk(0, 1)
    assert 0.0 <= float(u.min()) and float(u.max()) < 1.0
    assert abs(float(u.mean()) - 0.5) < se and abs(float(u.var()) - 1 / 12) < se
    z = blk(1, 2)
    assert abs(float(z.mean())) < se and abs(float(z.var()) - 1.0) < 3 * se
    assert abs(float((z ** 4).mean()) - 3.0) < 0.2            # kurtosis of a normal
    zz = out[:, B:3 * B].double().reshape(T, B, 2).cpu()     # the two words of a rollout are independent
    assert abs(float((zz[..., 0] * zz[..., 1]).mean())) < 2 * se
    e = blk(3, 1)
    assert float(e.min()) >= 0.0 and abs(float(e.mean()) - 1.0) < 2 * se
    g = blk(4, 1)
    assert abs(float(g.mean()) - 0.5772156649) < 2 * se and abs(float(g.var()) - math.pi ** 2 / 6) < 0.1
    lg = blk(5, 3)
    assert abs(float(lg.mean())) < 3 * se and abs(float(lg.var()) - math.pi ** 2 / 3) < 0.15
    lp = blk(8, 1)
    assert abs(float(lp.mean())) < 2 * se and abs(float(lp.var()) - 2.0) < 0.15
    c = blk(9, 1)
    assert abs(float(c.median())) < 0.05 and abs(float(c.quantile(0.75)) - 1.0) < 0.05
    assert bool(torch.isfinite(out).all())
    # Gamma sites (kind 7): per-word concentrations, Marsaglia-Tsang rejection on per-element streams
    from scipy import stats
    alphas = [0.3, 1.0, 2.5, 40.0]
    sites = [("normal", 0, 1), ("gamma", 1, 4)]
    conc = torch.tensor([1.0] + alphas, device=cuda_device)
    gout = torch.zeros(T, 5 * B, device=cuda_device)
    engine.draw_noise(T, B, sites, 99, draws, gout, conc)
    again = torch.zeros_like(gout)
    engine.draw_noise(T, B, sites, 99, draws, again, conc)
    assert torch.equal(gout, again) and bool(torch.isfinite(gout).all())
    gam = gout[:, B:5 * B].double().reshape(T, B, 4).cpu()
    for k, a in enumerate(alphas):
        x = gam[..., k].reshape(-1)
        assert float(x.min()) > 0.0
        assert abs(float(x.mean()) - a) < 6.0 * math.sqrt(a / n), (a, float(x.mean()))
        assert abs(float(x.var()) / a - 1.0) < 0.06, (a, float(x.var()))
        ks = stats.kstest(x.numpy()[::7], "gamma", args=(a,))
        assert ks.statistic < 0.012, (a, ks)
    assert abs(float((gam[..., 1] * gam[..., 2]).mean()) - 1.0 * 2.5) < 0.05     # independent words


def test_drp_returns_loss_and_transpose_split(cuda_device):
    """rk_drp_returns_loss (discounted returns, -mean - coeff * mean entropy) and rk_transpose_split
    (W^T and its TF32 hi / lo parts) against torch."""
    import torch
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(5)
    T, B = 37, 1000
    reward = torch.randn(T * B, generator=g).to(cuda_device)
    entropy = torch.randn(T, B, generator=g).to(cuda_device)
    disc = (0.97 ** torch.arange(T, dtype=torch.float32)).to(cuda_device)
    coeff = torch.tensor([0.3], device=cuda_device)
    returns, ent, loss = (torch.zeros(B, device=cuda_device), torch.zeros(B, device=cuda_device),
                          torch.zeros(1, device=cuda_device))
    engine.drp_returns_loss(T, B, reward, disc, entropy, coeff, returns, ent, loss)
    want_r = (reward.view(T, B).double() * disc.double().view(T, 1)).sum(0)
    want_l = -want_r.mean() - 0.3 * entropy.double().sum() / B
    torch.testing.assert_close(returns.double(), want_r, rtol=1e-5, atol=1e-5)
    assert abs(float(loss[0]) - float(want_l)) <= 1e-5 * max(1.0, abs(float(want_l)))
    engine.drp_returns_loss(T, B, reward, None, None, None, returns, None, loss)
    want_r = reward.view(T, B).double().sum(0)
    torch.testing.assert_close(returns.double(), want_r, rtol=1e-5, atol=1e-5)
    assert abs(float(loss[0]) + float(want_r.mean())) <= 1e-5
    for r, c in ((6, 256), (256, 10), (70, 33), (256, 256)):
        W = torch.randn(r, c, generator=g).to(cuda_device)
        WT, hi, lo = (torch.zeros(c, r, device=cuda_device) for _ in range(3))
        engine.transpose_split(r, c, W.reshape(-1), WT.reshape(-1), hi.reshape(-1), lo.reshape(-1))
        assert torch.equal(WT, W.t().contiguous())
        assert torch.equal(hi + lo, WT)
        assert bool(((hi.view(torch.int32) & 0x1FFF) == 0).all())        # TF32-exact high parts
        WT2 = torch.zeros(c, r, device=cuda_device)
        engine.transpose_split(r, c, W.reshape(-1), WT2.reshape(-1), None, None)
        assert torch.equal(WT2, WT)


def test_native_utilities_match_autograd(cuda_device):
    """rk_utility_loss (closed-form cotangents, one CTA) against torch autograd of the planner's own
    utility definitions (core/planner.py, restating planner.py:2161-2247)."""
    import torch
    from pyrddlgym_jax_b200.core import engine, planner
    g = torch.Generator().manual_seed(3)
    for B in (1000, 40000):
        r = (torch.randn(B, generator=g) * 3.0 + 1.5).to(cuda_device)
        cases = [("mean_var", planner.mean_variance_utility, dict(beta=0.7), 0.7, 0.0),
                 ("mean_std", planner.mean_deviation_utility, dict(beta=0.7), 0.7, 0.0),
                 ("mean_semivar", planner.mean_semivariance_utility, dict(beta=1.3), 1.3, 0.0),
                 ("mean_semidev", planner.mean_semideviation_utility, dict(beta=1.3), 1.3, 0.0),
                 ("sharpe", planner.sharpe_utility, dict(risk_free=0.2, eps=1e-6), 0.2, 1e-6),
                 ("entropic", planner.entropic_utility, dict(beta=0.4), 0.4, 0.0)]
        for kind, fn, kw, p0, p1 in cases:
            x = r.double().clone().requires_grad_(True)
            want = -fn(x, **kw)
            (gw,) = torch.autograd.grad(want, x)
            loss, gret = torch.zeros(1, device=cuda_device), torch.zeros(B, device=cuda_device)
            engine.utility_loss(kind, B, r, p0, p1, 1.0, loss, gret)
            assert abs(float(loss[0]) - float(want)) <= 2e-5 * max(1.0, abs(float(want))), kind
            torch.testing.assert_close(gret.double(), gw, rtol=2e-4, atol=2e-5 * float(gw.abs().max()))
        x = r.double().clone().requires_grad_(True)
        want = -(torch.sign(x) * torch.log1p(x.abs())).mean()
        (gw,) = torch.autograd.grad(want, x)
        loss, gret = torch.zeros(1, device=cuda_device), torch.zeros(B, device=cuda_device)
        engine.utility_loss("symlog_mean", B, r, 0.0, 0.0, 0.5, loss, gret)
        assert abs(float(loss[0]) - float(want)) <= 1e-5
        torch.testing.assert_close(gret.double(), 0.5 * gw, rtol=1e-5, atol=1e-9)


def test_sorted_utilities_match_autograd(cuda_device):
    """rk_utility_sorted_loss (var, cvar, cvar_ste, expectile, gini: one device sort, closed-form
    cotangents) against torch autograd of the planner's utility definitions in fp64 (restating
    planner.py:2195-2229), with ties in the return vector, and the sharded form: each of two ranks'
    slices of the cotangent of the global vector."""
    import torch
    from pyrddlgym_jax_b200.core import engine, planner
    g = torch.Generator().manual_seed(11)
    for B in (7, 1000, 40000):
        r = torch.randn(B, generator=g) * 3.0 + 1.5
        r[: B // 3] = torch.round(r[: B // 3])                       # tie groups
        r = r.to(cuda_device)
        work = engine.utility_sorted_work(B, cuda_device)
        cases = [("var", planner.var_utility, dict(alpha=0.3), 0.3, 0),
                 ("var", planner.var_utility, dict(alpha=1.0), 1.0, 0),
                 ("cvar", planner.cvar_utility, dict(alpha=0.25), 0.25, 0),
                 ("cvar_ste", planner.cvar_ste_utility, dict(alpha=0.1), 0.1, 0),
                 ("expectile", planner.expectile_utility, dict(alpha=0.2, max_iter=30), 0.2, 30),
                 ("expectile", planner.expectile_utility, dict(alpha=0.7, max_iter=0), 0.7, 0),
                 ("gini", planner.gini_utility, dict(gamma=0.6), 0.6, 0)]
        for kind, fn, kw, p0, it in cases:
            if kind == "gini" and B > 5000:
                continue                                             # the autograd form is O(B^2)
            # the C ABI takes the parameter as a float: the fp64 reference sees the same rounded value
            p0 = float(np.float32(p0))
            kw = {k: (p0 if k in ("alpha", "gamma") else v) for k, v in kw.items()}
            x = r.double().clone().requires_grad_(True)
            want = -fn(x, **kw)
            (gw,) = torch.autograd.grad(want, x)
            loss, gret = torch.zeros(1, device=cuda_device), torch.zeros(B, device=cuda_device)
            engine.utility_sorted_loss(kind, B, r, p0, it, 1.0, loss, gret, work)
            assert abs(float(loss[0]) - float(want)) <= 2e-5 * max(1.0, abs(float(want))), (kind, B)
            torch.testing.assert_close(gret.double(), gw, rtol=2e-4, atol=2e-5 * float(gw.abs().max()),
                                       msg=lambda m: f"{kind} B={B}: {m}")
            # two shards of the same global vector
            if B >= 1000:
                h = B // 2
                for b0, n in ((0, h), (h, B - h)):
                    part = torch.zeros(n, device=cuda_device)
                    engine.utility_sorted_loss(kind, B, r, p0, it, 0.5, None, part, work, b0, n)
                    torch.testing.assert_close(part, 0.5 * gret[b0:b0 + n], rtol=0, atol=0)
    # the closed-form family, sharded
    B = 1000
    r = (torch.randn(B, generator=g) * 2.0).to(cuda_device)
    loss, full = torch.zeros(1, device=cuda_device), torch.zeros(B, device=cuda_device)
    engine.utility_loss("mean_var", B, r, 0.7, 0.0, 1.0, loss, full)
    part = torch.zeros(300, device=cuda_device)
    engine.utility_loss("mean_var", B, r, 0.7, 0.0, 1.0, None, part, 500, 300)
    torch.testing.assert_close(part, full[500:800], rtol=0, atol=0)


@pytest.mark.parametrize("utility,kw", [("cvar", {"alpha": 0.3}), ("gini", {"gamma": 0.2}),
                                        ("expectile", {"alpha": 0.3})])
def test_sorted_utility_planner_matches_eager(cuda_device, utility, kw):
    """A planner with an order-statistic utility runs as one CUDA graph and follows the autograd path."""
    import torch
    import bench
    from pyrddlgym_jax_b200.core import logic
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner, JaxStraightLinePlan
    model = bench.workload_model("quadcopter")
    plans = []
    for native in (True, False):
        pl = JaxBackpropPlanner(model, JaxStraightLinePlan(), batch_size_train=64, batch_size_test=64,
                                rollout_horizon=10, optimizer="rmsprop",
                                optimizer_kwargs={"learning_rate": 0.05}, pgpe=None,
                                compiler=logic.DefaultJaxRDDLCompilerWithGrad,
                                compiler_kwargs={"sigmoid_weight": 10.}, utility=utility,
                                utility_kwargs=kw)
        assert pl._native_utility is not None
        if not native:
            pl._native_utility, pl._utility_eager = None, True
        pl.initialize(7)
        for _ in range(3):
            pl.iterate()
        assert (pl._graph is not None) == native
        plans.append((pl.params[0].cpu().clone(), pl.read_stats()[0].copy()))
    torch.testing.assert_close(plans[0][0], plans[1][0], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(plans[0][1], plans[1][1], rtol=1e-4)


def test_native_utility_keeps_the_graph_and_matches_eager(cuda_device):
    """A planner with utility='mean_var' runs its iteration as one CUDA graph (native cotangent) and
    produces the same plan as the autograd path (RK-independent check: same inputs, two planners)."""
    import torch
    import bench
    from pyrddlgym_jax_b200.core import logic
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner, JaxStraightLinePlan
    model = bench.workload_model("quadcopter")            # no draw sites: both planners see the same inputs
    plans = []
    for native in (True, False):
        pl = JaxBackpropPlanner(model, JaxStraightLinePlan(), batch_size_train=64, batch_size_test=64,
                                rollout_horizon=10, optimizer="rmsprop",
                                optimizer_kwargs={"learning_rate": 0.05}, pgpe=None,
                                compiler=logic.DefaultJaxRDDLCompilerWithGrad,
                                compiler_kwargs={"sigmoid_weight": 10.}, utility="mean_var",
                                utility_kwargs={"beta": 0.01})
        assert pl._native_utility is not None
        if not native:
            pl._native_utility, pl._utility_eager = None, True
        pl.initialize(7)
        assert pl.train_noise is None
        for _ in range(3):
            pl.iterate()
        assert (pl._graph is not None) == native
        plans.append((pl.params[0].cpu().clone(), pl.read_stats()[0].copy()))
    torch.testing.assert_close(plans[0][0], plans[1][0], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("rows,n", [(7, 33), (512, 1600), (1024, 40), (1025, 40), (16384, 1000), (8192, 720)])
@pytest.mark.parametrize("kind", ["sgd", "rmsprop", "adam"])
def test_fused_reduce_optimizer_equals_separate_launches(cuda_device, rows, n, kind):
    """rk_reduce_optimizer_step (partial-row reduction + optimiser step + box + all-zero flag in one launch)
    against rk_reduce_partials followed by rk_optimizer_step: gradient, parameters, optimiser state and
    flag agree to the bit, over several steps, and the counters are left at zero."""
    import torch
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(rows * 31 + n)
    dev = cuda_device
    hyper = torch.tensor([0.03, 0.9, 1e-8, 0.999, 0.0, 0.0 if kind != "sgd" else 0.5, 1.0,
                          1.0 - 0.9, 1.0 - 0.999], dtype=torch.float32, device=dev)
    n_state = {"sgd": 1, "rmsprop": 1, "adam": 2}[kind]
    p0 = torch.randn(n, generator=g).to(dev)
    lo, hi = (p0 - 0.05).contiguous(), (p0 + 0.04).contiguous()
    pa, pb = p0.clone(), p0.clone()
    sa, sb = torch.zeros(n_state * n, device=dev), torch.zeros(n_state * n, device=dev)
    ga, gb = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    fa = torch.full((1,), -1, dtype=torch.int32, device=dev)
    fb = torch.full((1,), -1, dtype=torch.int32, device=dev)
    counters = torch.zeros(2, dtype=torch.int32, device=dev)
    for step in range(4):
        gp = (torch.randn(rows, n, generator=g) * (0.0 if step == 2 else 1e-2)).to(dev)
        gp2 = gp.clone()                  # the large path reduces in place
        engine.reduce_partials(gp.reshape(-1), rows, n, ga)
        engine.optimizer_step(kind, hyper, n, pa, sa, ga, lo, hi, fa)
        engine.reduce_optimizer_step(gp2.reshape(-1), rows, n, gb, kind, hyper, pb, sb, lo, hi, fb, counters)
        torch.cuda.synchronize()
        assert torch.equal(ga, gb) and torch.equal(pa, pb) and torch.equal(sa, sb), step
        assert int(fa) == int(fb) == (1 if step == 2 else 0)
        assert counters.tolist() == [0, 0]
        hyper[6] += 1.0


@pytest.mark.parametrize("n", [700, 40000])
@pytest.mark.parametrize("kind", ["sgd", "rmsprop", "adam"])
def test_chain_step_recurrences_both_paths(cuda_device, n, kind):
    """rk_optimizer_chain_step on a short vector (single launch) and a long one (statistics launch + apply
    launch): clip_by_global_norm -> optimiser -> reduce_on_plateau scale -> ema with debias -> box, against
    the same recurrences written out in torch (fp64 where it matters), over several steps; with add_noise
    on, the perturbation has the requested deviation."""
    import torch
    from oracle import planner as OP
    from pyrddlgym_jax_b200.core import engine
    from pyrddlgym_jax_b200.core.planner import plateau_init, plateau_kwargs
    dev = cuda_device
    g = torch.Generator().manual_seed(n + len(kind))
    lr, d1, eps, b2, clip, decay = 0.05, 0.9, 1e-8, 0.999, 0.5, 0.8
    hyper = torch.tensor([lr, d1, eps, b2, clip, 0.0, 1.0, 1.0 - d1, 1.0 - b2], dtype=torch.float32, device=dev)
    pk = plateau_kwargs({"factor": 0.5, "patience": 1, "rtol": 1e-3})
    cfg = {"noise": None, "ema_decay": decay, "plateau": pk}
    p0 = torch.randn(n, generator=g)
    lo, hi = (p0 - 0.2), (p0 + 0.2)
    params, state = p0.clone().to(dev), torch.zeros((2 if kind == "adam" else 1) * n, device=dev)
    ema, count, scratch = torch.zeros(n, device=dev), torch.zeros(1, device=dev), torch.zeros(8, device=dev)
    plat = plateau_init(1, dev)[0]
    loss = torch.zeros(1, device=dev)
    zero = torch.zeros(1, dtype=torch.int32, device=dev)
    want, st_mu, st_nu, ema_w = p0.double().clone(), torch.zeros(n).double(), torch.zeros(n).double(), \
        torch.zeros(n).double()
    ref = OP.reduce_on_plateau_init()
    for t in range(1, 5):
        grad = torch.randn(n, generator=g) * 0.1
        lval = 5.0 - 0.5 * t if t < 3 else 4.0            # improves twice, then stalls
        loss.fill_(lval)
        hyper[6] = float(t)
        engine.optimizer_chain_step(kind, hyper, n, params, state, grad.to(dev), lo.to(dev), hi.to(dev), zero,
                                    cfg, count, plat, ema, loss, 0, scratch)
        torch.cuda.synchronize()
        gd = grad.double()
        gn = float(gd.norm())
        gd = gd * min(1.0, clip / gn)
        if kind == "rmsprop":
            st_nu = d1 * st_nu + (1 - d1) * gd * gd
            upd = gd / torch.sqrt(st_nu + eps)
        elif kind == "adam":
            st_mu = d1 * st_mu + (1 - d1) * gd
            st_nu = b2 * st_nu + (1 - b2) * gd * gd
            upd = (st_mu / (1 - d1 ** t)) / (torch.sqrt(st_nu / (1 - b2 ** t)) + eps)
        else:
            upd = gd
        ref = OP.reduce_on_plateau_step(ref, lval, **pk)
        delta = -lr * upd * ref["scale"]
        ema_w = decay * ema_w + (1 - decay) * delta
        want = torch.minimum(torch.maximum(want + ema_w / (1 - decay ** t), lo.double()), hi.double())
        torch.testing.assert_close(params.cpu().double(), want, rtol=2e-5, atol=2e-6)
        assert abs(float(plat[0]) - ref["scale"]) < 1e-7 and float(count) == t and int(zero) == 0
    assert ref["scale"] < 1.0
    # add_noise: N(0, eta / t^gamma) on every component, nothing else moving (sgd, lr 1, zero gradient)
    hyper2 = torch.tensor([1.0, 0, 0, 0, 0, 0, 1, 1, 1], dtype=torch.float32, device=dev)
    p2, cnt2 = torch.zeros(n, device=dev), torch.zeros(1, device=dev)
    engine.optimizer_chain_step("sgd", hyper2, n, p2, None, torch.zeros(n, device=dev), None, None, zero,
                                {"noise": (0.04, 0.55), "ema_decay": None, "plateau": None}, cnt2, None, None,
                                loss, 17, scratch)
    torch.cuda.synchronize()
    assert int(zero) == 1                                   # the flag is about the RAW gradient
    sd = float(p2.std())
    assert abs(sd - 0.2) < 0.2 * (4.0 / (2 * n) ** 0.5 + 0.02) and abs(float(p2.mean())) < 4 * 0.2 / n ** 0.5
