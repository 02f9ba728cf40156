"""Deep reactive policy: host-side pytree / oracle checks on CPU, and GPU parity of the whole
closed-loop update (dense layers, heads, T=1 rollout launches chained through GSIN/GS0, weight
gradients) against the torch-CPU oracle restatement of planner.py:1122-1534."""
import math

import numpy as np
import pytest
import torch

from oracle import planner as OP
from oracle.model_eval import OracleModel
from pyrddlgym_jax_b200.core import logic
from pyrddlgym_jax_b200.core.compiler import JaxRDDLCompiler
from pyrddlgym_jax_b200.core.drp import JaxDeepReactivePolicy
from pyrddlgym_jax_b200.core.layout import noise_as_list

from .helpers import fixture_model


def compiled_plan(topology=(8, 8), stochastic=True):
    m = fixture_model("powergen")
    plan = JaxDeepReactivePolicy(topology=list(topology), stochastic=stochastic)
    plan.compile(logic.DefaultJaxRDDLCompilerWithGrad(m), JaxRDDLCompiler(m), {}, 5)
    return m, plan


def test_pytree_roundtrip_and_layer_names():
    m, plan = compiled_plan()
    flat = plan.initial_params(torch.Generator().manual_seed(3))
    tree = plan.params_to_pytree(flat)
    names = set(tree["params"])
    assert names == {"dense_0", "dense_1", "dense_curProd_mu", "dense_curProd_log_sigma"}
    assert tuple(tree["params"]["dense_0"]["kernel"].shape) == (6, 8)
    assert torch.equal(plan.pytree_to_params(tree), flat)
    # biases start at zero, kernels follow variance_scaling(2, fan_in, truncated normal)
    assert float(tree["params"]["dense_1"]["bias"].abs().max()) == 0.0
    k = tree["params"]["dense_1"]["kernel"]
    assert float(k.abs().max()) <= 2.0 * np.sqrt(2.0 / 8) / 0.87962566103423978 + 1e-6


def test_input_order_matches_reference_sorting():
    """planner.py:1302-1313: sorted fluent names, non-bool first."""
    m, plan = compiled_plan()
    assert OP.drp_input_order(m) == ["prevProd", "temperature"]
    assert list(plan.input_perm) == list(range(6))


def test_host_test_policy_matches_oracle():
    m, plan = compiled_plan()
    flat = plan.initial_params(torch.Generator().manual_seed(5)) * 3
    tree = plan.params_to_pytree(flat)["params"]
    fls = {"prevProd": torch.tensor([[1., 2., 0., .5, 3.]]), "temperature": torch.tensor([15.])}
    want = OP.drp_test_actions(m, tree, fls, m.bounds, 2)["curProd"].numpy().reshape(-1)
    vec = np.concatenate([fls[s].numpy().reshape(-1) for s in m.state_fluents])
    got = plan.test_actions_host(flat, vec)["curProd"].reshape(-1)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)


def oracle_update(m, plan, flat, B, T, model_noise, pol_noise, gamma, ent_coeff, relax):
    tree = plan.params_to_pytree(flat)["params"]
    P = {k: {kk: vv.clone().requires_grad_(True) for kk, vv in v.items()} for k, v in tree.items()}
    om = OracleModel(m, relax)
    ents = []

    def pol(t, fls):
        noise = {"curProd": pol_noise[t].reshape(B, 5)}       # fluent-major block [B][n]
        a, e = OP.drp_actions(m, P, {k: fls[k] for k in m.state_fluents}, noise, 1.0, m.bounds,
                              len(plan._topology), plan._stochastic)
        ents.append(e)
        return a
    out = OP.rollout(om, pol, B, T, model_noise)
    loss = OP.mean_loss(out["reward"], gamma) - ent_coeff * torch.stack(ents, 1).sum(1).mean()
    loss.backward()
    grads = {k: {kk: (vv.grad if vv.grad is not None else torch.zeros_like(vv))
                 for kk, vv in v.items()} for k, v in P.items()}
    return out, float(loss.detach()), plan.pytree_to_params(grads)


@pytest.mark.gpu
@pytest.mark.parametrize("topology,B", [((8, 8), 37), ((32, 16), 5), ((64, 64), 200),
                                        ((256, 256), 300), ((64, 128, 32), 129)])
def test_drp_update_matches_oracle(cuda_device, topology, B):
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    m = fixture_model("powergen")
    T = 4
    plan = JaxDeepReactivePolicy(topology=list(topology))
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None,
                            use_cuda_graph=False)
    pl.initialize(7)
    g = torch.Generator().manual_seed(11)
    flat = plan.initial_params(g)
    flat[plan.seg["bh"][0]:plan.seg["bh"][0] + 5] = torch.tensor([2., 3., 4., 1., 5.])   # mu bias
    flat[plan.seg["bh"][0] + 5:plan.seg["bh"][0] + 10] = -1.0                              # log_sigma
    pl.params[0].copy_(flat)
    fw = pl.train_fwd.program
    noise = torch.randn(T, fw.N * B, generator=g)
    pol_noise = torch.randn(T, 5 * B, generator=g)
    pl.train_noise.copy_(noise.reshape(-1))
    pl.drp.pol_noise.copy_(pol_noise)
    pl.drp.ent_coeff.fill_(0.05)
    pl.update()
    torch.cuda.synchronize()
    model_noise = [noise_as_list(fw.noise_layout, noise, t, B) for t in range(T)]
    out, loss, grad = oracle_update(m, plan, flat, B, T, model_noise, pol_noise, 1.0, 0.05,
                                    pl.compiled.relax_config())
    r_ref = out["reward"].detach().numpy().T
    np.testing.assert_allclose(pl.train_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())
    assert abs(float(pl.train_loss[0]) - loss) <= 1e-5 * abs(loss)
    got = pl.grad[0].cpu().numpy()
    np.testing.assert_allclose(got, grad.numpy(), rtol=1e-4, atol=1e-4 * np.abs(grad.numpy()).max())


@pytest.mark.gpu
def test_drp_test_policy_matches_oracle(cuda_device):
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    m = fixture_model("powergen")
    T, B = 5, 33
    plan = JaxDeepReactivePolicy(topology=[16, 16])
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, pgpe=None,
                            use_cuda_graph=False)
    pl.initialize(3)
    te = pl.test_fwd.program
    g = torch.Generator().manual_seed(2)
    noise = torch.randn(T, te.N * B, generator=g)
    pl.test_noise.copy_(noise.reshape(-1))
    flat = pl.params[0].cpu().clone()
    pl.test_loss()
    torch.cuda.synchronize()
    tree = plan.params_to_pytree(flat)["params"]
    om = OracleModel(m, None)
    out = OP.rollout(om, lambda t, fls: OP.drp_test_actions(
        m, tree, {k: fls[k] for k in m.state_fluents}, m.bounds, 2), B, T,
        [noise_as_list(te.noise_layout, noise, t, B) for t in range(T)])
    r_ref = out["reward"].numpy().T
    np.testing.assert_allclose(pl.test_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())
    assert abs(float(pl.test_loss_buf[0]) + float(out["reward"].sum(1).mean())) <= 1e-3


@pytest.mark.gpu
def test_drp_optimises_with_cuda_graph(cuda_device):
    """A few captured iterations run, losses are finite and the policy moves."""
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    m = fixture_model("powergen")
    plan = JaxDeepReactivePolicy(topology=[32, 32])
    pl = JaxBackpropPlanner(m, plan, batch_size_train=64, rollout_horizon=10, optimizer="rmsprop",
                            optimizer_kwargs={"learning_rate": 1e-3}, pgpe=None)
    cb = None
    for cb in pl.optimize_generator(key=1, epochs=6, print_summary=False, print_progress=False):
        pass
    assert np.isfinite(cb["train_return"]).all() and np.isfinite(cb["test_return"]).all()
    assert "dense_0" in cb["best_params"]["params"]


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(37, 10, 6), (128, 128, 8), (200, 64, 64), (1, 16, 300),
                                   (300, 257, 129), (5, 5, 5), (1, 200, 3000), (5, 256, 2100),
                                   (700, 10, 256), (256, 10, 4000), (17, 17, 600)])
@pytest.mark.parametrize("a_trans", [False, True])
def test_sgemm_against_torch(cuda_device, M, N, K, a_trans):
    """rk_sgemm (all epilogues, accumulate, split-K) against torch fp32 matmul."""
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(M * 1000 + N * 10 + K)
    A = torch.randn(M, K, generator=g) / np.sqrt(K)      # O(1) pre-activations
    Bm = torch.randn(K, N, generator=g)
    bias = torch.randn(N, generator=g)
    aux = torch.rand(M, N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    Ad = (A.t().contiguous() if a_trans else A).to(cuda_device)
    lda = M if a_trans else K
    fl = engine.GEMM_A_TRANS if a_trans else 0
    ref = A.double() @ Bm.double()
    # the epilogue is applied to C_in + A*B when ACCUMULATE is set
    cases = [(engine.EPI_NONE, lambda x: x), (engine.EPI_BIAS, lambda x: x + bias),
             (engine.EPI_BIAS_TANH, lambda x: torch.tanh(x + bias)),
             (engine.EPI_DTANH, lambda x: x * (1 - aux.double() ** 2))]
    for split in (1, 3, -8):
        ws = torch.zeros(abs(split) * M * N, device=cuda_device)
        for epi, want in cases:
            for acc in (0, engine.GEMM_ACCUMULATE):
                C = C0.clone().to(cuda_device)
                engine.sgemm(fl | epi | acc, M, N, K, Ad, lda, Bm.to(cuda_device), N, C, N,
                             bias=bias.to(cuda_device), aux=aux.to(cuda_device), ldaux=N,
                             workspace=ws, split_k=split)
                torch.cuda.synchronize()
                w = want(ref + (C0.double() if acc else 0))
                np.testing.assert_allclose(C.cpu().numpy(), w.float().numpy(), rtol=2e-5,
                                           atol=2e-5 * float(w.abs().max()))


@pytest.mark.gpu
def test_action_heads_against_torch(cuda_device):
    from pyrddlgym_jax_b200.core import engine
    B, A = 45, 3
    g = torch.Generator().manual_seed(4)
    heads = torch.randn(B, 2 * A, generator=g) * 3
    heads[:, A:] *= 4                                   # exercise the log_sigma clip
    segs = [(0, 2), (2, 1)]                              # two action fluents: [B][2] and [B][1]
    noise_f = [torch.randn(B, n, generator=g) for _, n in segs]
    gact_f = [torch.randn(B, n, generator=g) for _, n in segs]
    noise = torch.cat(noise_f, 1).t().contiguous()      # [A][B] view used by the torch model
    gact = torch.cat(gact_f, 1).t().contiguous()
    lo, hi = torch.tensor([-1., 0., -5.]), torch.tensor([1., 2., 5.])
    fm = lambda blocks: torch.cat([b.reshape(-1) for b in blocks])     # noqa: E731
    ls_lo, ls_hi = float(np.log(1e-2)), float(np.log(1e2))
    h = heads.clone().requires_grad_(True)
    ls = torch.clamp(h[:, A:], ls_lo, ls_hi)
    act = torch.minimum(torch.maximum(h[:, :A] + torch.exp(ls) * noise.t(), lo), hi)
    (act * gact.t()).sum().backward()
    dev = cuda_device
    out, ent, gh = torch.zeros(A * B, device=dev), torch.zeros(B, device=dev), torch.zeros(B, 2 * A, device=dev)
    engine.drp_actions_forward(B, A, True, segs, heads.to(dev), fm(noise_f).to(dev), 1.0,
                               lo.to(dev), hi.to(dev), ls_lo, ls_hi, out, ent)
    engine.drp_actions_backward(B, A, True, segs, heads.to(dev), fm(noise_f).to(dev), 1.0,
                                lo.to(dev), hi.to(dev), ls_lo, ls_hi, fm(gact_f).to(dev), gh)
    torch.cuda.synchronize()
    want = fm([act.detach()[:, o:o + n] for o, n in segs])
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ent.cpu().numpy(), ls.detach().sum(1).numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(gh.cpu().numpy(), h.grad.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (1000, 256, 256), (333, 64, 128), (4096, 256, 32),
                                   (32768, 256, 256)])
def test_tcgemm_3xtf32_against_fp64(cuda_device, M, N, K):
    """rk_tcgemm (tcgen05 + TMEM + TMA, 3xTF32 split) against an fp64 reference: fp32-grade
    accuracy (a single-pass TF32 product would be ~1e-3 off)."""
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) / np.sqrt(K)).to(cuda_device)
    Bt = torch.randn(N, K, generator=g).to(cuda_device)
    bias = torch.randn(N, generator=g).to(cuda_device)
    aux = torch.rand(M, N, generator=g).to(cuda_device)
    Bh, Bl = torch.empty_like(Bt), torch.empty_like(Bt)
    engine.tf32_split(Bt, Bh, Bl)
    torch.cuda.synchronize()
    assert torch.equal(Bh + Bl, Bt)                                      # the split is exact
    assert int((Bh.view(torch.int32) & 0x1FFF).abs().max()) == 0         # hi is TF32-exact
    ref = A.double() @ Bt.double().t()
    wants = {0: ref, 1: ref + bias, 2: torch.tanh(ref + bias), 3: ref * (1 - aux.double() ** 2)}
    for epi, want in wants.items():
        C = torch.full((M, N), float("nan"), device=cuda_device)
        engine.tcgemm(epi, M, N, K, A, K, Bh, Bl, K, C, N, bias=bias, aux=aux, ldaux=N)
        torch.cuda.synchronize()
        err = float((C.double() - want).abs().max())
        # ~1e-6 of the pre-activation scale (dropped lo*lo term + fp32 accumulation over K)
        assert err <= 4e-6 * max(1.0, float(ref.abs().max())), (epi, err)


@pytest.mark.gpu
def test_layer0_colsum_rowdot_against_torch(cuda_device):
    """The three DRP helper kernels: first layer from fluent-major state blocks (+ packed [x|1]),
    deterministic column sums, one-pass skinny products."""
    from pyrddlgym_jax_b200.core import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(9)
    B, H = 777, 96
    segs = [(0, 5), (5, 1), (6, 3)]
    D = 9
    blocks = [torch.randn(B, n, generator=g) for _, n in segs]
    x_fm = torch.cat([b.reshape(-1) for b in blocks]).to(dev)
    X = torch.cat(blocks, 1)
    W0, b0 = torch.randn(D, H, generator=g) / 3, torch.randn(H, generator=g)
    out, xp = torch.empty(B, H, device=dev), torch.empty(B, D + 1, device=dev)
    engine.drp_layer0(B, D, H, segs, x_fm, W0.to(dev), b0.to(dev), out, xp)
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), torch.tanh(X @ W0 + b0).numpy(), rtol=2e-6, atol=2e-6)
    assert torch.equal(xp.cpu(), torch.cat([X, torch.ones(B, 1)], 1))
    # column sums
    G = torch.randn(B, H, generator=g)
    y0 = torch.randn(H, generator=g)
    for acc in (False, True):
        y = y0.clone().to(dev)
        ws = torch.zeros(64 * H, device=dev)
        engine.colsum(B, H, G.to(dev), H, y, acc, ws, 64)
        torch.cuda.synchronize()
        want = G.double().sum(0) + (y0.double() if acc else 0)
        np.testing.assert_allclose(y.cpu().numpy(), want.float().numpy(), rtol=1e-5, atol=1e-4)
    # skinny products
    for N, K in ((10, 256), (5, 64), (1, 96), (16, 33)):
        A = torch.randn(B, K, generator=g) / np.sqrt(K)
        W = torch.randn(K, N + 3, generator=g)            # a column window of a wider matrix
        bias = torch.randn(N, generator=g)
        o0 = torch.randn(B, N, generator=g)
        for acc in (False, True):
            o = o0.clone().to(dev)
            engine.rowdot(acc, B, N, K, A.to(dev), K, W.to(dev)[:, 2:].contiguous() if False else
                          W.to(dev).view(-1)[2:], N + 3, o, N, bias=bias.to(dev))
            torch.cuda.synchronize()
            want = A.double() @ W[:, 2:2 + N].double() + bias + (o0.double() if acc else 0)
            np.testing.assert_allclose(o.cpu().numpy(), want.float().numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 256, 4000), (64, 64, 1000), (256, 128, 32768),
                                   (96, 32, 37)])
def test_tcgemm_tn_against_fp64(cuda_device, M, N, K):
    """rk_tcgemm_tn: C (+)= A^T B over the batch axis on tcgen05 (MN-major tiles, in-kernel
    3xTF32 split, deterministic split-K) against fp64."""
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(M + 3 * N + K)
    A = torch.randn(K, M, generator=g).to(cuda_device)
    Bm = (torch.randn(K, N, generator=g) / np.sqrt(K)).to(cuda_device)
    C0 = torch.randn(M, N, generator=g).to(cuda_device)
    ws = torch.zeros(148 * M * N, device=cuda_device)
    ref = A.double().t() @ Bm.double()
    for acc in (False, True):
        C = C0.clone()
        engine.tcgemm_tn(acc, M, N, K, A, M, Bm, N, C, ws, 148)
        torch.cuda.synchronize()
        want = ref + (C0.double() if acc else 0)
        err = float((C.double() - want).abs().max())
        assert err <= 4e-6 * max(1.0, float(ref.abs().max())), (acc, err)
    C1, C2 = C0.clone(), C0.clone()
    engine.tcgemm_tn(False, M, N, K, A, M, Bm, N, C1, ws, 148)
    engine.tcgemm_tn(False, M, N, K, A, M, Bm, N, C2, ws, 148)
    torch.cuda.synchronize()
    assert torch.equal(C1, C2)                        # deterministic


@pytest.mark.gpu
@pytest.mark.parametrize("B,hl,NH", [(1000, 256, 10), (77, 64, 3), (4096, 512, 16), (300, 96, 1)])
def test_fused_head_backward_against_fp64(cuda_device, B, hl, NH):
    """rk_drp_head_backward == the four separate products it replaces (fp64 reference)."""
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(B + hl)
    Hl = torch.tanh(torch.randn(B, hl, generator=g)).cuda()
    G = torch.randn(B, NH, generator=g).cuda()
    Wh = (torch.randn(hl, NH, generator=g) * 0.2).cuda()
    gz = torch.zeros(B, hl, device="cuda")
    n = hl + hl * NH + NH
    tail0 = torch.randn(n, generator=g).cuda()
    tail = tail0.clone()
    ws = torch.zeros(engine.drp_head_backward_ctas(B) * n, device="cuda")
    engine.drp_head_backward(B, hl, NH, Hl, G, Wh, gz, tail, ws)
    torch.cuda.synchronize()
    H64, G64, W64 = Hl.double(), G.double(), Wh.double()
    gz_ref = (G64 @ W64.t()) * (1 - H64 * H64)
    want = torch.cat([gz_ref.sum(0), (H64.t() @ G64).reshape(-1), G64.sum(0)]) + tail0.double()
    torch.testing.assert_close(gz.double(), gz_ref, rtol=1e-5, atol=1e-5)
    scale = float(want.abs().max())
    torch.testing.assert_close(tail.double(), want, rtol=1e-5, atol=2e-6 * max(scale, 1.0) * B ** 0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("B,segs,H", [(1000, [(0, 5), (5, 1)], 256), (77, [(0, 1), (1, 3), (4, 4)], 64),
                                      (513, [(0, 2)], 96)])
def test_fused_layer0_backward_against_fp64(cuda_device, B, segs, H):
    """rk_drp_layer0_backward == [x | 1]^T gz0 and gs += gz0 W0^T on the fluent-major state."""
    from pyrddlgym_jax_b200.core import engine
    D = sum(n for _, n in segs)
    g = torch.Generator().manual_seed(B + H)
    X = torch.randn(B, D, generator=g)                       # row-major copy for the reference
    x_fm = torch.cat([X[:, o:o + n].reshape(-1) for o, n in segs]).cuda()
    gz0 = torch.randn(B, H, generator=g).cuda()
    W0 = (torch.randn(D, H, generator=g) * 0.3).cuda()
    gs0 = torch.randn(D * B, generator=g).cuda()
    gs = gs0.clone()
    n = (D + 1) * H
    gw0 = torch.randn(n, generator=g).cuda()
    gw = gw0.clone()
    ws = torch.zeros(engine.drp_head_backward_ctas(B) * n, device="cuda")
    engine.drp_layer0_backward(B, D, H, segs, x_fm, gz0, W0.t().contiguous(), gs, gw, ws)
    torch.cuda.synchronize()
    X64, g64, W64 = X.double().cuda(), gz0.double(), W0.double()
    want_w = torch.cat([(X64.t() @ g64).reshape(-1), g64.sum(0)]) + gw0.double()
    dx = g64 @ W64.t()                                       # [B, D]
    want_s = torch.cat([dx[:, o:o + k].reshape(-1) for o, k in segs]) + gs0.double()
    torch.testing.assert_close(gw.double(), want_w, rtol=1e-5,
                               atol=2e-6 * max(float(want_w.abs().max()), 1.0) * B ** 0.5)
    torch.testing.assert_close(gs.double(), want_s, rtol=1e-5, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("B,segs,h0,h1,NH", [(300, [(0, 5), (5, 1)], 256, 256, 10),
                                              (129, [(0, 16)], 64, 128, 16),
                                              (1000, [(0, 1), (1, 2)], 32, 32, 1),
                                              (4096, [(0, 3), (3, 4), (7, 2)], 128, 256, 7),
                                              (40000, [(0, 2), (2, 4)], 256, 256, 10)])
@pytest.mark.parametrize("keep", [True, False])
def test_fused_mlp2_forward_against_fp64(cuda_device, B, segs, h0, h1, NH, keep):
    """rk_drp_mlp2_forward (layer 0 computed into the tcgen05 operand tiles, hidden layer 3xTF32, head
    product in the TMEM epilogue) against an fp64 evaluation of the same two-hidden-layer network."""
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(B + h0 + NH)
    D = sum(n for _, n in segs)
    x = torch.randn(B, D, generator=g, dtype=torch.float64) * 1.5
    W0 = torch.randn(D, h0, generator=g, dtype=torch.float64) * (1.0 / D ** 0.5)
    b0 = torch.randn(h0, generator=g, dtype=torch.float64) * 0.2
    W1 = torch.randn(h0, h1, generator=g, dtype=torch.float64) * (1.5 / h0 ** 0.5)
    b1 = torch.randn(h1, generator=g, dtype=torch.float64) * 0.2
    Wh = torch.randn(h1, NH, generator=g, dtype=torch.float64) * (1.0 / h1 ** 0.5)
    bh = torch.randn(NH, generator=g, dtype=torch.float64)
    f32 = lambda t: t.float().contiguous().to(cuda_device)      # noqa: E731
    xs, W0s, b0s, W1s, b1s, Whs, bhs = (f32(t) for t in (x, W0, b0, W1, b1, Wh, bh))
    # reference on the fp32-rounded inputs
    xr, W0r, b0r, W1r, b1r, Whr, bhr = (t.double().cpu() for t in (xs, W0s, b0s, W1s, b1s, Whs, bhs))
    H0r = torch.tanh(xr @ W0r + b0r)
    H1r = torch.tanh(H0r @ W1r + b1r)
    want = H1r @ Whr + bhr
    # fluent-major state blocks
    x_fm = torch.cat([xs[:, off:off + n].contiguous().reshape(-1) for off, n in segs])
    W1T = W1s.t().contiguous()
    hi, lo = torch.empty_like(W1T), torch.empty_like(W1T)
    engine.tf32_split(W1T.reshape(-1), hi.reshape(-1), lo.reshape(-1), W1T.numel())
    H0 = torch.zeros(B, h0, device=cuda_device) if keep else None
    H1 = torch.zeros(B, h1, device=cuda_device) if keep else None
    heads = torch.zeros(B, NH, device=cuda_device)
    img_f = torch.zeros(engine.drp_mlp2_image_floats(D, h0, h1, NH, False), device=cuda_device)
    img_b = torch.zeros(engine.drp_mlp2_image_floats(D, h0, h1, NH, True), device=cuda_device)
    engine.drp_mlp2_prepare(D, h0, h1, NH, W0s.reshape(-1), Whs.reshape(-1), img_f, img_b)
    engine.drp_mlp2_forward("tanh", B, D, h0, h1, NH, segs, x_fm, img_f, b0s, hi.reshape(-1),
                            lo.reshape(-1), b1s, bhs, H0, H1, heads)
    torch.cuda.synchronize()
    torch.testing.assert_close(heads.double().cpu(), want, rtol=2e-5, atol=2e-5)
    if keep:
        # tanh on the MUFU units: absolute error <= 2e-7 per activation, and within one fp32 ulp of the
        # exact value where the unit saturates (the adjoint's 1 - h^2 depends on the last bits there)
        torch.testing.assert_close(H0.double().cpu(), H0r, rtol=1e-6, atol=5e-7)
        sat = H0r.abs() > 0.99
        assert float((H0.double().cpu() - H0r)[sat].abs().max()) <= 6.1e-8
        # the tensor core truncates when it aligns products to the accumulator (not IEEE round-to-nearest):
        # over the 96 accumulating MMAs of a 256-wide layer that is up to ~1e-5 of the pre-activation
        torch.testing.assert_close(H1.double().cpu(), H1r, rtol=1e-5, atol=1.5e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("act", ["tanh", "softplus"])
@pytest.mark.parametrize("B,segs,h0,h1,NH", [(300, [(0, 5), (5, 1)], 256, 256, 10),
                                              (129, [(0, 16)], 64, 128, 16),
                                              (1000, [(0, 1), (1, 2)], 32, 32, 1),
                                              (40000, [(0, 3), (3, 4), (7, 2)], 128, 256, 7)])
def test_fused_mlp2_backward_against_fp64(cuda_device, B, segs, h0, h1, NH, act):
    """rk_drp_mlp2_backward: gz1 = (g Wh^T) act'(H1), gz0 = (gz1 W1^T) act'(H0), gs += gz0 W0^T against
    fp64 on the same fp32 inputs (the second and third products run as 3xTF32 on the tensor cores)."""
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(B + h0 + NH + 1)
    D = sum(n for _, n in segs)
    f32 = lambda t: t.float().contiguous().to(cuda_device)      # noqa: E731
    if act == "tanh":
        H0 = torch.tanh(torch.randn(B, h0, generator=g, dtype=torch.float64))
        H1 = torch.tanh(torch.randn(B, h1, generator=g, dtype=torch.float64))
        dact = lambda y: 1 - y * y                                  # noqa: E731
    else:
        H0 = torch.nn.functional.softplus(torch.randn(B, h0, generator=g, dtype=torch.float64))
        H1 = torch.nn.functional.softplus(torch.randn(B, h1, generator=g, dtype=torch.float64))
        dact = lambda y: 1 - torch.exp(-y)                          # noqa: E731
    W0 = torch.randn(D, h0, generator=g, dtype=torch.float64) * (1.0 / D ** 0.5)
    W1 = torch.randn(h0, h1, generator=g, dtype=torch.float64) * (1.5 / h0 ** 0.5)
    Wh = torch.randn(h1, NH, generator=g, dtype=torch.float64) * (1.0 / h1 ** 0.5)
    G = torch.randn(B, NH, generator=g, dtype=torch.float64)
    gs0 = torch.randn(D * B, generator=g, dtype=torch.float64)
    H0s, H1s, W0s, W1s, Whs, Gs, gs = (f32(t) for t in (H0, H1, W0, W1, Wh, G, gs0))
    H0r, H1r, W0r, W1r, Whr, Gr, gsr = (t.double().cpu() for t in (H0s, H1s, W0s, W1s, Whs, Gs, gs))
    gz1r = (Gr @ Whr.t()) * dact(H1r)
    gz0r = (gz1r @ W1r.t()) * dact(H0r)
    dx = gz0r @ W0r.t()
    want_gs = gsr + torch.cat([dx[:, off:off + n].contiguous().reshape(-1) for off, n in segs])
    hi, lo = torch.empty_like(W1s), torch.empty_like(W1s)
    engine.tf32_split(W1s.reshape(-1), hi.reshape(-1), lo.reshape(-1), W1s.numel())
    gz0 = torch.zeros(B, h0, device=cuda_device)
    gz1 = torch.zeros(B, h1, device=cuda_device)
    img_f = torch.zeros(engine.drp_mlp2_image_floats(D, h0, h1, NH, False), device=cuda_device)
    img_b = torch.zeros(engine.drp_mlp2_image_floats(D, h0, h1, NH, True), device=cuda_device)
    engine.drp_mlp2_prepare(D, h0, h1, NH, W0s.reshape(-1), Whs.reshape(-1), img_f, img_b)
    engine.drp_mlp2_backward(act, B, D, h0, h1, NH, segs, Gs, H0s, H1s, img_b, hi.reshape(-1),
                             lo.reshape(-1), gz0, gz1, gs)
    torch.cuda.synchronize()
    s1, s0, sx = float(gz1r.abs().max()), float(gz0r.abs().max()), float(dx.abs().max())
    torch.testing.assert_close(gz1.double().cpu(), gz1r, rtol=1e-5, atol=1e-6 * s1)
    torch.testing.assert_close(gz0.double().cpu(), gz0r, rtol=1e-5, atol=2e-6 * s0)
    torch.testing.assert_close(gs.double().cpu(), want_gs, rtol=1e-5, atol=2e-6 * sx)


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(256, 256, 40000), (64, 32, 1000), (160, 96, 3000), (256, 128, 17)])
def test_wgrad_hidden_against_fp64(cuda_device, M, N, K):
    """rk_drp_wgrad_hidden: dW += A^T B over all K rows in one launch (tcgen05 3xTF32, MN-major
    operands), db += column sums of B taken by the converter warps."""
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(K, M, generator=g).to(cuda_device)
    Bm = torch.randn(K, N, generator=g).to(cuda_device)
    dW0 = torch.randn(M, N, generator=g).to(cuda_device)
    db0 = torch.randn(N, generator=g).to(cuda_device)
    dW, db = dW0.clone(), db0.clone()
    ws = torch.zeros(engine.drp_wgrad_ctas() * (M * N + N), device=cuda_device)
    engine.drp_wgrad_hidden(True, M, N, K, A, Bm, dW, db, ws)
    torch.cuda.synchronize()
    want = dW0.double() + A.double().t() @ Bm.double()
    want_b = db0.double() + Bm.double().sum(0)
    # entries are sums of K products of unit normals (|entry| ~ sqrt(K)); the tensor core's truncating
    # accumulator leaves ~4e-6 of that
    torch.testing.assert_close(dW.double(), want, rtol=1e-5, atol=6e-6 * K ** 0.5)
    torch.testing.assert_close(db.double(), want_b, rtol=1e-5, atol=1e-6 * K ** 0.5)
    dW2 = torch.empty_like(dW)
    engine.drp_wgrad_hidden(False, M, N, K, A, Bm, dW2, None, ws)        # overwrite, no bias sum
    torch.cuda.synchronize()
    torch.testing.assert_close(dW2.double(), want - dW0.double(), rtol=1e-5, atol=6e-6 * K ** 0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("M,NH,K", [(256, 10, 40000), (32, 16, 500), (128, 1, 3001)])
def test_wgrad_heads_against_fp64(cuda_device, M, NH, K):
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(M + NH + K)
    H = torch.randn(K, M, generator=g).to(cuda_device)
    G = torch.randn(K, NH, generator=g).to(cuda_device)
    d0 = torch.randn(M, NH, generator=g).to(cuda_device)
    b0 = torch.randn(NH, generator=g).to(cuda_device)
    d, db = d0.clone(), b0.clone()
    ws = torch.zeros(engine.drp_wgrad_ctas() * (M + 1) * 16, device=cuda_device)
    engine.drp_wgrad_heads(True, M, NH, K, H, G, d, db, ws)
    torch.cuda.synchronize()
    torch.testing.assert_close(db.double(), b0.double() + G.double().sum(0), rtol=1e-5, atol=1e-6 * K ** 0.5)
    want = d0.double() + H.double().t() @ G.double()
    torch.testing.assert_close(d.double(), want, rtol=1e-5, atol=6e-6 * K ** 0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("M,segs,T,Bt,extra", [(256, [(0, 5), (5, 1)], 7, 3000, 0),
                                                (64, [(0, 15)], 3, 100, 2),
                                                (128, [(0, 1), (1, 2), (3, 4)], 2, 1001, 1)])
def test_wgrad_input_against_fp64(cuda_device, M, segs, T, Bt, extra):
    """rk_drp_wgrad_input: dW0 += x^T gz0 and db0 += column sums of gz0 with x gathered from the
    fluent-major state tape of every epoch (``extra`` further state words the policy does not read)."""
    from pyrddlgym_jax_b200.core import engine
    g = torch.Generator().manual_seed(M + T + Bt)
    D = sum(n for _, n in segs)
    S = D + extra
    x = torch.randn(T, Bt, S, generator=g)
    gz0 = torch.randn(T * Bt, M, generator=g).to(cuda_device)
    all_segs = list(segs) + ([(D, extra)] if extra else [])
    tape = torch.cat([torch.cat([x[t][:, off:off + n].contiguous().reshape(-1) for off, n in all_segs])
                      for t in range(T)]).to(cuda_device)
    dW0_0 = torch.randn(D, M, generator=g).to(cuda_device)
    db0_0 = torch.randn(M, generator=g).to(cuda_device)
    dW0, db0 = dW0_0.clone(), db0_0.clone()
    ws = torch.zeros(engine.drp_wgrad_ctas() * M * 16, device=cuda_device)
    engine.drp_wgrad_input(True, M, D, T, Bt, S, segs, gz0, tape, dW0, db0, ws)
    torch.cuda.synchronize()
    X = x[:, :, :D].reshape(T * Bt, D).double().to(cuda_device)
    want = dW0_0.double() + X.t() @ gz0.double()
    want_b = db0_0.double() + gz0.double().sum(0)
    K = T * Bt
    torch.testing.assert_close(dW0.double(), want, rtol=1e-5, atol=6e-6 * K ** 0.5)
    torch.testing.assert_close(db0.double(), want_b, rtol=1e-5, atol=1e-6 * K ** 0.5)


def _segment_noise(plan, m, pol_noise_t, B):
    """{action fluent: [B, n]} views of one step's fluent-major policy-noise row."""
    out = {}
    for var, (off, n, _) in plan.layout.items():
        out[var] = pol_noise_t[off * B:(off + n) * B].reshape(B, n)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("stochastic", [True, False])
def test_drp_bool_heads_match_oracle(cuda_device, stochastic):
    """Boolean action heads (planner.py:1371-1391): sigmoid(w (logit + logistic noise)) in the
    train policy with the Bernoulli entropy (value AND gradient), (p > 0.5) in the test policy;
    boolean state fluents feed the first layer as 0 / 1."""
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    m = fixture_model("wildfire_free")
    T, B = 3, 21
    plan = JaxDeepReactivePolicy(topology=[16, 16], stochastic=stochastic, sigmoid_weight=2.0)
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None,
                            use_cuda_graph=False)
    pl.initialize(7)
    g = torch.Generator().manual_seed(11)
    flat = plan.initial_params(g)
    pl.params[0].copy_(flat)
    fw = pl.train_fwd.program
    noise = torch.rand(T, fw.N * B, generator=g)
    pl.train_noise.copy_(noise.reshape(-1))
    pol_noise = torch.randn(T, plan.A * B, generator=g)
    if stochastic:
        pl.drp.pol_noise.copy_(pol_noise)
    pl.drp.ent_coeff.fill_(0.05)
    pl.update()
    torch.cuda.synchronize()
    tree = plan.params_to_pytree(flat)["params"]
    assert "dense_cut_out_logit" in tree and "dense_cut_out_mu" not in tree
    P = {k: {kk: vv.clone().requires_grad_(True) for kk, vv in v.items()} for k, v in tree.items()}
    om = OracleModel(m, pl.compiled.relax_config())
    ents = []

    def pol(t, fls):
        a, e = OP.drp_actions(m, P, {k: fls[k] for k in m.state_fluents},
                              _segment_noise(plan, m, pol_noise[t], B), 1.0, m.bounds, 2,
                              stochastic, sigmoid_weight=2.0)
        ents.append(e)
        return a
    out = OP.rollout(om, pol, B, T, [noise_as_list(fw.noise_layout, noise, t, B) for t in range(T)])
    loss = OP.mean_loss(out["reward"], float(m.discount))
    if stochastic:
        loss = loss - 0.05 * torch.stack(ents, 1).sum(1).mean()
    loss.backward()
    grads = {k: {kk: (vv.grad if vv.grad is not None else torch.zeros_like(vv))
                 for kk, vv in v.items()} for k, v in P.items()}
    want = plan.pytree_to_params(grads).numpy()
    r_ref = out["reward"].detach().numpy().T
    np.testing.assert_allclose(pl.train_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())
    assert abs(float(pl.train_loss[0]) - float(loss)) <= 2e-5 * abs(float(loss))
    got = pl.grad[0].cpu().numpy()
    live = np.abs(want) > 0          # the unused log_sigma columns of boolean heads get no gradient
    np.testing.assert_allclose(got[live], want[live], rtol=1e-4, atol=1e-4 * np.abs(want).max())
    assert np.all(got[~live] == 0)
    # test policy on the exact model
    te = pl.test_fwd.program
    tnoise = torch.rand(T, te.N * B, generator=g)
    pl.test_noise.copy_(tnoise.reshape(-1))
    pl.test_loss()
    torch.cuda.synchronize()
    om = OracleModel(m, None)
    out = OP.rollout(om, lambda t, fls: OP.drp_test_actions(
        m, tree, {k: fls[k] for k in m.state_fluents}, m.bounds, 2, stochastic,
        sigmoid_weight=2.0), B, T,
        [noise_as_list(te.noise_layout, tnoise, t, B) for t in range(T)])
    np.testing.assert_allclose(pl.test_reward[0].view(T, B).cpu().numpy(),
                               out["reward"].numpy().T, rtol=1e-5, atol=1e-4)


@pytest.mark.gpu
def test_drp_wrap_non_bool_heads_match_oracle(cuda_device):
    """wrap_non_bool heads (planner.py:1411-1415): the sigmoid / softplus box wrap replaces the clip
    in the train policy (value and gradient); the test policy applies wrap then clip."""
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    m = fixture_model("powergen")
    T, B = 3, 19
    plan = JaxDeepReactivePolicy(topology=[16, 16], wrap_non_bool=True)
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None,
                            use_cuda_graph=False)
    pl.initialize(7)
    g = torch.Generator().manual_seed(11)
    flat = plan.initial_params(g)
    pl.params[0].copy_(flat)
    fw = pl.train_fwd.program
    noise = torch.randn(T, fw.N * B, generator=g)
    pol_noise = torch.randn(T, 5 * B, generator=g)
    pl.train_noise.copy_(noise.reshape(-1))
    pl.drp.pol_noise.copy_(pol_noise)
    pl.drp.ent_coeff.fill_(0.05)
    pl.update()
    torch.cuda.synchronize()
    tree = plan.params_to_pytree(flat)["params"]
    P = {k: {kk: vv.clone().requires_grad_(True) for kk, vv in v.items()} for k, v in tree.items()}
    om = OracleModel(m, pl.compiled.relax_config())
    ents = []

    def pol(t, fls):
        a, e = OP.drp_actions(m, P, {k: fls[k] for k in m.state_fluents},
                              {"curProd": pol_noise[t].reshape(B, 5)}, 1.0, m.bounds, 2, True,
                              wrap_non_bool=True)
        ents.append(e)
        return a
    out = OP.rollout(om, pol, B, T, [noise_as_list(fw.noise_layout, noise, t, B) for t in range(T)])
    loss = OP.mean_loss(out["reward"], float(m.discount)) - 0.05 * torch.stack(ents, 1).sum(1).mean()
    loss.backward()
    grads = {k: {kk: (vv.grad if vv.grad is not None else torch.zeros_like(vv))
                 for kk, vv in v.items()} for k, v in P.items()}
    want = plan.pytree_to_params(grads).numpy()
    assert abs(float(pl.train_loss[0]) - float(loss.detach())) <= 2e-5 * abs(float(loss.detach()))
    np.testing.assert_allclose(pl.grad[0].cpu().numpy(), want, rtol=1e-4,
                               atol=1e-4 * np.abs(want).max())
    # host-side test policy against the oracle's
    fls = {"prevProd": torch.tensor([[1., 2., 0., .5, 3.]]), "temperature": torch.tensor([15.])}
    ref = OP.drp_test_actions(m, tree, fls, m.bounds, 2, wrap_non_bool=True)["curProd"].numpy()
    vec = np.concatenate([fls[s].numpy().reshape(-1) for s in m.state_fluents])
    got = plan.test_actions_host(flat, vec)["curProd"].reshape(-1)
    np.testing.assert_allclose(got, ref.reshape(-1), rtol=1e-5, atol=1e-5)


# ---- input LayerNorm (planner.py:1300-1319) -----------------------------------------------------
def _with_grad(tree):
    if isinstance(tree, dict):
        return {k: _with_grad(v) for k, v in tree.items()}
    return tree.clone().requires_grad_(True)


def _grads(tree):
    if isinstance(tree, dict):
        return {k: _grads(v) for k, v in tree.items()}
    return tree.grad if tree.grad is not None else torch.zeros_like(tree)


def _normalised_plan(fixture, per_layer, topology=(16, 16)):
    m = fixture_model(fixture)
    plan = JaxDeepReactivePolicy(topology=list(topology), normalize=True,
                                 normalize_per_layer=per_layer)
    plan.compile(logic.DefaultJaxRDDLCompilerWithGrad(m), JaxRDDLCompiler(m), {}, 5)
    return m, plan


def test_oracle_layernorm_against_torch():
    """drp_state_vector restates flax nn.LayerNorm (biased variance, epsilon 1e-6)."""
    m = fixture_model("powergen")
    g = torch.Generator().manual_seed(1)
    fls = {"prevProd": torch.randn(7, 5, generator=g) * 3 + 1, "temperature": torch.randn(7, generator=g)}
    scale, bias = torch.randn(6, generator=g), torch.randn(6, generator=g)
    got = OP.drp_state_vector(m, {"inputs": {"norm": {"scale": scale, "bias": bias}}}, fls, True)
    x = torch.cat([fls["prevProd"], fls["temperature"].reshape(7, 1)], 1)
    want = torch.nn.functional.layer_norm(x.double(), (6,), scale.double(), bias.double(), 1e-6)
    np.testing.assert_allclose(got.numpy(), want.float().numpy(), rtol=1e-5, atol=1e-6)
    # a size-1 fluent turns per-layer normalisation off (planner.py:1268-1274)
    assert OP.drp_norm_layers(m, True, True) == [("norm", ["prevProd", "temperature"])]


@pytest.mark.parametrize("fixture,per_layer", [("powergen", False), ("quadcopter", True),
                                               ("quadcopter", False)])
def test_layernorm_pytree_and_host_policy(fixture, per_layer):
    m, plan = _normalised_plan(fixture, per_layer)
    g = torch.Generator().manual_seed(4)
    flat = plan.initial_params(g)
    so, sb = plan.seg["ln_scale"][0], plan.seg["ln_bias"][0]
    assert torch.all(flat[so:so + plan.Dn] == 1) and torch.all(flat[sb:sb + plan.Dn] == 0)
    flat[so:so + 2 * plan.Dn] += torch.randn(2 * plan.Dn, generator=g) * 0.3
    tree = plan.params_to_pytree(flat)
    mods = set(tree["params"]["inputs"])
    non_bool = [v for v in OP.drp_input_order(m) if m.variable_ranges[v] != "bool"]
    if per_layer:
        assert mods == {"norm_" + v.replace("-", "_") for v in non_bool}
    else:
        assert mods == {"norm"}
        assert tree["params"]["inputs"]["norm"]["scale"].numel() == plan.Dn
    assert torch.equal(plan.pytree_to_params(tree), flat)
    assert OP.drp_norm_layers(m, True, per_layer) == plan.norm_layers
    fls = {v: torch.randn(1, m.variable_size(v), generator=g) * 2 for v in m.state_fluents}
    ref = OP.drp_test_actions(m, tree["params"], fls, plan.bounds, 2, normalize=True,
                              normalize_per_layer=per_layer)
    vec = np.concatenate([fls[s].numpy().reshape(-1) for s in m.state_fluents])
    got = plan.test_actions_host(flat, vec)
    for var in m.action_fluents:
        np.testing.assert_allclose(got[var].reshape(-1), ref[var].numpy().reshape(-1), rtol=1e-5,
                                   atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("B,segs,groups", [(1000, [(0, 5), (5, 1)], [1, 1]),
                                           (77, [(0, 4), (4, 3), (7, 2), (9, 4)], [1, 0, 2, 1]),
                                           (129, [(0, 1), (1, 6)], [0, 1])])
def test_state_layernorm_kernels_against_fp64(cuda_device, B, segs, groups):
    from pyrddlgym_jax_b200.core import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    D = sum(n for _, n in segs)
    Dn = sum(n for (_, n), k in zip(segs, groups) if k)
    eps = 1e-6
    # rows with well-separated words: E[x^2] - E[x]^2 in fp32 (the reference's use_fast_variance)
    # loses digits when the words of a group nearly coincide, which says nothing about the kernel
    blocks = [(torch.randn(B, n, generator=g) * 0.4 + torch.linspace(-1.5, 1.5, n) + 0.6 * (f - 1))
              .double().requires_grad_(True) for f, (_, n) in enumerate(segs)]
    scale = (torch.randn(Dn, generator=g)).double().requires_grad_(True)
    bias = (torch.randn(Dn, generator=g)).double().requires_grad_(True)
    gyb = [torch.randn(B, n, generator=g).double() for _, n in segs]
    gx0 = [torch.randn(B, n, generator=g) for _, n in segs]
    # fp64 reference: parameters follow the normalised words in fluent order
    jpos, j = [], 0
    for (_, n), k in zip(segs, groups):
        jpos.append(j)
        j += n if k else 0
    ys = list(blocks)
    for grp in sorted(set(groups) - {0}):
        mem = [f for f, k in enumerate(groups) if k == grp]
        x = torch.cat([blocks[f] for f in mem], 1)
        idx = torch.cat([torch.arange(jpos[f], jpos[f] + segs[f][1]) for f in mem])
        mean = x.mean(1, keepdim=True)
        var = (x * x).mean(1, keepdim=True) - mean * mean
        y = (x - mean) * torch.rsqrt(var + eps) * scale[idx] + bias[idx]
        o = 0
        for f in mem:
            ys[f] = y[:, o:o + segs[f][1]]
            o += segs[f][1]
    sum((y * gg).sum() for y, gg in zip(ys, gyb)).backward()
    fm = lambda bl: torch.cat([b.detach().float().reshape(-1) for b in bl]).to(dev)   # noqa: E731
    x_fm, gy_fm, gx = fm(blocks), fm(gyb), fm(gx0)
    y_fm = torch.empty_like(x_fm)
    sc, bi = scale.detach().float().to(dev), bias.detach().float().to(dev)
    engine.state_layernorm_forward(B, D, segs, groups, eps, x_fm, sc, bi, y_fm)
    gsb0 = torch.randn(2 * Dn, generator=g)
    gsb = gsb0.clone().to(dev)
    ws = torch.zeros(engine.state_layernorm_ctas(B) * 2 * Dn, device=dev)
    engine.state_layernorm_backward(B, D, segs, groups, eps, x_fm, sc, gy_fm, gx, gsb, ws)
    torch.cuda.synchronize()
    np.testing.assert_allclose(y_fm.cpu().numpy(), fm(ys).cpu().numpy(), rtol=1e-5, atol=1e-5)
    want_gx = torch.cat([(b.grad + g0.double()).float().reshape(-1) for b, g0 in zip(blocks, gx0)])
    np.testing.assert_allclose(gx.cpu().numpy(), want_gx.numpy(), rtol=1e-4,
                               atol=1e-4 * float(want_gx.abs().max()))
    want_sb = torch.cat([scale.grad, bias.grad]).float() + gsb0
    np.testing.assert_allclose(gsb.cpu().numpy(), want_sb.numpy(), rtol=1e-4,
                               atol=1e-4 * float(want_sb.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("fixture,per_layer,eps", [("powergen", False, 1e-6),
                                                   ("quadcopter", True, 0.05),
                                                   ("quadcopter", False, 1e-6)])
def test_drp_layernorm_update_matches_oracle(cuda_device, fixture, per_layer, eps):
    """normalize / normalize_per_layer: rewards, loss and the whole gradient (dense layers, heads,
    LayerNorm scale and bias) of one update, and the exact test rollouts, against the oracle.
    (The drones of the Quadcopter fixture start with equal words in most fluents: a per-fluent variance
    of ~0 under epsilon 1e-6 amplifies fp32 rounding a thousandfold in any implementation, so the
    per-layer case runs with normalizer_kwargs epsilon = 0.05.)"""
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    m = fixture_model(fixture)
    T, B = 3, 23
    plan = JaxDeepReactivePolicy(topology=[16, 16], normalize=True, normalize_per_layer=per_layer,
                                 normalizer_kwargs={"use_bias": True, "use_scale": True, "epsilon": eps})
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None,
                            use_cuda_graph=False)
    pl.initialize(7)
    g = torch.Generator().manual_seed(11)
    flat = plan.initial_params(g)
    so = plan.seg["ln_scale"][0]
    flat[so:so + 2 * plan.Dn] += torch.randn(2 * plan.Dn, generator=g) * 0.2
    pl.params[0].copy_(flat)
    fw = pl.train_fwd.program
    noise = torch.randn(T, fw.N * B, generator=g)
    pol_noise = torch.randn(T, plan.A * B, generator=g)
    if pl.train_noise is not None:
        pl.train_noise.copy_(noise.reshape(-1))
    pl.drp.pol_noise.copy_(pol_noise)
    pl.drp.ent_coeff.fill_(0.05)
    pl.update()
    torch.cuda.synchronize()
    tree = plan.params_to_pytree(flat)["params"]
    P = _with_grad(tree)
    om = OracleModel(m, pl.compiled.relax_config())
    ents = []

    def pol(t, fls):
        a, e = OP.drp_actions(m, P, {k: fls[k] for k in m.state_fluents},
                              _segment_noise(plan, m, pol_noise[t], B), 1.0, plan.bounds, 2, True,
                              normalize=True, normalize_per_layer=per_layer, norm_eps=eps)
        ents.append(e)
        return a
    out = OP.rollout(om, pol, B, T, [noise_as_list(fw.noise_layout, noise, t, B) for t in range(T)])
    loss = OP.mean_loss(out["reward"], float(m.discount)) - 0.05 * torch.stack(ents, 1).sum(1).mean()
    loss.backward()
    want = plan.pytree_to_params(_grads(P)).numpy()
    r_ref = out["reward"].detach().numpy().T
    np.testing.assert_allclose(pl.train_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())
    assert abs(float(pl.train_loss[0]) - float(loss.detach())) <= 2e-5 * abs(float(loss.detach()))
    got = pl.grad[0].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * np.abs(want).max())
    assert np.abs(want[so:so + 2 * plan.Dn]).max() > 0
    # exact test rollouts with the normalised test policy
    te = pl.test_fwd.program
    tnoise = torch.randn(T, te.N * B, generator=g)
    if pl.test_noise is not None:
        pl.test_noise.copy_(tnoise.reshape(-1))
    pl.test_loss()
    torch.cuda.synchronize()
    out = OP.rollout(OracleModel(m, None), lambda t, fls: OP.drp_test_actions(
        m, tree, {k: fls[k] for k in m.state_fluents}, plan.bounds, 2, normalize=True,
        normalize_per_layer=per_layer, norm_eps=eps), B, T,
        [noise_as_list(te.noise_layout, tnoise, t, B) for t in range(T)])
    r_ref = out["reward"].numpy().T
    np.testing.assert_allclose(pl.test_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())


# ---- dense_topk head + GumbelSoftmaxTopK (planner.py:1027-1080, 1435-1447) ------------------------
def _topk_model(allowed):
    m = fixture_model("wildfire")
    m.max_allowed_actions = allowed
    return m


def test_oracle_gumbel_softmax_topk_known_answers():
    g = torch.Generator().manual_seed(2)
    logits = torch.randn(5, 7, generator=g)
    gum = -torch.log(-torch.log(torch.rand(5, 7, generator=g)))
    # allowed == 1: a softmax of the perturbed logits (the weight is not applied, planner.py:1045)
    a = OP.gumbel_softmax_topk(logits, gum, 1, 3.0, 1.0)
    np.testing.assert_allclose(a.numpy(), torch.softmax(logits + gum, 1).numpy(), rtol=1e-6)
    # soft top-k: total mass = allowed (single entries may pass 1: the reference masks with the
    # accumulated khot every round); a large weight approaches the k-hot
    k = OP.gumbel_softmax_topk(logits, gum, 3, 1.0, 1.0)
    np.testing.assert_allclose(k.sum(1).numpy(), 3.0, rtol=1e-5)
    assert float(k.min()) >= 0.0
    logits = torch.stack([torch.randperm(7, generator=g).float() for _ in range(5)])
    sharp = OP.gumbel_softmax_topk(logits, 0.0, 3, 200.0, 1.0)
    hard = OP.gumbel_softmax_topk(logits, 0.0, 3, 200.0, 0.0)
    assert torch.equal(hard.sum(1), torch.full((5,), 3.0))
    want = torch.zeros_like(logits).scatter_(1, torch.topk(logits, 3, dim=1).indices, 1.0)
    assert torch.equal(hard, want)
    np.testing.assert_allclose(sharp.numpy(), hard.numpy(), atol=1e-3)


@pytest.mark.parametrize("allowed", [1, 3])
def test_topk_pytree_and_host_policy(allowed):
    m = _topk_model(allowed)
    plan = JaxDeepReactivePolicy(topology=[16, 16], softmax_weight=2.0)
    plan.compile(logic.DefaultJaxRDDLCompilerWithGrad(m), JaxRDDLCompiler(m), {}, 5)
    assert plan.use_topk and set(plan.action_kinds) == {3}
    flat = plan.initial_params(torch.Generator().manual_seed(8)) * 2
    tree = plan.params_to_pytree(flat)["params"]
    assert set(tree) == {"dense_0", "dense_1", "dense_topk"}
    assert tuple(tree["dense_topk"]["kernel"].shape) == (16, plan.A)
    back = plan.pytree_to_params({"params": tree})
    live = torch.ones(plan.n_params, dtype=torch.bool)       # unused log_sigma columns carry no weights
    off, r, c = plan.seg["Wh"]
    live[off:off + r * c].view(r, c)[:, plan.A:] = False
    live[plan.seg["bh"][0] + plan.A:plan.seg["bh"][0] + c] = False
    assert torch.equal(back[live], flat[live])
    g = torch.Generator().manual_seed(3)
    fls = {v: (torch.rand(1, m.variable_size(v), generator=g) > 0.5) for v in m.state_fluents}
    ref = OP.drp_test_actions(m, tree, fls, plan.bounds, 2, softmax_weight=2.0)
    vec = np.concatenate([fls[s].numpy().reshape(-1) for s in m.state_fluents]).astype(np.float32)
    got = plan.test_actions_host(flat, vec)
    n_on = 0
    for var in m.action_fluents:
        assert np.array_equal(got[var].reshape(-1), ref[var].numpy().reshape(-1))
        n_on += int(got[var].sum())
    assert n_on <= allowed and (allowed == 1 or n_on == allowed)


@pytest.mark.gpu
@pytest.mark.parametrize("allowed,stochastic", [(1, True), (3, True), (2, False), (8, True)])
def test_topk_head_kernels_against_oracle(cuda_device, allowed, stochastic):
    """rk_drp_topk_forward / backward against torch autograd of the oracle's GumbelSoftmaxTopK
    (value, entropy, logits cotangent incl. the entropy term) and the exact k-hot of the test form."""
    from pyrddlgym_jax_b200.core import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(allowed)
    B, w, coeff = 203, 1.7, 0.3
    segs = [(0, 2), (2, 9), (11, 1), (12, 6)]
    kinds = [0, 3, 2, 4]                      # a real and an int fluent around two pooled ones
    A = 18
    ld = 2 * A if stochastic else A
    pooled_cols = list(range(2, 11)) + list(range(12, 18))
    heads = torch.randn(B, ld, generator=g)
    noise_fm = -torch.log(-torch.log(torch.rand(A * B, generator=g)))
    gact_fm = torch.randn(A * B, generator=g)
    blk = lambda v, f: v[segs[f][0] * B:(segs[f][0] + segs[f][1]) * B].reshape(B, segs[f][1])  # noqa: E731
    logits = heads[:, pooled_cols].clone().requires_grad_(True)
    gum = torch.cat([blk(noise_fm, 1), blk(noise_fm, 3)], 1) if stochastic else 0.0
    out = OP.gumbel_softmax_topk(logits, gum, allowed, w, 1.0)
    acts = [out[:, :9], 1.0 - out[:, 9:]]
    loss = (acts[0] * blk(gact_fm, 1)).sum() + (acts[1] * blk(gact_fm, 3)).sum()
    ent = torch.zeros(B)
    if stochastic:
        lp = torch.log_softmax(w * logits, 1)
        ent = -(torch.exp(lp) * lp).sum(1)
        loss = loss - coeff * ent.sum() / B
    loss.backward()
    actions = torch.full((A * B,), 7.0, device=dev)
    ent_dev = torch.full((B,), 0.5, device=dev)
    h_dev, n_dev = heads.to(dev), noise_fm.to(dev)
    engine.drp_topk_forward(B, A, stochastic, segs, kinds, h_dev, n_dev if stochastic else None, True,
                            allowed, w, actions, ent_dev)
    gheads = torch.full((B, ld), 9.0, device=dev)
    engine.drp_topk_backward(B, A, stochastic, segs, kinds, h_dev, n_dev if stochastic else None,
                             allowed, w, gact_fm.to(dev), gheads,
                             ent_coeff=torch.tensor([coeff], device=dev), inv_batch=1.0 / B)
    torch.cuda.synchronize()
    a_host = actions.cpu()
    for f, want in ((1, acts[0]), (3, acts[1])):
        np.testing.assert_allclose(blk(a_host, f).numpy(), want.detach().numpy(), rtol=1e-5, atol=1e-6)
    assert torch.all(blk(a_host, 0) == 7.0) and torch.all(blk(a_host, 2) == 7.0)   # not touched
    np.testing.assert_allclose(ent_dev.cpu().numpy(), 0.5 + ent.detach().numpy(), rtol=1e-5, atol=1e-5)
    gh = gheads.cpu()
    want = logits.grad.numpy()
    np.testing.assert_allclose(gh[:, pooled_cols].numpy(), want, rtol=1e-4,
                               atol=1e-4 * np.abs(want).max())
    other = [c for c in range(A) if c not in pooled_cols]
    assert torch.all(gh[:, other] == 9.0)
    if stochastic:
        assert torch.all(gh[:, [A + c for c in pooled_cols]] == 0.0)
    # test form: train = 0, typed words
    typed = torch.zeros(A * B, dtype=torch.int32, device=dev)
    engine.drp_topk_forward(B, A, stochastic, segs, kinds, h_dev, None, False, allowed, w, typed, None,
                            typed=True)
    torch.cuda.synchronize()
    hard = OP.gumbel_softmax_topk(heads[:, pooled_cols], 0.0, allowed, w, 0.0)
    t_host = typed.cpu()
    assert torch.equal(blk(t_host, 1), (hard[:, :9] > 0.5).int())
    assert torch.equal(blk(t_host, 3), ((1.0 - hard[:, 9:]) > 0.5).int())


@pytest.mark.gpu
@pytest.mark.parametrize("allowed", [1, 3])
def test_drp_topk_update_matches_oracle(cuda_device, allowed):
    """A DRP over Wildfire with a binding max-nondef-actions: rewards, loss and gradient of one
    update and the exact test rollouts against the oracle."""
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    m = _topk_model(allowed)
    T, B = 3, 21
    plan = JaxDeepReactivePolicy(topology=[16, 16], softmax_weight=2.0)
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None,
                            use_cuda_graph=False)
    pl.initialize(7)
    g = torch.Generator().manual_seed(11)
    flat = plan.initial_params(g)
    pl.params[0].copy_(flat)
    fw = pl.train_fwd.program
    noise = torch.rand(T, fw.N * B, generator=g)
    pl.train_noise.copy_(noise.reshape(-1))
    pol_noise = -torch.log(-torch.log(torch.rand(T, plan.A * B, generator=g)))
    pl.drp.pol_noise.copy_(pol_noise)
    pl.drp.ent_coeff.fill_(0.05)
    pl.update()
    torch.cuda.synchronize()
    tree = plan.params_to_pytree(flat)["params"]
    P = _with_grad(tree)
    om = OracleModel(m, pl.compiled.relax_config())
    ents = []

    def pol(t, fls):
        a, e = OP.drp_actions(m, P, {k: fls[k] for k in m.state_fluents},
                              _segment_noise(plan, m, pol_noise[t], B), 1.0, plan.bounds, 2, True,
                              softmax_weight=2.0)
        ents.append(e)
        return a
    out = OP.rollout(om, pol, B, T, [noise_as_list(fw.noise_layout, noise, t, B) for t in range(T)])
    loss = OP.mean_loss(out["reward"], float(m.discount)) - 0.05 * torch.stack(ents, 1).sum(1).mean()
    loss.backward()
    want = plan.pytree_to_params(_grads(P)).numpy()
    r_ref = out["reward"].detach().numpy().T
    np.testing.assert_allclose(pl.train_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())
    assert abs(float(pl.train_loss[0]) - float(loss.detach())) <= 2e-5 * abs(float(loss.detach()))
    got = pl.grad[0].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * np.abs(want).max())
    assert np.abs(want).max() > 0
    te = pl.test_fwd.program
    tnoise = torch.rand(T, te.N * B, generator=g)
    pl.test_noise.copy_(tnoise.reshape(-1))
    pl.test_loss()
    torch.cuda.synchronize()
    out = OP.rollout(OracleModel(m, None), lambda t, fls: OP.drp_test_actions(
        m, tree, {k: fls[k] for k in m.state_fluents}, plan.bounds, 2, softmax_weight=2.0), B, T,
        [noise_as_list(te.noise_layout, tnoise, t, B) for t in range(T)])
    np.testing.assert_allclose(pl.test_reward[0].view(T, B).cpu().numpy(), out["reward"].numpy().T,
                               rtol=1e-5, atol=1e-4)


@pytest.mark.gpu
def test_drp_sigma_entropy_grad_matches_oracle(cuda_device):
    """sigma_entropy_grad=True (planner.py:1414-1417): the entropy term reaches the log_sigma heads."""
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    m = fixture_model("powergen")
    T, B = 3, 19
    plan = JaxDeepReactivePolicy(topology=[16, 16], sigma_entropy_grad=True)
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None,
                            use_cuda_graph=False)
    pl.initialize(7)
    g = torch.Generator().manual_seed(11)
    flat = plan.initial_params(g)
    pl.params[0].copy_(flat)
    fw = pl.train_fwd.program
    noise = torch.randn(T, fw.N * B, generator=g)
    pol_noise = torch.randn(T, 5 * B, generator=g)
    pl.train_noise.copy_(noise.reshape(-1))
    pl.drp.pol_noise.copy_(pol_noise)
    pl.drp.ent_coeff.fill_(0.5)
    pl.update()
    torch.cuda.synchronize()
    P = _with_grad(plan.params_to_pytree(flat)["params"])
    om = OracleModel(m, pl.compiled.relax_config())
    ents = []

    def pol(t, fls):
        a, e = OP.drp_actions(m, P, {k: fls[k] for k in m.state_fluents},
                              {"curProd": pol_noise[t].reshape(B, 5)}, 1.0, m.bounds, 2, True,
                              sigma_entropy_grad=True)
        ents.append(e)
        return a
    out = OP.rollout(om, pol, B, T, [noise_as_list(fw.noise_layout, noise, t, B) for t in range(T)])
    loss = OP.mean_loss(out["reward"], float(m.discount)) - 0.5 * torch.stack(ents, 1).sum(1).mean()
    loss.backward()
    want = plan.pytree_to_params(_grads(P)).numpy()
    np.testing.assert_allclose(pl.grad[0].cpu().numpy(), want, rtol=1e-4,
                               atol=1e-4 * np.abs(want).max())
    ob = plan.seg["bh"][0]
    np.testing.assert_allclose(want[ob + 5:ob + 10].sum(), pl.grad[0][ob + 5:ob + 10].sum().item(),
                               rtol=1e-4)


# ---- FiLM time conditioning (planner.py:1083-1119, 1353-1363) --------------------------------------
def _film_plan(kind, topology=(16, 16), **kw):
    from pyrddlgym_jax_b200.core.drp import FixedTimeEmbedding, SinusoidalTimeEmbedding
    m = fixture_model("powergen")
    emb = {"sin": (SinusoidalTimeEmbedding, {"dim": 6}), "fixed": (FixedTimeEmbedding, {"max_t": 9, "dim": 4})}
    plan = JaxDeepReactivePolicy(topology=list(topology), time_dependent=True,
                                 time_embedding=emb[kind][0], time_embedding_kwargs=emb[kind][1], **kw)
    return m, plan, emb[kind][1]["dim"]


def test_sinusoidal_embedding_known_answer():
    """dim = 4: freqs 1, 1e-2, 1e-4, 1e-6; interleaved (sin, cos) pairs cut to 4 entries use the
    first two frequencies only (planner.py:1114-1119)."""
    from pyrddlgym_jax_b200.core.drp import SinusoidalTimeEmbedding
    tab = SinusoidalTimeEmbedding(dim=4).table(3)
    want = [[np.sin(t), np.cos(t), np.sin(t * 1e-2), np.cos(t * 1e-2)] for t in range(3)]
    np.testing.assert_allclose(tab.numpy(), np.asarray(want, np.float32), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(OP.drp_time_embedding({}, 2, "sin", 4).numpy(), tab[2].numpy(), rtol=1e-6)


@pytest.mark.parametrize("kind", ["sin", "fixed"])
def test_film_pytree_and_host_policy(kind):
    m, plan, dim = _film_plan(kind)
    plan.compile(logic.DefaultJaxRDDLCompilerWithGrad(m), JaxRDDLCompiler(m), {}, 5)
    g = torch.Generator().manual_seed(6)
    flat = plan.initial_params(g)
    tree = plan.params_to_pytree(flat)["params"]
    assert {"film_0", "film_1"} <= set(tree)
    assert ("fixed_time_embed" in tree) == (kind == "fixed")
    assert tuple(tree["film_1"]["gamma"]["kernel"].shape) == (dim, 16)
    assert float(tree["film_0"]["beta"]["bias"].abs().max()) == 0.0
    assert torch.equal(plan.pytree_to_params({"params": tree}), flat)
    fls = {"prevProd": torch.tensor([[1., 2., 0., .5, 3.]]), "temperature": torch.tensor([15.])}
    vec = np.concatenate([fls[s].numpy().reshape(-1) for s in m.state_fluents])
    for step in (0, 3):
        ref = OP.drp_test_actions(m, tree, fls, m.bounds, 2, step=step, time_embedding=kind,
                                  time_dim=dim)["curProd"].numpy().reshape(-1)
        got = plan.test_actions_host(flat, vec, step=step)["curProd"].reshape(-1)
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)
    a0 = plan.test_actions_host(flat, vec, step=0)["curProd"]
    a3 = plan.test_actions_host(flat, vec, step=3)["curProd"]
    assert not np.allclose(a0, a3)          # the policy depends on the decision epoch


@pytest.mark.gpu
@pytest.mark.parametrize("B,H", [(1000, 256), (77, 48), (4100, 33)])
def test_film_kernels_against_fp64(cuda_device, B, H):
    from pyrddlgym_jax_b200.core import engine
    dev = cuda_device
    g = torch.Generator().manual_seed(B)
    z = torch.tanh(torch.randn(B, H, generator=g))
    gamma, beta = torch.randn(H, generator=g), torch.randn(H, generator=g)
    gout = torch.randn(B, H, generator=g)
    out = torch.empty(B, H, device=dev)
    engine.film_forward(B, H, z.to(dev), gamma.to(dev), beta.to(dev), out)
    dgb0 = torch.randn(2 * H, generator=g)
    dgb = dgb0.clone().to(dev)
    gz = gout.clone().to(dev)                  # in place
    ws = torch.zeros(engine.film_chunks(B) * 2 * H, device=dev)
    engine.film_backward(B, H, gz, z.to(dev), gamma.to(dev), gz, dgb, ws)
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), (gamma * z + beta).numpy(), rtol=1e-6, atol=1e-6)
    want = (gout.double() * gamma.double() * (1 - z.double() ** 2)).float()
    np.testing.assert_allclose(gz.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
    want_gb = torch.cat([(gout.double() * z.double()).sum(0), gout.double().sum(0)]).float() + dgb0
    np.testing.assert_allclose(dgb.cpu().numpy(), want_gb.numpy(), rtol=1e-5, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,topology,B,normalize", [("sin", (16, 16), 23, False),
                                                       ("fixed", (32,), 19, False),
                                                       ("fixed", (64, 64, 32), 130, True)])
def test_drp_film_update_matches_oracle(cuda_device, kind, topology, B, normalize):
    """time_dependent policies: rewards, loss and the whole gradient (dense layers, heads, FiLM
    gamma / beta layers, the learned embedding table) of one update and the exact test rollouts
    against the oracle; B = 130 with 64-wide layers runs the tensor-core layers."""
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    m, plan, dim = _film_plan(kind, topology, normalize=normalize)
    T, L = 4, len(topology)
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None,
                            use_cuda_graph=False)
    pl.initialize(7)
    g = torch.Generator().manual_seed(11)
    flat = plan.initial_params(g)
    if kind == "fixed":
        o, r, c = plan.seg["temb"]
        flat[o:o + r * c] = torch.randn(r * c, generator=g)
    pl.params[0].copy_(flat)
    fw = pl.train_fwd.program
    noise = torch.randn(T, fw.N * B, generator=g)
    pol_noise = torch.randn(T, 5 * B, generator=g)
    pl.train_noise.copy_(noise.reshape(-1))
    pl.drp.pol_noise.copy_(pol_noise)
    pl.drp.ent_coeff.fill_(0.05)
    pl.update()
    torch.cuda.synchronize()
    tree = plan.params_to_pytree(flat)["params"]
    P = _with_grad(tree)
    om = OracleModel(m, pl.compiled.relax_config())
    ents = []

    def pol(t, fls):
        a, e = OP.drp_actions(m, P, {k: fls[k] for k in m.state_fluents},
                              {"curProd": pol_noise[t].reshape(B, 5)}, 1.0, m.bounds, L, True,
                              normalize=normalize, step=t, time_embedding=kind, time_dim=dim)
        ents.append(e)
        return a
    out = OP.rollout(om, pol, B, T, [noise_as_list(fw.noise_layout, noise, t, B) for t in range(T)])
    loss = OP.mean_loss(out["reward"], float(m.discount)) - 0.05 * torch.stack(ents, 1).sum(1).mean()
    loss.backward()
    want = plan.pytree_to_params(_grads(P)).numpy()
    r_ref = out["reward"].detach().numpy().T
    np.testing.assert_allclose(pl.train_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())
    assert abs(float(pl.train_loss[0]) - float(loss.detach())) <= 2e-5 * abs(float(loss.detach()))
    got = pl.grad[0].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * np.abs(want).max())
    o, r, c = plan.seg["fgW0"]
    assert np.abs(want[o:o + r * c]).max() > 0
    if kind == "fixed":
        o, r, c = plan.seg["temb"]
        rows = want[o:o + r * c].reshape(r, c)
        assert np.abs(rows[:T]).max() > 0 and np.all(rows[T:] == 0)
    te = pl.test_fwd.program
    tnoise = torch.randn(T, te.N * B, generator=g)
    pl.test_noise.copy_(tnoise.reshape(-1))
    pl.test_loss()
    torch.cuda.synchronize()
    out = OP.rollout(OracleModel(m, None), lambda t, fls: OP.drp_test_actions(
        m, tree, {k: fls[k] for k in m.state_fluents}, m.bounds, L, normalize=normalize, step=t,
        time_embedding=kind, time_dim=dim), B, T,
        [noise_as_list(te.noise_layout, tnoise, t, B) for t in range(T)])
    r_ref = out["reward"].numpy().T
    np.testing.assert_allclose(pl.test_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())


# ---- hidden-layer activations other than tanh (planner.py:1128, 1358-1361) -------------------------
ACTS = ["relu", "sigmoid", "softplus", "elu", "leaky_relu"]


@pytest.mark.parametrize("act", ACTS + ["gelu", "silu"])
def test_activation_host_policy_matches_oracle(act):
    m = fixture_model("powergen")
    plan = JaxDeepReactivePolicy(topology=[8, 8], activation=act)
    plan.compile(logic.DefaultJaxRDDLCompilerWithGrad(m), JaxRDDLCompiler(m), {}, 5)
    flat = plan.initial_params(torch.Generator().manual_seed(5)) * 3
    tree = plan.params_to_pytree(flat)["params"]
    fls = {"prevProd": torch.tensor([[1., 2., 0., .5, 3.]]), "temperature": torch.tensor([15.])}
    want = OP.drp_test_actions(m, tree, fls, m.bounds, 2, activation=act)["curProd"].numpy().reshape(-1)
    vec = np.concatenate([fls[s].numpy().reshape(-1) for s in m.state_fluents])
    got = plan.test_actions_host(flat, vec)["curProd"].reshape(-1)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
    with pytest.raises(NotImplementedError):
        JaxDeepReactivePolicy(activation="hard_tanh")


@pytest.mark.gpu
@pytest.mark.parametrize("act", ACTS)
def test_activation_epilogues_against_torch(cuda_device, act):
    """Every kernel with a fused activation, under each activation: SIMT and tensor-core layer
    (bias + activation, activation adjoint from the stored output), the first layer, the fused head
    adjoint and the FiLM adjoint."""
    from pyrddlgym_jax_b200.core import engine
    dev = cuda_device
    f = OP.DRP_ACTIVATIONS[act]
    g = torch.Generator().manual_seed(3)
    engine.drp_set_activation(act)

    def dact(pre):      # derivative at the pre-activation, fp64
        p = pre.double().clone().requires_grad_(True)
        OP.DRP_ACTIVATIONS[act](p).sum().backward()
        return p.grad
    M, N, K = 200, 64, 48
    A = torch.randn(M, K, generator=g) / np.sqrt(K) * 2
    W = torch.randn(K, N, generator=g)
    bias = torch.randn(N, generator=g)
    pre = A.double() @ W.double() + bias.double()
    Y = torch.empty(M, N, device=dev)
    engine.sgemm(engine.EPI_BIAS_TANH, M, N, K, A.to(dev), K, W.to(dev), N, Y, N, bias=bias.to(dev))
    torch.cuda.synchronize()
    np.testing.assert_allclose(Y.cpu().numpy(), f(pre).float().numpy(), rtol=2e-5, atol=2e-5)
    # adjoint epilogue: dX = (dZ W2^T) * act'(stored output)
    dZ = torch.randn(M, K, generator=g)
    W2T = torch.randn(K, N, generator=g)
    dX = torch.empty(M, N, device=dev)
    engine.sgemm(engine.EPI_DTANH, M, N, K, dZ.to(dev), K, W2T.to(dev), N, dX, N, aux=Y, ldaux=N)
    torch.cuda.synchronize()
    want = (dZ.double() @ W2T.double()) * dact(pre)
    keep = (pre.abs() > 1e-3).numpy()            # away from the kinks of relu / leaky_relu / elu
    np.testing.assert_allclose(dX.cpu().numpy()[keep], want.float().numpy()[keep], rtol=2e-4, atol=2e-4)
    # tensor-core layer
    M2, N2, K2 = 384, 64, 64
    A2 = torch.randn(M2, K2, generator=g) / 4
    Wt = torch.randn(N2, K2, generator=g)                  # [out][in]
    b2 = torch.randn(N2, generator=g)
    hi, lo = torch.empty(N2 * K2, device=dev), torch.empty(N2 * K2, device=dev)
    engine.tf32_split(Wt.to(dev).view(-1), hi, lo, N2 * K2)
    Y2 = torch.empty(M2, N2, device=dev)
    engine.tcgemm(2, M2, N2, K2, A2.to(dev), K2, hi, lo, K2, Y2, N2, bias=b2.to(dev))
    torch.cuda.synchronize()
    pre2 = A2.double() @ Wt.double().t() + b2.double()
    np.testing.assert_allclose(Y2.cpu().numpy(), f(pre2).float().numpy(), rtol=2e-5, atol=2e-5)
    dZ2 = torch.randn(M2, K2, generator=g)
    dX2 = torch.empty(M2, N2, device=dev)
    engine.tcgemm(3, M2, N2, K2, dZ2.to(dev), K2, hi, lo, K2, dX2, N2, aux=Y2, ldaux=N2)
    torch.cuda.synchronize()
    want2 = (dZ2.double() @ Wt.double().t()) * dact(pre2)
    keep2 = (pre2.abs() > 1e-3).numpy()
    np.testing.assert_allclose(dX2.cpu().numpy()[keep2], want2.float().numpy()[keep2], rtol=2e-4,
                               atol=2e-4)
    # first layer from fluent-major blocks (both kernels: D < 16 and D >= 16)
    for segs in ([(0, 5), (5, 1)], [(0, 12), (12, 9)]):
        B, H = 333, 96
        D = sum(n for _, n in segs)
        blocks = [torch.randn(B, n, generator=g) for _, n in segs]
        X = torch.cat(blocks, 1)
        W0, b0 = torch.randn(D, H, generator=g) / 3, torch.randn(H, generator=g)
        out = torch.empty(B, H, device=dev)
        engine.drp_layer0(B, D, H, segs, torch.cat([b.reshape(-1) for b in blocks]).to(dev),
                          W0.to(dev), b0.to(dev), out, None)
        torch.cuda.synchronize()
        np.testing.assert_allclose(out.cpu().numpy(), f(X.double() @ W0.double() + b0.double()).float().numpy(),
                                   rtol=2e-5, atol=2e-5)
    # fused head adjoint: gz = (G Wh^T) * act'(Hl)
    B, hl, NH = 500, 64, 6
    preH = torch.randn(B, hl, generator=g)
    Hl = f(preH)
    G = torch.randn(B, NH, generator=g)
    Wh = torch.randn(hl, NH, generator=g)
    gz = torch.empty(B, hl, device=dev)
    tail = torch.zeros(hl + hl * NH + NH, device=dev)
    ws = torch.zeros(engine.drp_head_backward_ctas(B) * (hl + hl * NH + NH), device=dev)
    engine.drp_head_backward(B, hl, NH, Hl.to(dev), G.to(dev), Wh.to(dev), gz, tail, ws)
    torch.cuda.synchronize()
    want = (G.double() @ Wh.double().t()) * dact(preH)
    keepH = (preH.abs() > 1e-3).numpy()
    np.testing.assert_allclose(gz.cpu().numpy()[keepH], want.float().numpy()[keepH], rtol=2e-4, atol=2e-4)
    # FiLM adjoint
    gamma = torch.randn(hl, generator=g)
    gout = torch.randn(B, hl, generator=g)
    gzf = gout.clone().to(dev)
    dgb = torch.zeros(2 * hl, device=dev)
    engine.film_backward(B, hl, gzf, Hl.to(dev), gamma.to(dev), gzf, dgb,
                         torch.zeros(engine.film_chunks(B) * 2 * hl, device=dev))
    torch.cuda.synchronize()
    want = gout.double() * gamma.double() * dact(preH)
    np.testing.assert_allclose(gzf.cpu().numpy()[keepH], want.float().numpy()[keepH], rtol=2e-4, atol=2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("act,topology,B,film", [("relu", (16, 16), 23, False),
                                                 ("softplus", (64, 64), 130, False),
                                                 ("elu", (32, 16), 19, True),
                                                 ("sigmoid", (64, 64, 32), 129, False),
                                                 # gelu / silu: pre-activation tape, layer-wise path
                                                 ("gelu", (16, 16), 23, False),
                                                 ("gelu", (64, 64), 130, False),
                                                 ("silu", (64, 32, 32), 129, False),
                                                 ("silu", (256, 256), 256, False)])
def test_drp_activation_update_matches_oracle(cuda_device, act, topology, B, film):
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    m = fixture_model("powergen")
    T, L = 4, len(topology)
    plan = JaxDeepReactivePolicy(topology=list(topology), activation=act, time_dependent=film)
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None,
                            use_cuda_graph=False)
    pl.initialize(7)
    g = torch.Generator().manual_seed(11)
    flat = plan.initial_params(g)
    pl.params[0].copy_(flat)
    fw = pl.train_fwd.program
    noise = torch.randn(T, fw.N * B, generator=g)
    pol_noise = torch.randn(T, 5 * B, generator=g)
    pl.train_noise.copy_(noise.reshape(-1))
    pl.drp.pol_noise.copy_(pol_noise)
    pl.drp.ent_coeff.fill_(0.05)
    pl.update()
    torch.cuda.synchronize()
    tree = plan.params_to_pytree(flat)["params"]
    P = _with_grad(tree)
    om = OracleModel(m, pl.compiled.relax_config())
    ents = []
    tkw = dict(time_embedding="sin", time_dim=8) if film else {}

    def pol(t, fls):
        a, e = OP.drp_actions(m, P, {k: fls[k] for k in m.state_fluents},
                              {"curProd": pol_noise[t].reshape(B, 5)}, 1.0, m.bounds, L, True,
                              activation=act, step=t, **tkw)
        ents.append(e)
        return a
    out = OP.rollout(om, pol, B, T, [noise_as_list(fw.noise_layout, noise, t, B) for t in range(T)])
    loss = OP.mean_loss(out["reward"], float(m.discount)) - 0.05 * torch.stack(ents, 1).sum(1).mean()
    loss.backward()
    want = plan.pytree_to_params(_grads(P)).numpy()
    r_ref = out["reward"].detach().numpy().T
    np.testing.assert_allclose(pl.train_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())
    assert abs(float(pl.train_loss[0]) - float(loss.detach())) <= 2e-5 * abs(float(loss.detach()))
    got = pl.grad[0].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * np.abs(want).max())
    te = pl.test_fwd.program
    tnoise = torch.randn(T, te.N * B, generator=g)
    pl.test_noise.copy_(tnoise.reshape(-1))
    pl.test_loss()
    torch.cuda.synchronize()
    out = OP.rollout(OracleModel(m, None), lambda t, fls: OP.drp_test_actions(
        m, tree, {k: fls[k] for k in m.state_fluents}, m.bounds, L, activation=act, step=t, **tkw),
        B, T, [noise_as_list(te.noise_layout, tnoise, t, B) for t in range(T)])
    r_ref = out["reward"].numpy().T
    np.testing.assert_allclose(pl.test_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())


# ---- StaticNormalizer preprocessor (planner.py:284-338, 1293-1296) ---------------------------------
INVARIANT_RDDL = """
domain bounds_toy {
    types { tank : object; };
    pvariables {
        CAP(tank) : { non-fluent, real, default = 10.0 };
        level(tank) : { state-fluent, real, default = 1.0 };
        heat : { state-fluent, real, default = 0.0 };
        free : { state-fluent, real, default = 0.0 };
        pump(tank) : { action-fluent, real, default = 0.0 };
    };
    cpfs {
        level'(?t) = level(?t) + pump(?t);
        heat' = heat;
        free' = free;
    };
    reward = sum_{?t : tank} [ level(?t) ];
    state-invariants {
        forall_{?t : tank} [ level(?t) >= 0.5 ];
        forall_{?t : tank} [ level(?t) <= CAP(?t) ];
        heat >= -2.0;
        3.0 >= heat;
        free >= 0.0;
    };
    action-preconditions {
        forall_{?t : tank} [ pump(?t) >= 0.0 ];
        forall_{?t : tank} [ pump(?t) <= 1.0 ];
    };
}
non-fluents nf { domain = bounds_toy; objects { tank : {t1, t2}; };
    non-fluents { CAP(t2) = 20.0; }; }
instance inst { domain = bounds_toy; non-fluents = nf; max-nondef-actions = pos-inf;
    horizon = 3; discount = 1.0; }
"""


def test_static_normalizer_reads_state_invariants():
    from pyrddlgym_jax_b200.core.planner import StaticNormalizer
    from pyrddlgym_jax_b200.rddl.model import load_model
    m = load_model(INVARIANT_RDDL)
    np.testing.assert_array_equal(m.state_bounds["level"][0], [0.5, 0.5])
    np.testing.assert_array_equal(m.state_bounds["level"][1], [10.0, 20.0])
    assert m.state_bounds["heat"] == (np.float32(-2.0), np.float32(3.0)) or \
        (float(m.state_bounds["heat"][0]) == -2.0 and float(m.state_bounds["heat"][1]) == 3.0)
    pre = StaticNormalizer(fluent_bounds={"free": (1.0, 5.0)})
    pre.compile(logic.DefaultJaxRDDLCompilerWithGrad(m))
    stats = pre.initialize()
    assert set(stats) == {"level", "heat", "free"}      # 'free' only through the user's bounds
    out = pre.transform({"level": torch.tensor([[5.25, 10.25]]), "heat": torch.tensor([0.5])}, stats)
    np.testing.assert_allclose(out["level"].numpy(), [[0.5, 0.5]], rtol=1e-6)
    np.testing.assert_allclose(out["heat"].numpy(), [0.5], rtol=1e-6)
    # a half-open box is not used
    pre2 = StaticNormalizer()
    pre2.compile(logic.DefaultJaxRDDLCompilerWithGrad(m))
    assert set(pre2.initialize()) == {"level", "heat"}


POWERGEN_BOUNDS = {"prevProd": (np.zeros(5, np.float32), np.array([5., 6., 7., 8., 10.], np.float32)),
                   "temperature": (-30.0, 40.0)}


def test_preprocessor_host_policy_matches_oracle():
    from pyrddlgym_jax_b200.core.planner import StaticNormalizer
    m = fixture_model("powergen")
    plan = JaxDeepReactivePolicy(topology=[8, 8], normalize=True)
    pre = StaticNormalizer(fluent_bounds=POWERGEN_BOUNDS)
    plan.compile(logic.DefaultJaxRDDLCompilerWithGrad(m), JaxRDDLCompiler(m), {}, 5, preprocessor=pre)
    flat = plan.initial_params(torch.Generator().manual_seed(5)) * 3
    tree = plan.params_to_pytree(flat)["params"]
    fls = {"prevProd": torch.tensor([[1., 2., 0., .5, 3.]]), "temperature": torch.tensor([15.])}
    want = OP.drp_test_actions(m, tree, fls, m.bounds, 2, normalize=True,
                               preprocess=pre.initialize())["curProd"].numpy().reshape(-1)
    vec = np.concatenate([fls[s].numpy().reshape(-1) for s in m.state_fluents])
    got = plan.test_actions_host(flat, vec)["curProd"].reshape(-1)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("normalize", [False, True])
def test_drp_preprocessor_update_matches_oracle(cuda_device, normalize):
    """preprocessor=StaticNormalizer: min-max scaled inputs (before the optional LayerNorm) in the
    train and test policies; rewards, loss, gradient and test rollouts against the oracle."""
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner, StaticNormalizer
    m = fixture_model("powergen")
    T, B = 4, 27
    plan = JaxDeepReactivePolicy(topology=[16, 16], normalize=normalize)
    pre = StaticNormalizer(fluent_bounds=POWERGEN_BOUNDS)
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None, preprocessor=pre,
                            use_cuda_graph=False)
    pl.initialize(7)
    g = torch.Generator().manual_seed(11)
    flat = plan.initial_params(g)
    pl.params[0].copy_(flat)
    fw = pl.train_fwd.program
    noise = torch.randn(T, fw.N * B, generator=g)
    pol_noise = torch.randn(T, 5 * B, generator=g)
    pl.train_noise.copy_(noise.reshape(-1))
    pl.drp.pol_noise.copy_(pol_noise)
    pl.drp.ent_coeff.fill_(0.05)
    pl.update()
    torch.cuda.synchronize()
    stats = pre.initialize()
    tree = plan.params_to_pytree(flat)["params"]
    P = _with_grad(tree)
    om = OracleModel(m, pl.compiled.relax_config())
    ents = []

    def pol(t, fls):
        a, e = OP.drp_actions(m, P, {k: fls[k] for k in m.state_fluents},
                              {"curProd": pol_noise[t].reshape(B, 5)}, 1.0, m.bounds, 2, True,
                              normalize=normalize, preprocess=stats)
        ents.append(e)
        return a
    out = OP.rollout(om, pol, B, T, [noise_as_list(fw.noise_layout, noise, t, B) for t in range(T)])
    loss = OP.mean_loss(out["reward"], float(m.discount)) - 0.05 * torch.stack(ents, 1).sum(1).mean()
    loss.backward()
    want = plan.pytree_to_params(_grads(P)).numpy()
    r_ref = out["reward"].detach().numpy().T
    np.testing.assert_allclose(pl.train_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())
    assert abs(float(pl.train_loss[0]) - float(loss.detach())) <= 2e-5 * abs(float(loss.detach()))
    np.testing.assert_allclose(pl.grad[0].cpu().numpy(), want, rtol=1e-4,
                               atol=1e-4 * np.abs(want).max())
    te = pl.test_fwd.program
    tnoise = torch.randn(T, te.N * B, generator=g)
    pl.test_noise.copy_(tnoise.reshape(-1))
    pl.test_loss()
    torch.cuda.synchronize()
    out = OP.rollout(OracleModel(m, None), lambda t, fls: OP.drp_test_actions(
        m, tree, {k: fls[k] for k in m.state_fluents}, m.bounds, 2, normalize=normalize,
        preprocess=stats), B, T, [noise_as_list(te.noise_layout, tnoise, t, B) for t in range(T)])
    r_ref = out["reward"].numpy().T
    np.testing.assert_allclose(pl.test_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("utility,kw", [("mean_var", {"beta": 0.8}), ("cvar", {"alpha": 0.3})])
def test_drp_utility_update_matches_oracle(cuda_device, utility, kw):
    """A utility other than the mean with a deep reactive policy (planner.py:2817-2818 applies it to the
    returns of either plan kind): train loss, gradient and test loss against the oracle, whose loss is
    -utility(returns) - ent_coeff * mean entropy through torch autograd."""
    from pyrddlgym_jax_b200.core import planner as PL
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    m = fixture_model("powergen")
    T, B = 4, 27
    plan = JaxDeepReactivePolicy(topology=[16, 16])
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None, utility=utility,
                            utility_kwargs=kw, use_cuda_graph=False)
    assert pl._native_utility is not None
    pl.initialize(7)
    g = torch.Generator().manual_seed(13)
    flat = plan.initial_params(g)
    pl.params[0].copy_(flat)
    fw = pl.train_fwd.program
    noise = torch.randn(T, fw.N * B, generator=g)
    pol_noise = torch.randn(T, 5 * B, generator=g)
    pl.train_noise.copy_(noise.reshape(-1))
    pl.drp.pol_noise.copy_(pol_noise)
    pl.drp.ent_coeff.fill_(0.05)
    pl.update()
    torch.cuda.synchronize()
    tree = plan.params_to_pytree(flat)["params"]
    P = _with_grad(tree)
    om = OracleModel(m, pl.compiled.relax_config())
    ents = []

    def pol(t, fls):
        a, e = OP.drp_actions(m, P, {k: fls[k] for k in m.state_fluents},
                              {"curProd": pol_noise[t].reshape(B, 5)}, 1.0, m.bounds, 2, True)
        ents.append(e)
        return a
    out = OP.rollout(om, pol, B, T, [noise_as_list(fw.noise_layout, noise, t, B) for t in range(T)])
    R = OP.returns_from_rewards(out["reward"], float(m.discount))
    fn = PL.UTILITY_LOOKUP[utility]
    loss = -fn(R, **kw) - 0.05 * torch.stack(ents, 1).sum(1).mean()
    loss.backward()
    want = plan.pytree_to_params(_grads(P)).numpy()
    assert abs(float(pl.train_loss[0]) - float(loss.detach())) <= 2e-5 * abs(float(loss.detach()))
    np.testing.assert_allclose(pl.grad[0].cpu().numpy(), want, rtol=1e-4,
                               atol=1e-4 * np.abs(want).max())
    pl.test_loss()
    torch.cuda.synchronize()
    want_test = -fn(pl.test_returns[0].cpu(), **kw)
    assert abs(float(pl.test_loss_buf[0]) - float(want_test)) <= 2e-5 * abs(float(want_test))


def test_observation_view_and_host_policy():
    """Partially observed models (planner.py:1300-1313): the carried copies of the observ-fluents are the
    only inputs of the policy -- parameter pytree, initialiser fan-in and host test policy ignore the
    hidden state words."""
    from pyrddlgym_jax_b200.rddl.model import observation_view
    from pyrddlgym_jax_b200.core.compiler import JaxRDDLCompiler
    base = fixture_model("pomdp")
    m = observation_view(base)
    assert list(m.state_fluents) == ["level", "sensor@seen", "alarm@seen"]
    assert m.next_state["sensor@seen"] == "sensor" and m.observed_inputs == ["sensor@seen", "alarm@seen"]
    assert observation_view(fixture_model("reservoir")).state_fluents.keys() == \
        fixture_model("reservoir").state_fluents.keys()
    plan = JaxDeepReactivePolicy(topology=[8, 8])
    with pytest.raises(ValueError):
        plan.compile(JaxRDDLCompiler(base), JaxRDDLCompiler(base), {}, 4)
    plan.compile(JaxRDDLCompiler(m), JaxRDDLCompiler(m), {}, 4)
    assert plan.D == 9 and plan.D_in == 6 and plan.frozen_segs == [(0, 3)]
    g = torch.Generator().manual_seed(3)
    flat = plan.initial_params(g)
    off, r, c = plan.seg["W0"]
    W0 = flat[off:off + r * c].reshape(r, c)
    assert float(W0[:3].abs().max()) == 0.0 and float(W0[3:].abs().min()) > 0.0
    assert abs(float(W0[3:].std()) - math.sqrt(2.0 / 6)) < 0.25          # fan-in 6, not 9
    tree = plan.params_to_pytree(flat)["params"]
    assert tuple(tree["dense_0"]["kernel"].shape) == (6, 8)             # non-bool (sensor) rows, then alarm
    assert torch.equal(plan.pytree_to_params({"params": tree}), flat)
    # the host policy does not react to the hidden level words
    obs = np.array([0.4, 0.8, 0.1, 1.0, 0.0, 1.0], np.float32)
    a = plan.test_actions_host(flat, np.concatenate([np.zeros(3, np.float32), obs]))
    b = plan.test_actions_host(flat, np.concatenate([np.full(3, 7.0, np.float32), obs]))
    np.testing.assert_array_equal(a["pump"], b["pump"])


@pytest.mark.gpu
@pytest.mark.parametrize("topology,normalize", [((16, 16), False), ((8, 8), True)])
def test_drp_pomdp_update_matches_oracle(cuda_device, topology, normalize):
    """A deep reactive policy over observ-fluents: rewards, loss and parameter gradient of one update and
    the exact test rollouts against the oracle, whose policy sees the sensors of the previous step (the
    initial defaults at step 0) and never the hidden levels."""
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    base = fixture_model("pomdp")
    T, B = 5, 21
    plan = JaxDeepReactivePolicy(topology=list(topology), normalize=normalize)
    pl = JaxBackpropPlanner(base, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None, use_cuda_graph=False)
    m = pl.rddl                                        # the observation view
    assert m is not base and plan.frozen_segs
    pl.initialize(7)
    g = torch.Generator().manual_seed(17)
    flat = plan.initial_params(g)
    pl.params[0].copy_(flat)
    fw = pl.train_fwd.program
    noise = torch.randn(T, fw.N * B, generator=g)
    pol_noise = torch.randn(T, 3 * B, generator=g)
    pl.train_noise.copy_(noise.reshape(-1))
    pl.drp.pol_noise.copy_(pol_noise)
    pl.drp.ent_coeff.fill_(0.05)
    pl.update()
    torch.cuda.synchronize()
    tree = plan.params_to_pytree(flat)["params"]
    P = _with_grad(tree)
    om = OracleModel(m, pl.compiled.relax_config())
    ents = []
    seen = ("sensor@seen", "alarm@seen")

    def pol(t, fls):
        a, e = OP.drp_actions(m, P, {k: fls[k] for k in seen},
                              {"pump": pol_noise[t].reshape(B, 3)}, 1.0, m.bounds, 2, True,
                              normalize=normalize)
        ents.append(e)
        return a
    out = OP.rollout(om, pol, B, T, [noise_as_list(fw.noise_layout, noise, t, B) for t in range(T)])
    loss = OP.mean_loss(out["reward"], float(m.discount)) - 0.05 * torch.stack(ents, 1).sum(1).mean()
    loss.backward()
    want = plan.pytree_to_params(_grads(P)).numpy()
    r_ref = out["reward"].detach().numpy().T
    np.testing.assert_allclose(pl.train_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())
    assert abs(float(pl.train_loss[0]) - float(loss.detach())) <= 2e-5 * abs(float(loss.detach()))
    got = pl.grad[0].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * np.abs(want).max())
    off, r, c = plan.seg["W0"]
    assert float(np.abs(got[off:off + 3 * c]).max()) == 0.0          # hidden rows: no gradient
    assert float(np.abs(want).max()) > 0
    te = pl.test_fwd.program
    tnoise = torch.randn(T, te.N * B, generator=g)
    pl.test_noise.copy_(tnoise.reshape(-1))
    pl.test_loss()
    torch.cuda.synchronize()
    ref = OP.rollout(OracleModel(m, None), lambda t, fls: OP.drp_test_actions(
        m, tree, {k: fls[k] for k in seen}, m.bounds, 2, normalize=normalize), B, T,
        [noise_as_list(te.noise_layout, tnoise, t, B) for t in range(T)])
    r_ref = ref["reward"].detach().numpy().T
    np.testing.assert_allclose(pl.test_reward[0].view(T, B).cpu().numpy(), r_ref, rtol=1e-5,
                               atol=1e-5 * np.abs(r_ref).max())


@pytest.mark.gpu
def test_drp_pomdp_optimize_and_act(cuda_device):
    """End to end on the partially observed fixture: a few planner iterations (CUDA graph), returns
    improve, and get_action takes the observation dict a POMDP environment hands out."""
    from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
    base = fixture_model("pomdp")
    plan = JaxDeepReactivePolicy(topology=[16, 16])
    pl = JaxBackpropPlanner(base, plan, batch_size_train=256, optimizer="rmsprop",
                            optimizer_kwargs={"learning_rate": 0.01}, pgpe=None)
    first = last = None
    for cb in pl.optimize_generator(key=5, epochs=40, print_summary=False, print_progress=False):
        first = cb if first is None else first
        last = cb
        assert np.isfinite(cb["train_return"]).all() and np.isfinite(cb["test_return"]).all()
    assert last["best_return"] > first["best_return"]
    obs = {"sensor": np.array([0.2, 0.7, 0.4], np.float32), "alarm": np.array([True, False, False])}
    act = pl.get_action(None, last["best_params"], 0, obs)
    assert set(act) == {"pump"} and act["pump"].shape == (3,)
    assert (act["pump"] >= 0.0).all() and (act["pump"] <= 1.0).all()
    tree = last["best_params"]["params"]
    assert tuple(np.asarray(tree["dense_0"]["kernel"]).shape) == (6, 16)      # sensors + alarms only
