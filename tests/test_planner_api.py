"""Drop-in surface: the .cfg format (planner.py:101-244), the planner / controller API and the
callback contract (planner.py:3368-3392) -- CPU tests for the host logic, GPU tests end to end."""
import numpy as np
import pytest
import torch

from pyrddlgym_jax_b200.core import logic
from pyrddlgym_jax_b200.core.planner import (JaxBackpropPlanner, JaxDeepReactivePolicy,
                                             JaxOfflineController, JaxOnlineController,
                                             JaxPlannerStatus, JaxStraightLinePlan,
                                             NoImprovementStoppingRule, load_config_from_string)

from .helpers import fixture_model

# the text of three configs shipped by the reference (examples/configs/*.cfg)
QUADCOPTER_SLP = """
[Compiler]
method='DefaultJaxRDDLCompilerWithGrad'
sigmoid_weight=10

[Planner]
method='JaxStraightLinePlan'
method_kwargs={}
optimizer='rmsprop'
optimizer_kwargs={'learning_rate': 0.03}
batch_size_train=1
batch_size_test=1
pgpe=None

[Optimize]
key=42
epochs=100000
train_seconds=360
"""
HVAC_DRP = """
[Compiler]
method='DefaultJaxRDDLCompilerWithGrad'
bernoulli_sigmoid_weight=5
sigmoid_weight=5

[Planner]
method='JaxDeepReactivePolicy'
method_kwargs={'topology': [64, 64]}
optimizer='rmsprop'
optimizer_kwargs={'learning_rate': 0.001}
batch_size_train=1
batch_size_test=1
pgpe=None

[Optimize]
key=42
epochs=6000
train_seconds=90
"""
WILDFIRE_SLP = """
[Compiler]
method='DefaultJaxRDDLCompilerWithGrad'
sigmoid_weight=100
bernoulli_sigmoid_weight=100

[Planner]
method='JaxStraightLinePlan'
method_kwargs={'sigmoid_weight': 10.0}
optimizer='rmsprop'
optimizer_kwargs={'learning_rate': 0.01}
batch_size_train=32
batch_size_test=32
pgpe=None

[Optimize]
key=42
epochs=1000
train_seconds=30
stopping_rule='NoImprovementStoppingRule'
stopping_rule_kwargs={'patience': 7}
"""


def test_load_config_resolves_names_like_the_reference():
    planner_args, plan_kwargs, train_args = load_config_from_string(QUADCOPTER_SLP)
    assert planner_args["compiler"] is logic.DefaultJaxRDDLCompilerWithGrad
    assert planner_args["compiler_kwargs"] == {"sigmoid_weight": 10}
    assert isinstance(planner_args["plan"], JaxStraightLinePlan)
    assert planner_args["optimizer"] == "rmsprop"
    assert planner_args["optimizer_kwargs"] == {"learning_rate": 0.03}
    assert planner_args["batch_size_train"] == 1 and planner_args["pgpe"] is None
    assert train_args == {"key": 42, "epochs": 100000, "train_seconds": 360}
    planner_args, plan_kwargs, _ = load_config_from_string(HVAC_DRP)
    assert isinstance(planner_args["plan"], JaxDeepReactivePolicy)
    assert plan_kwargs == {"topology": [64, 64]} and planner_args["plan"]._topology == [64, 64]
    planner_args, plan_kwargs, train_args = load_config_from_string(WILDFIRE_SLP)
    assert planner_args["plan"]._sigmoid_weight == 10.0
    assert isinstance(train_args["stopping_rule"], NoImprovementStoppingRule)
    assert train_args["stopping_rule"].patience == 7


def test_load_config_rejects_unknown_names():
    with pytest.raises(ValueError):
        load_config_from_string(QUADCOPTER_SLP.replace("DefaultJaxRDDLCompilerWithGrad", "Nope"))
    with pytest.raises(ValueError):
        load_config_from_string(QUADCOPTER_SLP.replace("JaxStraightLinePlan", "NoSuchPlan"))
    with pytest.raises(ValueError):
        load_config_from_string(QUADCOPTER_SLP.replace("'rmsprop'", "'lion'"))


def test_compiler_classes_expose_reference_kwargs():
    m = fixture_model("reservoir")
    c = logic.DefaultJaxRDDLCompilerWithGrad(m, sigmoid_weight=7., use_if_else_ste=False)
    kw = c.get_kwargs()
    for key in ("sigmoid_weight", "use_sigmoid_ste", "use_tanh_ste", "argmax_weight", "use_logic_ste",
                "floor_weight", "round_weight", "use_if_else_ste", "switch_weight",
                "bernoulli_sigmoid_weight", "use_ste_bernoulli", "discrete_softmax_weight",
                "binomial_nbins", "poisson_nbins"):
        assert key in kw, key
    assert kw["sigmoid_weight"] == 7. and kw["use_if_else_ste"] is False


def test_status_machine_matches_reference():
    """planner.py:2277-2304."""
    assert [s.value for s in JaxPlannerStatus] == [0, 1, 2, 3, 4, 5, 6]
    terminal = {s.name for s in JaxPlannerStatus if s.is_terminal()}
    assert terminal == {"STOPPING_RULE_REACHED", "INVALID_GRADIENT", "TIME_BUDGET_REACHED",
                        "ITER_BUDGET_REACHED"}


CALLBACK_KEYS = {"iteration", "elapsed_time", "progress", "status", "key", "test_key", "train_return",
                 "test_return", "best_return", "pgpe_return", "last_iteration_improved",
                 "pgpe_improved", "params", "best_params", "best_index", "pgpe_params",
                 "model_params", "policy_hyperparams", "grad", "best_grad", "train_log", "test_log",
                 "planner_state"}


@pytest.mark.gpu
def test_optimize_generator_callback_contract(cuda_device):
    planner_args, _, train_args = load_config_from_string(QUADCOPTER_SLP)
    planner_args.update(batch_size_train=64, batch_size_test=64, rollout_horizon=20)
    m = fixture_model("quadcopter")
    pl = JaxBackpropPlanner(m, **planner_args)
    train_args.update(epochs=25, print_summary=False, print_progress=False)
    returns, last = [], None
    for cb in pl.optimize_generator(**train_args):
        assert CALLBACK_KEYS <= set(cb), CALLBACK_KEYS - set(cb)
        returns.append(float(cb["train_return"][0]))
        last = cb
    assert last["iteration"] == 24 and last["status"] == JaxPlannerStatus.ITER_BUDGET_REACHED
    assert returns[-1] > returns[0]                       # the plan improves
    assert last["best_return"] >= -1e30 and np.isfinite(last["best_return"])
    mu = last["best_params"]["real"]["thrust"]["mu"]
    assert tuple(mu.shape) == (20, 4)
    assert tuple(last["train_log"]["reward"].shape) == (1, 64, 20)
    assert tuple(last["test_log"]["reward"].shape) == (1, 64, 20)
    # actions from the returned plan respect the action box (test policy, planner.py:842-868)
    act = pl.get_action(None, last["best_params"], 3, None)
    assert set(act) == set(m.action_fluents) and np.all(act["thrust"] >= 0) and np.all(act["thrust"] <= 20)


@pytest.mark.gpu
def test_offline_and_online_controllers(cuda_device):
    m = fixture_model("reservoir")
    pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=32, rollout_horizon=8,
                            optimizer_kwargs={"learning_rate": 0.2}, pgpe=None)
    off = JaxOfflineController(pl, key=42, epochs=10, print_summary=False, print_progress=False)
    state = {"rlevel": np.array([50., 60., 40.], np.float32)}
    a0 = off.sample_action(state)
    a1 = off.sample_action(state)
    assert set(a0) == {"release"} and a0["release"].shape == (3,) and off.step == 2
    off.reset()
    assert off.step == 0
    on = JaxOnlineController(pl, key=3, epochs=5, print_summary=False, print_progress=False)
    b0 = on.sample_action(state)
    assert on.guess is not None and b0["release"].shape == (3,)       # warm start for the next epoch
    b1 = on.sample_action({"rlevel": np.array([55., 58., 42.], np.float32)})
    assert np.all(b1["release"] >= 0)


@pytest.mark.gpu
def test_wildfire_plan_respects_max_nondef_actions(cuda_device):
    """After every update the SOGBOFA projection leaves at most max-nondef-actions (= 1) boolean
    actions on under the test policy (planner.py:914-959, 856)."""
    planner_args, _, train_args = load_config_from_string(WILDFIRE_SLP)
    planner_args.update(batch_size_train=64, batch_size_test=64, rollout_horizon=10)
    m = fixture_model("wildfire")
    pl = JaxBackpropPlanner(m, **planner_args)
    train_args.pop("stopping_rule")
    train_args.update(epochs=8, print_summary=False, print_progress=False)
    for cb in pl.optimize_generator(**train_args):
        # callback['params'] are device tensors (views of the live plan), [P, T, *objs]
        flat = torch.cat([v["logit"][0].cpu().reshape(10, -1)
                          for v in cb["params"]["bool"].values()], 1)
        assert int((flat > 0).sum(1).max()) <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["quadcopter", "wildfire"])
def test_pipelined_graph_equals_sequential_iterations(cuda_device, key, monkeypatch):
    """The captured, software-pipelined iteration (opt -> test || next gradient) produces the
    same losses and plan, iteration by iteration, as plain update-then-test launches."""
    m = fixture_model(key)
    monkeypatch.setenv("RK_NOISE_IN_GRAPH", "0")        # fixed noise in both runs

    def run(graph):
        pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=96, rollout_horizon=12,
                                optimizer_kwargs={"learning_rate": 0.05}, pgpe=None,
                                use_cuda_graph=graph)
        pl.initialize(5)
        out = []
        for _ in range(6):
            pl.iterate()                                 # fixed noise: initialise() drew it once
            tr, te, _ = pl.read_stats()
            out.append((float(tr[0]), float(te[0]), pl.params[0].cpu().clone(),
                        pl.grad_used[0].cpu().clone()))
        return out
    a, b = run(True), run(False)
    for (tr1, te1, p1, g1), (tr2, te2, p2, g2) in zip(a, b):
        assert tr1 == tr2 and te1 == te2
        assert torch.equal(p1, p2) and torch.equal(g1, g2)


def test_pgpe_formulas_known_answers():
    """Super-symmetric sample, mean / sigma gradient estimates of GaussianPGPE against an
    independent numpy restatement of planner.py:1888-1916, 1954-2003."""
    from pyrddlgym_jax_b200.core.pgpe import GaussianPGPE
    pg = GaussianPGPE()
    sigma = torch.tensor([1.0, 0.5, 2.0, 1.0, 0.3])
    eps = torch.tensor([0.3, -0.5, 3.5, 0.0, 0.3])             # |eps| <, =, > sigma, zero
    got = pg.eps_star(sigma, eps).numpy()
    s, e = sigma.numpy().astype(np.float64), eps.numpy().astype(np.float64)
    a = (s - np.abs(e)) / s
    aa = np.abs(a)
    c1, c2, c3 = -0.06655, -0.9706, 0.124
    with np.errstate(all="ignore"):
        neg = np.exp(np.where(aa == 1., 2 * c1 + c2, c1 * aa * (aa * aa - 1) / np.log(aa) + c2 * aa))
        pos = np.exp(aa) / np.power(np.maximum(1 - aa ** 3, np.finfo(np.float32).tiny), c3 * aa)
    want = np.where(e == 0, 0., np.sign(e) * 0.67449 * s * np.where(a < 0, neg, pos))
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=1e-7)
    r = torch.tensor([-10., -14., -11., -12.])
    m = torch.tensor(-9.)
    es = torch.tensor(want, dtype=torch.float32)
    g_mu, g_sigma = pg.grads(eps, es, sigma, r, m, 0.01)
    s1 = max(1e-5, -9 - (-10 - 14) / 2)
    s2 = max(1e-5, -9 - (-11 - 12) / 2)
    want_mu = -(((-10 + 14) / (2 * s1)) * e + ((-11 + 12) / (2 * s2)) * want)
    tau = e if (-10 - 14) >= (-11 - 12) else want
    sc = max(1e-5, -9 - (-10 - 14 - 11 - 12) / 4)
    want_sigma = -((((-10 - 14) - (-11 - 12)) / (4 * sc)) * (tau ** 2 / s ** 3 - 1 / s) + 0.01 / s)
    np.testing.assert_allclose(g_mu.numpy(), want_mu, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(g_sigma.numpy(), want_sigma, rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["adam", "rmsprop", "sgd"])
def test_pgpe_meta_optimiser_chain(name):
    """The meta-optimiser chain of GaussianPGPE (planner.py:1786-1792): optimiser -> ema with debias, against
    the recurrences written out."""
    from pyrddlgym_jax_b200.core.pgpe import GaussianPGPE
    pg = GaussianPGPE(optimizer=name, ema_decay=0.8, optimizer_kwargs_mu={"learning_rate": 0.1})
    n = 5
    pg.adam = {k: (torch.zeros(1, n), torch.zeros(1, n)) for k in ("mu", "sigma")}
    pg.ema = {k: torch.zeros(1, n) for k in ("mu", "sigma")}
    x, ema, v, m = torch.zeros(n), torch.zeros(n), torch.zeros(n), torch.zeros(n)
    gen = torch.Generator().manual_seed(0)
    for t in range(1, 5):
        pg.step = t
        g = torch.randn(n, generator=gen)
        got = pg._adam("mu", 0, x, g, pg.optimizer_kwargs_mu, None)
        if name == "adam":
            m = 0.9 * m + 0.1 * g
            v = 0.999 * v + 0.001 * g * g
            upd = (m / (1 - 0.9 ** t)) / ((v / (1 - 0.999 ** t)).sqrt() + 1e-8)
        elif name == "rmsprop":
            v = 0.9 * v + 0.1 * g * g
            upd = g / torch.sqrt(v + 1e-8)
        else:
            upd = g
        ema = 0.8 * ema + 0.2 * (-0.1 * upd)
        torch.testing.assert_close(got, x + ema / (1 - 0.8 ** t), rtol=1e-6, atol=1e-7)
        x = got
    with pytest.raises(NotImplementedError):
        GaussianPGPE(optimizer="lion")


def test_load_config_builds_pgpe():
    from pyrddlgym_jax_b200.core.pgpe import GaussianPGPE
    text = QUADCOPTER_SLP.replace("pgpe=None", "pgpe='GaussianPGPE'\npgpe_kwargs={'init_sigma': 0.5, 'batch_size': 2}")
    planner_args, _, _ = load_config_from_string(text)
    assert isinstance(planner_args["pgpe"], GaussianPGPE)
    assert planner_args["pgpe"].init_sigma == 0.5 and planner_args["pgpe"].batch_size == 2


@pytest.mark.gpu
@pytest.mark.parametrize("plan_kind", ["slp", "drp"])
def test_pgpe_runs_next_to_backprop_and_can_take_over(cuda_device, plan_kind):
    """With a zero learning rate the backprop plan never moves, so any improvement of the best
    return must come from the PGPE instance replacing it (planner.py:3283-3295)."""
    from pyrddlgym_jax_b200.core.pgpe import GaussianPGPE
    m = fixture_model("powergen" if plan_kind == "drp" else "reservoir")
    plan = JaxDeepReactivePolicy(topology=[16, 16]) if plan_kind == "drp" else JaxStraightLinePlan()
    pl = JaxBackpropPlanner(m, plan, batch_size_train=64, rollout_horizon=8, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0},
                            pgpe=GaussianPGPE(init_sigma=0.3, optimizer_kwargs_mu={"learning_rate": 0.2}))
    first, last, took_over = None, None, False
    for cb in pl.optimize_generator(key=4, epochs=30, print_summary=False, print_progress=False):
        first = cb if first is None else first
        last = cb
        took_over |= bool(cb["pgpe_improved"])
        assert cb["pgpe_return"] is not None and np.isfinite(cb["pgpe_return"]).all()
        assert cb["pgpe_params"] is not None
    assert took_over
    assert last["best_return"] > first["best_return"]
    if plan_kind == "slp":
        # batched meta-gradients (planner.py:2045-2058): every member scales by max(r_max before the
        # batch, its own rewards); the running maximum afterwards is the maximum over the members
        pg = GaussianPGPE(init_sigma=0.3, batch_size=3, optimizer_kwargs_mu={"learning_rate": 0.2})
        pl = JaxBackpropPlanner(m, plan, batch_size_train=64, rollout_horizon=8, optimizer="sgd",
                                optimizer_kwargs={"learning_rate": 0.0}, pgpe=pg)
        seen = []
        for cb in pl.optimize_generator(key=4, epochs=6, print_summary=False, print_progress=False):
            assert np.isfinite(cb["pgpe_return"]).all()
            seen.append(float(pg.r_max[0]))
        assert all(b >= a for a, b in zip(seen, seen[1:])) and np.isfinite(seen[-1])
        # projected samples, rmsprop meta-optimiser with an EMA of its updates (planner.py:1786-1792, 1927-1940)
        pg = GaussianPGPE(init_sigma=0.3, optimizer="rmsprop", ema_decay=0.9, project_samples=True,
                          optimizer_kwargs_mu={"learning_rate": 0.05})
        pl = JaxBackpropPlanner(m, plan, batch_size_train=64, rollout_horizon=8, optimizer="sgd",
                                optimizer_kwargs={"learning_rate": 0.0}, pgpe=pg)
        rets = [float(np.max(cb["pgpe_return"])) for cb in pl.optimize_generator(
            key=4, epochs=25, print_summary=False, print_progress=False)]
        assert np.isfinite(rets).all() and max(rets[-5:]) > rets[0]


@pytest.mark.gpu
def test_steps_per_update_equals_repeated_updates(cuda_device):
    """steps_per_update = K performs K optimiser steps per iteration (planner.py:2905-2914): two
    iterations with K = 3 land on the same plan as six iterations with K = 1 (fixed noise)."""
    m = fixture_model("quadcopter")

    def run(K, iters, graph):
        pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=64, rollout_horizon=10,
                                steps_per_update=K, optimizer="adam",
                                optimizer_kwargs={"learning_rate": 0.02}, pgpe=None,
                                use_cuda_graph=graph)
        pl.initialize(9)
        for _ in range(iters):
            pl.iterate()
        torch.cuda.synchronize()
        return pl.params[0].cpu().clone()
    ref = run(1, 6, False)
    for graph in (False, True):
        assert torch.equal(run(3, 2, graph), ref)
        assert torch.equal(run(1, 6, graph), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["reservoir", "wildfire"])
def test_stochastic_plan_update_matches_oracle(cuda_device, key):
    """One planner update of a stochastic straight-line plan (planner.py:811-834, 2811-2819):
    loss (returns + entropy regulariser) and gradient w.r.t. mu / log_sigma / logits against the
    oracle's autograd on the same parameters and supplied noise."""
    from oracle import planner as OP
    from oracle.model_eval import OracleModel
    from pyrddlgym_jax_b200.core.layout import noise_as_list
    m = fixture_model(key)
    B, T = 16, 5
    plan = JaxStraightLinePlan(stochastic=True, sigma_entropy_grad=(key == "reservoir"))
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None,
                            use_cuda_graph=False, start_entropy_coeff=0.3, end_entropy_coeff=0.3)
    pl.initialize(5)
    flat = pl.params[0].cpu().view(T, plan.A).clone()   # the update's concurrency projection moves them
    pl.iterate()
    torch.cuda.synchronize()
    tree = plan.params_to_pytree(flat)
    P = {}
    for name in m.action_fluents:
        if m.variable_ranges[name] == "bool":
            P[name] = tree["bool"][name]["logit"].reshape(T, -1).clone().requires_grad_(True)
        else:
            P[name] = tree["real"][name]["mu"].reshape(T, -1).clone().requires_grad_(True)
            P[name + "/log_sigma"] = tree["real"][name]["log_sigma"].reshape(T, -1).clone(
                ).requires_grad_(True)
    om = OracleModel(m, pl.compiled.relax_config())
    layout = pl.train_fwd.program.noise_layout
    noise = pl.train_noise.cpu().view(T, -1)
    steps = [noise_as_list(layout, noise, t, B) for t in range(T)]

    def pol(t, fls, pnoise):
        return OP.slp_train_actions(m, P, t, stochastic=True, pnoise=pnoise)
    pol.n_noise = len(OP.slp_policy_noise_spec(m))
    ref = OP.rollout(om, pol, B, T, steps)
    ent = sum(OP.slp_entropy(m, P, t, sigma_entropy_grad=(key == "reservoir")) for t in range(T))
    loss = OP.mean_loss(ref["reward"], float(m.discount)) - 0.3 * ent
    loss.backward()
    from tests.helpers import pack_params
    want = pack_params(m, {k: (v.grad if v.grad is not None else torch.zeros_like(v))
                           for k, v in P.items()})
    got = pl.grad_used[0].cpu().view(T, plan.A)
    np.testing.assert_allclose(float(pl.train_loss[0]), float(loss), rtol=2e-5)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-4,
                               atol=1e-4 * float(want.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["quadcopter", "reservoir"])
def test_host_driven_step_equals_separate_calls(cuda_device, key):
    """iterate_from_host (host copies captured in the iteration's CUDA graph: one launch, one
    synchronisation per replanning step) == load_initial_state + params upload + iterate +
    read_stats, bit for bit."""
    m = fixture_model(key)

    def make():
        pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=32, rollout_horizon=8,
                                optimizer_kwargs={"learning_rate": 0.05}, pgpe=None)
        pl.initialize(3)
        return pl
    a, b = make(), make()
    pa = a.params.cpu().pin_memory()
    pb = b.params.cpu().pin_memory()
    for it in range(4):
        a.load_initial_state(a.init_train_host, a.init_test_host)
        a.params.copy_(pa, non_blocking=True)
        a.iterate()
        ta, sa, za = a.read_stats()
        pa.copy_(a.params.cpu())
        tb, sb, zb, out = b.iterate_from_host(b.init_train_host, b.init_test_host, pb)
        pb.copy_(out)
        assert np.array_equal(ta, tb) and np.array_equal(sa, sb) and np.array_equal(za, zb), it
        assert torch.equal(pa, pb), it


@pytest.mark.gpu
def test_wrap_softmax_plan_runs_and_respects_the_constraint(cuda_device):
    """wrap_softmax with max-nondef-actions = 1 (planner.py:764-808): pooled '__bool__' logits, no
    projection, and a test policy that can switch on at most one boolean action per step."""
    m = fixture_model("wildfire")
    plan = JaxStraightLinePlan(wrap_softmax=True, softmax_weight=5.0)
    pl = JaxBackpropPlanner(m, plan, batch_size_train=64, rollout_horizon=6,
                            optimizer_kwargs={"learning_rate": 0.5}, pgpe=None)
    pl.initialize(1)
    assert plan.stack_bool and not plan.needs_concurrency_projection
    p0 = pl.params[0].cpu().clone()
    for _ in range(5):
        pl.iterate()
    tr, te, _ = pl.read_stats()
    assert np.isfinite(tr).all() and np.isfinite(te).all()
    flat = pl.params[0].cpu()
    assert not torch.equal(flat, p0)
    tree = plan.params_to_pytree(flat.view(6, plan.A))
    assert set(tree["bool"]) == {"__bool__"} and tree["bool"]["__bool__"]["logit"].shape == (6, plan.A)
    back = plan.pytree_to_params(tree)
    assert torch.equal(back.reshape(-1), flat.reshape(-1))
    for t in range(6):
        acts = plan.test_actions(flat.view(6, plan.A)[t].numpy())
        on = sum(int(np.sum(a != plan.noop[name])) for name, a in acts.items())
        assert on <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
def test_optimizer_chain_noise_and_ema(cuda_device, graph):
    """add_noise and ema links of the optax chain (planner.py:2366-2386).  With eta = 0 the noise
    link is the identity, and the EMA link is checked against its recurrence: the applied step is
    the debiased moving average of the plain optimiser steps (sgd: -lr * clipped gradient)."""
    m = fixture_model("quadcopter")

    def make(**kw):
        pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=32, rollout_horizon=6,
                                optimizer="sgd", optimizer_kwargs={"learning_rate": 0.01},
                                clip_grad=0.5, pgpe=None, use_cuda_graph=graph, **kw)
        pl.initialize(2)
        return pl
    decay = 0.8
    pl = make(ema_decay=decay, noise_kwargs={"eta": 0.0, "gamma": 0.55, "seed": 1})
    ema = torch.zeros_like(pl.params[0]).cpu()
    for it in range(1, 5):
        before = pl.params[0].cpu().clone()
        pl.iterate()
        torch.cuda.synchronize()
        g = pl.grad_used[0].cpu()
        gn = float(torch.linalg.vector_norm(g))
        step = -0.01 * g * min(1.0, 0.5 / gn)
        ema = decay * ema + (1 - decay) * step
        want = before + ema / (1 - decay ** it)
        lo, hi = pl.box_lo.cpu(), pl.box_hi.cpu()
        want = torch.minimum(torch.maximum(want, lo), hi)
        torch.testing.assert_close(pl.params[0].cpu(), want, rtol=1e-5, atol=1e-6)
    noisy = make(noise_kwargs={"eta": 1.0, "gamma": 0.55, "seed": 3})
    plain = make()
    for _ in range(2):
        noisy.iterate()
        plain.iterate()
    torch.cuda.synchronize()
    assert not torch.equal(noisy.params, plain.params)
    assert torch.isfinite(noisy.params).all()


@pytest.mark.gpu
def test_symlog_reward_matches_oracle(cuda_device):
    """use_symlog_reward (planner.py:2761-2763): the train loss is -mean(sign(R) log1p|R|)."""
    from oracle import planner as OP
    from oracle.model_eval import OracleModel
    m = fixture_model("quadcopter")
    B, T = 24, 6
    plan = JaxStraightLinePlan()
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None,
                            use_cuda_graph=False, use_symlog_reward=True)
    pl.initialize(4)
    flat = pl.params[0].cpu().view(T, plan.A).clone()
    pl.iterate()
    torch.cuda.synchronize()
    tree = plan.params_to_pytree(flat)
    P = {name: tree["real"][name]["mu"].reshape(T, -1).clone().requires_grad_(True)
         for name in m.action_fluents}
    om = OracleModel(m, pl.compiled.relax_config())
    ref = OP.rollout(om, lambda t, fls: OP.slp_train_actions(m, P, t), B, T, None)
    R = OP.returns_from_rewards(ref["reward"], float(m.discount))
    loss = -(torch.sign(R) * torch.log1p(R.abs())).mean()
    loss.backward()
    from tests.helpers import pack_params
    want = pack_params(m, {k: v.grad for k, v in P.items()})
    np.testing.assert_allclose(float(pl.train_loss[0]), float(loss.detach()), rtol=2e-5)
    got = pl.grad_used[0].cpu().view(T, plan.A)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-4,
                               atol=1e-4 * float(want.abs().max()))


# ---- reduce_on_plateau link of the optimiser chain (planner.py:2375-2384) -------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kw", [{}, {"factor": 0.5, "patience": 3, "cooldown": 2, "min_scale": 0.2},
                                {"patience": 2, "accumulation_size": 3, "factor": 0.7, "rtol": 0.05},
                                {"patience": 1, "atol": 0.5, "factor": 0.25}])
def test_plateau_state_machine_matches_oracle(cuda_device, kw):
    """The reduce_on_plateau state of rk_optimizer_chain_step against the oracle's restatement of
    optax.contrib.reduce_on_plateau, loss by loss (a frozen one-word sgd step carries the link)."""
    from oracle import planner as OP
    from pyrddlgym_jax_b200.core import engine
    from pyrddlgym_jax_b200.core.planner import plateau_init, plateau_kwargs
    full = plateau_kwargs(kw)
    g = torch.Generator().manual_seed(5)
    # a loss that improves, stalls, jumps and improves again
    losses = torch.cat([torch.linspace(10, 5, 8), 5 + 0.3 * torch.rand(25, generator=g),
                        torch.linspace(5, 1, 6), 1 + 0.2 * torch.rand(20, generator=g)])
    dev = cuda_device
    st = plateau_init(1, dev)[0]
    z = lambda *s: torch.zeros(*s, device=dev)      # noqa: E731
    hyper = torch.tensor([0.0, 0, 0, 0, 0, 0, 1, 1, 1], dtype=torch.float32, device=dev)
    params, grad, count, scratch, loss = z(1), z(1), z(1), z(8), z(1)
    cfg = {"noise": None, "ema_decay": None, "plateau": full}
    ref = OP.reduce_on_plateau_init()
    scales = []
    for i, v in enumerate(losses):
        loss.fill_(float(v))
        engine.optimizer_chain_step("sgd", hyper, 1, params, None, grad, None, None, None, cfg, count, st,
                                    None, loss, 0, scratch)
        ref = OP.reduce_on_plateau_step(ref, float(v), **full)
        got = st.cpu().tolist()
        want = [ref[k] for k in ("scale", "best_value", "plateau_count", "cooldown_count", "count",
                                 "avg_value")]
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
        scales.append(got[0])
        assert float(count) == i + 1
    assert scales[-1] < 1.0 and min(scales) >= full["min_scale"]
    with pytest.raises(TypeError):
        plateau_kwargs({"patients": 3})


@pytest.mark.gpu
def test_reduce_on_plateau_scales_the_updates(cuda_device):
    """sgd on Reservoir with a plateau test that can never see an improvement after the first step
    (atol 1e9, patience 2): the scale halves every second update, in the eager form and inside the
    captured graph alike, and follows the oracle's state machine on the losses the planner saw."""
    from oracle import planner as OP
    m = fixture_model("reservoir")
    kw = {"factor": 0.5, "patience": 2, "atol": 1e9}
    for graph in (False, True):
        pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=8, rollout_horizon=5,
                                optimizer="sgd", optimizer_kwargs={"learning_rate": 1e-3}, pgpe=None,
                                reduce_on_plateau_kwargs=kw, use_cuda_graph=graph)
        scales = []
        for cb in pl.optimize_generator(key=3, epochs=8, print_summary=False, print_progress=False):
            assert np.isfinite(cb["train_return"]).all()
            scales.append(float(pl._chain_plateau[0, 0]))
        assert all(b <= a for a, b in zip(scales, scales[1:])), scales
        assert scales[-1] <= 0.25 and set(scales) <= {1.0, 0.5, 0.25, 0.125, 0.0625}, scales
    pl2 = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=8, rollout_horizon=5,
                             optimizer="sgd", optimizer_kwargs={"learning_rate": 0.05}, pgpe=None,
                             reduce_on_plateau_kwargs={"factor": 0.5, "patience": 1, "rtol": 1e-3},
                             use_cuda_graph=False)
    pl2.initialize(3)
    ref = OP.reduce_on_plateau_init()
    for _ in range(6):
        before = pl2.params[0].clone()
        pl2.update()
        torch.cuda.synchronize()
        ref = OP.reduce_on_plateau_step(ref, float(pl2.train_loss[0]), factor=0.5, patience=1,
                                        rtol=1e-3)
        assert abs(float(pl2._chain_plateau[0, 0]) - ref["scale"]) < 1e-7
        # sgd: update = -lr * scale * grad (inside the box)
        want = before - 0.05 * ref["scale"] * pl2.grad_used[0]
        if pl2.box_lo is not None:
            want = torch.maximum(want, pl2.box_lo)
        if pl2.box_hi is not None:
            want = torch.minimum(want, pl2.box_hi)
        np.testing.assert_allclose(pl2.params[0].cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol=1e-6)


# ---- stochastic plans with wrap_softmax (planner.py:797-806) ---------------------------------------
def _plan_params_by_fluent(m, plan, flat):
    P = {}
    for name, (off, n, prange) in plan.layout.items():
        P[name] = flat[:, off:off + n].clone().requires_grad_(True)
        if name in plan.sigma_layout:
            so, sn = plan.sigma_layout[name]
            P[name + "/log_sigma"] = flat[:, so:so + sn].clone().requires_grad_(True)
    return P


@pytest.mark.parametrize("key,kw", [("wildfire", {"wrap_softmax": True, "softmax_weight": 3.0}),
                                    ("wildfire_free", {"sigmoid_weight": 2.0}),
                                    ("reservoir", {"sigma_entropy_grad": True})])
def test_entropy_terms_match_oracle_autograd(key, kw):
    """JaxStraightLinePlan.entropy_terms (value and gradient of the horizon's entropy, computed from
    the parameters alone) against autograd of the oracle's per-step entropy."""
    from oracle import planner as OP
    from pyrddlgym_jax_b200.core.compiler import JaxRDDLCompiler
    m = fixture_model(key)
    T = 4
    plan = JaxStraightLinePlan(stochastic=True, **kw)
    plan.compile(logic.DefaultJaxRDDLCompilerWithGrad(m), JaxRDDLCompiler(m), {}, T)
    flat = torch.randn(T, plan.A, generator=torch.Generator().manual_seed(1))
    H, dH = plan.entropy_terms(flat)
    P = _plan_params_by_fluent(m, plan, flat)
    ent = sum(OP.slp_entropy(m, P, t, sigmoid_weight=kw.get("sigmoid_weight", 1.0),
                             sigma_entropy_grad=kw.get("sigma_entropy_grad", False),
                             pooled_softmax=plan.stack_bool,
                             softmax_weight=kw.get("softmax_weight", 1.0)) for t in range(T))
    ent.backward()
    from tests.helpers import pack_params
    want = pack_params(m, {k: (v.grad if v.grad is not None else torch.zeros_like(v))
                           for k, v in P.items()})
    np.testing.assert_allclose(float(H), float(ent.detach()), rtol=1e-5)
    np.testing.assert_allclose(dH.numpy(), want.numpy(), rtol=1e-4, atol=1e-6)
    assert float(dH.abs().max()) > 0


@pytest.mark.gpu
def test_stochastic_wrap_softmax_update_matches_oracle(cuda_device):
    """One update of a stochastic plan whose boolean parameters are pooled under one softmax:
    Gumbel-perturbed logits per rollout, softmax entropy regulariser; loss and gradient against the
    oracle's autograd on the same parameters and supplied noise."""
    from oracle import planner as OP
    from oracle.model_eval import OracleModel
    from pyrddlgym_jax_b200.core.layout import noise_as_list
    m = fixture_model("wildfire")
    B, T = 16, 5
    plan = JaxStraightLinePlan(stochastic=True, wrap_softmax=True, softmax_weight=3.0)
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                            optimizer_kwargs={"learning_rate": 0.0}, pgpe=None,
                            use_cuda_graph=False, start_entropy_coeff=0.3, end_entropy_coeff=0.3)
    pl.initialize(5)
    pl.params[0].copy_(torch.randn(T * plan.A, generator=torch.Generator().manual_seed(2)))
    flat = pl.params[0].cpu().view(T, plan.A).clone()
    pl.iterate()
    torch.cuda.synchronize()
    P = _plan_params_by_fluent(m, plan, flat)
    om = OracleModel(m, pl.compiled.relax_config())
    layout = pl.train_fwd.program.noise_layout
    noise = pl.train_noise.cpu().view(T, -1)
    steps = [noise_as_list(layout, noise, t, B) for t in range(T)]

    def pol(t, fls, pnoise):
        return OP.slp_train_actions(m, P, t, stochastic=True, pnoise=pnoise, pooled_softmax=True,
                                    softmax_weight=3.0)
    pol.n_noise = len(OP.slp_policy_noise_spec(m, pooled_softmax=True))
    ref = OP.rollout(om, pol, B, T, steps)
    ent = sum(OP.slp_entropy(m, P, t, pooled_softmax=True, softmax_weight=3.0) for t in range(T))
    loss = OP.mean_loss(ref["reward"], float(m.discount)) - 0.3 * ent
    loss.backward()
    from tests.helpers import pack_params
    want = pack_params(m, {k: (v.grad if v.grad is not None else torch.zeros_like(v))
                           for k, v in P.items()})
    got = pl.grad_used[0].cpu().view(T, plan.A)
    np.testing.assert_allclose(float(pl.train_loss[0]), float(loss.detach()), rtol=2e-5)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-4,
                               atol=1e-4 * float(want.abs().max()))


# ---- utilities (planner.py:2161-2247) ----------------------------------------------------------------
def test_utilities_known_answers():
    """Closed forms on a small return vector, written from the reference's formulas with numpy."""
    from pyrddlgym_jax_b200.core.planner import UTILITY_LOOKUP, _utility_fn
    assert set(UTILITY_LOOKUP) == {"mean", "mean_var", "mean_std", "mean_semivar", "mean_semidev",
                                   "sharpe", "entropic", "exponential", "var", "cvar", "cvar_ste",
                                   "expectile", "gini"}
    r = np.array([3.0, -1.0, 4.0, 1.0, -5.0, 9.0, 2.0, 6.0], np.float64)
    t = torch.tensor(r, dtype=torch.float32)
    mu, beta, alpha = r.mean(), 0.7, 0.25
    down = np.minimum(0.0, r - mu)
    var = np.percentile(r, 100 * alpha)
    mask = r <= var
    ex = mu
    for _ in range(30):
        w = np.where(r > ex, alpha, 1 - alpha)
        ex = (w * r).sum() / w.sum()
    want = {
        ("mean_var", "beta"): mu - 0.5 * beta * r.var(),
        ("mean_std", "beta"): mu - 0.5 * beta * r.std(),
        ("mean_semivar", "beta"): mu - 0.5 * beta * (down ** 2).mean(),
        ("mean_semidev", "beta"): mu - 0.5 * beta * np.sqrt((down ** 2).mean()),
        ("entropic", "beta"): -np.log(np.mean(np.exp(-beta * r))) / beta,
        ("var", "alpha"): var,
        ("cvar", "alpha"): var - np.maximum(var - r, 0).mean() / alpha,
        ("cvar_ste", "alpha"): (mask * r).sum() / max(1, mask.sum()),
        ("expectile", "alpha"): ex,
        ("gini", "gamma"): mu - 0.3 * 2.0 * np.maximum(-(r[:, None] - r[None, :]), 0).mean(),
    }
    for (name, key), w in want.items():
        val = {"beta": beta, "alpha": alpha, "gamma": 0.3}[key]
        got = float(_utility_fn(name, {key: val})(t))
        assert abs(got - w) <= 1e-5 * max(1.0, abs(w)), (name, got, w)
    got = float(_utility_fn("sharpe", {"risk_free": 0.5})(t))
    assert abs(got - (mu - 0.5) / (r.std() + 1e-12)) < 1e-5
    # every utility is differentiable w.r.t. the returns (what the adjoint kernel is fed)
    for name, kw in (("cvar", {"alpha": 0.25}), ("expectile", {"alpha": 0.3}), ("gini", {"gamma": 0.5}),
                     ("var", {"alpha": 0.4})):
        x = t.clone().requires_grad_(True)
        (g,) = torch.autograd.grad(_utility_fn(name, kw)(x), x)
        assert torch.isfinite(g).all() and float(g.abs().sum()) > 0, name


def test_get_action_info_matches_reference_contract():
    """planner.py:423-472: shapes carry the horizon, booleans have no box, user bounds override the
    RDDL ones, cond_lists are the four finite-ness masks."""
    from pyrddlgym_jax_b200.core.planner import get_action_info
    m = fixture_model("powergen")
    comp = logic.DefaultJaxRDDLCompilerWithGrad(m)
    shapes, bounds, conds = get_action_info(comp, {}, 7)
    assert shapes == {"curProd": (7, 5)}
    lo, hi = bounds["curProd"]
    np.testing.assert_array_equal(lo, np.asarray(m.bounds["curProd"][0], np.float32) * np.ones(5, np.float32))
    assert [c.shape for c in conds["curProd"]] == [(5,)] * 4
    assert np.all(np.sum(np.stack(conds["curProd"]), 0) == 1)        # exactly one branch per element
    _, b2, c2 = get_action_info(comp, {"curProd": (-np.inf, 3.0)}, 7)
    assert np.all(b2["curProd"][1] == 3.0) and np.all(c2["curProd"][2])
    w = fixture_model("wildfire")
    _, bw, cw = get_action_info(logic.DefaultJaxRDDLCompilerWithGrad(w), {}, 3)
    assert all(v == (None, None) for v in bw.values()) and cw == {}


def test_compiler_relaxation_info():
    """model_parameter_info / overriden_ops_info (compiler.py:804-823): one entry per relaxed
    expression with its weight, grouped by the mixin class that relaxes it."""
    c = logic.DefaultJaxRDDLCompilerWithGrad(fixture_model("reservoir"), sigmoid_weight=7.0)
    info = c.model_parameter_info()
    assert info and all(set(v) == {"id", "rddl_op", "init_value"} for v in info.values())
    assert all(v["init_value"] == {"sigmoid_weight": 7.0} for v in info.values()
               if v["rddl_op"].startswith("relational"))
    over = c.overriden_ops_info()
    assert "SigmoidRelational" in over
    assert sorted(i for ids in over["SigmoidRelational"].values() for i in ids) == \
        sorted(k for k, v in info.items() if v["rddl_op"].startswith("relational"))
    w = logic.DefaultJaxRDDLCompilerWithGrad(fixture_model("wildfire"))
    over = w.overriden_ops_info()
    assert set(over) == {"ReparameterizedSigmoidBernoulli", "ProductNormLogical", "LinearIfElse"}
    assert len(over["ReparameterizedSigmoidBernoulli"]["randomvar Bernoulli"]) == 1
    # relaxable operations over non-fluents stay exact (compiler.py:825-832)
    d = logic.DefaultJaxRDDLCompilerWithGrad(fixture_model("discrete"))
    assert "randomvar Discrete" in d.exact_ops_info()
    assert "randomvar UnnormDiscrete" in d.overriden_ops_info()["GumbelSoftmaxDiscrete"]
    ids = {i for ops in over.values() for v in ops.values() for i in v}
    assert not ids & {i for v in w.exact_ops_info().values() for i in v}


@pytest.mark.gpu
def test_as_optimization_problem(cuda_device):
    """planner.py:2962-3031: loss / gradient of the train objective as functions of a 1-D vector.
    Quadcopter is deterministic, so a small step against the gradient must lower the loss and the
    directional derivative must agree with a central difference."""
    m = fixture_model("quadcopter")
    pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=8, rollout_horizon=12,
                            pgpe=None, use_cuda_graph=False)
    loss_fn, grad_fn, guess, unravel = pl.as_optimization_problem(key=4)
    assert guess.shape == (12 * pl.A,) and guess.dtype == np.float64
    f0, g0 = loss_fn(guess), grad_fn(guess)
    assert np.isfinite(f0) and g0.shape == guess.shape and np.abs(g0).max() > 0
    d = g0 / np.linalg.norm(g0)
    eps = 1e-2
    slope = float(g0 @ d)
    noise = 1e-6 * max(1.0, abs(f0))            # fp32 rounding of the loss value
    if eps * slope > 50 * noise:                # the step is visible above the rounding
        fd = (loss_fn(guess + eps * d) - loss_fn(guess - eps * d)) / (2 * eps)
        assert abs(fd - slope) <= 5e-2 * abs(slope) + 50 * noise / eps, (fd, slope)
        assert loss_fn(guess - eps * d) < f0
    tree = unravel(guess)
    assert tuple(tree["real"]["thrust"]["mu"].shape) == (12, 4)
    with pytest.raises(ValueError):
        JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=8, rollout_horizon=12,
                           pgpe=None, parallel_updates=2,
                           use_cuda_graph=False).as_optimization_problem(key=1)


@pytest.mark.gpu
def test_optimize_prints_summaries(cuda_device, capsys):
    """optimize() with the reference's default print_summary=True / print_hyperparams=True
    (planner.py:3132-3147): system line, relaxation table, hyper-parameter listing."""
    m = fixture_model("reservoir")
    pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=8, rollout_horizon=5, pgpe=None)
    cb = pl.optimize(key=1, epochs=3, print_progress=False, print_hyperparams=True)
    out = capsys.readouterr().out
    assert "SigmoidRelational" in out and "relational" in out
    assert "optimizer hyper-parameters" in out and "batch_size_train  =8" in out
    assert "Summary of optimization" in out and cb["iteration"] == 2


DRP_FULL = """
[Compiler]
method='DefaultJaxRDDLCompilerWithGrad'

[Planner]
method='JaxDeepReactivePolicy'
method_kwargs={'topology': [32, 32], 'activation': 'relu', 'normalize': True, 'time_dependent': True, 'time_embedding': 'FixedTimeEmbedding', 'time_embedding_kwargs': {'max_t': 50, 'dim': 4}, 'action_noise_dist': 'laplace'}
optimizer='adam'
optimizer_kwargs={'learning_rate': 0.001}
preprocessor='StaticNormalizer'
preprocessor_kwargs={'fluent_bounds': {'temperature': (-30.0, 40.0)}}
reduce_on_plateau_kwargs={'factor': 0.5, 'patience': 5}
dashboard=False
batch_size_train=8
pgpe=None

[Optimize]
key=1
epochs=10
"""


def test_load_config_resolves_drp_options_and_preprocessor():
    """planner.py:157-173, 200-217: activation / time embedding / noise law by name, the preprocessor
    class with its kwargs, the dashboard key dropped."""
    from pyrddlgym_jax_b200.core.drp import FixedTimeEmbedding
    from pyrddlgym_jax_b200.core.planner import StaticNormalizer
    planner_args, plan_kwargs, train_args = load_config_from_string(DRP_FULL)
    plan = planner_args["plan"]
    assert isinstance(plan, JaxDeepReactivePolicy) and plan._activation == "relu"
    assert isinstance(plan.time_embed, FixedTimeEmbedding) and plan.time_embed.max_t == 50
    assert plan._action_noise_dist == "laplace" and plan._normalize
    assert isinstance(planner_args["preprocessor"], StaticNormalizer)
    assert planner_args["preprocessor"].fluent_bounds == {"temperature": (-30.0, 40.0)}
    assert "dashboard" not in planner_args and "preprocessor_kwargs" not in planner_args
    assert planner_args["reduce_on_plateau_kwargs"] == {"factor": 0.5, "patience": 5}
    for bad in (DRP_FULL.replace("'relu'", "'hard_tanh'"), DRP_FULL.replace("FixedTimeEmbedding", "Nope"),
                DRP_FULL.replace("'StaticNormalizer'", "'JaxPlan'")):
        with pytest.raises(ValueError):
            load_config_from_string(bad)


REFERENCE_CONFIGS = "/root/reference/pyRDDLGym_jax/examples/configs"


@pytest.mark.skipif(not __import__("os").path.isdir(REFERENCE_CONFIGS),
                    reason="the reference tree is only present in the build container")
def test_every_shipped_reference_config_loads():
    """Drop-in check of the .cfg format: every hyper-parameter file the reference ships
    (examples/configs/*.cfg) loads and instantiates its plan; the three tuning_*.cfg templates hold
    placeholders that the reference's tuner substitutes first."""
    import glob
    from pyrddlgym_jax_b200.core.planner import load_config
    files = sorted(glob.glob(REFERENCE_CONFIGS + "/*.cfg"))
    assert len(files) >= 20
    loaded = 0
    for f in files:
        if __import__("os").path.basename(f).startswith("tuning_"):
            continue
        planner_args, plan_kwargs, train_args = load_config(f)
        assert isinstance(planner_args["plan"], (JaxStraightLinePlan, JaxDeepReactivePolicy)), f
        assert "compiler_kwargs" in planner_args and isinstance(train_args, dict)
        loaded += 1
    assert loaded >= 20


@pytest.mark.gpu
def test_parallel_updates_branches_equal_sequential(cuda_device, monkeypatch):
    """parallel_updates = 3 (planner.py:2920-2929 vmaps the update over the instances): the captured
    iteration runs the instances as concurrent graph branches with private scratch; plans, losses and
    test losses are bit-identical to the instances run one after another."""
    m = fixture_model("quadcopter")      # no draw sites: both planners see identical inputs
    out = []
    for branches in ("1", "0"):
        monkeypatch.setenv("RK_PARALLEL_BRANCHES", branches)
        pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=32, rollout_horizon=6,
                                optimizer="rmsprop", optimizer_kwargs={"learning_rate": 0.05}, pgpe=None,
                                parallel_updates=3)
        assert pl._branches == (branches == "1")
        pl.initialize(5)
        hist = []
        for _ in range(4):
            pl.redraw_noise()
            pl.iterate()
            hist.append([x.copy() for x in pl.read_stats()])
        out.append((pl.params.cpu().clone(), hist))
    assert not torch.equal(out[0][0][0], out[0][0][1])          # the instances differ from each other
    assert torch.equal(out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)


@pytest.mark.gpu
def test_zoom_line_search_link(cuda_device):
    """line_search_kwargs (planner.py:2373-2374): every update's step size satisfies the sufficient-decrease
    and curvature conditions along the optimiser's direction, checked by re-evaluating the relaxed loss and
    its gradient at the accepted point on the same (deterministic) rollouts; the loss goes down every
    iteration, and a too large learning rate that plain sgd diverges with is tamed."""
    m = fixture_model("quadcopter")
    kw = dict(batch_size_train=16, rollout_horizon=8, optimizer="sgd", pgpe=None)
    pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), optimizer_kwargs={"learning_rate": 5.0},
                            line_search_kwargs={"max_linesearch_steps": 12, "curv_rtol": 0.5}, **kw)
    pl.initialize(3)
    losses = []
    for _ in range(5):
        x0 = pl.params[0].clone()
        pl.iterate()
        torch.cuda.synchronize()
        info = pl.line_search_info
        f0, g0 = float(pl.train_loss[0]), pl.grad_used[0].clone()
        losses.append(f0)
        if info["learning_rate"] == 0.0:
            continue
        x1 = pl.params[0].clone()
        u = (x1 - x0) / info["learning_rate"]
        inside = True
        if pl.box_lo is not None:      # the Wolfe test is about the un-projected step
            inside = bool(((x0 + info["learning_rate"] * u >= pl.box_lo) &
                           (x0 + info["learning_rate"] * u <= pl.box_hi)).all())
        slope0 = float(torch.dot(g0, u))
        assert slope0 < 0
        pl._train_grad_kernels(0)                    # loss and gradient at the accepted point
        torch.cuda.synchronize()
        f1, slope1 = float(pl._train_loss_tmp[0]), float(torch.dot(pl.grad[0], u))
        if inside:
            assert f1 <= f0 + 1e-4 * info["learning_rate"] * slope0 + 1e-5 * abs(f0)
            if info["converged"]:
                assert abs(slope1) <= 0.5 * abs(slope0) * (1 + 1e-3)
        assert info["num_linesearch_steps"] <= 12
    assert all(b <= a + 1e-6 * abs(a) for a, b in zip(losses, losses[1:])), losses
    assert losses[-1] < losses[0]
    plain = JaxBackpropPlanner(m, JaxStraightLinePlan(), optimizer_kwargs={"learning_rate": 5.0}, **kw)
    plain.initialize(3)
    for _ in range(5):
        plain.iterate()
    final_plain = float(plain.read_stats()[0][0])
    assert not np.isfinite(final_plain) or losses[-1] <= final_plain


@pytest.mark.gpu
@pytest.mark.parametrize("key,utility,kw", [("quadcopter", "mean", {}), ("wildfire_free", "mean", {}),
                                            ("quadcopter", "mean_var", {"beta": 0.3})])
def test_critic_terminal_value_matches_oracle(cuda_device, key, utility, kw):
    """critic (planner.py:2721-2745, 2758-2760): returns += gamma^T critic(final state, step-0 action).
    Train loss and plan gradient (through the critic's state cotangent, which seeds the adjoint kernel,
    and through the step-0 action) against the oracle's rollout + torch autograd; the test loss adds the
    critic on the exact final state.  The critic is a torch callable here (no JAX in this build)."""
    from oracle import planner as OP
    from oracle.model_eval import OracleModel
    from pyrddlgym_jax_b200.core import planner as PL
    from tests.helpers import pack_params
    m = fixture_model(key)
    B, T = 24, 5
    plan = JaxStraightLinePlan()
    state_names = list(m.state_fluents)
    wvec = {s: 0.3 + 0.1 * i for i, s in enumerate(state_names)}

    def critic(cp, obs, action):
        v = 0.0
        for s in state_names:
            x = obs[s].to(torch.float32)
            v = v - cp["scale"] * wvec[s] * (x.reshape(x.shape[0], -1) - 0.25).square().sum(1)
        for a in action.values():
            v = v + 0.05 * a.to(torch.float32).sum()
        return v
    cp = {"scale": 0.7}
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, batch_size_test=B, rollout_horizon=T,
                            optimizer="sgd", optimizer_kwargs={"learning_rate": 0.0}, pgpe=None,
                            critic=critic, utility=utility, utility_kwargs=kw)
    pl.initialize(4)
    pl.critic_params = cp
    gen = torch.Generator().manual_seed(9)
    flat = torch.randn(T, plan.A, generator=gen) * 0.3
    pl.params[0].copy_(flat.reshape(-1))
    fw = pl.train_fwd.program
    noise = None
    if fw.N:
        from pyrddlgym_jax_b200.core.layout import draw_noise, noise_as_list
        noise = draw_noise(fw.noise_layout, T, B, gen)
        pl.train_noise.copy_(noise.reshape(-1))
    pl.iterate()
    torch.cuda.synchronize()
    assert pl._graph is None                      # host-driven mode
    tree = plan.params_to_pytree(flat)
    P = {}
    for name in m.action_fluents:
        kind = "bool" if m.variable_ranges[name] == "bool" else "real"
        leaf = tree[kind][name]["logit" if kind == "bool" else "mu"]
        P[name] = leaf.reshape(T, -1).clone().requires_grad_(True)
    om = OracleModel(m, pl.compiled.relax_config())
    nl = None if noise is None else [noise_as_list(fw.noise_layout, noise, t, B) for t in range(T)]
    ref = OP.rollout(om, lambda t, fls: OP.slp_train_actions(m, P, t), B, T, nl)
    a0 = {k: v.reshape(-1)[:m.variable_size(k)].reshape(m.variable_shape(k))
          for k, v in OP.slp_train_actions(m, P, 0).items()}
    gamma = float(m.discount)
    R = OP.returns_from_rewards(ref["reward"], gamma) + gamma ** T * critic(cp, ref["final"], a0)
    loss = -(R.mean() if utility == "mean" else PL.UTILITY_LOOKUP[utility](R, **kw))
    loss.backward()
    want = pack_params(m, {k: v.grad for k, v in P.items()}).numpy()
    assert abs(float(pl.train_loss[0]) - float(loss.detach())) <= 2e-5 * max(1.0, abs(float(loss.detach())))
    got = pl.grad_used[0].cpu().view(T, plan.A).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * float(np.abs(want).max()))
    assert float(np.abs(want[0]).max()) > 0
    # test loss: exact rollouts + critic on the exact final state and the test policy's step-0 action
    te = pl.test_fwd.program
    finals = unpack_test_final(pl, te, B)
    acts = plan.test_actions(flat[0].numpy())
    acts = {k: torch.as_tensor(np.asarray(v)) for k, v in acts.items()}
    Rt = pl.test_reward[0].view(T, B).cpu() * (gamma ** torch.arange(T, dtype=torch.float32)).view(T, 1)
    Rt = Rt.sum(0) + gamma ** T * critic(cp, finals, acts)
    want_t = -(Rt.mean() if utility == "mean" else PL.UTILITY_LOOKUP[utility](Rt, **kw))
    assert abs(float(pl.test_loss_buf[0]) - float(want_t)) <= 2e-5 * max(1.0, abs(float(want_t)))


def unpack_test_final(pl, te, B):
    from pyrddlgym_jax_b200.core.layout import unpack_fluent_major
    shapes = {name: tuple(pl.rddl.variable_shape(name)) for name in te.state_layout}
    return {k: v.cpu() for k, v in unpack_fluent_major(te.state_layout, pl.test_final[0], B, shapes).items()}


@pytest.mark.gpu
def test_reward_mask(cuda_device):
    """planner_state.reward_mask (planner.py:2802-2809): rewards[:, t] * mask[t] inside the loss.  The loss is
    the masked sum of the logged rewards; with the tail masked out, the plan parameters of the masked steps
    receive no gradient (their actions only reach rewards that no longer count)."""
    m = fixture_model("quadcopter")
    B, T = 32, 8
    pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=B, rollout_horizon=T,
                            optimizer="sgd", optimizer_kwargs={"learning_rate": 0.0}, pgpe=None)
    mask = np.array([1, 1, 0.5, 1, 1, 0, 0, 0], np.float32)
    pl.set_reward_mask(mask)
    pl.initialize(3)
    pl.params[0].copy_(torch.randn(pl.n_params, generator=torch.Generator().manual_seed(2)) * 0.1)
    for _ in range(2):
        pl.iterate()
    torch.cuda.synchronize()
    r = pl.train_reward[0].view(T, B).cpu().double()
    want = -float((r * torch.from_numpy(mask).double().view(T, 1)).sum(0).mean())
    assert abs(float(pl.train_loss[0]) - want) <= 1e-5 * abs(want)
    rt = pl.test_reward[0].view(T, B).cpu().double()
    want_t = -float((rt * torch.from_numpy(mask).double().view(T, 1)).sum(0).mean())
    assert abs(float(pl.test_loss_buf[0]) - want_t) <= 1e-5 * abs(want_t)
    g = pl.grad_used[0].cpu().view(T, -1)
    assert float(g[:5].abs().max()) > 0 and float(g[5:].abs().max()) == 0.0
    pl.set_reward_mask(None)
    pl.iterate()
    pl.iterate()
    torch.cuda.synchronize()
    assert float(pl.grad_used[0].cpu().view(T, -1)[5:].abs().max()) > 0
    with pytest.raises(ValueError):
        pl.set_reward_mask(np.ones(T + 1))
    with pytest.raises(ValueError):
        pl.set_reward_mask(np.ones((T, 1)))


@pytest.mark.gpu
def test_closure_seams_as_callables(cuda_device):
    """plan.projection / plan.initializer / test_policy and planner.train_rollouts / test_rollouts
    (planner.py:386-420, 2673-2698): the reference's jitted closures as plain callables over the parameter
    pytree."""
    m = fixture_model("wildfire")                   # max-nondef-actions binds: concurrency projection
    plan = JaxStraightLinePlan()
    B, T = 16, 5
    pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, pgpe=None)
    pl.initialize(3)
    tree, hyper = plan.initializer(7)
    assert set(hyper) == set(m.action_fluents)
    flat = plan.pytree_to_params(tree)
    assert tuple(flat.shape) == (T, plan.A)
    wild = plan.params_to_pytree(torch.randn(T, plan.A, generator=torch.Generator().manual_seed(1)) * 3)
    proj, converged = plan.projection(wild)
    assert converged
    pflat = plan.pytree_to_params(proj)
    for t in range(T):
        acts = plan.test_actions(pflat[t].numpy())
        on = sum(int(np.sum(a != plan.noop[name])) for name, a in acts.items())
        assert on <= m.max_allowed_actions
    a0 = plan.test_policy(None, proj, hyper, 0, {})
    assert set(a0) == set(m.action_fluents)
    log = pl.test_rollouts(proj)
    assert tuple(log["reward"].shape) == (1, B, T) and tuple(log["error"].shape) == (1, B, T)
    assert torch.isfinite(log["reward"]).all() and log["fluents"]
    tlog = pl.train_rollouts(proj)
    assert tuple(tlog["reward"].shape) == (1, B, T) and torch.isfinite(tlog["reward"]).all()
    # the exact test rollouts of a projected plan and of the same plan set through update/test agree
    pl.params[0].copy_(pflat.reshape(-1).to(pl.device))
    pl.test_loss()
    torch.cuda.synchronize()
    again = pl._test_log()["reward"]
    assert torch.equal(again, log["reward"])


def test_entry_point_resolution_and_defaults():
    """Command line (entry_point.py:6-57, examples/run_plan.py): bundled fixtures resolve by name, the three
    method names carry the reference's default settings, a .cfg path is loaded, bad input is rejected."""
    import os
    from pyrddlgym_jax_b200 import entry_point as EP
    d, i = EP.resolve_model_files("Reservoir_Continuous", "1")
    assert d.endswith("Reservoir_Continuous/domain.rddl") and i.endswith("instance1.rddl")
    d2, i2 = EP.resolve_model_files(d, i)
    assert (d2, i2) == (d, i)
    with pytest.raises(FileNotFoundError):
        EP.resolve_model_files("NoSuchDomain", "1")
    with pytest.raises(FileNotFoundError):
        EP.resolve_model_files("Reservoir_Continuous", "99")
    pa, pk, ta, online = EP.load_method("slp")
    assert pa["batch_size_train"] == 32 and pa["optimizer_kwargs"] == {"learning_rate": 0.01}
    assert ta["epochs"] == 30000 and ta["train_seconds"] == 60 and not online
    assert isinstance(pa["plan"], JaxStraightLinePlan)
    pa, pk, ta, online = EP.load_method("replan", train_seconds=0.25)
    assert online and pa["rollout_horizon"] == 5 and ta["train_seconds"] == 0.25 and ta["print_summary"] is False
    pa, _, _, _ = EP.load_method("drp")
    assert isinstance(pa["plan"], JaxDeepReactivePolicy) and pa["optimizer_kwargs"] == {"learning_rate": 0.0001}
    with pytest.raises(ValueError):
        EP.load_method("nope")
    with pytest.raises(ValueError):
        EP.run_tune("Reservoir_Continuous", "1", "nope")
    # the tuning templates parse once their tags are replaced
    from pyrddlgym_jax_b200.core.planner import load_config_from_string
    cfg = EP._TUNE_CFG.format(plan="JaxDeepReactivePolicy", method_kwargs="{'topology': [LAYER1_TUNE, LAYER1_TUNE]}",
                              extra="rollout_horizon=ROLLOUT_HORIZON_TUNE\n", seconds=3)
    for tag, val in (("MODEL_WEIGHT_TUNE", "10.0"), ("LEARNING_RATE_TUNE", "0.01"), ("LAYER1_TUNE", "32"),
                     ("ROLLOUT_HORIZON_TUNE", "5")):
        cfg = cfg.replace(tag, val)
    pa, pk, ta = load_config_from_string(cfg)
    assert pa["rollout_horizon"] == 5 and pk["topology"] == [32, 32] and ta["train_seconds"] == 3
