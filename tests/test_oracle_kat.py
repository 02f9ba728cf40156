"""Known-answer tests that pin the CPU oracle to the reference's closed forms.

Each case evaluates one RDDL expression through oracle/model_eval.py and compares value and
gradient with the formula written in the reference source (file:line in each case; SURVEY
Appendix A), computed independently here in float64 numpy.  These are what stands between the
oracle and the real reference, which cannot be executed offline."""
import numpy as np
import pytest
import torch

from oracle.model_eval import OracleModel
from pyrddlgym_jax_b200.core import logic
from pyrddlgym_jax_b200.rddl.model import load_model

TEMPLATE = """
domain kat {{
    types {{ obj : object; }};
    pvariables {{
        C(obj) : {{ non-fluent, real, default = 0.5 }};
        x(obj) : {{ state-fluent, real, default = 0.0 }};
        y(obj) : {{ state-fluent, real, default = 0.0 }};
        a(obj) : {{ action-fluent, real, default = 0.0 }};
    }};
    cpfs {{
        x'(?o) = {expr};
        y'(?o) = y(?o);
    }};
    reward = sum_{{?o : obj}} [ x'(?o) ];
}}
non-fluents nf {{ domain = kat; objects {{ obj : {{o1, o2, o3, o4}}; }}; non-fluents {{ C(o2) = 1.5; }}; }}
instance inst {{ domain = kat; non-fluents = nf; init-state {{ x(o1) = 0.0; }}; max-nondef-actions = pos-inf; horizon = 3; discount = 1.0; }}
"""

X = np.array([[0.3, -1.2, 2.5, 0.51], [1.7, 0.49, -0.2, 3.9]])
Y = np.array([[0.1, 0.4, 2.6, -0.3], [1.7, -2.0, 0.8, 4.4]])
U = np.array([[0.2, 0.9, 0.55, 0.4], [0.7, 0.1, 0.3, 0.95]])


def sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def evaluate(expr, relax_kw=None, cls=logic.DefaultJaxRDDLCompilerWithGrad, noise=None,
             exact=False):
    """-> (value [B,4], dvalue/dx [B,4], dvalue/dy [B,4]) of x' for one step."""
    m = load_model(TEMPLATE.format(expr=expr))
    relax = None if exact else cls(m, **(relax_kw or {})).relax_config()
    om = OracleModel(m, relax)
    B = X.shape[0]
    fls, nfls = om.init_fls_nfls(B)
    x = torch.tensor(X, dtype=torch.float32, requires_grad=True)
    y = torch.tensor(Y, dtype=torch.float32, requires_grad=True)
    fls = dict(fls)
    fls["x"], fls["y"] = x, y
    act = {"a": torch.zeros(B, 4)}
    out, _, err, _ = om.step(fls, nfls, act, noise, B)
    v = out["x"]
    if exact or not v.requires_grad:
        return v.detach().numpy(), None, None
    gx, gy = torch.autograd.grad(v.sum(), [x, y], allow_unused=True)
    z = lambda g: np.zeros_like(X) if g is None else g.numpy()   # noqa: E731
    return v.detach().numpy(), z(gx), z(gy)


W = 3.0
CASES = [
    # expr, soft value f(x,y), d/dx, d/dy, hard value, kwargs (weight), STE flag name
    ("x(?o) > y(?o)", lambda x, y: sigmoid(W * (x - y)),                      # logic.py:269-272
     lambda x, y: W * sigmoid(W * (x - y)) * (1 - sigmoid(W * (x - y))),
     lambda x, y: -W * sigmoid(W * (x - y)) * (1 - sigmoid(W * (x - y))),
     lambda x, y: (x > y) * 1.0, {"sigmoid_weight": W}, "use_sigmoid_ste"),
    ("x(?o) <= y(?o)", lambda x, y: sigmoid(W * (y - x)),                     # logic.py:311-314
     lambda x, y: -W * sigmoid(W * (y - x)) * (1 - sigmoid(W * (y - x))),
     lambda x, y: W * sigmoid(W * (y - x)) * (1 - sigmoid(W * (y - x))),
     lambda x, y: (x <= y) * 1.0, {"sigmoid_weight": W}, "use_sigmoid_ste"),
    ("x(?o) == y(?o)", lambda x, y: 1 - np.tanh(W * (y - x)) ** 2,            # logic.py:325-328
     lambda x, y: 2 * np.tanh(W * (y - x)) * (1 - np.tanh(W * (y - x)) ** 2) * W,
     lambda x, y: -2 * np.tanh(W * (y - x)) * (1 - np.tanh(W * (y - x)) ** 2) * W,
     lambda x, y: (x == y) * 1.0, {"sigmoid_weight": W}, "use_tanh_ste"),
    ("sgn[x(?o)]", lambda x, y: np.tanh(W * x),                               # logic.py:353-356
     lambda x, y: W * (1 - np.tanh(W * x) ** 2), lambda x, y: 0 * x,
     lambda x, y: np.sign(x), {"sigmoid_weight": W}, "use_tanh_ste"),
    ("floor[x(?o)]",                                                         # logic.py:801-804, 816
     lambda x, y: np.floor(x + .5) - 1 + .5 * (1 + np.tanh(W * (x - np.floor(x + .5))) / np.tanh(W / 2)),
     lambda x, y: .5 * W * (1 - np.tanh(W * (x - np.floor(x + .5))) ** 2) / np.tanh(W / 2),
     lambda x, y: 0 * x, lambda x, y: np.floor(x), {"floor_weight": W}, "use_floor_ste"),
    ("ceil[x(?o)]",                                                          # logic.py:827-830
     lambda x, y: -(np.floor(-x + .5) - 1 + .5 * (1 + np.tanh(W * (-x - np.floor(-x + .5))) / np.tanh(W / 2))),
     lambda x, y: .5 * W * (1 - np.tanh(W * (-x - np.floor(-x + .5))) ** 2) / np.tanh(W / 2),
     lambda x, y: 0 * x, lambda x, y: np.ceil(x), {"floor_weight": W}, "use_floor_ste"),
    ("round[x(?o)]",                                                         # logic.py:880-882, 894
     lambda x, y: np.floor(x) + .5 + .5 * np.tanh(W * (x - np.floor(x) - .5)) / np.tanh(W / 2),
     lambda x, y: .5 * W * (1 - np.tanh(W * (x - np.floor(x) - .5)) ** 2) / np.tanh(W / 2),
     lambda x, y: 0 * x, lambda x, y: np.round(x), {"round_weight": W}, "use_round_ste"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("ste", [False, True])
def test_relaxed_op_closed_form(case, ste):
    expr, soft, dx, dy, hard, kw, flag = case
    v, gx, gy = evaluate(expr, {**kw, flag: ste})
    want = hard(X, Y) if ste else soft(X, Y)
    np.testing.assert_allclose(v, want, rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(gx, dx(X, Y), rtol=2e-5, atol=2e-6)     # gradient is the soft one
    np.testing.assert_allclose(gy, dy(X, Y), rtol=2e-5, atol=2e-6)


def test_product_tnorm_closed_forms():
    """logic.py:441, 454, 467, 480, 493, 506 on soft truth values p = (x > 0), q = (y > 0)."""
    kw = {"sigmoid_weight": 1.0, "use_sigmoid_ste": False}
    p, q = sigmoid(X), sigmoid(Y)
    forms = {
        "~(x(?o) > 0.0)": 1 - p,
        "(x(?o) > 0.0) ^ (y(?o) > 0.0)": p * q,
        "(x(?o) > 0.0) | (y(?o) > 0.0)": 1 - (1 - p) * (1 - q),
        "(x(?o) > 0.0) => (y(?o) > 0.0)": 1 - p * (1 - q),
        "(x(?o) > 0.0) <=> (y(?o) > 0.0)": (1 - p * (1 - q)) * (1 - q * (1 - p)),
    }
    for expr, want in forms.items():
        v, _, _ = evaluate(expr, kw)
        np.testing.assert_allclose(v, want, rtol=3e-6, atol=1e-6, err_msg=expr)


def test_forall_exists_are_products():
    """logic.py:521, 534: forall = prod x, exists = 1 - prod (1 - x)."""
    kw = {"sigmoid_weight": 1.0, "use_sigmoid_ste": False}
    p = sigmoid(X)
    v, _, _ = evaluate("forall_{?u : obj} [ x(?u) > 0.0 ]", kw)
    np.testing.assert_allclose(v, np.repeat(p.prod(1, keepdims=True), 4, 1), rtol=3e-6)
    v, gx, _ = evaluate("exists_{?u : obj} [ x(?u) > 0.0 ]", kw)
    np.testing.assert_allclose(v, np.repeat(1 - (1 - p).prod(1, keepdims=True), 4, 1), rtol=3e-6)
    # d/dx_j of 4 copies of 1 - prod(1 - p_i) = 4 * prod_{i != j}(1 - p_i) * p_j (1 - p_j)
    others = np.stack([np.delete(1 - p, j, 1).prod(1) for j in range(4)], 1)
    np.testing.assert_allclose(gx, 4 * others * p * (1 - p), rtol=2e-5)


def test_godel_and_lukasiewicz_closed_forms():
    """logic.py:573-653 (min / max) and logic.py:692-744 (relu forms)."""
    class Godel(logic.SigmoidRelational, logic.GodelNormLogical):
        pass

    class Luka(logic.SigmoidRelational, logic.LukasiewiczNormLogical):
        pass
    kw = {"sigmoid_weight": 1.0, "use_sigmoid_ste": False}
    p, q = sigmoid(X), sigmoid(Y)
    relu = lambda z: np.maximum(z, 0)      # noqa: E731
    for cls, forms in ((Godel, {"^": np.minimum(p, q), "|": np.maximum(p, q)}),
                       (Luka, {"^": relu(p + q - 1), "|": 1 - relu(1 - p - q)})):
        for op, want in forms.items():
            v, _, _ = evaluate(f"(x(?o) > 0.0) {op} (y(?o) > 0.0)", kw, cls=cls)
            np.testing.assert_allclose(v, want, rtol=3e-6, atol=1e-6, err_msg=f"{cls.__name__} {op}")


@pytest.mark.parametrize("ste", [False, True])
def test_linear_if_else(ste):
    """logic.py:933-936: c' = c + sg(1[c > 0.5] - c); value c' a + (1 - c') b."""
    kw = {"sigmoid_weight": W, "use_sigmoid_ste": False, "use_if_else_ste": ste}
    v, gx, gy = evaluate("if (x(?o) > 0.0) then y(?o) * 2.0 else -y(?o)", kw)
    c = sigmoid(W * X)
    cv = (c > 0.5) * 1.0 if ste else c
    np.testing.assert_allclose(v, cv * (2 * Y) + (1 - cv) * (-Y), rtol=3e-6, atol=2e-6)
    np.testing.assert_allclose(gx, W * c * (1 - c) * (2 * Y + Y), rtol=2e-5, atol=1e-5)  # fp32 s(1-s)
    np.testing.assert_allclose(gy, cv * 2 - (1 - cv), rtol=2e-5, atol=2e-6)


def test_exact_if_thresholds_at_half_and_fold_order():
    """compiler.py:1640 (if = where(pred > 0.5)), compiler.py:1140-1143 (left fold)."""
    v, _, _ = evaluate("if (x(?o) > y(?o)) then 1.0 else -1.0", exact=True)
    np.testing.assert_array_equal(v, np.where(X > Y, 1.0, -1.0).astype(np.float32))
    v, _, _ = evaluate("x(?o) + y(?o) + C(?o) + 1e8 - 1e8", exact=True)
    f = np.float32
    want = ((((f(X) + f(Y)) + f(np.array([0.5, 1.5, 0.5, 0.5]))) + f(1e8)) - f(1e8))
    np.testing.assert_array_equal(v, want)


@pytest.mark.parametrize("ste", [False, True])
def test_sigmoid_bernoulli(ste):
    """logic.py:1125-1129: sigma(w (p - U)); STE value 1[p > U]."""
    kw = {"bernoulli_sigmoid_weight": W, "use_ste_bernoulli": ste}
    noise = [torch.tensor(U, dtype=torch.float32)]
    v, gx, _ = evaluate("Bernoulli(1.0 / (1.0 + exp[-x(?o)]))", kw, noise=noise)
    p = sigmoid(X)
    s = sigmoid(W * (p - U))
    np.testing.assert_allclose(v, (p > U) * 1.0 if ste else s, rtol=3e-6, atol=1e-6)
    np.testing.assert_allclose(gx, W * s * (1 - s) * p * (1 - p), rtol=3e-5, atol=1e-6)


def test_gumbel_sigmoid_bernoulli():
    """logic.py:1177-1179: z = logit(clip(p)) + Logistic; sigma(w z)."""
    class G(logic.SigmoidRelational, logic.ProductNormLogical, logic.LinearIfElse,
            logic.GumbelSigmoidBernoulli):
        pass
    lg = np.log(U) - np.log1p(-U)
    noise = [torch.tensor(lg, dtype=torch.float32)]
    v, gx, _ = evaluate("Bernoulli(1.0 / (1.0 + exp[-x(?o)]))", {"bernoulli_sigmoid_weight": W},
                        cls=G, noise=noise)
    p = sigmoid(X)
    z = np.log(p) - np.log1p(-p) + lg
    s = sigmoid(W * z)
    np.testing.assert_allclose(v, s, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(gx, W * s * (1 - s), rtol=1e-4, atol=1e-6)   # dz/dx = 1


def test_exact_bernoulli_and_uniform_normal():
    """compiler.py:1874 (U < p), 1793-1794 (a + (b-a) U), 1813-1816 (m + sqrt(v) Z)."""
    noise = [torch.tensor(U, dtype=torch.float32)]
    v, _, _ = evaluate("Bernoulli(C(?o) * 0.5)", exact=True, noise=noise)
    np.testing.assert_array_equal(v, (U < np.array([0.25, 0.75, 0.25, 0.25])).astype(np.float32))
    v, _, _ = evaluate("Uniform(x(?o), x(?o) + 2.0)", exact=True, noise=noise)
    np.testing.assert_allclose(v, X + 2 * U, rtol=1e-6, atol=1e-6)
    v, _, _ = evaluate("Normal(x(?o), 4.0)", exact=True, noise=noise)
    np.testing.assert_allclose(v, X + 2 * U, rtol=1e-6, atol=1e-6)


def test_aggregations_and_functions_exact():
    """compiler.py:1303-1348, 1421-1529."""
    v, _, _ = evaluate("(sum_{?u : obj} [ x(?u) ]) + (max_{?u : obj} [ y(?u) ]) * 0.0 "
                       "+ (prod_{?u : obj} [ C(?u) ])", exact=True)
    want = X.sum(1, keepdims=True) + 0.5 * 1.5 * 0.5 * 0.5
    np.testing.assert_allclose(v, np.repeat(want, 4, 1), rtol=1e-6)
    v, _, _ = evaluate("pow[x(?o), 2] + abs[y(?o)] + min[x(?o), y(?o)] + hypot[x(?o), y(?o)]",
                       exact=True)
    np.testing.assert_allclose(v, X ** 2 + np.abs(Y) + np.minimum(X, Y) + np.hypot(X, Y), rtol=2e-6)
    v, _, _ = evaluate("log[exp[x(?o)] + 1.0, 2.0] + sqrt[abs[y(?o)]] + mod[x(?o), 0.7]",
                       exact=True)
    np.testing.assert_allclose(v, np.log(np.exp(X) + 1) / np.log(2.0) + np.sqrt(np.abs(Y))
                               + np.mod(X, 0.7), rtol=3e-6, atol=1e-6)


def test_gradient_finite_differences_through_a_rollout():
    """Central differences (fp64 inputs, fp32 model) of the oracle's reservoir return."""
    from oracle import planner as OP
    from .helpers import fixture_model
    from .parity import make_params
    m = fixture_model("reservoir")
    comp = logic.DefaultJaxRDDLCompilerWithGrad(m, use_sigmoid_ste=False, use_if_else_ste=False,
                                                sigmoid_weight=0.2)
    om = OracleModel(m, comp.relax_config())
    T, B = 3, 2
    params = make_params(m, "reservoir", T)
    torch.manual_seed(0)
    noise = [[torch.rand(B, 3) * 4 + 1] for _ in range(T)]

    def loss_of(p):
        out = OP.rollout(om, lambda t, fls: OP.slp_train_actions(m, p, t), B, T, noise)
        return OP.mean_loss(out["reward"], 1.0)
    P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    loss_of(P).backward()
    g = P["release"].grad.numpy()
    eps = 2e-2
    for (t, i) in [(0, 0), (1, 2), (2, 1)]:
        hi = {k: v.clone() for k, v in params.items()}
        lo = {k: v.clone() for k, v in params.items()}
        hi["release"][t, i] += eps
        lo["release"][t, i] -= eps
        fd = (float(loss_of(hi)) - float(loss_of(lo))) / (2 * eps)
        assert abs(fd - g[t, i]) <= 2e-2 * max(1.0, abs(g[t, i])), (t, i, fd, g[t, i])


def test_optimizers_known_answers():
    """optax rmsprop / adam closed forms + clip_by_global_norm."""
    from oracle import planner as OP
    p, g = torch.tensor([1.0, -2.0]), torch.tensor([0.5, -4.0])
    p1, nu = OP.rmsprop_step(p, g, torch.zeros(2), lr=0.1)
    nu_w = 0.1 * np.array([0.25, 16.0])
    np.testing.assert_allclose(nu.numpy(), nu_w, rtol=1e-6)
    np.testing.assert_allclose(p1.numpy(), np.array([1.0, -2.0]) - 0.1 * np.array([0.5, -4.0])
                               / np.sqrt(nu_w + 1e-8), rtol=1e-6)
    p1, mu, nu = OP.adam_step(p, g, torch.zeros(2), torch.zeros(2), 1, lr=0.1)
    # first adam step moves every coordinate by ~lr against the gradient sign
    np.testing.assert_allclose(p1.numpy(), np.array([1.0 - 0.1, -2.0 + 0.1]), rtol=1e-5)
    (c,) = OP.clip_by_global_norm([g], 1.0)
    np.testing.assert_allclose(c.numpy(), g.numpy() / np.sqrt(0.25 + 16.0), rtol=1e-6)


def test_projection_known_answers():
    """planner.py:499-527 (sorting) and 540-616 (SOGBOFA) on hand-checkable rows."""
    from oracle import planner as OP
    noop = {"a": False}
    box = {"a": (-10.0, 10.0)}
    row = {"a": torch.tensor([3.0, -1.0, 2.0, 0.5])}
    new, _ = OP.sorting_project_row(row, noop, 1, True, 0.0, box)      # 2nd largest = 2.0
    np.testing.assert_allclose(new["a"].numpy(), [1.0, -3.0, 0.0, -1.5])
    new, _ = OP.sorting_project_row(row, noop, 3, True, 0.0, box)      # 4th largest = -1 < 0
    np.testing.assert_allclose(new["a"].numpy(), [3.0, -1.0, 2.0, 0.5])
    acts = {"a": torch.tensor([0.9, 0.8, 0.1, 0.0])}                   # sum 1.8, allowed 1
    new, ok = OP.sogbofa_project_row(acts, noop, 1, 100, 1e-6, 1 - 1e-6)
    assert ok
    # surplus .8 over k=3 -> [.6333, .5333, 0, 0]: sum 1.1667, k=2 -> -.0833 each -> [.55, .45]
    np.testing.assert_allclose(float(new["a"].sum()), 1.0, atol=2e-6)
    # only one action may stay above 0.5: the first active one is reset to 0.5
    assert int((new["a"] > 0.5).sum()) <= 1


SWITCH_DOMAIN = """
domain sw {
    types { obj : object; gear : {@a, @b, @c}; };
    pvariables {
        x(obj) : { state-fluent, real, default = 0.0 };
        g(obj) : { state-fluent, gear, default = @a };
        u(obj) : { action-fluent, real, default = 0.0 };
    };
    cpfs {
        x'(?o) = switch (g(?o)) { case @a : 1.0 + x(?o), case @c : 5.0 * x(?o), default : -2.0 };
        g'(?o) = g(?o);
    };
    reward = sum_{?o : obj} [ x'(?o) ];
}
non-fluents nf { domain = sw; objects { obj : {o1, o2, o3}; }; non-fluents { }; }
instance inst { domain = sw; non-fluents = nf; init-state { x(o1) = 0.5; x(o2) = 0.25; x(o3) = 2.0; g(o2) = @b; g(o3) = @c; }; max-nondef-actions = pos-inf; horizon = 2; discount = 1.0; }
"""


@pytest.mark.parametrize("ste", [True, False])
def test_switch_exact_and_softmax(ste):
    """compiler.py:1661-1684 (take_along_axis over the stacked cases) and logic.py:983-1002
    (cases weighted by softmax_k(-w |pred - k|), straight-through to the hard case)."""
    m = load_model(SWITCH_DOMAIN)
    x0 = np.array([0.5, 0.25, 2.0])
    cases = np.stack([1.0 + x0, np.full(3, -2.0), 5.0 * x0])          # [K, obj]
    pred = np.array([0, 1, 2])
    hard = cases[pred, np.arange(3)]
    om = OracleModel(m, None)
    fls, nfls = om.init_fls_nfls(2)
    out, _, _, _ = om.step(dict(fls), nfls, {"u": torch.zeros(2, 3)}, None, 2)
    np.testing.assert_allclose(out["x"].numpy(), np.stack([hard, hard]), rtol=1e-6)
    w = 2.0
    relax = logic.DefaultJaxRDDLCompilerWithGrad(m, switch_weight=w, use_switch_ste=ste).relax_config()
    om = OracleModel(m, relax)
    fls, nfls = om.init_fls_nfls(2)
    out, _, _, _ = om.step(dict(fls), nfls, {"u": torch.zeros(2, 3)}, None, 2)
    logits = -w * np.abs(pred[None, :] - np.arange(3)[:, None])
    sm = np.exp(logits) / np.exp(logits).sum(0, keepdims=True)
    soft = (sm * cases).sum(0)
    np.testing.assert_allclose(out["x"].detach().numpy()[0], hard if ste else soft, rtol=1e-5)


def test_count_and_gamma_family_sampler_moments():
    """The exact samplers fed with supplied noise reproduce the laws the reference draws from
    (compiler.py:1880-2147): Poisson / Binomial means and variances, and the means of Pareto,
    Student, ChiSquare and Beta through the fixture's shock term."""
    from .parity import forward_case
    B = 6000
    c = forward_case("emu", "distributions", B, 1, False, seed=7)
    fin = c["final"]
    level0 = np.array([0.2, 1.0, 1.8])
    lam = 0.5 + level0 ** 2
    q = fin["queue"].numpy().astype(np.float64)
    np.testing.assert_allclose(q.mean(0), lam, rtol=0.06)
    np.testing.assert_allclose(q.var(0), lam, rtol=0.10)
    p = 0.2 + 0.2 * np.minimum(level0, 3.0)
    s = fin["served"].numpy().astype(np.float64)
    assert s.min() >= 0 and s.max() <= 5
    np.testing.assert_allclose(s.mean(0), 5 * p, rtol=0.04)
    np.testing.assert_allclose(s.var(0), 5 * p * (1 - p), rtol=0.10)
    # NegativeBinomial(r, p) of the exact model: tfp's convention, mean r p / (1 - p)
    pn = 0.3 + 0.1 * np.minimum(level0, 3.0)
    nb = fin["backlog"].numpy().astype(np.float64)
    np.testing.assert_allclose(nb.mean(0), 3 * pn / (1 - pn), rtol=0.08)
    np.testing.assert_allclose(nb.var(0), 3 * pn / (1 - pn) ** 2, rtol=0.15)
    np.testing.assert_array_equal(nb, c["ref"]["final"]["backlog"].numpy())
    # the emulated program and the oracle agree draw for draw
    np.testing.assert_array_equal(q, c["ref"]["final"]["queue"].numpy())
    np.testing.assert_array_equal(s, c["ref"]["final"]["served"].numpy())


def test_gamma_family_closed_forms():
    """ChiSquare = 2 Gamma(df/2), Student = Z sqrt((df/2)/Gamma(df/2)), Beta = Ga/(Ga+Gb),
    Pareto = scale exp(Exp(1)/shape): each evaluated on hand-written noise."""
    Z = torch.tensor(U * 2 - 1, dtype=torch.float32)
    G = torch.tensor(U * 3 + 0.2, dtype=torch.float32)
    G2 = torch.tensor(Y ** 2 + 0.1, dtype=torch.float32)
    v, _, _ = evaluate("ChiSquare(3.0)", exact=True, noise=[G[:, 0]])     # scalar draw per rollout
    np.testing.assert_allclose(v, np.repeat(2 * G[:, :1].numpy(), 4, 1), rtol=1e-6)
    v, _, _ = evaluate("Student(C(?o) * 8.0)", exact=True, noise=[Z, G])
    df = np.array([4.0, 12.0, 4.0, 4.0])
    np.testing.assert_allclose(v, Z.numpy() * np.sqrt(0.5 * df / G.numpy()), rtol=1e-5)
    v, _, _ = evaluate("Beta(C(?o), 2.0)", exact=True, noise=[G, G2])
    np.testing.assert_allclose(v, G.numpy() / (G.numpy() + G2.numpy()), rtol=1e-6)
    v, _, _ = evaluate("Pareto(C(?o) * 4.0, x(?o) + 3.0)", exact=True, noise=[G])
    np.testing.assert_allclose(v, (X + 3.0) * np.exp(G.numpy() / np.array([2.0, 6.0, 2.0, 2.0])),
                               rtol=2e-6)


def test_fluent_shape_gamma_sampler_moments():
    """The in-step Gamma sampler (Marsaglia-Tsang on supplied normal / uniform draws, used when the
    shape is a fluent) reproduces the Gamma law: mean and variance of Gamma(alpha, scale) for shapes
    below and above one, the mean of Beta and of the Gamma-Poisson mixture, through the
    Gamma_Fluent_Toy fixture's first step; the emulated program and the oracle agree draw for draw."""
    from .parity import forward_case
    B = 20000
    c = forward_case("emu", "gammafluent", B, 1, False, seed=11)
    fin = c["final"]
    level0 = np.array([0.3, 1.0, 1.7])
    rate = np.array([0.5, 0.8, 0.5])
    feed = np.clip(c["params"]["feed"].numpy()[0], 0.0, 1.0)            # test policy clips to the box
    alpha = 0.4 + level0 ** 2 + feed
    # level' = 0.7 level + 0.1 inflow + 0.2 mix - 0.01 count, inflow = Gamma + 0.05 Chi2 + 0.02 Student
    lv = fin["level"].numpy().astype(np.float64)
    inflow = (lv - 0.7 * level0 - 0.2 * 0.5) / 0.1
    want_mean = alpha * rate + 0.05 * (1.0 + level0)                    # E Student = 0 (df > 1)
    want_var = alpha * rate ** 2 + 0.05 ** 2 * 2 * (1.0 + level0) \
        + 0.02 ** 2 * (3 + level0 ** 2) / (1 + level0 ** 2)
    np.testing.assert_allclose(inflow.mean(0), want_mean, rtol=0.03)
    np.testing.assert_allclose(inflow.var(0), want_var, rtol=0.08)
    mix = fin["mix"].numpy().astype(np.float64).reshape(-1)
    a, b = 0.5 + 0.5, 1.0 + level0.sum()
    assert abs(mix.mean() - a / (a + b)) < 0.01 and mix.min() > 0 and mix.max() < 1
    cnt = fin["count"].numpy().astype(np.float64)
    r, p = 1.5 + level0, 0.4
    np.testing.assert_allclose(cnt.mean(0), r * p / (1 - p), rtol=0.06)
    np.testing.assert_array_equal(cnt, c["ref"]["final"]["count"].numpy())
    np.testing.assert_allclose(lv, c["ref"]["final"]["level"].numpy(), rtol=1e-5, atol=1e-5)


def _soft_floor(x, w):
    """logic.py:801-804."""
    r = np.floor(x + 0.5)
    return r - 1 + 0.5 * (1 + np.tanh(w * (x - r)) / np.tanh(w / 2))


def test_remaining_relational_and_soft_division_closed_forms():
    """`~=`, `<`, `>=` (logic.py:283, 297, 339) and soft `div` / `mod` (logic.py:841, 855-858):
    soft_floor of the quotient, x - y * soft_floor(x / y)."""
    v, gx, gy = evaluate("x(?o) ~= y(?o)", {"sigmoid_weight": W, "use_tanh_ste": False})
    t = np.tanh(W * (Y - X))
    np.testing.assert_allclose(v, t ** 2, rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(gy, 2 * t * (1 - t ** 2) * W, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(gx, -2 * t * (1 - t ** 2) * W, rtol=2e-5, atol=2e-6)
    for expr, z in (("x(?o) < y(?o)", Y - X), ("x(?o) >= y(?o)", X - Y)):
        v, gx, _ = evaluate(expr, {"sigmoid_weight": W, "use_sigmoid_ste": False})
        np.testing.assert_allclose(v, sigmoid(W * z), rtol=2e-6, atol=2e-6)
        sign = -1.0 if "<" in expr else 1.0
        np.testing.assert_allclose(gx, sign * W * sigmoid(W * z) * (1 - sigmoid(W * z)), rtol=2e-5,
                                   atol=2e-6)
    yy = "(2.0 + y(?o) * y(?o))"                                  # a positive divisor
    D = 2.0 + Y ** 2
    v, gx, _ = evaluate(f"div[x(?o), {yy}]", {"floor_weight": W, "use_floor_ste": False})
    np.testing.assert_allclose(v, _soft_floor(X / D, W), rtol=2e-6, atol=2e-6)
    q = X / D
    dq = 0.5 * W * (1 - np.tanh(W * (q - np.floor(q + 0.5))) ** 2) / np.tanh(W / 2)
    np.testing.assert_allclose(gx, dq / D, rtol=2e-5, atol=2e-6)
    v, _, _ = evaluate(f"mod[x(?o), {yy}]", {"floor_weight": W, "use_floor_ste": False})
    np.testing.assert_allclose(v, X - D * _soft_floor(X / D, W), rtol=2e-6, atol=2e-5)
    v, _, _ = evaluate(f"div[x(?o), {yy}]", {"floor_weight": W, "use_floor_ste": True})
    np.testing.assert_allclose(v, np.floor(X / D), rtol=0, atol=1e-6)


def test_geometric_reparameterised_closed_form():
    """ReparameterizedGeometric (logic.py:1046-1047): 1 + soft_floor(-E / safe_log(1 - p), w) with an
    Exp(1) draw E; the exact model floors log U / log(1 - p) + 1 (compiler.py:2001-2003)."""
    Ex = torch.tensor(-np.log(U), dtype=torch.float32)           # Exp(1) draws from the uniforms
    p = "(0.2 + 0.1 * x(?o) * x(?o) / (1.0 + x(?o) * x(?o)))"
    P = 0.2 + 0.1 * X ** 2 / (1 + X ** 2)
    v, gx, _ = evaluate(f"Geometric({p})", {"geometric_floor_weight": W}, noise=[Ex])
    ratio = -Ex.numpy().astype(np.float64) / np.log(1 - P)
    np.testing.assert_allclose(v, 1 + _soft_floor(ratio, W), rtol=1e-5, atol=1e-5)
    assert np.abs(gx).max() > 0
    Un = torch.tensor(U, dtype=torch.float32)
    v, _, _ = evaluate(f"Geometric({p})", exact=True, noise=[Un])
    np.testing.assert_array_equal(v, np.floor(np.log(U) / np.log(1 - P)) + 1)


def test_location_scale_and_power_law_samplers_closed_forms():
    """Gumbel / Laplace / Cauchy are location + scale * standard draw (compiler.py:2060-2103);
    Gompertz(s, r) = ln(1 + E) / s / r with E ~ Exp(1) as coded at compiler.py:2123-2124; Kumaraswamy(a, b) =
    (1 - U^(1/b))^(1/a) (compiler.py:2163-2164); Weibull = scale * E^(1/shape) (compiler.py:1857)."""
    Z = torch.tensor(Y, dtype=torch.float32)                      # any standard draws
    for name in ("Gumbel", "Laplace", "Cauchy"):
        v, _, _ = evaluate(f"{name}(C(?o), 2.0)", exact=True, noise=[Z])
        loc = np.array([0.5, 1.5, 0.5, 0.5])
        np.testing.assert_allclose(v, loc + 2.0 * Y, rtol=1e-6, atol=1e-6)
    Ex = torch.tensor(-np.log(U), dtype=torch.float32)
    E64 = Ex.numpy().astype(np.float64)
    shape = np.array([1.0, 3.0, 1.0, 1.0])
    # the draw takes the shape of the SCALE argument (compiler.py:2121)
    v, _, _ = evaluate("Gompertz(3.0, C(?o) * 2.0)", exact=True, noise=[Ex])
    # the CODE computes log(1 + E) / shape / scale (its comment states ln(1 + E / s) / r): follow the code
    np.testing.assert_allclose(v, np.log(1 + E64) / 3.0 / shape, rtol=1e-5, atol=1e-6)
    Un = torch.tensor(U, dtype=torch.float32)
    v, _, _ = evaluate("Kumaraswamy(C(?o) * 2.0, 3.0)", exact=True, noise=[Un])
    np.testing.assert_allclose(v, (1 - U ** (1 / 3.0)) ** (1 / shape), rtol=1e-5, atol=1e-6)
    v, _, _ = evaluate("Weibull(C(?o) * 4.0, 1.5)", exact=True, noise=[Ex])
    np.testing.assert_allclose(v, 1.5 * E64 ** (1 / (2 * shape)), rtol=1e-5, atol=1e-6)


def test_exponential_poisson_relaxation_closed_form():
    """ExponentialPoisson (logic.py:1508-1531): sum_k sigmoid(w (1 - cumsum_k(E_k / rate))) over
    nbins Exp(1) draws while cdf(nbins, rate) >= min_cdf, rate + sqrt(rate) z above it; value and the
    derivative w.r.t. the rate."""
    K, w = 6, 4.0
    g = np.random.default_rng(3)
    Es = [torch.tensor(g.exponential(size=X.shape), dtype=torch.float32) for _ in range(K)]
    Zn = torch.tensor(g.normal(size=X.shape), dtype=torch.float32)
    rate_expr = "(0.3 + 0.2 * x(?o) * x(?o))"
    lam = 0.3 + 0.2 * X ** 2                                     # <= 3.4: cdf(6, 3.4) = 0.94 < 0.999 there
    v, gx, _ = evaluate(f"Poisson({rate_expr})", {"poisson_nbins": K, "poisson_comparison_weight": w},
                        noise=Es + [Zn])
    from scipy import stats
    small = stats.poisson.cdf(K, lam) >= 0.999
    assert small.any() and (~small).any()                        # both branches are exercised
    times = np.cumsum([e.numpy().astype(np.float64) / lam for e in Es], axis=0)
    s = sigmoid(w * (1 - times))
    soft = s.sum(0)
    normal = lam + np.sqrt(lam) * Zn.numpy()
    np.testing.assert_allclose(v, np.where(small, soft, normal), rtol=1e-5, atol=1e-5)
    # d soft / d lam = sum_k s (1 - s) w times_k / lam ; d normal / d lam = 1 + z / (2 sqrt lam)
    dsoft = (s * (1 - s) * w * times / lam).sum(0)
    dnorm = 1 + Zn.numpy() / (2 * np.sqrt(lam))
    dlam_dx = 0.4 * X
    np.testing.assert_allclose(gx, np.where(small, dsoft, dnorm) * dlam_dx, rtol=2e-4, atol=1e-5)


def test_gumbel_softmax_binomial_relaxation_closed_form():
    """GumbelSoftmaxBinomial (logic.py:1381-1434): soft arg-max, sum_k k softmax(w (g_k + log pmf_k)),
    over nbins bins of Gumbel noise plus the Binomial log-pmf (bins above n masked out), for n < nbins."""
    from scipy import special

    class Comp(logic.SigmoidRelational, logic.ProductNormLogical, logic.LinearIfElse,
               logic.GumbelSoftmaxBinomial):
        pass
    K, w, n = 8, 3.0, 5
    g = np.random.default_rng(5)
    G = torch.tensor(g.gumbel(size=X.shape + (K,)), dtype=torch.float32)
    Zn = torch.tensor(g.normal(size=X.shape[:1]), dtype=torch.float32)   # shaped like the trials argument
    P = 0.2 + 0.6 * sigmoid(X)
    v, gx, _ = evaluate(f"Binomial({n}, 0.2 + 0.6 / (1.0 + exp[-x(?o)]))",
                        {"binomial_nbins": K, "binomial_softmax_weight": w}, cls=Comp, noise=[G, Zn])
    ks = np.arange(K)
    logpmf = (special.gammaln(n + 1) - special.gammaln(np.minimum(ks, n) + 1)
              - special.gammaln(n - np.minimum(ks, n) + 1)
              + np.minimum(ks, n) * np.log(P[..., None]) + (n - np.minimum(ks, n)) * np.log(1 - P[..., None]))
    logpmf = np.where(ks <= n, logpmf, -np.inf)
    z = w * (G.numpy().astype(np.float64) + logpmf)
    sm = np.exp(z - z.max(-1, keepdims=True))
    sm /= sm.sum(-1, keepdims=True)
    np.testing.assert_allclose(v, (ks * sm).sum(-1), rtol=2e-5, atol=2e-5)
    # d/dp: softmax chain through d logpmf_k / dp = k / p - (n - k) / (1 - p)
    dl = np.where(ks <= n, ks / P[..., None] - (n - ks) / (1 - P[..., None]), 0.0)
    mean_k = (ks * sm).sum(-1, keepdims=True)
    dv_dp = (w * sm * (ks - mean_k) * dl).sum(-1)
    dp_dx = 0.6 * sigmoid(X) * (1 - sigmoid(X))
    np.testing.assert_allclose(gx, dv_dp * dp_dx, rtol=2e-4, atol=2e-5)


DISCRETE_DOMAIN = """
domain dk {
    types { obj : object; lvl : {@low, @mid, @high}; };
    pvariables {
        x(obj) : { state-fluent, real, default = 0.0 };
        m(obj) : { state-fluent, lvl, default = @low };
        u(obj) : { action-fluent, real, default = 0.0 };
    };
    cpfs {
        m'(?o) = Discrete(lvl, @low : 0.2 + 0.3 * x(?o), @mid : 0.5 - 0.3 * x(?o), @high : 0.3);
        x'(?o) = x(?o);
    };
    reward = sum_{?o : obj} [ x'(?o) ];
}
non-fluents nf { domain = dk; objects { obj : {o1, o2, o3}; }; non-fluents { }; }
instance inst { domain = dk; non-fluents = nf; init-state { x(o1) = 0.5; x(o2) = 0.25; x(o3) = 1.0; }; max-nondef-actions = pos-inf; horizon = 2; discount = 1.0; }
"""


def test_gumbel_softmax_discrete_closed_form():
    """GumbelSoftmaxDiscrete (logic.py:1258-1267): soft arg-max over the cases of Gumbel noise plus
    safe_log of the case probabilities (no straight-through); the exact model takes the arg-max of
    the same perturbed logits (compiler.py:2212)."""
    m = load_model(DISCRETE_DOMAIN)
    x0 = np.array([0.5, 0.25, 1.0])
    prob = np.stack([0.2 + 0.3 * x0, 0.5 - 0.3 * x0, np.full(3, 0.3)], -1)      # [obj, K]
    g = np.random.default_rng(9)
    G = torch.tensor(g.gumbel(size=(2, 3, 3)), dtype=torch.float32)
    w = 2.5
    relax = logic.DefaultJaxRDDLCompilerWithGrad(m, discrete_softmax_weight=w).relax_config()
    om = OracleModel(m, relax)
    fls, nfls = om.init_fls_nfls(2)
    out, _, _, _ = om.step(dict(fls), nfls, {"u": torch.zeros(2, 3)}, [G], 2)
    z = w * (G.numpy().astype(np.float64) + np.log(prob)[None])
    sm = np.exp(z - z.max(-1, keepdims=True))
    sm /= sm.sum(-1, keepdims=True)
    np.testing.assert_allclose(out["m"].detach().numpy(), (np.arange(3) * sm).sum(-1), rtol=1e-5,
                               atol=1e-5)
    om = OracleModel(m, None)
    fls, nfls = om.init_fls_nfls(2)
    out, _, _, _ = om.step(dict(fls), nfls, {"u": torch.zeros(2, 3)}, [G], 2)
    hard = np.argmax(G.numpy().astype(np.float64) + np.log(prob)[None], -1)
    np.testing.assert_array_equal(out["m"].numpy(), hard)


ARGMAX_DOMAIN = """
domain ak {
    types { obj : object; };
    pvariables {
        x(obj)   : { state-fluent, real, default = 0.0 };
        hi(obj)  : { state-fluent, obj, default = $o1 };
        lo(obj)  : { state-fluent, obj, default = $o1 };
        u(obj)   : { action-fluent, real, default = 0.0 };
    };
    cpfs {
        hi'(?o) = argmax_{?p : obj} [ x(?p) * x(?o) ];
        lo'(?o) = argmin_{?p : obj} [ x(?p) * x(?o) ];
        x'(?o) = x(?o);
    };
    reward = sum_{?o : obj} [ x'(?o) ];
}
non-fluents nf { domain = ak; objects { obj : {o1, o2, o3, o4}; }; non-fluents { }; }
instance inst { domain = ak; non-fluents = nf; init-state { x(o1) = 0.5; x(o2) = -0.25; x(o3) = 1.0; x(o4) = 0.8; }; max-nondef-actions = pos-inf; horizon = 2; discount = 1.0; }
"""


@pytest.mark.parametrize("ste", [True, False])
def test_softmax_argmax_closed_form(ste):
    """SoftmaxArgmax (logic.py:377-417): sum_i i softmax(w x)_i over the aggregated axis, arg-min on
    -x; straight-through to the exact index; the exact model returns the index itself."""
    m = load_model(ARGMAX_DOMAIN)
    x0 = np.array([0.5, -0.25, 1.0, 0.8])
    body = x0[None, :] * x0[:, None]                      # [o, p]
    om = OracleModel(m, None)
    fls, nfls = om.init_fls_nfls(2)
    out, _, _, _ = om.step(dict(fls), nfls, {"u": torch.zeros(2, 4)}, None, 2)
    np.testing.assert_array_equal(out["hi"].numpy()[0], body.argmax(1))
    np.testing.assert_array_equal(out["lo"].numpy()[0], body.argmin(1))
    w = 3.0
    relax = logic.DefaultJaxRDDLCompilerWithGrad(m, argmax_weight=w, use_argmax_ste=ste).relax_config()
    om = OracleModel(m, relax)
    fls, nfls = om.init_fls_nfls(2)
    out, _, _, _ = om.step(dict(fls), nfls, {"u": torch.zeros(2, 4)}, None, 2)

    def soft(z):
        e = np.exp(w * z - (w * z).max(1, keepdims=True))
        return (np.arange(4) * e / e.sum(1, keepdims=True)).sum(1)
    want_hi = body.argmax(1) if ste else soft(body)
    want_lo = body.argmin(1) if ste else soft(-body)
    np.testing.assert_allclose(out["hi"].detach().numpy()[0], want_hi, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["lo"].detach().numpy()[0], want_lo, rtol=1e-5, atol=1e-6)


def test_gumbel_softmax_poisson_relaxation_closed_form():
    """GumbelSoftmaxPoisson (logic.py:1618-1641): soft arg-max over nbins bins of Gumbel noise plus the
    Poisson log-pmf k log(rate) - rate - lgamma(k + 1), below the rate where cdf(nbins) >= min_cdf."""
    from scipy import special, stats

    class Comp(logic.SigmoidRelational, logic.ProductNormLogical, logic.LinearIfElse,
               logic.GumbelSoftmaxPoisson):
        pass
    K, w = 9, 2.0
    g = np.random.default_rng(8)
    G = torch.tensor(g.gumbel(size=X.shape + (K,)), dtype=torch.float32)
    Zn = torch.tensor(g.normal(size=X.shape), dtype=torch.float32)
    lam = 0.4 + 0.1 * X ** 2                                     # <= 1.9: cdf(9, 1.9) > 0.999
    assert (stats.poisson.cdf(K, lam) >= 0.999).all()
    v, gx, _ = evaluate("Poisson(0.4 + 0.1 * x(?o) * x(?o))",
                        {"poisson_nbins": K, "poisson_softmax_weight": w}, cls=Comp, noise=[G, Zn])
    ks = np.arange(K)
    logpmf = ks * np.log(lam[..., None]) - lam[..., None] - special.gammaln(ks + 1)
    z = w * (G.numpy().astype(np.float64) + logpmf)
    sm = np.exp(z - z.max(-1, keepdims=True))
    sm /= sm.sum(-1, keepdims=True)
    np.testing.assert_allclose(v, (ks * sm).sum(-1), rtol=2e-5, atol=2e-5)
    dl = ks / lam[..., None] - 1.0                                # d logpmf_k / d rate
    mean_k = (ks * sm).sum(-1, keepdims=True)
    dv = (w * sm * (ks - mean_k) * dl).sum(-1)
    np.testing.assert_allclose(gx, dv * 0.2 * X, rtol=2e-4, atol=2e-5)
