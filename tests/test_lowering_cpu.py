"""CPU tests: lowering + differentiation + assembler, executed by the numpy ISA emulator,
against the torch oracle (same tolerances the GPU tests use)."""
import pytest

from pyrddlgym_jax_b200.core import logic

from .parity import check_forward, check_gradient, forward_case, gradient_case

KEYS = ["quadcopter", "reservoir", "wildfire", "marsrover", "powergen", "discrete",
        "distributions", "ops", "samplers", "matrix", "pomdp"]


@pytest.mark.parametrize("key", KEYS)
@pytest.mark.parametrize("relaxed", [False, True])
def test_forward_emulated(key, relaxed):
    check_forward(forward_case("emu", key, B=13, T=6, relaxed=relaxed))


@pytest.mark.parametrize("key", KEYS)
@pytest.mark.parametrize("gamma", [1.0, 0.9])
def test_gradient_emulated(key, gamma):
    check_gradient(gradient_case("emu", key, B=11, T=5, gamma=gamma))


class _Godel(logic.SigmoidRelational, logic.GodelNormLogical, logic.LinearIfElse,
             logic.GumbelSigmoidBernoulli, logic.SoftFloor, logic.SoftRound):
    pass


class _Luka(logic.SigmoidRelational, logic.LukasiewiczNormLogical, logic.LinearIfElse,
            logic.DeterminizedBernoulli):
    pass


@pytest.mark.parametrize("cls,kw", [
    (_Godel, dict(sigmoid_weight=3.0, use_logic_ste=True, use_ste_bernoulli=True)),
    (_Godel, dict(sigmoid_weight=3.0, use_sigmoid_ste=False, use_if_else_ste=False)),
    (_Luka, dict(sigmoid_weight=2.0)),
    (logic.DefaultJaxRDDLCompilerWithGrad, dict(sigmoid_weight=100.0, bernoulli_sigmoid_weight=100.0)),
])
def test_relaxation_variants_emulated(cls, kw):
    """Other t-norms / Bernoulli relaxations / STE switches on the Wildfire and Reservoir CPFs."""
    for key in ("wildfire", "reservoir"):
        c = gradient_case("emu", key, B=7, T=4, compiler_cls=cls, **kw)
        check_forward(c)
        check_gradient(c)


class _GodelOps(logic.SigmoidRelational, logic.SoftmaxArgmax, logic.GodelNormLogical,
                logic.SoftFloor, logic.SoftRound, logic.LinearIfElse, logic.SoftmaxSwitch,
                logic.ReparameterizedGeometric):
    pass


class _LukaOps(logic.SigmoidRelational, logic.SoftmaxArgmax, logic.LukasiewiczNormLogical,
               logic.SoftFloor, logic.SoftRound, logic.LinearIfElse, logic.SoftmaxSwitch,
               logic.DeterminizedGeometric):
    pass


OPS_VARIANTS = [
    (logic.DefaultJaxRDDLCompilerWithGrad, dict(use_sigmoid_ste=False, use_tanh_ste=False,
                                                use_argmax_ste=False, use_floor_ste=False,
                                                use_round_ste=False, use_if_else_ste=False)),
    (logic.DefaultJaxRDDLCompilerWithGrad, dict(use_logic_ste=True, sigmoid_weight=4.0,
                                                argmax_weight=3.0, floor_weight=6.0, round_weight=7.0)),
    (_GodelOps, dict(sigmoid_weight=3.0)),
    (_GodelOps, dict(sigmoid_weight=3.0, use_logic_ste=True)),
    (_LukaOps, dict(sigmoid_weight=2.0)),
    (_LukaOps, dict(sigmoid_weight=2.0, use_logic_ste=True, use_sigmoid_ste=False)),
]


@pytest.mark.parametrize("cls,kw", OPS_VARIANTS)
def test_ops_relaxation_variants_emulated(cls, kw):
    """Every relaxation family (SigmoidRelational incl. == ~= sgn, SoftmaxArgmax, the three t-norms
    incl. => <=> xor, SoftFloor / SoftRound incl. div mod, ReparameterizedGeometric) with and without
    straight-through estimators, on the fixtures that use every operator of the instruction set."""
    for key in ("ops", "samplers"):
        c = gradient_case("emu", key, B=7, T=4, compiler_cls=cls, **kw)
        check_forward(c)
        check_gradient(c)
        assert float(abs(c["grad"]).max()) > 0


@pytest.mark.parametrize("key", ["reservoir", "quadcopter", "powergen"])
def test_wrap_non_bool_emulated(key):
    """Box-wrapped real actions (_jax_bound_action, planner.py:620-628, 755-762, 862): sigmoid
    for two-sided boxes, softplus for one-sided ones, in the train and the test policy."""
    check_forward(forward_case("emu", key, B=9, T=5, relaxed=False, wrap_non_bool=True))
    c = gradient_case("emu", key, B=9, T=5, wrap_non_bool=True)
    check_forward(c)
    check_gradient(c)


@pytest.mark.parametrize("key,wrap", [("reservoir", False), ("wildfire", False), ("quadcopter", True)])
def test_stochastic_plan_emulated(key, wrap):
    """Stochastic straight-line plans (planner.py:811-834): logistic noise on boolean logits,
    mu + exp(clip(log_sigma)) * noise on the others; gradients reach mu, log_sigma and logits."""
    check_forward(forward_case("emu", key, B=9, T=5, relaxed=False, stochastic=True,
                               wrap_non_bool=wrap))
    c = gradient_case("emu", key, B=9, T=5, stochastic=True, wrap_non_bool=wrap)
    check_forward(c)
    check_gradient(c)
    ls = [k for k in c["P"] if k.endswith("/log_sigma")]
    if ls:
        assert any(c["P"][k].grad is not None and float(c["P"][k].grad.abs().sum()) > 0 for k in ls)


class _GumbelCounts(logic.SigmoidRelational, logic.ProductNormLogical, logic.LinearIfElse,
                    logic.SoftmaxSwitch, logic.GumbelSoftmaxPoisson, logic.GumbelSoftmaxBinomial):
    pass


def test_gumbel_softmax_count_relaxations_emulated():
    """GumbelSoftmaxPoisson / GumbelSoftmaxBinomial (logic.py:1381-1434, 1618-1641): soft arg-max
    over the bins of Gumbel noise + log pmf, with the normal-approximation branch; 12 bins keep
    the emulated program small."""
    c = gradient_case("emu", "distributions", B=5, T=3, compiler_cls=_GumbelCounts,
                      poisson_nbins=12, binomial_nbins=12, poisson_softmax_weight=3.0)
    check_forward(c)
    check_gradient(c)


def test_pooled_softmax_plan_emulated():
    """wrap_softmax (planner.py:764-808, 848-853): one softmax over the pooled boolean parameters
    of a step, 1 - p for fluents whose no-op is true, (p > 0.5) in the test policy."""
    check_forward(forward_case("emu", "wildfire", B=7, T=4, relaxed=False, pooled_softmax=True,
                               param_scale=3.0))
    c = gradient_case("emu", "wildfire", B=7, T=4, pooled_softmax=True, param_scale=3.0)
    check_forward(c)
    check_gradient(c)


def test_stochastic_pooled_softmax_plan_emulated():
    """stochastic + wrap_softmax (planner.py:797-806): Gumbel noise on the pooled boolean logits
    before the softmax, per rollout."""
    c = gradient_case("emu", "wildfire", B=7, T=4, pooled_softmax=True, stochastic=True,
                      param_scale=2.0)
    check_forward(c)
    check_gradient(c)
    acts = [c["P"][k].grad for k in c["P"]]
    assert all(g is not None and float(g.abs().sum()) > 0 for g in acts)


def test_gamma_family_with_fluent_shapes_emulated():
    """Gamma / ChiSquare / Student / Beta / NegativeBinomial whose shape is a fluent: sampled in the
    step (Marsaglia-Tsang on supplied normal / uniform draws); exact and relaxed programs against the
    oracle, gradients through the shapes included."""
    check_forward(forward_case("emu", "gammafluent", B=9, T=4, relaxed=False))
    c = gradient_case("emu", "gammafluent", B=9, T=4)
    check_forward(c)
    check_gradient(c)
    assert float(abs(c["grad"]).max()) > 0


def test_matrix_and_random_vector_syntax():
    """Parser / tracer surface of the a14 families (compiler.py:2247-2434): expression kinds, scoping of
    the row / column / sample variables, and the rejections."""
    from pyrddlgym_jax_b200.rddl.parser import _Parser, RDDLParseError
    from pyrddlgym_jax_b200.rddl.expr import walk
    e = _Parser("det_{?a : dim, ?b : dim}[ cov(?a, ?b) ] + 1.0").parse_expr()
    kinds = [n.etype for n in walk(e)]
    assert ("matrix", "det") in kinds and ("pvar", "cov") in kinds
    e = _Parser("inverse[row=?i, col=?j][ cov(?i, ?j) * 2.0 ]").parse_expr()
    assert e.etype == ("matrix", "inverse") and e.args[0] == ("?i", "?j")
    e = _Parser("MultivariateNormal[?i, ?j](mu(?i), cov(?i, ?j))").parse_expr()
    assert e.etype == ("randomvector", "MultivariateNormal") and e.args[0] == ("?i", "?j") and len(e.args) == 3
    e = _Parser("Dirichlet[?i](ALPHA(?i))").parse_expr()
    assert e.etype == ("randomvector", "Dirichlet") and e.args[0] == ("?i",)
    with pytest.raises(RDDLParseError):
        _Parser("det_{?a : dim}[ cov(?a, ?a) ]").parse_expr()
    with pytest.raises(RDDLParseError):
        _Parser("cholesky[col=?i, row=?j][ cov(?i, ?j) ]").parse_expr()
    # lowering: row and column must be two distinct variables of the scope
    from pyrddlgym_jax_b200.core.lowering import Lowerer, RDDLNotImplementedError
    from pyrddlgym_jax_b200.rddl.trace import RDDLTraceError
    from .helpers import fixture_model
    m = fixture_model("matrix")
    lw = Lowerer(m, None)
    env = {}
    scope = [("?i", "dim"), ("?j", "dim")]
    bad = _Parser("inverse[row=?i, col=?i][ BASE(?i, ?j) ]").parse_expr()
    with pytest.raises(RDDLTraceError):
        lw.lower(bad, scope, env)
    multi = _Parser("Multinomial[?i](?i, ALPHA(?i))").parse_expr()     # trials vary over the sample axis
    with pytest.raises(RDDLTraceError):
        lw.lower(multi, scope, env)
    assert RDDLNotImplementedError is not None


SWITCH_RDDL = """
domain switch_toy {
    types { site : object; level : {@low, @mid, @high}; };
    pvariables {
        x(site)    : { state-fluent, real, default = 0.3 };
        mode(site) : { state-fluent, level, default = @low };
        push(site) : { action-fluent, real, default = 0.0 };
    };
    cpfs {
        mode'(?s) = if (x(?s) > 0.6) then @high else (if (x(?s) > 0.35) then @mid else @low);
        x'(?s) = x(?s) + switch (if (push(?s) > 0.5) then @high else mode(?s)) {
                     case @low : 0.15, case @mid : 0.5 * push(?s), default : -0.2 * x(?s) };
    };
    reward = - (sum_{?s : site} [ pow[x'(?s) - 0.5, 2] ]);
    action-preconditions { forall_{?s : site} [ push(?s) >= -1.0 ]; forall_{?s : site} [ push(?s) <= 1.0 ]; };
}
"""
SWITCH_INST = """
non-fluents nf { domain = switch_toy; objects { site : {s1, s2, s3}; }; }
instance inst { domain = switch_toy; non-fluents = nf; init-state { x(s2) = 0.5; x(s3) = 0.8; mode(s3) = @high; };
    max-nondef-actions = pos-inf; horizon = 5; discount = 1.0; }
"""


def test_switch_over_a_computed_predicate(tmp_path):
    """switch whose predicate is an expression (compiler.py:1661-1684 compiles any predicate; the enum type
    comes from the case labels here): exact and relaxed rollouts and the gradient, emulator vs oracle."""
    from . import helpers
    (tmp_path / "domain.rddl").write_text(SWITCH_RDDL)
    (tmp_path / "instance1.rddl").write_text(SWITCH_INST)
    helpers.FIXTURES["switch_toy"] = (str(tmp_path), "instance1")
    try:
        check_forward(forward_case("emu", "switch_toy", B=9, T=5, relaxed=False))
        c = gradient_case("emu", "switch_toy", B=9, T=5)
        check_forward(c)
        check_gradient(c)
        assert float(abs(c["grad"]).max()) > 0
    finally:
        helpers.FIXTURES.pop("switch_toy", None)
