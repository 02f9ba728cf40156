"""Relaxation configuration classes: the drop-in mirror of ``pyRDDLGym_jax/core/logic.py``.

In the reference every relaxation is a mixin subclass of the compiler that overrides one
family of ``_jax_*`` methods and is composed by MRO into
``DefaultJaxRDDLCompilerWithGrad`` (logic.py:1771-1789).  The class names and keyword
arguments below are identical (logic.py:90-94, 247-249, 364-365, 427, 546, 665, 788-789,
867-868, 906, 947-948, 1013-1015, 1094-1096, 1140-1143, 1230-1232, 1362-1365, 1492-1495,
1599-1603), so a ``[Compiler]`` cfg section (planner.py:130-137) resolves unchanged.

Here a class does not build closures: it contributes (a) its keyword values and (b) a tag
naming which relaxation variant handles an operator family.  ``relax_config()`` flattens
the MRO into one plain dict that the instruction lowering (core/lowering.py) and the CPU
oracle both read.  Python's MRO gives the same precedence the reference gets from method
resolution: the left-most base that defines a family wins.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Set

from .compiler import JaxRDDLCompiler


class JaxRDDLCompilerWithGrad(JaxRDDLCompiler):
    """All fluents REAL; non-differentiable ops replaced per the mixed-in relaxations
    (logic.py:83-236)."""

    _RELAX: Dict[str, str] = {}
    relaxed = True

    def __init__(self, *args, cpfs_without_grad: Optional[Set[str]] = None,
                 print_warnings: bool = True, sample_callback: Optional[Callable] = None,
                 **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.cpfs_without_grad = set() if cpfs_without_grad is None else set(cpfs_without_grad)
        self.print_warnings = print_warnings
        self.sample_callback = sample_callback

    def get_kwargs(self) -> Dict[str, Any]:
        kwargs = super().get_kwargs()
        kwargs["cpfs_without_grad"] = self.cpfs_without_grad
        kwargs["print_warnings"] = self.print_warnings
        kwargs["sample_callback"] = self.sample_callback
        return kwargs

    # ---- information about the relaxed operations (compiler.py:804-832) ---------------------
    def _lowered(self):
        """A lowering of the relaxed model with externally supplied actions (cached): the source
        of the model parameters and of the relaxed / exact bookkeeping."""
        if getattr(self, "_lowered_cache", None) is None:
            from .program import PlanSpec
            self._lowered_cache = self.builder(
                PlanSpec(kind="external", bounds=self.rddl.bounds), True).low
        return self._lowered_cache

    def model_parameter_info(self) -> Dict[int, Dict[str, Any]]:
        """id -> {'id', 'rddl_op', 'init_value': {parameter name: value}} (compiler.py:804-814)."""
        low = self._lowered()
        out: Dict[int, Dict[str, Any]] = {}
        for (eid, name), default, _ in low.sg.mparams:
            op = low.mparam_ops.get((eid, name), "policy")
            out.setdefault(eid, {"id": eid, "rddl_op": op, "init_value": {}})["init_value"][name] = default
        return out

    def overriden_ops_info(self) -> Dict[str, Dict[str, List[int]]]:
        """relaxation class -> {rddl op: [expression ids]} (compiler.py:816-823): the class is the
        mixin of this compiler's MRO that declares the relaxation family of the operation."""
        owner: Dict[str, str] = {}
        for klass in type(self).__mro__:
            for family in vars(klass).get("_RELAX", {}):
                owner.setdefault(family, klass.__name__)
        out: Dict[str, Dict[str, List[int]]] = {}
        for eid, (family, op) in self._lowered().relaxed_ops.items():
            out.setdefault(owner.get(family, type(self).__name__), {}).setdefault(op, []).append(eid)
        return out

    def exact_ops_info(self) -> Dict[str, List[int]]:
        """rddl op -> [expression ids] of the relaxable operations evaluated exactly, because their
        arguments are not fluent or the compiler has no relaxation for them (compiler.py:825-832)."""
        out: Dict[str, List[int]] = {}
        for eid, (_, op) in self._lowered().exact_ops.items():
            out.setdefault(op, []).append(eid)
        return out

    def relax_config(self) -> Dict[str, Any]:
        """Flatten the MRO into {family: variant} + every keyword value."""
        cfg: Dict[str, Any] = {}
        for klass in reversed(type(self).__mro__):
            cfg.update(getattr(klass, "_RELAX", {}) if "_RELAX" in vars(klass) else {})
        out = {k: v for k, v in self.get_kwargs().items()
               if k not in ("sample_callback", "python_functions", "logger")}
        out["variants"] = cfg
        return out


class SigmoidRelational(JaxRDDLCompilerWithGrad):
    """Comparisons as sigmoids / tanh (logic.py:244-358)."""
    _RELAX = {"relational": "sigmoid", "sgn": "tanh"}

    def __init__(self, *args, sigmoid_weight: float = 10., use_sigmoid_ste: bool = True,
                 use_tanh_ste: bool = True, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.sigmoid_weight = float(sigmoid_weight)
        self.use_sigmoid_ste = use_sigmoid_ste
        self.use_tanh_ste = use_tanh_ste

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs.update(sigmoid_weight=self.sigmoid_weight, use_sigmoid_ste=self.use_sigmoid_ste,
                      use_tanh_ste=self.use_tanh_ste)
        return kwargs


class SoftmaxArgmax(JaxRDDLCompilerWithGrad):
    """argmin/argmax as softmax-weighted index sums (logic.py:361-417)."""
    _RELAX = {"argmax": "softmax"}

    def __init__(self, *args, argmax_weight: float = 10., use_argmax_ste: bool = True,
                 **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.argmax_weight = float(argmax_weight)
        self.use_argmax_ste = use_argmax_ste

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs.update(argmax_weight=self.argmax_weight, use_argmax_ste=self.use_argmax_ste)
        return kwargs


class _LogicalBase(JaxRDDLCompilerWithGrad):
    def __init__(self, *args, use_logic_ste: bool = False, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.use_logic_ste = use_logic_ste

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs["use_logic_ste"] = self.use_logic_ste
        return kwargs


class ProductNormLogical(_LogicalBase):
    """Product t-norm (logic.py:424-540)."""
    _RELAX = {"logic": "product"}


class GodelNormLogical(_LogicalBase):
    """Goedel t-norm, min/max (logic.py:543-659)."""
    _RELAX = {"logic": "godel"}


class LukasiewiczNormLogical(_LogicalBase):
    """Lukasiewicz t-norm (logic.py:662-778)."""
    _RELAX = {"logic": "lukasiewicz"}


class SoftFloor(JaxRDDLCompilerWithGrad):
    """floor / ceil / div / mod via tanh-smoothed staircase (logic.py:785-861)."""
    _RELAX = {"floor": "soft"}

    def __init__(self, *args, floor_weight: float = 10., use_floor_ste: bool = True,
                 **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.floor_weight = float(floor_weight)
        self.use_floor_ste = use_floor_ste

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs.update(floor_weight=self.floor_weight, use_floor_ste=self.use_floor_ste)
        return kwargs


class SoftRound(JaxRDDLCompilerWithGrad):
    """round via tanh-smoothed staircase (logic.py:864-896)."""
    _RELAX = {"round": "soft"}

    def __init__(self, *args, round_weight: float = 10., use_round_ste: bool = True,
                 **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.round_weight = float(round_weight)
        self.use_round_ste = use_round_ste

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs.update(round_weight=self.round_weight, use_round_ste=self.use_round_ste)
        return kwargs


class LinearIfElse(JaxRDDLCompilerWithGrad):
    """if-then-else as a linear blend (logic.py:903-941)."""
    _RELAX = {"if": "linear"}

    def __init__(self, *args, use_if_else_ste: bool = True, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.use_if_else_ste = use_if_else_ste

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs["use_if_else_ste"] = self.use_if_else_ste
        return kwargs


class SoftmaxSwitch(JaxRDDLCompilerWithGrad):
    """switch via softmax over case distance (logic.py:944-1003)."""
    _RELAX = {"switch": "softmax"}

    def __init__(self, *args, switch_weight: float = 10., use_switch_ste: bool = True,
                 **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.switch_weight = float(switch_weight)
        self.use_switch_ste = use_switch_ste

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs.update(switch_weight=self.switch_weight, use_switch_ste=self.use_switch_ste)
        return kwargs


class ReparameterizedGeometric(JaxRDDLCompilerWithGrad):
    """Geometric by inverse CDF with a soft floor (logic.py:1010-1052)."""
    _RELAX = {"geometric": "reparam"}

    def __init__(self, *args, geometric_floor_weight: float = 10., geometric_eps: float = 1e-14,
                 **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.geometric_floor_weight = float(geometric_floor_weight)
        self.geometric_eps = float(geometric_eps)

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs.update(geometric_floor_weight=self.geometric_floor_weight,
                      geometric_eps=self.geometric_eps)
        return kwargs


class DeterminizedGeometric(JaxRDDLCompilerWithGrad):
    """Geometric replaced by its mean (logic.py:1055-1084)."""
    _RELAX = {"geometric": "determinized"}


class ReparameterizedSigmoidBernoulli(JaxRDDLCompilerWithGrad):
    """Bernoulli as sigmoid(w (p - U)) (logic.py:1091-1134)."""
    _RELAX = {"bernoulli": "sigmoid"}

    def __init__(self, *args, bernoulli_sigmoid_weight: float = 10.,
                 use_ste_bernoulli: bool = False, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.bernoulli_sigmoid_weight = float(bernoulli_sigmoid_weight)
        self.use_ste_bernoulli = use_ste_bernoulli

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs.update(bernoulli_sigmoid_weight=self.bernoulli_sigmoid_weight,
                      use_ste_bernoulli=self.use_ste_bernoulli)
        return kwargs


class GumbelSigmoidBernoulli(JaxRDDLCompilerWithGrad):
    """Bernoulli as sigmoid(w (logit p + Logistic)) (logic.py:1137-1187)."""
    _RELAX = {"bernoulli": "gumbel"}

    def __init__(self, *args, bernoulli_sigmoid_weight: float = 10.,
                 use_ste_bernoulli: bool = False, bernoulli_prob_eps: float = 1e-12,
                 **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.bernoulli_sigmoid_weight = float(bernoulli_sigmoid_weight)
        self.use_ste_bernoulli = use_ste_bernoulli
        self.bernoulli_prob_eps = bernoulli_prob_eps

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs.update(bernoulli_sigmoid_weight=self.bernoulli_sigmoid_weight,
                      use_ste_bernoulli=self.use_ste_bernoulli,
                      bernoulli_prob_eps=self.bernoulli_prob_eps)
        return kwargs


class DeterminizedBernoulli(JaxRDDLCompilerWithGrad):
    """Bernoulli replaced by its mean (logic.py:1190-1219)."""
    _RELAX = {"bernoulli": "determinized"}


class GumbelSoftmaxDiscrete(JaxRDDLCompilerWithGrad):
    """Discrete via Gumbel-softmax, soft value, no STE (logic.py:1227-1297)."""
    _RELAX = {"discrete": "gumbel"}

    def __init__(self, *args, discrete_softmax_weight: float = 10., discrete_eps: float = 1e-14,
                 **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.discrete_softmax_weight = float(discrete_softmax_weight)
        self.discrete_eps = float(discrete_eps)

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs.update(discrete_softmax_weight=self.discrete_softmax_weight,
                      discrete_eps=self.discrete_eps)
        return kwargs


class DeterminizedDiscrete(JaxRDDLCompilerWithGrad):
    """Discrete replaced by its mean (logic.py:1300-1351)."""
    _RELAX = {"discrete": "determinized"}


class GumbelSoftmaxBinomial(JaxRDDLCompilerWithGrad):
    """Binomial: Gumbel-softmax below nbins, normal approximation above (logic.py:1358-1444)."""
    _RELAX = {"binomial": "gumbel"}

    def __init__(self, *args, binomial_nbins: int = 100, binomial_softmax_weight: float = 10.,
                 binomial_eps: float = 1e-14, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.binomial_nbins = binomial_nbins
        self.binomial_softmax_weight = float(binomial_softmax_weight)
        self.binomial_eps = float(binomial_eps)

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs.update(binomial_nbins=self.binomial_nbins,
                      binomial_softmax_weight=self.binomial_softmax_weight,
                      binomial_eps=self.binomial_eps)
        return kwargs


class DeterminizedBinomial(JaxRDDLCompilerWithGrad):
    """Binomial replaced by its mean (logic.py:1447-1482)."""
    _RELAX = {"binomial": "determinized"}


class ExponentialPoisson(JaxRDDLCompilerWithGrad):
    """Poisson / NegativeBinomial via the exponential inter-arrival trick
    (logic.py:1489-1592)."""
    _RELAX = {"poisson": "exponential"}

    def __init__(self, *args, poisson_nbins: int = 100, poisson_comparison_weight: float = 10.,
                 poisson_min_cdf: float = 0.999, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.poisson_nbins = poisson_nbins
        self.poisson_comparison_weight = float(poisson_comparison_weight)
        self.poisson_min_cdf = float(poisson_min_cdf)

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs.update(poisson_nbins=self.poisson_nbins,
                      poisson_comparison_weight=self.poisson_comparison_weight,
                      poisson_min_cdf=self.poisson_min_cdf)
        return kwargs


class GumbelSoftmaxPoisson(JaxRDDLCompilerWithGrad):
    """Poisson via truncated Gumbel-softmax (logic.py:1595-1708)."""
    _RELAX = {"poisson": "gumbel"}

    def __init__(self, *args, poisson_nbins: int = 100, poisson_softmax_weight: float = 10.,
                 poisson_min_cdf: float = 0.999, poisson_eps: float = 1e-14, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.poisson_nbins = poisson_nbins
        self.poisson_softmax_weight = float(poisson_softmax_weight)
        self.poisson_min_cdf = float(poisson_min_cdf)
        self.poisson_eps = float(poisson_eps)

    def get_kwargs(self):
        kwargs = super().get_kwargs()
        kwargs.update(poisson_nbins=self.poisson_nbins,
                      poisson_softmax_weight=self.poisson_softmax_weight,
                      poisson_min_cdf=self.poisson_min_cdf, poisson_eps=self.poisson_eps)
        return kwargs


class DeterminizedPoisson(JaxRDDLCompilerWithGrad):
    """Poisson / NegativeBinomial replaced by their means (logic.py:1711-1768)."""
    _RELAX = {"poisson": "determinized"}


class DefaultJaxRDDLCompilerWithGrad(SigmoidRelational, SoftmaxArgmax,
                                     ProductNormLogical,
                                     SoftFloor, SoftRound,
                                     LinearIfElse, SoftmaxSwitch,
                                     ReparameterizedGeometric,
                                     ReparameterizedSigmoidBernoulli,
                                     GumbelSoftmaxDiscrete, GumbelSoftmaxBinomial,
                                     ExponentialPoisson):
    """Default composition (logic.py:1771-1789)."""

    def get_kwargs(self) -> Dict[str, Any]:
        kwargs: Dict[str, Any] = {}
        for base in type(self).__bases__:
            if base.__name__ != "object":
                kwargs = {**kwargs, **base.get_kwargs(self)}
        return kwargs


class GumbelBernoulliJaxRDDLCompilerWithGrad(SigmoidRelational, SoftmaxArgmax,
                                            ProductNormLogical,
                                            SoftFloor, SoftRound,
                                            LinearIfElse, SoftmaxSwitch,
                                            ReparameterizedGeometric,
                                            GumbelSigmoidBernoulli,
                                            GumbelSoftmaxDiscrete, GumbelSoftmaxBinomial,
                                            ExponentialPoisson):
    """The default composition with the Gumbel-sigmoid Bernoulli relaxation (logic.py:1137-1186)
    in place of the sigmoid one: the composite class BASELINE.json's Wildfire configuration
    needs ("FuzzyLogic + Gumbel-softmax Bernoulli"); the reference's cfg loader can only name
    existing classes (planner.py:135-136), so a user would define exactly this."""

    get_kwargs = DefaultJaxRDDLCompilerWithGrad.get_kwargs
