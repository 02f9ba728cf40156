"""``JaxBackpropPlanner`` and friends: the drop-in Python surface of
pyRDDLGym_jax/core/planner.py, driving the B200 rollout kernels instead of XLA.

Kept seams: ``JaxStraightLinePlan`` kwargs (planner.py:636-648),
``JaxBackpropPlanner`` kwargs (planner.py:2399-2429), ``optimize`` /
``optimize_generator`` and the callback dict they yield (planner.py:3073-3087, 3368-3392),
``get_action`` (planner.py:3527-3533), ``JaxPlannerStatus`` (planner.py:2277-2304),
``load_config`` (planner.py:229-244) and the controllers (planner.py:3632-3831).

One planner iteration = exactly what the reference does per epoch with ``pgpe=None``
(planner.py:3236-3262): ``update`` (relaxed forward rollouts with a state tape, loss,
backprop-through-time, gradient reduction, optimiser step, projection) followed by
``test_loss`` (exact-model forward rollouts with the updated plan).  On the device that is
6 kernel launches, replayed from one CUDA graph.
"""
from __future__ import annotations

import configparser
import math
import pickle
import time
from abc import ABCMeta, abstractmethod
from ast import literal_eval
from collections import deque
from enum import Enum
from typing import Any, Callable, Dict, Iterable, List, Optional, Tuple, Union

import numpy as np
import torch

from .linesearch import zoom_kwargs, zoom_linesearch
from . import engine, logic, parallel
from .compiler import JaxRDDLCompiler
from .layout import draw_noise, pack_fluent_major, unpack_fluent_major
from .program import PlanSpec, action_layout

Kwargs = Dict[str, Any]


# =====================================================================================
# initialisers / optimisers named in .cfg files (planner.py:140-186 resolve them with getattr
# on jax / optax; those modules do not exist here, so the names map to local objects)
# =====================================================================================

from .planner_base import (JaxPlan, Preprocessor, StaticNormalizer,  # noqa: E402,F401
                           _Initializer, initializers)


OPTIMIZERS = {"sgd": "sgd", "rmsprop": "rmsprop", "adam": "adam"}


# =====================================================================================
# plans
# =====================================================================================

class JaxStraightLinePlan(JaxPlan):
    """Open-loop plan: one parameter per action element per step (planner.py:631-1018)."""

    def __init__(self, initializer=None, wrap_sigmoid: bool = True, min_action_prob: float = 1e-6,
                 sigmoid_weight: float = 1.0, wrap_non_bool: bool = False,
                 wrap_softmax: bool = False, softmax_weight: float = 1.0,
                 use_new_projection: bool = False, max_constraint_iter: int = 100,
                 stochastic: bool = False, sigma_range: Tuple[float, float] = (1e-5, 1e3),
                 sigma_entropy_grad: bool = False, action_noise_dist: Any = "normal") -> None:
        super().__init__()
        self._initializer_base = initializer if initializer is not None else initializers.normal()
        self._wrap_sigmoid = wrap_sigmoid
        self._min_action_prob = min_action_prob
        self._sigmoid_weight = sigmoid_weight
        self._wrap_non_bool = wrap_non_bool
        self._wrap_softmax = wrap_softmax
        self._softmax_weight = softmax_weight
        self._use_new_projection = use_new_projection
        self._max_constraint_iter = max_constraint_iter
        self._stochastic = stochastic
        self._sigma_range = sigma_range
        self._sigma_entropy_grad = sigma_entropy_grad
        self._action_noise_dist = action_noise_dist
        if action_noise_dist not in ("normal", "logistic", "laplace", "gumbel", "cauchy", "uniform"):
            raise NotImplementedError(f"action_noise_dist <{action_noise_dist}>: name one of the "
                                      "noise kinds the engine draws (e.g. 'normal')")

    def __str__(self) -> str:
        return (f"[INFO] policy hyper-parameters:\n"
                f"    initializer ={self._initializer_base}\n"
                f"    wrap_sigmoid={self._wrap_sigmoid}\n"
                f"    wrap_sigmoid_min_prob={self._min_action_prob}\n"
                f"    sigmoid_weight={self._sigmoid_weight}\n"
                f"    use_new_projection={self._use_new_projection}\n"
                f"    max_projection_iters={self._max_constraint_iter}\n")

    def compile(self, compiled, test_compiled, _bounds, horizon, preprocessor=None) -> None:
        rddl = compiled.rddl
        self.rddl = rddl
        self.horizon = horizon
        self.layout, self.A_act = action_layout(rddl)
        self.A = self.A_act
        # stochastic plans: one log_sigma block per non-boolean fluent after the action rows
        self.sigma_layout: Dict[str, Tuple[int, int]] = {}
        if self._stochastic:
            for name, (off, n, prange) in self.layout.items():
                if prange != "bool":
                    self.sigma_layout[name] = (self.A, n)
                    self.A += n
        bounds = {}
        for name, (off, n, prange) in self.layout.items():      # get_action_info, planner.py:423-472
            if prange == "bool":
                bounds[name] = (None, None)
            elif prange in rddl.enum_types:
                shape = rddl.variable_shape(name)
                bounds[name] = (np.zeros(shape, np.float32),
                                np.full(shape, len(rddl.type_to_objects[prange]) - 1, np.float32))
            else:
                lo, hi = rddl.bounds[name]
                lo, hi = _bounds.get(name, (lo, hi))
                shape = rddl.variable_shape(name)
                bounds[name] = (np.broadcast_to(np.asarray(lo, np.float32), shape).copy(),
                                np.broadcast_to(np.asarray(hi, np.float32), shape).copy())
        self.bounds = bounds
        self.noop = {var: bool(np.ravel(v)[0]) if rddl.variable_ranges[var] == "bool" else v
                     for var, v in rddl.action_fluents.items()}
        n_bool = sum(n for _, n, r in self.layout.values() if r == "bool")
        # wrap_softmax: one softmax over the pooled boolean parameters of a step when the
        # concurrency constraint binds (planner.py:764-808, 924-933); ignored otherwise
        self.stack_bool = bool(self._wrap_softmax) and rddl.max_allowed_actions < n_bool
        if self.stack_bool and rddl.max_allowed_actions > 1:
            raise ValueError("SLPs with wrap_softmax currently do not support action cardinality "
                             f"constraint {rddl.max_allowed_actions} > 1.")
        self.spec = PlanSpec(kind="slp", wrap_sigmoid=self._wrap_sigmoid,
                             pooled_softmax=self.stack_bool,
                             softmax_weight=float(self._softmax_weight),
                             noop_true=tuple(v for v, (_, _, r) in self.layout.items()
                                             if r == "bool" and self.noop[v]),
                             wrap_non_bool=self._wrap_non_bool,
                             stochastic=self._stochastic, sigma_range=tuple(self._sigma_range),
                             action_noise=self._action_noise_dist,
                             sigmoid_weight=self._sigmoid_weight,
                             bounds={k: v for k, v in bounds.items() if v[0] is not None})
        self.bool_threshold = 0.0 if (self._wrap_sigmoid or self.stack_bool) else 0.5
        # flat box for the projection (planner.py:878-904)
        lo = np.full((horizon, self.A), -np.inf, np.float32)
        hi = np.full((horizon, self.A), np.inf, np.float32)
        for name, (off, n, prange) in self.layout.items():
            if prange == "bool" and self.stack_bool:
                continue                                # pooled logits are not projected
            if prange == "bool":
                p = self._min_action_prob
                if self._wrap_sigmoid:
                    l, h = np.log(p / (1 - p)) / self._sigmoid_weight, \
                        np.log((1 - p) / p) / self._sigmoid_weight
                else:
                    l, h = p, 1 - p
                lo[:, off:off + n], hi[:, off:off + n] = l, h
            elif not self._wrap_non_bool:           # planner.py:898-899: wrapped means stay free
                lo[:, off:off + n] = bounds[name][0].reshape(-1)
                hi[:, off:off + n] = bounds[name][1].reshape(-1)
        for name, (off, n) in self.sigma_layout.items():     # planner.py:900-902
            lo[:, off:off + n] = np.log(self._sigma_range[0])
            hi[:, off:off + n] = np.log(self._sigma_range[1])
        self.box = (lo, hi)
        # concurrency projection (planner.py:914-959): sort-based or SOGBOFA, per time step
        self.max_allowed_actions = int(min(rddl.max_allowed_actions, n_bool))
        self.needs_concurrency_projection = rddl.max_allowed_actions < n_bool and not self.stack_bool
        self.projection_method = "sorting" if self._use_new_projection else "sogbofa"
        self.bool_segs = [(off, n, bool(self.noop[name]), float(self._sigmoid_weight))
                          for name, (off, n, prange) in self.layout.items() if prange == "bool"]

    def initial_params(self, generator: torch.Generator) -> torch.Tensor:
        """planner.py:974-1008: initialiser + threshold shift + box projection -> [T, A]."""
        out = torch.zeros(self.horizon, self.A)
        for name, (off, n, prange) in self.layout.items():
            p = self._initializer_base((self.horizon, n), generator)
            if prange == "bool":
                p = p + self.bool_threshold
            out[:, off:off + n] = p
            if name in self.sigma_layout:                    # planner.py:990-993
                so, sn = self.sigma_layout[name]
                out[:, so:so + sn] = self._initializer_base((self.horizon, n), generator)
        lo, hi = self.box
        return torch.minimum(torch.maximum(out, torch.from_numpy(lo)), torch.from_numpy(hi))

    def params_to_pytree(self, flat: torch.Tensor) -> Dict[str, Any]:
        """[..., T, A] -> {'bool': {var: {'logit'}}, 'real': {var: {'mu', 'log_sigma'}}}."""
        tree: Dict[str, Any] = {"bool": {}, "real": {}}
        if self.stack_bool:         # planner.py:1000-1005: one [T, n_bool] tensor under '__bool__'
            tree["bool"]["__bool__"] = {"logit": torch.cat(
                [flat[..., off:off + n] for _, (off, n, r) in self.layout.items() if r == "bool"],
                dim=-1)}
        for name, (off, n, prange) in self.layout.items():
            v = flat[..., off:off + n].reshape(*flat.shape[:-1], *self.rddl.variable_shape(name))
            if prange == "bool":
                if not self.stack_bool:
                    tree["bool"][name] = {"logit": v}
            else:
                ls = None
                if name in self.sigma_layout:
                    so, sn = self.sigma_layout[name]
                    ls = flat[..., so:so + sn].reshape(*flat.shape[:-1],
                                                       *self.rddl.variable_shape(name))
                tree["real"][name] = {"mu": v, "log_sigma": ls}
        return tree

    def entropy_terms(self, params: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Entropy of a stochastic plan summed over the horizon, and its gradient w.r.t. the
        parameters [T, A] (planner.py:812-830): Bernoulli entropy of sigmoid(w logit) for wrapped
        booleans (safe_log, whose tangent is 1 / (x + 1e-12), compiler.py:65-75), sum of the
        clipped log_sigma for the others -- differentiated only when sigma_entropy_grad."""
        H = params.new_zeros(())
        dH = torch.zeros_like(params)
        eps = 1e-12
        if self.stack_bool:     # planner.py:797-800: entropy of softmax(w pooled logits) per step
            cols = torch.as_tensor(np.concatenate(
                [np.arange(off, off + n) for (off, n, r) in self.layout.values() if r == "bool"]),
                device=params.device)
            w = float(self._softmax_weight)
            lp = torch.log_softmax(w * params[:, cols], dim=1)
            pr = torch.exp(lp)
            Ht = -(pr * lp).sum(1, keepdim=True)
            H = H + Ht.sum()
            dH[:, cols] = -w * pr * (lp + Ht)
        for name, (off, n, prange) in self.layout.items():
            if prange == "bool":
                if not self._wrap_sigmoid or self.stack_bool:
                    continue
                w = float(self._sigmoid_weight)
                pr = torch.sigmoid(w * params[:, off:off + n])
                q = 1.0 - pr
                lp = torch.where(pr > 0, torch.log(pr.clamp_min(1e-45)), torch.full_like(pr, -np.inf))
                lq = torch.where(q > 0, torch.log(q.clamp_min(1e-45)), torch.full_like(q, -np.inf))
                H = H - torch.sum(torch.nan_to_num(pr * lp) + torch.nan_to_num(q * lq))
                df = torch.nan_to_num(lp, neginf=0.0) + pr / (pr + eps) \
                    - torch.nan_to_num(lq, neginf=0.0) - q / (q + eps)
                dH[:, off:off + n] = -df * w * pr * q
            elif name in self.sigma_layout:
                so, sn = self.sigma_layout[name]
                lo, hi = float(np.log(self._sigma_range[0])), float(np.log(self._sigma_range[1]))
                ls = params[:, so:so + sn]
                H = H + torch.sum(ls.clamp(lo, hi))
                if self._sigma_entropy_grad:
                    dH[:, so:so + sn] = ((ls >= lo) & (ls <= hi)).to(params.dtype)
        return H, dH

    def pytree_to_params(self, tree: Dict[str, Any]) -> torch.Tensor:
        out = torch.zeros(self.horizon, self.A)
        pooled, pos = None, 0
        if self.stack_bool:
            pooled = torch.as_tensor(np.asarray(tree["bool"]["__bool__"]["logit"]),
                                     dtype=torch.float32).reshape(-1, sum(
                                         n for _, n, r in self.layout.values() if r == "bool"))
        for name, (off, n, prange) in self.layout.items():
            if prange == "bool" and pooled is not None:
                out[:, off:off + n] = pooled[-self.horizon:, pos:pos + n]
                pos += n
                continue
            leaf = tree["bool"][name]["logit"] if prange == "bool" else tree["real"][name]["mu"]
            out[:, off:off + n] = torch.as_tensor(np.asarray(leaf), dtype=torch.float32
                                                  ).reshape(-1, n)[-self.horizon:]
            if name in self.sigma_layout and tree["real"][name].get("log_sigma") is not None:
                so, sn = self.sigma_layout[name]
                out[:, so:so + sn] = torch.as_tensor(
                    np.asarray(tree["real"][name]["log_sigma"]), dtype=torch.float32
                ).reshape(-1, sn)[-self.horizon:]
        return out

    def test_actions(self, flat_row: np.ndarray) -> Dict[str, np.ndarray]:
        """Test policy at one step (planner.py:842-868) -- host side, for get_action."""
        acts = {}
        soft = None
        if self.stack_bool:         # planner.py:848-853: softmax over the pooled logits, > 0.5
            z = np.concatenate([np.asarray(flat_row[off:off + n], np.float64)
                                for _, (off, n, r) in self.layout.items() if r == "bool"])
            z = np.exp(self._softmax_weight * (z - z.max()))
            soft, pos = z / z.sum(), 0
        for name, (off, n, prange) in self.layout.items():
            p = np.asarray(flat_row[off:off + n], np.float32).reshape(self.rddl.variable_shape(name))
            if prange == "bool" and soft is not None:
                a = soft[pos:pos + n].reshape(self.rddl.variable_shape(name))
                pos += n
                acts[name] = ((1.0 - a) if self.noop[name] else a) > 0.5
            elif prange == "bool":
                acts[name] = p > self.bool_threshold
            else:
                if self._wrap_non_bool:                 # _jax_bound_action, planner.py:620-628
                    lo, hi = self.bounds[name]
                    fl, fh = np.isfinite(lo), np.isfinite(hi)
                    lo0, hi0 = np.where(fl, lo, 0.0), np.where(fh, hi, 0.0)
                    sp = lambda x: np.logaddexp(0.0, x)      # noqa: E731
                    p = np.where(fl & fh, lo0 + (hi0 - lo0) / (1.0 + np.exp(-p)),
                                 np.where(fl, lo0 + sp(p), np.where(fh, hi0 - sp(-p), p))
                                 ).astype(np.float32)
                a = np.minimum(np.maximum(p, self.bounds[name][0]), self.bounds[name][1])
                if prange != "real":
                    a = np.round(a).astype(np.int32)
                acts[name] = a
        return acts

    def guess_next_epoch(self, params):
        """Shift the plan one step (planner.py:1011-1018)."""
        def shift(x):
            if x is None:
                return None
            x = torch.as_tensor(np.asarray(x))
            return torch.cat([x[1:], x[-1:]], dim=0)
        return {k: {var: {leaf: shift(v) for leaf, v in d.items()} for var, d in sub.items()}
                for k, sub in params.items()}


# =====================================================================================
# supporting classes (planner.py:2260-2338)
# =====================================================================================

class RollingMean:
    """Mean of the last ``window_size`` observations of a stream (floats or arrays): a fixed ring of
    slots and a running sum, so an update is O(1) whatever the window."""

    def __init__(self, window_size: int) -> None:
        self._window_size = int(window_size)
        self._slots: List[Any] = []
        self._next = 0
        self._sum: Any = 0

    def update(self, x):
        if len(self._slots) < self._window_size:
            self._slots.append(x)
            self._sum = self._sum + x
        else:
            self._sum = self._sum + x - self._slots[self._next]
            self._slots[self._next] = x
        self._next = (self._next + 1) % self._window_size
        return self._sum / len(self._slots)


class JaxPlannerStatus(Enum):
    """Outcome of one planner update (same members and values as planner.py:2276-2303)."""
    NORMAL = 0
    STOPPING_RULE_REACHED = 1
    NO_PROGRESS = 2
    PRECONDITION_POSSIBLY_UNSATISFIED = 3
    INVALID_GRADIENT = 4
    TIME_BUDGET_REACHED = 5
    ITER_BUDGET_REACHED = 6

    def is_terminal(self) -> bool:
        return self in _TERMINAL_STATUS

    def get_type(self) -> str:
        return _STATUS_TYPE[self]


_TERMINAL_STATUS = frozenset({JaxPlannerStatus.STOPPING_RULE_REACHED, JaxPlannerStatus.INVALID_GRADIENT,
                              JaxPlannerStatus.TIME_BUDGET_REACHED, JaxPlannerStatus.ITER_BUDGET_REACHED})
_STATUS_TYPE = {JaxPlannerStatus.NORMAL: "success", JaxPlannerStatus.STOPPING_RULE_REACHED: "success",
                JaxPlannerStatus.NO_PROGRESS: "warning",
                JaxPlannerStatus.PRECONDITION_POSSIBLY_UNSATISFIED: "warning",
                JaxPlannerStatus.INVALID_GRADIENT: "error",
                JaxPlannerStatus.TIME_BUDGET_REACHED: "info", JaxPlannerStatus.ITER_BUDGET_REACHED: "info"}


class JaxPlannerStoppingRule(metaclass=ABCMeta):
    """Interface of the stopping rules ``optimize_generator`` consults after every update."""

    @abstractmethod
    def reset(self) -> None:
        ...

    @abstractmethod
    def monitor(self, callback: Dict[str, Any]) -> bool:
        ...


class NoImprovementStoppingRule(JaxPlannerStoppingRule):
    """Stops once ``best_return`` has not improved for ``patience`` consecutive updates."""

    def __init__(self, patience: int) -> None:
        self.patience = patience
        self.reset()

    def reset(self) -> None:
        self.callback = None                    # the callback that held the best return so far
        self.iters_since_last_update = 0

    def monitor(self, callback: Dict[str, Any]) -> bool:
        improved = self.callback is None or callback["best_return"] > self.callback["best_return"]
        if improved:
            self.callback = callback
        self.iters_since_last_update = 0 if improved else self.iters_since_last_update + 1
        return self.iters_since_last_update >= self.patience

    def __str__(self) -> str:
        return f"No improvement for {self.patience} iterations"


# =====================================================================================
# utilities other than the mean (planner.py:2161-2247) -- evaluated with torch on the [B]
# return vector only (never on the rollout path)
# =====================================================================================

def entropic_utility(returns: torch.Tensor, beta: float) -> torch.Tensor:
    return -(torch.logsumexp(-beta * returns, 0) - np.log(returns.numel())) / beta


def mean_variance_utility(returns: torch.Tensor, beta: float) -> torch.Tensor:
    return returns.mean() - 0.5 * beta * returns.var(unbiased=False)


def mean_deviation_utility(returns: torch.Tensor, beta: float) -> torch.Tensor:
    return returns.mean() - 0.5 * beta * returns.std(unbiased=False)


def mean_semideviation_utility(returns: torch.Tensor, beta: float) -> torch.Tensor:
    mu = returns.mean()
    return mu - 0.5 * beta * torch.sqrt(torch.clamp(returns - mu, max=0.).square().mean())


def mean_semivariance_utility(returns: torch.Tensor, beta: float) -> torch.Tensor:
    mu = returns.mean()
    return mu - 0.5 * beta * torch.clamp(returns - mu, max=0.).square().mean()


def sharpe_utility(returns: torch.Tensor, risk_free: float = 0.0, eps: float = 1e-12) -> torch.Tensor:
    return (returns.mean() - risk_free) / (returns.std(unbiased=False) + eps)


def var_utility(returns: torch.Tensor, alpha: float) -> torch.Tensor:
    """sj.percentile in its default (hard, linear interpolation) mode."""
    return torch.quantile(returns, alpha)


def cvar_utility(returns: torch.Tensor, alpha: float) -> torch.Tensor:
    var = torch.quantile(returns, alpha)
    return var - torch.relu(var - returns).mean() / alpha


def cvar_ste_utility(returns: torch.Tensor, alpha: float) -> torch.Tensor:
    mask = (returns <= torch.quantile(returns.detach(), alpha)).to(returns.dtype)
    return (mask * returns).sum() / torch.clamp(mask.sum(), min=1.)


def expectile_utility(returns: torch.Tensor, alpha: float, max_iter: int = 30) -> torch.Tensor:
    """planner.py:2213-2220: fixed-point iterations from the mean, differentiated through."""
    val = returns.mean()
    for _ in range(int(max_iter)):
        w = torch.where(returns > val, torch.full_like(returns, alpha),
                        torch.full_like(returns, 1.0 - alpha))
        val = (w * returns).sum() / w.sum()
    return val


def gini_utility(returns: torch.Tensor, gamma: float) -> torch.Tensor:
    """planner.py:2223-2229: mean - gamma * 2 * mean_ij relu(r_j - r_i)."""
    return returns.mean() - gamma * 2.0 * torch.relu(
        returns.unsqueeze(0) - returns.unsqueeze(1)).mean()


# the reference's table (planner.py:2233-2247); 'mean' is handled by the loss kernel
UTILITY_LOOKUP = {
    "mean": torch.mean,
    "mean_var": mean_variance_utility,
    "mean_std": mean_deviation_utility,
    "mean_semivar": mean_semivariance_utility,
    "mean_semidev": mean_semideviation_utility,
    "sharpe": sharpe_utility,
    "entropic": entropic_utility,
    "exponential": entropic_utility,
    "var": var_utility,
    "cvar": cvar_utility,
    "cvar_ste": cvar_ste_utility,
    "expectile": expectile_utility,
    "gini": gini_utility,
}


def _utility_fn(name: str, kwargs: Kwargs) -> Callable[[torch.Tensor], torch.Tensor]:
    if name not in UTILITY_LOOKUP:
        raise NotImplementedError(f"Utility function <{name}> is not supported.")
    fn = UTILITY_LOOKUP[name]
    return lambda r: fn(r, **kwargs)


def get_action_info(compiled, user_bounds, horizon: int):
    """Shapes, box bounds and finite-ness masks of the action fluents (planner.py:423-472):
    shapes[name] = (horizon, *objects); bounds[name] = (None, None) for booleans, (0, n - 1) for
    enums, the RDDL / user box otherwise; cond_lists[name] = [both finite, lower only, upper only,
    neither] as boolean masks (the branches of the box wrap, planner.py:620-628)."""
    rddl = compiled.rddl
    shapes, bounds, cond_lists = {}, {}, {}
    for name in rddl.action_fluents:
        prange = rddl.variable_ranges[name]
        if prange not in ("bool", "int", "real") and prange not in rddl.enum_types:
            raise TypeError(f"Invalid range <{prange}> of action-fluent <{name}>.")
        shape = tuple(rddl.variable_shape(name))
        shapes[name] = (horizon,) + shape
        if prange == "bool":
            bounds[name] = (None, None)
            continue
        if prange in rddl.enum_types:
            lower = np.zeros(shape)
            upper = np.ones(shape) * (len(rddl.type_to_objects[prange]) - 1)
        else:
            lower, upper = rddl.bounds[name]
        lower, upper = (user_bounds or {}).get(name, (lower, upper))
        lower = np.broadcast_to(np.asarray(lower, np.float32), shape).copy()
        upper = np.broadcast_to(np.asarray(upper, np.float32), shape).copy()
        lf, uf = np.isfinite(lower), np.isfinite(upper)
        cond_lists[name] = [lf & uf, lf & ~uf, ~lf & uf, ~lf & ~uf]
        bounds[name] = (lower, upper)
    return shapes, bounds, cond_lists


# =====================================================================================
# the planner
# =====================================================================================

# ---- optax.contrib.reduce_on_plateau (third-party, optax >= 0.2.7; planner.py:2375-2384) ----------
def plateau_kwargs(kw: Dict[str, Any]) -> Dict[str, float]:
    out = {"factor": 0.1, "patience": 10, "rtol": 1e-4, "atol": 0.0, "cooldown": 0,
           "accumulation_size": 1, "min_scale": 0.0}
    extra = set(kw) - set(out)
    if extra:
        raise TypeError(f"reduce_on_plateau_kwargs: unknown arguments {sorted(extra)}")
    out.update(kw)
    return out


def plateau_init(P: int, device) -> torch.Tensor:
    """[P][6] = scale, best_value, plateau_count, cooldown_count, count, avg_value."""
    st = torch.zeros(P, 6, dtype=torch.float32, device=device)
    st[:, 0] = 1.0
    st[:, 1] = math.inf
    return st


class JaxBackpropPlanner:
    """Gradient-based planner over batched differentiable rollouts (planner.py:2395)."""

    def __init__(self, rddl, plan: JaxPlan, batch_size_train: int = 32,
                 batch_size_test: Optional[int] = None, rollout_horizon: Optional[int] = None,
                 parallel_updates: int = 1, steps_per_update: int = 1,
                 action_bounds: Optional[Dict[str, Tuple[Any, Any]]] = None,
                 optimizer: Union[str, Callable] = "rmsprop",
                 optimizer_kwargs: Optional[Kwargs] = None, clip_grad: Optional[float] = None,
                 line_search_kwargs: Optional[Kwargs] = None, noise_kwargs: Optional[Kwargs] = None,
                 ema_decay: Optional[float] = None, reduce_on_plateau_kwargs: Optional[Kwargs] = None,
                 pgpe: Optional[Any] = None,
                 compiler=logic.DefaultJaxRDDLCompilerWithGrad,
                 compiler_kwargs: Optional[Kwargs] = None, use_symlog_reward: bool = False,
                 utility: Union[Callable, str] = "mean", utility_kwargs: Optional[Kwargs] = None,
                 start_entropy_coeff: float = 1e-1, end_entropy_coeff: float = 1e-5,
                 critic: Optional[Callable] = None, logger: Optional[Any] = None,
                 dashboard: Optional[Any] = None, dashboard_viz: Optional[Any] = None,
                 preprocessor: Optional[Any] = None, python_functions: Optional[Dict] = None,
                 cache_full_train_rollouts: bool = False, cache_full_test_rollouts: bool = False,
                 device: Optional[Union[str, torch.device]] = None, use_cuda_graph: bool = True
                 ) -> None:
        self.noise_kwargs = dict(noise_kwargs) if noise_kwargs is not None else None
        self.ema_decay = None if ema_decay is None else float(ema_decay)
        self.reduce_on_plateau_kwargs = None if reduce_on_plateau_kwargs is None \
            else plateau_kwargs(reduce_on_plateau_kwargs)
        # optax.scale_by_zoom_linesearch (planner.py:2373-2374): host-driven trials, core/linesearch.py
        self.line_search_kwargs = None if line_search_kwargs is None else zoom_kwargs(line_search_kwargs)
        # terminal value function (planner.py:2721-2745).  The reference takes a JAX callable; this build has
        # no JAX, so the critic is a TORCH callable with the same arguments, batched over the rollouts:
        #   critic(critic_params, obs: {fluent: [B, *objects]}, action: {fluent: [*objects]}) -> [B]
        # evaluated on the final state of every rollout and on the plan's step-0 action, differentiated by
        # torch autograd; its state cotangent seeds the adjoint kernel (straight-line plans, one GPU, no graph)
        if critic is not None and not callable(critic):
            raise TypeError("critic must be a callable (critic_params, obs, action) -> [B] values")
        self.critic_fn = critic
        self.critic_params = None
        for name, val in (("dashboard", dashboard),):
            if val is not None:
                raise NotImplementedError(f"{name} is not supported by the B200 planner yet")
        self.use_symlog_reward = bool(use_symlog_reward)
        self.preprocessor = preprocessor
        self.rddl = rddl
        self.plan = plan
        self.batch_size_train = batch_size_train
        self.batch_size_test = batch_size_train if batch_size_test is None else batch_size_test
        self.steps_per_update = steps_per_update
        self.parallel_updates = parallel_updates
        self.horizon = rddl.horizon if rollout_horizon is None else rollout_horizon
        self._action_bounds = action_bounds or {}
        opt_name = optimizer if isinstance(optimizer, str) else getattr(optimizer, "__name__", "")
        if opt_name not in OPTIMIZERS:
            raise NotImplementedError(f"optimizer <{opt_name}> (supported: {list(OPTIMIZERS)})")
        self.optimizer_name = opt_name
        self.optimizer_kwargs = {"learning_rate": 0.1} if optimizer_kwargs is None else dict(optimizer_kwargs)
        self.clip_grad = clip_grad
        # NOTE: the reference's constructor default is pgpe=GaussianPGPE(); every shipped
        # straight-line config passes pgpe=None, and so does this default
        self.pgpe = pgpe
        self.use_pgpe = pgpe is not None
        if isinstance(utility, str):
            utility = utility.lower()
            if utility not in UTILITY_LOOKUP:
                raise NotImplementedError(
                    f"Utility function <{utility}> is not supported, must be one of "
                    f"{list(UTILITY_LOOKUP.keys())}.")
        self.utility = utility
        self.utility_kwargs = utility_kwargs or {}
        self._native_utility = None          # (kind, p0, p1) of rk_utility_loss, set in _allocate
        self.compiler_type = compiler
        self.compiler_kwargs = compiler_kwargs or {}
        self.start_entropy_coeff = start_entropy_coeff
        self.end_entropy_coeff = end_entropy_coeff
        self.logger = logger
        self.cache_full_train_rollouts = cache_full_train_rollouts
        self.cache_full_test_rollouts = cache_full_test_rollouts
        self.use_cuda_graph = use_cuda_graph
        if not torch.cuda.is_available():
            raise engine.RkError("JaxBackpropPlanner needs a CUDA device: the rollout kernels "
                                 "have no CPU fallback")
        self.device = torch.device(device if device is not None else
                                   f"cuda:{torch.cuda.current_device()}")
        # multi-GPU: every rank owns batch_size_train rollouts; the plan is replicated and
        # the gradient is sum-all-reduced once per iteration
        self.rank, self.world_size = parallel.world()
        # every utility of the reference's table has a closed-form cotangent and runs as native kernels
        # (csrc/rk_optim.cu:rk_utility_loss, csrc/rk_utility.cu); a utility is a function of the GLOBAL
        # return vector, so with several GPUs the shards are all-gathered first -- inside the
        # iteration's graph -- and every rank keeps its own slice of the cotangent
        if isinstance(self.utility, str):
            kw = self.utility_kwargs
            if self.utility == "sharpe" and set(kw) <= {"risk_free", "eps"}:
                self._native_utility = (self.utility, float(kw.get("risk_free", 0.0)), float(kw.get("eps", 1e-12)))
            elif self.utility in engine.UTILITY_KINDS and set(kw) == {"beta"}:
                self._native_utility = (self.utility, float(kw["beta"]), 0.0)
            elif self.utility in ("var", "cvar", "cvar_ste") and set(kw) == {"alpha"}:
                self._native_utility = (self.utility, float(kw["alpha"]), 0)
            elif self.utility == "expectile" and "alpha" in kw and set(kw) <= {"alpha", "max_iter"}:
                self._native_utility = (self.utility, float(kw["alpha"]), int(kw.get("max_iter", 30)))
            elif self.utility == "gini" and set(kw) == {"gamma"}:
                self._native_utility = (self.utility, float(kw["gamma"]), 0)
        self._utility_eager = self.utility != "mean" and self._native_utility is None
        self._compile()

    # ---------------------------------------------------------------------------------
    def _compile(self):
        if getattr(self.plan, "kind", "slp") == "drp" and getattr(self.rddl, "observ_fluents", None):
            # a closed-loop policy over a partially observed model reads the observ-fluents of the
            # previous step (planner.py:1300-1313): they travel in the state vector as carried copies
            from ..rddl.model import observation_view
            self.rddl_observed = self.rddl
            self.rddl = observation_view(self.rddl)
        rddl = self.rddl
        self.compiled = self.compiler_type(rddl, **self.compiler_kwargs)
        self.print_warnings = getattr(self.compiled, "print_warnings", True)
        self.compiled.compile()
        self.test_compiled = JaxRDDLCompiler(rddl)
        self.test_compiled.compile()
        # only the DRP input layer reads the preprocessor (planner.py:1293-1296)
        preprocessor = self.preprocessor if getattr(self.plan, "kind", "slp") == "drp" else None
        self.plan.compile(self.compiled, self.test_compiled, self._action_bounds, self.horizon,
                          preprocessor=preprocessor)
        spec = self.plan.spec
        self.gamma = float(rddl.discount)
        # the programs read the per-step weights of the rewards from a buffer only when asked to: an
        # undiscounted model compiles without it unless a reward mask needs it (set_reward_mask)
        gamma = self.gamma if (self.gamma != 1.0 or not getattr(self, "_weighted_rewards", False)) else 0.5
        train_log = tuple(rddl.state_fluents) if self.cache_full_train_rollouts else ()
        test_log = (tuple(rddl.next_state.values()) + tuple(rddl.action_fluents)) \
            if self.cache_full_test_rollouts else ()
        self.is_drp = getattr(self.plan, "kind", "slp") == "drp"
        if self.critic_fn is not None:
            plan = self.plan
            if self.is_drp or self.world_size > 1 or self.parallel_updates > 1 or self.use_pgpe \
                    or getattr(plan, "_stochastic", False) or getattr(plan, "_wrap_non_bool", False) \
                    or getattr(plan, "_wrap_softmax", False):
                raise NotImplementedError(
                    "critic: built for deterministic straight-line plans (sigmoid-wrapped booleans, plain "
                    "real / int parameters) on one GPU, parallel_updates = 1, pgpe = None")
        if self.is_drp and self._utility_eager:
            raise NotImplementedError(
                "deep reactive policies take the built-in utilities by name (native cotangent kernels); a "
                "callable utility is differentiated by torch autograd on the straight-line path only")
        with torch.cuda.device(self.device):
            tb = self.compiled.builder(spec, True, (), self.batch_size_train)
            if self.is_drp:
                # closed loop: one decision epoch per launch, the MLP between the launches; the
                # state slot passed as s0 is the tape entry, the adjoint chains through GSIN/GS0
                self.train_fwd = engine.DeviceProgram(tb.forward(write_tape=False, gamma=gamma))
                self.train_bwd = engine.DeviceProgram(
                    tb.backward(gamma=gamma, grad_s0=True, state_ct_in=True))
            else:   # adjoint first: it decides what the forward program tapes for it
                bwd_prog = tb.backward(gamma=gamma, state_ct_in=self.critic_fn is not None)
                self.train_fwd = engine.DeviceProgram(tb.forward(write_tape=True, gamma=gamma))
                self.train_bwd = engine.DeviceProgram(bwd_prog)
            eb = self.test_compiled.builder(spec, False, test_log, self.batch_size_test)
            self.test_fwd = engine.DeviceProgram(eb.forward(write_tape=False, gamma=gamma))
        self.A = self.plan.A
        self.T = self.horizon
        self.n_params = self.plan.n_params if self.is_drp else self.T * self.A
        self._alloc()
        self.drp = DrpPipeline(self) if self.is_drp else None
        # JaxPlan.projection / initializer / test_policy (planner.py:386-420) resolve to the planner's seams
        self.plan.projection, self.plan.initializer = self.projection, self.initializer
        self.plan.test_policy = self.test_policy
        if self.use_pgpe:
            self.pgpe.compile(self)

    def _alloc(self):
        dev, T, A = self.device, self.T, self.A
        Bt, Be = self.batch_size_train, self.batch_size_test
        P = self.parallel_updates
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)   # noqa: E731
        fw, bw, te = self.train_fwd.program, self.train_bwd.program, self.test_fwd.program
        NP = self.n_params
        self.params = z(P, NP)
        n_state = {"sgd": 1, "rmsprop": 1, "adam": 2}[self.optimizer_name]
        self.opt_state = z(P, n_state * NP)
        self.grad = z(P, NP)
        self.nwarps = self.train_bwd.num_warps(Bt)
        self.gradp = None if self.is_drp else z(self.nwarps * T * A)
        self.tape = None if self.is_drp else z(T * fw.tape_words * Bt)
        self.train_reward = z(P, T * Bt)
        self.train_returns = z(P, Bt)
        self.train_err = z(P, T * Bt, dt=torch.int32)
        # the three per-iteration scalars the host reads (planner.py:3261-3262, 2916) share one
        # buffer so that a planner iteration costs a single device->host copy
        self._stats = z(3 * P, dt=torch.int32)
        self._stats_host = torch.zeros(3 * P, dtype=torch.int32).pin_memory()
        self.train_loss = self._stats[0:P].view(torch.float32)
        self.gret = z(Bt)
        if self.critic_fn is not None:
            self.train_final, self.gs_in = z(fw.S * Bt), z(fw.S * Bt)
        if self._native_utility is not None:
            if self.world_size > 1:
                self._ret_all, self._ret_all_test = z(self.world_size * Bt), z(self.world_size * Be)
            if self._native_utility[0] in engine.UTILITY_SORTED_KINDS:
                self._utility_work = engine.utility_sorted_work(self.world_size * max(Bt, Be), dev)
        self.test_reward = z(P, T * Be)
        self.test_returns = z(P, Be)
        self.test_err = z(P, T * Be, dt=torch.int32)
        self.test_loss_buf = self._stats[P:2 * P].view(torch.float32)
        self.test_final = z(P, te.S * Be)
        self.test_traj = z(P, T * te.TRAJ * Be) if te.TRAJ else None
        self.zero_flag = self._stats[2 * P:3 * P]
        self.converged = torch.ones(P, dtype=torch.int32, device=dev)
        self.train_noise = z(T * fw.N * Bt) if fw.N else None
        self.test_noise = z(T * te.N * Be) if te.N else None
        self.disc = None if (self.gamma == 1.0 and not getattr(self, "_weighted_rewards", False)) \
            else torch.pow(torch.tensor(self.gamma, device=dev),
                           torch.arange(T, device=dev)).float().contiguous()
        lo, hi = self.plan.box
        self.box_lo = None if lo is None else torch.from_numpy(lo.reshape(-1)).to(dev).contiguous()
        self.box_hi = None if hi is None else torch.from_numpy(hi.reshape(-1)).to(dev).contiguous()
        kw = self.optimizer_kwargs
        lr = float(kw.get("learning_rate", 0.1))
        if self.optimizer_name == "rmsprop":
            h = [lr, kw.get("decay", 0.9), kw.get("eps", 1e-8), 0., 0., 0., 1.]
        elif self.optimizer_name == "adam":
            h = [lr, kw.get("b1", 0.9), kw.get("eps", 1e-8), kw.get("b2", 0.999), 0., 0., 1.]
        else:
            h = [lr, 0., 0., 0., 0., kw.get("momentum", 0.) or 0., 1.]
        h[4] = float(self.clip_grad) if self.clip_grad is not None else 0.
        # optax forms (1 - decay) from its Python-float rates in double precision: pass the complements
        h += [1.0 - float(h[1]), 1.0 - float(h[3])]
        self.hyper = torch.tensor(h, dtype=torch.float32, device=dev)
        self.train_mparams = torch.from_numpy(fw.mparam_defaults.copy()).to(dev) \
            if len(fw.mparam_defaults) else None
        import os as _os
        self._graph = None
        self._graph_host = None
        # the noise draws (torch's Philox generator, registered with the graph) replay with the
        # iteration instead of costing a dozen eager launches per step; RK_NOISE_IN_GRAPH=0 disables
        self._noise_in_graph = (_os.environ.get("RK_NOISE_IN_GRAPH", "1") != "0" and not self.is_drp
                                and (self.train_noise is not None or self.test_noise is not None))
        self._primed = False
        # software pipelining of the captured iteration (opt -> test || next gradient).  All NCCL
        # calls stay on the main branch (gradient all-reduce in the train stage, loss all-reduce
        # after the join); the DRP branches share scratch buffers, so only straight-line plans
        # run the two branches concurrently
        self._pipeline = _os.environ.get("RK_PIPELINE", "1") != "0"
        self._overlap = self._pipeline and not self.is_drp
        self._train_loss_tmp = z(P)
        # parallel_updates (planner.py:2920-2929 vmaps the update over P optimiser instances): the
        # instances are independent kernel chains, so the captured iteration runs them as P concurrent
        # branches, each with its own tape / partials / cotangent scratch.  Deep policies (shared MLP
        # scratch), several GPUs and steps_per_update > 1 (shared step counter) keep the instances in order.
        self._branches = (P > 1 and not self.is_drp and self.world_size == 1
                          and self.steps_per_update <= 1
                          and _os.environ.get("RK_PARALLEL_BRANCHES", "1") != "0")
        if self._branches:
            self._scratch = [(self.tape, self.gradp, self.gret)] + [
                (z(T * fw.tape_words * Bt), z(self.nwarps * T * A), z(Bt)) for _ in range(P - 1)]
        self._use_chain = (self.noise_kwargs is not None or self.ema_decay is not None
                           or self.reduce_on_plateau_kwargs is not None)
        # the tail of the adjoint fused with the optimiser (csrc/rk_optim.cu:rk_reduce_optimizer_step): in
        # the pipelined order the partial rows of a replay's adjoint are reduced by the NEXT replay's
        # optimiser kernel -- one launch less on the critical chain.  Needs the plain chain: one GPU (the
        # exchange wants the reduced gradient), no global-norm clipping, no extra links, K = P = 1, no term
        # added to the gradient after the reduction (stochastic plans, critic)
        self._fuse_tail = (self._pipeline and not self.is_drp and self.world_size == 1 and P == 1
                           and self.steps_per_update <= 1 and self.clip_grad is None
                           and not self._use_chain and self.line_search_kwargs is None
                           and self.critic_fn is None and not getattr(self.plan, "_stochastic", False)
                           and not self._utility_eager
                           and _os.environ.get("RK_FUSE_TAIL", "1") != "0")
        self._tail_counters = torch.zeros(2, dtype=torch.int32, device=dev)
        self._defer_reduce = False
        if self.reduce_on_plateau_kwargs is not None and self.world_size > 1:
            raise NotImplementedError("reduce_on_plateau_kwargs with more than one GPU: the plateau "
                                      "test needs the all-reduced loss before the optimiser step")
        if self._use_chain:
            self._chain_count = z(P)
            self._chain_plateau = plateau_init(P, dev)
            self._chain_ema = z(P, NP)
            self._chain_scratch = z(P, 8)
        # multi-GPU exchange: one message per iteration.  In the pipelined order the
        # gradient is not needed before the NEXT replay's optimiser step, so it travels together
        # with the two loss sums after the branches join: [grad | train loss | test loss], summed
        # by the one-shot NVLink peer kernel (csrc/rk_comm.cu) or, failing that, NCCL.
        self._peer_ar = None
        self._fold_msg = False
        self._defer_ar = False
        if self.world_size > 1:
            self._fold_msg = (self._pipeline and P == 1 and self.steps_per_update <= 1)
            self._msg = z(NP + 4)
            if self._fold_msg:
                self.grad = self._msg[:NP].view(1, NP)
            self._peer_ar = parallel.PeerAllReduce.create(NP + 4, self.device)
        self.ent_coeff = torch.full((1,), float(self.start_entropy_coeff), device=self.device)
        if self.use_pgpe:
            self._pg_reward, self._pg_returns = z(T * Be), z(Be)
            self._pg_err, self._pg_final = z(T * Be, dt=torch.int32), z(te.S * Be)
            self._pg_loss = z(P)
        self.s0_train, self.s0_test = z(fw.S * Bt), z(te.S * Be)
        self._init_train, self._init_test = z(fw.S), z(te.S)
        self.grad_used = z(P, NP)
        self._gen = torch.Generator(device=dev)
        # counter-based noise (one launch per program and redraw): seed and the device-side draw number
        self._philox = _os.environ.get("RK_PHILOX", "1") != "0" and torch.device(dev).type == "cuda"
        self._noise_seed = 42
        self._draws = torch.zeros(1, dtype=torch.int64, device=dev)
        # [train loss | test loss] sums over the ranks, followed by the two loop-control words every
        # rank votes on (time budget reached, progress): they travel with the gradient exchange, so all
        # replicas leave optimize_generator at the same iteration and anneal the same entropy coefficient
        self._pair_ctl = z(2 * P + 2)
        self.loss_pair = self._pair_ctl[:2 * P].view(2, P)
        self.ctl = self._pair_ctl[2 * P:]
        self.ctl_in = z(2)
        self._ctl_host = torch.zeros(2, dtype=torch.float32).pin_memory() \
            if torch.cuda.is_available() else torch.zeros(2, dtype=torch.float32)

    # ---- reward mask (planner_state.reward_mask, planner.py:2802-2809) ------------------------------
    def set_reward_mask(self, mask) -> None:
        """Per-step weights of the rewards inside the loss (train and test): rewards[:, t] * mask[t] before
        discounting.  None removes the mask.  The weights fold into the discount vector the rollout
        kernels already take, so nothing else changes; plans with an entropy term (stochastic=True) would
        need the same weights on their per-step entropies and are rejected."""
        T = self.T
        base = torch.pow(torch.tensor(float(self.gamma), dtype=torch.float64),
                         torch.arange(T, dtype=torch.float64))
        if mask is None:
            self.disc = None if (self.gamma == 1.0 and not getattr(self, "_weighted_rewards", False)) \
                else base.float().to(self.device).contiguous()
        else:
            m = torch.as_tensor(np.asarray(mask), dtype=torch.float64)
            if m.dim() != 1:
                raise ValueError("Reward mask must be 1-dimensional.")
            if m.numel() != T:
                raise ValueError("Reward mask size must match reward dimension.")
            if getattr(self.plan, "_stochastic", False):
                raise NotImplementedError("reward_mask with a stochastic plan (per-step entropy weights)")
            if self.gamma == 1.0 and not getattr(self, "_weighted_rewards", False):
                # an undiscounted model was compiled without the weight buffer: rebuild the programs with it
                # (this re-creates the planner's device state, like a new planner: call before initialize)
                self._weighted_rewards = True
                if getattr(self, "rddl_observed", None) is not None:
                    self.rddl = self.rddl_observed
                self._compile()
            self.disc = (base * m).float().to(self.device).contiguous()
        self.reward_mask = mask
        self._graph = self._graph_host = None        # the discount vector is baked into the captured launches
        self._primed = False

    # ---- model parameters (relaxation weights), planner.py:3165-3168 ---------------------
    def model_parameter_info(self) -> Dict[Any, float]:
        fw = self.train_fwd.program
        return {k: float(v) for k, v in zip(fw.mparam_keys, fw.mparam_defaults)}

    def set_model_params(self, values: Dict[Any, float]):
        fw = self.train_fwd.program
        cur = self.train_mparams.cpu().numpy().copy()
        for i, k in enumerate(fw.mparam_keys):
            if k in values:
                cur[i] = values[k]
        self.train_mparams.copy_(torch.from_numpy(cur))

    # ---- state initialisation (init_sim_state, compiler.py:729-771) ------------------------
    def _init_words(self, layout, subs) -> torch.Tensor:
        """The un-batched initial state as [S] words in the kernels' layout (pinned host)."""
        vals = {}
        for name in layout:
            v = np.asarray(self.rddl.init_values[name])
            key = name[:-len("@seen")] if name.endswith("@seen") else name     # carried observation
            if subs is not None and key in subs:
                v = np.reshape(np.asarray(subs[key]), v.shape).astype(v.dtype)
            t = torch.from_numpy(np.ascontiguousarray(v))
            vals[name] = t.unsqueeze(0)
        return pack_fluent_major(layout, vals, 1).pin_memory()

    def load_initial_state(self, train_words: torch.Tensor, test_words: torch.Tensor):
        """Host -> device transfer of the S initial-state words of each model and on-device
        tiling into the [B] rollouts (init_sim_state, compiler.py:729-771): the per-call input
        path of optimize() and of every online-replanning step."""
        for words, dev_words, prog, B, s0 in (
                (train_words, self._init_train, self.train_fwd.program, self.batch_size_train,
                 self.s0_train),
                (test_words, self._init_test, self.test_fwd.program, self.batch_size_test,
                 self.s0_test)):
            dev_words.copy_(words, non_blocking=True)
            segs = [(off, n) for (off, n, _) in prog.state_layout.values()]
            engine.tile_state(B, segs, dev_words, s0)

    def _draw(self, prog, T, B, out, tag: int = 0):
        """Fresh noise for every draw site of ``prog``: one Philox launch (csrc/rk_optim.cu:rk_draw_noise,
        keyed by the optimize() key, this rank and ``tag``; the draw number advances once per redraw) --
        or, with RK_PHILOX=0, the torch-generator form the parity tests use (core/layout.py:draw_noise)."""
        if out is None:
            return
        sites = [(kind, off, n) for kind, _, n, off, _ in prog.noise_layout]
        if self._philox and all(k in engine.NOISE_KINDS for k, _, _ in sites) and len(sites) <= 64:
            engine.draw_noise(T, B, sites, self._noise_seed * 4 + tag, self._draws, out,
                              self._gamma_conc(prog))
            return
        out.copy_(draw_noise(prog.noise_layout, T, B, self._gen, self.device).reshape(-1))

    def _philox_all(self) -> bool:
        """Every draw site of the train and test programs is served by the counter-based kernel."""
        def ok(prog):
            kinds = [k for k, *_ in prog.noise_layout]
            return all(k in engine.NOISE_KINDS for k in kinds) and len(kinds) <= 64
        return bool(self._philox and ok(self.train_fwd.program) and ok(self.test_fwd.program))

    def _gamma_conc(self, prog) -> Optional[torch.Tensor]:
        """[N] concentrations of the program's Gamma draw sites (build-time shapes), uploaded
        once; None when the program draws no Gamma noise."""
        cache = self.__dict__.setdefault("_conc_cache", {})
        key = id(prog)
        if key not in cache:
            conc = None
            for kind, _, n, off, extra in prog.noise_layout:
                if kind == "gamma":
                    if conc is None:
                        conc = np.ones(prog.N, dtype=np.float32)
                    conc[off:off + n] = np.maximum(
                        np.broadcast_to(np.asarray(extra, np.float32).reshape(-1), (n,)), 1e-6)
            cache[key] = None if conc is None else torch.from_numpy(conc).to(self.device)
        return cache[key]

    # ---- the two device calls of an iteration -----------------------------------------------
    def _train_grad_kernels(self, p: int):
        """Relaxed rollouts, loss and its gradient w.r.t. the plan for parallel instance p
        (the value_and_grad of planner.py:2873-2882); the loss goes to a staging word."""
        T, A, B = self.T, self.A, self.batch_size_train
        params = self.params[p]
        loss_out = self._train_loss_tmp[p:p + 1]
        if self.is_drp:
            self.drp.update_kernels(p, loss_out)
        else:
            critic = self.critic_fn is not None
            self.train_fwd.forward(B, T, self.s0_train, policy=params, noise=self.train_noise,
                                   mparams=self.train_mparams, disc=self.disc,
                                   reward=self.train_reward[p], returns=self.train_returns[p],
                                   err=self.train_err[p], tape=self.tape,
                                   final=self.train_final if critic else None)
            if critic:      # returns += gamma^T critic(final state, step-0 action)   (planner.py:2758-2760)
                crit = self._critic_values(params, self.train_final, self.train_fwd.program, B, train=True)
                self.train_returns[p].add_(crit["values"].detach() * crit["scale"])
            main = torch.cuda.current_stream(self.device)
            if getattr(self, "_published", None) is not None:
                main.wait_event(self._published)
            side = getattr(self, "_loss_stream", None)
            if self.utility == "mean" and self.use_symlog_reward:
                # planner.py:2761-2763: returns -> sign(R) log1p(|R|) in the TRAIN loss only;
                # the cotangent of the batch mean becomes -(1 / B) / (1 + |R_b|)
                engine.utility_loss("symlog_mean", B, self.train_returns[p], 0.0, 0.0,
                                    1.0 / self.world_size, loss_out, self.gret)
            elif self._native_utility is not None:
                # closed-form cotangent on the device: no autograd, the iteration stays one CUDA graph
                self._native_utility_kernels(self.train_returns[p], loss_out)
            elif self.utility == "mean" and side is not None \
                    and not getattr(self.plan, "_stochastic", False):
                # the loss value is not an input of the adjoint (its cotangent is the constant
                # -1/B): reduce it on the side stream, off the critical chain
                done = torch.cuda.Event()
                done.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(done)
                    engine.mean_loss(self.train_returns[p], B, loss_out, None)
            elif self.utility == "mean":
                engine.mean_loss(self.train_returns[p], B, loss_out, None)
            else:
                # a utility other than the mean is a function of the GLOBAL return vector
                # (planner.py:2161-2247, 2817-2818): gather the shards, differentiate, keep this rank's
                # slice of the cotangent -- the summed gradients are then exact
                r = parallel.allgather_returns(self.train_returns[p]).detach().clone().requires_grad_(True)
                fn = self.utility if callable(self.utility) else _utility_fn(self.utility, self.utility_kwargs)
                with torch.enable_grad():
                    loss = -fn(r, **(self.utility_kwargs if callable(self.utility) else {}))
                    (g,) = torch.autograd.grad(loss, r)
                loss_out.copy_(loss.detach().reshape(1))
                self.gret.copy_(g[self.rank * B:(self.rank + 1) * B])
            gp0 = None
            if critic:      # d loss / d critic_b = gret_b gamma^T: back through the critic by autograd
                (self.gret * crit["values"] * crit["scale"]).sum().backward()
                self.gs_in.zero_()
                for name, (off, n, _) in self.train_fwd.program.state_layout.items():
                    g = crit["obs"][name].grad
                    if g is not None:
                        self.gs_in[off * B:(off + n) * B].copy_(g.reshape(-1))
                gp0 = crit["row"].grad
            self.train_bwd.backward(B, T, self.tape, self.gret, policy=params,
                                    noise=self.train_noise, mparams=self.train_mparams,
                                    disc=self.disc, gradp=self.gradp,
                                    gs_in=self.gs_in if critic else None)
            if not self._defer_reduce:      # (deferred: the next optimiser launch reduces the rows itself)
                engine.reduce_partials(self.gradp, self.nwarps, T * A, self.grad[p])
            if gp0 is not None:       # the critic also sees the plan's step-0 action
                self.grad[p][:A].add_(gp0)
        if self.world_size > 1 and not self._defer_ar:   # the exchange step of the path
            self._allreduce(self.grad[p])
        if not self.is_drp and getattr(self.plan, "_stochastic", False):
            # entropy regulariser of a stochastic plan (planner.py:2812-2813): the same for every
            # rollout, so value and gradient come from the parameters alone
            H, dH = self.plan.entropy_terms(params.view(T, A))
            loss_out.sub_(self.ent_coeff * H)
            # with the deferred exchange every rank's gradient is still to be SUMMED over the ranks: the
            # regulariser (the same on every rank) goes in once, not world_size times
            scale = 1.0 / self.world_size if (self.world_size > 1 and self._defer_ar) else 1.0
            self.grad[p].view(T, A).sub_(self.ent_coeff * dH * scale)

    def _critic_values(self, params: torch.Tensor, final: torch.Tensor, prog, B: int, train: bool):
        """critic(critic_params, final observation, step-0 action) per rollout (planner.py:2721-2745).
        train: float observations and the relaxed train action, both autograd leaves; test: the exact
        model's typed state and the test policy's action."""
        T, A = self.T, self.A
        shapes = {name: tuple(self.rddl.variable_shape(name)) for name in prog.state_layout}
        obs = unpack_fluent_major(prog.state_layout, final, B, shapes)
        plan = self.plan
        row = params.view(T, A)[0].detach().clone()
        if train:
            obs = {k: v.detach().clone().requires_grad_(True) for k, v in obs.items()}
            row.requires_grad_(True)
            action = {}
            for name, (off, n, prange) in plan.layout.items():
                a = row[off:off + n]
                if prange == "bool" and plan._wrap_sigmoid:            # planner.py:741-745, 818
                    a = torch.sigmoid(float(plan._sigmoid_weight) * a)
                action[name] = a.reshape(self.rddl.variable_shape(name))
        else:
            acts = plan.test_actions(row.cpu().numpy())
            action = {k: torch.as_tensor(np.asarray(v), device=self.device) for k, v in acts.items()}
        with torch.enable_grad():
            values = self.critic_fn(self.critic_params, obs, action)
        values = torch.as_tensor(values, dtype=torch.float32, device=self.device).reshape(-1)
        if values.numel() != B:
            raise ValueError("Critic value must be a scalar per rollout.")
        return {"values": values, "obs": obs, "row": row, "scale": float(self.gamma) ** T}

    def _native_utility_kernels(self, local_returns: torch.Tensor, loss_out: torch.Tensor,
                                test: bool = False):
        """loss = -U(global returns) and -- for the train batch -- this rank's slice of its cotangent
        (planner.py:2161-2247, 2817-2818; the test loss applies the same utility, planner.py:2695-2698).
        Several GPUs: one all-gather of the B returns per rank (4 B bytes, NCCL, captured with the
        iteration); every rank then holds the GLOBAL loss, which the [grad | loss] exchange sums and
        read_stats divides by the world size again."""
        kind, p0, p1 = self._native_utility
        B, ws = (self.batch_size_test if test else self.batch_size_train), self.world_size
        r = local_returns
        if ws > 1:
            import torch.distributed as dist
            r = self._ret_all_test if test else self._ret_all
            dist.all_gather_into_tensor(r, local_returns)
        gret = None if test else self.gret
        if kind in engine.UTILITY_SORTED_KINDS:
            engine.utility_sorted_loss(kind, B * ws, r, p0, int(p1), 1.0, loss_out, gret,
                                       self._utility_work, self.rank * B, 0 if test else B)
        else:
            engine.utility_loss(kind, B * ws, r, p0, p1, 1.0, loss_out, gret, self.rank * B,
                                0 if test else B)

    def _publish_kernels(self, p: int):
        """Publish the loss and the gradient of the update being applied (callback['grad'])."""
        self.train_loss[p:p + 1].copy_(self._train_loss_tmp[p:p + 1])
        self.grad_used[p].copy_(self.grad[p])

    def _opt_kernels(self, p: int, publish: bool = True):
        """optax update + apply_updates + projection (planner.py:2885-2899) from the gradient the
        train stage left in self.grad; publishes that stage's loss."""
        T, A = self.T, self.A
        params = self.params[p]
        if self._defer_reduce:
            engine.reduce_optimizer_step(self.gradp, self.nwarps, T * A, self.grad[p], self.optimizer_name,
                                         self.hyper, params, self.opt_state[p], self.box_lo, self.box_hi,
                                         self.zero_flag[p:p + 1], self._tail_counters)
            if publish:
                self._publish_kernels(p)
        elif publish:
            self._publish_kernels(p)
        if self._defer_reduce:
            pass
        elif self.line_search_kwargs is not None:
            self._line_search_update(p)
        elif not self._use_chain:
            engine.optimizer_step(self.optimizer_name, self.hyper, self.n_params, params,
                                  self.opt_state[p], self.grad[p], self.box_lo, self.box_hi,
                                  self.zero_flag[p:p + 1])
        else:
            self._opt_chain_kernels(p)
        plan = self.plan
        if getattr(plan, "needs_concurrency_projection", False):     # planner.py:947-954
            self.converged[p:p + 1].fill_(1)
            engine.project_slp(plan.projection_method, T, A, params, plan.bool_segs,
                               plan.max_allowed_actions, plan._max_constraint_iter,
                               plan._wrap_sigmoid, plan._min_action_prob,
                               self.converged[p:p + 1])

    def _line_search_update(self, p: int):
        """optimiser -> scale_by_zoom_linesearch -> apply_updates -> box projection (planner.py:2373-2374,
        2886-2899).  The optimiser's update u is the search direction; every trial step size re-runs the
        relaxed rollouts and the adjoint at params + eta u on the noise of this update (the reference
        hands value_fn the same sim_state) and the host reads back (loss, <grad, u>).  Runs outside the
        CUDA graph, in update-then-test order."""
        if self._use_chain or self.world_size > 1:
            raise NotImplementedError("line_search_kwargs together with noise_kwargs / ema_decay / "
                                      "reduce_on_plateau_kwargs or several GPUs")
        params, grad = self.params[p], self.grad[p]
        x0, g0 = params.clone(), grad.clone()
        f0 = float(self._train_loss_tmp[p])
        hyper = self.hyper.clone()
        engine.optimizer_step(self.optimizer_name, hyper, self.n_params, params, self.opt_state[p], grad,
                              None, None, self.zero_flag[p:p + 1])
        u = params - x0
        slope0 = float(torch.dot(g0, u))

        def phi(eta: float):
            params.copy_(x0 + eta * u)
            self._train_grad_kernels(p)
            return float(self._train_loss_tmp[p]), float(torch.dot(self.grad[p], u))

        kw = self.line_search_kwargs
        state = self.__dict__.setdefault("_ls_stepsize", {})
        prev = state.get(p, 1.0)
        init = 1.0 if (kw["initial_guess_strategy"] == "one" or prev <= 0.0) else prev
        eta, f_eta, evals, ok = zoom_linesearch(phi, f0, slope0, init, **kw)
        state[p] = eta
        self.line_search_info = {"learning_rate": eta, "value": f_eta, "num_linesearch_steps": evals,
                                 "converged": ok}
        params.copy_(x0 + eta * u)
        if self.box_lo is not None:
            torch.maximum(params, self.box_lo, out=params)
        if self.box_hi is not None:
            torch.minimum(params, self.box_hi, out=params)
        grad.copy_(g0)                       # what the callback publishes: loss and gradient at x0
        self._train_loss_tmp[p:p + 1].fill_(f0)

    def _opt_chain_kernels(self, p: int):
        """The optional links of the optax chain (planner.py:2366-2386) around the optimiser, one native
        launch (csrc/rk_optim.cu:rk_optimizer_chain_step):
        clip_by_global_norm -> add_noise(eta, gamma) -> optimiser -> reduce_on_plateau -> ema(decay),
        then apply_updates and the box projection.  add_noise: g + N(0, eta / t^gamma) with the
        1-based step count t (counter-based draws keyed by noise_kwargs['seed']); reduce_on_plateau: the
        updates are scaled by a factor that shrinks when the train loss stops improving; ema: the
        debiased exponential moving average of the updates is what is applied."""
        nk = self.noise_kwargs
        cfg = {"noise": None if nk is None else (float(nk.get("eta", 0.01)), float(nk.get("gamma", 0.55))),
               "ema_decay": self.ema_decay, "plateau": self.reduce_on_plateau_kwargs}
        engine.optimizer_chain_step(
            self.optimizer_name, self.hyper, self.n_params, self.params[p], self.opt_state[p],
            self.grad[p], self.box_lo, self.box_hi, self.zero_flag[p:p + 1], cfg,
            self._chain_count[p:p + 1],
            self._chain_plateau[p] if self.reduce_on_plateau_kwargs is not None else None,
            self._chain_ema[p] if self.ema_decay is not None else None,
            self._train_loss_tmp[p:p + 1], int((nk or {}).get("seed", 0)) * 1000003 + p,
            self._chain_scratch[p])

    def _update_kernels(self, p: int):
        """planner.py:2868-2917 on the device for parallel instance p."""
        self._train_grad_kernels(p)
        self._opt_kernels(p)

    def _test_kernels(self, p: int, params: Optional[torch.Tensor] = None, scratch: bool = False,
                      loss_out: Optional[torch.Tensor] = None):
        """planner.py:2695-2698: exact-model rollouts, forward only.  ``scratch`` evaluates a
        candidate plan (PGPE samples) without touching the logged test buffers."""
        T, B = self.T, self.batch_size_test
        loss_out = self.test_loss_buf[p:p + 1] if loss_out is None else loss_out
        if scratch:
            reward, returns, err, final = self._pg_reward, self._pg_returns, self._pg_err, self._pg_final
        else:
            reward, returns, err, final = (self.test_reward[p], self.test_returns[p],
                                           self.test_err[p], self.test_final[p])
        if self.is_drp:
            self.drp.test_kernels(p, params, (reward, returns, err, final, loss_out))
            return
        self.test_fwd.forward(B, T, self.s0_test, policy=self.params[p] if params is None else params,
                              noise=self.test_noise, disc=self.disc,
                              reward=reward, returns=returns, err=err, final=final,
                              traj=None if (self.test_traj is None or scratch) else self.test_traj[p])
        if self.critic_fn is not None:      # the test loss is the same loss function (planner.py:2695-2698)
            crit = self._critic_values(self.params[p] if params is None else params, final,
                                       self.test_fwd.program, B, train=False)
            returns.add_(crit["values"].detach() * crit["scale"])
        self._test_loss_kernels(returns, loss_out)

    def _test_loss_kernels(self, returns: torch.Tensor, loss_out: torch.Tensor):
        """-utility(test returns) (planner.py:2695-2698: the test loss is the same loss function over the
        exact rollouts, without the symlog transform)."""
        if self.utility == "mean":
            engine.mean_loss(returns, self.batch_size_test, loss_out, None)
        elif self._native_utility is not None:
            self._native_utility_kernels(returns, loss_out, test=True)
        else:
            r = parallel.allgather_returns(returns)
            fn = self.utility if callable(self.utility) else _utility_fn(self.utility, self.utility_kwargs)
            loss_out.copy_((-fn(r, **(self.utility_kwargs if callable(self.utility) else {}))).reshape(1))

    def project_params(self, params: torch.Tensor, p: int = 0):
        """plan.projection on a flat parameter vector: box clip (planner.py:878-910) and, when
        the instance limits concurrent actions, the max-nondef-actions projection (914-959)."""
        if self.box_lo is not None:
            torch.maximum(params, self.box_lo, out=params)
            torch.minimum(params, self.box_hi, out=params)
        plan = self.plan
        if getattr(plan, "needs_concurrency_projection", False):
            self.converged[p:p + 1].fill_(1)
            engine.project_slp(plan.projection_method, self.T, self.A, params, plan.bool_segs,
                               plan.max_allowed_actions, plan._max_constraint_iter,
                               plan._wrap_sigmoid, plan._min_action_prob, self.converged[p:p + 1])

    def _iteration_kernels(self):
        for p in range(self.parallel_updates):
            for _ in range(max(1, self.steps_per_update)):      # planner.py:2905-2914
                self._update_kernels(p)
                if self.optimizer_name == "adam" and self.steps_per_update > 1:
                    self.hyper[6:7].add_(1.0)
            self._test_kernels(p)
        self._finish_iteration()

    def _allreduce(self, vec: torch.Tensor):
        if self._peer_ar is not None and vec.numel() <= self._peer_ar.n_max:
            self._peer_ar(vec)
        else:
            parallel.allreduce_gradient(vec)

    def _finish_iteration(self):
        if self.optimizer_name == "adam" and self.steps_per_update <= 1:
            self.hyper[6:7].add_(1.0)
        if self.world_size > 1:
            if self._defer_ar and self._peer_ar is not None and self._msg.numel() <= self._peer_ar.n_max:
                # [grad | train loss | test loss | votes] gathered by the exchange kernel itself
                self._peer_ar.gather(self._msg[:self.n_params],
                                     [self.train_loss, self.test_loss_buf, self.ctl_in], self._pair_ctl)
            elif self._defer_ar:        # [grad | train loss | test loss] in one message
                NP = self.n_params
                self._msg[NP:NP + 1].copy_(self.train_loss)
                self._msg[NP + 1:NP + 2].copy_(self.test_loss_buf)
                self._msg[NP + 2:].copy_(self.ctl_in)
                self._allreduce(self._msg)
                self._pair_ctl.copy_(self._msg[NP:])
            else:
                self.loss_pair[0].copy_(self.train_loss)
                self.loss_pair[1].copy_(self.test_loss_buf)
                self.ctl.copy_(self.ctl_in)
                self._allreduce(self._pair_ctl)

    def _pipelined_kernels(self, side: Optional[torch.cuda.Stream]):
        """One iteration in software-pipelined order: the optimiser step for the gradient the
        previous replay left behind, then -- concurrently when ``side`` is given -- the exact
        test rollouts of the new plan and the relaxed rollouts + gradient for the NEXT step.
        Both read the plan only, and at B = 4096 each fills a fraction of the machine
        (one warp per SM sub-partition), so they overlap almost perfectly.  Results per
        iteration are identical to update-then-test (planner.py:3245-3260)."""
        P = self.parallel_updates
        self._defer_reduce = self._fuse_tail
        # fresh noise of this replay, drawn where it is consumed: the train tensor on the main branch
        # (next to the optimiser step, which it does not depend on), the test tensor at the head of the
        # test branch -- the two counter-based launches overlap instead of preceding the iteration
        draw = getattr(self, "_draw_in_body", False)
        if draw:
            self._draw(self.train_fwd.program, self.T, self.batch_size_train, self.train_noise, 0)
            if side is None:
                self._draw(self.test_fwd.program, self.T, self.batch_size_test, self.test_noise, 1)
        # the two small copies that publish the previous stage's loss / gradient only READ what the
        # optimiser reads, so with a side stream they leave the critical chain (K = 1 only: with
        # more updates per iteration the intermediate stages overwrite what they read)
        off_chain = side is not None and self.steps_per_update <= 1
        for p in range(P):
            self._opt_kernels(p, publish=not off_chain)
            for _ in range(max(1, self.steps_per_update) - 1):   # K updates per iteration
                if self.optimizer_name == "adam":
                    self.hyper[6:7].add_(1.0)
                self._train_grad_kernels(p)
                self._opt_kernels(p)
            if self.optimizer_name == "adam" and self.steps_per_update > 1:
                self.hyper[6:7].add_(1.0)
        main = torch.cuda.current_stream(self.device)
        published = None
        if side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                if draw:
                    self._draw(self.test_fwd.program, self.T, self.batch_size_test, self.test_noise, 1)
                if off_chain:
                    for p in range(P):
                        self._publish_kernels(p)
                    published = torch.cuda.Event()
                    published.record(side)
                for p in range(P):
                    self._test_kernels(p)
        else:
            for p in range(P):
                self._test_kernels(p)
        self._defer_ar = self._fold_msg
        self._published = published      # the train stage overwrites what the copies read
        self._loss_stream = side if off_chain else None
        for p in range(P):
            self._train_grad_kernels(p)
        self._published = None
        self._loss_stream = None
        if side is not None:
            main.wait_stream(side)
        if draw:
            engine.noise_advance(self._draws)
        self._finish_iteration()
        self._defer_ar = False
        self._defer_reduce = False

    def _use_scratch(self, p: int):
        if self._branches:
            self.tape, self.gradp, self.gret = self._scratch[p]

    def _branch_kernels(self):
        """One iteration with parallel_updates > 1: instance p runs `optimiser step -> exact test rollouts
        -> relaxed rollouts + gradient for the next step` on its own stream, i.e. as its own branch of the
        captured graph; the branches join before the statistics are read.  Per-instance results are
        those of the sequential order (the instances share nothing but read-only inputs)."""
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_pstreams", None) is None:
            self._pstreams = [torch.cuda.Stream(device=self.device) for _ in range(self.parallel_updates)]
        for p, sp in enumerate(self._pstreams):
            sp.wait_stream(main)
            with torch.cuda.stream(sp):
                self._use_scratch(p)
                self._opt_kernels(p)
                self._test_kernels(p)
                self._train_grad_kernels(p)
        for sp in self._pstreams:
            main.wait_stream(sp)
        self._use_scratch(0)
        self._finish_iteration()

    def _capture(self, from_host: bool = False):
        """Capture one full iteration into a CUDA graph (after a warm-up on a side stream).
        ``from_host``: the graph also carries the host<->device copies of one optimize() call
        (initial-state words and plan in, status words and plan out, through pinned staging
        buffers), so a whole replanning step is ONE graph launch and ONE synchronisation."""
        backup = (self.params.clone(), self.opt_state.clone(), self.hyper.clone())
        chain = [t.clone() for t in (self._chain_count, self._chain_ema, self._chain_plateau)] \
            if hasattr(self, "_chain_count") else None
        s = torch.cuda.Stream(device=self.device)
        side = torch.cuda.Stream(device=self.device) if self._overlap else None
        inner0 = (lambda: self._pipelined_kernels(side)) if self._pipeline else self._iteration_kernels
        if self._branches and self._pipeline:
            inner0 = self._branch_kernels
        self._draw_in_body = False
        if self._noise_in_graph and self._philox_all() and self._pipeline \
                and not (self._branches and self._pipeline):
            self._draw_in_body = True            # the pipelined body draws its own noise (see there)
            inner = inner0
        elif self._noise_in_graph:
            def inner():
                self.redraw_noise(force=True)
                inner0()
        else:
            inner = inner0
        if from_host:
            pin = lambda t: torch.empty_like(t, device="cpu").pin_memory()   # noqa: E731
            self._h_init_train, self._h_init_test = pin(self._init_train), pin(self._init_test)
            self._h_params_in, self._h_params_out = pin(self.params), pin(self.params)
            self._h_pair = pin(self.loss_pair)
            self._h_init_train.copy_(self.init_train_host)
            self._h_init_test.copy_(self.init_test_host)
            self._h_params_in.copy_(self.params)

            def body():
                self.load_initial_state(self._h_init_train, self._h_init_test)
                self.params.copy_(self._h_params_in, non_blocking=True)
                inner()
                self._h_params_out.copy_(self.params, non_blocking=True)
                self._stats_host.copy_(self._stats, non_blocking=True)
                if self.world_size > 1:
                    self._h_pair.copy_(self.loss_pair, non_blocking=True)
        else:
            body = inner
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            body()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        if self._noise_in_graph:
            g.register_generator_state(self._gen)
        with torch.cuda.graph(g):
            body()
        self.params.copy_(backup[0])
        self.opt_state.copy_(backup[1])
        self.hyper.copy_(backup[2])
        if chain is not None:
            self._chain_count.copy_(chain[0])
            self._chain_ema.copy_(chain[1])
            self._chain_plateau.copy_(chain[2])
        if from_host:
            self._graph_host = g
        else:
            self._graph = g
        self._primed = False

    def iterate_from_host(self, init_train: torch.Tensor, init_test: torch.Tensor,
                          params: torch.Tensor):
        """One replanning step driven from host buffers (the per-call path of optimize() and of
        JaxOnlineController): initial-state words of both models and the plan to start from go
        in, (train loss, test loss, zero-gradient flags, plan) come back -- one graph launch and
        one synchronisation.  Returns (train [P], test [P], zero [P], plan [P, NP] pinned view)."""
        if self._utility_eager or self.line_search_kwargs is not None or self.critic_fn is not None \
                or not self.use_cuda_graph:
            self.load_initial_state(init_train, init_test)
            self.params.copy_(params, non_blocking=True)
            self.iterate()
            train, test, zero = self.read_stats()
            return train, test, zero, self.params.cpu()
        if getattr(self, "_graph_host", None) is None:
            self._capture(from_host=True)
        self._h_init_train.copy_(init_train)
        self._h_init_test.copy_(init_test)
        self._h_params_in.copy_(params)
        if self._pipeline and not self._primed:
            self.load_initial_state(self._h_init_train, self._h_init_test)
            self.params.copy_(self._h_params_in, non_blocking=True)
            self._defer_reduce = self._fuse_tail
            for p in range(self.parallel_updates):     # gradient for the first optimiser step
                self._use_scratch(p)
                self._train_grad_kernels(p)
            self._use_scratch(0)
            self._defer_reduce = False
            self._primed = True
        self._graph_host.replay()
        torch.cuda.current_stream(self.device).synchronize()
        P = self.parallel_updates
        raw = self._stats_host.numpy()
        train = raw[0:P].view(np.float32).copy()
        test = raw[P:2 * P].view(np.float32).copy()
        zero = raw[2 * P:3 * P] > 0
        if self.world_size > 1:
            pair = self._h_pair.numpy() / self.world_size
            train, test = pair[0].copy(), pair[1].copy()
        return train, test, zero, self._h_params_out

    def iterate(self):
        """One planner iteration on the device: update + test_loss (planner.py:3245-3260)."""
        if self._utility_eager or self.line_search_kwargs is not None or self.critic_fn is not None \
                or not self.use_cuda_graph:
            self._iteration_kernels()
            return
        if self._graph is None:
            self._capture()
        if self._pipeline and not self._primed:
            self._defer_reduce = self._fuse_tail
            for p in range(self.parallel_updates):     # gradient for the first optimiser step
                self._use_scratch(p)
                self._train_grad_kernels(p)
            self._use_scratch(0)
            self._defer_reduce = False
            self._primed = True
        self._graph.replay()

    # ---- the reference-facing jitted callables ----------------------------------------------
    def initialize(self, key: int, subs=None, guess=None):
        """Initial simulator and planner state (planner.py:3159-3182, 2824-2851)."""
        with torch.cuda.device(self.device):
            gen = torch.Generator().manual_seed(int(key))
            self._gen.manual_seed(int(key))
            self._noise_seed = int(key) + 7919 * self.rank
            self._draws.zero_()
            self.init_train_host = self._init_words(self.train_fwd.program.state_layout, subs)
            self.init_test_host = self._init_words(self.test_fwd.program.state_layout, subs)
            self.load_initial_state(self.init_train_host, self.init_test_host)
            for p in range(self.parallel_updates):
                if guess is not None:
                    flat = self.plan.pytree_to_params(guess)
                else:
                    flat = self.plan.initial_params(gen)
                self.params[p].copy_(flat.reshape(-1))
            self.opt_state.zero_()
            if self._use_chain:         # the optax chain state is re-initialised too (P:2824-2851)
                self._chain_count.zero_()
                self._chain_ema.zero_()
                self._chain_plateau.copy_(plateau_init(self.parallel_updates, self.device))
            self._primed = False
            self.hyper[6] = 1.0
            engine.mean_loss(self.train_returns[0], self.batch_size_train, None, self.gret)
            if self.world_size > 1:          # d(-mean over the GLOBAL batch)/d return_b
                self.gret.mul_(1.0 / self.world_size)
            if self._branches:
                for _, _, gret in self._scratch[1:]:
                    gret.copy_(self.gret)
                self._gen.manual_seed(int(key) + 7919 * (self.rank + 1))
            self._draw(self.train_fwd.program, self.T, self.batch_size_train, self.train_noise, 0)
            self._draw(self.test_fwd.program, self.T, self.batch_size_test, self.test_noise, 1)
            if self._philox:
                engine.noise_advance(self._draws)

    def _plan_view(self, flat: torch.Tensor) -> torch.Tensor:
        """[..., n_params] in the shape the plan's pytree conversion expects."""
        if self.is_drp:
            return flat
        return flat.view(*flat.shape[:-1], self.T, self.A)

    def read_stats(self):
        """(train_loss [P], test_loss [P], zero_grads [P]) of the last iteration: one pinned
        device->host copy + the iteration's only host synchronisation."""
        P = self.parallel_updates
        self._stats_host.copy_(self._stats, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        raw = self._stats_host.numpy()
        train = raw[0:P].view(np.float32).copy()
        test = raw[P:2 * P].view(np.float32).copy()
        zero = raw[2 * P:3 * P] > 0
        if self.world_size > 1:
            both = self._pair_ctl.cpu().numpy()
            pair = both[:2 * P].reshape(2, P) / self.world_size
            train, test = pair[0], pair[1]
            self._agreed_ctl = (float(both[2 * P]), float(both[2 * P + 1]) / self.world_size)
        return train, test, zero

    def vote_loop_control(self, budget_reached: bool, progress_time: float):
        """Multi-GPU: this rank's view of the loop control (its own wall clock) for the exchange of the
        next iteration; ``agreed_loop_control`` returns what all ranks agreed on."""
        if self.world_size > 1:
            self._ctl_host[0] = 1.0 if budget_reached else 0.0
            self._ctl_host[1] = float(progress_time)
            self.ctl_in.copy_(self._ctl_host, non_blocking=True)

    def agreed_loop_control(self):
        """(any rank's time budget reached, mean progress over the ranks) of the last iteration."""
        votes, progress = getattr(self, "_agreed_ctl", (0.0, 0.0))
        return votes > 0.5, progress

    def redraw_noise(self, force: bool = False):
        """Fresh supplied-noise tensors for the next iteration (the reference splits its PRNG
        key once per update and once per test evaluation, planner.py:2880, 3257).  Once the
        iteration graph has been captured with the draws inside it (``_noise_in_graph``), every
        replay redraws by itself and this call is a no-op."""
        if not force and self._noise_in_graph and \
                (self._graph is not None or self._graph_host is not None):
            return
        with torch.cuda.device(self.device):
            self._draw(self.train_fwd.program, self.T, self.batch_size_train, self.train_noise, 0)
            self._draw(self.test_fwd.program, self.T, self.batch_size_test, self.test_noise, 1)
            if self.drp is not None:
                self.drp.redraw_noise(self._gen)
            if self._philox:
                engine.noise_advance(self._draws)

    def update(self):
        for p in range(self.parallel_updates):
            self._update_kernels(p)

    # ---- summaries printed by optimize() (planner.py:2563-2628) ----------------------------------
    def summarize_system(self) -> str:
        name = torch.cuda.get_device_name(self.device) if torch.cuda.is_available() else "n/a"
        return (f"\nStarting the B200 planner behind the JaxPlan API on device {self.device} ({name}), "
                f"rank {self.rank} of {self.world_size}\n")

    def summarize_relaxations(self) -> str:
        """Table of the non-differentiable operations and the classes that relax them."""
        try:
            info = getattr(self.compiled, "overriden_ops_info", lambda: {})()
        except Exception:           # a summary must never stop an optimisation
            info = {}
        if not info:
            return ""
        out = ("[INFO] Some RDDL operations are non-differentiable and will be approximated "
               "as follows:\n")
        try:
            exact = getattr(self.compiled, "exact_ops_info", lambda: {})()
        except Exception:
            exact = {}
        for klass, ops in info.items():
            out += f"    {klass}:\n"
            for op, ids in ops.items():
                out += f"        {op} [{len(ids)} relaxed, {len(exact.get(op, []))} exact]\n"
        return out

    def summarize_hyperparameters(self) -> str:
        utility = self.utility if isinstance(self.utility, str) else getattr(
            self.utility, "__name__", str(self.utility))
        out = (f"[INFO] objective hyper-parameters:\n"
               f"    utility_fn         ={utility}\n"
               f"    utility args       ={self.utility_kwargs}\n"
               f"    use_symlog         ={self.use_symlog_reward}\n"
               f"    use_critic         ={self.critic_fn is not None}\n"
               f"    lookahead          ={self.horizon}\n"
               f"    user_action_bounds ={self._action_bounds}\n"
               f"    start_entropy_coeff={self.start_entropy_coeff}\n"
               f"    end_entropy_coeff  ={self.end_entropy_coeff}\n"
               f"[INFO] optimizer hyper-parameters:\n"
               f"    optimizer         ={self.optimizer_name}\n"
               f"    optimizer args    ={self.optimizer_kwargs}\n"
               f"    clip_gradient     ={self.clip_grad}\n"
               f"    line_search_kwargs={self.line_search_kwargs}\n"
               f"    noise_kwargs      ={self.noise_kwargs}\n"
               f"    ema_decay         ={self.ema_decay}\n"
               f"    reduce_on_plateau ={self.reduce_on_plateau_kwargs}\n"
               f"    batch_size_train  ={self.batch_size_train}\n"
               f"    batch_size_test   ={self.batch_size_test}\n"
               f"    steps_per_update  ={self.steps_per_update}\n"
               f"    parallel_updates  ={self.parallel_updates}\n"
               f"    preprocessor      ={self.preprocessor}\n")
        out += str(self.plan)
        if self.use_pgpe:
            out += str(self.pgpe)
        out += "test compiler:\n"
        for k, v in self.test_compiled.get_kwargs().items():
            out += f"    {k}={v}\n"
        out += "train compiler:\n"
        for k, v in self.compiled.get_kwargs().items():
            out += f"    {k}={v}\n"
        return out

    def as_optimization_problem(self, key: Optional[int] = None, critic_params=None,
                                print_values: bool = False):
        """(loss_fn, grad_fn, guess_1d, unravel_fn) for off-the-shelf optimisers such as scipy
        (planner.py:2962-3031): loss and gradient of the relaxed train objective as functions of a
        1-D parameter vector, each call on fresh noise like the reference splits its key per call;
        unravel_fn maps the vector to the plan's parameter pytree.  Action constraints are the
        caller's, as in the reference."""
        if self.parallel_updates > 1:
            raise ValueError("Cannot compile static optimization problem when parallel_updates "
                             "is not None.")
        if self.world_size > 1:
            raise NotImplementedError("as_optimization_problem runs on one GPU")
        self.critic_params = critic_params
        if key is None:
            key = round(time.time() * 1000) % (2 ** 31)
        self.initialize(int(key))
        guess = self.params[0].detach().cpu().numpy().astype(np.float64)

        def _evaluate(params_1d):
            flat = torch.as_tensor(np.asarray(params_1d, dtype=np.float32).reshape(-1))
            self.params[0].copy_(flat.to(self.device))
            self.redraw_noise(force=True)
            self._train_grad_kernels(0)
            torch.cuda.synchronize(self.device)
            return float(self._train_loss_tmp[0]), \
                self.grad[0].detach().cpu().numpy().astype(np.float64)

        def _loss_function(params_1d):
            loss, _ = _evaluate(params_1d)
            if print_values:
                print(f"Loss: {loss:.6f}")
            return loss

        def _grad_function(params_1d):
            return _evaluate(params_1d)[1]

        def _unravel(params_1d):
            flat = torch.as_tensor(np.asarray(params_1d, dtype=np.float32).reshape(-1))
            return self.plan.params_to_pytree(self._plan_view(flat))
        return _loss_function, _grad_function, guess, _unravel

    def test_loss(self):
        for p in range(self.parallel_updates):
            self._test_kernels(p)

    # ---- the reference's closure seams as plain callables (planner.py:386-420, 2673-2698) ----------
    # The reference hands out jitted closures over (sim_state, planner_state) pytrees; here the same work
    # is done by kernels, and these thin callables take / return the parameter pytree directly.
    def projection(self, params: Dict[str, Any]):
        """plan.projection (planner.py:878-959): box clip and, when the instance limits concurrent actions,
        the max-nondef-actions projection, on the device.  Returns (projected pytree, converged)."""
        flat = self.plan.pytree_to_params(params).reshape(-1).to(self.device).contiguous()
        self.project_params(flat, 0)
        torch.cuda.synchronize(self.device)
        return (self.plan.params_to_pytree(self._plan_view(flat.cpu())), bool(self.converged[0].item()))

    def initializer(self, key: int):
        """plan.initializer (planner.py:974-1009): (initial parameter pytree after the projection,
        policy hyper-parameters)."""
        flat = self.plan.initial_params(torch.Generator().manual_seed(int(key))).reshape(-1)
        tree, _ = self.projection(self.plan.params_to_pytree(self._plan_view(flat)))
        hyper = {var: self.plan._sigmoid_weight for var in self.rddl.action_fluents} \
            if not self.is_drp else {"softmax_weight": self.plan._softmax_output_weight,
                                     "sigmoid_weight": self.plan._sigmoid_weight}
        return tree, hyper

    def test_policy(self, key, params: Dict[str, Any], policy_hyperparams, step: int, fls: Dict[str, Any]):
        """The test policy at one decision epoch (planner.py:842-868, 1482-1499): the same call as
        get_action."""
        return self.get_action(key, params, step, fls, policy_hyperparams)

    def _rollouts(self, params: Optional[Dict[str, Any]], train: bool) -> Dict[str, Any]:
        if self.is_drp:
            raise NotImplementedError("train_rollouts / test_rollouts of a closed-loop policy: use update() "
                                      "/ test_loss(), which walk the policy network between the steps")
        if params is not None:
            self.params[0].copy_(self.plan.pytree_to_params(params).reshape(-1).to(self.device))
        if train:
            T, B = self.T, self.batch_size_train
            self.train_fwd.forward(B, T, self.s0_train, policy=self.params[0], noise=self.train_noise,
                                   mparams=self.train_mparams, disc=self.disc,
                                   reward=self.train_reward[0], returns=self.train_returns[0],
                                   err=self.train_err[0], tape=self.tape)
            log = self._train_log()
            log.pop("grad", None)
        else:
            self._test_kernels(0)
            log = self._test_log()
        torch.cuda.synchronize(self.device)
        return log

    def train_rollouts(self, params: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        """Relaxed-model rollouts of the (given) plan on the current noise: the log of planner.py:2673-2680
        (``reward`` / ``error`` as [P, B, T])."""
        return self._rollouts(params, True)

    def test_rollouts(self, params: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        """Exact-model rollouts of the (given) plan: the log of planner.py:2681-2687 (with ``fluents``)."""
        return self._rollouts(params, False)

    # ---------------------------------------------------------------------------------
    def _train_log(self) -> Dict[str, Any]:
        P, T, B = self.parallel_updates, self.T, self.batch_size_train
        return {"reward": self.train_reward.view(P, T, B).transpose(1, 2),
                "error": self.train_err.view(P, T, B).transpose(1, 2),
                "grad": self.plan.params_to_pytree(self._plan_view(self.grad_used)),
                "fluents": {}}

    def _test_log(self) -> Dict[str, Any]:
        P, T, B = self.parallel_updates, self.T, self.batch_size_test
        te = self.test_fwd.program
        log = {"reward": self.test_reward.view(P, T, B).transpose(1, 2),
               "error": self.test_err.view(P, T, B).transpose(1, 2)}
        if self.test_traj is not None:
            shapes = {f: self.rddl.variable_shape(f) for f in te.traj_layout}
            fl = unpack_fluent_major(te.traj_layout, self.test_traj.view(P, T, te.TRAJ * B), B,
                                     shapes, lead=(P, T))
            log["fluents"] = {k: v.transpose(1, 2) for k, v in fl.items()}
        else:
            shapes = {f: self.rddl.variable_shape(f) for f in te.state_layout}
            fl = unpack_fluent_major(te.state_layout, self.test_final, B, shapes, lead=(P,))
            log["fluents"] = {k: v.unsqueeze(2) for k, v in fl.items()}
        return log

    def optimize(self, *args, **kwargs) -> Optional[Dict[str, Any]]:
        """Run the generator to completion and return the last callback (planner.py:3033)."""
        it = self.optimize_generator(*args, **kwargs)
        callback = None
        for callback in it:
            pass
        return callback

    def optimize_generator(self, key: Optional[int] = None, epochs: int = 999999,
                           train_seconds: float = 120., model_params: Optional[Dict] = None,
                           policy_hyperparams: Optional[Dict] = None, subs: Optional[Dict] = None,
                           guess: Optional[Dict] = None, critic_params=None,
                           print_summary: bool = True, print_progress: bool = True,
                           print_hyperparams: bool = False, dashboard_id=None,
                           stopping_rule: Optional[JaxPlannerStoppingRule] = None,
                           test_rolling_window: int = 10, tqdm_position: Optional[int] = None
                           ) -> Iterable[Dict[str, Any]]:
        """The optimisation loop of planner.py:3073-3469 (pgpe=None branch)."""
        start_time = time.time()
        elapsed_outside_loop = 0.
        if key is None:
            key = round(time.time() * 1000) % (2 ** 31)
        if hasattr(key, "__len__"):
            key = int(np.asarray(key).ravel()[-1])
        if model_params is not None:
            self.set_model_params(model_params)
        if print_summary:           # planner.py:3132-3134
            print(self.summarize_system())
            print(self.summarize_relaxations())
        if print_hyperparams:       # planner.py:3135-3147
            print(self.summarize_hyperparameters())
            print(f"[INFO] optimize call hyper-parameters:\n"
                  f"    PRNG key           ={key}\n"
                  f"    max_iterations     ={epochs}\n"
                  f"    max_seconds        ={train_seconds}\n"
                  f"    model_params       ={model_params}\n"
                  f"    policy_hyper_params={policy_hyperparams}\n"
                  f"    override_subs_dict ={subs is not None}\n"
                  f"    provide_param_guess={guess is not None}\n"
                  f"    test_rolling_window={test_rolling_window}\n")
        self.initialize(key, subs, guess)
        self.critic_params = critic_params
        P = self.parallel_updates
        if policy_hyperparams is not None:
            hyper = policy_hyperparams
        elif self.is_drp:           # planner.py:1519-1521
            hyper = {"softmax_weight": self.plan._softmax_output_weight,
                     "sigmoid_weight": self.plan._sigmoid_weight}
        else:                       # planner.py:976 (+ the pooled softmax weight, planner.py:1003)
            hyper = {var: self.plan._sigmoid_weight for var in self.rddl.action_fluents}
            if getattr(self.plan, "stack_bool", False):
                hyper["__bool__"] = self.plan._softmax_weight
        if policy_hyperparams is None:      # preprocessor statistics travel with them (P:1519-1521)
            hyper.update(getattr(self.plan, "hyperparams_extra", {}))
        best_params = self.plan.params_to_pytree(self._plan_view(self.params[0]).clone().cpu())
        best_loss, best_grad, best_index = np.inf, None, 0
        last_iter_improve = 0
        rolling = RollingMean(test_rolling_window)
        rolling_pgpe = RollingMean(test_rolling_window)
        pbest_loss = np.full(P, np.inf)
        pgpe_return = None
        if self.use_pgpe:
            self.pgpe.initialize(self.params, self._gen)
        status = JaxPlannerStatus.NORMAL
        if stopping_rule is not None:
            stopping_rule.reset()
        shown_train, shown_test = set(), set()
        it = -1
        progress_percent = 0.
        for it in range(epochs):
            status = JaxPlannerStatus.NORMAL
            with torch.cuda.device(self.device):
                if self.world_size > 1:     # this rank's clock, agreed on through the exchange
                    local = time.time() - start_time - elapsed_outside_loop
                    self.vote_loop_control(local >= train_seconds, local / train_seconds)
                if self.train_noise is not None or self.test_noise is not None \
                        or (self.drp is not None and self.drp.pol_noise is not None):
                    self.redraw_noise()
                if self.drp is not None:      # planner.py:2775-2778, 2812-2813
                    self.drp.ent_coeff.fill_(self.start_entropy_coeff * (
                        self.end_entropy_coeff / self.start_entropy_coeff) ** (progress_percent / 100.))
                elif getattr(self.plan, "_stochastic", False):      # planner.py:2775-2778
                    self.ent_coeff.fill_(self.start_entropy_coeff * (
                        self.end_entropy_coeff / self.start_entropy_coeff) ** (progress_percent / 100.))
                self.iterate()
                train_loss, test_loss, zero_grads = self.read_stats()       # the host sync
            test_loss_smooth = rolling.update(test_loss)
            pgpe_improve = False
            if self.use_pgpe:                           # planner.py:3270-3296
                with torch.cuda.device(self.device):
                    self.pgpe.update(progress_percent)
                    for p in range(P):
                        self._test_kernels(p, params=self.pgpe.mu[p], scratch=True,
                                           loss_out=self._pg_loss[p:p + 1])
                    pgpe_loss = self._pg_loss.cpu().numpy()
                pgpe_loss_smooth = rolling_pgpe.update(pgpe_loss)
                pgpe_return = -pgpe_loss_smooth
                mask = np.less(pgpe_loss_smooth, pbest_loss) | ~np.isfinite(train_loss)
                if np.any(mask):                        # hard replacement (planner.py:2932-2949)
                    for p in np.nonzero(mask)[0]:
                        self.params[p].copy_(self.pgpe.mu[p])
                    test_loss = np.where(mask, pgpe_loss, test_loss)
                    test_loss_smooth = np.where(mask, pgpe_loss_smooth, test_loss_smooth)
                    self._primed = False                # the pending gradient is for the old plan
                    pgpe_improve = True
            pbest_loss = np.minimum(pbest_loss, test_loss_smooth)
            best_index = int(np.argmin(test_loss_smooth))
            if test_loss_smooth[best_index] < best_loss:
                best_params = self.plan.params_to_pytree(
                    self._plan_view(self.params[best_index]).clone().cpu())
                best_grad = self.plan.params_to_pytree(
                    self._plan_view(self.grad_used[best_index]).clone().cpu())
                best_loss = float(test_loss_smooth[best_index])
                last_iter_improve = it
            if (not pgpe_improve) and np.all(zero_grads):
                status = JaxPlannerStatus.NO_PROGRESS
            if np.all(~np.isfinite(train_loss)):
                status = JaxPlannerStatus.INVALID_GRADIENT
            if print_progress and it % 50 == 0:
                for name, buf, shown in (("Training", self.train_err, shown_train),
                                         ("Testing", self.test_err, shown_test)):
                    for code in torch.unique(buf).cpu().numpy():
                        if int(code) not in shown:
                            shown.add(int(code))
                            for msg in JaxRDDLCompiler.get_error_messages(int(code)):
                                print(f"[FAIL] {name} model error: {msg}")
            elapsed = time.time() - start_time - elapsed_outside_loop
            progress_time = elapsed / train_seconds
            if self.world_size > 1:     # every rank acts on the SAME budget flag and progress
                budget_reached, progress_time = self.agreed_loop_control()
            else:
                budget_reached = elapsed >= train_seconds
            if budget_reached:
                status = JaxPlannerStatus.TIME_BUDGET_REACHED
            if it >= epochs - 1:
                status = JaxPlannerStatus.ITER_BUDGET_REACHED
            progress_it = it / (epochs - 1) if epochs > 1 else 0
            progress_percent = 100 * min(1, max(0, progress_time, progress_it))
            callback = {
                "iteration": it, "elapsed_time": elapsed, "progress": progress_percent,
                "status": status, "key": key, "test_key": key,
                "train_return": -train_loss, "test_return": -test_loss_smooth,
                "best_return": -best_loss, "pgpe_return": pgpe_return,
                "last_iteration_improved": last_iter_improve, "pgpe_improved": pgpe_improve,
                "params": self.plan.params_to_pytree(self._plan_view(self.params)),
                "best_params": best_params, "best_index": best_index,
                "pgpe_params": (self.plan.params_to_pytree(self._plan_view(self.pgpe.mu))
                                if self.use_pgpe else None),
                "model_params": self.model_parameter_info(), "policy_hyperparams": hyper,
                "grad": self.plan.params_to_pytree(self._plan_view(self.grad_used)),
                "best_grad": best_grad, "train_log": self._train_log(),
                "test_log": self._test_log(), "planner_state": self,
            }
            if stopping_rule is not None and stopping_rule.monitor(callback):
                callback["status"] = status = JaxPlannerStatus.STOPPING_RULE_REACHED
            if print_progress and (it % 100 == 0 or status.is_terminal()):
                print(f"{it:6d} it | {-np.min(train_loss):13.5f} train | "
                      f"{-np.min(test_loss_smooth):13.5f} test | {-best_loss:13.5f} best | "
                      f"{status.value} status | {(it + 1) / (elapsed + 1e-6):.2f}it/s")
            t_out = time.time()
            yield callback
            elapsed_outside_loop += time.time() - t_out
            if status.is_terminal():
                break
        if print_summary:
            print(f"[INFO] Summary of optimization:\n    status:         {status}\n"
                  f"    time:           {time.time() - start_time - elapsed_outside_loop:.2f} seconds\n"
                  f"    iterations:     {it}\n    best objective: {-best_loss:.5f}\n")

    def get_action(self, key, params: Dict[str, Any], step: int, state: Dict[str, Any],
                   policy_hyperparams: Optional[Dict] = None, history=None,
                   copy_state: bool = True) -> Dict[str, Any]:
        """Action of the test policy at one decision epoch (planner.py:3527-3586)."""
        flat = self.plan.pytree_to_params(params)
        if self.is_drp:
            def words(name):        # a carried observation is looked up under the observ-fluent's name;
                key = name[:-len("@seen")] if name.endswith("@seen") else name     # hidden words read 0
                if key in state:
                    return np.asarray(state[key], np.float32).reshape(-1)
                return np.zeros(self.rddl.variable_size(name), np.float32)
            vec = np.concatenate([words(name) for name in self.rddl.state_fluents])
            return self.plan.test_actions_host(flat, vec, step=int(step))
        step = min(int(step), flat.shape[0] - 1)
        return self.plan.test_actions(flat[step].numpy())


# =====================================================================================
# config files (planner.py:101-244)
# =====================================================================================

def _config_sections(parser: "configparser.RawConfigParser"):
    """(parser, {section: {key: python literal}}) of a case-preserving .cfg parser."""
    return parser, {sec: {key: literal_eval(text) for key, text in parser.items(sec)}
                    for sec in parser.sections()}


def _new_cfg_parser() -> "configparser.RawConfigParser":
    parser = configparser.RawConfigParser()
    parser.optionxform = str                  # keys are keyword-argument names: keep their case
    return parser


def _parse_config_file(path: str):
    parser = _new_cfg_parser()
    if not parser.read(path):
        raise FileNotFoundError(f"No config file found at {path}.")
    return _config_sections(parser)


def _parse_config_string(value: str):
    parser = _new_cfg_parser()
    parser.read_string(value)
    return _config_sections(parser)


def _getattr_any(packages, item):
    """First non-None attribute ``item`` among ``packages`` (None when nobody has it)."""
    found = (getattr(pkg, item, None) for pkg in packages)
    return next((val for val in found if val is not None), None)


def _load_config(config, args):
    compiler_kwargs = {k: args["Compiler"][k] for k in args.get("Compiler", {})}
    planner_args = {k: args["Planner"][k] for k in args.get("Planner", {})}
    train_args = {k: args["Optimize"][k] for k in args.get("Optimize", {})}
    if "method" in compiler_kwargs:                             # planner.py:130-137
        name = compiler_kwargs.pop("method")
        cls = getattr(logic, name, None)
        if cls is None:
            raise ValueError(f"Invalid compiler class <{name}>.")
        planner_args["compiler"] = cls
    planner_args["compiler_kwargs"] = compiler_kwargs
    plan_method = planner_args.pop("method")
    plan_kwargs = planner_args.pop("method_kwargs", {})
    if "initializer" in plan_kwargs:                            # planner.py:140-155
        init = getattr(initializers, plan_kwargs["initializer"], None)
        kw = plan_kwargs.pop("initializer_kwargs", {})
        plan_kwargs["initializer"] = init(**kw) if callable(init) and kw else (
            init() if callable(init) and not isinstance(init, _Initializer) else init)
    import sys
    this = sys.modules[__name__]
    from . import drp as drp_mod
    act = plan_kwargs.get("activation")                         # planner.py:157-164
    if isinstance(act, str) and act not in engine.ACTIVATIONS:
        raise ValueError(f"Invalid activation <{act}>.")
    emb = plan_kwargs.get("time_embedding")                     # planner.py:166-169
    if isinstance(emb, str):
        if emb not in drp_mod.TIME_EMBEDDINGS:
            raise ValueError(f"Invalid time embedding <{emb}>.")
        plan_kwargs["time_embedding"] = drp_mod.TIME_EMBEDDINGS[emb]
    plan_cls = getattr(this, plan_method, None)
    if plan_cls is None:
        from . import drp
        plan_cls = getattr(drp, plan_method, None)
    if plan_cls is None:
        raise ValueError(f"Invalid plan class <{plan_method}>.")
    planner_args["plan"] = plan_cls(**plan_kwargs)
    opt = planner_args.get("optimizer")
    if isinstance(opt, str) and opt not in OPTIMIZERS:          # planner.py:180-186
        raise ValueError(f"Invalid optimizer <{opt}>.")
    pgpe_method = planner_args.get("pgpe", None)                # planner.py:188-198
    pgpe_kwargs = planner_args.pop("pgpe_kwargs", {})
    if pgpe_method is not None:
        from . import pgpe as pgpe_mod
        cls = getattr(pgpe_mod, pgpe_method, None)
        if cls is None:
            raise ValueError(f"Invalid PGPE class <{pgpe_method}>.")
        planner_args["pgpe"] = cls(**pgpe_kwargs)
    preproc = planner_args.get("preprocessor", None)            # planner.py:200-205
    preproc_kwargs = planner_args.pop("preprocessor_kwargs", {})
    if isinstance(preproc, str):
        cls = getattr(this, preproc, None)
        if cls is None or not (isinstance(cls, type) and issubclass(cls, Preprocessor)):
            raise ValueError(f"Invalid preprocessor class <{preproc}>.")
        planner_args["preprocessor"] = cls(**preproc_kwargs)
    if "dashboard" in planner_args:                             # planner.py:212-217: no dashboard here
        del planner_args["dashboard"]
    if "stopping_rule" in train_args:                           # planner.py:219-226
        rule = getattr(this, train_args["stopping_rule"])
        train_args["stopping_rule"] = rule(**train_args.pop("stopping_rule_kwargs", {}))
    return planner_args, plan_kwargs, train_args


def load_config(path: str):
    """(planner_args, plan_kwargs, train_args) from a .cfg file (planner.py:229-235)."""
    return _load_config(*_parse_config_file(path))


def load_config_from_string(value: str):
    return _load_config(*_parse_config_string(value))


# =====================================================================================
# controllers (planner.py:3632-3831)
# =====================================================================================

class JaxOfflineController:
    """Optimise once, then replay the plan (planner.py:3632-3726)."""
    use_tensor_obs = True

    def __init__(self, planner: JaxBackpropPlanner, key: Optional[int] = None,
                 eval_hyperparams: Optional[Dict] = None, params: Optional[Any] = None,
                 train_on_reset: bool = False, save_path: Optional[str] = None,
                 **train_kwargs) -> None:
        self.planner = planner
        self.key = key if key is not None else round(time.time() * 1000) % (2 ** 31)
        self.eval_hyperparams = eval_hyperparams
        self.train_on_reset = train_on_reset
        self.train_kwargs = train_kwargs
        self.params_given = params is not None
        self.hyperparams_given = eval_hyperparams is not None
        if isinstance(params, str):
            with open(params, "rb") as f:
                data = pickle.load(f)
            params, self.eval_hyperparams = data["params"], data["hyperparams"]
            self.hyperparams_given = True
        self.step = 0
        self.callback = None
        if not self.train_on_reset and not self.params_given:
            self.callback = self.planner.optimize(key=self.key, **self.train_kwargs)
            params = self.callback["best_params"]
            if not self.hyperparams_given:
                self.eval_hyperparams = self.callback["policy_hyperparams"]
            if save_path is not None:
                with open(save_path, "wb") as f:
                    pickle.dump({"params": params, "hyperparams": self.eval_hyperparams}, f)
        self.params = params

    def sample_action(self, state: Dict[str, Any]) -> Dict[str, Any]:
        actions = self.planner.get_action(self.key, self.params, self.step, state,
                                          self.eval_hyperparams)
        self.step += 1
        return actions

    def reset(self) -> None:
        self.step = 0
        if self.train_on_reset and not self.params_given:
            self.callback = self.planner.optimize(key=self.key, **self.train_kwargs)
            self.params = self.callback["best_params"]
            if not self.hyperparams_given:
                self.eval_hyperparams = self.callback["policy_hyperparams"]


class JaxOnlineController:
    """Re-optimise at every decision epoch with a warm start (planner.py:3729-3831)."""
    use_tensor_obs = True

    def __init__(self, planner: JaxBackpropPlanner, key: Optional[int] = None,
                 eval_hyperparams: Optional[Dict] = None, warm_start: bool = True,
                 max_attempts: int = 3, **train_kwargs) -> None:
        self.planner = planner
        self.key = key if key is not None else round(time.time() * 1000) % (2 ** 31)
        self.eval_hyperparams = eval_hyperparams
        self.warm_start = warm_start
        self.max_attempts = max_attempts
        self.train_kwargs = train_kwargs
        self.reset()

    def sample_action(self, state: Dict[str, Any]) -> Dict[str, Any]:
        planner = self.planner
        callback = planner.optimize(key=self.key, guess=self.guess, subs=state,
                                    **self.train_kwargs)
        self.key += 1
        params = callback["best_params"]
        hyper = self.eval_hyperparams if self.eval_hyperparams is not None \
            else callback["policy_hyperparams"]
        actions = planner.get_action(self.key, params, 0, state, hyper)
        if self.warm_start:
            self.guess = planner.plan.guess_next_epoch(params)
        self.callback = callback
        return actions

    def reset(self) -> None:
        self.guess = None
        self.callback = None


from .drp import DrpPipeline, JaxDeepReactivePolicy  # noqa: E402,F401
from .pgpe import GaussianPGPE, PGPE  # noqa: E402,F401
