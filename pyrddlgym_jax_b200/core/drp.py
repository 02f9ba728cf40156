"""Deep reactive policy (closed-loop MLP plan) on the B200 kernels.

``JaxDeepReactivePolicy`` mirrors planner.py:1122-1534 of the reference (constructor keywords,
parameter pytree with flax layer names ``dense_{i}`` / ``dense_{var}_mu`` /
``dense_{var}_log_sigma``); ``DrpPipeline`` is the device side of one planner iteration for
such a plan.  Because the policy reads the current state, the horizon is walked one decision
epoch per launch:

    forward  t = 0..T-1 :  dense layers (rk_sgemm, bias+tanh fused) -> heads -> actions
                           (rk_drp_actions_forward) -> one T=1 launch of the rollout kernel
    backward t = T-1..0 :  T=1 adjoint launch (state cotangent chained through GSIN/GS0, action
                           cotangent out) -> heads/dense adjoints (rk_sgemm with the tanh
                           adjoint fused; weight gradients accumulated with a deterministic
                           split-K) -> input cotangent added to the state cotangent

Hidden activations of every step are kept in HBM (T*B*H floats per layer) rather than
recomputed: at B=32768, T=100, H=256 that is 3.4 GB per layer, far below the 180 GB available.
All of it is captured into the iteration's CUDA graph by the planner.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import engine
from .planner_base import JaxPlan, initializers
from .program import PlanSpec, action_layout


HOST_ACTIVATIONS = {
    "tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid,
    "softplus": torch.nn.functional.softplus, "elu": torch.nn.functional.elu,
    "leaky_relu": lambda x: torch.nn.functional.leaky_relu(x, 0.01),
    "gelu": lambda x: torch.nn.functional.gelu(x, approximate="tanh"),      # jax.nn.gelu's default form
    "silu": torch.nn.functional.silu, "swish": torch.nn.functional.silu}


class SinusoidalTimeEmbedding:
    """Sinusoidal positional encoding of the decision epoch (planner.py:1108-1119)."""
    kind = "sin"

    def __init__(self, dim: int = 8, name: str = "sin_time_embed") -> None:
        self.dim, self.name = int(dim), name

    def table(self, T: int) -> torch.Tensor:
        """Rows 0..T-1: (sin, cos) pairs of t * 10000^(-2 i / dim), interleaved, first dim entries."""
        freqs = 10000.0 ** (-2.0 * torch.arange(self.dim, dtype=torch.float32) / self.dim)
        angles = torch.arange(T, dtype=torch.float32).reshape(T, 1) * freqs
        return torch.stack([torch.sin(angles), torch.cos(angles)], dim=-1).reshape(T, -1)[:, :self.dim] \
            .contiguous()


class FixedTimeEmbedding:
    """Learnable table of time embeddings (planner.py:1095-1105)."""
    kind = "fixed"

    def __init__(self, max_t: int = 500, dim: int = 8, init: Any = None,
                 name: str = "fixed_time_embed") -> None:
        if init is not None:
            raise NotImplementedError("FixedTimeEmbedding(init=...): only the default normal(0.01)")
        self.max_t, self.dim, self.name = int(max_t), int(dim), name


TIME_EMBEDDINGS = {"SinusoidalTimeEmbedding": SinusoidalTimeEmbedding,
                   "FixedTimeEmbedding": FixedTimeEmbedding}


class JaxDeepReactivePolicy(JaxPlan):
    """A deep reactive policy network (planner.py:1122)."""

    kind = "drp"
    history_dependent = False

    def __init__(self, topology: Optional[Sequence[int]] = None, activation: Any = "tanh",
                 initializer: Any = None, normalize: bool = False,
                 normalize_per_layer: bool = False, normalizer_kwargs: Optional[Dict] = None,
                 sigmoid_weight: float = 1.0, wrap_non_bool: bool = False,
                 softmax_weight: float = 1.0, stochastic: bool = True,
                 sigma_range: Tuple[float, float] = (1e-5, 1e3), sigma_entropy_grad: bool = False,
                 action_noise_dist: Any = "normal", time_dependent: bool = False,
                 time_embedding: Any = None, time_embedding_kwargs: Optional[Dict] = None) -> None:
        super().__init__()
        self._topology = list(topology) if topology is not None else [64, 64]
        # the reference takes a jax.nn function (planner.py:1128, 144-150 resolve the .cfg name);
        # here the name selects the epilogue of the layer kernels
        act = getattr(activation, "__name__", activation)
        if act not in engine.ACTIVATIONS:
            raise NotImplementedError(f"activation <{act}>: one of {sorted(engine.ACTIVATIONS)} "
                                      "(planner.py:1128)")
        self._activation = act
        self._wrap_non_bool = bool(wrap_non_bool)
        self._time_dependent = bool(time_dependent)
        if time_embedding is None:
            time_embedding = SinusoidalTimeEmbedding
        if isinstance(time_embedding, str):
            time_embedding = TIME_EMBEDDINGS[time_embedding]
        self._time_embedding = time_embedding
        self._time_embedding_kwargs = dict(time_embedding_kwargs or {})
        self.time_embed = time_embedding(**self._time_embedding_kwargs) if self._time_dependent \
            else None
        self._sigma_entropy_grad = bool(sigma_entropy_grad)
        # the reference takes a jax.random sampler (planner.py:1139, 1418-1420); here its name
        dist = getattr(action_noise_dist, "__name__", action_noise_dist)
        if dist not in ("normal", "laplace", "logistic", "cauchy", "gumbel"):
            raise NotImplementedError(f"action_noise_dist <{dist}> is not built")
        self._action_noise_dist = dist
        self._initializer_base = initializer
        self._sigmoid_weight = sigmoid_weight
        self._softmax_output_weight = softmax_weight
        self._stochastic = stochastic
        self._sigma_range = sigma_range
        self._normalize = bool(normalize)
        self._normalize_per_layer = bool(normalize_per_layer)
        if normalizer_kwargs is None:       # planner.py:1179-1181
            normalizer_kwargs = {"use_bias": True, "use_scale": True}
        extra = set(normalizer_kwargs) - {"use_bias", "use_scale", "epsilon"}
        if extra or not normalizer_kwargs.get("use_bias", True) \
                or not normalizer_kwargs.get("use_scale", True):
            raise NotImplementedError(
                f"normalizer_kwargs {normalizer_kwargs}: only epsilon (with scale and bias) is built")
        self._normalizer_kwargs = dict(normalizer_kwargs)
        self._norm_eps = float(normalizer_kwargs.get("epsilon", 1e-6))     # flax nn.LayerNorm

    def __str__(self) -> str:
        return (f"[INFO] policy hyper-parameters:\n"
                f"    topology          ={self._topology}\n"
                f"    activation_fn     ={self._activation}\n"
                f"    apply_input_norm  ={self._normalize}\n"
                f"    input_norm_layerwise={self._normalize_per_layer}\n"
                f"    input_norm_args   ={self._normalizer_kwargs}\n"
                f"    stochastic        ={self._stochastic}\n"
                f"    sigma_range       ={self._sigma_range}\n")

    # ---------------------------------------------------------------------------------
    def compile(self, compiled, test_compiled, _bounds, horizon, preprocessor=None) -> None:
        rddl = compiled.rddl
        self.rddl = rddl
        self.horizon = horizon
        self.layout, self.A = action_layout(rddl)
        # head kind per action fluent (planner.py:1368-1409): 0 real, 1 bool (sigmoid of a logit
        # head), 2 int / enum (real head, rounded by the test policy)
        # with a binding max-nondef-actions the boolean fluents share the 'dense_topk' head
        # (planner.py:1322-1326, 1435-1447): kind 3, or 4 when the fluent's default is true
        self.action_kinds = [1 if prange == "bool" else (0 if prange == "real" else 2)
                             for (_, _, prange) in self.layout.values()]
        n_bool = sum(n for _, n, r in self.layout.values() if r == "bool")
        self.use_topk = bool(n_bool and rddl.max_allowed_actions < n_bool)
        self.allowed = int(rddl.max_allowed_actions)
        if self.use_topk:
            if n_bool > engine.TOPK_MAX_POOLED or self.allowed > engine.TOPK_MAX_ALLOWED:
                raise NotImplementedError(
                    f"dense_topk head: {n_bool} pooled boolean actions / max-nondef-actions "
                    f"{self.allowed} exceed the kernel limits ({engine.TOPK_MAX_POOLED} / "
                    f"{engine.TOPK_MAX_ALLOWED})")
            self.action_kinds = [
                (4 if bool(np.ravel(np.asarray(rddl.action_fluents[v]))[0]) else 3)
                if prange == "bool" else k
                for k, (v, (_, _, prange)) in zip(self.action_kinds, self.layout.items())]
        # partially observed models (planner.py:1300-1313: the policy reads the observ-fluents only): the
        # planner hands over rddl.model.observation_view, where every observ-fluent has a carried copy in
        # the state vector; the copies are the network's inputs, the hidden state words keep structurally
        # zero first-layer rows (no initial weight, no gradient, absent from the parameter pytree)
        if rddl.observ_fluents and not getattr(rddl, "observed_inputs", None):
            raise ValueError("a partially observed model needs rddl.model.observation_view(model) "
                             "(JaxBackpropPlanner applies it)")
        inputs = list(getattr(rddl, "observed_inputs", None) or rddl.state_fluents)
        bounds = {}
        for name, (_, _, prange) in self.layout.items():
            shape = rddl.variable_shape(name)
            if prange == "bool":
                bounds[name] = (np.full(shape, -np.inf, np.float32), np.full(shape, np.inf, np.float32))
                continue
            if prange in rddl.enum_types:
                lo, hi = 0.0, float(len(rddl.type_to_objects[prange]) - 1)
            else:
                lo, hi = rddl.bounds[name]
            lo, hi = _bounds.get(name, (lo, hi))
            bounds[name] = (np.broadcast_to(np.asarray(lo, np.float32), shape).copy(),
                            np.broadcast_to(np.asarray(hi, np.float32), shape).copy())
        self.bounds = bounds
        self.spec = PlanSpec(kind="external", bounds=bounds)
        # input vector: the state words in the kernels' fluent-major order; `input_perm` maps
        # the reference's row order (sorted names, non-bool first, planner.py:1302-1313) to it
        off, pos = 0, {}
        for name in rddl.state_fluents:
            n = rddl.variable_size(name)
            pos[name] = (off, n)
            off += n
        self.D = off
        self.state_segs = [(pos[v][0], pos[v][1]) for v in rddl.state_fluents]
        # exact-model state words of bool / int fluents are int32: converted before the first layer
        self.state_is_int = [rddl.variable_ranges[v] != "real" for v in rddl.state_fluents]
        self.action_segs = [(o, n) for (o, n, _) in self.layout.values()]
        names = sorted(inputs)
        ref = [v for v in names if rddl.variable_ranges[v] != "bool"] + \
              [v for v in names if rddl.variable_ranges[v] == "bool"]
        self.input_perm = np.concatenate([np.arange(pos[v][0], pos[v][0] + pos[v][1]) for v in ref])
        self.D_in = int(len(self.input_perm))
        # word ranges of the state vector the network must not look at (hidden state of a POMDP)
        self.frozen_segs = [pos[v] for v in rddl.state_fluents if v not in inputs]
        # optional state preprocessing before everything else (planner.py:1293-1296): per state word
        # (x - lo) / (hi - lo); words without stats keep lo = 0, hi = 1
        self.pre_lo = self.pre_hi = None
        if preprocessor is not None:
            preprocessor.compile(compiled)
            stats = preprocessor.initialize()
            lo_w, hi_w = np.zeros(self.D, np.float32), np.ones(self.D, np.float32)
            for v, (lower, upper) in stats.items():
                o, n = pos[v] if v in pos else pos[v + "@seen"]
                shape = rddl.variable_shape(v)
                lo_w[o:o + n] = np.broadcast_to(lower, shape).reshape(-1)
                hi_w[o:o + n] = np.broadcast_to(upper, shape).reshape(-1)
            if stats:
                self.pre_lo, self.pre_hi = lo_w, hi_w
            self.hyperparams_extra = {preprocessor.HYPERPARAMS_KEY: stats}
        # input LayerNorm (planner.py:1259-1282, 1300-1319): `norm_layers` = [(flax name, fluents)]
        # in the reference's order, `norm_groups[f]` = group of state fluent f for the kernels
        # (0 = not normalised), `norm_perm[name]` = kernel positions of that module's scale / bias
        non_bool = [v for v in names if rddl.variable_ranges[v] != "bool"]
        per_layer = self._normalize_per_layer
        self.norm_layers: List[Tuple[str, List[str]]] = []
        if self._normalize:
            if per_layer and any(pos[v][1] == 1 for v in non_bool):
                if getattr(compiled, "print_warnings", False):
                    print("[WARN] Cannot apply layer norm to a state-fluent of size 1: "
                          "setting normalize_per_layer = False.")
                per_layer = False
            if per_layer:
                self.norm_layers = [("norm_" + v.replace("-", "_"), [v]) for v in non_bool]
            elif sum(pos[v][1] for v in non_bool) > 1:
                self.norm_layers = [("norm", non_bool)]
            elif non_bool and getattr(compiled, "print_warnings", False):
                print("[WARN] Cannot apply layer norm to state-fluents of total size 1: "
                      "setting normalize = False.")
        group_of = {v: g + 1 for g, (_, members) in enumerate(self.norm_layers) for v in members}
        self.norm_groups = [group_of.get(v, 0) for v in rddl.state_fluents]
        kpos, j = {}, 0
        for v in rddl.state_fluents:
            if v in group_of:
                kpos[v] = np.arange(j, j + pos[v][1])
                j += pos[v][1]
        self.Dn = j
        self.norm_perm = {name: np.concatenate([kpos[v] for v in members])
                          for name, members in self.norm_layers}
        # flat parameter vector: [W_i | b_i]* then [W_heads | b_heads] (then [ln_scale | ln_bias])
        self.NH = self.A * (2 if self._stochastic else 1)
        dims = [self.D] + self._topology
        self.segments: List[Tuple[str, int, int, int]] = []     # (name, offset, rows, cols)
        o = 0
        for i in range(len(self._topology)):
            self.segments.append((f"W{i}", o, dims[i], dims[i + 1]))
            o += dims[i] * dims[i + 1]
            self.segments.append((f"b{i}", o, 1, dims[i + 1]))
            o += dims[i + 1]
        self.segments.append(("Wh", o, dims[-1], self.NH))
        o += dims[-1] * self.NH
        self.segments.append(("bh", o, 1, self.NH))
        o += self.NH
        if self.Dn:
            self.segments.append(("ln_scale", o, 1, self.Dn))
            self.segments.append(("ln_bias", o + self.Dn, 1, self.Dn))
            o += 2 * self.Dn
        # FiLM time conditioning (planner.py:1083-1092, 1353-1363): per hidden layer the kernels and
        # biases of the gamma / beta Dense layers over the time embedding; then the learned table
        self.film = self.time_embed is not None
        if self.film:
            dim = self.time_embed.dim
            if self.time_embed.kind == "fixed" and horizon > self.time_embed.max_t + 1:
                raise ValueError(f"FixedTimeEmbedding(max_t={self.time_embed.max_t}) is shorter "
                                 f"than the horizon {horizon}")
            for i, h in enumerate(self._topology):
                for nm, r in ((f"fgW{i}", dim), (f"fgb{i}", 1), (f"fbW{i}", dim), (f"fbb{i}", 1)):
                    self.segments.append((nm, o, r, h))
                    o += r * h
            if self.time_embed.kind == "fixed":
                self.segments.append(("temb", o, self.time_embed.max_t + 1, dim))
                o += (self.time_embed.max_t + 1) * dim
        self.n_params = o
        self.seg = {name: (off, r, c) for name, off, r, c in self.segments}
        self.box = (None, None)
        self.needs_concurrency_projection = False

    def num_params(self, horizon: int) -> int:
        return self.n_params

    def initial_params(self, generator: torch.Generator) -> torch.Tensor:
        """variance_scaling(2.0, fan_in, truncated_normal) kernels, zero biases
        (planner.py:1129-1130, 1520-1529)."""
        flat = torch.zeros(self.n_params)
        for name, off, r, c in self.segments:
            if name.startswith("W"):
                rows = self.D_in if name == "W0" else r        # fan-in = the inputs the policy sees
                if self._initializer_base is not None:
                    w = self._initializer_base((rows, c), generator)
                else:
                    std = math.sqrt(2.0 / rows) / 0.87962566103423978
                    w = torch.empty(rows, c)
                    torch.nn.init.trunc_normal_(w, mean=0., std=1., a=-2., b=2., generator=generator)
                    w = w * std
                if rows != r:                                   # hidden state words: zero rows
                    full = torch.zeros(r, c)
                    full[torch.as_tensor(np.sort(self.input_perm))] = w
                    w = full
                flat[off:off + r * c] = w.reshape(-1)
            elif name == "ln_scale":
                flat[off:off + c] = 1.0
            elif name.startswith("fgW") or name.startswith("fbW"):     # flax Dense: lecun_normal
                w = torch.empty(r, c)
                torch.nn.init.trunc_normal_(w, mean=0., std=1., a=-2., b=2., generator=generator)
                flat[off:off + r * c] = (w * (math.sqrt(1.0 / r) / 0.87962566103423978)).reshape(-1)
            elif name == "temb":                                        # nn.initializers.normal()
                flat[off:off + r * c] = torch.randn(r * c, generator=generator) * 0.01
        return flat

    # ---- flat vector <-> flax-style pytree ------------------------------------------------
    def params_to_pytree(self, flat: torch.Tensor) -> Dict[str, Any]:
        lead = flat.shape[:-1]
        tree: Dict[str, Any] = {}
        for name, idx in self.norm_perm.items():       # DRPStateLayer 'inputs' / LayerNorm modules
            so, sb = self.seg["ln_scale"][0], self.seg["ln_bias"][0]
            idx = torch.as_tensor(idx)
            tree.setdefault("inputs", {})[name] = {"scale": flat[..., so + idx],
                                                   "bias": flat[..., sb + idx]}
        for i in range(len(self._topology)):
            off, r, c = self.seg[f"W{i}"]
            W = flat[..., off:off + r * c].reshape(*lead, r, c)
            if i == 0:
                W = W[..., torch.as_tensor(self.input_perm), :]
            ob, _, cb = self.seg[f"b{i}"]
            tree[f"dense_{i}"] = {"kernel": W, "bias": flat[..., ob:ob + cb]}
        off, r, c = self.seg["Wh"]
        Wh = flat[..., off:off + r * c].reshape(*lead, r, c)
        ob = self.seg["bh"][0]
        bh = flat[..., ob:ob + c]
        if self.film:
            blk = lambda nm: (lambda o, r, c: flat[..., o:o + r * c].reshape(*lead, r, c) if r > 1  # noqa
                              else flat[..., o:o + c])(*self.seg[nm])
            for i in range(len(self._topology)):
                tree[f"film_{i}"] = {"gamma": {"kernel": blk(f"fgW{i}"), "bias": blk(f"fgb{i}")},
                                     "beta": {"kernel": blk(f"fbW{i}"), "bias": blk(f"fbb{i}")}}
            if self.time_embed.kind == "fixed":
                tree[self.time_embed.name] = {"kernel": blk("temb")}
        if self.use_topk:               # planner.py:1436-1438: the pooled boolean logits
            cols = torch.as_tensor(self._pooled_columns())
            tree["dense_topk"] = {"kernel": Wh[..., cols], "bias": bh[..., cols]}
        for var, (aoff, n, prange) in self.layout.items():
            nm = var.replace("-", "_")
            if prange == "bool" and self.use_topk:
                continue
            if prange == "bool":        # planner.py:1375-1377: a single logit head
                tree[f"dense_{nm}_logit"] = {"kernel": Wh[..., aoff:aoff + n],
                                             "bias": bh[..., aoff:aoff + n]}
                continue
            tree[f"dense_{nm}_mu"] = {"kernel": Wh[..., aoff:aoff + n], "bias": bh[..., aoff:aoff + n]}
            if self._stochastic:
                tree[f"dense_{nm}_log_sigma"] = {
                    "kernel": Wh[..., self.A + aoff:self.A + aoff + n],
                    "bias": bh[..., self.A + aoff:self.A + aoff + n]}
        return {"params": tree}

    def pytree_to_params(self, tree: Dict[str, Any]) -> torch.Tensor:
        tree = tree.get("params", tree)
        flat = torch.zeros(self.n_params)
        as_t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32)   # noqa: E731
        for name, idx in self.norm_perm.items():
            so, sb = self.seg["ln_scale"][0], self.seg["ln_bias"][0]
            idx = torch.as_tensor(idx)
            flat[so + idx] = as_t(tree["inputs"][name]["scale"]).reshape(-1)
            flat[sb + idx] = as_t(tree["inputs"][name]["bias"]).reshape(-1)
        for i in range(len(self._topology)):
            off, r, c = self.seg[f"W{i}"]
            W = as_t(tree[f"dense_{i}"]["kernel"]).reshape(-1, c)
            if i == 0:
                Wi = torch.zeros(r, c)
                Wi[torch.as_tensor(self.input_perm)] = W
                W = Wi
            flat[off:off + r * c] = W.reshape(-1)
            ob, _, cb = self.seg[f"b{i}"]
            flat[ob:ob + cb] = as_t(tree[f"dense_{i}"]["bias"]).reshape(-1)
        if self.film:
            def put(nm, x):
                o, r, c = self.seg[nm]
                flat[o:o + r * c] = as_t(x).reshape(-1)
            for i in range(len(self._topology)):
                put(f"fgW{i}", tree[f"film_{i}"]["gamma"]["kernel"])
                put(f"fgb{i}", tree[f"film_{i}"]["gamma"]["bias"])
                put(f"fbW{i}", tree[f"film_{i}"]["beta"]["kernel"])
                put(f"fbb{i}", tree[f"film_{i}"]["beta"]["bias"])
            if self.time_embed.kind == "fixed":
                put("temb", tree[self.time_embed.name]["kernel"])
        off, r, c = self.seg["Wh"]
        Wh = torch.zeros(r, c)
        bh = torch.zeros(c)
        if self.use_topk:
            cols = torch.as_tensor(self._pooled_columns())
            Wh[:, cols] = as_t(tree["dense_topk"]["kernel"]).reshape(r, len(cols))
            bh[cols] = as_t(tree["dense_topk"]["bias"]).reshape(-1)
        for var, (aoff, n, prange) in self.layout.items():
            nm = var.replace("-", "_")
            if prange == "bool" and self.use_topk:
                continue
            if prange == "bool":
                Wh[:, aoff:aoff + n] = as_t(tree[f"dense_{nm}_logit"]["kernel"]).reshape(r, n)
                bh[aoff:aoff + n] = as_t(tree[f"dense_{nm}_logit"]["bias"]).reshape(-1)
                continue
            Wh[:, aoff:aoff + n] = as_t(tree[f"dense_{nm}_mu"]["kernel"]).reshape(r, n)
            bh[aoff:aoff + n] = as_t(tree[f"dense_{nm}_mu"]["bias"]).reshape(-1)
            if self._stochastic:
                Wh[:, self.A + aoff:self.A + aoff + n] = \
                    as_t(tree[f"dense_{nm}_log_sigma"]["kernel"]).reshape(r, n)
                bh[self.A + aoff:self.A + aoff + n] = \
                    as_t(tree[f"dense_{nm}_log_sigma"]["bias"]).reshape(-1)
        flat[off:off + r * c] = Wh.reshape(-1)
        ob = self.seg["bh"][0]
        flat[ob:ob + c] = bh
        return flat

    def _pooled_columns(self) -> np.ndarray:
        """Head columns of the pooled boolean actions in the order of the reference's softmax
        output (layer_sizes order, planner.py:1453-1465)."""
        return np.concatenate([np.arange(aoff, aoff + n)
                               for (aoff, n, prange) in self.layout.values() if prange == "bool"])

    def guess_next_epoch(self, params):
        return params                      # planner.py:1531-1534: warm start from the same policy

    def time_embedding_rows(self, flat: torch.Tensor, T: int) -> torch.Tensor:
        """[T, dim] embeddings of the decision epochs 0..T-1 (a view of the parameters when learned)."""
        if self.time_embed.kind == "fixed":
            o, r, c = self.seg["temb"]
            return flat[o:o + r * c].view(r, c)[:T]
        return self.time_embed.table(T).to(flat.device)

    def test_actions_host(self, flat: torch.Tensor, state_vec: np.ndarray,
                          step: int = 0) -> Dict[str, np.ndarray]:
        """Test policy on one state (controllers' get_action, planner.py:3527): a handful of
        tiny mat-vecs on the host."""
        h = torch.as_tensor(state_vec, dtype=torch.float32).reshape(1, -1).clone()
        if self.pre_lo is not None:
            lo_w, hi_w = torch.from_numpy(self.pre_lo), torch.from_numpy(self.pre_hi)
            h = (h - lo_w) / (hi_w - lo_w)
        if self.Dn:
            so, sb = self.seg["ln_scale"][0], self.seg["ln_bias"][0]
            words = {g: [] for g in range(1, len(self.norm_layers) + 1)}
            j = 0
            for (off, n), g in zip(self.state_segs, self.norm_groups):
                if g:
                    words[g].append((off, n, j))
                    j += n
            for g, blocks in words.items():
                cols = torch.cat([torch.arange(o, o + n) for o, n, _ in blocks])
                par = torch.cat([torch.arange(jj, jj + n) for _, n, jj in blocks])
                x = h[0, cols]
                mean = x.mean()
                var = torch.clamp((x * x).mean() - mean * mean, min=0.0)
                h[0, cols] = (x - mean) * torch.rsqrt(var + self._norm_eps) * flat[so + par] \
                    + flat[sb + par]
        for i in range(len(self._topology)):
            off, r, c = self.seg[f"W{i}"]
            ob = self.seg[f"b{i}"][0]
            h = HOST_ACTIVATIONS[self._activation](
                h @ flat[off:off + r * c].reshape(r, c) + flat[ob:ob + c])
            if self.film:
                emb = self.time_embedding_rows(flat, int(step) + 1)[int(step)].reshape(1, -1)
                vec = lambda nm: (lambda o, rr, cc: flat[o:o + rr * cc].reshape(rr, cc))(*self.seg[nm])  # noqa
                gamma = emb @ vec(f"fgW{i}") + vec(f"fgb{i}")
                beta = emb @ vec(f"fbW{i}") + vec(f"fbb{i}")
                h = gamma * h + beta
        off, r, c = self.seg["Wh"]
        ob = self.seg["bh"][0]
        out = (h @ flat[off:off + r * c].reshape(r, c) + flat[ob:ob + c]).reshape(-1)
        acts = {}
        if self.use_topk:               # train = 0 (planner.py:1039-1045, 1066-1078), then > 0.5
            cols = torch.as_tensor(self._pooled_columns())
            lg = out[cols]
            if self.allowed == 1:
                pooled = torch.softmax(lg, 0)
            else:
                pooled = torch.zeros_like(lg)
                lg = lg.clone()
                for _ in range(self.allowed):
                    i = int(torch.argmax(lg))
                    lg[i] = -math.inf
                    pooled[i] = 1.0
            out = out.clone()
            out[cols] = pooled
        for var, (aoff, n, prange) in self.layout.items():
            lo, hi = self.bounds[var]
            a = out[aoff:aoff + n].numpy().reshape(self.rddl.variable_shape(var))
            if prange == "bool" and self.use_topk:
                if bool(np.ravel(np.asarray(self.rddl.action_fluents[var]))[0]):
                    a = 1.0 - a
                acts[var] = a > 0.5
                continue
            if prange == "bool":        # sigmoid(w * logit) > 0.5  (planner.py:1489)
                acts[var] = (self._sigmoid_weight * a) > 0.0
                continue
            if self._wrap_non_bool:     # planner.py:1411-1415
                fl, fh = np.isfinite(lo), np.isfinite(hi)
                lo0, hi0 = np.where(fl, lo, 0.0), np.where(fh, hi, 0.0)
                sp = lambda x: np.logaddexp(0.0, x)      # noqa: E731
                a = np.where(fl & fh, lo0 + (hi0 - lo0) / (1.0 + np.exp(-a)),
                             np.where(fl, lo0 + sp(a), np.where(fh, hi0 - sp(-a), a))).astype(np.float32)
            a = np.minimum(np.maximum(a, lo), hi)
            acts[var] = a if prange == "real" else np.round(a).astype(np.int32)
        return acts


class DrpPipeline:
    """Device buffers + launch sequence of one planner iteration with a DRP (see module doc)."""

    def __init__(self, planner):
        self.pl = planner
        plan: JaxDeepReactivePolicy = planner.plan
        self.plan = plan
        dev = planner.device
        self.dev = dev
        T, Bt, Be = planner.T, planner.batch_size_train, planner.batch_size_test
        fw, te = planner.train_fwd.program, planner.test_fwd.program
        self.S, self.A, self.NH, self.D = fw.S, plan.A, plan.NH, plan.D
        assert self.D == self.S
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)   # noqa: E731
        self.hid = plan._topology
        self.tape = z((T + 1) * self.S * Bt)
        self.H = [z(T, Bt, h) for h in self.hid]
        self.heads = z(T, Bt, self.NH)
        self.act = z(T, self.A * Bt)
        self.entropy = z(T, Bt)
        self.ent_b = z(Bt)                             # per-rollout entropy sums (loss reduction scratch)
        self._u_loss, self._m_loss = z(1), z(1)        # -utility(returns), -mean(returns) of an update
        self.pol_noise = z(T, self.A * Bt) if plan._stochastic else None
        self.gact = z(self.A * Bt)
        self.gs = [z(self.S * Bt), z(self.S * Bt)]
        self.gheads = z(Bt, self.NH)
        self.gz = [z(Bt, h) for h in self.hid]
        self.ones = torch.ones(max(Bt, Be), dtype=torch.float32, device=dev)
        # weight-gradient GEMMs reduce over the batch: deterministic split-K, slice count chosen
        # by the library (negative = "up to"), workspace sized for the largest product
        max_split = int(max(1, min(128, Bt // 256)))
        self.max_split = max_split
        self.split_k = -max_split if max_split > 1 else 1
        big = max([(self.D + 1) * self.hid[0]] + [a * b for a, b in zip(self.hid[:-1], self.hid[1:])]
                  + [self.hid[-1] * self.NH])
        self.ws = z(max_split * big) if max_split > 1 else None
        self.wT = z(plan.n_params)                    # transposed kernels, same offsets
        self.xp = z(T, Bt * (self.D + 1))             # [x | 1] per step: fused dW0 / db0 operand
        self.cs_chunks = int(max(1, min(64, Bt // 64)))
        self.cs_ws = z(self.cs_chunks * max(max(self.hid), self.NH))
        # fused streaming backward kernels (one pass over [B][h] instead of four / three)
        import os as _os0
        fuse = _os0.environ.get("RK_DRP_FUSED", "1") != "0"
        # gelu / silu: the derivative needs the pre-activation, so every layer is z = X W + b (kept) then
        # y = act(z) (kept): the layer-wise path with a second tape; the fused kernels keep act(z) only
        self.zact = plan._activation in engine.Z_ACTIVATIONS
        if self.zact:
            if plan.time_embed is not None:
                raise NotImplementedError(f"activation <{plan._activation}> with time_dependent=True")
            fuse = False
            self.Z = [z(T, Bt, h) for h in self.hid]
        self._Z = None                      # pre-activation buffers of the epoch being evaluated
        hl = self.hid[-1]
        self.fused_head = fuse and len(self.hid) >= 2 and self.NH <= 16 and hl <= 1024
        self.fused_l0 = fuse and self.D <= 16
        # whole forward of a 2-hidden-layer network in one launch per decision epoch (rk_drp_mlp2_forward)
        self.fused_fwd = (fuse and self.fused_l0 and len(self.hid) == 2 and self.D <= 16 and self.NH <= 16
                          and Bt >= 128 and self.hid[0] % 16 == 0 and self.hid[0] <= 256
                          and self.hid[1] % 32 == 0 and 32 <= self.hid[1] <= 256
                          and _os0.environ.get("RK_TCGEMM", "1") != "0"
                          and _os0.environ.get("RK_DRP_FUSED_FWD", "1") != "0")
        self.fused_bwd = self.fused_fwd and self.fused_head and self.hid[0] % 32 == 0 \
            and self.D <= 15 and _os0.environ.get("RK_DRP_FUSED_BWD", "1") != "0"
        if self.fused_fwd:
            # operand images of the fused kernels (W0 / Wh; rebuilt by refresh_splits)
            shp = (self.D, self.hid[0], self.hid[1], self.NH)
            self.img_f = z(engine.drp_mlp2_image_floats(*shp, False))
            self.img_b = z(engine.drp_mlp2_image_floats(*shp, True))
        if self.fused_bwd:
            # pre-activation cotangents and head cotangents of EVERY epoch stay in HBM: the weight
            # gradients are one reduction over all T * B rows after the reverse sweep (rk_wgrad.cu)
            self.gz_tape = [z(T, Bt, h) for h in self.hid]
            self.gheads_tape = z(T, Bt, self.NH)
            self.wg_ws = z(engine.drp_wgrad_ctas() * (self.hid[0] * self.hid[1] + self.hid[1]))
        ctas = engine.drp_head_backward_ctas(Bt)
        self.fb_ws = z(ctas * max(hl + hl * self.NH + self.NH, (self.D + 1) * self.hid[0]))
        # tensor-core path (tcgen05, 3xTF32) for hidden layers whose shapes fit the MMA tiles
        import os as _os
        ok = lambda a, b: a % 32 == 0 and b % 32 == 0 and 32 <= a <= 256 and 32 <= b <= 256  # noqa
        self.tc_layers = [i for i in range(1, len(self.hid))
                          if _os.environ.get("RK_TCGEMM", "1") != "0" and Bt >= 128
                          and ok(self.hid[i - 1], self.hid[i])]
        if self.tc_layers:
            self.w_hi, self.w_lo = z(plan.n_params), z(plan.n_params)       # split of W  [in][out]
            self.wT_hi, self.wT_lo = z(plan.n_params), z(plan.n_params)     # split of W^T [out][in]
        lo = np.concatenate([plan.bounds[v][0].reshape(-1) for v in plan.layout])
        hi = np.concatenate([plan.bounds[v][1].reshape(-1) for v in plan.layout])
        self.lo = torch.from_numpy(lo.astype(np.float32)).to(dev)
        self.hi = torch.from_numpy(hi.astype(np.float32)).to(dev)
        self.ls_lo, self.ls_hi = math.log(plan._sigma_range[0]), math.log(plan._sigma_range[1])
        # exact test model: a single slot of activations is enough (forward only)
        self.t_state = [z(te.S * Be), z(te.S * Be)]
        self.t_H = [z(Be, h) for h in self.hid]
        self.t_Z = [z(Be, h) for h in self.hid] if self.zact else None
        self.t_heads = z(Be, self.NH)
        self.t_act = z(self.A * Be)
        self.t_xf = z(te.S * Be) if any(plan.state_is_int) else None
        # entropy-regulariser coefficient: starts at start_entropy_coeff (planner.py:2775-2778, progress 0);
        # optimize_generator anneals it every iteration
        self.ent_coeff = torch.full((1,), float(planner.start_entropy_coeff), device=dev)
        # input LayerNorm: normalised inputs of every step (operands of the first layer's weight
        # gradient), cotangent of the normalised input, per-CTA partials of [dscale | dbias]
        self.Dn = plan.Dn
        self.pre = plan.pre_lo is not None
        if self.Dn or self.pre:
            self.gy = z(self.S * Bt)
        if self.Dn:
            self.xn = z(T, self.S * Bt)
            self.ln_ws = z(engine.state_layernorm_ctas(Bt) * 2 * self.Dn)
            self.t_xn = z(te.S * Be)
        if self.pre:        # scaled inputs of every step (LayerNorm / first-layer operands)
            self.pre_lo = torch.from_numpy(plan.pre_lo).to(dev)
            self.pre_hi = torch.from_numpy(plan.pre_hi).to(dev)
            self.xa = z(T, self.S * Bt)
            self.gya = z(self.S * Bt)
            self.t_xa = z(te.S * Be)

        # FiLM: conditioned activations of every step (inputs of the next layer), the per-step gamma /
        # beta vectors of this update and their cotangents [dgamma | dbeta]
        self.film = plan.film
        if self.film:
            self.fused_head = False       # the head adjoint's (1 - H^2) factor moves behind FiLM
            self.Hf = [z(T, Bt, h) for h in self.hid]
            self.t_Hf = [z(Be, h) for h in self.hid]
            self.gam = [z(T, h) for h in self.hid]
            self.bet = [z(T, h) for h in self.hid]
            self.dgb = [z(T, 2 * h) for h in self.hid]
            self.film_ws = z(engine.film_chunks(max(Bt, Be)) * 2 * max(self.hid))
            self.sin_table = None if plan.time_embed.kind == "fixed" \
                else plan.time_embed.table(T).to(dev)

    def film_tables(self, params):
        """gamma_t = Dense(t_emb), beta_t = Dense(t_emb) for every decision epoch of the horizon
        (planner.py:1088-1091): [T, dim] x [dim, h] products, once per update."""
        plan = self.plan
        emb = self.sin_table if self.sin_table is not None \
            else plan.time_embedding_rows(params, self.pl.T)
        for i, h in enumerate(self.hid):
            dim = emb.shape[1]
            torch.addmm(self._W(params, f"fgb{i}"), emb, self._W(params, f"fgW{i}").view(dim, h),
                        out=self.gam[i])
            torch.addmm(self._W(params, f"fbb{i}"), emb, self._W(params, f"fbW{i}").view(dim, h),
                        out=self.bet[i])
        return emb

    def film_param_grads(self, params, grad, emb):
        """Chain the per-step [dgamma | dbeta] sums into the FiLM Dense layers and, when it is
        learned, the rows of the time-embedding table."""
        plan = self.plan
        T, dim = emb.shape
        for i, h in enumerate(self.hid):
            dg, db = self.dgb[i][:, :h], self.dgb[i][:, h:]
            self._W(grad, f"fgW{i}").view(dim, h).addmm_(emb.t(), dg)
            self._W(grad, f"fgb{i}").add_(dg.sum(0))
            self._W(grad, f"fbW{i}").view(dim, h).addmm_(emb.t(), db)
            self._W(grad, f"fbb{i}").add_(db.sum(0))
            if plan.time_embed.kind == "fixed":
                rows = self._W(grad, "temb").view(-1, dim)[:T]
                rows.addmm_(dg, self._W(params, f"fgW{i}").view(dim, h).t())
                rows.addmm_(db, self._W(params, f"fbW{i}").view(dim, h).t())

    def layernorm_forward(self, params, B, x, y):
        plan = self.plan
        engine.state_layernorm_forward(B, self.D, plan.state_segs, plan.norm_groups, plan._norm_eps,
                                       x, self._W(params, "ln_scale"), self._W(params, "ln_bias"), y)

    # ---- views into a flat parameter / gradient vector ---------------------------------------
    def _W(self, flat, name):
        off, r, c = self.plan.seg[name]
        return flat[off:off + r * c]

    def refresh_transposes(self, params: torch.Tensor):
        """W^T of every Dense kernel and, for the tensor-core layers, the TF32 hi / lo parts of W and W^T
        (once per update, native kernels: rk_transpose_split / rk_tf32_split)."""
        tc = {f"W{i}" for i in self.tc_layers}
        for name, off, r, c in self.plan.segments:
            if name.startswith("W"):
                n = r * c
                split = name in tc
                engine.transpose_split(r, c, params[off:off + n], self.wT[off:off + n],
                                       self.wT_hi[off:off + n] if split else None,
                                       self.wT_lo[off:off + n] if split else None)
                if split:
                    engine.tf32_split(params[off:off + n], self.w_hi[off:off + n], self.w_lo[off:off + n], n)
        self.refresh_splits(params)

    def refresh_splits(self, params: torch.Tensor):
        """Operand images of the fused kernels (once per update)."""
        if self.fused_fwd and self.tc_layers:
            engine.drp_mlp2_prepare(self.D, self.hid[0], self.hid[1], self.NH, self._W(params, "W0"),
                                    self._W(params, "Wh"), self.img_f, self.img_b)

    # ---- the MLP ------------------------------------------------------------------------------
    def mlp_forward(self, params, B, x_fm, H, heads, xp=None, film=None, keep_hidden=True):
        """x_fm: the state in the kernels' fluent-major layout (fluent f = block [B][n_f] at word
        offset off_f * B); H: per-layer outputs [B][h]; heads [B][NH]; xp: optional [B][D+1]
        packed copy [x | 1] for the backward pass; keep_hidden=False (exact test sweep): the
        hidden activations are not needed afterwards."""
        h0 = self.hid[0]
        if self.fused_fwd and film is None and 1 in self.tc_layers:
            engine.drp_mlp2_forward(
                self.plan._activation, B, self.D, h0, self.hid[1], self.NH, self.plan.state_segs, x_fm,
                self.img_f, self._W(params, "b0"), self._W(self.wT_hi, "W1"),
                self._W(self.wT_lo, "W1"), self._W(params, "b1"), self._W(params, "bh"),
                H[0] if keep_hidden else None, H[1] if keep_hidden else None, heads)
            return
        if self.zact:
            self._mlp_forward_z(params, B, x_fm, H, heads, xp)
            return
        engine.drp_layer0(B, self.D, h0, self.plan.state_segs, x_fm, self._W(params, "W0"),
                          self._W(params, "b0"), H[0], xp)
        fan, src = h0, H[0]
        if film is not None:            # film = (decision epoch, conditioned-activation buffers)
            t, Hf = film
            engine.film_forward(B, h0, H[0], self.gam[0][t], self.bet[0][t], Hf[0])
            src = Hf[0]
        for i in range(1, len(self.hid)):
            h = self.hid[i]
            if i in self.tc_layers:     # Y = tanh(X W + b) as X * (W^T)^T on the tensor cores
                engine.tcgemm(2, B, h, fan, src, fan, self._W(self.wT_hi, f"W{i}"),
                              self._W(self.wT_lo, f"W{i}"), fan, H[i], h,
                              bias=self._W(params, f"b{i}"))
            else:
                engine.sgemm(engine.EPI_BIAS_TANH, B, h, fan, src, fan, self._W(params, f"W{i}"),
                             h, H[i], h, bias=self._W(params, f"b{i}"))
            fan, src = h, H[i]
            if film is not None:
                engine.film_forward(B, h, H[i], self.gam[i][t], self.bet[i][t], Hf[i])
                src = Hf[i]
        if self.NH <= 16:
            engine.rowdot(False, B, self.NH, fan, src, fan, self._W(params, "Wh"), self.NH, heads,
                          self.NH, bias=self._W(params, "bh"))
        else:
            engine.sgemm(engine.EPI_BIAS, B, self.NH, fan, src, fan, self._W(params, "Wh"),
                         self.NH, heads, self.NH, bias=self._W(params, "bh"))

    def _mlp_forward_z(self, params, B, x_fm, H, heads, xp):
        """Layers whose activation keeps its pre-activation (gelu, silu): z_i = X W_i + b_i into self._Z[i]
        (bias-only epilogues), y_i = act(z_i) into H[i] (rk_act_apply)."""
        act, Z = self.plan._activation, self._Z
        engine.drp_layer0(B, self.D, self.hid[0], self.plan.state_segs, x_fm, self._W(params, "W0"),
                          self._W(params, "b0"), Z[0], xp, act=engine.ACT_IDENTITY)
        engine.act_apply(act, B * self.hid[0], Z[0], H[0])
        fan = self.hid[0]
        for i in range(1, len(self.hid)):
            h = self.hid[i]
            if i in self.tc_layers:
                engine.tcgemm(1, B, h, fan, H[i - 1], fan, self._W(self.wT_hi, f"W{i}"),
                              self._W(self.wT_lo, f"W{i}"), fan, Z[i], h, bias=self._W(params, f"b{i}"))
            else:
                engine.sgemm(engine.EPI_BIAS, B, h, fan, H[i - 1], fan, self._W(params, f"W{i}"),
                             h, Z[i], h, bias=self._W(params, f"b{i}"))
            engine.act_apply(act, B * h, Z[i], H[i])
            fan = h
        if self.NH <= 16:
            engine.rowdot(False, B, self.NH, fan, H[-1], fan, self._W(params, "Wh"), self.NH, heads,
                          self.NH, bias=self._W(params, "bh"))
        else:
            engine.sgemm(engine.EPI_BIAS, B, self.NH, fan, H[-1], fan, self._W(params, "Wh"),
                         self.NH, heads, self.NH, bias=self._W(params, "bh"))

    def wgrad_all(self, grad, x_tape):
        """Weight gradients of the fused path: one reduction over all T * B (epoch, rollout) rows of the
        activation / cotangent tapes (``x_tape``: the first layer's inputs of every epoch, fluent-major)."""
        pl = self.pl
        T, B = pl.T, pl.batch_size_train
        h0, h1 = self.hid
        K = T * B
        engine.drp_wgrad_hidden(True, h0, h1, K, self.H[0], self.gz_tape[1], self._W(grad, "W1"),
                                self._W(grad, "b1"), self.wg_ws)
        engine.drp_wgrad_heads(True, h1, self.NH, K, self.H[1], self.gheads_tape, self._W(grad, "Wh"),
                               self._W(grad, "bh"), self.wg_ws)
        engine.drp_wgrad_input(True, h0, self.D, T, B, self.S, self.plan.state_segs, self.gz_tape[0],
                               x_tape, self._W(grad, "W0"), self._W(grad, "b0"), self.wg_ws)

    def mlp_backward(self, grad, B, x_fm, H, gs_cur, xp, params_w=None, film=None, t=0):
        """Consumes self.gheads [B][NH]; accumulates weight gradients into ``grad`` and adds the
        input cotangent to the fluent-major state cotangent ``gs_cur``."""
        AT, ACC = engine.GEMM_A_TRANS, engine.GEMM_ACCUMULATE
        sk, ws = self.split_k, self.ws
        last = len(self.hid) - 1
        hl = self.hid[last]
        if self.fused_bwd and film is None and 1 in self.tc_layers:
            # one launch for the whole input-side adjoint of the epoch: gz1, gz0 (kept for the weight
            # gradients, see wgrad_all) and the state cotangent
            engine.drp_mlp2_backward(
                self.plan._activation, B, self.D, self.hid[0], hl, self.NH, self.plan.state_segs,
                self.gheads_tape[t], H[0], H[1], self.img_b, self._W(self.w_hi, "W1"),
                self._W(self.w_lo, "W1"), self.gz_tape[0][t], self.gz_tape[1][t], gs_cur)
            return
        if film is not None:    # layer inputs are the conditioned activations; z = H stays the tanh
            t, Hf = film
            engine.sgemm(AT | ACC, hl, self.NH, B, Hf[last], hl, self.gheads, self.NH,
                         self._W(grad, "Wh"), self.NH, workspace=ws, split_k=sk)
            engine.colsum(B, self.NH, self.gheads, self.NH, self._W(grad, "bh"), True, self.cs_ws,
                          self.cs_chunks)
            engine.sgemm(engine.EPI_NONE, B, hl, self.NH, self.gheads, self.NH,
                         self._W(self.wT, "Wh"), hl, self.gz[last], hl)
            engine.film_backward(B, hl, self.gz[last], H[last], self.gam[last][t], self.gz[last],
                                 self.dgb[last][t], self.film_ws)
        elif self.fused_head:     # gz, dWh, dbh and db_last in one pass over H[last] / gheads
            engine.drp_head_backward(B, hl, self.NH, H[last], self.gheads, self._W(params_w, "Wh"),
                                     self.gz[last], grad[self.plan.seg[f"b{last}"][0]:], self.fb_ws)
        else:
            engine.sgemm(AT | ACC, hl, self.NH, B, H[last], hl, self.gheads, self.NH,
                         self._W(grad, "Wh"), self.NH, workspace=ws, split_k=sk)
            engine.colsum(B, self.NH, self.gheads, self.NH, self._W(grad, "bh"), True, self.cs_ws,
                          self.cs_chunks)
            engine.sgemm(engine.EPI_DTANH, B, hl, self.NH, self.gheads, self.NH,
                         self._W(self.wT, "Wh"), hl, self.gz[last], hl,
                         aux=(self._Z if self.zact else H)[last], ldaux=hl)
        AUX = self._Z if self.zact else H     # what the activation adjoint reads: z (gelu, silu) or act(z)
        for i in range(last, -1, -1):
            h = self.hid[i]
            fan = self.D if i == 0 else self.hid[i - 1]
            if i == 0 and self.fused_l0:   # dW0, db0 and the state cotangent in one pass over gz[0]
                engine.drp_layer0_backward(B, self.D, h, self.plan.state_segs, x_fm, self.gz[0],
                                           self._W(self.wT, "W0"), gs_cur, self._W(grad, "W0"),
                                           self.fb_ws)
                continue
            if i == 0:      # [x | 1]^T dZ: the rows of dW0 and, in row D, db0 (contiguous in grad)
                engine.sgemm(AT | ACC, fan + 1, h, B, xp, fan + 1, self.gz[0], h,
                             self._W(grad, "W0"), h, workspace=ws, split_k=sk)
            else:
                src = H[i - 1] if film is None else Hf[i - 1]
                if i in self.tc_layers and self.ws is not None:     # dW = H^T dZ on the tensor cores
                    engine.tcgemm_tn(True, fan, h, B, src, fan, self.gz[i], h,
                                     self._W(grad, f"W{i}"), self.ws, self.max_split)
                else:
                    engine.sgemm(AT | ACC, fan, h, B, src, fan, self.gz[i], h,
                                 self._W(grad, f"W{i}"), h, workspace=ws, split_k=sk)
                if not (i == last and self.fused_head):     # db_last came with the head backward
                    engine.colsum(B, h, self.gz[i], h, self._W(grad, f"b{i}"), True, self.cs_ws,
                                  self.cs_chunks)
            if i > 0 and film is not None:       # dX = dZ W^T, then back through FiLM and tanh
                if i in self.tc_layers:
                    engine.tcgemm(0, B, fan, h, self.gz[i], h, self._W(self.w_hi, f"W{i}"),
                                  self._W(self.w_lo, f"W{i}"), h, self.gz[i - 1], fan)
                else:
                    engine.sgemm(engine.EPI_NONE, B, fan, h, self.gz[i], h,
                                 self._W(self.wT, f"W{i}"), fan, self.gz[i - 1], fan)
                engine.film_backward(B, fan, self.gz[i - 1], H[i - 1], self.gam[i - 1][t],
                                     self.gz[i - 1], self.dgb[i - 1][t], self.film_ws)
            elif i > 0 and i in self.tc_layers:    # dX = (dZ W^T) * (1 - H^2): "Bt" is W itself
                engine.tcgemm(3, B, fan, h, self.gz[i], h, self._W(self.w_hi, f"W{i}"),
                              self._W(self.w_lo, f"W{i}"), h, self.gz[i - 1], fan,
                              aux=AUX[i - 1], ldaux=fan)
            elif i > 0:
                engine.sgemm(engine.EPI_DTANH, B, fan, h, self.gz[i], h, self._W(self.wT, f"W{i}"),
                             fan, self.gz[i - 1], fan, aux=AUX[i - 1], ldaux=fan)
            else:           # input cotangent, added block by block to the state cotangent
                W0T = self._W(self.wT, "W0")
                for off, n in self.plan.state_segs:
                    if n <= 16:
                        engine.rowdot(True, B, n, h, self.gz[0], h, W0T[off:], fan,
                                      gs_cur[off * B:], n)
                    else:
                        engine.sgemm(ACC, B, n, h, self.gz[0], h, W0T[off:], fan,
                                     gs_cur[off * B:], n)

    # ---- one update (planner.py:2868-2917) -----------------------------------------------------
    def update_kernels(self, p: int, loss_out: Optional[torch.Tensor] = None):
        pl, plan = self.pl, self.plan
        T, B, S, A = pl.T, pl.batch_size_train, self.S, self.A
        params, grad = pl.params[p], pl.grad[p]
        fwd, bwd = pl.train_fwd, pl.train_bwd
        N = fwd.program.N
        engine.drp_set_activation(plan._activation)
        self.refresh_transposes(params)          # W^T and the hi / lo splits of this update's W
        emb = self.film_tables(params) if self.film else None
        self.tape[:S * B].copy_(pl.s0_train)
        for t in range(T):
            x = self.tape[t * S * B:(t + 1) * S * B]
            Ht = [Hl[t] for Hl in self.H]
            if self.zact:
                self._Z = [Zl[t] for Zl in self.Z]
            xin = x
            if self.pre:
                engine.state_scale(B, self.D, plan.state_segs, self.pre_lo, self.pre_hi, x, self.xa[t])
                xin = self.xa[t]
            if self.Dn:
                self.layernorm_forward(params, B, xin, self.xn[t])
                xin = self.xn[t]
            self.mlp_forward(params, B, xin, Ht, self.heads[t], self.xp[t],
                             film=(t, [Hl[t] for Hl in self.Hf]) if self.film else None)
            engine.drp_actions_forward(
                B, A, plan._stochastic, plan.action_segs, self.heads[t],
                None if self.pol_noise is None else self.pol_noise[t], 1.0, self.lo, self.hi,
                self.ls_lo, self.ls_hi, self.act[t], self.entropy[t], kinds=plan.action_kinds,
                bool_weight=plan._sigmoid_weight, wrap=plan._wrap_non_bool)
            if plan.use_topk:
                engine.drp_topk_forward(
                    B, A, plan._stochastic, plan.action_segs, plan.action_kinds, self.heads[t],
                    None if self.pol_noise is None else self.pol_noise[t], True, plan.allowed,
                    plan._softmax_output_weight, self.act[t], self.entropy[t])
            fwd.forward(B, 1, x, actions=self.act[t],
                        noise=None if pl.train_noise is None else pl.train_noise[t * N * B:],
                        mparams=pl.train_mparams,
                        disc=None if pl.disc is None else pl.disc[t:],
                        reward=pl.train_reward[p][t * B:], returns=pl.train_returns[p],
                        err=pl.train_err[p][t * B:], final=self.tape[(t + 1) * S * B:])
        # returns and loss (planner.py:2747-2765, 2811-2819); entropy regulariser is value-only
        loss_out = pl.train_loss[p:p + 1] if loss_out is None else loss_out
        engine.drp_returns_loss(T, B, pl.train_reward[p], pl.disc, self.entropy if plan._stochastic else None,
                                self.ent_coeff, pl.train_returns[p], self.ent_b, loss_out)
        if pl.utility != "mean" or pl.use_symlog_reward:
            # a utility other than the mean (planner.py:2161-2247, 2817-2818): the return term of the loss
            # and the return cotangents the adjoint sweep is fed come from the utility kernels
            if pl.utility == "mean":
                engine.utility_loss("symlog_mean", B, pl.train_returns[p], 0.0, 0.0, 1.0 / pl.world_size,
                                    self._u_loss, pl.gret)
            else:
                pl._native_utility_kernels(pl.train_returns[p], self._u_loss)
            engine.mean_loss(pl.train_returns[p], B, self._m_loss, None)
            loss_out.add_(self._u_loss - self._m_loss)
        # ---- backward sweep
        grad.zero_()
        self.gs[0].zero_()
        if self.film:
            for d in self.dgb:
                d.zero_()
        cur = 0
        fused = self.fused_bwd and not self.film and 1 in self.tc_layers
        for t in range(T - 1, -1, -1):
            x = self.tape[t * S * B:(t + 1) * S * B]
            nxt, out = self.gs[cur], self.gs[cur ^ 1]
            if self.zact:
                self._Z = [Zl[t] for Zl in self.Z]
            gh = self.gheads_tape[t] if fused else self.gheads
            bwd.backward(B, 1, x, pl.gret, actions=self.act[t],
                         noise=None if pl.train_noise is None else pl.train_noise[t * N * B:],
                         mparams=pl.train_mparams, disc=None if pl.disc is None else pl.disc[t:],
                         gact=self.gact, gs0=out, gs_in=nxt)
            engine.drp_actions_backward(
                B, A, plan._stochastic, plan.action_segs, self.heads[t],
                None if self.pol_noise is None else self.pol_noise[t], 1.0, self.lo, self.hi,
                self.ls_lo, self.ls_hi, self.gact, gh, kinds=plan.action_kinds,
                bool_weight=plan._sigmoid_weight, ent_coeff=self.ent_coeff,
                inv_batch=1.0 / (B * pl.world_size), wrap=plan._wrap_non_bool,
                sigma_entropy_grad=plan._sigma_entropy_grad)
            if plan.use_topk:
                engine.drp_topk_backward(
                    B, A, plan._stochastic, plan.action_segs, plan.action_kinds, self.heads[t],
                    None if self.pol_noise is None else self.pol_noise[t], plan.allowed,
                    plan._softmax_output_weight, self.gact, gh,
                    ent_coeff=self.ent_coeff, inv_batch=1.0 / (B * pl.world_size))
            film = (t, [Hl[t] for Hl in self.Hf]) if self.film else None
            if self.Dn or self.pre:
                # MLP input cotangent -> gy, then back through LayerNorm (+ dscale, dbias) and the
                # static scaling into the state cotangent
                self.gy.zero_()
                xin = self.xn[t] if self.Dn else self.xa[t]
                self.mlp_backward(grad, B, xin, [Hl[t] for Hl in self.H], self.gy, self.xp[t], params,
                                  film=film, t=t)
                g_in = self.gy
                if self.Dn:
                    dst = out
                    if self.pre:
                        self.gya.zero_()
                        dst = self.gya
                    engine.state_layernorm_backward(
                        B, self.D, plan.state_segs, plan.norm_groups, plan._norm_eps,
                        self.xa[t] if self.pre else x, self._W(params, "ln_scale"), self.gy, dst,
                        self._W(grad, "ln_scale"), self.ln_ws)
                    g_in = dst
                if self.pre:
                    engine.state_scale(B, self.D, plan.state_segs, self.pre_lo, self.pre_hi, g_in, out,
                                       backward=True)
            else:
                self.mlp_backward(grad, B, x, [Hl[t] for Hl in self.H], out, self.xp[t], params,
                                  film=film, t=t)
            cur ^= 1
        if fused:
            self.wgrad_all(grad, self.xn if self.Dn else (self.xa if self.pre else self.tape))
        if self.film:
            self.film_param_grads(params, grad, emb)
        if plan.frozen_segs:
            # hidden state words of a partially observed model: their first-layer rows are structural
            # zeros, not parameters (planner.py:1300-1313 never feeds them to the network)
            off, _, c = plan.seg["W0"]
            for o, n in plan.frozen_segs:
                grad[off + o * c:off + (o + n) * c].zero_()

    def test_kernels(self, p: int, params: Optional[torch.Tensor] = None, out=None):
        """Exact-model rollouts with the test policy (train = 0, planner.py:1482-1499).
        ``out`` = (reward, returns, err, final, loss) overrides the logged test buffers."""
        pl, plan = self.pl, self.plan
        if out is None:
            out = (pl.test_reward[p], pl.test_returns[p], pl.test_err[p], pl.test_final[p],
                   pl.test_loss_buf[p:p + 1])
        o_reward, o_returns, o_err, o_final, o_loss = out
        T, B, A = pl.T, pl.batch_size_test, self.A
        params = pl.params[p] if params is None else params
        te = pl.test_fwd
        N = te.program.N
        engine.drp_set_activation(plan._activation)
        self.refresh_transposes(params)
        if self.film:
            self.film_tables(params)
        self.t_state[0].copy_(pl.s0_test)
        cur = 0
        for t in range(T):
            x, nxt = self.t_state[cur], self.t_state[cur ^ 1]
            xin = x
            if any(plan.state_is_int):      # bool / int state words of the exact model -> fp32
                xin = self.t_xf
                for (off, n), is_int in zip(plan.state_segs, plan.state_is_int):
                    blk = x[off * B:(off + n) * B]
                    xin[off * B:(off + n) * B].copy_(blk.view(torch.int32) if is_int else blk)
            if self.pre:
                engine.state_scale(B, self.D, plan.state_segs, self.pre_lo, self.pre_hi, xin,
                                   self.t_xa)
                xin = self.t_xa
            if self.Dn:
                self.layernorm_forward(params, B, xin, self.t_xn)
                xin = self.t_xn
            self._Z = self.t_Z
            self.mlp_forward(params, B, xin, self.t_H, self.t_heads,
                             film=(t, self.t_Hf) if self.film else None, keep_hidden=False)
            engine.drp_actions_forward(B, A, plan._stochastic, plan.action_segs, self.t_heads,
                                       None, 0.0, self.lo,
                                       self.hi, self.ls_lo, self.ls_hi, self.t_act, None,
                                       kinds=plan.action_kinds, bool_weight=plan._sigmoid_weight,
                                       typed=True, wrap=plan._wrap_non_bool)
            if plan.use_topk:
                engine.drp_topk_forward(B, A, plan._stochastic, plan.action_segs, plan.action_kinds,
                                        self.t_heads, None, False, plan.allowed,
                                        plan._softmax_output_weight, self.t_act, None, typed=True)
            te.forward(B, 1, x, actions=self.t_act,
                       noise=None if pl.test_noise is None else pl.test_noise[t * N * B:],
                       disc=None if pl.disc is None else pl.disc[t:],
                       reward=o_reward[t * B:], returns=o_returns,
                       err=o_err[t * B:], final=nxt)
            cur ^= 1
        o_final.copy_(self.t_state[cur])
        engine.drp_returns_loss(T, B, o_reward, pl.disc, None, None, o_returns, None, o_loss)
        if pl.utility != "mean":                     # the test loss applies the utility too (P:2695-2698)
            pl._test_loss_kernels(o_returns, o_loss)

    def redraw_noise(self, gen: torch.Generator):
        pl = self.pl
        if self.pol_noise is not None and getattr(pl, "_philox", False):
            # one Philox launch for the whole [T][A * B] policy-noise tensor: the action-noise law, with
            # Logistic noise on boolean logits (planner.py:1386) and Gumbel noise on pooled logits (1041, 1051)
            dist = self.plan._action_noise_dist
            sites = [("logistic" if kind == 1 else ("gumbel" if kind in (3, 4) else dist), off, n)
                     for (off, n), kind in zip(self.plan.action_segs, self.plan.action_kinds)]
            if len(sites) <= 64 and all(k in engine.NOISE_KINDS for k, _, _ in sites):
                engine.draw_noise(pl.T, pl.batch_size_train, sites, pl._noise_seed * 4 + 2, pl._draws,
                                  self.pol_noise)
                return
        if self.pol_noise is not None:
            dist = self.plan._action_noise_dist
            if dist == "normal":
                self.pol_noise.normal_(generator=gen)
            else:                       # standard laws as jax.random draws them
                u = torch.rand(self.pol_noise.shape, generator=gen,
                               device=self.pol_noise.device).clamp_(1e-7, 1 - 1e-7)
                if dist == "logistic":
                    self.pol_noise.copy_(torch.log(u) - torch.log1p(-u))
                elif dist == "gumbel":
                    self.pol_noise.copy_(-torch.log(-torch.log(u)))
                elif dist == "cauchy":
                    self.pol_noise.copy_(torch.tan(math.pi * (u - 0.5)))
                else:                   # laplace
                    self.pol_noise.copy_(torch.sign(u - 0.5) * -torch.log1p(-2 * (u - 0.5).abs()))
            B = self.pl.batch_size_train
            for (off, n), kind in zip(self.plan.action_segs, self.plan.action_kinds):
                if kind in (1, 3, 4):
                    blk = self.pol_noise[:, off * B:(off + n) * B]
                    u = torch.rand(blk.shape, generator=gen, device=blk.device).clamp_(1e-7, 1 - 1e-7)
                    if kind == 1:       # boolean logits take Logistic(0, 1) noise (planner.py:1386)
                        blk.copy_(torch.log(u) - torch.log1p(-u))
                    else:               # pooled logits take Gumbel(0, 1) noise (planner.py:1041, 1051)
                        blk.copy_(-torch.log(-torch.log(u)))

    # ---- measurement hooks of bench.py ------------------------------------------------------------
    def tape_bytes(self) -> int:
        """Bytes of hidden activations + state tape one update keeps in HBM for its backward sweep."""
        n = sum(H.numel() for H in self.H) + self.tape.numel()
        if self.fused_bwd:
            n += sum(g.numel() for g in self.gz_tape) + self.gheads_tape.numel()
        return 4 * n

    def _bench_roofline_fused(self, time_kernel, peaks, peak_src):
        """Fused path: the dominant kernel is rk_mlp2_kernel (one launch per decision epoch and direction).
        Timed alone: the train-sweep forward (keeps H0 / H1), the backward, and the three one-shot
        weight-gradient launches.  Algorithmic FLOPs of the forward launch = 2 B (D h0 + h0 h1 + h1 NH),
        fp32-equivalent; the hidden layer runs as 3 TF32 passes at half the bf16 rate, so the hardware
        ceiling of the kernel is peak / 6 (stated in the note)."""
        pl, plan = self.pl, self.plan
        B, T = pl.batch_size_train, pl.T
        params = pl.params[0]
        h0, h1 = self.hid
        act = plan._activation
        self.refresh_transposes(params)
        x = self.tape[:self.S * B]
        H = [Hl[0] for Hl in self.H]

        def k_fwd():
            engine.drp_mlp2_forward(act, B, self.D, h0, h1, self.NH, plan.state_segs, x, self.img_f,
                                    self._W(params, "b0"), self._W(self.wT_hi, "W1"),
                                    self._W(self.wT_lo, "W1"), self._W(params, "b1"),
                                    self._W(params, "bh"), H[0], H[1], self.heads[0])

        def k_bwd():
            engine.drp_mlp2_backward(act, B, self.D, h0, h1, self.NH, plan.state_segs,
                                     self.gheads_tape[0], H[0], H[1], self.img_b, self._W(self.w_hi, "W1"),
                                     self._W(self.w_lo, "W1"), self.gz_tape[0][0], self.gz_tape[1][0],
                                     self.gs[0])

        grad = torch.zeros_like(params)

        def k_wgrad():
            self.wgrad_all(grad, self.tape)
        t_fwd, t_bwd, t_wg = time_kernel(k_fwd), time_kernel(k_bwd), time_kernel(k_wgrad)
        self.gs[0].zero_()
        flops_fwd = 2.0 * B * (self.D * h0 + h0 * h1 + h1 * self.NH)
        flops_iter = T * flops_fwd * (3 + 1)          # train fwd + dX + dW, + the test forward
        tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        achieved_tf = flops_fwd / (t_fwd * 1e-3) / 1e12
        roofline = {"bound": "tensor",
                    "kernel": "rk_mlp2_kernel<forward> (layer 0 + hidden layer tcgen05 3xTF32 + heads, one launch "
                              "per decision epoch; train sweep, H0 / H1 kept)",
                    "achieved": achieved_tf, "peak": tensor_peak, "unit": "TFLOP/s",
                    "frac": achieved_tf / tensor_peak,
                    # dram__bytes of one launch from the committed ncu capture of this shape
                    # (profiles/r02_ncu_mlp2.md: 1.39 MB read + 14.73 MB written by the end of the launch;
                    # the rest of the 67 MB of H0 / H1 is still in the 126 MB L2 then)
                    "traffic": 16116000 if (B, h0, h1) == (32768, 256, 256) else None,
                    "peak_source": peak_src,
                    "algorithmic_flops": flops_fwd, "executed_tf32_tflops": 3 * achieved_tf,
                    "mlp_tflop_per_iteration": flops_iter / 1e12,
                    "hbm_bytes_per_launch": 4.0 * B * (h0 + h1 + self.NH + self.D),
                    "note": "algorithmic (fp32-equivalent) FLOPs of the three layers over the launch time, "
                            "against the sustained bf16 tensor peak; 3 TF32 passes per product at half the "
                            "bf16 rate put the hardware ceiling of the hidden layer at peak/6, and 256 row "
                            "tiles on 148 SMs (two per SM) at 0.86 of that"}
        # launches per iteration: per epoch (policy, actions, rollout step) x {train fwd, train bwd, test};
        # weight gradients: 3 products + 4 reductions + column sums; transposes / splits / images / losses
        n_launch = T * 3 * 3 + 10 + 24
        return roofline, {"mlp2_forward": t_fwd, "mlp2_backward": t_bwd, "weight_gradients_all_epochs": t_wg}, \
            n_launch

    def bench_roofline(self, time_kernel, peaks, peak_src):
        """(roofline, kernels_ms, launches per iteration) of the dominant kernel of a DRP iteration: the
        last hidden dense layer (the only dense contraction of the path), timed alone through
        ``time_kernel`` (CUDA events, L2 flushed).  Algorithmic FLOPs are the fp32-equivalent
        2 * B * K * N of Y = act(X W + b); the tensor cores execute 3 TF32 passes per product (the
        error-compensated split the 1e-5 / 1e-4 parity tolerances need) at half the bf16 rate, so
        the hardware ceiling of this kernel is peak / 6."""
        pl = self.pl
        B, T = pl.batch_size_train, pl.T
        params = pl.params[0]
        hid = self.hid
        if self.fused_bwd and not self.film and 1 in self.tc_layers:
            return self._bench_roofline_fused(time_kernel, peaks, peak_src)
        li = len(hid) - 1
        h_in, h_out = (hid[-2] if len(hid) > 1 else self.D), hid[-1]
        src = self.H[-2][0] if len(hid) > 1 else None
        use_tc = li in self.tc_layers

        def k_gemm():
            if use_tc:
                engine.tcgemm(2, B, h_out, h_in, src, h_in, self._W(self.wT_hi, f"W{li}"),
                              self._W(self.wT_lo, f"W{li}"), h_in, self.H[-1][0], h_out,
                              bias=self._W(params, f"b{li}"))
            else:
                engine.sgemm(engine.EPI_BIAS_TANH, B, h_out, h_in, src, h_in,
                             self._W(params, f"W{li}"), h_out, self.H[-1][0], h_out,
                             bias=self._W(params, f"b{li}"))
        t_gemm = time_kernel(k_gemm) if src is not None else float("nan")
        dims = [self.D] + list(hid) + [self.NH]
        mlp_flops = 2.0 * B * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
        flops_iter = T * mlp_flops * (3 + 1)          # train fwd + dX + dW, + the test forward
        tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        achieved_tf = 2.0 * B * h_in * h_out / (t_gemm * 1e-3) / 1e12
        roofline = {"bound": "tensor",
                    "kernel": ("rk_tcgemm_kernel (hidden dense layer, tcgen05 3xTF32, bias+tanh "
                               "epilogue)" if use_tc else "rk_sgemm_kernel (hidden dense layer, fp32 SIMT)"),
                    "achieved": achieved_tf, "peak": tensor_peak, "unit": "TFLOP/s",
                    "frac": achieved_tf / tensor_peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_flops": 2.0 * B * h_in * h_out,
                    "executed_tf32_tflops": (3 if use_tc else 1) * achieved_tf,
                    "mlp_tflop_per_iteration": flops_iter / 1e12,
                    "note": "algorithmic (fp32-equivalent) FLOPs of Y = tanh(X W + b) over the kernel "
                            "time, against the sustained bf16 tensor peak; 3 TF32 passes per product at "
                            "half the bf16 rate put the hardware ceiling of this kernel at peak/6"}
        n_state = len(self.plan.state_segs)
        L = len(hid)
        n_launch = T * ((n_state + L) + 2) \
            + T * (2 + 2 * 2 + 1 + L * 2 + (n_state * 2 - 1) * 2 + (L - 1) + n_state) \
            + T * ((n_state + L) + 2) + 12
        return roofline, {"hidden_layer_gemm": t_gemm}, n_launch
