"""CPU ORACLE (test infrastructure, not product code) -- PARITY UNPINNED.

A torch-CPU fp32 restatement of the reference's expression semantics:
  * exact model      : pyRDDLGym_jax/core/compiler.py:936-2241
  * relaxed model    : pyRDDLGym_jax/core/logic.py:83-1789
  * one step / scan  : pyRDDLGym_jax/core/compiler.py:360-376, 490-619
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package.

"Parity unpinned": the reference ships no tests, golden vectors or fixtures, and none of its
dependencies (jax, optax, flax, softjax, pyRDDLGym) can be imported here, so this oracle is
pinned only by closed-form known-answer tests (tests/test_oracle_kat.py) and finite
differences, not by outputs of the reference itself.

Conventions
-----------
* Every value is a torch tensor of shape [B or 1, *keepdims(scope)]: size 1 on scope axes
  the expression does not depend on, so broadcasting reproduces the reference's
  expand_dims/broadcast_to (compiler.py:1073-1075).
* Randomness is an explicit input: ``noise`` is a list with one tensor per draw site, in
  evaluation order (CPF level order, then depth-first left-to-right as the reference splits
  its key, e.g. compiler.py:1792), each shaped [B, *draw_shape] where draw_shape is the shape
  the reference passes to jax.random (the shape of the distribution's first argument:
  full scope if that argument involves any pvariable, scalar if it is a pure constant).
* jax.lax.stop_gradient == ``.detach()``.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from pyrddlgym_jax_b200.rddl.expr import Expr
from pyrddlgym_jax_b200.rddl.trace import Tracer

REAL = torch.float32
INT = torch.int32

ERR = {  # compiler.py:861-888
    "CAST": 1 << 0, "UNIFORM": 1 << 1, "NORMAL": 1 << 2, "EXPONENTIAL": 1 << 3,
    "WEIBULL": 1 << 4, "BERNOULLI": 1 << 5, "POISSON": 1 << 6, "GAMMA": 1 << 7,
    "BETA": 1 << 8, "GEOMETRIC": 1 << 9, "PARETO": 1 << 10, "STUDENT": 1 << 11,
    "GUMBEL": 1 << 12, "LAPLACE": 1 << 13, "CAUCHY": 1 << 14, "GOMPERTZ": 1 << 15,
    "CHISQUARE": 1 << 16, "KUMARASWAMY": 1 << 17, "DISCRETE": 1 << 18, "KRON": 1 << 19,
    "DIRICHLET": 1 << 20, "MULTIVARIATE_STUDENT": 1 << 21, "MULTINOMIAL": 1 << 22,
    "BINOMIAL": 1 << 23, "NEGATIVE_BINOMIAL": 1 << 24,
}
COUNT_NBINS = 100      # truncation of the exact Poisson / Binomial samplers (see _random)


# ---- custom-derivative helpers ------------------------------------------------------
class _SafeLog(torch.autograd.Function):
    """compiler.py:65-75: value log(x) (-inf at x<=0), tangent xdot / (x + eps)."""

    @staticmethod
    def forward(ctx, x, eps):
        ctx.save_for_backward(x)
        ctx.eps = eps
        pos = x > 0
        return torch.where(pos, torch.log(torch.where(pos, x, torch.ones_like(x))),
                           torch.full_like(x, -math.inf))

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g / (x + ctx.eps), None


def safe_log(x, eps=1e-12):
    return _SafeLog.apply(x, eps)


class _StableTanh(torch.autograd.Function):
    """logic.py:57-73."""

    @staticmethod
    def forward(ctx, x):
        ax = torch.abs(x)
        e = torch.expm1(2. * ax)
        small = torch.where(ax < 20., e / (e + 2.), 1. - 2. * torch.exp(-2. * ax))
        t = torch.sign(x) * small
        ctx.save_for_backward(t)
        return t

    @staticmethod
    def backward(ctx, g):
        (t,) = ctx.saved_tensors
        return g * (1. - t * t)


def stable_tanh(x):
    return _StableTanh.apply(x)


class _SafeSqrt(torch.autograd.Function):
    """softjax ``sj.sqrt`` (assumed): sqrt value, derivative 0 instead of inf/nan at x<=0."""

    @staticmethod
    def forward(ctx, x):
        y = torch.sqrt(x)
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        return torch.where(x > 0, g * 0.5 / torch.where(x > 0, y, torch.ones_like(y)),
                           torch.zeros_like(g))


def soft_argmax(logits, w):
    """sum_k k softmax(w logits)_k over the last axis (SoftmaxArgmax.soft_argmax, logic.py:377-379)."""
    lit = torch.arange(logits.shape[-1], dtype=REAL)
    return torch.sum(lit * torch.softmax(w * logits, dim=-1), dim=-1)


def safe_sqrt(x):
    return _SafeSqrt.apply(x)


class _SafeDiv(torch.autograd.Function):
    """softjax ``sj.div`` (assumed): x / y value, zero cotangents where y == 0."""

    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x, y)
        return x / y

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        ok = y != 0
        ys = torch.where(ok, y, torch.ones_like(y))
        gx = torch.where(ok, g / ys, torch.zeros_like(g))
        gy = torch.where(ok, -g * x / (ys * ys), torch.zeros_like(g))
        return _unbroadcast(gx, x.shape), _unbroadcast(gy, y.shape)


def _unbroadcast(g, shape):
    while g.dim() > len(shape):
        g = g.sum(0)
    for i, s in enumerate(shape):
        if s == 1 and g.shape[i] != 1:
            g = g.sum(i, keepdim=True)
    return g


def safe_div(x, y):
    x, y = torch.broadcast_tensors(_f(x), _f(y))
    return _SafeDiv.apply(x, y)


def _f(x):
    return x if x.dtype == REAL else x.to(REAL)


def _at_least_int(x):
    return x.to(INT) if x.dtype == torch.bool else x


def _ste(soft, hard):
    """soft + stop_gradient(hard - soft)  (e.g. logic.py:272)."""
    return soft + (hard.to(soft.dtype) - soft).detach()


def _fix_int(x):
    return x.to(INT) if x.dtype == torch.int64 else x


class NoiseCursor:
    """Hands out the supplied noise tensors in draw order, recording the draw spec."""

    def __init__(self, noise: Optional[List[torch.Tensor]]):
        self.noise = noise
        self.pos = 0
        self.spec: List[Tuple[str, Tuple[int, ...]]] = []

    def draw(self, kind: str, shape: Tuple[int, ...], batch: int, extra=None) -> torch.Tensor:
        self.spec.append((kind, tuple(shape)) if extra is None else (kind, tuple(shape), extra))
        if self.noise is None:
            raise RuntimeError("oracle requires a supplied noise list")
        z = self.noise[self.pos]
        self.pos += 1
        assert tuple(z.shape) == (batch,) + tuple(shape), (z.shape, batch, shape, kind)
        return z


class OracleModel:
    """Evaluates one RDDL model (exact or relaxed) the way the reference's compiled closures
    do.  ``relax`` is None for the exact model (JaxRDDLCompiler) or the flattened relaxation
    config (core/logic.py relax_config()) for JaxRDDLCompilerWithGrad subclasses.
    """

    def __init__(self, model, relax: Optional[Dict[str, Any]] = None):
        self.model = model
        self.relax = relax
        self.relaxed = relax is not None
        self.tracer = Tracer(model)
        self.variants = (relax or {}).get("variants", {})
        # model_params: (expr.id, name) -> float   (logic.py:266 aux['params'][id])
        self.model_params: Dict[Tuple[int, str], float] = {}
        self.cpf_order = [c for lv in model.levels.values() for c in lv]

    # ------------------------------------------------------------------------------
    def init_fls_nfls(self, batch: int, subs: Optional[Dict[str, np.ndarray]] = None):
        """compiler.py:668-727 (+ float cast, logic.py:144-147)."""
        m = self.model
        fls, nfls = {}, {}
        for name, val in m.init_values.items():
            if name.endswith("'"):
                continue
            if subs is not None and name in subs:
                val = np.reshape(np.asarray(subs[name]), np.shape(val)).astype(val.dtype)
            t = torch.from_numpy(np.ascontiguousarray(val))
            if self.relaxed:
                t = t.to(REAL)
            if name in m.non_fluents:
                nfls[name] = t
            else:
                fls[name] = t.unsqueeze(0).repeat(batch, *([1] * t.dim())).contiguous()
        for s, sp in m.next_state.items():
            fls[sp] = fls[s]
        return fls, nfls

    def param(self, expr: Expr, name: str, default: float) -> float:
        return self.model_params.setdefault((expr.id, name), float(default))

    # ------------------------------------------------------------------------------
    def _shape(self, scope):
        return self.tracer.scope_sizes(scope)

    def _ref_full(self, e: Expr) -> bool:
        """Whether the reference's value for this node carries the full scope shape
        (True as soon as any pvariable / free variable is involved) or is a 0-d constant."""
        fam, op = e.etype
        if fam == "constant":
            return False
        if fam == "pvar":
            name, params = e.args
            return not (params is None and self.tracer.is_object(name))
        if fam in ("aggregation", "matrix", "randomvector"):
            return True
        if fam == "control" and op == "switch":
            return True
        if fam == "randomvar" and op in ("Discrete", "UnnormDiscrete"):
            return True
        return any(self._ref_full(a) for a in e.args)

    def ev(self, e: Expr, scope, fls, nfls, cur: NoiseCursor, B: int):
        """-> (value [B|1, *keepdims], err [B] int64).  Dispatch mirrors compiler.py:936-964."""
        fam, op = e.etype
        z = torch.zeros(B, dtype=torch.int64)
        rank = len(scope)
        if fam == "constant":      # compiler.py:1000-1008
            dt = {"bool": torch.bool, "int": INT, "real": REAL}[op]
            if self.relaxed and False:
                dt = REAL
            return torch.tensor(e.args, dtype=dt).reshape((1,) * (1 + rank)), z
        if fam == "pvar":          # compiler.py:1017-1081
            acc = self.tracer.access(e, scope)
            if acc.kind == "freevar":
                n = acc.keep_shape[acc.axes[0]]
                return torch.arange(n, dtype=INT).reshape((1,) + acc.keep_shape), z
            if acc.kind == "object":
                return torch.tensor(acc.value, dtype=INT).reshape((1,) * (1 + rank)), z
            src = fls[acc.name] if acc.name in fls else nfls[acc.name]
            batched = acc.name in fls
            err = z
            if acc.nested:        # compiler.py:1046-1064 gather by computed indices
                sizes = self._shape(scope)
                vshape = self.model.variable_shape(acc.name)
                base = torch.from_numpy(acc.index).reshape(
                    (1,) + tuple(sizes[i] if i in acc.axes else 1 for i in range(rank)))
                flat = base
                for d, sub in acc.nested:
                    v, er = self.ev(sub, scope, fls, nfls, cur, B)
                    err = err | er
                    flat = flat + v.to(torch.int64) * int(acc.strides[d])
                total = int(np.prod(vshape))
                flat = flat.clamp(0, total - 1)
                full = (B if (batched or flat.shape[0] == B) else 1,) + tuple(sizes)
                flat = flat.expand(full)
                if batched:
                    s2 = src.reshape(B, -1)
                    out = torch.gather(s2, 1, flat.reshape(full[0], -1).expand(B, -1))
                    return out.reshape((B,) + tuple(sizes)), err
                out = src.reshape(-1)[flat.reshape(-1)]
                return out.reshape(full), err
            if acc.identity:          # whole tensor in declaration order: a view
                if batched:
                    return src.reshape((B,) + acc.keep_shape), err
                return src.reshape((1,) + acc.keep_shape), err
            idx = torch.from_numpy(acc.index.reshape(-1))
            if batched:
                out = src.reshape(B, -1)[:, idx].reshape((B,) + acc.keep_shape)
            else:
                out = src.reshape(-1)[idx].reshape((1,) + acc.keep_shape)
            return out, err
        if fam == "arithmetic":
            return self._arithmetic(e, scope, fls, nfls, cur, B)
        if fam == "relational":
            return self._relational(e, scope, fls, nfls, cur, B)
        if fam == "boolean":
            return self._logical(e, scope, fls, nfls, cur, B)
        if fam == "aggregation":
            return self._aggregation(e, scope, fls, nfls, cur, B)
        if fam == "func":
            return self._function(e, scope, fls, nfls, cur, B)
        if fam == "control":
            return self._control(e, scope, fls, nfls, cur, B)
        if fam == "randomvar":
            return self._random(e, scope, fls, nfls, cur, B)
        if fam == "matrix":
            return self._matrix(e, scope, fls, nfls, cur, B)
        if fam == "randomvector":
            return self._random_vector(e, scope, fls, nfls, cur, B)
        raise NotImplementedError(f"expression type {e.etype} is not supported")

    # -- matrix algebra: compiler.py:2383-2434 (jnp.linalg on the last two axes) ------------------
    def _matrix(self, e, scope, fls, nfls, cur, B):
        _, op = e.etype
        if op == "det":                                   # compiler.py:2398-2407
            *tv, body = e.args
            inner = list(scope) + list(tv)
            x, err = self.ev(body, inner, fls, nfls, cur, B)
            x = _f(x).expand((x.shape[0],) + tuple(self._shape(inner)))
            return torch.linalg.det(x), err
        (row, col), body = e.args                         # compiler.py:2409-2434
        names = [v for v, _ in scope]
        r_ax, c_ax = 1 + names.index(row), 1 + names.index(col)
        x, err = self.ev(body, scope, fls, nfls, cur, B)
        x = _f(x).expand((x.shape[0],) + tuple(self._shape(scope)))
        m = torch.movedim(x, (r_ax, c_ax), (-2, -1))
        fn = {"inverse": torch.linalg.inv, "pinverse": torch.linalg.pinv,
              "cholesky": torch.linalg.cholesky}[op]
        return torch.movedim(fn(m), (-2, -1), (r_ax, c_ax)), err

    # -- random vectors: compiler.py:2247-2377 -----------------------------------------------------
    def _random_vector(self, e, scope, fls, nfls, cur, B):
        _, name = e.etype
        svars, *args = e.args
        names = [v for v, _ in scope]
        i_ax = 1 + names.index(svars[0])
        sizes = tuple(self._shape(scope))

        def oob(cond):
            c = cond.reshape(cond.shape[0], -1).all(1)
            return (~(c.expand(B) if c.shape[0] != B else c)).to(torch.int64)

        if name in ("MultivariateNormal", "MultivariateStudent"):
            inner = list(scope) + [(svars[1], scope[i_ax - 1][1])]
            mean, err = self.ev(args[0], scope, fls, nfls, cur, B)
            cov, er = self.ev(args[1], inner, fls, nfls, cur, B)
            err = err | er
            mean = _f(mean).expand((mean.shape[0],) + sizes)
            cov = _f(cov).expand((cov.shape[0],) + tuple(self._shape(inner)))
            Z = cur.draw("normal", sizes, B)                           # compiler.py:2276-2278
            if name == "MultivariateStudent":                          # compiler.py:2307-2321
                df, er = self.ev(args[2], scope, fls, nfls, cur, B)
                df = _f(df).expand((df.shape[0],) + sizes)
                err = err | er | oob(df > 0) * ERR["MULTIVARIATE_STUDENT"]
                G = cur.draw("gamma", sizes, B, (0.5 * df[0]).detach().numpy()) \
                    if not (df.requires_grad or df.shape[0] == B) else None
                if G is None:
                    raise NotImplementedError("MultivariateStudent with a fluent df")
                Z = Z * torch.sqrt(0.5 * df / G)
            L = torch.linalg.cholesky(torch.movedim(cov, (i_ax, cov.dim() - 1), (-2, -1)))
            z = torch.movedim(Z, i_ax, -1).unsqueeze(-1)
            out = torch.matmul(L, z)[..., 0]                            # compiler.py:2281-2283
            return torch.movedim(out, -1, i_ax) + mean, err
        if name == "Dirichlet":                                         # compiler.py:2331-2348
            alpha, err = self.ev(args[0], scope, fls, nfls, cur, B)
            alpha = _f(alpha).expand((alpha.shape[0],) + sizes)
            err = err | oob(alpha > 0) * ERR["DIRICHLET"]
            if alpha.shape[0] == B and B > 1:
                raise NotImplementedError("Dirichlet with a fluent concentration")
            G = cur.draw("gamma", sizes, B, alpha[0].detach().numpy())
            return G / G.sum(i_ax, keepdim=True), err
        if name == "Multinomial":                                       # compiler.py:2350-2377
            # same law as the tfp sampler from COUNT_NBINS supplied uniforms (see core/lowering.py)
            trials, err = self.ev(args[0], scope, fls, nfls, cur, B)
            prob, er = self.ev(args[1], scope, fls, nfls, cur, B)
            trials = trials.to(INT) if trials.dtype.is_floating_point else _at_least_int(trials)
            prob = _f(prob).expand((prob.shape[0],) + sizes)
            cum = torch.cumsum(prob, i_ax)
            total = cum.select(i_ax, sizes[i_ax - 1] - 1).unsqueeze(i_ax)
            ok = (prob >= 0) & ((total - 1.0).abs() <= 1e-5) & (trials >= 0)
            err = err | er | oob(ok) * ERR["MULTINOMIAL"]
            n = sizes[i_ax - 1]
            counts = torch.zeros((B,) + sizes, dtype=INT)
            for k in range(COUNT_NBINS):
                U = cur.draw("uniform", sizes, B).select(i_ax, 0).unsqueeze(i_ax)
                below = U < cum.detach()
                prev = torch.cat([torch.zeros_like(below.narrow(i_ax, 0, 1)),
                                  below.narrow(i_ax, 0, n - 1)], i_ax)
                last = torch.arange(n).reshape([n if d == i_ax else 1 for d in range(below.dim())]) == n - 1
                hit = torch.where(last, ~prev, below & ~prev)
                counts = counts + (hit & (k < trials)).to(INT)
            return (counts.to(REAL) if self.relaxed else counts), err
        raise NotImplementedError(f"Distribution {name} is not supported.")

    def _args(self, e, scope, fls, nfls, cur, B, args=None):
        vals, err = [], torch.zeros(B, dtype=torch.int64)
        for a in (e.args if args is None else args):
            v, er = self.ev(a, scope, fls, nfls, cur, B)
            vals.append(v)
            err = err | er
        return vals, err

    def _cast_err(self, x, dtype_ok: bool, B: int):
        """check_dtype branches of compiler.py:1094-1096, 1112-1117: static dtype test."""
        if dtype_ok:
            return torch.zeros(B, dtype=torch.int64)
        return torch.full((B,), ERR["CAST"], dtype=torch.int64)

    # -- arithmetic: compiler.py:1150-1183 (exact in both models) -----------------------
    def _arithmetic(self, e, scope, fls, nfls, cur, B):
        _, op = e.etype
        vals, err = self._args(e, scope, fls, nfls, cur, B)
        vals = [_at_least_int(v) for v in vals]
        if op == "-" and len(vals) == 1:
            return torch.neg(vals[0]), err
        fn = {"+": torch.add, "-": torch.sub, "*": torch.mul, "/": torch.true_divide}[op]
        out = vals[0]
        for v in vals[1:]:          # left fold, compiler.py:1140-1143
            out = fn(out, v)
        return out, err

    # -- relational: compiler.py:1189-1227 ; logic.py:262-344 ---------------------------
    def _relational(self, e, scope, fls, nfls, cur, B):
        _, op = e.etype
        (x, y), err = self._args(e, scope, fls, nfls, cur, B)
        x, y = _at_least_int(x), _at_least_int(y)
        exact = {">=": torch.ge, "<=": torch.le, ">": torch.gt, "<": torch.lt,
                 "==": torch.eq, "~=": torch.ne}[op]
        if not (self.relaxed and self.variants.get("relational") == "sigmoid"
                and self.tracer.is_fluent(e)):
            return exact(x, y), err
        w = self.param(e, "sigmoid_weight", self.relax["sigmoid_weight"])
        x, y = _f(x), _f(y)
        if op in (">", ">="):
            soft = torch.sigmoid(w * (x - y))
            ste = self.relax["use_sigmoid_ste"]
        elif op in ("<", "<="):
            soft = torch.sigmoid(w * (y - x))
            ste = self.relax["use_sigmoid_ste"]
        elif op == "==":
            soft = 1. - torch.square(stable_tanh(w * (y - x)))
            ste = self.relax["use_tanh_ste"]
        else:
            soft = torch.square(stable_tanh(w * (y - x)))
            ste = self.relax["use_tanh_ste"]
        return (_ste(soft, exact(x, y)) if ste else soft), err

    # -- logical: compiler.py:1233-1274 ; logic.py:424-778 -----------------------------
    def _logical(self, e, scope, fls, nfls, cur, B):
        _, op = e.etype
        vals, err = self._args(e, scope, fls, nfls, cur, B)
        variant = self.variants.get("logic") if self.relaxed else None
        if variant is None or not self.tracer.is_fluent(e):
            ok = all(v.dtype == torch.bool for v in vals)
            err = err | self._cast_err(None, ok, B)
            b = [v != 0 if v.dtype != torch.bool else v for v in vals]
            if op == "~" and len(b) == 1:
                return torch.logical_not(b[0]), err
            if op == "~":
                return torch.logical_xor(b[0], b[1]), err
            if op in ("^", "&", "|"):
                fn = torch.logical_and if op != "|" else torch.logical_or
                out = b[0]
                for v in b[1:]:
                    out = fn(out, v)
                return out, err
            if op == "=>":
                return torch.logical_or(torch.logical_not(b[0]), b[1]), err
            if op == "<=>":
                return torch.eq(vals[0], vals[1]), err
            raise NotImplementedError(op)
        ste = self.relax["use_logic_ste"]
        relu = torch.relu

        def hard2(kind, x, y):
            if kind == "and":
                return (x > 0.5) & (y > 0.5)
            if kind == "or":
                return (x > 0.5) | (y > 0.5)
            if kind == "xor":
                return (x > 0.5) ^ (y > 0.5)
            if kind == "implies":
                return (x <= 0.5) | (y > 0.5)
            return ((x <= 0.5) | (y > 0.5)) & ((y <= 0.5) | (x > 0.5))

        def soft2(kind, x, y):
            if variant == "product":       # logic.py:454, 467, 480, 493, 506
                return {"and": lambda: x * y,
                        "or": lambda: 1. - (1. - x) * (1. - y),
                        "xor": lambda: (1. - (1. - x) * (1. - y)) * (1. - x * y),
                        "implies": lambda: 1. - x * (1. - y),
                        "equiv": lambda: (1. - x * (1. - y)) * (1. - y * (1. - x))}[kind]()
            if variant == "godel":         # logic.py:573, 586, 599, 612, 625
                return {"and": lambda: torch.minimum(x, y),
                        "or": lambda: torch.maximum(x, y),
                        "xor": lambda: torch.minimum(torch.maximum(x, y),
                                                     1. - torch.minimum(x, y)),
                        "implies": lambda: torch.maximum(1. - x, y),
                        "equiv": lambda: torch.minimum(torch.maximum(1. - x, y),
                                                       torch.maximum(1. - y, x))}[kind]()
            # lukasiewicz: logic.py:692, 705, 718, 731, 744
            return {"and": lambda: relu(x + y - 1.),
                    "or": lambda: 1. - relu(1. - x - y),
                    "xor": lambda: relu(1. - torch.abs(1. - x - y)),
                    "implies": lambda: 1. - relu(x - y),
                    "equiv": lambda: relu(1. - torch.abs(x - y))}[kind]()

        def binop(kind, x, y):
            # jnp.multiply etc. promote bool/int operands; work in REAL
            x, y = _f(x), _f(y)
            s = soft2(kind, x, y)
            return _ste(s, hard2(kind, x, y)) if ste else s

        if op == "~" and len(vals) == 1:      # logic.py:436-447 (same for all three norms)
            x = _f(vals[0])
            s = 1. - x
            return (_ste(s, x <= 0.5) if ste else s), err
        if op == "~":
            return binop("xor", vals[0], vals[1]), err
        if op in ("^", "&", "|"):
            out = vals[0]
            for v in vals[1:]:
                out = binop("and" if op != "|" else "or", out, v)
            return out, err
        if op == "=>":
            return binop("implies", vals[0], vals[1]), err
        if op == "<=>":
            return binop("equiv", vals[0], vals[1]), err
        raise NotImplementedError(op)

    # -- aggregation: compiler.py:1280-1348 ; logic.py:381-417, 516-540 ... --------------
    def _aggregation(self, e, scope, fls, nfls, cur, B):
        _, op = e.etype
        *tv, body = e.args
        inner = list(scope) + list(tv)
        x, err = self.ev(body, inner, fls, nfls, cur, B)
        sizes = self._shape(inner)
        x = x.expand((x.shape[0],) + tuple(sizes))
        axes = tuple(range(1 + len(scope), 1 + len(inner)))
        fluent = self.tracer.is_fluent(e)
        is_bool = op in ("forall", "exists")
        if is_bool:
            variant = self.variants.get("logic") if self.relaxed else None
            if variant is None or not fluent:
                err = err | self._cast_err(None, x.dtype == torch.bool, B)
                xb = x if x.dtype == torch.bool else x != 0
                out = xb
                for ax in reversed(axes):
                    out = out.all(ax) if op == "forall" else out.any(ax)
                return out, err
            xf = _f(x)
            ste = self.relax["use_logic_ste"]
            if variant == "product":         # logic.py:521, 534
                soft = _prod(xf, axes) if op == "forall" else 1. - _prod(1. - xf, axes)
            elif variant == "godel":         # logic.py:640, 653
                soft = _amin(xf, axes) if op == "forall" else _amax(xf, axes)
            else:                            # logic.py:759, 772
                soft = torch.relu(torch.sum(xf - 1., dim=axes) + 1.) if op == "forall" \
                    else 1. - torch.relu(torch.sum(-xf, dim=axes) + 1.)
            if ste:
                hb = xf > 0.5
                for ax in reversed(axes):
                    hb = hb.all(ax) if op == "forall" else hb.any(ax)
                soft = _ste(soft, hb)
            return soft, err
        x = _at_least_int(x)
        if op in ("argmin", "argmax"):
            if len(axes) != 1:
                raise NotImplementedError("argmin/argmax over more than one variable")
            ax = axes[0]
            if self.relaxed and self.variants.get("argmax") == "softmax" and fluent:
                w = self.param(e, "argmax_weight", self.relax["argmax_weight"])
                xs = _f(x) if op == "argmax" else -_f(x)
                lit = torch.arange(xs.shape[ax], dtype=REAL).reshape(
                    [-1 if i == ax else 1 for i in range(xs.dim())])
                soft = torch.sum(lit * torch.softmax(w * xs, dim=ax), dim=ax)  # logic.py:377-379
                if self.relax["use_argmax_ste"]:
                    soft = _ste(soft, torch.argmax(xs, dim=ax))
                return soft, err
            fn = torch.argmax if op == "argmax" else torch.argmin
            return fn(x, dim=ax).to(INT), err
        if op == "sum":
            return _fix_int(torch.sum(x, dim=axes)), err
        if op == "avg":
            return torch.mean(_f(x), dim=axes), err
        if op == "prod":
            return _fix_int(_prod(x, axes)), err
        if op == "minimum":
            return _amin(x, axes), err
        if op == "maximum":
            return _amax(x, axes), err
        raise NotImplementedError(op)

    # -- functions: compiler.py:1354-1529 ; logic.py:346-358, 785-896 --------------------
    def _function(self, e, scope, fls, nfls, cur, B):
        _, op = e.etype
        vals, err = self._args(e, scope, fls, nfls, cur, B)
        vals = [_at_least_int(v) for v in vals]
        fluent = self.tracer.is_fluent(e)
        R = self.relax if self.relaxed else {}
        V = self.variants if self.relaxed else {}
        if len(vals) == 1:
            x = vals[0]
            if op == "sgn" and V.get("sgn") == "tanh" and fluent:     # logic.py:346-358
                w = self.param(e, "sigmoid_weight", R["sigmoid_weight"])
                soft = stable_tanh(w * _f(x))
                return (_ste(soft, torch.sign(_f(x))) if R["use_tanh_ste"] else soft), err
            if op in ("floor", "ceil") and V.get("floor") == "soft" and fluent:
                w = self.param(e, "floor_weight", R["floor_weight"])
                xf = _f(x)
                if op == "floor":                                      # logic.py:806-818
                    soft, hard = soft_floor(xf, w), torch.floor(xf)
                else:                                                  # logic.py:820-832
                    soft, hard = -soft_floor(-xf, w), torch.ceil(xf)
                return (_ste(soft, hard) if R["use_floor_ste"] else soft), err
            if op == "round" and V.get("round") == "soft" and fluent:  # logic.py:884-896
                w = self.param(e, "round_weight", R["round_weight"])
                xf = _f(x)
                soft = soft_round(xf, w)
                return (_ste(soft, torch.round(xf)) if R["use_round_ste"] else soft), err
            if op in ("abs", "sgn", "round", "floor", "ceil"):
                fn = {"abs": torch.abs, "sgn": torch.sign, "round": torch.round,
                      "floor": torch.floor, "ceil": torch.ceil}[op]
                return fn(x), err
            xf = _f(x)
            if op == "sqrt":
                return safe_sqrt(xf), err                              # compiler.py:1487
            fn = {"cos": torch.cos, "sin": torch.sin, "tan": torch.tan, "acos": torch.acos,
                  "asin": torch.asin, "atan": torch.atan, "cosh": torch.cosh,
                  "sinh": torch.sinh, "tanh": torch.tanh, "exp": torch.exp, "ln": torch.log,
                  "lngamma": torch.lgamma,
                  "gamma": lambda t: torch.exp(torch.lgamma(t))}[op]
            return fn(xf), err
        x, y = vals
        if op in ("div", "mod") and V.get("floor") == "soft" and fluent:  # logic.py:834-861
            w = self.param(e, "floor_weight", R["floor_weight"])
            xf, yf = _f(x), _f(y)
            d = soft_floor(safe_div(xf, yf), w)
            if R["use_floor_ste"]:
                d = _ste(d, torch.floor_divide(xf, yf))
            return (d if op == "div" else xf - yf * d), err
        if op == "div":
            return torch.floor_divide(x, y), err
        if op == "mod":
            return torch.remainder(x, y), err
        if op == "fmod":
            return torch.fmod(x, y), err
        if op == "min":
            return torch.minimum(*_promote(x, y)), err
        if op == "max":
            return torch.maximum(*_promote(x, y)), err
        if op == "pow":
            return torch.pow(x, y), err
        if op == "log":
            return torch.log(_f(x)) / torch.log(_f(y)), err
        if op == "hypot":
            return torch.hypot(_f(x), _f(y)), err
        raise NotImplementedError(op)

    # -- control: compiler.py:1617-1685 ; logic.py:903-1003 ------------------------------
    def _control(self, e, scope, fls, nfls, cur, B):
        _, op = e.etype
        if op == "if":
            pred_e = e.args[0]
            (c, a, b), err = self._args(e, scope, fls, nfls, cur, B)
            if self.relaxed and self.variants.get("if") == "linear" \
                    and self.tracer.is_fluent(pred_e):
                cf = _f(c)
                if self.relax["use_if_else_ste"]:                      # logic.py:933-935
                    cf = _ste(cf, cf > 0.5)
                # sj.where(c, a, b) taken as the linear blend c*a + (1-c)*b (logic.py:936)
                return cf * _f(a) + (1. - cf) * _f(b), err
            err = err | self._cast_err(None, c.dtype == torch.bool, B)
            a, b = _promote(a, b)
            return torch.where(c > 0.5, a, b), err                     # compiler.py:1640
        # switch
        pred_e, cases = e.args
        p, err = self.ev(pred_e, scope, fls, nfls, cur, B)
        objs = self.model.type_to_objects[self._enum_type_of(pred_e, cases)]
        default = dict(cases).get(None)
        vals = []
        for o in objs:
            ce = dict(cases).get(o, default)
            if ce is None:
                raise ValueError(f"switch is missing case <{o}> and has no default")
            v, er = self.ev(ce, scope, fls, nfls, cur, B)
            err = err | er
            vals.append(v)
        vals = torch.broadcast_tensors(*[_at_least_int(v) for v in vals])
        dt = REAL if any(v.dtype == REAL for v in vals) else vals[0].dtype
        stack = torch.stack([v.to(dt) for v in vals], dim=0)
        pb = p.expand(stack.shape[1:]) if p.shape != stack.shape[1:] else p
        if self.relaxed and self.variants.get("switch") == "softmax" \
                and self.tracer.is_fluent(pred_e):                     # logic.py:983-1002
            w = self.param(e, "switch_weight", self.relax["switch_weight"])
            lit = torch.arange(len(objs), dtype=REAL).reshape([-1] + [1] * (stack.dim() - 1))
            logits = -w * torch.abs(_f(pb).unsqueeze(0) - lit)
            soft = torch.sum(_f(stack) * torch.softmax(logits, dim=0), dim=0)
            if self.relax["use_switch_ste"]:
                hard = torch.gather(_f(stack), 0, pb.to(torch.int64).unsqueeze(0))[0]
                soft = _ste(soft, hard)
            return soft, err
        out = torch.gather(stack, 0, pb.to(torch.int64).unsqueeze(0))[0]   # compiler.py:1680-1683
        return out, err

    def _enum_type_of(self, e: Expr, cases=()) -> str:
        if e.etype[0] == "pvar":
            name = e.args[0]
            if name in self.model.variable_ranges:
                return self.model.variable_ranges[name]
        types = {self.model.object_to_type.get(lbl) for lbl, _ in cases if lbl is not None}
        if len(types) == 1 and None not in types:          # a computed predicate: the labels' enum type
            return types.pop()
        raise NotImplementedError("switch predicate must be enum-valued")

    # -- random variables: compiler.py:1701-2241 ; logic.py:1010-1768 --------------------
    def _draw_shape(self, first_arg: Expr, scope):
        return tuple(self._shape(scope)) if self._ref_full(first_arg) else ()

    def _draw(self, cur, kind, first_arg, scope, B, extra=None):
        shp = self._draw_shape(first_arg, scope)
        zz = cur.draw(kind, shp, B, extra)
        return zz if shp else zz.reshape((B,) + (1,) * len(scope))

    def _random(self, e, scope, fls, nfls, cur, B):
        _, name = e.etype
        R = self.relax if self.relaxed else {}
        V = self.variants if self.relaxed else {}

        def oob(cond):   # jnp.logical_not(jnp.all(cond)) per rollout
            c = cond.reshape(cond.shape[0], -1).all(1)
            if c.shape[0] != B:
                c = c.expand(B)
            return (~c).to(torch.int64)

        if name == "KronDelta":
            v, err = self.ev(e.args[0], scope, fls, nfls, cur, B)
            if self.relaxed:
                return v, err                                      # logic.py:232-236
            ok = v.dtype in (torch.bool, INT)
            return v, err | self._cast_err(None, ok, B) * (ERR["KRON"] // ERR["CAST"])
        if name == "DiracDelta":
            v, err = self.ev(e.args[0], scope, fls, nfls, cur, B)
            return _f(v), err                                      # compiler.py:1772-1777
        if name in ("Discrete", "UnnormDiscrete"):
            return self._discrete(e, scope, fls, nfls, cur, B, name == "UnnormDiscrete")
        vals, err = self._args(e, scope, fls, nfls, cur, B)
        a0 = e.args[0]
        if name == "Uniform":                                      # compiler.py:1779-1798
            lb, ub = vals
            U = self._draw(cur, "uniform", a0, scope, B)
            return lb + (ub - lb) * U, err | oob(lb <= ub) * ERR["UNIFORM"]
        if name == "Normal":                                       # compiler.py:1800-1820
            mean, var = vals
            Z = self._draw(cur, "normal", a0, scope, B)
            return mean + safe_sqrt(_f(var)) * Z, err | oob(var >= 0) * ERR["NORMAL"]
        if name == "Exponential":                                  # compiler.py:1822-1839
            (scale,) = vals
            E = self._draw(cur, "exponential", a0, scope, B)
            return scale * E, err | oob(scale > 0) * ERR["EXPONENTIAL"]
        if name == "Weibull":                                      # compiler.py:1841-1860
            shape, scale = vals
            E = self._draw(cur, "exponential", a0, scope, B)
            return scale * torch.pow(E, 1. / _f(shape)), \
                err | oob((shape > 0) & (scale > 0)) * ERR["WEIBULL"]
        def gamma_sampler(shape, scale):
            """standard Gamma(scale * shape) for a fluent shape from supplied draws (the reference
            calls jax.random.gamma, whose bits cannot be reproduced):
            Marsaglia-Tsang on alpha + 1, first accepted of 4 proposals, times U^(1 / alpha)."""
            alpha = _f(shape) if scale == 1.0 else np.float32(scale) * _f(shape)
            d = alpha + np.float32(2.0 / 3.0)
            c = 1.0 / torch.sqrt(9.0 * d)
            props = []
            for _ in range(4):
                z = self._draw(cur, "normal", a0, scope, B)
                u = self._draw(cur, "uniform", a0, scope, B)
                t = 1.0 + c * z
                v = t * t * t
                logv = torch.log(torch.clamp(v, min=1e-30))
                bound = 0.5 * (z * z) + d - d * v + d * logv
                ok = (t > 0) & (torch.log(u) < bound)
                props.append((ok, d * torch.clamp(v, min=1e-30)))
            x = props[-1][1]
            for ok, xv in reversed(props[:-1]):
                x = torch.where(ok, xv, x)
            ub = self._draw(cur, "uniform", a0, scope, B)
            return x * torch.exp(torch.log(ub) / alpha)

        def gamma_draw(shape, shape_e, scale=1.0):
            if not self.tracer.is_static(shape_e):
                return gamma_sampler(shape, scale)
            c0 = shape[0].numpy() * np.float32(scale)
            conc = np.broadcast_to(c0, self._shape(scope)).copy() if self._ref_full(a0) \
                else np.asarray(c0, dtype=np.float32).reshape(-1)[0].reshape(())
            return self._draw(cur, "gamma", a0, scope, B, extra=conc)

        if name == "Gamma":                                        # compiler.py:1899-1919
            shape, scale = vals
            G = gamma_draw(shape, e.args[0])
            return scale * G, err | oob((shape > 0) & (scale > 0)) * ERR["GAMMA"]
        if name == "ChiSquare":                                    # compiler.py:2130-2147
            (df,) = vals
            G = gamma_draw(_f(df), e.args[0], 0.5)
            return 2.0 * G, err | oob(df > 0) * ERR["CHISQUARE"]
        if name == "Student":              # compiler.py:2028-2044; jax.random.t = Z sqrt(h / Gamma(h))
            (df,) = vals
            Z = self._draw(cur, "normal", a0, scope, B)
            G = gamma_draw(_f(df), e.args[0], 0.5)
            return Z * torch.sqrt(0.5 * _f(df) / G), err | oob(df > 0) * ERR["STUDENT"]
        if name == "Beta":                 # compiler.py:1969-1986; Ga / (Ga + Gb)
            a, b = vals
            Ga = gamma_draw(_f(a), e.args[0])
            Gb = gamma_draw(_f(b), e.args[1])
            return Ga / (Ga + Gb), err | oob((a > 0) & (b > 0)) * ERR["BETA"]
        if name == "Pareto":               # compiler.py:2007-2026; jax.random.pareto = exp(Exp(1)/b)
            shape, scale = vals
            E = self._draw(cur, "exponential", a0, scope, B)
            return scale * torch.exp(E / _f(shape)), \
                err | oob((shape > 0) & (scale > 0)) * ERR["PARETO"]
        def poisson_sample(rate, variant, perr):
            if variant == "determinized":                          # logic.py:1733-1741
                return rate, perr
            if variant == "exponential":                           # logic.py:1508-1531
                from scipy import optimize, stats
                K = int(R["poisson_nbins"])
                w = self.param(e, "poisson_comparison_weight", R["poisson_comparison_weight"])
                cut = float(optimize.brentq(
                    lambda r: stats.poisson.cdf(K, r) - float(R["poisson_min_cdf"]),
                    1e-6, 10.0 * K + 10))
                cum, cnt = 0.0, 0.0
                for _ in range(K):
                    cum = cum + self._draw(cur, "exponential", a0, scope, B) / rate
                    cnt = cnt + torch.sigmoid(w * (1.0 - cum))
                Z = self._draw(cur, "normal", a0, scope, B)
                return torch.where(rate.detach() <= cut, cnt, rate + safe_sqrt(rate) * Z), perr
            if variant == "gumbel":                                # logic.py:1618-1641
                from scipy import optimize, stats
                K = int(R["poisson_nbins"])
                w = self.param(e, "poisson_softmax_weight", R["poisson_softmax_weight"])
                eps = self.param(e, "poisson_eps", R["poisson_eps"])
                cut = float(optimize.brentq(
                    lambda r: stats.poisson.cdf(K, r) - float(R["poisson_min_cdf"]),
                    1e-6, 10.0 * K + 10))
                ks = torch.arange(K, dtype=REAL)
                r_ = rate.unsqueeze(-1)
                term = torch.where(ks == 0, torch.zeros(()), ks * safe_log(r_, eps))
                log_prob = term - r_ - torch.lgamma(ks + 1)
                shp = tuple(self._shape(scope)) if self._ref_full(a0) else ()
                g = cur.draw("gumbel", shp + (K,), B)
                if not shp:
                    g = g.reshape((B,) + (1,) * len(scope) + (K,))
                cnt = soft_argmax(g + log_prob, w)
                Z = self._draw(cur, "normal", a0, scope, B)
                return torch.where(rate.detach() <= cut, cnt, rate + safe_sqrt(rate) * Z), perr
            if variant is not None:
                raise NotImplementedError(f"Poisson relaxation <{variant}> is not built")
            # exact: inverse CDF of one uniform, truncated at COUNT_NBINS terms (jax.random.poisson's
            # own bits cannot be reproduced from supplied noise; same law below the truncation)
            U = self._draw(cur, "uniform", a0, scope, B)
            pk = torch.exp(-rate)
            cdf, cnt = pk, torch.zeros(())
            for k in range(1, COUNT_NBINS + 1):
                cnt = cnt + (U > cdf).to(REAL)
                pk = pk * (rate * np.float32(1.0 / k))
                cdf = cdf + pk
            return (cnt if self.relaxed else cnt.to(INT)), perr
        if name == "Poisson":
            (rate,) = vals
            rate = _f(rate)
            perr = err | oob(rate >= 0) * ERR["POISSON"]
            variant = V.get("poisson") if self.tracer.is_fluent(a0) else None
            return poisson_sample(rate, variant, perr)
        if name == "Binomial":
            trials, prob = _f(vals[0]), _f(vals[1])
            perr = err | oob((prob >= 0) & (prob <= 1) & (trials >= 0)) * ERR["BINOMIAL"]
            fluent = self.tracer.is_fluent(e.args[0]) or self.tracer.is_fluent(e.args[1])
            variant = V.get("binomial") if fluent else None
            if variant == "determinized":                          # logic.py:1476
                return trials * prob, perr
            if variant == "gumbel":                                # logic.py:1381-1434
                K = int(R["binomial_nbins"])
                w = self.param(e, "binomial_softmax_weight", R["binomial_softmax_weight"])
                eps = self.param(e, "binomial_eps", R["binomial_eps"])
                ks0 = torch.arange(K, dtype=REAL)
                n_, p_ = trials.unsqueeze(-1), prob.unsqueeze(-1)
                in_support = ks0 <= n_
                ks = torch.minimum(ks0, n_)
                term_k = torch.where(ks == 0, torch.zeros(()), ks * safe_log(p_, eps))
                term_nk = torch.where(ks == n_, torch.zeros(()), (n_ - ks) * safe_log(1. - p_, eps))
                log_prob = (torch.lgamma(n_ + 1) - torch.lgamma(ks + 1) - torch.lgamma(n_ - ks + 1)
                            + term_k + term_nk)
                log_prob = torch.where(in_support, log_prob, torch.full((), -float("inf")))
                shp = tuple(self._shape(scope)) \
                    if (self._ref_full(e.args[0]) or self._ref_full(e.args[1])) else ()
                g = cur.draw("gumbel", shp + (K,), B)
                if not shp:
                    g = g.reshape((B,) + (1,) * len(scope) + (K,))
                cnt = soft_argmax(g + log_prob, w)
                Z = self._draw(cur, "normal", a0, scope, B)
                var = trials * torch.clamp(prob * (1. - prob), 0., 1.)
                normal = trials * prob + safe_sqrt(var) * Z
                return torch.where(trials.detach() < K, cnt, normal), perr
            if variant is not None:
                raise NotImplementedError(f"Binomial relaxation <{variant}> is not built")
            cnt = torch.zeros(())
            for j in range(COUNT_NBINS):     # successes among the first `trials` uniform draws
                U = self._draw(cur, "uniform", a0, scope, B)
                cnt = cnt + ((float(j) < trials) & (U < prob)).to(REAL)
            return (cnt if self.relaxed else cnt.to(INT)), perr
        if name == "NegativeBinomial":
            trials, prob = _f(vals[0]), _f(vals[1])
            perr = err | oob((prob >= 0) & (prob <= 1) & (trials > 0)) * ERR["NEGATIVE_BINOMIAL"]
            fluent = self.tracer.is_fluent(e.args[0]) or self.tracer.is_fluent(e.args[1])
            variant = V.get("poisson") if (self.relaxed and fluent) else None
            if variant == "determinized":
                return (1.0 / prob - 1.0) * trials, perr           # logic.py:1762
            G = gamma_draw(trials, e.args[0])
            if variant is None:        # exact: tfp NegativeBinomial(total_count, probs) as a mixture
                rate = G * prob / (1.0 - prob)
            else:                      # logic.py:1583-1585, 1698-1700
                rate = (1.0 / prob - 1.0) * G
            return poisson_sample(rate, variant, perr)
        if name == "Gumbel":                                       # compiler.py:2046-2065
            mean, scale = vals
            g = self._draw(cur, "gumbel", a0, scope, B)
            return mean + scale * g, err | oob(scale > 0) * ERR["GUMBEL"]
        if name == "Laplace":                                      # compiler.py:2067-2086
            mean, scale = vals
            g = self._draw(cur, "laplace", a0, scope, B)
            return mean + scale * g, err | oob(scale > 0) * ERR["LAPLACE"]
        if name == "Cauchy":                                       # compiler.py:2088-2107
            mean, scale = vals
            g = self._draw(cur, "cauchy", a0, scope, B)
            return mean + scale * g, err | oob(scale > 0) * ERR["CAUCHY"]
        if name == "Gompertz":                                     # compiler.py:2109-2128
            shape, scale = vals
            E = self._draw(cur, "exponential", e.args[1], scope, B)
            return safe_div(safe_div(torch.log(1. + E), shape), scale), \
                err | oob((shape > 0) & (scale > 0)) * ERR["GOMPERTZ"]
        if name == "Kumaraswamy":                                  # compiler.py:2149-2168
            a, b = vals
            U = self._draw(cur, "uniform", a0, scope, B)
            one = torch.ones(())
            return torch.pow(1.0 - torch.pow(U, safe_div(one, b)), safe_div(one, a)), \
                err | oob((a > 0) & (b > 0)) * ERR["KUMARASWAMY"]
        if name == "Bernoulli":
            (p,) = vals
            variant = V.get("bernoulli") if self.tracer.is_fluent(a0) else None
            perr = err | oob((p >= 0) & (p <= 1)) * ERR["BERNOULLI"]
            if variant is None:                                    # compiler.py:1862-1878
                U = self._draw(cur, "uniform", a0, scope, B)
                return U < p, perr
            if variant == "sigmoid":                               # logic.py:1122-1133
                w = self.param(e, "bernoulli_sigmoid_weight", R["bernoulli_sigmoid_weight"])
                U = self._draw(cur, "uniform", a0, scope, B)
                pf = _f(p)
                s = torch.sigmoid(w * (pf - U))
                if R["use_ste_bernoulli"]:
                    s = _ste(s, pf > U)
                return s, perr
            if variant == "gumbel":                                # logic.py:1172-1186
                w = self.param(e, "bernoulli_sigmoid_weight", R["bernoulli_sigmoid_weight"])
                eps = self.param(e, "bernoulli_prob_eps", R["bernoulli_prob_eps"])
                L = self._draw(cur, "logistic", a0, scope, B)
                pc = torch.clamp(_f(p), eps, 1. - eps)
                logit = torch.log(pc / (1. - pc)) + L
                s = torch.sigmoid(w * logit)
                if R["use_ste_bernoulli"]:
                    s = _ste(s, logit > 0)
                return s, perr
            return _f(p), perr                                     # logic.py:1212-1218
        if name == "Geometric":
            (p,) = vals
            variant = V.get("geometric") if self.tracer.is_fluent(a0) else None
            perr = err | oob((p >= 0) & (p <= 1)) * ERR["GEOMETRIC"]
            if variant is None:       # jax.random.geometric: floor(log(U)/log1p(-p)) + 1
                U = self._draw(cur, "uniform", a0, scope, B)
                return (torch.floor(torch.log(U) / torch.log1p(-_f(p))) + 1).to(INT), perr
            if variant == "reparam":                               # logic.py:1042-1051
                w = self.param(e, "geometric_floor_weight", R["geometric_floor_weight"])
                eps = self.param(e, "geometric_eps", R["geometric_eps"])
                E = self._draw(cur, "exponential", a0, scope, B)
                return 1. + soft_floor(safe_div(-E, safe_log(1. - _f(p), eps)), w), perr
            return safe_div(torch.ones(()), _f(p)), perr           # logic.py:1077-1083
        raise NotImplementedError(f"distribution {name} is not supported by the oracle")

    def _discrete(self, e, scope, fls, nfls, cur, B, unnorm):
        etype_name, cases = e.args                                 # compiler.py:2203-2215
        objs = self.model.type_to_objects[etype_name]
        table = dict(cases)
        vals, err = self._args(e, scope, fls, nfls, cur, B, args=[table[o] for o in objs])
        sizes = self._shape(scope)
        vals = [_f(v).expand((v.shape[0],) + tuple(sizes)) for v in vals]
        nb = max(v.shape[0] for v in vals)
        prob = torch.stack([v.expand((nb,) + tuple(sizes)) for v in vals], dim=-1)
        if unnorm:
            prob = safe_div(prob, prob.sum(-1, keepdim=True))
        fluent = any(self.tracer.is_fluent(table[o]) for o in objs)
        bad = ~((prob >= 0).reshape(prob.shape[0], -1).all(1)
                & torch.isclose(prob.sum(-1), torch.ones(())).reshape(prob.shape[0], -1).all(1))
        err = err | (bad.expand(B).to(torch.int64) * ERR["DISCRETE"])
        variant = self.variants.get("discrete") if (self.relaxed and fluent) else None
        if variant == "determinized":                              # logic.py:1321-1327
            lit = torch.arange(len(objs), dtype=REAL)
            return torch.sum(lit * prob, dim=-1), err
        g = cur.draw("gumbel", tuple(sizes) + (len(objs),), B)
        if variant == "gumbel":                                    # logic.py:1258-1267
            w = self.param(e, "discrete_softmax_weight", self.relax["discrete_softmax_weight"])
            eps = self.param(e, "discrete_eps", self.relax["discrete_eps"])
            logits = g + safe_log(prob, eps)
            lit = torch.arange(len(objs), dtype=REAL)
            return torch.sum(lit * torch.softmax(w * logits, dim=-1), dim=-1), err
        return torch.argmax(g + safe_log(prob), dim=-1).to(INT), err   # compiler.py:2212

    # ------------------------------------------------------------------------------
    def step(self, fls, nfls, actions, noise: Optional[List[torch.Tensor]], B: int):
        """One transition, compiler.py:490-547.  Returns (new_fls, reward [B], err [B],
        noise spec)."""
        m = self.model
        fls = {**fls, **actions}
        cur = NoiseCursor(noise)
        err = torch.zeros(B, dtype=torch.int64)
        for name in self.cpf_order:                                # compiler.py:360-368
            params, expr = m.cpfs[name]
            v, er = self.ev(expr, list(params), fls, nfls, cur, B)
            err = err | er
            shape = (B,) + tuple(self._shape(params))
            v = v.expand(shape)
            if self.relaxed:                                       # logic.py:164
                out = _f(v)
                if name in (self.relax.get("cpfs_without_grad") or ()):
                    out = out.detach()
            else:                                                  # compiler.py:258-259, 971-983
                prange = m.variable_ranges[name]
                dt = {"real": REAL, "int": INT, "bool": torch.bool}.get(prange, INT)
                out = v.to(dt)
                if not _can_cast(v.dtype, dt):
                    bad = (out.to(v.dtype) != v).reshape(B, -1).any(1)
                    err = err | bad.to(torch.int64) * ERR["CAST"]
            fls[name] = out.contiguous()
        r, er = self.ev(m.reward, [], fls, nfls, cur, B)           # compiler.py:370-376, 521
        err = err | er
        reward = _f(r).reshape(r.shape[0]).expand(B)
        for s, sp in m.next_state.items():                         # compiler.py:524-525
            fls[s] = fls[sp]
        return fls, reward, err, cur.spec


def _can_cast(src, dst) -> bool:
    order = {torch.bool: 0, INT: 1, torch.int64: 1, REAL: 2}
    return order[src] <= order[dst]


def _promote(a, b):
    if a.dtype == b.dtype:
        return a, b
    dt = torch.promote_types(a.dtype, b.dtype)
    if dt == torch.float64:
        dt = REAL
    return a.to(dt), b.to(dt)


def _prod(x, axes):
    for ax in sorted(axes, reverse=True):
        x = torch.prod(x, dim=ax)
    return x


def _amin(x, axes):
    return torch.amin(x, dim=axes)


def _amax(x, axes):
    return torch.amax(x, dim=axes)


def soft_floor(x, w):
    """logic.py:800-804."""
    r = torch.floor(x + 0.5)
    d = x - r
    return r - 1. + 0.5 * (1. + stable_tanh(w * d) / stable_tanh(torch.tensor(w / 2., dtype=REAL)))


def soft_round(x, w):
    """logic.py:879-882."""
    m = torch.floor(x) + 0.5
    return m + 0.5 * stable_tanh(w * (x - m)) / stable_tanh(torch.tensor(w / 2., dtype=REAL))
