"""Object tracing: how every pvariable reference maps onto the axes of its scope.

Plays the role of pyRDDLGym's ``RDDLObjectsTracer`` as consumed by the reference
(compiler.py:141-144; ``cached_sim_info`` compiler.py:1020-1043, ``cached_is_fluent`` used
throughout logic.py).  Instead of (slices, axis, shape, op_code, op_args) recipes for
numpy, a reference ``f(?a, o3, ?b)`` inside a scope is resolved to ONE flat gather table over
the scope variables it actually uses -- the form both the instruction lowering (GATHER
operand maps) and the CPU oracle consume.

Values are laid out "keepdims": rank = len(scope), size-1 on every scope axis an
expression does not depend on, so elementwise broadcasting reproduces the reference's
``expand_dims`` + ``broadcast_to`` (compiler.py:1073-1075) lazily.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .expr import Expr

Scope = List[Tuple[str, str]]  # [(?var, type), ...]


class RDDLTraceError(Exception):
    pass


@dataclass
class Access:
    kind: str                       # 'freevar' | 'object' | 'var'
    name: str = ""
    axes: Tuple[int, ...] = ()      # scope positions used (sorted)
    keep_shape: Tuple[int, ...] = ()  # keepdims shape over the full scope
    index: Optional[np.ndarray] = None   # flat gather table, shape = sizes of `axes`
    identity: bool = False
    value: int = 0                  # object index (kind == 'object')
    nested: List[Tuple[int, Expr]] = field(default_factory=list)


class Tracer:
    def __init__(self, model, stochastic_is_fluent: bool = False):
        self.model = model
        self.stochastic_is_fluent = stochastic_is_fluent
        self._fluent: Dict[int, bool] = {}

    # ---------------------------------------------------------------------------------
    def scope_sizes(self, scope: Scope) -> Tuple[int, ...]:
        return tuple(len(self.model.type_to_objects[t]) for _, t in scope)

    def is_object(self, name: str) -> bool:
        return name in self.model.object_to_index and name not in self.model.variable_types

    def is_fluent(self, expr: Expr) -> bool:
        """``cached_is_fluent``: does the node (syntactically) depend on any fluent."""
        r = self._fluent.get(expr.id)
        if r is not None:
            return r
        fam, op = expr.etype
        m = self.model
        if fam == "constant":
            r = False
        elif fam == "pvar":
            name, params = expr.args
            if name.startswith("?") or (params is None and self.is_object(name)):
                r = False
            else:
                r = name not in m.non_fluents
                for p in params or ():
                    if isinstance(p, Expr) and self.is_fluent(p):
                        r = True
        elif fam == "aggregation":
            r = self.is_fluent(expr.args[-1])
        elif fam == "control" and op == "switch":
            pred, cases = expr.args
            r = self.is_fluent(pred) or any(self.is_fluent(e) for _, e in cases)
        elif fam == "matrix":
            r = self.is_fluent(expr.args[-1])
        elif fam == "randomvector":
            r = self.stochastic_is_fluent or any(self.is_fluent(a) for a in expr.args[1:])
        elif fam == "randomvar":
            if op in ("Discrete", "UnnormDiscrete"):
                args = [e for _, e in expr.args[1]]
            else:
                args = list(expr.args)
            r = self.stochastic_is_fluent or any(self.is_fluent(a) for a in args)
        else:
            r = any(self.is_fluent(a) for a in expr.args)
        self._fluent[expr.id] = r
        return r

    def is_static(self, expr: Expr) -> bool:
        """True when the value is the same for every rollout and step: no fluent, no draw."""
        from .expr import walk
        m = self.model
        for node in walk(expr):
            fam, op = node.etype
            if (fam == "randomvar" and op not in ("KronDelta", "DiracDelta")) or fam == "randomvector":
                return False
            if fam == "pvar":
                name, params = node.args
                if name.startswith("?") or (params is None and self.is_object(name)):
                    continue
                if name not in m.non_fluents:
                    return False
        return True

    # ---------------------------------------------------------------------------------
    def access(self, expr: Expr, scope: Scope) -> Access:
        """Resolve a pvar node against its scope (``cached_sim_info`` for pvars)."""
        m = self.model
        name, params = expr.args
        svars = [v for v, _ in scope]
        sizes = self.scope_sizes(scope)
        if name.startswith("?"):
            if name not in svars:
                raise RDDLTraceError(f"free variable {name} is not in scope {svars}")
            ax = svars.index(name)
            keep = tuple(sizes[i] if i == ax else 1 for i in range(len(scope)))
            return Access("freevar", name, (ax,), keep)
        if params is None and self.is_object(name):
            return Access("object", name, value=m.object_to_index[name])
        if name not in m.variable_types:
            raise RDDLTraceError(f"unknown pvariable <{name}>")
        params = params or []
        vshape = m.variable_shape(name)
        if len(params) != len(vshape):
            raise RDDLTraceError(f"<{name}> expects {len(vshape)} parameters, got {len(params)}")
        # classify every parameter
        axes_used: List[int] = []
        terms = []   # per var-dimension: ('axis', scope_pos) | ('const', idx) | ('nested', Expr)
        nested = []
        for d, p in enumerate(params):
            if isinstance(p, Expr):
                # a nested zero-arity reference that is really an object literal / free var
                if p.etype[0] == "pvar" and p.args[1] is None and \
                        (p.args[0].startswith("?") or self.is_object(p.args[0])):
                    p = p.args[0]
                else:
                    nested.append((d, p))
                    terms.append(("nested", p))
                    continue
            if p.startswith("?"):
                if p not in svars:
                    raise RDDLTraceError(f"free variable {p} is not in scope {svars}")
                ax = svars.index(p)
                axes_used.append(ax)
                terms.append(("axis", ax))
            elif self.is_object(p):
                terms.append(("const", m.object_to_index[p]))
            elif p in m.variable_types and not m.variable_params[p]:
                e = Expr(("pvar", p), (p, None))
                nested.append((d, e))
                terms.append(("nested", e))
            else:
                raise RDDLTraceError(f"cannot resolve argument {p!r} of <{name}>")
        axes = tuple(sorted(set(axes_used)))
        sub_sizes = tuple(sizes[a] for a in axes)
        strides = np.ones(len(vshape), dtype=np.int64)
        for d in range(len(vshape) - 2, -1, -1):
            strides[d] = strides[d + 1] * vshape[d + 1]
        grids = np.indices(sub_sizes, dtype=np.int64) if axes else np.zeros((0,), np.int64)
        index = np.zeros(sub_sizes, dtype=np.int64)
        for d, (kind, v) in enumerate(terms):
            if kind == "axis":
                index = index + grids[axes.index(v)] * strides[d]
            elif kind == "const":
                index = index + v * strides[d]
        keep = tuple(sizes[i] if i in axes else 1 for i in range(len(scope)))
        total = int(np.prod(vshape, dtype=np.int64))
        identity = (not nested) and index.size == total and \
            bool(np.array_equal(index.ravel(), np.arange(total)))
        acc = Access("var", name, axes, keep, index, identity, nested=nested)
        acc.strides = strides  # type: ignore[attr-defined]
        return acc


# -------------------------------------------------------------------------------------
# static (non-fluent, deterministic) evaluation in numpy, exact semantics of compiler.py
# -------------------------------------------------------------------------------------

def _fix(x):
    x = np.asarray(x)
    if x.dtype == np.float64:
        return x.astype(np.float32)
    if x.dtype == np.int64:
        return x.astype(np.int32)
    return x


def _at_least_int(x):
    x = np.asarray(x)
    return x.astype(np.int32) if x.dtype == np.bool_ else x


_UNARY = {
    "abs": np.abs, "sgn": np.sign, "round": np.round, "floor": np.floor, "ceil": np.ceil,
    "cos": np.cos, "sin": np.sin, "tan": np.tan, "acos": np.arccos, "asin": np.arcsin,
    "atan": np.arctan, "cosh": np.cosh, "sinh": np.sinh, "tanh": np.tanh, "exp": np.exp,
    "ln": np.log, "sqrt": np.sqrt,
}
_BINARY_F = {
    "div": np.floor_divide, "mod": np.mod, "fmod": np.fmod, "min": np.minimum,
    "max": np.maximum, "pow": np.power, "hypot": np.hypot,
}
_REL = {">=": np.greater_equal, "<=": np.less_equal, ">": np.greater, "<": np.less,
        "==": np.equal, "~=": np.not_equal}
_AGG = {"sum": np.sum, "avg": np.mean, "prod": np.prod, "minimum": np.min, "maximum": np.max,
        "forall": np.all, "exists": np.any, "argmin": np.argmin, "argmax": np.argmax}


def eval_static(model, expr: Expr, scope: Scope, as_float: bool = False,
                tracer: Optional[Tracer] = None, cast_errors: Optional[list] = None) -> np.ndarray:
    """Evaluate a non-fluent, deterministic expression with the exact semantics of
    compiler.py:1000-1685.  Result is keepdims-shaped over ``scope``.

    ``as_float=True`` reads non-fluents cast to REAL as the differentiable compiler does
    (logic.py:144-147).
    """
    tr = tracer or Tracer(model)

    def ev(e: Expr, sc: Scope):
        fam, op = e.etype
        if fam == "constant":
            if op == "bool":
                return np.asarray(bool(e.args))
            if op == "int":
                return np.asarray(e.args, dtype=np.int32)
            return np.asarray(e.args, dtype=np.float32)
        if fam == "pvar":
            acc = tr.access(e, sc)
            if acc.kind == "freevar":
                return np.arange(acc.keep_shape[acc.axes[0]], dtype=np.int32).reshape(acc.keep_shape)
            if acc.kind == "object":
                return np.asarray(acc.value, dtype=np.int32)
            if acc.name not in model.non_fluents:
                raise RDDLTraceError(f"<{acc.name}> is a fluent; expression is not static")
            if acc.nested:
                raise RDDLTraceError("nested static references are not supported")
            val = np.asarray(model.init_values[acc.name])
            if as_float:
                val = val.astype(np.float32)
            return val.reshape(-1)[acc.index].reshape(acc.keep_shape if sc else ())
        if fam == "arithmetic":
            args = [_at_least_int(ev(a, sc)) for a in e.args]
            if op == "-" and len(args) == 1:
                return _fix(np.negative(args[0]))
            fn = {"+": np.add, "-": np.subtract, "*": np.multiply, "/": np.true_divide}[op]
            out = args[0]
            for a in args[1:]:
                with np.errstate(all="ignore"):
                    out = _fix(fn(out, a))
            return out
        if fam == "relational":
            a, b = (_at_least_int(ev(x, sc)) for x in e.args)
            return _REL[op](a, b)
        if fam == "boolean":
            args = [ev(a, sc) for a in e.args]
            if cast_errors is not None and any(a.dtype != np.bool_ for a in args):
                cast_errors.append(e.id)     # check_dtype=bool, compiler.py:1112-1117
            if op == "~" and len(args) == 1:
                return np.logical_not(args[0])
            if op == "~":
                return np.logical_xor(args[0], args[1])
            if op in ("^", "&"):
                out = args[0]
                for a in args[1:]:
                    out = np.logical_and(out, a)
                return out
            if op == "|":
                out = args[0]
                for a in args[1:]:
                    out = np.logical_or(out, a)
                return out
            if op == "=>":
                return np.logical_or(np.logical_not(args[0]), args[1])
            if op == "<=>":
                return np.equal(args[0], args[1])
        if fam == "aggregation":
            *tv, body = e.args
            inner = sc + list(tv)
            val = ev(body, inner)
            sizes = tr.scope_sizes(inner)
            is_bool = op in ("forall", "exists")
            if is_bool and cast_errors is not None and val.dtype != np.bool_:
                cast_errors.append(e.id)
            if not is_bool:
                val = _at_least_int(val)
            val = np.broadcast_to(val, sizes)
            axes = tuple(range(len(sc), len(inner)))
            if op in ("argmin", "argmax"):
                if len(axes) != 1:
                    raise RDDLTraceError("argmin/argmax take exactly one variable")
                out = _AGG[op](val, axis=axes[0])
                return _fix(out).reshape(sizes[:len(sc)])
            out = _fix(_AGG[op](val, axis=axes))
            # keep rank == len(sc); collapse axes the result does not depend on is not
            # needed: result is full-size over the outer scope
            return out
        if fam == "func":
            args = [_at_least_int(ev(a, sc)) for a in e.args]
            with np.errstate(all="ignore"):
                if op in _UNARY:
                    return _fix(_UNARY[op](args[0]))
                if op == "log":
                    return _fix(np.log(args[0]) / np.log(args[1]))
                if op in _BINARY_F:
                    return _fix(_BINARY_F[op](args[0], args[1]))
                if op in ("lngamma", "gamma"):
                    from scipy import special
                    f = special.gammaln if op == "lngamma" else special.gamma
                    return _fix(f(args[0].astype(np.float32)))
        if fam == "control" and op == "if":
            c, a, b = (ev(x, sc) for x in e.args)
            if cast_errors is not None and c.dtype != np.bool_:
                cast_errors.append(e.id)     # compiler.py:1642-1643
            return _fix(np.where(c > 0.5, a, b))
        if fam == "randomvar" and op in ("KronDelta", "DiracDelta"):
            v = ev(e.args[0], sc)
            return v.astype(np.float32) if op == "DiracDelta" else v
        raise RDDLTraceError(f"expression {e.etype} is not supported in static evaluation")

    return ev(expr, list(scope))
