"""Lowering of CPF / reward expression trees to the SSA step graph.

One ``StepGraph`` restates what ``JaxRDDLCompiler._jax_cpfs`` + ``_jax_reward`` +
``compile_transition`` build as nested closures (compiler.py:360-376, 490-547), for either
the exact model (compiler.py:936-2241) or a relaxed model (logic.py:83-1789, selected by
the ``relax`` config of core/logic.py).  Differences from the reference are layout only:

* values are narrowed to the object tuples they depend on ("tuple-space narrowing") and
  broadcast lazily through operand gather maps;
* everything that is non-fluent and deterministic is folded on the host
  (rddl/trace.py:eval_static) into constant tables;
* every relaxation weight becomes a run-time *model parameter* slot, the counterpart of
  ``aux['params'][expr.id]`` (logic.py:266), so weights can be changed per ``optimize`` call
  (planner.py:3165-3168) without rebuilding the program.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..rddl.expr import Expr
from ..rddl.trace import Tracer, eval_static, RDDLTraceError
from .graph import Graph, V, Map, norm_map

ERR_BITS = {  # compiler.py:861-888
    "CAST": 0, "UNIFORM": 1, "NORMAL": 2, "EXPONENTIAL": 3, "WEIBULL": 4, "BERNOULLI": 5,
    "POISSON": 6, "GAMMA": 7, "BETA": 8, "GEOMETRIC": 9, "PARETO": 10, "STUDENT": 11,
    "GUMBEL": 12, "LAPLACE": 13, "CAUCHY": 14, "GOMPERTZ": 15, "CHISQUARE": 16,
    "KUMARASWAMY": 17, "DISCRETE": 18, "KRON": 19, "DIRICHLET": 20, "MULTIVARIATE_STUDENT": 21,
    "MULTINOMIAL": 22, "BINOMIAL": 23, "NEGATIVE_BINOMIAL": 24,
}

# exact Poisson / Binomial samplers are truncated at this many terms (the reference's relaxations
# default to the same number of bins, logic.py:1362, 1492)
COUNT_NBINS = 100


class RDDLNotImplementedError(NotImplementedError):
    pass


@dataclass
class TV:
    """A lowered tensor value: SSA value + the scope positions labelling its axes.

    ``idx`` (shape = sizes of ``vars``) maps each tuple to an element of ``v``; None means
    the identity (v.n == number of tuples).
    """
    v: V
    vars: Tuple[int, ...]
    idx: Optional[np.ndarray] = None
    full: bool = True        # reference value carries the full scope shape (vs 0-d constant)

    @property
    def dtype(self):
        return self.v.dtype


@dataclass
class NoiseSlot:
    kind: str
    shape: Tuple[int, ...]
    n: int
    value: V
    extra: Any = None
    offset: int = 0


@dataclass
class StepGraph:
    g: Graph
    state_in: Dict[str, V] = field(default_factory=dict)
    action_in: Dict[str, V] = field(default_factory=dict)
    next_state: Dict[str, V] = field(default_factory=dict)
    fluents: Dict[str, V] = field(default_factory=dict)       # every fluent value of the step
    reward: Optional[V] = None
    noise: List[NoiseSlot] = field(default_factory=list)
    mparams: List[Tuple[Tuple[int, str], float, V]] = field(default_factory=list)
    errs: List[Tuple[int, V]] = field(default_factory=list)
    static_err: int = 0
    relaxed: bool = False


class Lowerer:
    def __init__(self, model, relax: Optional[Dict[str, Any]] = None):
        self.m = model
        self.relax = relax
        self.relaxed = relax is not None
        self.variants = (relax or {}).get("variants", {})
        self.tr = Tracer(model)
        self.g = Graph()
        self.sg = StepGraph(self.g, relaxed=self.relaxed)
        self._mparam_index: Dict[Tuple[int, str], V] = {}
        self.mparam_ops: Dict[Tuple[int, str], str] = {}       # key -> 'family op' of its expression
        # bookkeeping for overriden_ops_info / exact_ops_info (compiler.py:816-832): expression id ->
        # (relaxation family, 'family op') of the relaxable operations that were relaxed / kept exact
        self.relaxed_ops: Dict[int, Tuple[str, str]] = {}
        self.exact_ops: Dict[int, Tuple[str, str]] = {}

    # ---------------------------------------------------------------------------------
    def dtype_of_range(self, prange: str) -> str:
        if self.relaxed:
            return "f"
        return {"real": "f", "int": "i", "bool": "b"}.get(prange, "i")

    def mparam(self, expr: Expr, name: str, default: float) -> V:
        key = (expr.id, name)
        v = self._mparam_index.get(key)
        if v is None:
            v = self.g.leaf("mparam", 1, "f", uniform=True, key=key, default=float(default),
                            slot=len(self.sg.mparams))
            self.sg.mparams.append((key, float(default), v))
            self._mparam_index[key] = v
            self.mparam_ops[key] = " ".join(str(x) for x in expr.etype)
        return v

    def sizes(self, scope) -> Tuple[int, ...]:
        return self.tr.scope_sizes(scope)

    # ---- tuple-space helpers -----------------------------------------------------------
    def n_of(self, vars_, sizes) -> int:
        return int(np.prod([sizes[p] for p in vars_], dtype=np.int64)) if vars_ else 1

    def operand(self, tv: TV, rvars: Tuple[int, ...], sizes) -> Tuple[V, Map]:
        """Gather map that presents ``tv`` over the tuple space ``rvars``."""
        rshape = tuple(sizes[p] for p in rvars)
        n_r = int(np.prod(rshape, dtype=np.int64)) if rshape else 1
        idx = tv.idx
        if idx is None:
            idx = np.arange(tv.v.n, dtype=np.int64).reshape(tuple(sizes[p] for p in tv.vars))
        # insert size-1 axes for variables of rvars that tv does not depend on
        shape = tuple(sizes[p] if p in tv.vars else 1 for p in rvars)
        table = np.broadcast_to(idx.reshape(shape), rshape).reshape(-1) if rshape \
            else idx.reshape(-1)
        if table.size != n_r:
            table = np.broadcast_to(table, (n_r,))
        return tv.v, norm_map(table, tv.v.n)

    def union(self, tvs: Sequence[TV]) -> Tuple[int, ...]:
        s = set()
        for t in tvs:
            s.update(t.vars)
        return tuple(sorted(s))

    def ew(self, op: str, tvs: Sequence[TV], sizes, dtype: Optional[str] = None, **attrs) -> TV:
        rvars = self.union(tvs)
        n = self.n_of(rvars, sizes)
        args = [self.operand(t, rvars, sizes) for t in tvs]
        v = self.g.op(op, args, n, dtype, **attrs)
        return TV(v, rvars, None, any(t.full for t in tvs))

    def const_tv(self, value, dtype=None, full=False) -> TV:
        return TV(self.g.const(value, dtype), (), None, full)

    def materialise(self, tv: TV, sizes) -> V:
        """An SSA value holding exactly tv's tuples in order (copy only when needed)."""
        if tv.idx is None:
            return tv.v
        v, m = self.operand(tv, tv.vars, sizes)
        if m is None:
            return v
        return self.g.op("MOV", [(v, m)], self.n_of(tv.vars, sizes), dtype=tv.dtype)

    # ---- dtype discipline (compiler.py:985-994, 1092-1093) -----------------------------
    def to_f(self, tv: TV, sizes) -> TV:
        if tv.dtype == "f":
            return tv
        return self.ew("I2F", [tv], sizes, "f")

    def at_least_int(self, tv: TV) -> TV:
        if tv.dtype == "b":
            return TV(self._retag(tv.v, "i"), tv.vars, tv.idx, tv.full)
        return tv

    def _retag(self, v: V, dtype: str) -> V:
        """bool and int share the 0/1 int32 representation: retagging is free."""
        if v.dtype == dtype:
            return v
        if v.const is not None:
            return self.g.const(v.const, dtype)
        r = self.g.add(V("sg", [(v, None)], n=v.n, dtype=dtype, uniform=v.uniform))
        r.attrs["retag"] = True
        return r

    def to_b(self, tv: TV, sizes) -> TV:
        if tv.dtype == "b":
            return tv
        return self.ew("F2B" if tv.dtype == "f" else "I2B", [tv], sizes, "b")

    def promote(self, tvs: Sequence[TV], sizes) -> Tuple[List[TV], str]:
        tvs = [self.at_least_int(t) for t in tvs]
        if any(t.dtype == "f" for t in tvs):
            return [self.to_f(t, sizes) for t in tvs], "f"
        return list(tvs), "i"

    def ste(self, soft: TV, hard: TV, sizes) -> TV:
        """soft + stop_gradient(hard - soft)  (logic.py:272 and friends)."""
        hard = self.to_f(self.at_least_int(hard), sizes)
        diff = self.ew("SUB", [hard, soft], sizes)
        blocked = TV(self.g.sg(diff.v), diff.vars, None, diff.full)
        return self.ew("ADD", [soft, blocked], sizes)

    def static_cast_err(self, ok: bool):
        if not ok:
            self.sg.static_err |= 1 << ERR_BITS["CAST"]

    def errchk(self, name: str, violation: TV, sizes):
        """Record ``err |= bit`` where ``violation`` (bool) holds for any element."""
        if violation.v.const is not None:
            if np.any(violation.v.const != 0):
                self.sg.static_err |= 1 << ERR_BITS[name]
            return
        self.sg.errs.append((ERR_BITS[name], self.materialise(violation, sizes)))

    # ==================================================================================
    _RELAX_FAMILY = {
        "relational": lambda op: "relational",
        "boolean": lambda op: "logic",
        "aggregation": lambda op: {"forall": "logic", "exists": "logic", "argmin": "argmax",
                                   "argmax": "argmax"}.get(op),
        "func": lambda op: {"sgn": "sgn", "floor": "floor", "ceil": "floor", "div": "floor",
                            "mod": "floor", "round": "round"}.get(op),
        "control": lambda op: {"if": "if", "switch": "switch"}.get(op),
        "randomvar": lambda op: {"Bernoulli": "bernoulli", "Geometric": "geometric",
                                 "Poisson": "poisson", "NegativeBinomial": "poisson",
                                 "Binomial": "binomial", "Discrete": "discrete",
                                 "UnnormDiscrete": "discrete"}.get(op),
    }

    def _note_relaxable(self, e: Expr) -> None:
        """Record whether a relaxable operation is relaxed (its deciding argument is fluent and the
        compiler has a relaxation for its family) or evaluated exactly.  Bookkeeping only."""
        try:
            fam, op = e.etype
            key = self._RELAX_FAMILY.get(fam, lambda _: None)(op)
            if key is None or e.id in self.relaxed_ops or e.id in self.exact_ops:
                return
            if fam == "control" or (fam == "randomvar" and op in ("Bernoulli", "Geometric", "Poisson")):
                fluent = self.tr.is_fluent(e.args[0])
            elif fam == "randomvar" and op in ("NegativeBinomial", "Binomial"):
                fluent = self.tr.is_fluent(e.args[0]) or self.tr.is_fluent(e.args[1])
            else:
                fluent = self.tr.is_fluent(e)
            name = f"{fam} {op}"
            if fluent and key in self.variants:
                self.relaxed_ops[e.id] = (key, name)
            else:
                self.exact_ops[e.id] = (key, name)
        except Exception:           # never let the bookkeeping stop a lowering
            pass

    def lower(self, e: Expr, scope, env: Dict[str, V]) -> TV:
        """Dispatch of compiler.py:936-964."""
        fam, op = e.etype
        sizes = self.sizes(scope)
        if self.relaxed:
            self._note_relaxable(e)
        if fam != "constant" and self.tr.is_static(e):
            cast_errors: list = []
            try:
                val = eval_static(self.m, e, scope, as_float=self.relaxed, tracer=self.tr,
                                  cast_errors=cast_errors)
            except RDDLTraceError:
                val = None
            if val is not None and cast_errors:
                self.sg.static_err |= 1 << ERR_BITS["CAST"]
            if val is not None:
                return self._static_tv(e, val, scope, sizes)
        if fam == "constant":
            dt = {"bool": "b", "int": "i", "real": "f"}[op]
            return self.const_tv(np.asarray(e.args), dt, full=False)
        if fam == "pvar":
            return self._pvar(e, scope, env, sizes)
        if fam == "arithmetic":
            return self._arithmetic(e, scope, env, sizes)
        if fam == "relational":
            return self._relational(e, scope, env, sizes)
        if fam == "boolean":
            return self._logical(e, scope, env, sizes)
        if fam == "aggregation":
            return self._aggregation(e, scope, env, sizes)
        if fam == "func":
            return self._function(e, scope, env, sizes)
        if fam == "control":
            return self._control(e, scope, env, sizes)
        if fam == "randomvar":
            return self._random(e, scope, env, sizes)
        if fam == "matrix":
            return self._matrix(e, scope, env, sizes)
        if fam == "randomvector":
            return self._random_vector(e, scope, env, sizes)
        raise RDDLNotImplementedError(f"Expression type {e.etype} is not supported.")

    def _static_tv(self, e, val: np.ndarray, scope, sizes) -> TV:
        val = np.asarray(val)
        full = self._ref_full(e)
        if val.ndim == 0:
            return self.const_tv(val, None, full)
        # keepdims array over the scope -> narrow to the axes it really varies over
        val = np.broadcast_to(val, val.shape) if val.ndim == len(scope) else val
        vars_ = tuple(i for i in range(len(scope)) if val.shape[i] > 1)
        flat = val.reshape(tuple(val.shape[i] for i in vars_) if vars_ else ())
        return TV(self.g.const(flat), vars_, None, full)

    def _ref_full(self, e: Expr) -> bool:
        fam, op = e.etype
        if fam == "constant":
            return False
        if fam == "pvar":
            name, params = e.args
            return not (params is None and self.tr.is_object(name))
        if fam in ("aggregation", "matrix", "randomvector") or (fam == "control" and op == "switch") \
                or (fam == "randomvar" and op in ("Discrete", "UnnormDiscrete")):
            return True
        return any(self._ref_full(a) for a in e.args)

    # ---- leaves: compiler.py:1017-1081 ------------------------------------------------
    def _pvar(self, e, scope, env, sizes) -> TV:
        acc = self.tr.access(e, scope)
        if acc.kind == "freevar":
            n = sizes[acc.axes[0]]
            return TV(self.g.const(np.arange(n, dtype=np.int32), "i"), acc.axes, None, True)
        if acc.kind == "object":
            return self.const_tv(np.int32(acc.value), "i", full=False)
        if acc.name not in env and acc.nested and acc.name in self.m.non_fluents:
            # a non-fluent indexed by a fluent: its table becomes a constant of the program
            vals = np.asarray(self.m.non_fluents[acc.name]).reshape(-1)
            prange = self.m.variable_ranges[acc.name]
            dt = "f" if (self.relaxed or prange == "real") else ("b" if prange == "bool" else "i")
            env = dict(env)
            env[acc.name] = self.g.const(vals.astype(np.float32 if dt == "f" else np.int32), dt)
        if acc.name not in env:
            raise RDDLTraceError(f"fluent <{acc.name}> is read before it is defined")
        if acc.nested:
            return self._nested_pvar(acc, scope, env, sizes)
        src = env[acc.name]
        idx = None if acc.identity else acc.index
        return TV(src, acc.axes, idx, True)

    def _nested_pvar(self, acc, scope, env, sizes) -> TV:
        """A pvariable whose arguments are themselves computed (compiler.py:1046-1064: gather by
        integer indices, no gradient w.r.t. the indices).  The index expressions range over the
        objects of their parameter type, so the gather is a select chain over the statically
        sliced candidates -- every candidate is a plain static gather of the source."""
        import itertools
        src = env[acc.name]
        vshape = self.m.variable_shape(acc.name)
        subs = []
        for d, sub in acc.nested:
            tv = self.lower(sub, scope, env)
            if tv.dtype == "f":                    # jnp.asarray(index, dtype=int): truncation
                tv = self.ew("F2I", [tv], sizes, "i")
            subs.append((d, self.at_least_int(tv)))
        ranges = [range(int(vshape[d])) for d, _ in subs]
        if int(np.prod([len(r) for r in ranges], dtype=np.int64)) > 64:
            raise RDDLNotImplementedError(
                f"nested arguments of <{acc.name}> range over more than 64 combinations")
        out = None
        for combo in reversed(list(itertools.product(*ranges))):
            off = sum(int(k) * int(acc.strides[d]) for k, (d, _) in zip(combo, subs))
            cand = TV(src, acc.axes, np.asarray(acc.index) + off, True)
            if out is None:             # out-of-range indices fall on the last candidate
                out = cand
                continue
            hit = None
            for k, (_, ti) in zip(combo, subs):
                h = self.ew("IEQ", [ti, self.const_tv(np.int32(k), "i")], sizes)
                hit = h if hit is None else self.ew("AND", [hit, h], sizes)
            out = self.ew("SELECT", [hit, cand, out], sizes, dtype=src.dtype)
        return out

    def _args(self, e, scope, env, args=None) -> List[TV]:
        return [self.lower(a, scope, env) for a in (e.args if args is None else args)]

    # ---- arithmetic: compiler.py:1150-1183 --------------------------------------------
    def _arithmetic(self, e, scope, env, sizes) -> TV:
        _, op = e.etype
        tvs = self._args(e, scope, env)
        if op == "-" and len(tvs) == 1:
            (x,), dt = self.promote(tvs, sizes)
            return self.ew("NEG" if dt == "f" else "INEG", [x], sizes)
        if op == "/":
            tvs = [self.to_f(self.at_least_int(t), sizes) for t in tvs]
            return self.ew("DIV", tvs, sizes)
        out = tvs[0]
        for t in tvs[1:]:
            (a, b), dt = self.promote([out, t], sizes)
            name = {"+": "ADD", "-": "SUB", "*": "MUL"}[op]
            out = self.ew(name if dt == "f" else "I" + name, [a, b], sizes)
        return out

    # ---- relational: compiler.py:1189-1227 ; logic.py:262-344 ---------------------------
    def _relational(self, e, scope, env, sizes) -> TV:
        _, op = e.etype
        (x, y), dt = self.promote(self._args(e, scope, env), sizes)
        name = {">=": "GE", "<=": "LE", ">": "GT", "<": "LT", "==": "EQ", "~=": "NE"}[op]
        hard = lambda a, b: self.ew(name if a.dtype == "f" else "I" + name, [a, b], sizes)  # noqa
        if not (self.relaxed and self.variants.get("relational") == "sigmoid"
                and self.tr.is_fluent(e)):
            return hard(x, y)
        w = TV(self.mparam(e, "sigmoid_weight", self.relax["sigmoid_weight"]), ())
        x, y = self.to_f(x, sizes), self.to_f(y, sizes)
        if op in (">", ">="):
            soft = self.ew("SIGMOID", [self.ew("MUL", [w, self.ew("SUB", [x, y], sizes)], sizes)],
                           sizes)
            use_ste = self.relax["use_sigmoid_ste"]
        elif op in ("<", "<="):
            soft = self.ew("SIGMOID", [self.ew("MUL", [w, self.ew("SUB", [y, x], sizes)], sizes)],
                           sizes)
            use_ste = self.relax["use_sigmoid_ste"]
        else:
            t = self.ew("STANH", [self.ew("MUL", [w, self.ew("SUB", [y, x], sizes)], sizes)], sizes)
            sq = self.ew("SQUARE", [t], sizes)
            soft = self.ew("SUB", [self.const_tv(np.float32(1.)), sq], sizes) if op == "==" else sq
            use_ste = self.relax["use_tanh_ste"]
        return self.ste(soft, hard(x, y), sizes) if use_ste else soft

    # ---- logical: compiler.py:1233-1274 ; logic.py:424-778 -----------------------------
    def _logical(self, e, scope, env, sizes) -> TV:
        _, op = e.etype
        tvs = self._args(e, scope, env)
        variant = self.variants.get("logic") if self.relaxed else None
        if variant is None or not self.tr.is_fluent(e):
            self.static_cast_err(all(t.dtype == "b" for t in tvs))
            b = [self.to_b(t, sizes) for t in tvs]
            if op == "~" and len(b) == 1:
                return self.ew("NOT", b, sizes)
            if op == "~":
                return self.ew("XOR", b, sizes)
            if op in ("^", "&", "|"):
                out = b[0]
                for t in b[1:]:
                    out = self.ew("AND" if op != "|" else "OR", [out, t], sizes)
                return out
            if op == "=>":
                return self.ew("OR", [self.ew("NOT", [b[0]], sizes), b[1]], sizes)
            if op == "<=>":
                (p, q), dt = self.promote(tvs, sizes)
                return self.ew("EQ" if dt == "f" else "IEQ", [p, q], sizes)
            raise RDDLNotImplementedError(op)
        use_ste = self.relax["use_logic_ste"]
        one = self.const_tv(np.float32(1.))
        half = self.const_tv(np.float32(0.5))
        E = lambda name, *a: self.ew(name, list(a), sizes)     # noqa: E731

        def hard2(kind, x, y):
            gx, gy = E("GT", x, half), E("GT", y, half)
            if kind == "and":
                return E("AND", gx, gy)
            if kind == "or":
                return E("OR", gx, gy)
            if kind == "xor":
                return E("XOR", gx, gy)
            lx, ly = E("LE", x, half), E("LE", y, half)
            if kind == "implies":
                return E("OR", lx, gy)
            return E("AND", E("OR", lx, gy), E("OR", ly, gx))

        def soft2(kind, x, y):
            if variant == "product":
                if kind == "and":
                    return E("MUL", x, y)
                if kind == "or":
                    return E("SUB", one, E("MUL", E("SUB", one, x), E("SUB", one, y)))
                if kind == "xor":
                    o = E("SUB", one, E("MUL", E("SUB", one, x), E("SUB", one, y)))
                    return E("MUL", o, E("SUB", one, E("MUL", x, y)))
                if kind == "implies":
                    return E("SUB", one, E("MUL", x, E("SUB", one, y)))
                a = E("SUB", one, E("MUL", x, E("SUB", one, y)))
                b = E("SUB", one, E("MUL", y, E("SUB", one, x)))
                return E("MUL", a, b)
            if variant == "godel":
                if kind == "and":
                    return E("MIN", x, y)
                if kind == "or":
                    return E("MAX", x, y)
                if kind == "xor":
                    return E("MIN", E("MAX", x, y), E("SUB", one, E("MIN", x, y)))
                if kind == "implies":
                    return E("MAX", E("SUB", one, x), y)
                return E("MIN", E("MAX", E("SUB", one, x), y), E("MAX", E("SUB", one, y), x))
            if kind == "and":
                return E("RELU", E("SUB", E("ADD", x, y), one))
            if kind == "or":
                return E("SUB", one, E("RELU", E("SUB", E("SUB", one, x), y)))
            if kind == "xor":
                return E("RELU", E("SUB", one, E("ABS", E("SUB", E("SUB", one, x), y))))
            if kind == "implies":
                return E("SUB", one, E("RELU", E("SUB", x, y)))
            return E("RELU", E("SUB", one, E("ABS", E("SUB", x, y))))

        def binop(kind, x, y):
            x = self.to_f(self.at_least_int(x), sizes)
            y = self.to_f(self.at_least_int(y), sizes)
            s = soft2(kind, x, y)
            return self.ste(s, hard2(kind, x, y), sizes) if use_ste else s

        if op == "~" and len(tvs) == 1:
            x = self.to_f(self.at_least_int(tvs[0]), sizes)
            s = E("SUB", one, x)
            return self.ste(s, E("LE", x, half), sizes) if use_ste else s
        if op == "~":
            return binop("xor", tvs[0], tvs[1])
        if op in ("^", "&", "|"):
            out = tvs[0]
            for t in tvs[1:]:
                out = binop("and" if op != "|" else "or", out, t)
            return out
        if op == "=>":
            return binop("implies", tvs[0], tvs[1])
        if op == "<=>":
            return binop("equiv", tvs[0], tvs[1])
        raise RDDLNotImplementedError(op)

    # ---- aggregation: compiler.py:1280-1348 ; logic.py:377-417, 516-540, ... ------------
    def _csr(self, body: TV, n_outer: int, inner_sizes) -> Tuple[Tuple[int, ...], np.ndarray, np.ndarray, int]:
        """CSR description of reducing ``body`` over scope positions >= n_outer."""
        dvars = tuple(p for p in body.vars if p < n_outer)
        rvars = tuple(range(n_outer, len(inner_sizes)))
        allv = dvars + rvars
        v, m = self.operand(body, allv, inner_sizes)
        n_d = self.n_of(dvars, inner_sizes)
        K = self.n_of(rvars, inner_sizes)
        from .graph import apply_map
        table = apply_map(m, n_d * K)
        ptr = np.arange(n_d + 1, dtype=np.int64) * K
        return dvars, ptr, table, K

    def _aggregation(self, e, scope, env, sizes) -> TV:
        _, op = e.etype
        *tv, body_e = e.args
        inner = list(scope) + list(tv)
        isz = self.sizes(inner)
        body = self.lower(body_e, inner, env)
        n_outer = len(scope)
        fluent = self.tr.is_fluent(e)
        dvars, ptr, idx, K = self._csr(body, n_outer, isz)
        red = lambda name, src, dt=None: TV(self.g.reduce(name, src, ptr, idx, dt), dvars, None)  # noqa
        src = body.v
        is_bool = op in ("forall", "exists")
        if is_bool:
            variant = self.variants.get("logic") if self.relaxed else None
            if variant is None or not fluent:
                self.static_cast_err(body.dtype == "b")
                b = self.to_b(body, isz)
                dvars, ptr, idx, K = self._csr(b, n_outer, isz)
                out = TV(self.g.reduce("RIMIN" if op == "forall" else "RIMAX", b.v, ptr, idx, "b"),
                         dvars, None)
                return out
            x = self.to_f(self.at_least_int(body), isz)
            dvars, ptr, idx, K = self._csr(x, n_outer, isz)
            one = self.const_tv(np.float32(1.))
            R = lambda name, t: TV(self.g.reduce(name, t.v, *self._csr(t, n_outer, isz)[1:3]),  # noqa
                                   dvars, None)
            if variant == "product":
                if op == "forall":
                    soft = R("RPROD", x)
                else:
                    soft = self.ew("SUB", [one, R("RPROD", self._mat(self.ew("SUB", [one, x], isz), isz))], sizes)
            elif variant == "godel":
                soft = R("RMIN" if op == "forall" else "RMAX", x)
            else:
                if op == "forall":
                    s = R("RSUM", self._mat(self.ew("SUB", [x, one], isz), isz))
                    soft = self.ew("RELU", [self.ew("ADD", [s, one], sizes)], sizes)
                else:
                    s = R("RSUM", self._mat(self.ew("NEG", [x], isz), isz))
                    soft = self.ew("SUB", [one, self.ew("RELU", [self.ew("ADD", [s, one], sizes)],
                                                        sizes)], sizes)
            if self.relax["use_logic_ste"]:
                hb = self._mat(self.ew("GT", [x, self.const_tv(np.float32(0.5))], isz), isz)
                hard = R("RIMIN" if op == "forall" else "RIMAX", hb)
                hard = TV(hard.v, hard.vars)
                soft = self.ste(soft, TV(self._retag(hard.v, "b"), hard.vars), sizes)
            return soft
        body = self.at_least_int(body)
        dvars, ptr, idx, K = self._csr(body, n_outer, isz)
        isf = body.dtype == "f"
        if op in ("argmin", "argmax"):
            if len(tv) != 1:
                raise RDDLNotImplementedError("argmin/argmax over more than one variable")
            if self.relaxed and self.variants.get("argmax") == "softmax" and fluent:
                w = TV(self.mparam(e, "argmax_weight", self.relax["argmax_weight"]), ())
                x = self.to_f(body, isz)
                if op == "argmin":
                    x = self.ew("NEG", [x], isz)
                x = TV(self._mat(x, isz).v, x.vars)
                soft = self._soft_argmax(x, w, n_outer, isz, sizes)
                if self.relax["use_argmax_ste"]:
                    dv, p2, i2, _ = self._csr(x, n_outer, isz)
                    hard = TV(self.g.reduce("RARGMAX", x.v, p2, i2, "i"), dv, None)
                    soft = self.ste(soft, hard, sizes)
                return soft
            if not isf:
                body = self.to_f(body, isz)
                dvars, ptr, idx, K = self._csr(body, n_outer, isz)
            return TV(self.g.reduce("RARGMAX" if op == "argmax" else "RARGMIN", body.v, ptr, idx,
                                    "i"), dvars, None)
        if op == "sum":
            return red("RSUM" if isf else "RISUM", body.v)
        if op == "avg":
            body = self.to_f(body, isz)
            dvars, ptr, idx, K = self._csr(body, n_outer, isz)
            s = TV(self.g.reduce("RSUM", body.v, ptr, idx), dvars, None)
            return self.ew("DIV", [s, self.const_tv(np.float32(K))], sizes)
        if op == "prod":
            return red("RPROD" if isf else "RIPROD", body.v)
        if op == "minimum":
            return red("RMIN" if isf else "RIMIN", body.v)
        if op == "maximum":
            return red("RMAX" if isf else "RIMAX", body.v)
        raise RDDLNotImplementedError(op)

    def _mat(self, tv: TV, sizes) -> TV:
        return TV(self.materialise(tv, sizes), tv.vars, None, tv.full)

    def _soft_argmax(self, x: TV, w: TV, n_outer: int, isz, sizes) -> TV:
        """sum_i i * softmax(w x)_i over the reduced axis  (logic.py:377-379)."""
        wx = self._mat(self.ew("MUL", [w, x], isz), isz)
        dv, ptr, idx, K = self._csr(wx, n_outer, isz)
        mx = TV(self.g.sg(self.g.reduce("RMAX", wx.v, ptr, idx)), dv, None)
        ex = self._mat(self.ew("EXP", [self.ew("SUB", [wx, mx], isz)], isz), isz)
        dv, ptr, idx, K = self._csr(ex, n_outer, isz)
        Z = TV(self.g.reduce("RSUM", ex.v, ptr, idx), dv, None)
        p = self.ew("DIV", [ex, Z], isz)
        lit = TV(self.g.const(np.arange(isz[-1], dtype=np.float32)), (len(isz) - 1,), None)
        term = self._mat(self.ew("MUL", [lit, p], isz), isz)
        dv, ptr, idx, K = self._csr(term, n_outer, isz)
        return TV(self.g.reduce("RSUM", term.v, ptr, idx), dv, None)

    # ---- matrix algebra and random vectors: compiler.py:2247-2434 --------------------------------
    # The reference hands [.., n, n] blocks to jnp.linalg.  Here n is the object count of a type, known
    # at build time, so every routine is unrolled over the matrix entries: entry (i, j) is a value over
    # the remaining tuple axes, and det / inverse / cholesky / L z become straight-line element-wise
    # arithmetic of the step graph (no new instruction, the adjoint comes for free).
    MATRIX_MAX_N = 6

    def _slice_axes(self, tv: TV, fixed: Dict[int, int], sizes) -> TV:
        """``tv`` restricted to scope position p = fixed[p]; the result no longer varies over them."""
        idx = tv.idx
        if idx is None:
            idx = np.arange(tv.v.n, dtype=np.int64).reshape(tuple(sizes[p] for p in tv.vars))
        take = tuple(fixed[p] if p in fixed else slice(None) for p in tv.vars)
        vars_ = tuple(p for p in tv.vars if p not in fixed)
        sub = np.asarray(idx[take])
        return TV(tv.v, vars_, sub.reshape(tuple(sizes[p] for p in vars_)), tv.full)

    def _stack_axes(self, parts: Dict[Tuple[int, ...], TV], axes: Tuple[int, ...], sizes) -> TV:
        """One value over ``axes`` + the parts' own axes whose slice at axes = key is parts[key]
        (select chain over constant one-hot masks; entries without a part read zero)."""
        shape = tuple(sizes[a] for a in axes)
        order = sorted(range(len(axes)), key=lambda k: axes[k])
        out = self.const_tv(np.float32(0.), full=True)
        for key, part in parts.items():
            mask = np.zeros(shape, dtype=np.int32)
            mask[key] = 1
            mtv = TV(self.g.const(np.transpose(mask, order).copy(), "b"),
                     tuple(axes[k] for k in order), None, True)
            out = self.ew("SELECT", [mtv, part, out], sizes, dtype="f")
        return out

    def _matrix_entries(self, body: TV, r_ax: int, c_ax: int, sizes):
        n = sizes[r_ax]
        if sizes[c_ax] != n:
            raise RDDLNotImplementedError("matrix operations need a square matrix")
        if n > self.MATRIX_MAX_N:
            raise RDDLNotImplementedError(
                f"matrix operations are unrolled over the entries: n = {n} > {self.MATRIX_MAX_N}")
        return n, [[self._slice_axes(body, {r_ax: i, c_ax: j}, sizes) for j in range(n)] for i in range(n)]

    def _minor_det(self, A, rows: Tuple[int, ...], cols: Tuple[int, ...], sizes, memo) -> TV:
        """Determinant of the sub-matrix A[rows][cols] by cofactor expansion along its first row."""
        key = (rows, cols)
        if key in memo:
            return memo[key]
        E = lambda nm, *x: self.ew(nm, list(x), sizes)     # noqa: E731
        if len(rows) == 1:
            out = A[rows[0]][cols[0]]
        else:
            out = None
            for k, c in enumerate(cols):
                sub = self._minor_det(A, rows[1:], cols[:k] + cols[k + 1:], sizes, memo)
                term = E("MUL", A[rows[0]][c], sub)
                if out is None:
                    out = term
                else:
                    out = E("SUB" if k % 2 else "ADD", out, term)
        memo[key] = out
        return out

    def _inverse_entries(self, A, n: int, sizes):
        """adj(A) / det(A), entry by entry (branch-free and differentiable; jnp.linalg.inv pivots, the
        result is the same matrix)."""
        E = lambda nm, *x: self.ew(nm, list(x), sizes)     # noqa: E731
        memo: Dict = {}
        full = tuple(range(n))
        det = self._minor_det(A, full, full, sizes, memo)
        if n == 1:
            return [[E("DIV", self.const_tv(np.float32(1.)), det)]]
        inv = [[None] * n for _ in range(n)]
        for i in range(n):
            for j in range(n):
                rows = full[:j] + full[j + 1:]
                cols = full[:i] + full[i + 1:]
                cof = self._minor_det(A, rows, cols, sizes, memo)
                if (i + j) % 2:
                    cof = E("NEG", cof)
                inv[i][j] = E("DIV", cof, det)
        return inv

    def _cholesky_entries(self, A, n: int, sizes):
        """Lower factor of A = L L^T (Cholesky-Banachiewicz); entries above the diagonal are zero."""
        E = lambda nm, *x: self.ew(nm, list(x), sizes)     # noqa: E731
        L = [[None] * n for _ in range(n)]
        for i in range(n):
            for j in range(i + 1):
                acc = A[i][j]
                for k in range(j):
                    acc = E("SUB", acc, E("MUL", L[i][k], L[j][k]))
                L[i][j] = E("SQRT", acc) if i == j else E("DIV", acc, L[j][j])
        return L

    def _matrix(self, e, scope, env, sizes) -> TV:
        _, op = e.etype
        F = lambda t, sz: self.to_f(self.at_least_int(t), sz)   # noqa: E731
        if op == "det":                               # compiler.py:2398-2407
            *tv, body_e = e.args
            inner = list(scope) + list(tv)
            isz = self.sizes(inner)
            body = F(self.lower(body_e, inner, env), isz)
            n, A = self._matrix_entries(body, len(scope), len(scope) + 1, isz)
            full = tuple(range(n))
            det = self._minor_det(A, full, full, isz, {})
            return TV(self._mat(det, isz).v, det.vars, None, True)
        (row, col), body_e = e.args                   # compiler.py:2409-2434
        svars = [v for v, _ in scope]
        if row not in svars or col not in svars or row == col:
            raise RDDLTraceError(f"{op}[row={row}, col={col}]: both must be distinct variables in scope")
        r_ax, c_ax = svars.index(row), svars.index(col)
        body = F(self.lower(body_e, scope, env), sizes)
        n, A = self._matrix_entries(body, r_ax, c_ax, sizes)
        if op == "cholesky":
            R = self._cholesky_entries(A, n, sizes)
        elif op == "inverse":
            R = self._inverse_entries(A, n, sizes)
        elif op == "pinverse":
            # full-rank square matrices: pinv(A) = (A^T A)^-1 A^T (jnp.linalg.pinv truncates small
            # singular values; a rank-deficient argument gives inf / nan here instead)
            E = lambda nm, *x: self.ew(nm, list(x), sizes)     # noqa: E731

            def dot(us, vs):
                acc = None
                for u, v in zip(us, vs):
                    t = E("MUL", u, v)
                    acc = t if acc is None else E("ADD", acc, t)
                return acc
            G = [[dot([A[k][i] for k in range(n)], [A[k][j] for k in range(n)]) for j in range(n)]
                 for i in range(n)]
            Gi = self._inverse_entries(G, n, sizes)
            R = [[dot(Gi[i], [A[j][k] for k in range(n)]) for j in range(n)] for i in range(n)]
        else:
            raise RDDLNotImplementedError(f"Matrix operation {op} is not supported.")
        parts = {(i, j): R[i][j] for i in range(n) for j in range(n) if R[i][j] is not None}
        return self._stack_axes(parts, (r_ax, c_ax), sizes)

    def _random_vector(self, e, scope, env, sizes) -> TV:
        """compiler.py:2247-2377.  Name[?i, (?j)](args): the draw is joint over the objects of ?i."""
        _, name = e.etype
        svars_, *arg_es = e.args
        names = [v for v, _ in scope]
        if svars_[0] not in names:
            raise RDDLTraceError(f"{name}[{svars_[0]}]: the sample variable must be in scope {names}")
        i_ax = names.index(svars_[0])
        n = sizes[i_ax]
        if n > self.MATRIX_MAX_N:
            raise RDDLNotImplementedError(f"{name} over {n} > {self.MATRIX_MAX_N} objects")
        E = lambda nm, *x: self.ew(nm, list(x), sizes)     # noqa: E731
        C = lambda x: self.const_tv(np.float32(x))         # noqa: E731
        F = lambda t, sz=sizes: self.to_f(self.at_least_int(t), sz)   # noqa: E731
        col = lambda tv, a: self._slice_axes(tv, {i_ax: a}, sizes)    # noqa: E731
        if name in ("MultivariateNormal", "MultivariateStudent"):
            if len(svars_) != 2:
                raise RDDLTraceError(f"{name}[?i, ?j](mean(?i), cov(?i, ?j)[, df]): name the column variable")
            inner = list(scope) + [(svars_[1], scope[i_ax][1])]
            isz = self.sizes(inner)
            mean = F(self.lower(arg_es[0], scope, env))
            cov = self.to_f(self.at_least_int(self.lower(arg_es[1], inner, env)), isz)
            _, Cm = self._matrix_entries(cov, i_ax, len(scope), isz)
            # entries of cov no longer vary over ?j: they are values over the outer scope again
            Cm = [[TV(t.v, t.vars, t.idx, t.full) for t in row] for row in Cm]
            L = self._cholesky_entries(Cm, n, sizes)
            Z = self._draw("normal", arg_es[0], scope, sizes) if self._ref_full(arg_es[0]) else None
            if Z is None:
                raise RDDLTraceError(f"{name}: the mean must vary over the sample variable")
            if name == "MultivariateStudent":           # compiler.py:2307-2321: independent t draws
                df = F(self.lower(arg_es[2], scope, env))
                G = self._gamma_draw(df, arg_es[0], scope, sizes, scale=0.5)
                self.errchk("MULTIVARIATE_STUDENT", E("LE", df, C(0.)), sizes)
                Z = E("MUL", Z, E("SQRT", E("DIV", E("MUL", C(0.5), df), G)))
            parts = {}
            for a in range(n):
                acc = col(mean, a)
                for b in range(a + 1):
                    acc = E("ADD", acc, E("MUL", L[a][b], col(Z, b)))
                parts[(a,)] = acc
            return self._stack_axes(parts, (i_ax,), sizes)
        if name == "Dirichlet":                          # compiler.py:2331-2348
            alpha = F(self.lower(arg_es[0], scope, env))
            G = self._gamma_draw(alpha, arg_es[0], scope, sizes)
            self.errchk("DIRICHLET", E("LE", alpha, C(0.)), sizes)
            total = None
            for a in range(n):
                total = col(G, a) if total is None else E("ADD", total, col(G, a))
            return self._stack_axes({(a,): E("DIV", col(G, a), total) for a in range(n)}, (i_ax,), sizes)
        if name == "Multinomial":                        # compiler.py:2350-2377 (tfp sampler)
            # the reference's tfp bits cannot be reproduced from supplied noise: same law from COUNT_NBINS
            # uniforms, trial k < trials falls into category a when cum_{a-1} <= U_k < cum_a (the last
            # category takes the remainder); counts are exact integers, no gradient (as in the reference)
            trials = self.lower(arg_es[0], scope, env)
            if trials.dtype == "f":
                trials = E("F2I", trials)
            trials = self.at_least_int(trials)
            prob = F(self.lower(arg_es[1], scope, env))
            if i_ax in trials.vars:
                raise RDDLTraceError("Multinomial: trials must not vary over the sample variable")
            cum, acc = [], None
            for a in range(n):
                acc = col(prob, a) if acc is None else E("ADD", acc, col(prob, a))
                cum.append(acc)
            if n < 2:
                raise RDDLTraceError("Multinomial needs at least two categories")
            bad = E("OR", E("LT", prob, C(0.)), E("ILT", trials, self.const_tv(np.int32(0), "i")))
            bad_sum = E("GT", E("ABS", E("SUB", cum[-1], C(1.))), C(1e-5))
            self.errchk("MULTINOMIAL", E("OR", bad, bad_sum), sizes)
            counts = [None] * n
            for k in range(COUNT_NBINS):
                U = col(self._draw("uniform", arg_es[1], scope, sizes), 0)
                live = E("ILT", self.const_tv(np.int32(k), "i"), trials)
                below_prev = None
                for a in range(n):
                    below = E("LT", U, cum[a]) if a < n - 1 else None
                    if a == 0:
                        hit = below
                    elif a < n - 1:
                        hit = E("AND", below, E("NOT", below_prev))
                    else:
                        hit = E("NOT", below_prev)
                    below_prev = below
                    inc = self.at_least_int(E("AND", hit, live))
                    counts[a] = inc if counts[a] is None else E("IADD", counts[a], inc)
            parts = {(a,): self.to_f(counts[a], sizes) for a in range(n)}
            out = self._stack_axes(parts, (i_ax,), sizes)
            if self.relaxed:
                return TV(self.g.sg(out.v), out.vars, out.idx, out.full)
            return self.ew("F2I", [out], sizes, "i")
        raise RDDLNotImplementedError(f"Distribution {name} is not supported.")

    # ---- functions: compiler.py:1354-1529 ; logic.py:346-358, 785-896 -------------------
    _UNARY_F = {"cos": "COS", "sin": "SIN", "tan": "TAN", "acos": "ACOS", "asin": "ASIN",
                "atan": "ATAN", "cosh": "COSH", "sinh": "SINH", "tanh": "TANH", "exp": "EXP",
                "ln": "LOG", "sqrt": "SQRT", "lngamma": "LGAMMA"}

    def soft_floor(self, x: TV, w: TV, sizes) -> TV:
        """logic.py:800-804."""
        E = lambda name, *a: self.ew(name, list(a), sizes)     # noqa: E731
        half, one = self.const_tv(np.float32(0.5)), self.const_tv(np.float32(1.))
        r = TV(self.g.sg(E("FLOOR", E("ADD", x, half)).v), x.vars)
        d = E("SUB", x, r)
        ratio = E("DIV", E("STANH", E("MUL", w, d)), E("STANH", E("DIV", w, self.const_tv(np.float32(2.)))))
        return E("ADD", E("SUB", r, one), E("MUL", half, E("ADD", one, ratio)))

    def soft_round(self, x: TV, w: TV, sizes) -> TV:
        """logic.py:879-882."""
        E = lambda name, *a: self.ew(name, list(a), sizes)     # noqa: E731
        half = self.const_tv(np.float32(0.5))
        m = E("ADD", TV(self.g.sg(E("FLOOR", x).v), x.vars), half)
        ratio = E("DIV", E("STANH", E("MUL", w, E("SUB", x, m))),
                  E("STANH", E("DIV", w, self.const_tv(np.float32(2.)))))
        return E("ADD", m, E("MUL", half, ratio))

    def _function(self, e, scope, env, sizes) -> TV:
        _, op = e.etype
        tvs = [self.at_least_int(t) for t in self._args(e, scope, env)]
        fluent = self.tr.is_fluent(e)
        R = self.relax if self.relaxed else {}
        Vv = self.variants if self.relaxed else {}
        E = lambda name, *a: self.ew(name, list(a), sizes)     # noqa: E731
        if len(tvs) == 1:
            x = tvs[0]
            if op == "sgn" and Vv.get("sgn") == "tanh" and fluent:
                w = TV(self.mparam(e, "sigmoid_weight", R["sigmoid_weight"]), ())
                xf = self.to_f(x, sizes)
                soft = E("STANH", E("MUL", w, xf))
                return self.ste(soft, E("SGN", xf), sizes) if R["use_tanh_ste"] else soft
            if op in ("floor", "ceil") and Vv.get("floor") == "soft" and fluent:
                w = TV(self.mparam(e, "floor_weight", R["floor_weight"]), ())
                xf = self.to_f(x, sizes)
                if op == "floor":
                    soft, hard = self.soft_floor(xf, w, sizes), E("FLOOR", xf)
                else:
                    soft = E("NEG", self.soft_floor(E("NEG", xf), w, sizes))
                    hard = E("CEIL", xf)
                return self.ste(soft, hard, sizes) if R["use_floor_ste"] else soft
            if op == "round" and Vv.get("round") == "soft" and fluent:
                w = TV(self.mparam(e, "round_weight", R["round_weight"]), ())
                xf = self.to_f(x, sizes)
                soft = self.soft_round(xf, w, sizes)
                return self.ste(soft, E("ROUND", xf), sizes) if R["use_round_ste"] else soft
            if op in ("abs", "sgn"):
                name = {"abs": "ABS", "sgn": "SGN"}[op]
                return E(name if x.dtype == "f" else "I" + name, x)
            if op in ("round", "floor", "ceil"):
                if x.dtype != "f":
                    return x
                return E(op.upper(), x)
            xf = self.to_f(x, sizes)
            if op == "gamma":
                return E("EXP", E("LGAMMA", xf))
            return E(self._UNARY_F[op], xf)
        x, y = tvs
        if op in ("div", "mod") and Vv.get("floor") == "soft" and fluent:
            w = TV(self.mparam(e, "floor_weight", R["floor_weight"]), ())
            xf, yf = self.to_f(x, sizes), self.to_f(y, sizes)
            d = self.soft_floor(E("DIV", xf, yf), w, sizes)
            if R["use_floor_ste"]:
                d = self.ste(d, E("FDIV", xf, yf), sizes)
            return d if op == "div" else E("SUB", xf, E("MUL", yf, d))
        (x, y), dt = self.promote([x, y], sizes)
        I = "" if dt == "f" else "I"
        if op == "div":
            return E(I + "FDIV", x, y)
        if op == "mod":
            return E(I + "MOD", x, y)
        if op == "fmod":
            return E(I + "FMOD", x, y)
        if op == "min":
            return E(I + "MIN", x, y)
        if op == "max":
            return E(I + "MAX", x, y)
        if op == "pow":
            return E(I + "POW", x, y)
        xf, yf = self.to_f(x, sizes), self.to_f(y, sizes)
        if op == "log":
            return E("DIV", E("LOG", xf), E("LOG", yf))
        if op == "hypot":
            return E("HYPOT", xf, yf)
        raise RDDLNotImplementedError(op)

    # ---- control: compiler.py:1617-1685 ; logic.py:903-941 ------------------------------
    def _control(self, e, scope, env, sizes) -> TV:
        _, op = e.etype
        E = lambda name, *x: self.ew(name, list(x), sizes)     # noqa: E731
        if op == "switch":
            return self._switch(e, scope, env, sizes)
        c, a, b = self._args(e, scope, env)
        if self.relaxed and self.variants.get("if") == "linear" and self.tr.is_fluent(e.args[0]):
            cf = self.to_f(self.at_least_int(c), sizes)
            if self.relax["use_if_else_ste"]:
                cf = self.ste(cf, E("GT", cf, self.const_tv(np.float32(0.5))), sizes)
            af = self.to_f(self.at_least_int(a), sizes)
            bf = self.to_f(self.at_least_int(b), sizes)
            return E("LERP", cf, af, bf)
        self.static_cast_err(c.dtype == "b")
        if c.dtype == "f":
            c = E("GT", c, self.const_tv(np.float32(0.5)))
        else:
            c = self.to_b(c, sizes)
        if a.dtype != b.dtype:
            (a, b), _ = self.promote([a, b], sizes)
        out = self.ew("SELECT", [c, a, b], sizes, dtype=a.dtype)
        return out

    def _switch(self, e, scope, env, sizes) -> TV:
        """switch over an enum-valued predicate.  Exact: the case selected by the integer
        predicate (take_along_axis over the stacked cases, compiler.py:1661-1684), here a chain
        of selects -- every case is evaluated, as in the reference.  Relaxed (SoftmaxSwitch,
        logic.py:959-1002): cases weighted by softmax_k(-w |pred - k|), straight-through to the
        hard selection when use_switch_ste."""
        pred_e, cases = e.args
        E = lambda name, *x: self.ew(name, list(x), sizes)     # noqa: E731
        C = lambda x: self.const_tv(np.float32(x))             # noqa: E731
        table = dict(cases)
        if pred_e.etype[0] == "pvar" and pred_e.args[0] in self.m.variable_ranges:
            etype = self.m.variable_ranges[pred_e.args[0]]
        else:
            # a computed predicate (if / argmax / nested expression): the enum type is the one the case
            # labels belong to (pyRDDLGym's tracer infers it from the predicate; the labels must agree)
            types = {self.m.object_to_type.get(lbl) for lbl in table if lbl is not None}
            if len(types) != 1 or None in types:
                raise RDDLNotImplementedError(
                    "switch over a computed predicate: the case labels must be objects of one enum type")
            etype = types.pop()
        objs = self.m.type_to_objects[etype]
        default = table.get(None)
        p = self.lower(pred_e, scope, env)
        vals = []
        for o in objs:
            ce = table.get(o, default)
            if ce is None:
                raise RDDLNotImplementedError(f"switch is missing case <{o}> and has no default")
            vals.append(self.lower(ce, scope, env))
        vals, dt = self.promote(vals, sizes)
        K = len(objs)

        def hard_select() -> TV:
            acc = vals[K - 1]
            for k in range(K - 2, -1, -1):
                hit = E("EQ", p, C(k)) if p.dtype == "f" else \
                    E("IEQ", self.at_least_int(p), self.const_tv(np.int32(k), "i"))
                acc = self.ew("SELECT", [hit, vals[k], acc], sizes, dtype=dt)
            return acc
        if not (self.relaxed and self.variants.get("switch") == "softmax"
                and self.tr.is_fluent(pred_e)):
            return hard_select()
        w = TV(self.mparam(e, "switch_weight", self.relax["switch_weight"]), ())
        pf = self.to_f(self.at_least_int(p), sizes)
        logits = [E("NEG", E("MUL", w, E("ABS", E("SUB", pf, C(k))))) for k in range(K)]
        top = logits[0]
        for lk in logits[1:]:
            top = E("MAX", top, lk)
        num = den = None
        for lk, vk in zip(logits, vals):
            ek = E("EXP", E("SUB", lk, top))
            term = E("MUL", ek, self.to_f(vk, sizes))
            num = term if num is None else E("ADD", num, term)
            den = ek if den is None else E("ADD", den, ek)
        soft = E("DIV", num, den)
        if self.relax["use_switch_ste"]:
            soft = self.ste(soft, hard_select(), sizes)
        return soft

    # ---- random variables: compiler.py:1701-2241 ; logic.py:1010-1768 --------------------
    GAMMA_PROPOSALS = 4     # Marsaglia-Tsang proposals per draw; each is accepted with p >= 0.95

    def _gamma_sampler(self, shape: TV, a0: Expr, scope, sizes, scale: float) -> TV:
        """standard Gamma(scale * shape) for a FLUENT shape, from supplied draws: Marsaglia-Tsang
        on alpha + 1 (d = alpha + 2/3, c = 1 / sqrt(9 d); proposal d (1 + c z)^3, accepted when
        1 + c z > 0 and log u < z^2 / 2 + d - d v + d log v), the first accepted of GAMMA_PROPOSALS
        proposals (the last one if none is), times U^(1 / alpha) to come back to alpha.  Draw
        order: (normal, uniform) per proposal, then the uniform of the power.  Differentiable in
        alpha through d, c and the exponent (the accept tests are constants)."""
        E = lambda nm, *x: self.ew(nm, list(x), sizes)     # noqa: E731
        C = lambda x: self.const_tv(np.float32(x))         # noqa: E731
        alpha = shape if scale == 1.0 else E("MUL", C(scale), shape)
        d = E("ADD", alpha, C(2.0 / 3.0))
        c = E("DIV", C(1.0), E("SQRT", E("MUL", C(9.0), d)))
        props = []
        for _ in range(self.GAMMA_PROPOSALS):
            z = self._draw("normal", a0, scope, sizes)
            u = self._draw("uniform", a0, scope, sizes)
            t = E("ADD", C(1.0), E("MUL", c, z))
            v = E("MUL", E("MUL", t, t), t)
            logv = E("LOG", E("MAX", v, C(1e-30)))
            bound = E("ADD", E("SUB", E("ADD", E("MUL", C(0.5), E("MUL", z, z)), d), E("MUL", d, v)),
                      E("MUL", d, logv))
            ok = E("AND", E("GT", t, C(0.0)), E("LT", E("LOG", u), bound))
            props.append((ok, E("MUL", d, E("MAX", v, C(1e-30)))))
        x = props[-1][1]
        for ok, xv in reversed(props[:-1]):
            x = self.ew("SELECT", [ok, xv, x], sizes, dtype="f")
        ub = self._draw("uniform", a0, scope, sizes)
        return E("MUL", x, E("EXP", E("DIV", E("LOG", ub), alpha)))

    def _gamma_draw(self, shape: TV, a0: Expr, scope, sizes, scale: float = 1.0) -> TV:
        """standard Gamma(scale * shape) noise: supplied from outside the kernel when the shape is
        known at build time, sampled in the step from supplied normal / uniform
        draws when it is a fluent."""
        if shape.v.const is None:
            return self._gamma_sampler(shape, a0, scope, sizes, scale)
        full = self._ref_full(a0)
        allv = tuple(range(len(sizes)))
        if full:
            from .graph import apply_map
            _, mp = self.operand(shape, allv, sizes)
            conc = shape.v.const[apply_map(mp, self.n_of(allv, sizes))]
        else:
            conc = shape.v.const[:1]
        conc = np.asarray(conc, dtype=np.float32) * np.float32(scale)
        return self._draw("gamma", a0, scope, sizes,
                          extra=conc.reshape(tuple(sizes) if full else ()))

    def _draw(self, kind: str, first_arg: Expr, scope, sizes, extra=None) -> TV:
        full = self._ref_full(first_arg)
        shape = tuple(sizes) if full else ()
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        v = self.g.leaf("noise", n, "f", slot=len(self.sg.noise))
        self.sg.noise.append(NoiseSlot(kind, shape, n, v, extra))
        return TV(v, tuple(range(len(sizes))) if full else (), None, full)

    def _random(self, e, scope, env, sizes) -> TV:
        _, name = e.etype
        R = self.relax if self.relaxed else {}
        Vv = self.variants if self.relaxed else {}
        E = lambda nm, *x: self.ew(nm, list(x), sizes)     # noqa: E731
        C = lambda x: self.const_tv(np.float32(x))         # noqa: E731
        F = lambda t: self.to_f(self.at_least_int(t), sizes)   # noqa: E731
        if name == "KronDelta":
            v = self.lower(e.args[0], scope, env)
            if not self.relaxed and v.dtype == "f":
                self.sg.static_err |= 1 << ERR_BITS["KRON"]
            return v
        if name == "DiracDelta":
            return F(self.lower(e.args[0], scope, env))
        if name in ("Discrete", "UnnormDiscrete"):
            return self._discrete(e, scope, env, sizes, name == "UnnormDiscrete")
        tvs = self._args(e, scope, env)
        a0 = e.args[0]
        if name == "Uniform":
            lb, ub = F(tvs[0]), F(tvs[1])
            U = self._draw("uniform", a0, scope, sizes)
            self.errchk("UNIFORM", E("GT", lb, ub), sizes)
            return E("ADD", lb, E("MUL", E("SUB", ub, lb), U))
        if name == "Normal":
            mean, var = F(tvs[0]), F(tvs[1])
            Z = self._draw("normal", a0, scope, sizes)
            self.errchk("NORMAL", E("LT", var, C(0.)), sizes)
            return E("ADD", mean, E("MUL", E("SQRT", var), Z))
        if name == "Exponential":
            scale = F(tvs[0])
            Ex = self._draw("exponential", a0, scope, sizes)
            self.errchk("EXPONENTIAL", E("LE", scale, C(0.)), sizes)
            return E("MUL", scale, Ex)
        if name == "Weibull":
            shape, scale = F(tvs[0]), F(tvs[1])
            Ex = self._draw("exponential", a0, scope, sizes)
            self.errchk("WEIBULL", E("OR", E("LE", shape, C(0.)), E("LE", scale, C(0.))), sizes)
            return E("MUL", scale, E("POW", Ex, E("DIV", C(1.), shape)))
        if name == "Gamma":
            shape, scale = F(tvs[0]), F(tvs[1])
            G = self._gamma_draw(shape, a0, scope, sizes)
            self.errchk("GAMMA", E("OR", E("LE", shape, C(0.)), E("LE", scale, C(0.))), sizes)
            return E("MUL", scale, G)
        if name == "ChiSquare":                 # compiler.py:2130-2147: 2 Gamma(df / 2)
            df = F(tvs[0])
            G = self._gamma_draw(df, a0, scope, sizes, scale=0.5)
            self.errchk("CHISQUARE", E("LE", df, C(0.)), sizes)
            return E("MUL", C(2.), G)
        if name == "Student":                   # compiler.py:2028-2044: Z sqrt((df/2) / Gamma(df/2))
            df = F(tvs[0])
            Z = self._draw("normal", a0, scope, sizes)
            G = self._gamma_draw(df, a0, scope, sizes, scale=0.5)
            self.errchk("STUDENT", E("LE", df, C(0.)), sizes)
            return E("MUL", Z, E("SQRT", E("DIV", E("MUL", C(0.5), df), G)))
        if name == "Beta":                      # compiler.py:1969-1986: Ga / (Ga + Gb)
            a, b = F(tvs[0]), F(tvs[1])
            Ga = self._gamma_draw(a, a0, scope, sizes)
            Gb = self._gamma_draw(b, a0, scope, sizes)
            self.errchk("BETA", E("OR", E("LE", a, C(0.)), E("LE", b, C(0.))), sizes)
            return E("DIV", Ga, E("ADD", Ga, Gb))
        if name == "Pareto":                    # compiler.py:2007-2026: scale exp(Exp(1) / shape)
            shape, scale = F(tvs[0]), F(tvs[1])
            Ex = self._draw("exponential", a0, scope, sizes)
            self.errchk("PARETO", E("OR", E("LE", shape, C(0.)), E("LE", scale, C(0.))), sizes)
            return E("MUL", scale, E("EXP", E("DIV", Ex, shape)))
        if name == "Poisson":
            return self._poisson(e, F(tvs[0]), a0, scope, sizes)
        if name == "Binomial":
            return self._binomial(e, tvs, scope, sizes)
        if name == "NegativeBinomial":
            trials, prob = F(tvs[0]), F(tvs[1])
            fluent = self.tr.is_fluent(e.args[0]) or self.tr.is_fluent(e.args[1])
            bad = E("OR", E("OR", E("LT", prob, C(0.)), E("GT", prob, C(1.))), E("LE", trials, C(0.)))
            self.errchk("NEGATIVE_BINOMIAL", bad, sizes)
            variant = Vv.get("poisson") if (self.relaxed and fluent) else None
            if variant == "determinized":
                return E("MUL", E("SUB", E("DIV", C(1.), prob), C(1.)), trials)   # logic.py:1762
            G = self._gamma_draw(trials, a0, scope, sizes)
            if variant is None:
                # exact: tfp NegativeBinomial(total_count, probs) (compiler.py:1945-1967) as its
                # Gamma-Poisson mixture, rate = Gamma(trials) p / (1 - p)
                rate = E("DIV", E("MUL", G, prob), E("SUB", C(1.), prob))
            else:
                # relaxed (logic.py:1583-1585, 1698-1700): rate = (1 / p - 1) Gamma(trials)
                rate = E("MUL", E("SUB", E("DIV", C(1.), prob), C(1.)), G)
            return self._poisson_sample(e, rate, a0, scope, sizes, variant)
        if name in ("Gumbel", "Laplace", "Cauchy"):
            mean, scale = F(tvs[0]), F(tvs[1])
            Gn = self._draw(name.lower(), a0, scope, sizes)
            self.errchk(name.upper(), E("LE", scale, C(0.)), sizes)
            return E("ADD", mean, E("MUL", scale, Gn))
        if name == "Gompertz":
            shape, scale = F(tvs[0]), F(tvs[1])
            Ex = self._draw("exponential", e.args[1], scope, sizes)
            self.errchk("GOMPERTZ", E("OR", E("LE", shape, C(0.)), E("LE", scale, C(0.))), sizes)
            return E("DIV", E("DIV", E("LOG1P", Ex), shape), scale)
        if name == "Kumaraswamy":
            a, b = F(tvs[0]), F(tvs[1])
            U = self._draw("uniform", a0, scope, sizes)
            self.errchk("KUMARASWAMY", E("OR", E("LE", a, C(0.)), E("LE", b, C(0.))), sizes)
            inner = E("SUB", C(1.), E("POW", U, E("DIV", C(1.), b)))
            return E("POW", inner, E("DIV", C(1.), a))
        if name == "Bernoulli":
            p = tvs[0]
            variant = Vv.get("bernoulli") if self.tr.is_fluent(a0) else None
            pf = F(p)
            self.errchk("BERNOULLI", E("OR", E("LT", pf, C(0.)), E("GT", pf, C(1.))), sizes)
            if variant is None:
                U = self._draw("uniform", a0, scope, sizes)
                return E("LT", U, pf)
            if variant == "sigmoid":
                w = TV(self.mparam(e, "bernoulli_sigmoid_weight", R["bernoulli_sigmoid_weight"]), ())
                U = self._draw("uniform", a0, scope, sizes)
                s = E("SIGMOID", E("MUL", w, E("SUB", pf, U)))
                return self.ste(s, E("GT", pf, U), sizes) if R["use_ste_bernoulli"] else s
            if variant == "gumbel":
                w = TV(self.mparam(e, "bernoulli_sigmoid_weight", R["bernoulli_sigmoid_weight"]), ())
                eps = TV(self.mparam(e, "bernoulli_prob_eps", R["bernoulli_prob_eps"]), ())
                L = self._draw("logistic", a0, scope, sizes)
                pc = E("MIN", E("MAX", pf, eps), E("SUB", C(1.), eps))
                logit = E("ADD", E("LOG", E("DIV", pc, E("SUB", C(1.), pc))), L)
                s = E("SIGMOID", E("MUL", w, logit))
                return self.ste(s, E("GT", logit, C(0.)), sizes) if R["use_ste_bernoulli"] else s
            return pf
        if name == "Geometric":
            p = F(tvs[0])
            variant = Vv.get("geometric") if self.tr.is_fluent(a0) else None
            self.errchk("GEOMETRIC", E("OR", E("LT", p, C(0.)), E("GT", p, C(1.))), sizes)
            if variant is None:
                U = self._draw("uniform", a0, scope, sizes)
                q = E("DIV", E("LOG", U), E("LOG1P", E("NEG", p)))
                return self.ew("F2I", [E("ADD", E("FLOOR", q), C(1.))], sizes, "i")
            if variant == "reparam":
                w = TV(self.mparam(e, "geometric_floor_weight", R["geometric_floor_weight"]), ())
                eps = TV(self.mparam(e, "geometric_eps", R["geometric_eps"]), ())
                Ex = self._draw("exponential", a0, scope, sizes)
                ratio = E("DIV", E("NEG", Ex), E("SAFELOG", E("SUB", C(1.), p), eps))
                return E("ADD", C(1.), self.soft_floor(ratio, w, sizes))
            return E("DIV", C(1.), p)
        raise RDDLNotImplementedError(f"Distribution {name} is not supported.")

    def _bins_scope(self, scope, K: int):
        """scope + one trailing axis of K count bins (a synthetic object type)."""
        tname = f"__bins{K}__"
        if tname not in self.m.type_to_objects:
            self.m.type_to_objects[tname] = [f"__bin{i}" for i in range(K)]
        inner = list(scope) + [("?__bin__", tname)]
        return inner, self.sizes(inner)

    def _gumbel_softmax_count(self, log_prob: TV, w: TV, scope, isz, sizes, full: bool = True) -> TV:
        """soft arg-max over the bins of Gumbel noise + log pmf (logic.py:1381-1384, 1626-1629).
        The noise has the shape of the log pmf: the scope's (when an argument carries it) + bins."""
        allv = tuple(range(len(isz)))
        gvars = allv if full else (len(isz) - 1,)
        gshape = tuple(isz[p] for p in gvars)
        n = int(np.prod(gshape, dtype=np.int64))
        gv = self.g.leaf("noise", n, "f", slot=len(self.sg.noise))
        self.sg.noise.append(NoiseSlot("gumbel", gshape, n, gv, None))
        G = TV(gv, gvars, None, True)
        logits = self._mat(self.ew("ADD", [G, log_prob], isz), isz)
        return self._soft_argmax(TV(logits.v, allv, None, True), w, len(scope), isz, sizes)

    def _poisson(self, e, rate: TV, a0, scope, sizes) -> TV:
        """Poisson(rate).  Exact programs invert the CDF of one uniform draw, truncated at
        COUNT_NBINS terms (the reference calls jax.random.poisson, compiler.py:1880-1897, whose
        bits cannot be reproduced from supplied noise; the law is the same for rates whose mass
        above COUNT_NBINS is negligible).  Relaxed (ExponentialPoisson, logic.py:1508-1531): the
        count of exponential inter-arrival times before time 1, each indicator a sigmoid; rates
        whose CDF at nbins falls below poisson_min_cdf use the normal approximation instead.
        DeterminizedPoisson returns the rate (logic.py:1733-1741)."""
        E = lambda nm, *x: self.ew(nm, list(x), sizes)     # noqa: E731
        C = lambda x: self.const_tv(np.float32(x))         # noqa: E731
        self.errchk("POISSON", E("LT", rate, C(0.)), sizes)
        variant = self.variants.get("poisson") if (self.relaxed and self.tr.is_fluent(a0)) else None
        return self._poisson_sample(e, rate, a0, scope, sizes, variant)

    def _poisson_sample(self, e, rate: TV, a0, scope, sizes, variant) -> TV:
        """A Poisson draw of the given rate under the given relaxation (shared with
        NegativeBinomial, a Gamma mixture of Poissons)."""
        E = lambda nm, *x: self.ew(nm, list(x), sizes)     # noqa: E731
        C = lambda x: self.const_tv(np.float32(x))         # noqa: E731
        if variant == "determinized":
            return rate
        if variant == "exponential":
            R = self.relax
            K = int(R["poisson_nbins"])
            w = TV(self.mparam(e, "poisson_comparison_weight", R["poisson_comparison_weight"]), ())
            # largest rate whose CDF at K still reaches min_cdf (the branch is under stop_gradient)
            from scipy import optimize, stats
            cut = float(optimize.brentq(
                lambda r: stats.poisson.cdf(K, r) - float(R["poisson_min_cdf"]), 1e-6, 10.0 * K + 10))
            cum = cnt = None
            for _ in range(K):
                Ex = self._draw("exponential", a0, scope, sizes)
                dt = E("DIV", Ex, rate)
                cum = dt if cum is None else E("ADD", cum, dt)
                ind = E("SIGMOID", E("MUL", w, E("SUB", C(1.), cum)))
                cnt = ind if cnt is None else E("ADD", cnt, ind)
            Z = self._draw("normal", a0, scope, sizes)
            normal = E("ADD", rate, E("MUL", E("SQRT", rate), Z))
            return self.ew("SELECT", [E("LE", rate, C(cut)), cnt, normal], sizes, dtype="f")
        if variant == "gumbel":           # GumbelSoftmaxPoisson, logic.py:1618-1641
            R = self.relax
            K = int(R["poisson_nbins"])
            w = TV(self.mparam(e, "poisson_softmax_weight", R["poisson_softmax_weight"]), ())
            eps = TV(self.mparam(e, "poisson_eps", R["poisson_eps"]), ())
            from scipy import optimize, special, stats
            cut = float(optimize.brentq(
                lambda r: stats.poisson.cdf(K, r) - float(R["poisson_min_cdf"]), 1e-6, 10.0 * K + 10))
            inner, isz = self._bins_scope(scope, K)
            EI = lambda nm, *x: self.ew(nm, list(x), isz)      # noqa: E731
            kpos = len(scope)
            ks = TV(self.g.const(np.arange(K, dtype=np.float32)), (kpos,), None)
            lgk = TV(self.g.const(special.gammaln(np.arange(K) + 1.0).astype(np.float32)), (kpos,), None)
            term = self.ew("SELECT", [EI("EQ", ks, C(0.)), C(0.),
                                      EI("MUL", ks, EI("SAFELOG", rate, eps))], isz, dtype="f")
            log_prob = EI("SUB", EI("SUB", term, rate), lgk)
            cnt = self._gumbel_softmax_count(log_prob, w, scope, isz, sizes, self._ref_full(a0))
            Z = self._draw("normal", a0, scope, sizes)
            normal = E("ADD", rate, E("MUL", E("SQRT", rate), Z))
            return self.ew("SELECT", [E("LE", rate, C(cut)), cnt, normal], sizes, dtype="f")
        if variant is not None:
            raise RDDLNotImplementedError(f"Poisson relaxation <{variant}> is not built")
        U = self._draw("uniform", a0, scope, sizes)
        pk = E("EXP", E("NEG", rate))
        cdf, cnt = pk, None
        for k in range(1, COUNT_NBINS + 1):
            over = self.to_f(self.at_least_int(E("GT", U, cdf)), sizes)
            cnt = over if cnt is None else E("ADD", cnt, over)
            pk = E("MUL", pk, E("MUL", rate, C(1.0 / k)))
            cdf = E("ADD", cdf, pk)
        return cnt if self.relaxed else self.ew("F2I", [cnt], sizes, "i")

    def _binomial(self, e, tvs, scope, sizes) -> TV:
        """Binomial(trials, prob).  Exact: the number of successes among the first ``trials`` of
        COUNT_NBINS supplied uniform draws (compiler.py:1921-1943 calls jax.random.binomial; same
        law for trials <= COUNT_NBINS).  DeterminizedBinomial: trials * prob (logic.py:1476)."""
        E = lambda nm, *x: self.ew(nm, list(x), sizes)     # noqa: E731
        C = lambda x: self.const_tv(np.float32(x))         # noqa: E731
        F = lambda t: self.to_f(self.at_least_int(t), sizes)   # noqa: E731
        trials, prob = F(tvs[0]), F(tvs[1])
        a0 = e.args[0]
        bad = E("OR", E("OR", E("LT", prob, C(0.)), E("GT", prob, C(1.))), E("LT", trials, C(0.)))
        self.errchk("BINOMIAL", bad, sizes)
        fluent = self.tr.is_fluent(e.args[0]) or self.tr.is_fluent(e.args[1])
        variant = self.variants.get("binomial") if (self.relaxed and fluent) else None
        if variant == "determinized":
            return E("MUL", trials, prob)
        if variant == "gumbel":           # GumbelSoftmaxBinomial, logic.py:1381-1434
            R = self.relax
            K = int(R["binomial_nbins"])
            w = TV(self.mparam(e, "binomial_softmax_weight", R["binomial_softmax_weight"]), ())
            eps = TV(self.mparam(e, "binomial_eps", R["binomial_eps"]), ())
            inner, isz = self._bins_scope(scope, K)
            EI = lambda nm, *x: self.ew(nm, list(x), isz)      # noqa: E731
            SEL = lambda c, a, b: self.ew("SELECT", [c, a, b], isz, dtype="f")   # noqa: E731
            kpos = len(scope)
            ks0 = TV(self.g.const(np.arange(K, dtype=np.float32)), (kpos,), None)
            in_support = EI("LE", ks0, trials)
            ks = EI("MIN", ks0, trials)
            q = E("SUB", C(1.), prob)
            term_k = SEL(EI("EQ", ks, C(0.)), C(0.), EI("MUL", ks, EI("SAFELOG", prob, eps)))
            nk = EI("SUB", trials, ks)
            term_nk = SEL(EI("EQ", ks, trials), C(0.), EI("MUL", nk, EI("SAFELOG", q, eps)))
            lg = lambda x: EI("LGAMMA", x)                      # noqa: E731
            comb = EI("SUB", EI("SUB", lg(EI("ADD", trials, C(1.))), lg(EI("ADD", ks, C(1.)))),
                      lg(EI("ADD", nk, C(1.))))
            log_prob = EI("ADD", EI("ADD", comb, term_k), term_nk)
            log_prob = SEL(in_support, log_prob, C(-np.inf))
            cnt = self._gumbel_softmax_count(
                log_prob, w, scope, isz, sizes,
                self._ref_full(e.args[0]) or self._ref_full(e.args[1]))
            Z = self._draw("normal", a0, scope, sizes)          # logic.py:1372-1378
            var = E("MUL", trials, E("MIN", E("MAX", E("MUL", prob, q), C(0.)), C(1.)))
            normal = E("ADD", E("MUL", trials, prob), E("MUL", E("SQRT", var), Z))
            small = E("LT", trials, C(float(K)))
            return self.ew("SELECT", [small, cnt, normal], sizes, dtype="f")
        if variant is not None:
            raise RDDLNotImplementedError(f"Binomial relaxation <{variant}> is not built")
        cnt = None
        for j in range(COUNT_NBINS):
            U = self._draw("uniform", a0, scope, sizes)
            hit = E("AND", E("LT", C(float(j)), trials), E("LT", U, prob))
            hf = self.to_f(self.at_least_int(hit), sizes)
            cnt = hf if cnt is None else E("ADD", cnt, hf)
        return cnt if self.relaxed else self.ew("F2I", [cnt], sizes, "i")

    def _discrete(self, e, scope, env, sizes, unnorm: bool) -> TV:
        """Discrete / UnnormDiscrete over an enumerated type: exact = categorical sample
        argmax_k(Gumbel_k + safe_log p_k) (compiler.py:2203-2215); relaxed = Gumbel-softmax
        soft arg-max, no straight-through (logic.py:1243-1267); determinised = sum_k k p_k
        (logic.py:1321-1327).  The K case probabilities are stacked along an extra trailing
        axis of the tuple space (element = scope index * K + k)."""
        etype_name, cases = e.args
        objs = self.m.type_to_objects[etype_name]
        K = len(objs)
        table = dict(cases)
        R = self.relax if self.relaxed else {}
        inner = list(scope) + [("?__case__", etype_name)]
        isz = self.sizes(inner)
        n_outer, kpos = len(scope), len(scope)
        allv = tuple(range(len(isz)))
        EI = lambda nm, *x: self.ew(nm, list(x), isz)      # noqa: E731
        C = lambda x: self.const_tv(np.float32(x))         # noqa: E731
        prob = None
        for k, o in enumerate(objs):
            pk = self.to_f(self.at_least_int(self.lower(table[o], scope, env)), sizes)
            onehot = TV(self.g.const(np.eye(K, dtype=np.float32)[k]), (kpos,), None)
            term = EI("MUL", onehot, pk)
            prob = term if prob is None else EI("ADD", prob, term)
        prob = self._mat(prob, isz)

        def total(x: TV) -> TV:
            dv, ptr, idx, _ = self._csr(x, n_outer, isz)
            return TV(self.g.reduce("RSUM", x.v, ptr, idx), dv, None)
        tot = total(prob)
        if unnorm:                                   # compiler.py:2197-2199 (sj.div == divide)
            prob = self._mat(EI("DIV", prob, tot), isz)
            tot = total(prob)
        # _jax_update_discrete_oob_error: any p < 0 or not allclose(sum p, 1)
        self.errchk("DISCRETE", EI("LT", prob, C(0.)), isz)
        dev = self.ew("ABS", [self.ew("SUB", [tot, C(1.)], sizes)], sizes)
        self.errchk("DISCRETE", self.ew("GT", [dev, C(1e-5 + 1e-8)], sizes), sizes)
        fluent = any(self.tr.is_fluent(table[o]) for o in objs)
        variant = self.variants.get("discrete") if (self.relaxed and fluent) else None
        lit = TV(self.g.const(np.arange(K, dtype=np.float32)), (kpos,), None)
        if variant == "determinized":
            return total(self._mat(EI("MUL", lit, prob), isz))
        n = int(np.prod(isz, dtype=np.int64))
        gv = self.g.leaf("noise", n, "f", slot=len(self.sg.noise))
        self.sg.noise.append(NoiseSlot("gumbel", tuple(isz), n, gv, None))
        G = TV(gv, allv, None, True)
        if variant == "gumbel":
            w = TV(self.mparam(e, "discrete_softmax_weight", R["discrete_softmax_weight"]), ())
            eps = TV(self.mparam(e, "discrete_eps", R["discrete_eps"]), ())
            logits = self._mat(EI("ADD", G, EI("SAFELOG", prob, eps)), isz)
            return self._soft_argmax(TV(logits.v, allv, None, True), w, n_outer, isz, sizes)
        logits = self._mat(EI("ADD", G, EI("SAFELOG", prob, C(1e-12))), isz)
        dv, ptr, idx, _ = self._csr(logits, n_outer, isz)
        return TV(self.g.reduce("RARGMAX", logits.v, ptr, idx, "i"), dv, None)

    # ==================================================================================
    def build(self, action_values: Dict[str, V], state_values: Optional[Dict[str, V]] = None
              ) -> StepGraph:
        """Lower one transition (compiler.py:490-547) given the action values."""
        m, sg = self.m, self.sg
        env: Dict[str, V] = {}
        for name in m.state_fluents:
            if state_values is not None and name in state_values:
                v = state_values[name]
            else:
                v = self.g.leaf("state_in", m.variable_size(name),
                                self.dtype_of_range(m.variable_ranges[name]), fluent=name)
            sg.state_in[name] = v
            env[name] = v
            env[m.next_state[name]] = v          # compiler.py:721-722: s' starts equal to s
        for name, v in action_values.items():
            sg.action_in[name] = v
            env[name] = v
        for name in list(m.interm_fluents) + list(m.derived_fluents) + list(m.observ_fluents):
            pass
        for lv in m.levels.values():
            for name in lv:                       # compiler.py:360-368
                params, expr = m.cpfs[name]
                scope = list(params)
                sizes = self.sizes(scope)
                tv = self.lower(expr, scope, env)
                all_vars = tuple(range(len(scope)))
                full = TV(tv.v, tv.vars, tv.idx)
                want = self.dtype_of_range(m.variable_ranges[name])
                if full.dtype != want:               # _jax_cast, compiler.py:971-983
                    full = self._cast(full, want, sizes)
                v, mp = self.operand(full, all_vars, sizes)
                if mp is not None or v.op in ("state_in", "action_in", "const", "mparam", "noise") \
                        or v.const is not None:
                    v = self.g.add(V("MOV", [(v, mp)], n=self.n_of(all_vars, sizes),
                                     dtype=want, uniform=False))
                if self.relaxed and name in (self.relax.get("cpfs_without_grad") or ()):
                    v = self.g.sg(v)
                env[name] = v
                sg.fluents[name] = v
        sizes = ()
        r = self.lower(m.reward, [], env)
        r = self.to_f(self.at_least_int(r), sizes)
        rv = self.materialise(r, sizes)
        if rv.uniform or rv.const is not None:
            rv = self.g.add(V("MOV", [(rv, None)], n=1, dtype="f", uniform=False))
        sg.reward = rv
        for s, sp in m.next_state.items():
            sg.next_state[s] = env[sp]
        return sg

    def _cast(self, tv: TV, want: str, sizes) -> TV:
        src = tv.dtype
        if want == "f":
            return self.to_f(self.at_least_int(tv), sizes)
        if want == "i":
            if src == "b":
                return self.at_least_int(tv)
            out = self.ew("F2I", [tv], sizes, "i")          # float -> int truncates
            back = self.ew("I2F", [out], sizes)
            self.errchk("CAST", self.ew("NE", [back, tv], sizes), sizes)
            return out
        # want bool
        out = self.to_b(tv, sizes)
        if src == "f":
            back = self.ew("I2F", [TV(self._retag(out.v, "i"), out.vars)], sizes)
            self.errchk("CAST", self.ew("NE", [back, tv], sizes), sizes)
        else:
            self.errchk("CAST", self.ew("INE", [TV(self._retag(out.v, "i"), out.vars), tv], sizes), sizes)
        return out
