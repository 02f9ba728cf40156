"""SSA value graph of one rollout step, and reverse-mode differentiation over it.

This is the host-side middle layer between the expression trees (rddl/) and the flat
instruction stream executed on the GPU (core/program.py -> csrc/rk_interp.cu).  What the
reference gets from tracing closures through ``jax.vmap`` / ``lax.scan`` / ``value_and_grad``
(compiler.py:551-619, planner.py:2873) is done here explicitly:

* every value knows its element count per rollout (``n``) and whether it is rollout-invariant
  (``uniform``): the "lanes over object tuples" layout of the kernel;
* operands carry gather maps, so slices / transposes / broadcasts of pvariables
  (compiler.py:1071-1079) cost no instruction;
* ``differentiate`` emits the adjoint of every instruction as more instructions of the same
  ISA, honouring stop_gradient (the STE pattern ``soft + sg(hard - soft)``, logic.py:272).
"""
from __future__ import annotations

import itertools
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

Map = Union[None, str, np.ndarray]      # None = identity, 'b' = broadcast element 0, table


class V:
    """One SSA value: ``n`` 32-bit words per rollout (or in total when ``uniform``)."""
    _ids = itertools.count()
    __slots__ = ("id", "op", "args", "n", "dtype", "uniform", "const", "attrs")

    def __init__(self, op: str, args: Sequence[Tuple["V", Map]] = (), n: int = 1,
                 dtype: str = "f", uniform: bool = False, const: Optional[np.ndarray] = None,
                 attrs: Optional[Dict[str, Any]] = None):
        self.id = next(V._ids)
        self.op = op
        self.args = list(args)
        self.n = int(n)
        self.dtype = dtype
        self.uniform = uniform
        self.const = const
        self.attrs = attrs or {}

    def __repr__(self):
        return f"%{self.id}:{self.op}[{self.n}{'u' if self.uniform else ''}{self.dtype}]"


NP_DT = {"f": np.float32, "i": np.int32, "b": np.int32}

_FOLD1 = {
    "NEG": np.negative, "ABS": np.abs, "SGN": np.sign, "FLOOR": np.floor, "CEIL": np.ceil,
    "ROUND": np.round, "EXP": np.exp, "LOG": np.log, "SIN": np.sin, "COS": np.cos,
    "TAN": np.tan, "TANH": np.tanh, "SQUARE": np.square, "MOV": lambda x: x,
    "RCP": lambda x: np.float32(1.) / x, "INEG": np.negative, "IABS": np.abs,
    "I2F": lambda x: x.astype(np.float32), "NOT": lambda x: 1 - x,
    "F2B": lambda x: (x != 0).astype(np.int32), "I2B": lambda x: (x != 0).astype(np.int32),
}
_FOLD2 = {
    "ADD": np.add, "SUB": np.subtract, "MUL": np.multiply, "DIV": np.divide,
    "MIN": np.minimum, "MAX": np.maximum, "POW": np.power,
    "IADD": np.add, "ISUB": np.subtract, "IMUL": np.multiply, "IMIN": np.minimum,
    "IMAX": np.maximum, "AND": np.bitwise_and, "OR": np.bitwise_or, "XOR": np.bitwise_xor,
    "LT": lambda a, b: (a < b).astype(np.int32), "LE": lambda a, b: (a <= b).astype(np.int32),
    "GT": lambda a, b: (a > b).astype(np.int32), "GE": lambda a, b: (a >= b).astype(np.int32),
    "EQ": lambda a, b: (a == b).astype(np.int32), "NE": lambda a, b: (a != b).astype(np.int32),
    "ILT": lambda a, b: (a < b).astype(np.int32), "ILE": lambda a, b: (a <= b).astype(np.int32),
    "IGT": lambda a, b: (a > b).astype(np.int32), "IGE": lambda a, b: (a >= b).astype(np.int32),
    "IEQ": lambda a, b: (a == b).astype(np.int32), "INE": lambda a, b: (a != b).astype(np.int32),
}

OUT_DTYPE = {}
for _op in ("LT", "LE", "GT", "GE", "EQ", "NE", "ILT", "ILE", "IGT", "IGE", "IEQ", "INE",
            "AND", "OR", "XOR", "NOT", "F2B", "I2B"):
    OUT_DTYPE[_op] = "b"
for _op in ("INEG", "IABS", "ISGN", "IADD", "ISUB", "IMUL", "IMIN", "IMAX", "IFDIV", "IMOD",
            "IFMOD", "IPOW", "F2I", "RISUM", "RIPROD", "RIMIN", "RIMAX", "RARGMAX", "RARGMIN"):
    OUT_DTYPE[_op] = "i"


def apply_map(m: Map, n_out: int) -> np.ndarray:
    if m is None:
        return np.arange(n_out)
    if isinstance(m, str):
        return np.zeros(n_out, dtype=np.int64)
    return np.asarray(m, dtype=np.int64)


def norm_map(table: np.ndarray, n_src: int) -> Map:
    """Canonicalise a gather table."""
    table = np.asarray(table, dtype=np.int64).ravel()
    if table.size == n_src and np.array_equal(table, np.arange(n_src)):
        return None
    if np.all(table == 0):
        return "b"
    return table


class Graph:
    """Append-only list of SSA values (creation order is a topological order)."""

    def __init__(self):
        self.nodes: List[V] = []
        self._consts: Dict[Tuple, V] = {}
        self._cse: Dict[Tuple, V] = {}      # value numbering: identical pure ops share a value

    @staticmethod
    def _map_key(m: Map):
        if m is None or isinstance(m, str):
            return m
        return np.asarray(m, dtype=np.int64).tobytes()

    def _numbered(self, key: Tuple, make) -> V:
        v = self._cse.get(key)
        if v is None:
            v = self.add(make())
            self._cse[key] = v
        return v

    def add(self, v: V) -> V:
        self.nodes.append(v)
        return v

    # ---- leaves ------------------------------------------------------------------
    def const(self, value, dtype: Optional[str] = None) -> V:
        arr = np.asarray(value)
        if dtype is None:
            dtype = "f" if arr.dtype.kind == "f" else ("b" if arr.dtype == np.bool_ else "i")
        arr = np.ascontiguousarray(arr.astype(NP_DT[dtype])).ravel()
        key = (dtype, arr.tobytes())
        v = self._consts.get(key)
        if v is None:
            v = self.add(V("const", n=arr.size, dtype=dtype, uniform=True, const=arr))
            self._consts[key] = v
        return v

    def leaf(self, op: str, n: int, dtype: str = "f", uniform: bool = False, **attrs) -> V:
        return self.add(V(op, n=n, dtype=dtype, uniform=uniform, attrs=attrs))

    # ---- operations ----------------------------------------------------------------
    def op(self, name: str, args: Sequence[Tuple[V, Map]], n: int, dtype: Optional[str] = None,
           **attrs) -> V:
        dtype = dtype or OUT_DTYPE.get(name, "f")
        # constant folding of the simple elementwise ops
        if all(a.const is not None for a, _ in args) and not attrs:
            vals = [a.const[apply_map(m, n)] for a, m in args]
            fn = _FOLD1.get(name) if len(vals) == 1 else _FOLD2.get(name)
            if fn is not None:
                with np.errstate(all="ignore"):
                    return self.const(np.asarray(fn(*vals)).astype(NP_DT[dtype]), dtype)
        uniform = all(a.uniform for a, _ in args)
        simp = self._simplify(name, args, n, dtype)
        if simp is not None:
            return simp
        if attrs:
            return self.add(V(name, args, n=n, dtype=dtype, uniform=uniform, attrs=attrs))
        key = (name, dtype, n, tuple((a.id, self._map_key(m)) for a, m in args))
        return self._numbered(key, lambda: V(name, args, n=n, dtype=dtype, uniform=uniform))

    def _simplify(self, name: str, args, n: int, dtype: str) -> Optional[V]:
        """Value-preserving rewrites (bit-identical results, so parity is unaffected):
        pow(x, 2) -> x*x, pow(x, 1) -> x (what XLA's algebraic simplifier and torch.pow do
        as well), x*1 -> x, x/1 -> x, x-0 -> x, -(-x) -> x."""
        def scalar_const(a: V):
            if a.const is None or a.const.size == 0:
                return None
            c = a.const
            return float(c[0]) if np.all(c == c[0]) else None

        def same_space(a: V, m: Map) -> bool:
            return m is None and a.n == n and a.dtype == dtype

        if name == "POW" and dtype == "f" and args[0][0].dtype == "f":
            y = scalar_const(args[1][0])
            if y == 2.0:
                return self.op("SQUARE", [args[0]], n)
            if y == 1.0 and same_space(*args[0]):
                return args[0][0]
        if name in ("MUL", "DIV") and dtype == "f":
            if scalar_const(args[1][0]) == 1.0 and same_space(*args[0]):
                return args[0][0]
            if name == "MUL" and scalar_const(args[0][0]) == 1.0 and same_space(*args[1]):
                return args[1][0]
        if name == "SUB" and dtype == "f":
            y = scalar_const(args[1][0])
            if y == 0.0 and not np.signbit(args[1][0].const[0]) and same_space(*args[0]):
                return args[0][0]
        if name == "NEG" and args[0][0].op == "NEG" and same_space(*args[0]) \
                and same_space(*args[0][0].args[0]):
            return args[0][0].args[0][0]
        return None

    def sg(self, a: V) -> V:
        """stop_gradient: aliases the operand's storage, blocks differentiation."""
        if a.const is not None:
            return a
        return self._numbered(("sg", a.id), lambda: V("sg", [(a, None)], n=a.n, dtype=a.dtype,
                                                      uniform=a.uniform))

    def reduce(self, name: str, a: V, ptr: np.ndarray, idx: np.ndarray,
               dtype: Optional[str] = None) -> V:
        n = len(ptr) - 1
        ptr, idx = np.asarray(ptr, np.int64), np.asarray(idx, np.int64)
        dtype = dtype or OUT_DTYPE.get(name, "f")
        key = (name, dtype, a.id, ptr.tobytes(), idx.tobytes())
        return self._numbered(key, lambda: V(name, [(a, None)], n=n, dtype=dtype,
                                             uniform=a.uniform, attrs={"ptr": ptr, "idx": idx}))


# =====================================================================================
# reverse-mode differentiation
# =====================================================================================

class NotDifferentiable(Exception):
    pass


def _active_set(nodes: List[V], wrt: Sequence[V]) -> set:
    active = {v.id for v in wrt}
    for v in nodes:
        if v.id in active:
            continue
        if v.op in ("sg", "const") or v.dtype != "f":
            continue
        if any(a.id in active for a, _ in v.args):
            active.add(v.id)
    return active


def differentiate(g: Graph, nodes: List[V], seeds: Sequence[Tuple[V, V]],
                  wrt: Sequence[V]) -> Dict[int, V]:
    """Emit adjoint instructions into ``g``.

    nodes : forward values in topological order
    seeds : [(value, cotangent V of the same n / uniform-ness), ...]
    wrt   : values whose cotangents are wanted
    Returns {value id: cotangent V} for every ``wrt`` value that receives any gradient.
    """
    active = _active_set(nodes, wrt)
    contrib: Dict[int, List[V]] = {}
    for tgt, s in seeds:
        if tgt.uniform and not s.uniform:
            # a rollout-invariant value (e.g. a next-state fluent that depends on the plan
            # only) seeded with a per-rollout cotangent: sum over the rollouts first, exactly
            # as push() does for operands, or later QSUMs would count uniform parts twice
            s = g.add(V("QSUM", [(s, None)], n=tgt.n, dtype="f", uniform=True))
        contrib.setdefault(tgt.id, []).append(s)
    F = lambda x: g.const(np.float32(x))     # noqa: E731

    def total(v: V) -> Optional[V]:
        lst = contrib.get(v.id)
        if not lst:
            return None
        acc = lst[0]
        for c in lst[1:]:
            acc = g.op("ADD", [(acc, None), (c, None)], v.n)
        contrib[v.id] = [acc]
        return acc

    def push(v: V, arg_index: int, ct: V):
        """Send cotangent ``ct`` (in v's element space) back through operand map."""
        a, m = v.args[arg_index]
        if a.id not in active:
            return
        n_r = v.n
        if not (m is None and a.n == n_r):
            table = apply_map(m, n_r)
            order = np.argsort(table, kind="stable")
            counts = np.bincount(table, minlength=a.n)
            if a.n == n_r and np.all(counts == 1):
                ct = g.op("MOV", [(ct, order.astype(np.int64))], a.n)   # inverse permutation
            else:
                ptr = np.concatenate([[0], np.cumsum(counts)])
                ct = g.reduce("RSUM", ct, ptr, order)
        if a.uniform and not ct.uniform:
            ct = g.add(V("QSUM", [(ct, None)], n=a.n, dtype="f", uniform=True))
        contrib.setdefault(a.id, []).append(ct)

    wrt_ids = {v.id for v in wrt}
    for v in reversed(nodes):
        if v.id not in active or v.id in wrt_ids and not v.args:
            continue
        if not v.args:
            continue
        ct = total(v)
        if ct is None:
            continue
        n = v.n
        op = v.op
        A = v.args
        sam = lambda i: (A[i][0], A[i][1])     # operand i as seen in v's space  # noqa: E731
        me = (v, None)
        c = (ct, None)

        def mul(x, y):
            return g.op("MUL", [x, y], n)

        if op in ("MOV", "sg_passthrough"):
            push(v, 0, ct)
        elif op == "NEG":
            push(v, 0, g.op("NEG", [c], n))
        elif op == "ABS":
            push(v, 0, mul(c, (g.op("SGN", [sam(0)], n), None)))
        elif op in ("SGN", "FLOOR", "CEIL", "ROUND", "FDIV"):
            pass
        elif op == "SQRT":     # softjax sj.sqrt: zero derivative at x <= 0
            half_over = g.op("DIV", [(F(0.5), "b"), me], n)
            d = mul(c, (half_over, None))
            pos = g.op("GT", [sam(0), (F(0.), "b")], n)
            push(v, 0, g.op("SELECT", [(pos, None), (d, None), (F(0.), "b")], n))
        elif op == "EXP":
            push(v, 0, mul(c, me))
        elif op == "EXPM1":
            push(v, 0, mul(c, (g.op("ADD", [me, (F(1.), "b")], n), None)))
        elif op == "LOG":
            push(v, 0, g.op("DIV", [c, sam(0)], n))
        elif op == "LOG1P":
            push(v, 0, g.op("DIV", [c, (g.op("ADD", [sam(0), (F(1.), "b")], n), None)], n))
        elif op == "SIN":
            push(v, 0, mul(c, (g.op("COS", [sam(0)], n), None)))
        elif op == "COS":
            push(v, 0, g.op("NEG", [(mul(c, (g.op("SIN", [sam(0)], n), None)), None)], n))
        elif op == "TAN":
            sq = g.op("SQUARE", [me], n)
            push(v, 0, mul(c, (g.op("ADD", [(sq, None), (F(1.), "b")], n), None)))
        elif op in ("ASIN", "ACOS"):
            sq = g.op("SQUARE", [sam(0)], n)
            den = g.op("SQRT", [(g.op("SUB", [(F(1.), "b"), (sq, None)], n), None)], n)
            d = g.op("DIV", [c, (den, None)], n)
            push(v, 0, d if op == "ASIN" else g.op("NEG", [(d, None)], n))
        elif op == "ATAN":
            sq = g.op("SQUARE", [sam(0)], n)
            push(v, 0, g.op("DIV", [c, (g.op("ADD", [(sq, None), (F(1.), "b")], n), None)], n))
        elif op == "SINH":
            push(v, 0, mul(c, (g.op("COSH", [sam(0)], n), None)))
        elif op == "COSH":
            push(v, 0, mul(c, (g.op("SINH", [sam(0)], n), None)))
        elif op in ("TANH", "STANH"):
            sq = g.op("SQUARE", [me], n)
            push(v, 0, mul(c, (g.op("SUB", [(F(1.), "b"), (sq, None)], n), None)))
        elif op == "SIGMOID":
            om = g.op("SUB", [(F(1.), "b"), me], n)
            push(v, 0, mul(c, (mul(me, (om, None)), None)))
        elif op == "RELU":
            pos = g.op("GT", [sam(0), (F(0.), "b")], n)
            push(v, 0, g.op("SELECT", [(pos, None), c, (F(0.), "b")], n))
        elif op == "RCP":
            push(v, 0, g.op("NEG", [(mul(c, (g.op("SQUARE", [me], n), None)), None)], n))
        elif op == "SQUARE":
            push(v, 0, mul(c, (mul((F(2.), "b"), sam(0)), None)))
        elif op == "ADD":
            push(v, 0, ct)
            push(v, 1, ct)
        elif op == "SUB":
            push(v, 0, ct)
            if A[1][0].id in active:
                push(v, 1, g.op("NEG", [c], n))
        elif op == "MUL":
            if A[0][0].id in active:
                push(v, 0, mul(c, sam(1)))
            if A[1][0].id in active:
                push(v, 1, mul(c, sam(0)))
        elif op == "DIV":
            if A[0][0].id in active:
                push(v, 0, g.op("DIV", [c, sam(1)], n))
            if A[1][0].id in active:
                q = g.op("DIV", [(mul(c, me), None), sam(1)], n)
                push(v, 1, g.op("NEG", [(q, None)], n))
        elif op in ("MIN", "MAX"):
            w = g.op("MINW" if op == "MIN" else "MAXW", [sam(0), sam(1)], n)
            gx = mul(c, (w, None))
            push(v, 0, gx)
            if A[1][0].id in active:
                push(v, 1, g.op("SUB", [c, (gx, None)], n))
        elif op == "POW":
            if A[0][0].id in active:
                ym1 = g.op("SUB", [sam(1), (F(1.), "b")], n)
                p = g.op("POW", [sam(0), (ym1, None)], n)
                push(v, 0, mul(c, (mul(sam(1), (p, None)), None)))
            if A[1][0].id in active:
                lg = g.op("LOG", [sam(0)], n)
                push(v, 1, mul(c, (mul(me, (lg, None)), None)))
        elif op == "HYPOT":
            for i in (0, 1):
                if A[i][0].id in active:
                    push(v, i, mul(c, (g.op("DIV", [sam(i), me], n), None)))
        elif op == "MOD":
            push(v, 0, ct)
            if A[1][0].id in active:
                fd = g.op("FDIV", [sam(0), sam(1)], n)
                push(v, 1, g.op("NEG", [(mul(c, (fd, None)), None)], n))
        elif op == "FMOD":
            push(v, 0, ct)
            if A[1][0].id in active:
                raise NotDifferentiable("fmod with a differentiable divisor")
        elif op == "SAFELOG":   # compiler.py:70-75: tangent xdot / (x + eps)
            den = g.op("ADD", [sam(0), sam(1)], n)
            push(v, 0, g.op("DIV", [c, (den, None)], n))
        elif op == "SELECT":
            z = (F(0.), "b")
            if A[1][0].id in active:
                push(v, 1, g.op("SELECT", [sam(0), c, z], n))
            if A[2][0].id in active:
                push(v, 2, g.op("SELECT", [sam(0), z, c], n))
        elif op == "LERP":      # a*b + (1-a)*c
            if A[0][0].id in active:
                push(v, 0, mul(c, (g.op("SUB", [sam(1), sam(2)], n), None)))
            if A[1][0].id in active:
                push(v, 1, mul(c, sam(0)))
            if A[2][0].id in active:
                push(v, 2, mul(c, (g.op("SUB", [(F(1.), "b"), sam(0)], n), None)))
        elif op == "FMA":
            if A[0][0].id in active:
                push(v, 0, mul(c, sam(1)))
            if A[1][0].id in active:
                push(v, 1, mul(c, sam(0)))
            push(v, 2, ct)
        elif op == "RSUM":
            src = A[0][0]
            ptr, idx = v.attrs["ptr"], v.attrs["idx"]
            seg = np.repeat(np.arange(n), np.diff(ptr))          # segment of each CSR entry
            counts = np.bincount(idx, minlength=src.n)
            if np.all(counts == 1):
                table = np.empty(src.n, dtype=np.int64)
                table[idx] = seg
                back = g.op("MOV", [(ct, table)], src.n)
            else:
                order = np.argsort(idx, kind="stable")
                ptr2 = np.concatenate([[0], np.cumsum(counts)])
                back = g.reduce("RSUM", ct, ptr2, seg[order])
            contrib.setdefault(src.id, []).append(back) if src.id in active else None
        elif op == "RPROD":
            src = A[0][0]
            ptr, idx = v.attrs["ptr"], v.attrs["idx"]
            seg = np.repeat(np.arange(n), np.diff(ptr))
            if not np.all(np.bincount(idx, minlength=src.n) == 1):
                raise NotDifferentiable("product over a broadcast operand")
            table = np.empty(src.n, dtype=np.int64)
            table[idx] = seg
            others = g.add(V("RPRODX", [(src, None)], n=src.n, dtype="f", uniform=src.uniform,
                             attrs={"ptr": ptr, "idx": idx, "seg": table}))
            back = g.op("MUL", [(ct, table), (others, None)], src.n)
            contrib.setdefault(src.id, []).append(back) if src.id in active else None
        elif op in ("RMIN", "RMAX"):   # ties share the cotangent evenly (JAX reduce_max rule)
            src = A[0][0]
            ptr, idx = v.attrs["ptr"], v.attrs["idx"]
            seg = np.repeat(np.arange(n), np.diff(ptr))
            if not np.all(np.bincount(idx, minlength=src.n) == 1):
                raise NotDifferentiable("min/max over a broadcast operand")
            table = np.empty(src.n, dtype=np.int64)
            table[idx] = seg
            eq = g.op("EQ", [(src, None), (v, table)], src.n)
            cnt = g.reduce("RSUM", g.op("I2F", [(eq, None)], src.n), ptr, idx)
            share = g.op("DIV", [c, (cnt, None)], n)
            back = g.op("SELECT", [(eq, None), (share, table), (F(0.), "b")], src.n)
            contrib.setdefault(src.id, []).append(back) if src.id in active else None
        elif op == "LGAMMA":   # d/dx ln Gamma(x) = digamma(x)  (jax.scipy.special.gammaln's JVP)
            push(v, 0, mul(c, (g.op("DIGAMMA", [sam(0)], n), None)))
        else:
            raise NotDifferentiable(f"no adjoint rule for op {op}")

    out = {}
    for w in wrt:
        t = total(w)
        if t is not None:
            out[w.id] = t
    return out


# =====================================================================================
# sparsification: compile-time known elements, identity pruning, element compaction
# =====================================================================================
# The reference evaluates aggregations over full object-tuple spaces, e.g. Wildfire's
# exists / sum over (x2, y2) of NEIGHBOR(x, y, x2, y2) ^ burning(x2, y2): 625 terms per step of
# which ~90% are multiplied by a non-fluent 0 (logic.py:454, compiler.py:1316).  Because the
# mask is known when the program is built, those terms are exactly 0 (sum) or 1 (product) and
# can be dropped without changing any result bit (for finite fluents), and the wide
# intermediate vectors shrink to the non-zero pattern.

_LEAF_OPS = {"const", "mparam", "state_in", "persist", "param_in", "act_in", "noise", "disc_in"}
_RED_OPS = {"RSUM", "RPROD", "RMIN", "RMAX", "RISUM", "RIPROD", "RIMIN", "RIMAX", "RARGMAX",
            "RARGMIN"}
_RED_FOLD = {"RSUM": np.sum, "RISUM": np.sum, "RPROD": np.prod, "RIPROD": np.prod,
             "RMIN": np.min, "RIMIN": np.min, "RMAX": np.max, "RIMAX": np.max}
_RED_IDENTITY = {"RSUM": 0.0, "RISUM": 0.0, "RPROD": 1.0, "RIPROD": 1.0}


def _reachable(g: Graph, roots: Sequence[V]) -> List[V]:
    seen, stack = set(), list(roots)
    while stack:
        v = stack.pop()
        if v.id in seen:
            continue
        seen.add(v.id)
        for a, _ in v.args:
            stack.append(a)
    return [v for v in g.nodes if v.id in seen]


def sparsify(g: Graph, roots: Sequence[V]) -> Dict[str, int]:
    """Rewrite, IN PLACE, the sub-graph under ``roots`` (see the section comment).  Roots and
    leaves keep their element spaces; interior values may shrink.  Returns statistics."""
    nodes = _reachable(g, roots)
    root_ids = {r.id for r in roots}
    stats = {"dropped_members": 0, "elements_before": 0, "elements_after": 0, "to_const": 0}

    def cast(vals, dtype):
        return np.asarray(vals).astype(NP_DT[dtype]).astype(np.float64)

    # ---- phase 1: known elements (forward), pruning identity members of sums / products -----
    KM: Dict[int, np.ndarray] = {}
    KV: Dict[int, np.ndarray] = {}
    for v in nodes:
        n = v.n
        if v.const is not None:
            KM[v.id], KV[v.id] = np.ones(n, bool), v.const.astype(np.float64)
            continue
        mask, vals = np.zeros(n, bool), np.zeros(n)
        if v.op in _LEAF_OPS or not v.args or v.op in ("RPRODX", "QSUM"):
            KM[v.id], KV[v.id] = mask, vals
            continue
        if v.op == "sg":
            a = v.args[0][0]
            KM[v.id], KV[v.id] = KM[a.id], KV[a.id]
            continue
        if v.op in _RED_OPS:
            src = v.args[0][0]
            km, kv = KM[src.id], KV[src.id]
            ptr, idx = v.attrs["ptr"], v.attrs["idx"]
            ident = _RED_IDENTITY.get(v.op)
            if v.op == "RIMAX" and src.dtype == "b":
                ident = 0.0
            if v.op == "RIMIN" and src.dtype == "b":
                ident = 1.0
            new_ptr, new_idx = [0], []
            for e in range(n):
                mem = idx[ptr[e]:ptr[e + 1]]
                if ident is not None and len(mem):
                    keep = ~(km[mem] & (kv[mem] == ident))
                    stats["dropped_members"] += int((~keep).sum())
                    mem = mem[keep]
                new_idx.append(mem)
                new_ptr.append(new_ptr[-1] + len(mem))
                if v.op in _RED_FOLD and (len(mem) == 0 or km[mem].all()):
                    if len(mem) == 0 and ident is None:
                        continue
                    mask[e] = True
                    vals[e] = ident if len(mem) == 0 else float(
                        cast(_RED_FOLD[v.op](cast(kv[mem], v.dtype)), v.dtype))
            v.attrs["ptr"] = np.asarray(new_ptr, np.int64)
            v.attrs["idx"] = (np.concatenate(new_idx) if new_idx else np.zeros(0)).astype(np.int64)
            KM[v.id], KV[v.id] = mask, vals
            continue
        # elementwise
        tabs = [apply_map(m, n) for _, m in v.args]
        ms = [KM[a.id][t] for (a, _), t in zip(v.args, tabs)]
        vs = [KV[a.id][t] for (a, _), t in zip(v.args, tabs)]
        allk = np.logical_and.reduce(ms)
        fn = (_FOLD1.get(v.op) if len(v.args) == 1 else _FOLD2.get(v.op)) if len(v.args) <= 2 else None
        if fn is not None and allk.any():
            with np.errstate(all="ignore"):
                ins = [cast(x, a.dtype).astype(NP_DT[a.dtype]) for x, (a, _) in zip(vs, v.args)]
                out = np.asarray(fn(*ins)).astype(NP_DT[v.dtype]).astype(np.float64)
            mask |= allk
            vals = np.where(allk, out, vals)
        if v.op in ("MUL", "IMUL", "AND"):          # annihilator (finite operands assumed)
            for m_, x_ in zip(ms, vs):
                z = m_ & (x_ == 0)
                mask |= z
                vals = np.where(z, 0.0, vals)
        if v.op == "OR":
            for m_, x_ in zip(ms, vs):
                z = m_ & (x_ != 0)
                mask |= z
                vals = np.where(z, 1.0, vals)
        KM[v.id], KV[v.id] = mask, vals

    # ---- phase 2: demand (backward) ------------------------------------------------------------
    need: Dict[int, np.ndarray] = {v.id: np.zeros(v.n, bool) for v in nodes}
    for r in roots:
        need[r.id][:] = True
        x = r
        while x.op == "sg":                            # a root alias pins its operand too
            x = x.args[0][0]
            need[x.id][:] = True
            root_ids.add(x.id)
    for v in reversed(nodes):
        nd = need[v.id]
        if not nd.any() or v.op in _LEAF_OPS or not v.args:
            continue
        if v.id not in root_ids and KM[v.id][nd].all():
            continue                                  # becomes a constant: needs no operand
        if v.op in _RED_OPS:
            src = v.args[0][0]
            ptr, idx = v.attrs["ptr"], v.attrs["idx"]
            for e in np.nonzero(nd)[0]:
                need[src.id][idx[ptr[e]:ptr[e + 1]]] = True
        elif v.op == "RPRODX":
            need[v.args[0][0].id][:] = True            # opaque: keeps its operand whole
            need[v.id][:] = True
        else:
            for a, m in v.args:
                need[a.id][apply_map(m, v.n)[nd]] = True

    # ---- phase 3: rebuild in place (forward) ----------------------------------------------------
    pos: Dict[int, np.ndarray] = {}
    for v in nodes:
        n = v.n
        nd = need[v.id]
        stats["elements_before"] += n if v.op not in _LEAF_OPS else 0
        if v.op in _LEAF_OPS or not v.args or v.id in root_ids or v.op == "RPRODX":
            pos[v.id] = np.arange(n)
            keep_all = True
        else:
            keep_all = bool(nd.all())
            pos[v.id] = np.arange(n) if keep_all else \
                np.where(nd, np.cumsum(nd) - 1, -1)
        if v.op in _LEAF_OPS or not v.args:
            continue
        if v.op == "sg":
            # stop_gradient aliases its operand's storage (program.py:root): it keeps exactly the
            # elements its operand keeps, whatever its own consumers asked for
            a = v.args[0][0]
            pos[v.id] = pos[a.id]
            v.n = a.n
            v.args = [(a, None)]
            if a.const is not None:
                v.op, v.args, v.const, v.uniform = "const", [], a.const.astype(NP_DT[v.dtype]), True
            continue
        if not nd.any():
            continue                                   # dead: no instruction will be emitted
        sel = np.arange(n) if keep_all else np.nonzero(nd)[0]
        if v.id not in root_ids and v.op != "RPRODX" and KM[v.id][sel].all():
            v.op, v.args, v.attrs = "const", [], {}
            v.const = KV[v.id][sel].astype(NP_DT[v.dtype])
            v.n, v.uniform = len(sel), True
            stats["to_const"] += 1
            stats["elements_after"] += v.n
            continue
        if v.op in _RED_OPS:
            src = v.args[0][0]
            ptr, idx = v.attrs["ptr"], v.attrs["idx"]
            new_ptr, new_idx = [0], []
            for e in sel:
                mem = pos[src.id][idx[ptr[e]:ptr[e + 1]]]
                assert (mem >= 0).all()
                new_idx.append(mem)
                new_ptr.append(new_ptr[-1] + len(mem))
            v.attrs["ptr"] = np.asarray(new_ptr, np.int64)
            v.attrs["idx"] = (np.concatenate(new_idx) if new_idx else np.zeros(0)).astype(np.int64)
        elif v.op == "RPRODX":
            src = v.args[0][0]
            assert (pos[src.id] == np.arange(src.n)).all()
        else:
            new_args = []
            for a, m in v.args:
                tab = pos[a.id][apply_map(m, n)[sel]]
                assert (tab >= 0).all(), (v, a)
                new_args.append((a, norm_map(tab, a.n if a.op in _LEAF_OPS or a.const is not None
                                             else int((pos[a.id] >= 0).sum()))))
            v.args = new_args
        v.n = len(sel)
        stats["elements_after"] += v.n
    # value numbering: re-key the surviving values with their new element spaces, so that the
    # adjoint rules emitted afterwards (cos x for d sin x, ...) still find the forward values
    g._cse.clear()
    for v in g.nodes:
        if v.const is not None or v.op in _LEAF_OPS or not v.args or v.op in ("QSUM", "RPRODX"):
            continue
        if v.op == "sg":
            if not v.attrs:
                g._cse.setdefault(("sg", v.args[0][0].id), v)
            continue
        if v.op in _RED_OPS:
            key = (v.op, v.dtype, v.args[0][0].id, v.attrs["ptr"].tobytes(), v.attrs["idx"].tobytes())
        elif v.attrs:
            continue
        else:
            key = (v.op, v.dtype, v.n, tuple((a.id, g._map_key(m)) for a, m in v.args))
        g._cse.setdefault(key, v)
    return stats
