"""Per-program kernel specialisation: instruction stream -> straight-line CUDA C.

The interpreter (csrc/rk_interp.cu) spends ~128 SASS instructions per interpreted instruction
on fetch / decode / dispatch / address arithmetic (profiles/r01_ncu_interp_first.md).  For a
given program all of that is known ahead of time, so this module unrolls the instruction
stream into one kernel per program:

* no fetch, decode or dispatch: every instruction becomes one C statement with its
  element count, addresses and gather tables baked in;
* values that only ever take part in single-pass instructions (n <= lanes per rollout) live
  in *registers* -- lane (q, e) holds element e of rollout q -- and are exchanged with
  ``__shfl_sync`` for gathers, aggregations and cross-rollout sums; only multi-pass values
  stay in the shared-memory register file of the interpreter layout;
* nvcc then schedules independent CPFs against each other (ILP), which the interpreter,
  executing one instruction at a time per warp, cannot.

The generated kernel has the interpreter's exact semantics (same program, same layouts, same
``RkLaunch`` argument block) and sits behind the same C ABI (rk_program_attach_cubin), so every
parity test runs against both back-ends.  This is the role XLA's jit plays for the reference
(planner.py:2687, 2930), done per RDDL domain/instance instead of per Python trace.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
from typing import Any, Dict, List, Tuple

import numpy as np

from . import isa
from .program import Program

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc")
JIT_DIR = os.path.join(CSRC, "jit")
KERNEL_NAME = "rk_spec_kernel"
CODEGEN_VERSION = 19

_F1 = {
    "NEG": "-x", "ABS": "fabsf(x)", "SGN": "sgnf(x)", "FLOOR": "floorf(x)", "CEIL": "ceilf(x)",
    "ROUND": "rintf(x)", "SQRT": "sqrtf(x)", "EXP": "expf(x)", "LOG": "logf(x)", "SIN": "sinf(x)",
    "COS": "cosf(x)", "TAN": "tanf(x)", "ASIN": "asinf(x)", "ACOS": "acosf(x)",
    "ATAN": "atanf(x)", "SINH": "sinhf(x)", "COSH": "coshf(x)", "TANH": "tanhf(x)",
    "LGAMMA": "lgammaf(x)", "SIGMOID": "(1.f / (1.f + expf(-x)))", "STANH": "stable_tanh(x)",
    "RELU": "fmaxf(x, 0.f)", "RCP": "(1.f / x)", "SQUARE": "(x * x)", "LOG1P": "log1pf(x)",
    "EXPM1": "expm1f(x)", "DIGAMMA": "rk_digamma(x)",
}
_F2 = {
    "ADD": "(x + y)", "SUB": "(x - y)", "MUL": "(x * y)", "DIV": "(x / y)", "MIN": "fminf(x, y)",
    "MAX": "fmaxf(x, y)", "POW": "powf(x, y)", "HYPOT": "hypotf(x, y)", "FMOD": "fmodf(x, y)",
    "MOD": "py_modf(x, y)", "FDIV": "floorf(x / y)",
    "SAFELOG": "((x > 0.f) ? logf(x) : -CUDART_INF_F + 0.f * y)",
    "MINW": "((x < y) ? 1.f : ((x == y) ? 0.5f : 0.f))",
    "MAXW": "((x > y) ? 1.f : ((x == y) ? 0.5f : 0.f))",
}
# relaxed programs: branch-free division / sqrt / trigonometry (rk_device.cuh), so that a whole
# step is one basic block the scheduler can interleave freely
_FAST = {
    "DIV": "rk_fdiv(x, y)", "RCP": "rk_fdiv(1.f, x)", "SQRT": "rk_fsqrt(x)",
    "SIGMOID": "rk_fdiv(1.f, 1.f + expf(-x))", "STANH": "rk_fstable_tanh(x)",
    "SIN": "rk_fsin(x)", "COS": "rk_fcos(x)", "TAN": "rk_ftan(x)",
}
_F2B = {"LT": "<", "LE": "<=", "GT": ">", "GE": ">=", "EQ": "==", "NE": "!="}
_I1 = {"INEG": "(-x)", "IABS": "abs(x)", "ISGN": "((x > 0) - (x < 0))", "NOT": "(x == 0)",
       "I2B": "(x != 0)"}
_I2 = {"IADD": "(x + y)", "ISUB": "(x - y)", "IMUL": "(x * y)", "IMIN": "min(x, y)",
       "IMAX": "max(x, y)", "IFDIV": "ifloordiv(x, y)", "IMOD": "ipymod(x, y)",
       "IFMOD": "((y == 0) ? 0 : (x % y))", "IPOW": "ipow(x, y)", "ILT": "(x < y)",
       "ILE": "(x <= y)", "IGT": "(x > y)", "IGE": "(x >= y)", "IEQ": "(x == y)",
       "INE": "(x != y)", "AND": "(x & y)", "OR": "(x | y)", "XOR": "(x ^ y)"}
_RED = {  # op -> (ctype, init, update)
    "RSUM": ("float", "0.f", "acc = acc + t"), "RPROD": ("float", "1.f", "acc = acc * t"),
    "RMIN": ("float", "CUDART_INF_F", "acc = fminf(acc, t)"),
    "RMAX": ("float", "-CUDART_INF_F", "acc = fmaxf(acc, t)"),
    "RISUM": ("int", "0", "acc = acc + t"), "RIPROD": ("int", "1", "acc = acc * t"),
    "RIMIN": ("int", "2147483647", "acc = min(acc, t)"),
    "RIMAX": ("int", "(-2147483647 - 1)", "acc = max(acc, t)"),
}


def _flit(x: float) -> str:
    x = np.float32(x)
    if np.isnan(x):
        return "CUDART_NAN_F"
    if np.isinf(x):
        return "CUDART_INF_F" if x > 0 else "(-CUDART_INF_F)"
    return f"__uint_as_float(0x{np.float32(x).view(np.uint32):08x}u)"


def _edge_color(fa: List[int], fb: List[int]) -> List[int]:
    """Proper edge colouring of the bipartite multigraph with one edge (fa[e], fb[e]) per element
    (fb[e] < 0: no second endpoint): no two edges at a node share a colour, max-degree colours
    (Koenig), by flipping two-coloured alternating paths."""
    E = len(fa)
    colA: Dict[int, Dict[int, int]] = {}
    colB: Dict[int, Dict[int, int]] = {}
    color = [-1] * E
    fb = [v if v >= 0 else -(e + 1) for e, v in enumerate(fb)]
    deg: Dict[Any, int] = {}
    for e in range(E):
        deg[("a", fa[e])] = deg.get(("a", fa[e]), 0) + 1
        deg[("b", fb[e])] = deg.get(("b", fb[e]), 0) + 1
    delta = max(deg.values()) if deg else 0

    def free(used: Dict[int, int]) -> int:
        return next(c for c in range(delta) if c not in used)

    for e in range(E):
        u, v = fa[e], fb[e]
        ua, vb = colA.setdefault(u, {}), colB.setdefault(v, {})
        a, b = free(ua), free(vb)
        if a != b and a in vb:
            # colour a is taken at v: flip the a / b alternating path that starts at v (it cannot
            # reach u, where a is free), after which a is free at v
            path, node, side, c = [], v, "b", a
            while True:
                used = colB[node] if side == "b" else colA[node]
                if c not in used:
                    break
                edge = used[c]
                path.append(edge)
                node = fa[edge] if side == "b" else fb[edge]
                side = "a" if side == "b" else "b"
                c = b if c == a else a
            for edge in path:
                del colA[fa[edge]][color[edge]]
                del colB[fb[edge]][color[edge]]
            for edge in path:
                color[edge] = b if color[edge] == a else a
                colA[fa[edge]][color[edge]] = edge
                colB[fb[edge]][color[edge]] = edge
        color[e] = a
        colA[u][a] = e
        colB[v][a] = e
    return color


class _EllFamily:
    """Register layout of a group of wide values (n > lanes per rollout) that are only combined
    element by element, filled by gathers from narrow values and consumed by segment reductions:
    element e lives in register ``slot[e]`` of lane ``lane[e]`` -- its segment under the family's first
    segmentation -- so those reductions are register adds; a second segmentation (the transposed
    reduction of an adjoint) sees at most one member per slot in every segment (edge colouring), i.e.
    one shuffle per slot.  ELL-style storage of the sparse (lane x lane) incidence the model's object
    relations define (Wildfire: cell x neighbour, 144 of 625 pairs, 8 slots)."""

    def __init__(self, fid: int, n: int, segA: List[int], segB, L: int):
        self.fid, self.n, self.L = fid, n, L
        self.lane = list(segA)
        self.segB = None if segB is None else list(segB)
        self.slot = _edge_color(self.lane, self.segB if self.segB is not None else [-1] * n)
        self.K = max(self.slot) + 1
        self.elem_at = -np.ones((L, self.K), dtype=np.int64)
        for e in range(n):
            self.elem_at[self.lane[e], self.slot[e]] = e
        self.elem_b = -np.ones((L, self.K), dtype=np.int64)
        if self.segB is not None:
            for e in range(n):
                if self.segB[e] >= 0:
                    assert self.elem_b[self.segB[e], self.slot[e]] < 0
                    self.elem_b[self.segB[e], self.slot[e]] = e


class _Gen:
    def __init__(self, prog: Program):
        self.p = prog
        self.ins = prog.meta["ins"]
        self.L, self.rpw = prog.lanes, prog.rpw
        self.n_pro, self.n_step, self.n_epi = prog.n_ins
        self.persistent = set(prog.meta["persistent"])
        self.err_base = prog.meta["err_base"]
        self._classify()
        # relaxed programs may trade last-bit identity for speed (reciprocal of constants, FMA
        # contraction); exact programs never do.  RK_FAST_RELAXED=0 switches it off.
        self.fast = bool(prog.relaxed) and os.environ.get("RK_FAST_RELAXED", "1") != "0"
        # sin / cos / tan: the branch-free Cody-Waite + minimax evaluation (rk_fsincos, ~1 ulp, the
        # accuracy class of libdevice's own fast path) in exact programs too -- libdevice's slow-
        # path branch per call splits the step into small basic blocks; the IEEE division and
        # sqrt of exact programs stay.  RK_FAST_TRIG=0 restores sincosf.
        self.fast_trig = self.fast or os.environ.get("RK_FAST_TRIG", "1") != "0"
        # per-step loads run this many steps ahead of their use (a cp.async ring in shared memory): one step of
        # a lone warp is shorter than an HBM round trip, so a distance of 1 leaves the step time
        # pinned at the memory latency
        self.pf_depth = max(1, int(os.environ.get("RK_PREFETCH", "3")))
        self.lines: List[str] = []
        self.tmp = 0

    # ---- register vs shared-memory residency ------------------------------------------
    def _classify(self):
        L = self.L
        self._plan_ell()
        smem = set()
        dtype: Dict[int, str] = {}
        for rec in self.ins:
            vals = ([rec["dst"]] if rec["dst"] else []) + rec["srcs"]
            multi = rec["n"] > L and id(rec) not in self.ell_rec
            for v in vals:
                if v["kind"] != "val":
                    continue
                dtype.setdefault(v["id"], v["dtype"])
                if v["id"] in self.ell_val:
                    continue
                if multi or v["n"] > L:
                    smem.add(v["id"])
            # a flat shared-memory copy is the cheapest way to move multi-word rows
        self.smem = smem
        self.dtype = dtype

    def _plan_ell(self):
        """Find the wide values that can live in registers (see _EllFamily)."""
        self.ell_val: Dict[int, _EllFamily] = {}
        self.ell_rec: Dict[int, _EllFamily] = {}
        self.ell_tables: List[str] = []
        L = self.L
        if os.environ.get("RK_ELL", "1") == "0" or L < 2:
            return
        img = self.p.meta["const_image"]
        wide = lambda v: v["kind"] == "val" and v["n"] > L          # noqa: E731
        parent: Dict[int, int] = {}

        def find(x):
            while parent.setdefault(x, x) != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x
        touching = []
        for rec in self.ins:
            vals = ([rec["dst"]] if rec["dst"] else []) + rec["srcs"]
            ids = [v["id"] for v in vals if wide(v)]
            if ids:
                touching.append((rec, ids))
                for i in ids[1:]:
                    parent[find(i)] = find(ids[0])
        comps: Dict[int, List] = {}
        for rec, ids in touching:
            comps.setdefault(find(ids[0]), []).append(rec)
        fid = 0
        for recs in comps.values():
            fam = self._ell_component(recs, img, fid)
            if fam is None:
                continue
            fid += 1
            for rec in recs:
                self.ell_rec[id(rec)] = fam
                for v in ([rec["dst"]] if rec["dst"] else []) + rec["srcs"]:
                    if wide(v):
                        self.ell_val[v["id"]] = fam

    def _ell_component(self, recs, img, fid):
        L = self.L
        n = None
        segs: List[Tuple[int, ...]] = []            # distinct element -> segment functions
        persistent = self.persistent

        def seg_of(ptr, idx, n_el):
            out = [-1] * n_el
            for sgi in range(len(ptr) - 1):
                for e in idx[ptr[sgi]:ptr[sgi + 1]]:
                    if out[int(e)] >= 0:
                        return None                 # an element in two segments
                    out[int(e)] = sgi
            return tuple(out)
        for rec in recs:
            op, dst, srcs = rec["op"], rec["dst"], rec["srcs"]
            vals = ([dst] if dst else []) + srcs
            for v in vals:
                if v["kind"] == "val" and v["n"] > L:
                    if v["uniform"] or v["id"] in persistent or (n is not None and v["n"] != n):
                        return None
                    n = v["n"]
            if op in _RED:
                src = srcs[0]
                if not (src["kind"] == "val" and src["n"] > L and src["map"] == isa.MAP_IDENT
                        and rec["n"] <= L and dst["n"] <= L and not dst["uniform"]):
                    return None
                ptr, idx = self.tables(rec)
                sg = seg_of(ptr, idx, src["n"])
                if sg is None:
                    return None
                rec["_seg"] = sg
                if sg not in segs:
                    segs.append(sg)
            elif op == "RPRODX":
                src = srcs[0]
                tab = rec.get("tab")
                if tab is None or not (src["kind"] == "val" and src["n"] > L and dst["n"] == src["n"]
                                       and src["map"] == isa.MAP_IDENT and src["id"] != dst["id"]):
                    return None
                ptr, idx = (np.asarray(tab[k]).astype(np.int64) for k in ("ptr", "idx"))
                sg = seg_of(ptr, idx, src["n"])
                if sg is None or min(sg) < 0 or len(ptr) - 1 > L:
                    return None
                rec["_seg"] = sg
                if sg not in segs:
                    segs.append(sg)
            elif op in ("LDB", "STB", "LDU", "STG", "STERR", "ERRCHK", "QSUM", "RARGMAX", "RARGMIN", "NOP"):
                return None
            else:                                   # element-wise
                if dst is None or dst["n"] <= L or rec["n"] != dst["n"]:
                    return None
                for v in srcs:
                    if v["kind"] == "val" and v["n"] > L:
                        if v["map"] != isa.MAP_IDENT:
                            return None
                    elif v["kind"] == "val":
                        if v["map"] == isa.MAP_IDENT:
                            return None             # a narrow value read element-for-element: n mismatch
                        if v["map"] != isa.MAP_BCAST and \
                                int(img[v["map"]: v["map"] + rec["n"]].max()) >= L:
                            return None
        if n is None or not segs or len(segs) > 2:
            return None
        total = [sg for sg in segs if min(sg) >= 0 and max(sg) < L]
        if not total:
            return None
        # RPRODX needs its segmentation to be the lane assignment
        px = [rec["_seg"] for rec in recs if rec["op"] == "RPRODX"]
        if len(set(px)) > 1 or (px and px[0] not in total):
            return None
        segA = px[0] if px else total[0]
        rest = [sg for sg in segs if sg != segA]
        segB = rest[0] if rest else None
        if segB is not None and max(segB) >= L:
            return None
        fam = _EllFamily(fid, n, list(segA), None if segB is None else list(segB), L)
        if fam.K > 12:
            return None
        fam.segA, fam.segB_t = segA, segB
        return fam

    def is_reg(self, v) -> bool:
        return v["kind"] == "val" and v["id"] not in self.smem

    @staticmethod
    def ctype(dt: str) -> str:
        return "float" if dt == "f" else "int"

    def var(self, v) -> str:
        return f"v{v['id']}"

    # ---- operand expressions --------------------------------------------------------------
    def from_bits(self, expr: str, dt: str) -> str:
        return f"__uint_as_float({expr})" if dt == "f" else f"(int)({expr})"

    def to_bits(self, expr: str, dt: str) -> str:
        return f"__float_as_uint({expr})" if dt == "f" else f"(uint32_t)({expr})"

    def map_index(self, v, e: str, n_r: int) -> str:
        """``e`` is always an in-range element index (callers clamp idle lanes to 0)."""
        mp = v["map"]
        if mp == isa.MAP_IDENT:
            return e
        if mp == isa.MAP_BCAST:
            return "0u"
        return f"C[{mp}u + {e}]"

    def opnd(self, v, e: str, n_r: int, want: str) -> str:
        """C expression of operand ``v`` for element ``e`` of an n_r-element instruction,
        typed ``want`` ('f' float, 'i' int -- bool shares int)."""
        want_c = "f" if want == "f" else "i"
        if v["kind"] == "const":
            c = v["const"]
            mp = v["map"]
            if mp == isa.MAP_BCAST or c.size == 1:
                if want_c == "f":
                    return _flit(float(c[0]))
                return f"({int(c[0])})"
            if e == "ec" and n_r <= self.L:
                # loop-invariant per-lane word of the constant region: read it once, before
                # the step loop (the compiler cannot hoist it past the shared-memory stores)
                name = f"hc{v['base']}_{mp}_{n_r}"
                self.hoisted[name] = (f"  const uint32_t {name} = C[{v['base']}u + "
                                      f"{self.map_index(v, f'(le < {n_r}u ? le : 0u)', n_r)}];")
                return self.from_bits(name, want_c)
            return self.from_bits(f"C[{v['base']}u + {self.map_index(v, e, n_r)}]", want_c)
        if v["kind"] == "mparam":
            name = f"hm{v['base']}"
            self.hoisted[name] = f"  const uint32_t {name} = C[{v['base']}u];"
            return self.from_bits(name, want_c)
        if self.is_reg(v):
            name = self.var(v)
            src_c = "f" if self.dtype[v["id"]] == "f" else "i"
            if v["map"] == isa.MAP_IDENT:
                ex = name
            elif v["map"] == isa.MAP_BCAST:
                ex = f"__shfl_sync(0xffffffffu, {name}, {'0u' if v['uniform'] else 'qL'})"
            else:
                base = "0u" if v["uniform"] else "qL"
                ex = f"__shfl_sync(0xffffffffu, {name}, {base} + {self.map_index(v, e, n_r)})"
            if src_c != want_c:
                ex = f"__uint_as_float((uint32_t)({ex}))" if want_c == "f" else f"(int)__float_as_uint({ex})"
            return ex
        # shared-memory resident value (q clamped: spare lanes must stay inside the warp region)
        qs = f" + qc * {v['sq']}u" if v["sq"] else ""
        return self.from_bits(f"R[{v['base']}u{qs} + {self.map_index(v, e, n_r)}]", want_c)

    def emit(self, s: str):
        self.lines.append(s)

    def recip_opnd(self, v, n_r: int) -> str:
        """1 / (constant operand) as a literal or a hoisted per-lane register."""
        c = v["const"]
        if v["map"] == isa.MAP_BCAST or c.size == 1:
            with np.errstate(all="ignore"):
                return _flit(float(np.float32(1.0) / np.float32(c[0])))
        name = f"hr{v['base']}_{v['map']}_{n_r}"
        idx = self.map_index(v, f"(le < {n_r}u ? le : 0u)", n_r)
        self.hoisted[name] = (f"  const float {name} = 1.f / __uint_as_float(C[{v['base']}u + {idx}]);")
        return name

    # ---- sin / cos / tan of one argument share a single range reduction -------------------
    def plan_trig(self, lst):
        """Group the SIN / COS / TAN instructions of one list by operand: a group of two or
        more becomes one sincosf() (tan = sin / cos), emitted where the first member stood."""
        self.trig_skip, self.trig_lead = set(), {}
        groups: Dict[Tuple, List[Any]] = {}
        gen: Dict[int, int] = {}          # value id -> how often it has been (re)assigned so far
        for rec in lst:
            if rec["op"] in ("SIN", "COS", "TAN") and rec["n"] <= self.L and self.is_reg(rec["dst"]):
                a = rec["srcs"][0]
                if a["kind"] == "val":      # same operand *and* same assignment of it
                    key = (a["id"], a["map"], rec["n"], gen.get(a["id"], 0))
                    groups.setdefault(key, []).append(rec)
            if rec["dst"] is not None and rec["dst"]["kind"] == "val":
                gen[rec["dst"]["id"]] = gen.get(rec["dst"]["id"], 0) + 1
        for recs in groups.values():
            if len({r["op"] for r in recs}) >= 2:
                self.trig_lead[id(recs[0])] = recs
                for r in recs[1:]:
                    self.trig_skip.add(id(r))

    def gen_trig_group(self, recs):
        n = recs[0]["n"]
        x = self.opnd(recs[0]["srcs"][0], "ec", n, "f")
        self.emit(f"  // sincos group: {' '.join(r['op'] for r in recs)} n={n}")
        asg = []
        for r in recs:
            ex = {"SIN": "sn_", "COS": "cs_",
                  "TAN": "rk_fdiv(sn_, cs_)" if self.fast else "(sn_ / cs_)"}[r["op"]]
            asg.append(f"{self.var(r['dst'])} = {ex};")
        self.emit(f"  {{ const uint32_t ec = le < {n}u ? le : 0u; float sn_, cs_; "
                  f"{'rk_fsincos' if self.fast_trig else 'sincosf'}({x}, &sn_, &cs_); " +
                  " ".join(asg) + " }")

    # ---- global-memory addressing: one 64-bit base per buffer per step, 32-bit offsets ----
    def gbase(self, buf: int, kind: str, tstep: bool, nxt: bool = False) -> str:
        """Name of the (per list, per step) base pointer of a global buffer.
        kind: 'b' per-rollout [W][B] blocks, 'u' uniform rows, 'g' gradient partial rows.
        nxt: the base for the step after this one (software prefetch of per-step loads)."""
        name = f"gn{buf}" if nxt else f"gp{buf}"
        if nxt:
            if buf not in self.next_bases:
                ex = "tn * K.W[%d] * (long long)K.B" % buf if kind == "b" else "tn * K.W[%d]" % buf
                self.next_bases[buf] = (f"  const uint32_t* const {name} = "
                                        f"(const uint32_t*)K.buf[{buf}] + {ex};")
            return name
        if buf not in self.bases:
            tt = "t" if tstep else "0ll"
            if kind == "b":
                ex = f"{tt} * K.W[{buf}] * (long long)K.B"
            elif kind == "u":
                ex = f"{tt} * K.W[{buf}]"
            else:
                ex = f"(gwarp * (long long)K.T + {tt}) * K.W[{buf}]"
            self.bases[buf] = f"  uint32_t* const {name} = (uint32_t*)K.buf[{buf}] + {ex};"
        return name

    def lane_off(self, n: int) -> str:
        """b0*n + q*n + le for an n-element per-rollout value (hoisted out of the step loop)."""
        self.lane_offs.add(n)
        return f"lo{n}"

    def tables(self, rec):
        img = self.p.meta["const_image"]
        n = rec["n"]
        if rec["op"] == "RPRODX":
            return None, None
        ptr = img[rec["aux0"]: rec["aux0"] + n + 1].astype(np.int64)
        idx = img[rec["aux1"]: rec["aux1"] + int(ptr[-1])].astype(np.int64)
        return ptr, idx

    # ---- one instruction ----------------------------------------------------------------------
    def gen_ins(self, rec):
        op, n, L = rec["op"], rec["n"], self.L
        dst, srcs = rec["dst"], rec["srcs"]
        self.emit(f"  // {op} n={n}")
        if op == "NOP":
            return
        if id(rec) in self.trig_skip:
            return
        if id(rec) in self.trig_lead:
            return self.gen_trig_group(self.trig_lead[id(rec)])
        if id(rec) in self.ell_rec:
            return self.gen_ell(rec, self.ell_rec[id(rec)])
        if op in ("LDB", "STB", "LDU", "STG", "STERR", "ERRCHK", "QSUM"):
            return self.gen_special(rec)
        if op in _RED or op in ("RARGMAX", "RARGMIN", "RPRODX"):
            return self.gen_reduce(rec)
        # ---- elementwise
        dt_out = dst["dtype"]
        single = n <= L and (self.is_reg(dst) or True)
        recip = None
        if op == "DIV" and self.fast and srcs[1]["kind"] == "const" and n <= L:
            # relaxed (train / adjoint) programs: x / c with a build-time constant c becomes
            # x * (1/c) -- one rounding of difference, inside the 1e-5 / 1e-4 parity tolerances;
            # exact programs keep the IEEE division (bit-exact bool / int contract)
            recip = self.recip_opnd(srcs[1], n)
        if recip is not None:
            self.emit(f"  {{ const uint32_t ec = le < {n}u ? le : 0u; const float x = "
                      f"{self.opnd(srcs[0], 'ec', n, 'f')}; {self.var(dst)} = x * {recip}; }}"
                      if self.is_reg(dst) else
                      f"  {{ const uint32_t ec = le < {n}u ? le : 0u; const float x = "
                      f"{self.opnd(srcs[0], 'ec', n, 'f')}; const float r = x * {recip}; "
                      f"if ({'q == 0u' if dst['uniform'] else f'q < {self.rpw}u'} && le < {n}u) "
                      f"R[{dst['base']}u{(' + q * %du' % dst['sq']) if dst['sq'] else ''} + le] = "
                      f"__float_as_uint(r); }}")
            if not self.is_reg(dst):
                self.emit("  __syncwarp();")
            return
        if self.fast and op in _FAST:
            body, tin = _FAST[op], ("ff" if op in _F2 else "f")
        elif self.fast_trig and op in ("SIN", "COS"):
            body, tin = _FAST[op], "f"
        elif self.fast_trig and op == "TAN":
            body, tin = "rk_ftan_ieee(x)", "f"
        elif op in _F1:
            body, tin = _F1[op], "f"
        elif op in _F2:
            body, tin = _F2[op], "ff"
        elif op in _F2B:
            body, tin = f"((x {_F2B[op]} y) ? 1 : 0)", "ff"
        elif op in _I1:
            body, tin = _I1[op], "i"
        elif op in _I2:
            body, tin = _I2[op], "ii"
        elif op == "MOV":
            body, tin = "x", dt_out if dt_out == "f" else "i"
        elif op == "SELECT":
            d = "f" if dt_out == "f" else "i"
            body, tin = "((x != 0) ? y : z)", "i" + d + d
        elif op == "FMA":
            body, tin = "__fadd_rn(__fmul_rn(x, y), z)", "fff"
        elif op == "LERP":
            body, tin = "__fadd_rn(__fmul_rn(x, y), __fmul_rn(__fsub_rn(1.f, x), z))", "fff"
        elif op == "I2F":
            body, tin = "(float)x", "i"
        elif op == "F2I":
            body, tin = "(int)x", "f"
        elif op == "F2B":
            body, tin = "((x != 0.f) ? 1 : 0)", "f"
        else:
            raise NotImplementedError(f"codegen: op {op}")
        names = "xyz"
        out_c = "f" if dt_out == "f" else "i"

        def stmt(e: str) -> List[str]:
            ls = []
            for k, tk in enumerate(tin):
                ls.append(f"const {self.ctype(tk)} {names[k]} = {self.opnd(srcs[k], e, n, tk)};")
            ls.append(f"const {self.ctype(out_c)} r = {body};")
            return ls

        uni = dst["uniform"]
        ec = f"const uint32_t ec = le < {n}u ? le : 0u;"
        if n <= L and self.is_reg(dst):
            self.emit("  { " + ec + " " + " ".join(stmt("ec")) + f" {self.var(dst)} = r; }}")
            return
        guard_q = "q == 0u" if uni else f"q < {self.rpw}u"
        qs = f" + q * {dst['sq']}u" if dst["sq"] else ""
        if n <= L:
            # single pass into shared memory: every lane evaluates (shuffles need all lanes)
            self.emit("  { " + ec + " " + " ".join(stmt("ec")) +
                      f" if ({guard_q} && le < {n}u) R[{dst['base']}u{qs} + le] = {self.to_bits('r', out_c)}; }}")
        else:
            self.emit(f"  if ({guard_q}) for (uint32_t e = le; e < {n}u; e += {L}u) {{ " +
                      " ".join(stmt("e")) +
                      f" R[{dst['base']}u{qs} + e] = {self.to_bits('r', out_c)}; }}")
        self.emit("  __syncwarp();")

    # ---- wide values in registers (see _EllFamily) ----------------------------------------------
    def ell_table(self, fam, tag: str, values: np.ndarray, ctype: str = "uint32_t") -> str:
        """A per-(slot, lane) table as a __constant__ array; returns its name."""
        name = f"ELL{fam.fid}_{tag}"
        if not any(name + "[" in t for t in self.ell_tables):
            flat = values.T.reshape(-1)             # [K][L]
            if ctype == "float":
                body = ", ".join(_flit(float(x)) .replace("__uint_as_float(", "").rstrip(")") for x in flat)
                self.ell_tables.append(f"__device__ const uint32_t {name}[{flat.size}] = {{{body}}};")
            else:
                body = ", ".join(f"{int(x)}u" for x in flat)
                self.ell_tables.append(f"__device__ const uint32_t {name}[{flat.size}] = {{{body}}};")
        return name

    def ell_lane_reg(self, fam, tag: str, values: np.ndarray, k: int) -> str:
        """Hoisted per-lane register holding values[lane][k]."""
        tab = self.ell_table(fam, tag, values)
        reg = f"el{fam.fid}_{tag}_{k}"
        self.hoisted[reg] = f"  const uint32_t {reg} = rk_ld_tab({tab} + {k * self.L}u + le);"
        return reg

    def ell_mask(self, fam, which: str) -> str:
        """Hoisted per-lane bit mask of the occupied slots (which = 'a': own elements, 'b': members of
        the lane's segment under the second segmentation)."""
        at = fam.elem_at if which == "a" else fam.elem_b
        bits = np.array([[sum(1 << k for k in range(fam.K) if at[l, k] >= 0)] for l in range(self.L)])
        tab = self.ell_table(fam, "m" + which, bits)
        reg = f"el{fam.fid}_m{which}"
        self.hoisted[reg] = f"  const uint32_t {reg} = rk_ld_tab({tab} + le);"
        return reg

    def ell_var(self, v, k: int) -> str:
        return f"v{v['id']}_{k}"

    def ell_opnd(self, fam, v, k: int, n_r: int, want: str) -> str:
        want_c = "f" if want == "f" else "i"
        L = self.L
        img = self.p.meta["const_image"]
        at = fam.elem_at[:, k]                      # element held by each lane in this slot (or -1)
        if v["kind"] == "val" and v["id"] in self.ell_val:
            name = self.ell_var(v, k)
            src_c = "f" if self.dtype[v["id"]] == "f" else "i"
            if src_c != want_c:
                return f"__uint_as_float((uint32_t)({name}))" if want_c == "f" else f"(int)__float_as_uint({name})"
            return name
        if v["kind"] == "mparam":
            return self.opnd(v, "ec", n_r, want)
        mp = v["map"]
        if v["kind"] == "const":
            c = v["const"].reshape(-1)
            if mp == isa.MAP_BCAST or c.size == 1:
                return self.opnd(v, "ec", n_r, want)
            idx = np.arange(n_r) if mp == isa.MAP_IDENT else img[mp: mp + n_r].astype(np.int64)
            vals = np.array([c[idx[e]] if e >= 0 else c[idx[max(at.max(), 0)]] for e in at])
            used = vals[at >= 0]
            if used.size and np.all(used == used[0]):
                return _flit(float(used[0])) if want_c == "f" else f"({int(used[0])})"
            full = np.zeros((L, fam.K), dtype=np.float64)
            for kk in range(fam.K):
                a2 = fam.elem_at[:, kk]
                full[:, kk] = [c[idx[e]] if e >= 0 else 0 for e in a2]
            tag = f"c{v['base']}_{mp}"
            if v["dtype"] == "f":
                tab = self.ell_table(fam, tag, full, "float")
            else:
                tab = self.ell_table(fam, tag, full.astype(np.int64) & 0xFFFFFFFF)
            reg = f"el{fam.fid}_{tag}_{k}"
            self.hoisted[reg] = f"  const uint32_t {reg} = rk_ld_tab({tab} + {k * L}u + le);"
            return self.from_bits(reg, want_c)
        # narrow value
        if mp == isa.MAP_BCAST:
            return self.opnd(v, "ec", n_r, want)
        tblm = img[mp: mp + n_r].astype(np.int64)
        srcs = np.array([tblm[e] if e >= 0 else 0 for e in at])
        lanes = np.arange(L)
        src_c = "f" if self.dtype[v["id"]] == "f" else "i"
        if self.is_reg(v):
            if np.all(srcs[at >= 0] == lanes[at >= 0]):
                ex = self.var(v)                    # the lane's own element: no shuffle
            else:
                full = np.zeros((L, fam.K), dtype=np.int64)
                for kk in range(fam.K):
                    a2 = fam.elem_at[:, kk]
                    full[:, kk] = [tblm[e] if e >= 0 else 0 for e in a2]
                reg = self.ell_lane_reg(fam, f"g{mp}", full, k)
                base = "0u" if v["uniform"] else "qL"
                ex = f"__shfl_sync(0xffffffffu, {self.var(v)}, {base} + {reg})"
            if src_c != want_c:
                ex = f"__uint_as_float((uint32_t)({ex}))" if want_c == "f" else f"(int)__float_as_uint({ex})"
            return ex
        full = np.zeros((L, fam.K), dtype=np.int64)
        for kk in range(fam.K):
            a2 = fam.elem_at[:, kk]
            full[:, kk] = [tblm[e] if e >= 0 else 0 for e in a2]
        reg = self.ell_lane_reg(fam, f"g{mp}", full, k)
        qs = f" + qc * {v['sq']}u" if v["sq"] else ""
        return self.from_bits(f"R[{v['base']}u{qs} + {reg}]", want_c)

    def gen_ell(self, rec, fam):
        op, n, L, K = rec["op"], rec["n"], self.L, fam.K
        dst, srcs = rec["dst"], rec["srcs"]
        if op in _RED:
            ct, init, upd = _RED[op]
            tin = "f" if ct == "float" else "i"
            src = srcs[0]
            second = rec["_seg"] != fam.segA
            mask = self.ell_mask(fam, "b" if second else "a")
            body = [f"{ct} acc = {init};"]
            for k in range(K):
                at = fam.elem_b[:, k] if second else fam.elem_at[:, k]
                if not np.any(at >= 0):
                    continue
                x = self.ell_opnd(fam, src, k, n, tin)
                if second:
                    # the member of this lane's segment that sits in slot k lives in lane lane[e]
                    lanes = np.zeros((L, K), dtype=np.int64)
                    for kk in range(K):
                        lanes[:, kk] = [fam.lane[e] if e >= 0 else 0 for e in fam.elem_b[:, kk]]
                    reg = self.ell_lane_reg(fam, "sb", lanes, k)
                    x = f"__shfl_sync(0xffffffffu, {x}, qL + {reg})"
                ok = "true" if np.all(at[:rec["n"]] >= 0) and rec["n"] == L else f"(({mask} >> {k}u) & 1u)"
                # the shuffle is executed by every lane; only its use is conditional
                body.append(f"{{ const {ct} s_ = {x}; const {ct} t = {ok} ? s_ : {init}; {upd}; }}")
            code = " ".join(body)
            out_c = "f" if ct == "float" else "i"
            if self.is_reg(dst):
                self.emit(f"  {{ {code} {self.var(dst)} = acc; }}")
            else:
                qs_d = f" + q * {dst['sq']}u" if dst["sq"] else ""
                self.emit(f"  {{ {code} if (q < {self.rpw}u && le < {rec['n']}u) "
                          f"R[{dst['base']}u{qs_d} + le] = {self.to_bits('acc', out_c)}; }}")
                self.emit("  __syncwarp();")
            return
        if op == "RPRODX":
            src = srcs[0]
            mask = self.ell_mask(fam, "a")
            ls = ["float run = 1.f;"]
            for k in range(K):
                ls.append(f"{self.ell_var(dst, k)} = run; "
                          f"run = run * ((({mask} >> {k}u) & 1u) ? {self.ell_opnd(fam, src, k, n, 'f')} : 1.f);")
            ls.append("run = 1.f;")
            for k in range(K - 1, -1, -1):
                ls.append(f"{self.ell_var(dst, k)} = {self.ell_var(dst, k)} * run; "
                          f"run = run * ((({mask} >> {k}u) & 1u) ? {self.ell_opnd(fam, src, k, n, 'f')} : 1.f);")
            self.emit("  { " + " ".join(ls) + " }")
            return
        body, tin, out_c = self._ew_body(rec)
        names = "xyz"
        for k in range(K):
            if not np.any(fam.elem_at[:, k] >= 0):
                continue
            ls = [f"const {self.ctype(tk)} {names[j]} = {self.ell_opnd(fam, srcs[j], k, n, tk)};"
                  for j, tk in enumerate(tin)]
            self.emit("  { " + " ".join(ls) + f" {self.ell_var(dst, k)} = {body}; }}")

    def _ew_body(self, rec):
        """(C expression in x, y, z; operand types; result type) of an element-wise instruction."""
        op, dst, srcs = rec["op"], rec["dst"], rec["srcs"]
        dt_out = dst["dtype"]
        if self.fast and op in _FAST:
            body, tin = _FAST[op], ("ff" if op in _F2 else "f")
        elif self.fast_trig and op in ("SIN", "COS"):
            body, tin = _FAST[op], "f"
        elif self.fast_trig and op == "TAN":
            body, tin = "rk_ftan_ieee(x)", "f"
        elif op in _F1:
            body, tin = _F1[op], "f"
        elif op in _F2:
            body, tin = _F2[op], "ff"
        elif op in _F2B:
            body, tin = f"((x {_F2B[op]} y) ? 1 : 0)", "ff"
        elif op in _I1:
            body, tin = _I1[op], "i"
        elif op in _I2:
            body, tin = _I2[op], "ii"
        elif op == "MOV":
            body, tin = "x", dt_out if dt_out == "f" else "i"
        elif op == "SELECT":
            d = "f" if dt_out == "f" else "i"
            body, tin = "((x != 0) ? y : z)", "i" + d + d
        elif op == "FMA":
            body, tin = "__fadd_rn(__fmul_rn(x, y), z)", "fff"
        elif op == "LERP":
            body, tin = "__fadd_rn(__fmul_rn(x, y), __fmul_rn(__fsub_rn(1.f, x), z))", "fff"
        elif op == "I2F":
            body, tin = "(float)x", "i"
        elif op == "F2I":
            body, tin = "(int)x", "f"
        elif op == "F2B":
            body, tin = "((x != 0.f) ? 1 : 0)", "f"
        else:
            raise NotImplementedError(f"codegen: op {op}")
        return body, tin, ("f" if dt_out == "f" else "i")

    # ---- reductions ---------------------------------------------------------------------------
    def gen_reduce(self, rec):
        op, n, L = rec["op"], rec["n"], self.L
        dst, src = rec["dst"], rec["srcs"][0]
        ptr, idx, kmax = rec["aux0"], rec["aux1"], rec.get("kmax", 0)
        uni = dst["uniform"]
        guard_q = "q == 0u" if uni else f"q < {self.rpw}u"
        qs_d = f" + q * {dst['sq']}u" if dst["sq"] else ""
        src_reg = self.is_reg(src)
        if op in _RED and src_reg and n <= L and self.gen_reduce_static(rec):
            return
        if op == "RPRODX" and self.gen_prodx_scan(rec):
            return
        if op in _RED:
            ct, init, upd = _RED[op]
            tin = "f" if ct == "float" else "i"
        elif op in ("RARGMAX", "RARGMIN"):
            ct, tin = "int", "f"
        else:
            ct, tin = "float", "f"

        def fetch(sidx: str, ok: str) -> str:
            """value of source element sidx (per-lane), evaluated by all lanes."""
            if src["kind"] == "const":
                return self.from_bits(f"C[{src['base']}u + {sidx}]", tin)
            if src_reg:
                base = "0u" if src["uniform"] else "qL"
                ex = f"__shfl_sync(0xffffffffu, {self.var(src)}, {base} + {sidx})"
                if (self.dtype[src["id"]] == "f") != (tin == "f"):
                    ex = (f"__uint_as_float((uint32_t)({ex}))" if tin == "f"
                          else f"(int)__float_as_uint({ex})")
                return ex
            qs = f" + qc * {src['sq']}u" if src["sq"] else ""
            rd = self.from_bits(f"R[{src['base']}u{qs} + {sidx}]", tin)
            return f"(({ok}) ? {rd} : ({ct})0)"

        def body(e: str, shuffle_safe: bool) -> List[str]:
            ls = [f"const bool act = ({e}) < {n}u && q < {self.rpw}u;",
                  f"const uint32_t j0 = act ? C[{ptr}u + ({e})] : 0u;",
                  f"const uint32_t j1 = act ? C[{ptr}u + ({e}) + 1u] : 0u;"]
            if op == "RPRODX":
                seg = rec["seg"]
                ls = [f"const bool act = ({e}) < {n}u && q < {self.rpw}u;",
                      f"const uint32_t sg_ = act ? C[{seg}u + ({e})] : 0u;",
                      f"const uint32_t j0 = act ? C[{ptr}u + sg_] : 0u;",
                      f"const uint32_t j1 = act ? C[{ptr}u + sg_ + 1u] : 0u;"]
                ls.append("float acc = 1.f;")
                loop_hdr = f"for (uint32_t k = 0; k < {kmax}u; ++k)" if shuffle_safe else \
                    "for (uint32_t k = 0; k < j1 - j0; ++k)"
                ls.append(loop_hdr + " { const bool ok = j0 + k < j1; "
                          f"const uint32_t s = ok ? C[{idx}u + j0 + k] : 0u; "
                          f"const float t = {fetch('s', 'ok')}; if (ok && s != ({e})) acc = acc * t; }}")
                return ls
            if op in ("RARGMAX", "RARGMIN"):
                sgn = "1.f" if op == "RARGMAX" else "-1.f"
                ls += ["float best = -CUDART_INF_F; int acc = 0;"]
                loop_hdr = f"for (uint32_t k = 0; k < {kmax}u; ++k)" if shuffle_safe else \
                    "for (uint32_t k = 0; k < j1 - j0; ++k)"
                ls.append(loop_hdr + " { const bool ok = j0 + k < j1; "
                          f"const uint32_t s = ok ? C[{idx}u + j0 + k] : 0u; "
                          f"const float t = {sgn} * {fetch('s', 'ok')}; "
                          "if (ok && t > best) { best = t; acc = (int)k; } }")
                return ls
            ls.append(f"{ct} acc = {init};")
            loop_hdr = f"for (uint32_t k = 0; k < {kmax}u; ++k)" if shuffle_safe else \
                "for (uint32_t k = 0; k < j1 - j0; ++k)"
            ls.append(loop_hdr + " { const bool ok = j0 + k < j1; "
                      f"const uint32_t s = ok ? C[{idx}u + j0 + k] : 0u; "
                      f"const {ct} t = {fetch('s', 'ok')}; if (ok) {{ {upd}; }} }}")
            return ls

        out_c = "f" if ct == "float" else "i"
        if n <= L:
            pre = "#pragma unroll\n" if False else ""
            code = " ".join(body("le", True))
            if self.is_reg(dst):
                self.emit(f"  {{ {code} {self.var(dst)} = acc; }}")
            else:
                self.emit(f"  {{ {code} if ({guard_q} && le < {n}u) "
                          f"R[{dst['base']}u{qs_d} + le] = {self.to_bits('acc', out_c)}; }}")
                self.emit("  __syncwarp();")
        else:
            assert not src_reg
            code = " ".join(body("e", False))
            self.emit(f"  if ({guard_q}) for (uint32_t e = le; e < {n}u; e += {L}u) {{ {code} "
                      f"R[{dst['base']}u{qs_d} + e] = {self.to_bits('acc', out_c)}; }}")
            self.emit("  __syncwarp();")

    def gen_prodx_scan(self, rec) -> bool:
        """Product of the OTHER members of each element's segment (the derivative of RPROD) as a
        prefix scan and a suffix scan: one lane owns one segment, so a value of n elements costs
        2 * max-degree steps in a single pass instead of n * max-degree.  Needs shared-memory
        operands and segments that partition the elements (checked here, else the generic loop)."""
        n, L = rec["n"], self.L
        dst, src = rec["dst"], rec["srcs"][0]
        tab = rec.get("tab")
        if tab is None or src["kind"] != "val" or self.is_reg(src) or self.is_reg(dst):
            return False
        ptr, idx, seg = (np.asarray(tab[k]).astype(np.int64) for k in ("ptr", "idx", "seg"))
        nseg = len(ptr) - 1
        if len(idx) != n or len(seg) != n or not np.array_equal(np.sort(idx), np.arange(n)):
            return False
        for g_ in range(nseg):
            if np.any(seg[idx[ptr[g_]:ptr[g_ + 1]]] != g_):
                return False
        guard_q = "q == 0u" if dst["uniform"] else f"q < {self.rpw}u"
        qs_d = f" + q * {dst['sq']}u" if dst["sq"] else ""
        qs_s = f" + qc * {src['sq']}u" if src["sq"] else ""
        P, I = rec["aux0"], rec["aux1"]
        self.emit(
            f"  if ({guard_q}) for (uint32_t sgi = le; sgi < {nseg}u; sgi += {L}u) {{ "
            f"const uint32_t j0 = C[{P}u + sgi], j1 = C[{P}u + sgi + 1u]; float run = 1.f; "
            f"for (uint32_t j = j0; j < j1; ++j) {{ const uint32_t s = C[{I}u + j]; "
            f"R[{dst['base']}u{qs_d} + s] = __float_as_uint(run); "
            f"run *= __uint_as_float(R[{src['base']}u{qs_s} + s]); }} run = 1.f; "
            f"for (uint32_t j = j1; j > j0; --j) {{ const uint32_t s = C[{I}u + j - 1u]; "
            f"R[{dst['base']}u{qs_d} + s] = __float_as_uint(__uint_as_float(R[{dst['base']}u{qs_d} + s]) * run); "
            f"run *= __uint_as_float(R[{src['base']}u{qs_s} + s]); }} }}")
        self.emit("  __syncwarp();")
        return True

    def gen_reduce_static(self, rec) -> bool:
        """Register-resident segment reduction with the CSR tables resolved at generation
        time: one shuffle per member, in the reference's (left-to-right) member order; lane
        indices that are the same for every element, or a fixed shift of the element index,
        cost no table load at all."""
        op, n, L = rec["op"], rec["n"], self.L
        dst, src = rec["dst"], rec["srcs"][0]
        ptr, idx = self.tables(rec)
        deg = np.diff(ptr)
        kmax = int(deg.max()) if len(deg) else 0
        if kmax > 64:
            return False
        ct, init, upd = _RED[op]
        tin = "f" if ct == "float" else "i"
        base = "0u" if src["uniform"] else "qL"
        conv = (self.dtype[src["id"]] == "f") != (tin == "f")
        if n == 1 and kmax >= 4 and not conv and src["uniform"] == dst["uniform"] \
                and self.is_reg(dst) and L >= 2 and (L & (L - 1)) == 0:
            # one segment (a reward-style total): xor-butterfly over the rollout's lanes, log2(L) shuffles
            # instead of one per member; lanes outside the segment contribute the identity.  (The order of
            # the additions differs from the member order -- as it does in XLA's own reductions.  A value
            # shared by all rollouts is replicated in every rollout's lanes, so each group reduces its copy.)
            members = [int(m) for m in idx[ptr[0]:ptr[1]]]
            if len(set(members)) == len(members) and max(members) < L:
                comb = {"RSUM": "a_ + b_", "RISUM": "a_ + b_", "RPROD": "a_ * b_", "RIPROD": "a_ * b_",
                        "RMIN": "fminf(a_, b_)", "RMAX": "fmaxf(a_, b_)", "RIMIN": "min(a_, b_)",
                        "RIMAX": "max(a_, b_)"}[op]
                if members == list(range(len(members))):
                    cond = f"le < {len(members)}u"
                else:
                    cond = f"(({sum(1 << m for m in members)}u >> le) & 1u)"
                steps = " ".join(
                    f"{{ const {ct} a_ = t_; const {ct} b_ = __shfl_xor_sync(0xffffffffu, t_, {o}u); t_ = {comb}; }}"
                    for o in [L >> (i + 1) for i in range(L.bit_length() - 1)])
                self.emit(f"  {{ {ct} t_ = ({cond}) ? {self.var(src)} : {init}; {steps} {self.var(dst)} = t_; }}")
                return True
        need_dyn = False
        body = [f"{ct} acc = {init};"]
        for k in range(kmax):
            valid = deg > k
            tab = np.array([idx[ptr[e] + k] if valid[e] else -1 for e in range(n)])
            tv = tab[valid]
            if np.all(tv == tv[0]):
                lane = f"{base} + {int(tv[0])}u"
            elif np.all(tv - np.arange(n)[valid] == tv[0] - np.arange(n)[valid][0]):
                sh = int(tv[0] - np.arange(n)[valid][0])
                lane = f"(uint32_t)((int)({base} + ec) + ({sh}))"
            else:
                need_dyn = True
                lane = f"{base} + (({k}u < dg_) ? C[{rec['aux1']}u + j0_ + {k}u] : 0u)"
            ex = f"__shfl_sync(0xffffffffu, {self.var(src)}, {lane})"
            if conv:
                ex = (f"__uint_as_float((uint32_t)({ex}))" if tin == "f"
                      else f"(int)__float_as_uint({ex})")
            if np.all(valid):
                body.append(f"{{ const {ct} t = {ex}; {upd}; }}")
            else:
                need_dyn = True
                body.append(f"{{ const {ct} t = {ex}; if ({k}u < dg_) {{ {upd}; }} }}")
        pre = f"const uint32_t ec = le < {n}u ? le : 0u; "
        if need_dyn:
            pre += (f"const uint32_t j0_ = C[{rec['aux0']}u + ec]; "
                    f"const uint32_t dg_ = C[{rec['aux0']}u + ec + 1u] - j0_; ")
        code = pre + " ".join(body)
        out_c = "f" if ct == "float" else "i"
        if self.is_reg(dst):
            self.emit(f"  {{ {code} {self.var(dst)} = acc; }}")
        else:
            guard_q = "q == 0u" if dst["uniform"] else f"q < {self.rpw}u"
            qs_d = f" + q * {dst['sq']}u" if dst["sq"] else ""
            self.emit(f"  {{ {code} if ({guard_q} && le < {n}u) "
                      f"R[{dst['base']}u{qs_d} + le] = {self.to_bits('acc', out_c)}; }}")
            self.emit("  __syncwarp();")
        return True

    # ---- memory / cross-rollout / error instructions ------------------------------------------------
    def gen_special(self, rec):
        op, n, L, rpw = rec["op"], rec["n"], self.L, self.rpw
        dst, srcs = rec["dst"], rec["srcs"]
        buf, off = rec["buf"], rec["off"]
        tt = "t" if rec["flags"] & isa.F_TSTEP else "0ll"
        if op == "QSUM":
            src = srcs[0]
            if self.is_reg(src) and rpw * L == 32 and (rpw & (rpw - 1)) == 0:
                # full warps: xor-butterfly over the rollouts of the warp (log2(rpw) shuffles);
                # rollouts past the end of the batch (ragged last warp) contribute zero: no branch
                fly = " ".join(f"acc += __shfl_xor_sync(0xffffffffu, acc, {o}u);"
                               for o in [L << i for i in range(rpw.bit_length() - 1)])
                code = f"float acc = (q < nvalid) ? {self.var(src)} : 0.f; {fly}"
            elif self.is_reg(src):
                code = (f"float acc = 0.f; for (uint32_t qq = 0; qq < {rpw}u; ++qq) {{ "
                        f"const float tq = __shfl_sync(0xffffffffu, {self.var(src)}, qq * {L}u + le); "
                        f"if (qq < nvalid) acc += tq; }}")
            else:
                code = (f"float acc = 0.f; if (le < {n}u) for (uint32_t qq = 0; qq < nvalid; ++qq) "
                        f"acc += __uint_as_float(R[{src['base']}u + qq * {src['sq']}u + le]);")
            if self.is_reg(dst):
                self.emit(f"  {{ {code} {self.var(dst)} = acc; }}")
            else:
                self.emit(f"  {{ {code} if (q == 0u && le < {n}u) R[{dst['base']}u + le] = __float_as_uint(acc); }}")
                self.emit("  __syncwarp();")
            return
        tstep = bool(rec["flags"] & isa.F_TSTEP)
        if op in ("LDB", "LDU") and tstep and self.in_step and self.is_reg(dst) and n <= L \
                and self.pf_depth > 1:
            # asynchronous prefetch ring: the words a step reads from global memory are copied
            # into a per-warp shared-memory ring D steps ahead (cp.async, one 4-byte word per
            # lane), so the L2 / HBM latency overlaps D - 1 whole steps instead of stalling the
            # lone warp of a scheduler; this step reads its slot and refills it for step t +- D
            dt = self.dtype[dst["id"]]
            if op == "LDB":
                gn = self.gbase(buf, "b", True, nxt=True)
                src = f"{gn} + ({off}u * Bu + {self.lane_off(n)})"
                cond = f"has{buf} && q < nvalid && le < {n}u"
            else:
                gn = self.gbase(buf, "u", True, nxt=True)
                src = f"{gn} + ({off}u + le)"
                cond = f"has{buf} && le < {n}u"
            i = len(self.pf_stmts)
            self.pf_stmts.append(f"    rk_cp_async4(ring_w + {i * 32}u, {src}, more && {cond});")
            self.emit(f"  {self.var(dst)} = {self.from_bits(f'ring_r[{i * 32}u]', dt)};")
            self.pf_last = len(self.lines)
            return
        if op == "LDB":
            gp = self.gbase(buf, "b", tstep)
            idx = f"{off}u * Bu + {self.lane_off(n)}"
            if self.is_reg(dst):
                dt = self.dtype[dst["id"]]
                self.emit(f"  {self.var(dst)} = (has{buf} && q < nvalid && le < {n}u) ? "
                          f"{self.from_bits(f'{gp}[{idx}]', dt)} : ({self.ctype(dt)})0;")
            else:
                self.emit(f"  if (has{buf}) {{ const uint32_t* src_ = {gp} + {off}u * Bu + b0u * {n}u; "
                          f"for (uint32_t i = lane; i < nvalid * {n}u; i += 32u) R[{dst['base']}u + i] = src_[i]; }}")
                self.emit("  __syncwarp();")
            return
        if op == "STB":
            src = srcs[0]
            gp = self.gbase(buf, "b", tstep)
            if n <= L:
                dt = src["dtype"] if src["kind"] != "val" else self.dtype[src["id"]]
                dtc = "f" if dt == "f" else "i"
                val = self.opnd(src, "ec", n, dtc)
                self.emit(f"  {{ const uint32_t ec = le < {n}u ? le : 0u; const {self.ctype(dtc)} sv = {val}; "
                          f"if (has{buf} && q < nvalid && le < {n}u) "
                          f"{gp}[{off}u * Bu + {self.lane_off(n)}] = {self.to_bits('sv', dtc)}; }}")
            elif src["kind"] == "val" and src["map"] == isa.MAP_IDENT and src["sq"] == n:
                self.emit(f"  if (has{buf}) {{ uint32_t* dst_ = {gp} + {off}u * Bu + b0u * {n}u; "
                          f"for (uint32_t i = lane; i < nvalid * {n}u; i += 32u) dst_[i] = R[{src['base']}u + i]; }}")
            else:
                val = self.opnd(src, "e", n, "i")
                self.emit(f"  if (has{buf} && q < nvalid) for (uint32_t e = le; e < {n}u; e += {L}u) "
                          f"{gp}[{off}u * Bu + b0u * {n}u + q * {n}u + e] = (uint32_t)({val});")
            return
        if op == "LDU":
            gp = self.gbase(buf, "u", tstep)
            if self.is_reg(dst):
                dt = self.dtype[dst["id"]]
                self.emit(f"  {self.var(dst)} = (has{buf} && le < {n}u) ? "
                          f"{self.from_bits(f'__ldg({gp} + {off}u + le)', dt)} : ({self.ctype(dt)})0;")
            else:
                self.emit(f"  if (has{buf}) for (uint32_t i = lane; i < {n}u; i += 32u) "
                          f"R[{dst['base']}u + i] = __ldg({gp} + {off}u + i);")
                self.emit("  __syncwarp();")
            return
        if op == "STG":
            src = srcs[0]
            gp = self.gbase(buf, "g", tstep)
            if n <= L:
                val = self.opnd(src, "ec", n, "f")
                self.emit(f"  {{ const uint32_t ec = le < {n}u ? le : 0u; const float sv = {val}; "
                          f"if (has{buf} && q == 0u && le < {n}u) {gp}[{off}u + le] = __float_as_uint(sv); }}")
            else:
                val = self.opnd(src, "e", n, "f")
                self.emit(f"  if (has{buf} && q == 0u) for (uint32_t e = le; e < {n}u; e += {L}u) "
                          f"{gp}[{off}u + e] = __float_as_uint({val});")
            return
        if op == "ERRCHK":
            src = srcs[0]
            bit = rec["aux0"]
            if n <= L:
                val = self.opnd(src, "ec", n, "i")
                self.emit(f"  {{ const uint32_t ec = le < {n}u ? le : 0u; const int cv = {val}; if (q < nvalid && le < {n}u && cv != 0) "
                          f"atomicOr(R + {self.err_base}u + q, {1 << bit}u); }}")
            else:
                val = self.opnd(src, "e", n, "i")
                self.emit(f"  if (q < nvalid) {{ bool bad = false; for (uint32_t e = le; e < {n}u; e += {L}u) "
                          f"bad |= ({val}) != 0; if (bad) atomicOr(R + {self.err_base}u + q, {1 << bit}u); }}")
            self.emit("  __syncwarp();")
            return
        if op == "STERR":
            gp = self.gbase(buf, "b", bool(rec["flags"] & isa.F_TSTEP))
            self.emit(f"  if (lane < nvalid) {{ if (has{buf}) {gp}[b0u + lane] = "
                      f"R[{self.err_base}u + lane] | {self.p.static_err}u; "
                      f"R[{self.err_base}u + lane] = 0u; }}")
            self.emit("  __syncwarp();")
            return
        raise NotImplementedError(op)

    # ---- whole kernel ---------------------------------------------------------------------------
    def source(self) -> str:
        p = self.p
        decl = []
        seen = set()
        for rec in self.ins:
            for v in ([rec["dst"]] if rec["dst"] else []) + rec["srcs"]:
                if v["kind"] == "val" and v["id"] not in self.smem and v["id"] not in seen:
                    seen.add(v["id"])
                    dt = self.dtype[v["id"]]
                    if v["id"] in self.ell_val:
                        decl.append("  " + self.ctype(dt) + " " + ", ".join(
                            f"v{v['id']}_{k} = ({self.ctype(dt)})0" for k in range(self.ell_val[v["id"]].K)) + ";")
                        continue
                    decl.append(f"  {self.ctype(dt)} v{v['id']} = ({self.ctype(dt)})0;")
        lists = [self.ins[:self.n_pro], self.ins[self.n_pro:self.n_pro + self.n_step],
                 self.ins[self.n_pro + self.n_step:]]
        bodies = []
        self.lane_offs = set()
        # ring budget: 4 warps x D stages x (loads x 32 lanes) words of static shared memory
        nld = sum(1 for rec in lists[1] if rec["op"] in ("LDB", "LDU")
                  and rec["flags"] & isa.F_TSTEP and rec["n"] <= self.L and self.is_reg(rec["dst"]))
        dyn = (p.const_words + 4 * p.warp_words) * 4
        while self.pf_depth > 1 and (4 * self.pf_depth * nld * 128 > 40 * 1024
                                     or dyn + 4 * self.pf_depth * nld * 128 > 200 * 1024):
            self.pf_depth -= 1
        self.pf_stmts, self.pf_last = [], 0
        self.hoisted: Dict[str, str] = {}
        self.next_bases: Dict[int, str] = {}
        backward = p.kind == "backward"
        for li, lst in enumerate(lists):
            self.lines = []
            self.bases: Dict[int, str] = {}
            self.in_step = li == 1
            self.plan_trig(lst)
            for rec in lst:
                self.gen_ins(rec)
            head = list(self.bases.values())
            if li == 1 and self.pf_stmts:
                D = self.pf_depth
                refill = [f"  {{ const long long tn = t {'-' if backward else '+'} {D}; "
                          f"const bool more = tn >= 0 && tn < (long long)K.T;"] \
                    + list(self.next_bases.values()) \
                    + [f"    uint32_t* const ring_w = ring + slot * {len(self.pf_stmts) * 32}u;"] \
                    + self.pf_stmts + ["    rk_cp_async_commit(); }",
                                       f"  slot = (slot + 1u == {D}u) ? 0u : slot + 1u;"]
                self.lines[self.pf_last:self.pf_last] = refill
                head = [f"  rk_cp_async_wait<{D - 1}>();",
                        f"  const uint32_t* const ring_r = ring + slot * {len(self.pf_stmts) * 32}u;"] + head
            bodies.append("\n".join(head + self.lines))
        prefetch0, ring_decl = "", ""
        if self.pf_stmts:
            t0 = "(long long)K.T - 1" if backward else "0ll"
            sgn = "-" if backward else "+"
            nld = len(self.pf_stmts)
            prefetch0 = "\n".join(
                "  { const long long tn = " + t0 + f" {sgn} {k}; "
                "const bool more = tn >= 0 && tn < (long long)K.T;\n"
                + "\n".join(list(self.next_bases.values())
                            + [f"    uint32_t* const ring_w = ring + {k * nld * 32}u;"]
                            + self.pf_stmts) + "\n    rk_cp_async_commit();\n  }"
                for k in range(self.pf_depth))
            ring_decl = (f"  __shared__ uint32_t ring_all[{4 * self.pf_depth * nld * 32}];\n"
                         f"  uint32_t* const ring = ring_all + warp * {self.pf_depth * nld * 32}u + lane;\n"
                         f"  uint32_t slot = 0u;")
        hoist = [f"  const bool has{b} = K.buf[{b}] != nullptr; (void)has{b};"
                 for b in range(len(isa.BUFFERS))]
        hoist += [f"  const uint32_t lo{n} = b0u * {n}u + q * {n}u + le; (void)lo{n};"
                  for n in sorted(self.lane_offs)]
        loop = "for (long long t = (long long)K.T - 1; t >= 0; --t)" if backward else \
            "for (long long t = 0; t < (long long)K.T; ++t)"
        src = f"""// GENERATED by pyrddlgym_jax_b200/core/codegen.py (version {CODEGEN_VERSION}) -- do not edit.
// program: kind={p.kind} relaxed={p.relaxed} lanes={p.lanes} rpw={p.rpw} S={p.S} A={p.A} N={p.N}
//          instructions={p.n_ins} register-resident values={len(seen)} shared-memory values={len(self.smem)}
#include "rk_launch.h"
#include "rk_device.cuh"
{chr(10).join(self.ell_tables)}
// per-lane layout tables are read ONCE, before the step loop (volatile: ptxas would otherwise
// re-materialise the load inside the loop -- a lane-indexed, i.e. serialised, access)
__device__ __forceinline__ uint32_t rk_ld_tab(const uint32_t* p) {{
  uint32_t v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}}

extern "C" __global__ void {KERNEL_NAME}(const RkLaunch K) {{
  extern __shared__ __align__(16) uint32_t smem[];
  const uint32_t* prog = K.prog;
  for (int i = threadIdx.x; i < K.const_words; i += blockDim.x) smem[i] = prog[RK_HEADER_WORDS + i];
  __syncthreads();
  if (K.mparams != nullptr)
    for (int i = threadIdx.x; i < K.n_mparams; i += blockDim.x)
      smem[K.mparam_off + i] = __float_as_uint(K.mparams[i]);
  __syncthreads();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const long long b0 = gwarp * {p.rpw}ll;
  if (b0 >= K.B) return;
  const uint32_t nvalid = (uint32_t)min({p.rpw}ll, (long long)K.B - b0);
  const uint32_t q = lane / {p.lanes}u, le = lane - q * {p.lanes}u;
  const uint32_t qc = q < {p.rpw}u ? q : 0u;
  const uint32_t qL = qc * {p.lanes}u;
  const uint32_t* C = smem;
  uint32_t* R = smem + {p.const_words}u + warp * {p.warp_words}u;
  for (uint32_t i = lane; i < {p.rpw}u; i += 32u) R[{self.err_base}u + i] = 0u;
  __syncwarp();
  const uint32_t Bu = (uint32_t)K.B, b0u = (uint32_t)b0;   // host checks W*B < 2^31
  (void)qL; (void)qc; (void)le; (void)nvalid; (void)gwarp; (void)Bu; (void)b0u;
{chr(10).join(hoist)}
{chr(10).join(self.hoisted.values())}
{chr(10).join(decl)}
{ring_decl}
  {{ const long long t = 0; (void)t;
{bodies[0]}
  }}
{prefetch0}
  {loop} {{
{bodies[1]}
  }}
  {{ const long long t = 0; (void)t;
{bodies[2]}
  }}
}}
"""
        return src


def generate_source(prog: Program) -> str:
    if prog.lanes == 1:         # a lane owns a whole rollout: registers only, no shuffles
        from .codegen_tpr import generate_source as generate_tpr
        return generate_tpr(prog)
    return _Gen(prog).source()


def nvcc_flags(relaxed: bool = False) -> List[str]:
    """Exact programs: -fmad=false, every RDDL multiply and add rounds separately like the
    oracle (bit-exact bool / int contract).  Relaxed (train / adjoint) programs let ptxas
    contract a*b+c into FMAs, as XLA does for the reference: fewer instructions, results differ
    in the last bits and stay inside the 1e-5 / 1e-4 tolerances.  RK_FAST_RELAXED=0 disables
    it, RK_FMAD=1 forces contraction everywhere."""
    fast = relaxed and os.environ.get("RK_FAST_RELAXED", "1") != "0"
    fmad = "true" if (fast or os.environ.get("RK_FMAD", "0") == "1") else "false"
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
             f"-fmad={fmad}", "-cubin"]
    cap = os.environ.get("RK_MAXREG", "")
    if cap:                                  # A/B switch: register cap of the generated kernels
        flags.append(f"-maxrregcount={int(cap)}")
    return flags


def compile_cubin(prog: Program, verbose: bool = False) -> Tuple[str, str]:
    """Generate + compile the specialised kernel (cached by content hash under csrc/jit/)."""
    src = generate_source(prog)
    deps = b""
    for h in ("rk_launch.h", "rk_device.cuh", "rk_isa.h"):
        with open(os.path.join(CSRC, h), "rb") as f:
            deps += f.read()
    flags = nvcc_flags(bool(prog.relaxed))
    key = hashlib.sha1(src.encode() + deps + " ".join(flags).encode()).hexdigest()[:16]
    os.makedirs(JIT_DIR, exist_ok=True)
    cu = os.path.join(JIT_DIR, f"rk_{key}.cu")
    cubin = os.path.join(JIT_DIR, f"rk_{key}.cubin")
    if not os.path.exists(cubin):
        # several ranks may generate the same kernel at once: private temporaries, atomic rename
        tag = f".{os.getpid()}.tmp"
        with open(cu + tag, "w") as f:
            f.write(src)
        os.replace(cu + tag, cu)
        cmd = [os.environ.get("NVCC", "nvcc"), *flags, "-I", CSRC,
               *(["-Xptxas", "-v"] if verbose else []), cu, "-o", cubin + tag]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on generated kernel:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
        os.replace(cubin + tag, cubin)
    return cubin, cu
