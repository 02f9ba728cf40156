"""Assembler: SSA step graph -> binary rollout program for csrc/rk_interp.cu.

A program has three instruction lists -- prologue (once), step (once per decision epoch,
t ascending for forward programs, descending for adjoint programs), epilogue (once) -- the
flat counterpart of ``compile_rollouts`` (compiler.py:621-666): the scan body of
compiler.py:561-593 is the step list, and the backward pass XLA derives from
``jax.value_and_grad`` (planner.py:2873) is a second program whose step list reloads s_t
from the state tape, recomputes the step and runs the adjoint instructions emitted by
core/graph.py:differentiate.

Register file layout (per warp, 32-bit words in shared memory): a value with ``n`` elements
per rollout occupies RPW*n consecutive words as [q][e] (RPW = rollouts per warp, lane = q*L+e
for e < L); rollout-invariant values occupy n words.  Constants, gather tables and model
parameters live in one CTA-wide constant region.
"""
from __future__ import annotations

import os as _os
import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import isa
from .graph import Graph, V, Map, apply_map, differentiate, sparsify
from .lowering import Lowerer, StepGraph, TV

HEADER_WORDS = 32


class ProgramError(Exception):
    pass


@dataclass
class Ins:
    op: str
    dst: Optional[V] = None
    srcs: List[Tuple[V, Map]] = field(default_factory=list)
    n: int = 0
    flags: int = 0
    aux0: int = 0
    aux1: int = 0
    tables: Dict[str, np.ndarray] = field(default_factory=dict)
    off: int = 0                # memory offset (words per rollout / per step)
    buf: int = 0


def root(v: V) -> V:
    while v.op == "sg":
        v = v.args[0][0]
    return v


@dataclass
class Program:
    blob: bytes
    kind: str                   # 'forward' | 'backward'
    relaxed: bool
    rpw: int
    lanes: int
    warp_words: int
    const_words: int
    n_ins: Tuple[int, int, int]
    S: int
    A: int
    N: int
    state_layout: Dict[str, Tuple[int, int, str]]      # fluent -> (offset, n, dtype)
    param_layout: Dict[str, Tuple[int, int, str]]      # action fluent -> (offset, n, range)
    noise_layout: List[Tuple[str, Tuple[int, ...], int, int, Any]]  # kind, shape, n, offset, extra
    mparam_keys: List[Tuple[Any, str]]
    mparam_defaults: np.ndarray
    traj_layout: Dict[str, Tuple[int, int, str]] = field(default_factory=dict)
    TRAJ: int = 0
    tape_words: int = 0         # words per rollout-step of the tape (S, or S + taped values)
    static_err: int = 0
    stats: Dict[str, Any] = field(default_factory=dict)
    meta: Dict[str, Any] = field(default_factory=dict)    # decoded instructions for codegen


class Assembler:
    def __init__(self):
        self.const_image: List[np.ndarray] = [np.array([0.0, 1.0], dtype=np.float32).view(np.uint32)]
        self.const_words = 2
        self._const_cache: Dict[bytes, int] = {}
        self.addr: Dict[int, Tuple[int, int]] = {}      # root V id -> (base, sq)

    def add_const(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.dtype.kind == "f":
            words = arr.astype(np.float32).view(np.uint32)
        else:
            words = arr.astype(np.int32).view(np.uint32)
        key = words.tobytes()
        off = self._const_cache.get(key)
        if off is None:
            off = self.const_words
            self.const_image.append(words.ravel())
            self.const_words += words.size
            self._const_cache[key] = off
        return off

    def table(self, arr: np.ndarray) -> int:
        return self.add_const(np.asarray(arr, dtype=np.int32))


class _Alloc:
    """First-fit interval allocator over a 1-D word space."""

    def __init__(self, start: int):
        self.free: List[List[int]] = [[start, 1 << 30]]
        self.high = start

    def alloc(self, size: int) -> int:
        for iv in self.free:
            if iv[1] - iv[0] >= size:
                base = iv[0]
                iv[0] += size
                if iv[0] == iv[1]:
                    self.free.remove(iv)
                self.high = max(self.high, base + size)
                return base
        raise ProgramError("register space exhausted")

    def release(self, base: int, size: int):
        self.free.append([base, base + size])
        self.free.sort()
        merged = [self.free[0]]
        for iv in self.free[1:]:
            if iv[0] <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], iv[1])
            else:
                merged.append(iv)
        self.free = merged


# =====================================================================================

@dataclass
class PlanSpec:
    """How actions are produced inside the step (the policy part of compiler.py:552-557)."""
    kind: str = "slp"                      # 'slp' | 'external'
    wrap_sigmoid: bool = True
    wrap_non_bool: bool = False
    sigmoid_weight: float = 1.0
    bounds: Optional[Dict[str, Tuple[np.ndarray, np.ndarray]]] = None
    # stochastic straight-line plan (planner.py:811-834): logistic noise on the boolean logits,
    # mu + exp(clip(log_sigma)) * noise for the others; the log_sigma rows follow the action rows
    stochastic: bool = False
    sigma_range: Tuple[float, float] = (1e-5, 1e3)
    action_noise: str = "normal"
    # softmax over ALL boolean action parameters of a step instead of per-element sigmoids
    # (planner.py:764-808, 848-853), used when max-nondef-actions = 1 < number of boolean actions;
    # fluents whose no-op value is true receive 1 - softmax
    pooled_softmax: bool = False
    softmax_weight: float = 1.0
    noop_true: Tuple[str, ...] = ()


def param_layout(model, spec: "PlanSpec") -> Tuple[Dict[str, Tuple[int, int, str]], int]:
    """Row layout of the plan parameters: the action layout, followed (stochastic plans) by one
    log_sigma block per non-boolean action fluent, keyed '<fluent>/log_sigma'."""
    layout, off = action_layout(model)
    layout = dict(layout)
    if spec.kind == "slp" and spec.stochastic:
        for name, (_, n, prange) in list(layout.items()):
            if prange != "bool":
                layout[name + "/log_sigma"] = (off, n, "real")
                off += n
    return layout, off


def action_layout(model) -> Tuple[Dict[str, Tuple[int, int, str]], int]:
    layout, off = {}, 0
    for name in model.action_fluents:
        n = model.variable_size(name)
        layout[name] = (off, n, model.variable_ranges[name])
        off += n
    return layout, off


def state_layout(model, relaxed: bool) -> Tuple[Dict[str, Tuple[int, int, str]], int]:
    layout, off = {}, 0
    for name in model.state_fluents:
        n = model.variable_size(name)
        dt = "f" if relaxed else {"real": "f", "int": "i", "bool": "b"}.get(
            model.variable_ranges[name], "i")
        layout[name] = (off, n, dt)
        off += n
    return layout, off


def _bound_action(g, p: V, lo: np.ndarray, hi: np.ndarray, n: int) -> V:
    """_jax_bound_action (planner.py:620-628): sigmoid wrap for two-sided boxes, softplus wraps for
    one-sided ones, identity when unbounded; the branch of every element is known at build time."""
    lo = np.asarray(lo, np.float32).ravel()
    hi = np.asarray(hi, np.float32).ravel()
    fl, fh = np.isfinite(lo), np.isfinite(hi)
    lo0, hi0 = np.where(fl, lo, 0).astype(np.float32), np.where(fh, hi, 0).astype(np.float32)
    C = lambda a: g.const(np.asarray(a, np.float32))       # noqa: E731
    O = lambda op, *a: g.op(op, [(x, None) for x in a], n)  # noqa: E731

    def softplus(x):            # log1p(exp(-|x|)) + relu(x)
        return O("ADD", O("LOG1P", O("EXP", O("NEG", O("ABS", x)))), O("RELU", x))
    out = p
    if np.any(~fl & fh):
        f_hi = O("SUB", C(hi0), softplus(O("NEG", p)))
        out = g.op("SELECT", [(g.const((~fl & fh).astype(np.int32), "b"), None), (f_hi, None),
                              (out, None)], n, "f")
    if np.any(fl & ~fh):
        f_lo = O("ADD", C(lo0), softplus(p))
        out = g.op("SELECT", [(g.const((fl & ~fh).astype(np.int32), "b"), None), (f_lo, None),
                              (out, None)], n, "f")
    if np.any(fl & fh):
        f_both = O("ADD", C(lo0), O("MUL", C(hi0 - lo0), O("SIGMOID", p)))
        out = g.op("SELECT", [(g.const((fl & fh).astype(np.int32), "b"), None), (f_both, None),
                              (out, None)], n, "f")
    return out


def build_policy_actions(low: Lowerer, spec: PlanSpec, train: bool) -> Tuple[Dict[str, V], Dict[str, V]]:
    """Action values fed to the transition + the parameter leaves they come from.

    train=True : JaxStraightLinePlan train policy, planner.py:786-838
    train=False: test policy, planner.py:842-868
    """
    m, g = low.m, low.g
    layout, _ = action_layout(m)
    playout, _ = param_layout(m, spec)
    stoch = spec.kind == "slp" and spec.stochastic and train
    actions, leaves = {}, {}

    def policy_noise(kind: str, name: str, n: int) -> V:     # drawn before the model's own noise
        from .lowering import NoiseSlot
        v = g.leaf("noise", n, "f", slot=len(low.sg.noise))
        low.sg.noise.append(NoiseSlot(kind, tuple(m.variable_shape(name)), n, v, None))
        return v
    pooled: Dict[str, V] = {}
    if spec.kind == "slp" and spec.pooled_softmax:
        # softmax_i(w p_i) over the pooled boolean parameters: per-fluent maxima / sums are
        # 1-element reductions, combined across fluents and broadcast back
        plist = []
        for name, (off, n, prange) in layout.items():
            if prange == "bool":
                p = g.leaf("param_in", n, "f", uniform=True, off=off, fluent=name)
                leaves[name] = p
                plist.append((name, n, p))
        if stoch:                                     # planner.py:797-805: logits + Gumbel noise
            plist = [(name, n, g.op("ADD", [(p, None), (policy_noise("gumbel", name, n), None)], n))
                     for name, n, p in plist]
        w = g.const(np.float32(spec.softmax_weight))
        z = {name: g.op("MUL", [(w, "b"), (p, None)], n) for name, n, p in plist}
        whole = lambda n: (np.array([0, n], np.int64), np.arange(n, dtype=np.int64))   # noqa: E731
        top = None
        for name, n, _ in plist:
            m1 = g.reduce("RMAX", z[name], *whole(n))
            top = m1 if top is None else g.op("MAX", [(top, None), (m1, None)], 1)
        top = g.sg(top)
        ex = {name: g.op("EXP", [(g.op("SUB", [(z[name], None), (top, "b")], n), None)], n)
              for name, n, _ in plist}
        tot = None
        for name, n, _ in plist:
            s1 = g.reduce("RSUM", ex[name], *whole(n))
            tot = s1 if tot is None else g.op("ADD", [(tot, None), (s1, None)], 1)
        for name, n, _ in plist:
            a = g.op("DIV", [(ex[name], None), (tot, "b")], n)
            if name in spec.noop_true:
                a = g.op("SUB", [(g.const(np.float32(1.0)), "b"), (a, None)], n)
            pooled[name] = a
    for name, (off, n, prange) in layout.items():
        if spec.kind == "external":
            dt = low.dtype_of_range(prange)
            leaf = g.leaf("act_in", n, dt, uniform=False, off=off, fluent=name)
            leaves[name] = leaf
            actions[name] = leaf
            continue
        if name in pooled:
            a = pooled[name]
            if not train:                             # planner.py:852: softmax output > 0.5
                a = g.op("GT", [(a, None), (g.const(np.float32(0.5)), "b")], n)
                want = low.dtype_of_range(prange)
                if a.dtype != want and want == "f":
                    a = g.op("I2F", [(a, None)], n)
            actions[name] = a
            continue
        p = g.leaf("param_in", n, "f", uniform=True, off=off, fluent=name)
        leaves[name] = p
        is_bool = prange == "bool"
        if train:
            if is_bool and spec.wrap_sigmoid:         # planner.py:741-745, 818
                w = g.leaf("mparam", 1, "f", uniform=True, key=("policy", name),
                           default=float(spec.sigmoid_weight), slot=len(low.sg.mparams))
                low.sg.mparams.append((("policy", name), float(spec.sigmoid_weight), w))
                logit = p
                if stoch:                             # planner.py:811-819
                    logit = g.op("ADD", [(p, None), (policy_noise("logistic", name, n), None)], n)
                a = g.op("SIGMOID", [(g.op("MUL", [(w, "b"), (logit, None)], n), None)], n)
            elif not is_bool:
                a = p
                if stoch:                             # planner.py:825-834
                    loff = playout[name + "/log_sigma"][0]
                    ls = g.leaf("param_in", n, "f", uniform=True, off=loff,
                                fluent=name + "/log_sigma")
                    leaves[name + "/log_sigma"] = ls
                    lo_s, hi_s = (np.float32(np.log(x)) for x in spec.sigma_range)
                    ls = g.op("MIN", [(g.op("MAX", [(ls, None), (g.const(lo_s), "b")], n), None),
                                      (g.const(hi_s), "b")], n)
                    z = policy_noise(spec.action_noise, name, n)
                    a = g.op("ADD", [(p, None), (g.op("MUL", [(g.op("EXP", [(ls, None)], n), None),
                                                              (z, None)], n), None)], n)
                if spec.wrap_non_bool:                # planner.py:755-762
                    a = _bound_action(g, a, *spec.bounds[name], n)
            else:
                a = p
            if not low.relaxed and a.dtype != low.dtype_of_range(prange):
                raise ProgramError("train policy requires a relaxed (all-REAL) model")
            actions[name] = a
        else:
            if is_bool:                               # planner.py:856
                thr = 0.0 if spec.wrap_sigmoid else 0.5
                a = g.op("GT", [(p, None), (g.const(np.float32(thr)), "b")], n)
            else:                                     # planner.py:860-865
                lo, hi = spec.bounds[name]
                if spec.wrap_non_bool:                # planner.py:862
                    p = _bound_action(g, p, lo, hi, n)
                lo = g.const(np.asarray(lo, dtype=np.float32).ravel())
                hi = g.const(np.asarray(hi, dtype=np.float32).ravel())
                a = g.op("MIN", [(g.op("MAX", [(p, None), (lo, None)], n), None), (hi, None)], n)
                if prange != "real":
                    a = g.op("F2I", [(g.op("ROUND", [(a, None)], n), None)], n, "i")
            want = low.dtype_of_range(prange)
            if a.dtype != want:
                if want == "f":
                    a = g.op("I2F", [(a, None)], n)
                else:
                    raise ProgramError(f"cannot feed {a.dtype} action into {want} fluent <{name}>")
            actions[name] = a
    return actions, leaves


class ProgramBuilder:
    """Builds forward / backward programs for one (model, relaxation, plan) triple."""

    def __init__(self, model, relax: Optional[Dict[str, Any]], spec: PlanSpec, train_policy: bool,
                 log_fluents: Sequence[str] = (), batch_hint: int = 0):
        self.batch_hint = batch_hint
        # the specialised kernels (core/codegen.py) have almost no per-instruction overhead
        import os as _os
        if _os.environ.get("RK_SPECIALIZE", "1") != "0":
            self.INS_OVERHEAD = 0.5
        self.model = model
        self.relax = relax
        self.spec = spec
        self.low = Lowerer(model, relax)
        self.g = self.low.g
        self.actions, self.param_leaves = build_policy_actions(self.low, spec, train_policy)
        self.sg: StepGraph = self.low.build(self.actions)
        self.slayout, self.S = state_layout(model, relax is not None)
        # straight-line plans: A is the width of one parameter row (actions + log_sigma blocks)
        self.alayout, self.A = param_layout(model, spec) if spec.kind == "slp" \
            else action_layout(model)
        off = 0
        for slot in self.sg.noise:
            slot.offset = off
            off += slot.n
        self.N = off
        self.log_fluents = list(log_fluents)

    # ---------------------------------------------------------------------------------
    def _needed(self, roots: Sequence[V], stop: Optional[set] = None) -> List[V]:
        """Values reachable from ``roots``; the operands of values in ``stop`` (taped values,
        which are loaded instead of computed) are not followed."""
        seen, stack = set(), [r for r in roots]
        while stack:
            v = stack.pop()
            if v.id in seen:
                continue
            seen.add(v.id)
            if stop is not None and root(v).id in stop and v.op != "sg":
                continue
            for a, _ in v.args:
                stack.append(a)
        return [v for v in self.g.nodes if v.id in seen]

    def _compute_ins(self, nodes: Sequence[V], tstep_params: bool = True) -> List[Ins]:
        """One instruction per value that needs computing, in topological order."""
        out = []
        taped = getattr(self, "_load_taped", None) or {}
        for v in nodes:
            if v.op in ("const", "mparam", "sg", "state_in", "persist"):
                continue
            if v.id in taped:          # adjoint program: forward value read back from the tape
                out.append(Ins("LDB", dst=v, n=v.n, flags=isa.F_TSTEP, buf=isa.BUF["TAPE"],
                               off=self.S + taped[v.id]))
                continue
            if v.op == "param_in":
                out.append(Ins("LDU", dst=v, n=v.n, flags=isa.F_UNIFORM | isa.F_TSTEP,
                               buf=isa.BUF["PARAMS"], off=v.attrs["off"]))
            elif v.op == "act_in":
                out.append(Ins("LDB", dst=v, n=v.n, flags=isa.F_TSTEP, buf=isa.BUF["ACT"],
                               off=v.attrs["off"]))
            elif v.op == "noise":
                slot = self.sg.noise[v.attrs["slot"]]
                out.append(Ins("LDB", dst=v, n=v.n, flags=isa.F_TSTEP, buf=isa.BUF["NOISE"],
                               off=slot.offset))
            elif v.op == "disc_in":
                out.append(Ins("LDU", dst=v, n=1, flags=isa.F_UNIFORM | isa.F_TSTEP,
                               buf=isa.BUF["DISC"], off=0))
            elif v.op in ("RSUM", "RPROD", "RMIN", "RMAX", "RISUM", "RIPROD", "RIMIN", "RIMAX",
                          "RARGMAX", "RARGMIN", "RPRODX"):
                tables = {"ptr": v.attrs["ptr"], "idx": v.attrs["idx"]}
                if v.op == "RPRODX":
                    tables["seg"] = v.attrs["seg"]
                out.append(Ins(v.op, dst=v, srcs=list(v.args), n=v.n, tables=tables,
                               flags=isa.F_UNIFORM if v.uniform else 0))
            else:
                out.append(Ins(v.op, dst=v, srcs=list(v.args), n=v.n,
                               flags=isa.F_UNIFORM if v.uniform else 0))
        return out

    # ---------------------------------------------------------------------------------
    def forward(self, write_tape: bool, gamma: float = 1.0, final_state: bool = True) -> Program:
        g, sg = self.g, self.sg
        ret = g.leaf("persist", 1, "f")
        pro: List[Ins] = []
        for name, (off, n, _) in self.slayout.items():
            pro.append(Ins("LDB", dst=sg.state_in[name], n=n, buf=isa.BUF["S0"], off=off))
        pro.append(Ins("MOV", dst=ret, srcs=[(g.const(np.float32(0.)), "b")], n=1))

        roots = [sg.reward] + list(sg.next_state.values()) + [c for _, c in sg.errs] \
            + [sg.fluents[f] if f in sg.fluents else sg.action_in[f] for f in self.log_fluents]
        step: List[Ins] = []
        extras = dict(getattr(self, "tape_extra", {})) if write_tape else {}
        self._fwd_built = True
        if write_tape:
            for name, (off, n, _) in self.slayout.items():
                step.append(Ins("STB", srcs=[(sg.state_in[name], None)], n=n, flags=isa.F_TSTEP,
                                buf=isa.BUF["TAPE"], off=off))
            roots += [v for v, _ in extras.values()]
        disc = None
        if gamma != 1.0:
            disc = g.leaf("disc_in", 1, "f", uniform=True)
            roots.append(disc)
        if not extras:      # (with a planned tape the adjoint builder already pruned the graph)
            self.sparsity = sparsify(g, roots) if self.SPARSIFY else {}
        self._tape_words = self.S + sum(v.n for v, _ in extras.values())
        step += self._compute_ins(self._needed(roots))
        for v, off in extras.values():      # forward values the adjoint reads back (full tape)
            step.append(Ins("STB", srcs=[(v, None)], n=v.n, flags=isa.F_TSTEP,
                            buf=isa.BUF["TAPE"], off=self.S + off))
        step.append(Ins("STB", srcs=[(sg.reward, None)], n=1, flags=isa.F_TSTEP,
                        buf=isa.BUF["REWARD"], off=0))
        if disc is not None:       # planner.py:2754-2757
            step.append(Ins("FMA", dst=ret, srcs=[(sg.reward, None), (disc, "b"), (ret, None)], n=1))
        else:
            step.append(Ins("ADD", dst=ret, srcs=[(ret, None), (sg.reward, None)], n=1))
        for bit, cond in sg.errs:
            step.append(Ins("ERRCHK", srcs=[(cond, None)], n=cond.n, aux0=bit))
        step.append(Ins("STERR", n=1, flags=isa.F_TSTEP, buf=isa.BUF["ERR"]))
        traj_layout, toff = {}, 0
        for f in self.log_fluents:
            v = sg.fluents[f] if f in sg.fluents else sg.action_in[f]
            step.append(Ins("STB", srcs=[(v, None)], n=v.n, flags=isa.F_TSTEP,
                            buf=isa.BUF["TRAJ"], off=toff))
            traj_layout[f] = (toff, v.n, v.dtype)
            toff += v.n
        self._traj_words = toff
        state_roots = {v.id for v in sg.state_in.values()}
        moves = []
        for name in self.slayout:
            nxt = sg.next_state[name]
            if root(nxt) is sg.state_in[name]:
                continue
            if root(nxt).id in state_roots:      # x' = y: stage through a temporary
                tmp = g.add(V("MOV", [(nxt, None)], n=nxt.n, dtype=nxt.dtype))
                step.append(Ins("MOV", dst=tmp, srcs=[(nxt, None)], n=nxt.n))
                nxt = tmp
            moves.append(Ins("MOV", dst=sg.state_in[name], srcs=[(nxt, None)],
                             n=sg.state_in[name].n))
        step += moves
        epi: List[Ins] = [Ins("STB", srcs=[(ret, None)], n=1, buf=isa.BUF["RETURNS"], off=0)]
        if final_state:
            for name, (off, n, _) in self.slayout.items():
                epi.append(Ins("STB", srcs=[(sg.state_in[name], None)], n=n,
                               buf=isa.BUF["FINAL"], off=off))
        persistent = list(sg.state_in.values()) + [ret]
        prog = self._assemble("forward", pro, step, epi, persistent)
        prog.traj_layout, prog.TRAJ = traj_layout, toff
        return prog

    # ---------------------------------------------------------------------------------
    def backward(self, gamma: float = 1.0, grad_s0: bool = False,
                 state_ct_in: bool = False) -> Program:
        """Adjoint program: d(sum_b gret[b] * return_b) / d(policy parameters).

        state_ct_in: start from a supplied cotangent of the final state (buffer GSIN) instead
        of zero -- used when the horizon is walked one step per launch with a closed-loop
        policy between the launches (deep reactive policy, planner.py:1468-1500)."""
        if self.relax is None:
            raise ProgramError("the adjoint program needs a relaxed (all-REAL) model")
        g, sg = self.g, self.sg
        gret = g.leaf("persist", 1, "f")
        carried = {name: g.leaf("persist", n, "f") for name, (_, n, _) in self.slayout.items()}
        pro: List[Ins] = [Ins("LDB", dst=gret, n=1, buf=isa.BUF["GRET"], off=0)]
        zero = g.const(np.float32(0.))
        for name, c in carried.items():
            if state_ct_in:
                off, n, _ = self.slayout[name]
                pro.append(Ins("LDB", dst=c, n=n, buf=isa.BUF["GSIN"], off=off))
            else:
                pro.append(Ins("MOV", dst=c, srcs=[(zero, "b")], n=c.n))

        if gamma != 1.0:
            disc = g.leaf("disc_in", 1, "f", uniform=True)
            gr = g.op("MUL", [(gret, None), (disc, "b")], 1)
        else:
            gr = gret
        seeds = [(sg.reward, gr)] + [(sg.next_state[name], carried[name]) for name in self.slayout]
        if self.SPARSIFY:       # differentiate the pruned forward graph (RPRODX tables follow it)
            sparsify(g, [sg.reward] + list(sg.next_state.values()))
        fwd_nodes = self._needed([sg.reward] + list(sg.next_state.values()))
        wrt = list(sg.state_in.values()) + list(self.param_leaves.values())
        n_before = len(g.nodes)
        cts = differentiate(g, fwd_nodes, seeds, wrt)

        step: List[Ins] = []
        for name, (off, n, _) in self.slayout.items():
            step.append(Ins("LDB", dst=sg.state_in[name], n=n, flags=isa.F_TSTEP,
                            buf=isa.BUF["TAPE"], off=off))
        outs = [cts[v.id] for v in wrt if v.id in cts]
        # prune with the forward roots kept alive too: the forward program may be assembled
        # from this graph afterwards (it must store exactly what this program loads)
        fwd_roots = [sg.reward] + list(sg.next_state.values()) + [c for _, c in sg.errs]
        self.sparsity = sparsify(g, outs + fwd_roots) if self.SPARSIFY else {}
        self._load_taped = self._plan_tape(outs, fwd_roots, state_ct_in)
        step_nodes = self._needed(outs, stop=set(self._load_taped))
        used_state = {root(a).id for v in step_nodes for a, _ in v.args} | \
            {root(o).id for o in outs}
        step = [i for i in step if not (i.op == "LDB" and i.buf == isa.BUF["TAPE"]
                                        and root(i.dst).id not in used_state)]
        step += self._compute_ins(step_nodes)
        gbuf = isa.BUF["GRADP"] if self.spec.kind == "slp" else isa.BUF["GACT"]
        for name, leaf in self.param_leaves.items():
            off, n, _ = self.alayout[name]
            ct = cts.get(leaf.id)
            src = (ct, None) if ct is not None else (zero, "b")
            if self.spec.kind == "slp":
                step.append(Ins("STG", srcs=[src], n=n, flags=isa.F_TSTEP | isa.F_UNIFORM,
                                buf=gbuf, off=off))
            else:
                step.append(Ins("STB", srcs=[src], n=n, flags=isa.F_TSTEP, buf=gbuf, off=off))
        for name, c in carried.items():
            ct = cts.get(sg.state_in[name].id)
            src = (ct, None) if ct is not None else (zero, "b")
            step.append(Ins("MOV", dst=c, srcs=[src], n=c.n))
        epi: List[Ins] = []
        if grad_s0:
            for name, (off, n, _) in self.slayout.items():
                epi.append(Ins("STB", srcs=[(carried[name], None)], n=n, buf=isa.BUF["GS0"], off=off))
        persistent = list(sg.state_in.values()) + [gret] + list(carried.values())
        self._traj_words = 0
        self._tape_words = self.S + sum(v.n for v, _ in self.tape_extra.values())
        prog = self._assemble("backward", pro, step, epi, persistent)
        self._load_taped = None
        prog.stats["adjoint_nodes"] = len(g.nodes) - n_before
        prog.stats["taped_values"] = len(self.tape_extra)
        return prog

    # ---------------------------------------------------------------------------------
    # measured (Quadcopter, B=4096, T=100): "full" makes the adjoint 10% faster (no recompute) but
    # the forward 40% slower (19 extra stores per lane-step), a net loss -> minimal tape by default
    TAPE_MODE = _os.environ.get("RK_TAPE", "state")     # state | full | auto

    def _plan_tape(self, outs: Sequence[V], fwd_roots: Sequence[V], external: bool) -> Dict[int, int]:
        """Decide what the adjoint reads back instead of recomputing.

        'state' : only s_t (S words per rollout-step), the step is recomputed -- the minimum of
                  HBM traffic, right when the batch fills the machine;
        'full'  : additionally every per-rollout forward value the reverse instructions read (the
                  "frontier"), no recompute -- right when the launch is latency-bound (few warps)
                  and the extra words are cheap next to the saved instructions.
        Returns {value id: word offset after the state words}; also sets self.tape_extra for the
        forward program, which must then be assembled AFTER this adjoint program."""
        self.tape_extra: Dict[int, Tuple[V, int]] = {}
        if self.TAPE_MODE == "state" or external or getattr(self, "_fwd_built", False):
            return {}
        fwd_ids = {v.id for v in self._needed(fwd_roots)}
        adj = [v for v in self._needed(outs) if v.id not in fwd_ids]
        leaf = ("const", "mparam", "state_in", "persist", "param_in", "act_in", "noise", "disc_in")
        frontier: Dict[int, V] = {}
        for v in adj:
            for a, _ in v.args:
                r = root(a)
                if r.id in fwd_ids and r.op not in leaf and r.const is None and not r.uniform:
                    frontier[r.id] = r
        words = sum(r.n for r in frontier.values())
        if self.TAPE_MODE != "full":
            B = max(int(self.batch_hint or 4096), 1)
            widest = max([1] + [v.n for v in self._needed(outs)])
            lanes = min(32, 1 << max(0, (widest - 1).bit_length()))
            nwarps = -(-B // max(1, 32 // lanes))
            latency_bound = nwarps <= 4 * 148 * 4            # under ~4 warps per sub-partition
            if not (latency_bound and words <= 4 * self.S):
                return {}
        off, plan = 0, {}
        for r in sorted(frontier.values(), key=lambda v: v.id):
            self.tape_extra[r.id] = (r, off)
            plan[r.id] = off
            off += r.n
        return plan

    # ---------------------------------------------------------------------------------
    # per-instruction fixed cost (fetch, decode, dispatch) in units of one element pass
    SPARSIFY = _os.environ.get("RK_SPARSIFY", "1") != "0"
    INS_OVERHEAD = 4.0
    MULTIPASS_PENALTY = 16.0       # measured: MarsRover L=4 (smem, 8 passes) 8.0 ms vs L=32 (registers) 1.7 ms
    SMEM_WORDS = 56 * 1024            # ~224 KB of shared memory per SM, in words
    FREE_WARPS_PER_SM = 8             # warps/SM that overlap almost perfectly (latency-bound)

    def _assemble(self, kind: str, pro: List[Ins], step: List[Ins], epi: List[Ins],
                  persistent: Sequence[V]) -> Program:
        """Try every lanes-per-rollout choice and keep the one with the lowest estimated
        time for the batch size this builder was given."""
        B = max(int(getattr(self, "batch_hint", 0) or 4096), 1)
        best, best_t, last_err = None, None, None
        import os as _os
        forced = int(_os.environ.get("RK_LANES", "0") or 0)     # tuning / experiments only
        for L in sorted({32 // r for r in range(1, 33)}):
            if forced and L != forced:
                continue
            try:
                prog = self._assemble_for(kind, pro, step, epi, persistent, L)
            except ProgramError as e:
                last_err = e
                continue
            # specialised kernels keep single-pass values (n <= L) in registers; multi-pass
            # values live in shared memory and cost an order of magnitude more per pass -- unless the
            # code generator can give them a register layout of K slots per lane (codegen._EllFamily)
            mp = self.MULTIPASS_PENALTY if self.INS_OVERHEAD < 1.0 else 1.0
            ell = self._ell_slots(prog) if self.INS_OVERHEAD < 1.0 else {}
            unit = [float(ell[k]) if k in ell else
                    (-(-max(i.n, 1) // L)) * (mp if i.n > L else 1.0) for k, i in enumerate(step)]
            cost = sum(u + self.INS_OVERHEAD for u in unit)
            prog.meta["unit_costs"] = unit
            nwarps = -(-B // prog.rpw)
            resident = max(1, min(64, self.SMEM_WORDS // max(prog.warp_words, 1)))
            conc = 148 * min(resident, self.FREE_WARPS_PER_SM)
            t = cost * max(1.0, nwarps / conc)
            if best is None or t < best_t * 0.999:
                best, best_t = prog, t
        if best is None:
            raise last_err
        # thread-per-rollout kernels (lanes = 1, core/codegen_tpr.py): a warp carries 32 rollouts and
        # every value lives in registers, so the step costs ~1.4 SASS instructions per ELEMENT instead
        # of 2 - 4 per instruction-pass of an L-lane layout (measured on the MarsRover adjoint: 54
        # against 332 per rollout-step) -- but the launch has 32 / rpw times fewer warps.  Estimated
        # time = issue slots per scheduler, a lone warp issuing at ipc < 1.
        if self.INS_OVERHEAD < 1.0 and best.lanes != 1 and not forced \
                and _os.environ.get("RK_TPR", "1") != "0":
            try:
                tpr = self._assemble_for(kind, pro, step, epi, persistent, 1)
            except ProgramError:
                tpr = None
            if tpr is not None:
                def est(prog, sass, ipc):
                    wps = (-(-B // prog.rpw)) / (148 * 4)
                    return sass * max(1.0 / ipc, wps)
                L = best.lanes
                sass_multi = 2.2 * sum(best.meta["unit_costs"])
                sass_tpr = sum((10.0 if i.op == "QSUM" else 1.4) * max(i.n, 1) for i in step)
                live = tpr.warp_words / 32.0                  # scalar values alive at once, per lane
                if live > 320:      # measured: Wildfire forward, 441 live scalars, spills and loses
                    sass_tpr *= 1.0 + (live - 320.0) / 40.0
                if est(tpr, sass_tpr, 0.7) < 0.8 * est(best, sass_multi, 0.4):
                    best = tpr
        return best

    @staticmethod
    def _ell_slots(prog: Program) -> Dict[int, int]:
        """{index in the step list: register slots per lane} of the instructions the code generator keeps
        in registers although they are wider than the lane group."""
        from .codegen import _Gen
        try:
            gen = _Gen(prog)
        except Exception:
            return {}
        n_pro, n_step, _ = prog.n_ins
        out = {}
        for k, rec in enumerate(prog.meta["ins"][n_pro:n_pro + n_step]):
            fam = gen.ell_rec.get(id(rec))
            if fam is not None:
                out[k] = fam.K
        return out

    def _assemble_for(self, kind: str, pro: List[Ins], step: List[Ins], epi: List[Ins],
                      persistent: Sequence[V], L: int) -> Program:
        asm = Assembler()
        sg = self.sg
        rpw = 32 // L

        # ---- constant region: model parameters first, then constants / tables on demand
        n_mp = len(sg.mparams)
        mp_defaults = np.array([d for _, d, _ in sg.mparams], dtype=np.float32)
        mparam_off = asm.const_words
        if n_mp:
            asm.const_image.append(mp_defaults.view(np.uint32).copy())
            asm.const_words += n_mp
        for i, (_, _, v) in enumerate(sg.mparams):
            asm.addr[v.id] = ((mparam_off + i) | isa.CONST_BIT, 0)

        def place_const(v: V):
            if v.id not in asm.addr:
                asm.addr[v.id] = (asm.add_const(v.const) | isa.CONST_BIT, 0)

        # ---- persistent registers
        alloc = _Alloc(0)

        def size_of(v: V) -> int:
            return v.n if v.uniform else v.n * rpw

        err_base = alloc.alloc(rpw)                # per-rollout error accumulators
        for v in persistent:
            asm.addr[v.id] = (alloc.alloc(size_of(v)), 0 if v.uniform else v.n)
        persist_high = alloc.high
        warp_words = persist_high

        encoded: List[np.ndarray] = []
        records: List[Dict[str, Any]] = []
        counts = []
        for lst in (pro, step, epi):
            al = _Alloc(persist_high)
            # liveness
            last_use: Dict[int, int] = {}
            for i, ins in enumerate(lst):
                for a, _ in ins.srcs:
                    last_use[root(a).id] = i
            live: Dict[int, Tuple[int, int]] = {}
            for i, ins in enumerate(lst):
                for a, _ in ins.srcs:
                    r = root(a)
                    if r.const is not None:
                        place_const(r)
                    elif r.id not in asm.addr:
                        raise ProgramError(f"value {r} used before definition in {kind} program")
                d = ins.dst
                if d is not None:
                    r = root(d)
                    if r.id not in asm.addr:
                        base = al.alloc(size_of(r))
                        asm.addr[r.id] = (base, 0 if r.uniform else r.n)
                        live[r.id] = (base, size_of(r))
                        if r.id not in last_use:
                            last_use[r.id] = i
                encoded.append(self._encode(asm, ins))
                records.append(self._record(asm, ins, encoded[-1]))
                for rid in [rid for rid, lu in last_use.items() if lu == i and rid in live]:
                    base, sz = live.pop(rid)
                    al.release(base, sz)
            warp_words = max(warp_words, al.high)
            counts.append(len(lst))
            # forget temp addresses so a stale one can never be read by the next list
            pers_ids = {p.id for p in persistent}
            for ins in lst:
                if ins.dst is not None and root(ins.dst).id not in pers_ids:
                    asm.addr.pop(root(ins.dst).id, None)

        if warp_words >= 0x8000 or warp_words + asm.const_words > self.SMEM_WORDS:
            raise ProgramError(f"warp register file too large ({warp_words} words)")
        if asm.const_words >= 0x8000:
            raise ProgramError(f"constant region too large ({asm.const_words} words)")

        pad = (-asm.const_words) % 4
        if pad:
            asm.const_image.append(np.zeros(pad, dtype=np.uint32))
            asm.const_words += pad
        const_img = np.concatenate(asm.const_image).astype(np.uint32)
        hdr = np.zeros(HEADER_WORDS, dtype=np.uint32)
        hdr[0] = isa.MAGIC
        hdr[1] = 1
        hdr[2:5] = counts
        hdr[5] = asm.const_words
        hdr[6] = mparam_off
        hdr[7] = n_mp
        hdr[8] = warp_words
        hdr[9] = rpw
        hdr[10] = L
        hdr[11] = self.S
        hdr[12] = self.A
        hdr[13] = self.N
        hdr[14] = (1 if self.relax is not None else 0) | (2 if kind == "backward" else 0)
        hdr[15] = sg.static_err
        hdr[16] = err_base
        hdr[17] = getattr(self, "_traj_words", 0)
        hdr[18] = getattr(self, "_tape_words", 0) or self.S
        ins_img = np.concatenate(encoded).astype(np.uint16) if encoded else np.zeros(0, np.uint16)
        blob = hdr.tobytes() + const_img.tobytes() + ins_img.tobytes()
        hist: Dict[str, int] = {}
        for ins in step:
            hist[ins.op] = hist.get(ins.op, 0) + 1
        return Program(
            blob=blob, kind=kind, relaxed=self.relax is not None, rpw=rpw, lanes=L,
            warp_words=warp_words, const_words=asm.const_words, n_ins=tuple(counts),
            S=self.S, A=self.A, N=self.N, state_layout=self.slayout, param_layout=self.alayout,
            noise_layout=[(s.kind, s.shape, s.n, s.offset, s.extra) for s in sg.noise],
            mparam_keys=[k for k, _, _ in sg.mparams], mparam_defaults=mp_defaults,
            static_err=sg.static_err, stats={"step_ops": hist},
            tape_words=int(hdr[18]),
            meta={"ins": records, "persistent": [root(p).id for p in persistent],
                  "err_base": err_base, "mparam_off": mparam_off,
                  "const_image": const_img.copy()})

    def _record(self, asm: Assembler, ins: Ins, w: np.ndarray) -> Dict[str, Any]:
        """Decoded form of one instruction (value ids, dtypes, constants) for core/codegen.py."""
        def val(v: V, m: Map = None, k: int = -1):
            r = root(v)
            base, sq = asm.addr[r.id]
            kind = "const" if r.const is not None else ("mparam" if r.op == "mparam" else "val")
            mp = int(w[7 + 3 * k]) if k >= 0 else 0
            return {"id": r.id, "kind": kind, "base": int(base) & 0x7FFF if kind != "val" else int(base),
                    "sq": int(sq), "n": r.n, "dtype": v.dtype, "uniform": r.uniform,
                    "map": mp, "const": None if r.const is None else r.const.copy()}
        rec = {"op": ins.op, "n": int(ins.n), "flags": int(ins.flags),
               "dst": None if ins.dst is None else val(ins.dst),
               "srcs": [val(a, m, k) for k, (a, m) in enumerate(ins.srcs[:3])],
               "aux0": int(w[14]), "aux1": int(w[15]), "buf": int(ins.buf), "off": int(ins.off),
               "seg": int(w[8]) if "seg" in ins.tables else 0}
        if ins.tables:
            rec["kmax"] = int(np.max(np.diff(ins.tables["ptr"]))) if len(ins.tables["ptr"]) > 1 else 0
            rec["tab"] = {k: np.asarray(t) for k, t in ins.tables.items()}    # for the code generator
        return rec

    def _encode(self, asm: Assembler, ins: Ins) -> np.ndarray:
        w = np.zeros(isa.INS_WORDS, dtype=np.uint16)
        w[0] = isa.OPCODE[ins.op]
        w[1] = ins.flags
        w[2] = ins.n
        if ins.n > 0xFFFF:
            raise ProgramError("value too large")
        if ins.dst is not None:
            base, sq = asm.addr[root(ins.dst).id]
            if base >= 0x8000 and not (base & isa.CONST_BIT and base < 0x10000):
                raise ProgramError("warp register file too large")
            if base >= 0x10000 or sq >= 0x10000:
                raise ProgramError("warp register file too large")
            w[3], w[4] = base, sq
        for k, (a, m) in enumerate(ins.srcs[:3]):
            base, sq = asm.addr[root(a).id]
            if base >= 0x10000 or sq >= 0x10000:
                raise ProgramError("register or constant space too large")
            if m is None:
                mp = isa.MAP_IDENT
            elif isinstance(m, str):
                mp = isa.MAP_BCAST
            else:
                mp = asm.table(m)
                if mp >= 0x8000:
                    raise ProgramError("constant region too large")
            w[5 + 3 * k: 8 + 3 * k] = (base, sq, mp)
        if ins.tables:
            tp, ti = asm.table(ins.tables["ptr"]), asm.table(ins.tables["idx"])
            ts = asm.table(ins.tables["seg"]) if "seg" in ins.tables else 0
            if max(tp, ti, ts) >= 0x8000:
                raise ProgramError("constant region too large")
            w[14], w[15] = tp, ti
            if "seg" in ins.tables:
                w[8] = ts
        elif ins.op in ("LDB", "STB", "LDU", "STG", "STERR"):
            w[14] = ins.buf
            w[15] = ins.off & 0xFFFF
            w[11] = (ins.off >> 16) & 0xFFFF
        elif ins.op == "ERRCHK":
            w[14] = ins.aux0
        return w
