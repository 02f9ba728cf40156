"""Lifted RDDL model: the surface of pyRDDLGym that the reference's hot path consumes.

Re-creates, for this build, the attributes listed in SURVEY Appendix B that
compiler.py / planner.py read from ``RDDLLiftedModel`` (variable_ranges, variable_types,
state_fluents, action_fluents, next_state, cpfs, reward, horizon, discount,
max_allowed_actions ...), the initial-value tensors of ``RDDLValueInitializer``
(compiler.py:130-131), the CPF evaluation levels of ``RDDLLevelAnalysis``
(compiler.py:136-137) and the box bounds of ``RDDLConstraints`` (compiler.py:152-159,
planner.py:450).
"""
from __future__ import annotations

import math
import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .expr import Expr, walk
from .parser import parse_rddl

NUMPY_TYPES = {"real": np.float32, "int": np.int32, "bool": np.bool_}


class RDDLModelError(Exception):
    pass


class RDDLLiftedModel:
    """A lifted (object-parameterised) RDDL domain + instance."""

    def __init__(self, domain: Dict[str, Any], nonfluents: Optional[Dict[str, Any]],
                 instance: Dict[str, Any]):
        self.domain_name = domain["name"]
        self.instance_name = instance["name"]
        self.policy = None

        # ---- types and objects
        self.enum_types = {t for t, v in domain["types"].items() if v is not None}
        self.type_to_objects: Dict[str, List[str]] = {}
        for t, vals in domain["types"].items():
            if vals is not None:
                self.type_to_objects[t] = [v.lstrip("@") for v in vals]
        for blk in (nonfluents or {}, instance):
            for t, objs in blk.get("objects", {}).items():
                cur = self.type_to_objects.setdefault(t, [])
                cur.extend(o for o in objs if o not in cur)
        for t in domain["types"]:
            if t not in self.type_to_objects:
                raise RDDLModelError(f"type <{t}> has no objects")
        self.object_to_index: Dict[str, int] = {}
        self.object_to_type: Dict[str, str] = {}
        for t, objs in self.type_to_objects.items():
            for i, o in enumerate(objs):
                self.object_to_index[o] = i
                self.object_to_type[o] = t

        # ---- pvariables
        self.variable_types: Dict[str, str] = {}
        self.variable_ranges: Dict[str, str] = {}
        self.variable_params: Dict[str, List[str]] = {}
        self.variable_defaults: Dict[str, Any] = {}
        self.state_fluents: Dict[str, Any] = {}
        self.action_fluents: Dict[str, Any] = {}
        self.interm_fluents: Dict[str, Any] = {}
        self.derived_fluents: Dict[str, Any] = {}
        self.observ_fluents: Dict[str, Any] = {}
        self.non_fluents: Dict[str, Any] = {}
        self.next_state: Dict[str, str] = {}
        self.prev_state: Dict[str, str] = {}
        for name, pv in domain["pvariables"].items():
            kind, prange = pv["kind"], pv["range"]
            if kind == "interm-fluent" or kind == "intermediate-fluent":
                kind = "interm-fluent"
            default = pv["default"]
            if default is None:
                default = {"real": 0.0, "int": 0, "bool": False}.get(prange)
                if default is None:  # enum/object-valued
                    default = self.type_to_objects[prange][0]
            if isinstance(default, str):
                default = default.lstrip("@")
            self.variable_types[name] = kind
            self.variable_ranges[name] = prange
            self.variable_params[name] = list(pv["params"])
            self.variable_defaults[name] = default
            target = {"state-fluent": self.state_fluents, "action-fluent": self.action_fluents,
                      "interm-fluent": self.interm_fluents, "derived-fluent": self.derived_fluents,
                      "observ-fluent": self.observ_fluents, "non-fluent": self.non_fluents}.get(kind)
            if target is None:
                raise RDDLModelError(f"unknown pvariable kind <{kind}> for <{name}>")
            target[name] = default
            if kind == "state-fluent":
                primed = name + "'"
                self.next_state[name] = primed
                self.prev_state[primed] = name
                self.variable_types[primed] = "next-state-fluent"
                self.variable_ranges[primed] = prange
                self.variable_params[primed] = list(pv["params"])
                self.variable_defaults[primed] = default

        # ---- cpfs / reward / constraints
        self.cpfs: Dict[str, Tuple[List[Tuple[str, str]], Expr]] = {}
        for name, params, expr in domain["cpfs"]:
            if name not in self.variable_types:
                raise RDDLModelError(f"cpf <{name}> is not a declared pvariable")
            ptypes = self.variable_params[name]
            if len(params) != len(ptypes):
                raise RDDLModelError(f"cpf <{name}> has wrong arity")
            self.cpfs[name] = (list(zip(params, ptypes)), expr)
        for s, sp in self.next_state.items():
            if sp not in self.cpfs:
                raise RDDLModelError(f"missing cpf for <{sp}>")
        self.reward: Expr = domain["reward"]
        self.invariants: List[Expr] = domain["invariants"]
        self.preconditions: List[Expr] = domain["preconditions"]
        self.terminations: List[Expr] = domain["terminations"]

        # ---- instance
        self.horizon = int(instance["horizon"])
        self.discount = float(instance["discount"])
        mna = instance["max-nondef-actions"]
        n_actions = sum(self.variable_size(a) for a in self.action_fluents)
        self.max_allowed_actions = n_actions if (isinstance(mna, float) and math.isinf(mna)) \
            else int(mna)

        # ---- initial values (RDDLValueInitializer, compiler.py:130-131)
        self.init_values: Dict[str, np.ndarray] = {}
        for name in self.variable_types:
            if name.endswith("'"):
                continue
            self.init_values[name] = self._default_tensor(name)
        for blk, allowed in ((nonfluents, self.non_fluents), (instance, None)):
            if blk is None:
                continue
            key = "values" if allowed is not None else "init-state"
            for name, args, val in blk.get(key, []):
                if name not in self.init_values:
                    raise RDDLModelError(f"assignment to unknown pvariable <{name}>")
                idx = tuple(self.object_to_index[a.lstrip("@")] for a in args)
                self.init_values[name][idx] = self._encode(name, val)
        for s, sp in self.next_state.items():
            self.init_values[sp] = self.init_values[s].copy()
        for name in self.non_fluents:
            self.non_fluents[name] = self.init_values[name]

        self.levels = self._compute_levels()
        self.bounds = self._extract_action_bounds()
        self.state_bounds = self._extract_state_bounds()

    # ---------------------------------------------------------------------------------
    def variable_shape(self, name: str) -> Tuple[int, ...]:
        return tuple(len(self.type_to_objects[t]) for t in self.variable_params[name])

    def variable_size(self, name: str) -> int:
        return int(np.prod(self.variable_shape(name), dtype=np.int64))

    def object_counts(self, types) -> Tuple[int, ...]:
        return tuple(len(self.type_to_objects[t]) for t in types)

    def _encode(self, name: str, val: Any):
        prange = self.variable_ranges[name]
        if prange in ("real", "int", "bool"):
            return val
        return self.object_to_index[str(val).lstrip("@")]  # enum/object-valued -> index

    def _default_tensor(self, name: str) -> np.ndarray:
        prange = self.variable_ranges[name]
        dtype = NUMPY_TYPES.get(prange, np.int32)
        return np.full(self.variable_shape(name),
                       self._encode(name, self.variable_defaults[name]), dtype=dtype)

    # -- level analysis (RDDLLevelAnalysis.compute_levels) -------------------------------
    def _compute_levels(self) -> Dict[int, List[str]]:
        deps: Dict[str, set] = {}
        for name, (_, expr) in self.cpfs.items():
            d = set()
            for node in walk(expr):
                if node.etype[0] == "pvar":
                    ref = node.args[0]
                    if ref in self.cpfs and ref != name:
                        d.add(ref)
            deps[name] = d
        level: Dict[str, int] = {}

        def visit(n, stack=()):
            if n in level:
                return level[n]
            if n in stack:
                raise RDDLModelError(f"cyclic cpf dependency through <{n}>")
            lv = 0
            for d in deps[n]:
                lv = max(lv, visit(d, stack + (n,)) + 1)
            level[n] = lv
            return lv

        for n in self.cpfs:
            visit(n)
        out: Dict[int, List[str]] = {}
        # observ-fluents are evaluated after everything else
        max_lv = max(level.values(), default=0)
        for n in self.cpfs:
            lv = level[n]
            if self.variable_types[n] == "observ-fluent":
                lv = max(lv, max_lv + 1)
            out.setdefault(lv, []).append(n)
        return dict(sorted(out.items()))

    # -- box bounds from action-preconditions (RDDLConstraints) --------------------------
    def _extract_action_bounds(self) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
        bounds = {}
        for a in self.action_fluents:
            shape = self.variable_shape(a)
            bounds[a] = (np.full(shape, -np.inf, dtype=np.float32),
                         np.full(shape, np.inf, dtype=np.float32))
        from .trace import eval_static  # late import (trace depends on this module)
        for pre in self.preconditions:
            self._bounds_from(pre, [], bounds, eval_static)
        return bounds

    def _extract_state_bounds(self) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
        """Box bounds of the state fluents from state-invariants and action-preconditions (the
        RDDLConstraints bounds that StaticNormalizer reads, planner.py:300-312)."""
        bounds = {}
        for s in self.state_fluents:
            shape = self.variable_shape(s)
            bounds[s] = (np.full(shape, -np.inf, dtype=np.float32),
                         np.full(shape, np.inf, dtype=np.float32))
        from .trace import eval_static
        for expr in list(self.invariants) + list(self.preconditions):
            self._bounds_from(expr, [], bounds, eval_static, self.state_fluents)
        return bounds

    def _bounds_from(self, expr, scope, bounds, eval_static, names=None):
        names = self.action_fluents if names is None else names
        fam, op = expr.etype
        if fam == "aggregation" and op == "forall":
            *tv, body = expr.args
            self._bounds_from(body, scope + list(tv), bounds, eval_static, names)
            return
        if fam == "boolean" and op in ("^", "&"):
            for a in expr.args:
                self._bounds_from(a, scope, bounds, eval_static, names)
            return
        if fam != "relational" or op not in ("<=", ">=", "<", ">"):
            return
        lhs, rhs = expr.args
        for act_side, other, sense in ((lhs, rhs, op), (rhs, lhs, {"<=": ">=", ">=": "<=",
                                                                    "<": ">", ">": "<"}[op])):
            if act_side.etype[0] != "pvar" or act_side.args[0] not in names:
                continue
            name, params = act_side.args
            params = params or []
            svars = [v for v, _ in scope]
            if any(not isinstance(p, str) or p not in svars for p in params) \
                    or len(set(params)) != len(params):
                continue
            try:
                val = eval_static(self, other, scope)   # tensor over scope, or raises if fluent
            except Exception:
                continue
            val = np.broadcast_to(np.asarray(val, dtype=np.float32),
                                  self.object_counts(t for _, t in scope))
            # reduce scope axes not used by the action, then order as the action's params
            used = [svars.index(p) for p in params]
            other_axes = tuple(i for i in range(len(scope)) if i not in used)
            lo, hi = bounds[name]
            if sense in ("<=", "<"):
                v = val.min(axis=other_axes) if other_axes else val
                v = self._reorder(v, used)
                np.minimum(hi, v, out=hi)
            else:
                v = val.max(axis=other_axes) if other_axes else val
                v = self._reorder(v, used)
                np.maximum(lo, v, out=lo)

    @staticmethod
    def _reorder(v, used):
        # v has axes in increasing scope-index order of `used`; reorder to the order in `used`
        if len(used) <= 1:
            return v
        order = np.argsort(used)          # position in v of each sorted axis
        inv = np.argsort(order)
        return np.transpose(v, inv)


def load_model(domain: str, instance: Optional[str] = None) -> RDDLLiftedModel:
    """Build a lifted model from RDDL text or file paths.

    ``domain`` / ``instance`` may each be a path or RDDL source text; a single text may hold
    all three blocks.
    """
    def read(x):
        if x is None:
            return ""
        if "\n" not in x and os.path.exists(x):
            with open(x) as f:
                return f.read()
        return x
    blocks = parse_rddl(read(domain) + "\n" + read(instance))
    if blocks["domain"] is None:
        raise RDDLModelError("no domain block found")
    if not blocks["instance"]:
        raise RDDLModelError("no instance block found")
    inst = blocks["instance"][0]
    nf = None
    for cand in blocks["non-fluents"]:
        if inst["non-fluents"] is None or cand["name"] == inst["non-fluents"]:
            nf = cand
    return RDDLLiftedModel(blocks["domain"], nf, inst)


SEEN = "@seen"


def observation_view(model: "RDDLLiftedModel") -> "RDDLLiftedModel":
    """POMDP view for closed-loop policies: every observ-fluent ``o`` gets a carried copy ``o@seen`` -- a
    state fluent whose next value is the observation the step computes -- so that the value a policy may
    look at in step t + 1 (the reference keeps it in ``sim_state.fls``, compiler.py:524-545; the policy
    reads the observ-fluents only, planner.py:1300-1313) travels through the rollout kernels' state
    vector and its cotangent through their adjoint.  Initial copies hold the observ-fluents' defaults.
    Models without observ-fluents are returned unchanged."""
    import copy
    if not model.observ_fluents:
        return model
    m = copy.copy(model)
    for attr in ("state_fluents", "next_state", "prev_state", "variable_types", "variable_ranges",
                 "variable_params", "variable_defaults", "init_values", "state_bounds"):
        setattr(m, attr, dict(getattr(model, attr)))
    for o, default in model.observ_fluents.items():
        s = o + SEEN
        m.state_fluents[s] = default
        m.next_state[s] = o
        m.prev_state[o] = s
        m.variable_types[s] = "state-fluent"
        for table in (m.variable_ranges, m.variable_params, m.variable_defaults):
            table[s] = table[o]
        m.init_values[s] = np.array(model.init_values[o], copy=True)
        shape = model.variable_shape(o)
        m.state_bounds[s] = (np.full(shape, -np.inf, dtype=np.float32),
                             np.full(shape, np.inf, dtype=np.float32))
    m.observed_inputs = [o + SEEN for o in model.observ_fluents]
    return m
