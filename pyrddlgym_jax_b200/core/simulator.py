"""Step-by-step simulator of an RDDL model on the exact rollout kernels (the reference's
``JaxRDDLSimulator``, core/simulator.py:43-254: the gym simulation backend).

The reference jit-compiles every CPF, the reward and every constraint separately and walks them on the host
each step (simulator.py:191-254).  Here one environment step is ONE launch of the exact-model program with
the caller's actions (``PlanSpec(kind='external')``, T = 1): the CPFs in level order, the reward, and the
constraints -- state invariants, action preconditions and termination conditions become boolean values of
the same step graph (``constraint_view``), logged next to the observ-fluents.  ``n_envs`` environments step
in the same launch (one rollout each).  Noise comes from the counter-based kernel, keyed by ``seed``.

Differences to note: error bits and constraint violations are read once per step (the reference raises from
the middle of its host loop); ``keep_tensors=False`` grounds names as ``fluent___obj1__obj2`` like pyRDDLGym.
"""
from __future__ import annotations

import copy
import time
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from ..rddl.expr import Expr
from . import engine
from .compiler import JaxRDDLCompiler
from .layout import pack_fluent_major, unpack_fluent_major
from .program import PlanSpec


class RDDLStateInvariantNotSatisfiedError(ValueError):
    pass


class RDDLActionPreconditionNotSatisfiedError(ValueError):
    pass


class RDDLInvalidExpressionError(ValueError):
    pass


def _primed(expr: Expr, model) -> Expr:
    """The expression with every state fluent replaced by its next-state fluent (a constraint on the state
    the step ARRIVES in)."""
    fam, _ = expr.etype
    if fam == "pvar":
        name, params = expr.args
        new = model.next_state.get(name, name)
        params = None if params is None else [(_primed(p, model) if isinstance(p, Expr) else p)
                                              for p in params]
        return Expr(("pvar", new), (new, params))
    if fam == "constant":
        return expr
    if isinstance(expr.args, (list, tuple)):
        args = [(_primed(a, model) if isinstance(a, Expr) else
                 ([(lbl, _primed(e, model)) for lbl, e in a] if isinstance(a, list) and a
                  and isinstance(a[0], tuple) and len(a[0]) == 2 and isinstance(a[0][1], Expr) else a))
                for a in expr.args]
        return Expr(expr.etype, type(expr.args)(args) if isinstance(expr.args, tuple) else args)
    return expr


def constraint_view(model):
    """The model with its constraints as boolean intermediate fluents of the step:
    ``__pre<i>`` (action preconditions, on the current state and the actions), ``__inv<i>`` (state
    invariants of the current state), ``__inv_next<i>`` and ``__term<i>`` (invariants / terminations of the
    state the step arrives in)."""
    m = copy.copy(model)
    for attr in ("interm_fluents", "variable_types", "variable_ranges", "variable_params",
                 "variable_defaults", "init_values", "cpfs"):
        setattr(m, attr, dict(getattr(model, attr)))
    names = []

    def add(name, expr):
        m.interm_fluents[name] = False
        m.variable_types[name] = "interm-fluent"
        m.variable_ranges[name] = "bool"
        m.variable_params[name] = []
        m.variable_defaults[name] = False
        m.init_values[name] = np.zeros((), dtype=bool)
        m.cpfs[name] = ([], expr)
        names.append(name)
    for i, e in enumerate(model.preconditions):
        add(f"__pre{i}", e)
    for i, e in enumerate(model.invariants):
        add(f"__inv{i}", e)
        add(f"__inv_next{i}", _primed(e, model))
    for i, e in enumerate(model.terminations):
        add(f"__term{i}", _primed(e, model))
    m.levels = m._compute_levels()
    m.constraint_fluents = names
    return m


class JaxRDDLSimulator:
    """A simulator for RDDL models on the exact rollout kernels (simulator.py:43)."""

    def __init__(self, rddl, key: Optional[int] = None, raise_error: bool = True, logger=None,
                 keep_tensors: bool = False, objects_as_strings: bool = True,
                 python_functions: Optional[Dict] = None, n_envs: int = 1, device=None,
                 **compiler_args) -> None:
        if python_functions:
            raise NotImplementedError("external Python functions cannot be lowered to the GPU")
        if not torch.cuda.is_available():
            raise engine.RkError("JaxRDDLSimulator needs a CUDA device: the rollout kernels have no CPU "
                                 "fallback")
        self.rddl = rddl
        self.raise_error = raise_error
        self.logger = logger
        self.keep_tensors = keep_tensors
        self.objects_as_strings = objects_as_strings
        self.compiler_args = compiler_args
        self.B = int(n_envs)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.key = round(time.time() * 1000) % (2 ** 31) if key is None else int(np.asarray(key).ravel()[-1])
        self._compile()

    def seed(self, seed: int) -> None:
        self.key = int(seed)
        self._draws.zero_()

    # ---- compilation (simulator.py:87-139) -----------------------------------------------------------
    def _compile(self):
        rddl = self.rddl
        self._view = constraint_view(rddl)
        self.compiled = JaxRDDLCompiler(self._view, **self.compiler_args)
        self.compiled.compile()
        self.init_values = dict(rddl.init_values)
        self.levels = rddl.levels
        logged = tuple(self._view.constraint_fluents) + tuple(rddl.observ_fluents) + \
            tuple(rddl.interm_fluents) + tuple(rddl.derived_fluents)
        spec = PlanSpec(kind="external", bounds=rddl.bounds)
        with torch.cuda.device(self.device):
            self.prog = engine.DeviceProgram(
                self.compiled.builder(spec, False, logged, self.B).forward(write_tape=False, gamma=1.0))
        p, B, dev = self.prog.program, self.B, self.device
        z = lambda n, dt=torch.float32: torch.zeros(n, dtype=dt, device=dev)     # noqa: E731
        self._state = [z(p.S * B), z(p.S * B)]
        self._cur = 0
        self._actions = z(max(p.A, 1) * B)
        self._noise = z(p.N * B) if p.N else None
        self._reward, self._returns = z(B), z(B)
        self._err = z(B, torch.int32)
        self._traj = z(p.TRAJ * B) if p.TRAJ else None
        self._draws = torch.zeros(1, dtype=torch.int64, device=dev)
        self._sites = [(kind, off, n) for kind, _, n, off, _ in p.noise_layout]
        conc = None
        for kind, _, n, off, extra in p.noise_layout:
            if kind not in engine.NOISE_KINDS:
                raise NotImplementedError(f"noise kind {kind}")
            if kind == "gamma":
                conc = np.ones(p.N, np.float32) if conc is None else conc
                conc[off:off + n] = np.maximum(np.broadcast_to(
                    np.asarray(extra, np.float32).reshape(-1), (n,)), 1e-6)
        self._conc = None if conc is None else torch.from_numpy(conc).to(dev)
        self.noop_actions = {a: np.asarray(rddl.init_values[a]) for a in rddl.action_fluents}
        self._pomdp = bool(rddl.observ_fluents)
        self.invariant_names = [f"Invariant {i}" for i in range(len(rddl.invariants))]
        self.precond_names = [f"Precondition {i}" for i in range(len(rddl.preconditions))]
        self.terminal_names = [f"Termination {i}" for i in range(len(rddl.terminations))]
        self.state: Optional[Dict[str, Any]] = None
        self._flags: Dict[str, np.ndarray] = {}
        self._last_reward = np.zeros(B, np.float32)
        self.reset()

    # ---- host <-> kernel layouts ---------------------------------------------------------------------
    def _pack(self, layout, values: Dict[str, Any]) -> torch.Tensor:
        vals = {}
        for name in layout:
            ref = np.asarray(self.rddl.init_values[name])
            v = values.get(name, ref)
            v = np.asarray(self._encode(name, v))
            v = np.broadcast_to(v.reshape((-1,) + ref.shape) if v.size == self.B * ref.size and self.B > 1
                                else v.reshape(ref.shape), (self.B,) + ref.shape).astype(ref.dtype)
            vals[name] = torch.from_numpy(np.ascontiguousarray(v))
        return pack_fluent_major(layout, vals, self.B).to(self.device)

    def _encode(self, name, v):
        prange = self.rddl.variable_ranges[name]
        if prange in self.rddl.enum_types or prange in self.rddl.type_to_objects:
            arr = np.asarray(v)
            if arr.dtype.kind in ("U", "S", "O"):
                return np.vectorize(lambda o: self.rddl.object_to_index[str(o).lstrip("@")])(arr)
        return v

    def _ground(self, name: str, values: np.ndarray) -> Dict[str, Any]:
        """pyRDDLGym's grounded naming: fluent___obj1__obj2 -> scalar."""
        params = self.rddl.variable_params[name]
        if not params:
            return {name: values.reshape(()).item()}
        objs = [self.rddl.type_to_objects[t] for t in params]
        out = {}
        for idx in np.ndindex(*values.shape):
            out[name + "___" + "__".join(objs[d][i] for d, i in enumerate(idx))] = values[idx].item()
        return out

    def _present(self, tensors: Dict[str, torch.Tensor]) -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        for name, t in tensors.items():
            v = t.cpu().numpy()
            v = v[0] if self.B == 1 else v
            prange = self.rddl.variable_ranges[name]
            if self.objects_as_strings and prange in self.rddl.type_to_objects and prange not in (
                    "real", "int", "bool"):
                objs = np.asarray(self.rddl.type_to_objects[prange], dtype=object)
                v = objs[np.asarray(v, dtype=np.int64)]
            if self.keep_tensors or self.B > 1:
                out[name] = v
            else:
                out.update(self._ground(name, np.asarray(v)))
        return out

    def _actions_from(self, actions: Dict[str, Any]) -> Dict[str, Any]:
        """Accepts lifted tensors or pyRDDLGym's grounded names (fluent___obj1__obj2)."""
        out = {a: np.array(v, copy=True) for a, v in self.noop_actions.items()}
        for k, v in actions.items():
            if k in out:
                out[k] = np.asarray(self._encode(k, v))
                continue
            name, _, rest = k.partition("___")
            if name not in out:
                raise KeyError(f"<{k}> is not an action-fluent of the model")
            idx = tuple(self.rddl.object_to_index[o] for o in rest.split("__")) if rest else ()
            out[name][idx] = self._encode(name, v)
        return out

    # ---- one launch ----------------------------------------------------------------------------------
    def _launch(self, actions: Dict[str, Any], advance: bool):
        p, B = self.prog.program, self.B
        if p.A:
            self._actions.copy_(self._pack(self._action_layout(), actions))
        if self._noise is not None:
            engine.draw_noise(1, B, self._sites, self.key, self._draws, self._noise, self._conc)
            engine.noise_advance(self._draws)
        x, nxt = self._state[self._cur], self._state[self._cur ^ 1]
        self.prog.forward(B, 1, x, actions=self._actions, noise=self._noise, reward=self._reward,
                          returns=self._returns, err=self._err, final=nxt, traj=self._traj)
        torch.cuda.synchronize(self.device)
        flags = {}
        logs = {}
        if self._traj is not None:
            shapes = {f: tuple(self._view.variable_shape(f)) for f in p.traj_layout}
            fl = unpack_fluent_major(p.traj_layout, self._traj, B, shapes)
            for name, t in fl.items():
                if name.startswith("__"):
                    flags[name] = t.reshape(B).cpu().numpy().astype(bool)
                else:
                    logs[name] = t
        self._flags, self._logs = flags, logs
        self._last_reward = self._reward.cpu().numpy().copy()
        self.handle_error_code(int(np.bitwise_or.reduce(self._err.cpu().numpy().astype(np.int64))), "step")
        if advance:
            self._cur ^= 1

    def _action_layout(self):
        """Action words of the exact model: (offset, words, dtype code) per action fluent."""
        from .program import action_layout
        code = {"real": "f", "int": "i", "bool": "b"}
        return {name: (off, n, code.get(prange, "i"))
                for name, (off, n, prange) in action_layout(self.rddl)[0].items()}

    def handle_error_code(self, error: int, msg: str) -> None:
        """simulator.py:141-149."""
        if self.raise_error and error:
            errors = JaxRDDLCompiler.get_error_messages(error)
            if errors:
                raise RDDLInvalidExpressionError(
                    f"Internal error in evaluation of {msg}:\n" +
                    "\n".join(f"{i + 1}. {s}" for i, s in enumerate(errors)))

    def _state_dict(self) -> Dict[str, Any]:
        p = self.prog.program
        shapes = {f: tuple(self.rddl.variable_shape(f)) for f in p.state_layout}
        return self._present(unpack_fluent_major(p.state_layout, self._state[self._cur], self.B, shapes))

    # ---- the reference's methods ---------------------------------------------------------------------
    def reset(self) -> Tuple[Dict[str, Any], bool]:
        """Initial state; the constraints of that state are evaluated by a probe launch with no-op actions
        that does not advance the environment (and does not consume the noise stream)."""
        p = self.prog.program
        self._cur = 0
        self._state[0].copy_(self._pack(p.state_layout, {}))
        draws = self._draws.clone()
        self._launch(self.noop_actions, advance=False)
        self._draws.copy_(draws)
        self.state = self._state_dict()
        obs = self.state
        if self._pomdp:
            obs = self._present({o: torch.as_tensor(np.broadcast_to(
                np.asarray(self.rddl.init_values[o]), (self.B,) + np.asarray(self.rddl.init_values[o]).shape)
                .copy()) for o in self.rddl.observ_fluents})
        return obs, False

    def check_state_invariants(self, silent: bool = False) -> bool:
        """simulator.py:151-162, on the current state (evaluated by the last launch)."""
        key = "__inv" if self._inv_of_current else "__inv_next"
        for i, loc in enumerate(self.invariant_names):
            ok = self._flags.get(f"{key}{i}")
            if ok is not None and not bool(ok.all()):
                if not silent:
                    raise RDDLStateInvariantNotSatisfiedError(f"{loc} is not satisfied.")
                return False
        return True

    def check_action_preconditions(self, actions: Dict[str, Any], silent: bool = False) -> bool:
        """simulator.py:164-179: evaluates the preconditions for ``actions`` on the current state (a probe
        launch that does not advance the environment)."""
        draws = self._draws.clone()
        self._launch(self._actions_from(actions), advance=False)
        self._draws.copy_(draws)
        self._inv_of_current = True
        for i, loc in enumerate(self.precond_names):
            ok = self._flags.get(f"__pre{i}")
            if ok is not None and not bool(ok.all()):
                if not silent:
                    raise RDDLActionPreconditionNotSatisfiedError(
                        f"{loc} is not satisfied for actions {actions}.")
                return False
        return True

    def check_terminal_states(self) -> bool:
        """simulator.py:181-189, on the state the last step arrived in."""
        return any(bool(self._flags.get(f"__term{i}", np.zeros(1, bool)).any())
                   for i in range(len(self.terminal_names)))

    def sample_reward(self) -> float:
        """simulator.py:191-196: the reward of the last launch."""
        return float(self._last_reward[0]) if self.B == 1 else self._last_reward

    def step(self, actions: Dict[str, Any]) -> Tuple[Dict[str, Any], Any, bool]:
        """simulator.py:198-254: CPFs in level order, reward, state <- next state, observation, done."""
        self._launch(self._actions_from(actions), advance=True)
        self._inv_of_current = False
        reward = self.sample_reward()
        self.state = self._state_dict()
        obs = self._present({o: self._logs[o] for o in self.rddl.observ_fluents}) if self._pomdp \
            else self.state
        done = self.check_terminal_states()
        return obs, reward, done

    _inv_of_current = True
