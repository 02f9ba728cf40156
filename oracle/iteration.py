"""One planner iteration of the CPU oracle -- TEST INFRASTRUCTURE, never on the product path.

Restates the reference's ``optimize_generator`` body (planner.py:3245-3260) with ``pgpe=None``:
``update`` = relaxed rollouts + loss + reverse-mode gradient (planner.py:2873-2882) + optax rmsprop
step + box / concurrency projection (planner.py:2885-2899), then ``test_loss`` = exact-model rollouts
of the new plan (planner.py:2695-2698).  torch-CPU fp32 + autograd; explicit noise tensors.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` (``cpu_baseline``,
``--impl reference``) import this module.  Parity of the oracle itself against JAX is UNPINNED
(DESIGN.md section 2).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from . import planner as OP
from .model_eval import OracleModel


class OracleIteration:
    """iteration = it() -> (train loss of the gradient applied, test loss of the new plan).

    ``layouts`` = (train noise layout, test noise layout) of the lowered programs (the draw sites in
    program order; ``core/layout.py``).  ``init`` overrides the initial plan (SLP: {action: [T, n]},
    DRP: flax-style parameter tree).  Every call draws its noise from ``generator`` unless ``noise``
    = {'train': [T, N*B], 'test': [T, N*B], 'policy': [T, A*B]} is given; what was used is kept in
    ``last_noise`` so that a caller can replay the same draws elsewhere."""

    def __init__(self, model, relax, layouts, B: int, T: int, lr: float, plan: str = "slp",
                 plan_sigmoid_weight: float = 1.0, topology=(256, 256), ent_coeff: float = 0.1,
                 policy_layout: Optional[Dict[str, Any]] = None, seed: int = 42, init=None):
        self.m, self.B, self.T, self.lr = model, B, T, lr
        self.om_train, self.om_test = OracleModel(model, relax), OracleModel(model, None)
        self.lay_train, self.lay_test = layouts
        self.plan, self.pw, self.topology = plan, float(plan_sigmoid_weight), list(topology)
        self.ent_coeff = float(ent_coeff)
        self.policy_layout = policy_layout
        self.g = torch.Generator().manual_seed(seed)
        self.bounds = model.bounds
        acts = list(model.action_fluents)
        if plan == "drp":
            self.tree = init if init is not None else OP.drp_init(model, self.topology, self.g)
            self.nu = {k: {kk: torch.zeros_like(vv) for kk, vv in v.items()}
                       for k, v in self.tree.items()}
        else:
            self.params = init if init is not None else {
                a: torch.randn(T, model.variable_size(a), generator=self.g) * 0.01 for a in acts}
            self.nu = {a: torch.zeros_like(p) for a, p in self.params.items()}
        n_bool = sum(model.variable_size(a) for a in acts if model.variable_ranges[a] == "bool")
        self.project = model.max_allowed_actions < n_bool
        self.last_noise: Dict[str, Optional[torch.Tensor]] = {}

    # ---- noise ------------------------------------------------------------------------------
    def _noise(self, tag: str, layout, given) -> Optional[List]:
        from pyrddlgym_jax_b200.core.layout import draw_noise, noise_as_list
        if not layout:
            self.last_noise[tag] = None
            return None
        z = given.get(tag) if given else None
        if z is None:
            z = draw_noise(layout, self.T, self.B, self.g)
        self.last_noise[tag] = z
        return [noise_as_list(layout, z, t, self.B) for t in range(self.T)]

    def _policy_noise(self, given):
        m, B, T = self.m, self.B, self.T
        A = sum(m.variable_size(a) for a in m.action_fluents)
        z = given.get("policy") if given else None
        if z is None:
            z = torch.randn(T, A * B, generator=self.g)
        self.last_noise["policy"] = z
        out = []
        if self.policy_layout is not None:
            for t in range(T):
                out.append({a: z[t, o * B:(o + n) * B].reshape(B, n)
                            for a, (o, n, _) in self.policy_layout.items()})
            return out
        for t in range(T):
            row, off = {}, 0
            for a in m.action_fluents:
                n = m.variable_size(a)
                row[a] = z[t, off * B:(off + n) * B].reshape(B, n)
                off += n
            out.append(row)
        return out

    # ---- one iteration ----------------------------------------------------------------------
    def __call__(self, noise: Optional[Dict[str, torch.Tensor]] = None):
        return self._drp(noise) if self.plan == "drp" else self._slp(noise)

    def _slp(self, noise):
        m, B, T = self.m, self.B, self.T
        P = {k: v.clone().requires_grad_(True) for k, v in self.params.items()}
        out = OP.rollout(self.om_train, lambda t, fls: OP.slp_train_actions(
            m, P, t, sigmoid_weight=self.pw), B, T, self._noise("train", self.lay_train, noise))
        loss = OP.mean_loss(out["reward"], m.discount)
        loss.backward()
        self.last_grad = {k: (P[k].grad if P[k].grad is not None else torch.zeros_like(P[k]))
                          for k in P}
        with torch.no_grad():
            new = {}
            for k in self.params:
                new[k], self.nu[k] = OP.rmsprop_step(self.params[k], self.last_grad[k], self.nu[k],
                                                     self.lr)
            new = OP.project_box(m, new, self.bounds, True, 1e-6, self.pw)
            if self.project:
                new, _ = OP.project_max_constraint(m, new, "sogbofa", int(m.max_allowed_actions),
                                                   True, 1e-6, self.pw)
            self.params.update(new)
            test = OP.rollout(self.om_test, lambda t, fls: OP.slp_test_actions(
                m, {k: v.reshape(T, *m.variable_shape(k)) for k, v in self.params.items()},
                t, self.bounds), B, T, self._noise("test", self.lay_test, noise))
            tl = OP.mean_loss(test["reward"], m.discount)
        return float(loss.detach()), float(tl)

    def _drp(self, noise):
        m, B, T, L = self.m, self.B, self.T, len(self.topology)
        P = {k: {kk: vv.clone().requires_grad_(True) for kk, vv in v.items()}
             for k, v in self.tree.items()}
        pn = self._policy_noise(noise)
        ents = []

        def pol(t, fls):
            a, e = OP.drp_actions(m, P, {k: fls[k] for k in m.state_fluents}, pn[t], 1.0,
                                  self.bounds, L)
            ents.append(e)
            return a
        out = OP.rollout(self.om_train, pol, B, T, self._noise("train", self.lay_train, noise))
        loss = OP.mean_loss(out["reward"], m.discount) \
            - self.ent_coeff * torch.stack(ents, 1).sum(1).mean()
        loss.backward()
        self.last_grad = {k: {kk: (vv.grad if vv.grad is not None else torch.zeros_like(vv))
                              for kk, vv in v.items()} for k, v in P.items()}
        with torch.no_grad():
            for k in self.tree:
                for kk in self.tree[k]:
                    self.tree[k][kk], self.nu[k][kk] = OP.rmsprop_step(
                        self.tree[k][kk], self.last_grad[k][kk], self.nu[k][kk], self.lr)
            test = OP.rollout(self.om_test, lambda t, fls: OP.drp_test_actions(
                m, self.tree, {k: fls[k] for k in m.state_fluents}, self.bounds, L),
                B, T, self._noise("test", self.lay_test, noise))
            tl = OP.mean_loss(test["reward"], m.discount)
        return float(loss.detach()), float(tl)
