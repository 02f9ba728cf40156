"""Generates tests/golden/drp/*.npz: frozen inputs and ORACLE outputs of the deep-reactive-policy
restatement (oracle/planner.py: drp_actions / drp_test_actions) with the optional input layers
and heads switched on.

    python -m tests.golden.make_golden_drp

Like make_golden.py these vectors come from the CPU oracle (the reference cannot run here,
DESIGN.md §2): they pin it against silent drift.  Parameters, states and noise are stored in full.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import planner as OP                                    # noqa: E402
from pyrddlgym_jax_b200.core import logic                            # noqa: E402
from pyrddlgym_jax_b200.core.compiler import JaxRDDLCompiler         # noqa: E402
from pyrddlgym_jax_b200.core.drp import (FixedTimeEmbedding,         # noqa: E402
                                         JaxDeepReactivePolicy, SinusoidalTimeEmbedding)
from pyrddlgym_jax_b200.core.planner import StaticNormalizer         # noqa: E402
from tests.helpers import fixture_model                              # noqa: E402

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "drp")
POWERGEN_BOUNDS = {"prevProd": (np.zeros(5, np.float32), np.array([5., 6., 7., 8., 10.], np.float32)),
                   "temperature": (-30.0, 40.0)}

# name -> (fixture, max-nondef-actions override, plan kwargs, oracle kwargs, preprocessor bounds)
CASES = {
    "powergen_norm_film_elu_pre": (
        "powergen", None,
        dict(topology=[12, 8], normalize=True, activation="elu", time_dependent=True,
             time_embedding=SinusoidalTimeEmbedding, time_embedding_kwargs={"dim": 6}),
        dict(normalize=True, activation="elu", time_embedding="sin", time_dim=6), POWERGEN_BOUNDS),
    "wildfire_topk3": (
        "wildfire", 3, dict(topology=[16, 8], softmax_weight=2.0), dict(softmax_weight=2.0), None),
    "quadcopter_perlayer_relu_fixed": (
        "quadcopter", None,
        dict(topology=[16], normalize=True, normalize_per_layer=True,
             normalizer_kwargs={"use_bias": True, "use_scale": True, "epsilon": 0.05},
             activation="relu", wrap_non_bool=True, time_dependent=True,
             time_embedding=FixedTimeEmbedding, time_embedding_kwargs={"max_t": 7, "dim": 4}),
        dict(normalize=True, normalize_per_layer=True, norm_eps=0.05, activation="relu",
             wrap_non_bool=True, time_embedding="fixed", time_dim=4), None),
}


def build(name):
    fixture, allowed, pkw, okw, pre_bounds = CASES[name]
    m = fixture_model(fixture)
    if allowed is not None:
        m.max_allowed_actions = allowed
    plan = JaxDeepReactivePolicy(**pkw)
    pre = StaticNormalizer(fluent_bounds=pre_bounds) if pre_bounds is not None else None
    plan.compile(logic.DefaultJaxRDDLCompilerWithGrad(m), JaxRDDLCompiler(m), {}, 6, preprocessor=pre)
    okw = dict(okw)
    if pre is not None:
        okw["preprocess"] = pre.initialize()
    return m, plan, okw


def evaluate(m, plan, okw, flat, states, noise, step):
    """Oracle outputs for one batch of states: train actions + entropy, test actions, and the
    gradient of sum(train actions * cot) - 0.1 * sum(entropy) w.r.t. the flat parameters."""
    tree = plan.params_to_pytree(flat)["params"]

    def with_grad(t):
        return {k: with_grad(v) for k, v in t.items()} if isinstance(t, dict) \
            else t.clone().requires_grad_(True)

    def grads(t):
        return {k: grads(v) for k, v in t.items()} if isinstance(t, dict) \
            else (t.grad if t.grad is not None else torch.zeros_like(t))
    P = with_grad(tree)
    L = len(plan._topology)
    acts, ent = OP.drp_actions(m, P, states, noise, 1.0, plan.bounds, L, True, step=step, **okw)
    loss = -0.1 * ent.sum()
    for i, v in enumerate(m.action_fluents):
        cot = torch.cos(torch.arange(acts[v].numel(), dtype=torch.float32) * (0.37 + i)).reshape(acts[v].shape)
        loss = loss + (acts[v] * cot).sum()
    loss.backward()
    test = OP.drp_test_actions(m, tree, states, plan.bounds, L, step=step,
                               **{k: v for k, v in okw.items() if k != "sigma_entropy_grad"})
    return acts, ent, test, plan.pytree_to_params(grads(P))


def main(only=()):
    os.makedirs(HERE, exist_ok=True)
    for name in CASES:
        if only and name not in only:
            continue
        m, plan, okw = build(name)
        g = torch.Generator().manual_seed(len(name))
        B, step = 5, 3
        flat = plan.initial_params(g)
        flat = flat + 0.05 * torch.randn(flat.shape, generator=g)       # biases, norms, tables off zero
        states, noise = {}, {}
        for v in m.state_fluents:
            x = torch.randn(B, m.variable_size(v), generator=g) * 2 + 1
            states[v] = (x > 1) if m.variable_ranges[v] == "bool" else x
        for v in m.action_fluents:
            u = torch.rand(B, m.variable_size(v), generator=g).clamp(1e-6, 1 - 1e-6)
            noise[v] = -torch.log(-torch.log(u)) if plan.use_topk else torch.randn(
                B, m.variable_size(v), generator=g)
        acts, ent, test, grad = evaluate(m, plan, okw, flat, states, noise, step)
        data = {"B": B, "step": step, "flat": flat.numpy(), "entropy": ent.detach().numpy(),
                "grad": grad.numpy()}
        for v in m.state_fluents:
            data["state_" + v] = states[v].numpy()
        for v in m.action_fluents:
            data["noise_" + v] = noise[v].numpy()
            data["train_" + v] = acts[v].detach().numpy()
            data["test_" + v] = test[v].numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, flat.numel(), {k: np.asarray(v).shape for k, v in data.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main(tuple(sys.argv[1:]))
