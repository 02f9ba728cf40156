"""Shared helpers for the parity tests (oracle vs emulator vs CUDA kernels)."""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

from pyrddlgym_jax_b200.rddl.model import load_model
from pyrddlgym_jax_b200.core import logic
from pyrddlgym_jax_b200.core.compiler import JaxRDDLCompiler
from pyrddlgym_jax_b200.core.program import PlanSpec, action_layout
from pyrddlgym_jax_b200.core.layout import pack_fluent_major, draw_noise, noise_as_list

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOMAINS = os.path.join(ROOT, "pyrddlgym_jax_b200", "domains")

FIXTURES = {
    "quadcopter": ("Quadcopter", "instance1"),
    "reservoir": ("Reservoir_Continuous", "instance1"),
    "wildfire": ("Wildfire_MDP_ippc2014", "instance5"),
    "marsrover": ("MarsRover_ippc2023", "instance5"),
    "powergen": ("PowerGen_Continuous", "instance1"),
    "discrete": ("Discrete_Toy", "instance1"),
    "distributions": ("Distributions_Toy", "instance1"),
    "wildfire_free": ("Wildfire_MDP_ippc2014", "instance5free"),    # no concurrency constraint
    "gammafluent": ("Gamma_Fluent_Toy", "instance1"),
    # every operator / sampler of SURVEY Appendix G that the domains above do not use
    "ops": ("Ops_Toy", "instance1"),
    "samplers": ("Samplers_Toy", "instance1"),
    # matrix algebra and random vectors
    "matrix": ("Matrix_Toy", "instance1"),
    # partially observed: noisy sensors and a boolean alarm over hidden tank levels
    "pomdp": ("Pomdp_Toy", "instance1"),
}


def fixture_model(key: str):
    d, i = FIXTURES[key]
    return load_model(os.path.join(DOMAINS, d, "domain.rddl"), os.path.join(DOMAINS, d, i + ".rddl"))


def init_params(model, T: int, seed: int = 42, scale: float = 0.01) -> Dict[str, torch.Tensor]:
    """normal(stddev) initial plan (planner.py:636), one tensor [T, n] per action fluent."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in model.action_fluents:
        n = model.variable_size(name)
        out[name] = torch.randn(T, n, generator=g) * scale
    return out


def pack_params(model, params: Dict[str, torch.Tensor]) -> torch.Tensor:
    """-> [T, A] rows; stochastic plans carry '<fluent>/log_sigma' entries after the actions."""
    from pyrddlgym_jax_b200.core.program import param_layout
    stochastic = any(k.endswith("/log_sigma") for k in params)
    layout, A = param_layout(model, PlanSpec(stochastic=stochastic))
    T = next(iter(params.values())).shape[0]
    flat = torch.zeros(T, A)
    for name, (off, n, _) in layout.items():
        flat[:, off:off + n] = params[name].reshape(T, n)
    return flat.contiguous()


def initial_state(model, layout, B: int) -> torch.Tensor:
    vals = {}
    for name in layout:
        v = torch.from_numpy(np.ascontiguousarray(model.init_values[name]))
        vals[name] = v.unsqueeze(0).repeat(B, *([1] * v.dim()))
    return pack_fluent_major(layout, vals, B)
