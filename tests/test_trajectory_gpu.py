"""Trajectory parity: five consecutive planner iterations of each BASELINE.json workload (compiler
class, relaxation weights, plan and optimiser exactly as ``bench.WORKLOADS``; reduced batch) on the
GPU through the public planner API, against the CPU oracle's iteration (oracle/iteration.py) on the
same initial plan and the same supplied noise.  After EVERY iteration: train loss, test loss and the
whole plan are compared.

Tolerances.  Losses: rtol 1e-5 on the first iteration (states / returns tolerance of north_star), then
2e-4 (they depend on a plan that went through optimiser steps).  Plan: rmsprop divides the gradient by
its own running magnitude, so a RELATIVE gradient error e of an entry becomes an ABSOLUTE parameter
error of about 3 * lr * e per step whatever the size of the entry; with the gradient tolerance 1e-4 that
is 3e-4 * lr per step, and the test allows PLAN_TOL * lr * iteration."""
import json
import os

import numpy as np
import pytest
import torch

import bench
from oracle.iteration import OracleIteration
from pyrddlgym_jax_b200.core import logic
from pyrddlgym_jax_b200.core.program import PlanSpec, param_layout

pytestmark = pytest.mark.gpu
PLAN_TOL = 1e-3
N_ITER = 5
# reduced batches (the oracle is a CPU program); horizons are the workloads' own
CASES = [("reservoir", 32), ("quadcopter", 64), ("wildfire", 48), ("marsrover", 64),
         ("powergen_drp", 160)]


def _planner(name, B, use_graph):
    from pyrddlgym_jax_b200.core.planner import (JaxBackpropPlanner, JaxDeepReactivePolicy,
                                                 JaxStraightLinePlan)
    wl = bench.WORKLOADS[name]
    model = bench.workload_model(name)
    plan_cls = JaxDeepReactivePolicy if wl.get("plan") == "drp" else JaxStraightLinePlan
    pl = JaxBackpropPlanner(
        model, plan_cls(**wl["plan_kwargs"]), batch_size_train=B, batch_size_test=B,
        rollout_horizon=wl["T"], optimizer="rmsprop", optimizer_kwargs={"learning_rate": wl["lr"]},
        pgpe=None, compiler=getattr(logic, wl["compiler"]), compiler_kwargs=wl["compiler_kwargs"],
        use_cuda_graph=use_graph)
    pl.initialize(42)
    return model, wl, pl


def _record(name, rows):
    out = os.environ.get("RK_TRAJECTORY_LOG")
    if out:
        with open(out, "a") as f:
            f.write(json.dumps({"workload": name, "rows": rows}) + "\n")


@pytest.mark.parametrize("name,B", CASES)
@pytest.mark.parametrize("use_graph", [False, True])
def test_five_iterations_match_oracle(cuda_device, name, B, use_graph):
    model, wl, pl = _planner(name, B, use_graph)
    T = wl["T"]
    fw, te = pl.train_fwd.program, pl.test_fwd.program
    noisy = bool(fw.N or te.N or pl.is_drp)
    if use_graph and noisy:
        pytest.skip("supplied-noise parity runs feed the noise between eager iterations")
    relax = pl.compiled.relax_config()
    flat0 = pl.params[0].cpu().clone()
    if pl.is_drp:
        init = {k: {kk: vv.clone() for kk, vv in v.items()}
                for k, v in pl.plan.params_to_pytree(flat0)["params"].items()}
        layout = None
    else:
        layout, A = param_layout(model, PlanSpec())
        init = {a: flat0.view(T, A)[:, off:off + n].clone() for a, (off, n, _) in layout.items()}
    it = OracleIteration(model, relax, (fw.noise_layout, te.noise_layout), B, T, wl["lr"],
                         plan="drp" if pl.is_drp else "slp",
                         plan_sigmoid_weight=float(wl["plan_kwargs"].get("sigmoid_weight", 1.0)),
                         topology=wl["plan_kwargs"].get("topology", ()),
                         ent_coeff=float(pl.start_entropy_coeff),
                         policy_layout=pl.plan.layout if pl.is_drp else None, init=init)
    rows = []
    for i in range(1, N_ITER + 1):
        want_train, want_test = it()
        z = it.last_noise
        if z.get("train") is not None:
            pl.train_noise.copy_(z["train"].reshape(-1))
        if z.get("test") is not None:
            pl.test_noise.copy_(z["test"].reshape(-1))
        if pl.is_drp:
            pl.drp.pol_noise.copy_(z["policy"])
        pl.iterate()
        train, test, _ = pl.read_stats()
        got = pl.params[0].cpu()
        if pl.is_drp:
            want = pl.plan.pytree_to_params(it.tree)
        else:
            want = torch.zeros_like(got).view(T, -1)
            for a, (off, n, _) in layout.items():
                want[:, off:off + n] = it.params[a].reshape(T, n)
            want = want.reshape(-1)
        dplan = float((got - want).abs().max())
        rows.append({"it": i, "train": [float(train[0]), want_train],
                     "test": [float(test[0]), want_test], "plan_maxdiff": dplan,
                     "plan_tol": PLAN_TOL * wl["lr"] * i})
        ltol = 1e-5 if i == 1 else 2e-4
        assert abs(float(train[0]) - want_train) <= ltol * max(1.0, abs(want_train)), rows[-1]
        assert abs(float(test[0]) - want_test) <= 2e-4 * max(1.0, abs(want_test)), rows[-1]
        assert dplan <= PLAN_TOL * wl["lr"] * i, rows[-1]
    _record(name + ("/graph" if use_graph else "/eager"), rows)
