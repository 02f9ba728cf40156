"""Runs a few eager (no CUDA graph) planner iterations of a bench workload, for ncu:

    python tools/profile_workload.py --workload wildfire --reps 3

The launches of one iteration of a straight-line plan are, in order: optimiser (+ projection),
exact test forward (rk_spec_kernel), loss, relaxed train forward (rk_spec_kernel), loss,
adjoint (rk_spec_kernel), partial reduction."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from pyrddlgym_jax_b200.core import logic  # noqa: E402
from pyrddlgym_jax_b200.core.planner import (JaxBackpropPlanner, JaxDeepReactivePolicy,  # noqa: E402
                                             JaxStraightLinePlan)

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="quadcopter", choices=sorted(bench.WORKLOADS))
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--horizon", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
model = bench.workload_model(a.workload)
B, T = (a.batch or wl["B"]), (a.horizon or wl["T"])
plan_cls = JaxDeepReactivePolicy if wl.get("plan") == "drp" else JaxStraightLinePlan
pl = JaxBackpropPlanner(model, plan_cls(**wl["plan_kwargs"]), batch_size_train=B, batch_size_test=B,
                        rollout_horizon=T, optimizer="rmsprop",
                        optimizer_kwargs={"learning_rate": wl["lr"]}, pgpe=None,
                        compiler=getattr(logic, wl["compiler"]), compiler_kwargs=wl["compiler_kwargs"],
                        use_cuda_graph=False)
pl.initialize(42)
for _ in range(a.reps):
    pl.redraw_noise()
    pl.iterate()
torch.cuda.synchronize()
print("ok", float(pl.train_loss[0]), float(pl.test_loss_buf[0]))
