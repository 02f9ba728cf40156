import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
from tests.helpers import fixture_model
from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner, JaxStraightLinePlan
m = fixture_model("reservoir")
res = {}
for mode in ("branches", "sequential", "eager", "p1"):
    os.environ["RK_PARALLEL_BRANCHES"] = "1" if mode == "branches" else "0"
    P = 1 if mode == "p1" else 3
    pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=16, rollout_horizon=6,
                            optimizer="rmsprop", optimizer_kwargs={"learning_rate": 0.05}, pgpe=None,
                            parallel_updates=P, use_cuda_graph=(mode != "eager"))
    pl.initialize(5)
    if mode == "p1":
        pl.params[0].copy_(res["init"][1])
    else:
        res["init"] = pl.params.cpu().clone()
    hist = []
    for _ in range(4):
        pl.redraw_noise()
        pl.iterate()
        hist.append([x.copy() for x in pl.read_stats()])
    res[mode] = (pl.params.cpu().clone(), hist)
    print(mode, pl._branches, [h[0] for h in hist])
for a in ("branches", "sequential"):
    print(a, "vs eager max diff", float((res[a][0] - res["eager"][0]).abs().max()))
print("p1 vs eager[1]", float((res["p1"][0][0] - res["eager"][0][1]).abs().max()),
      "vs branches[1]", float((res["p1"][0][0] - res["branches"][0][1]).abs().max()),
      "vs sequential[1]", float((res["p1"][0][0] - res["sequential"][0][1]).abs().max()))
