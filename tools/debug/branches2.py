import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
from tests.helpers import fixture_model
from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner, JaxStraightLinePlan
m = fixture_model("reservoir")
for mode in ("branches", "sequential"):
    os.environ["RK_PARALLEL_BRANCHES"] = "1" if mode == "branches" else "0"
    pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=16, rollout_horizon=6,
                            optimizer="rmsprop", optimizer_kwargs={"learning_rate": 0.05}, pgpe=None,
                            parallel_updates=3)
    pl.initialize(5)
    def show(tag):
        torch.cuda.synchronize()
        print(mode, tag, "noise", float(pl.train_noise.double().sum()), "draws", int(pl._draws),
              "params", float(pl.params.double().sum()), "tmp", pl._train_loss_tmp.cpu().numpy(),
              "train_loss", pl.train_loss.cpu().numpy(), "grad", float(pl.grad.double().abs().sum()))
    show("init")
    pl.redraw_noise(); show("redraw")
    pl._capture(); show("captured")
    for p in range(3):
        pl._use_scratch(p); pl._train_grad_kernels(p)
    pl._use_scratch(0); pl._primed = True; show("primed")
    pl._graph.replay(); show("replay1")
    pl._graph.replay(); show("replay2")
