import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
from tests.helpers import fixture_model
from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner, JaxStraightLinePlan
m = fixture_model("reservoir")
for mode in ("0", "0", "1", "1", "0"):
    os.environ["RK_PARALLEL_BRANCHES"] = mode
    pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=16, rollout_horizon=6,
                            optimizer="rmsprop", optimizer_kwargs={"learning_rate": 0.05}, pgpe=None,
                            parallel_updates=3)
    pl.initialize(5)
    torch.cuda.synchronize()
    print(mode, pl._branches, pl._philox, [k for k, *_ in pl.train_fwd.program.noise_layout],
          "noise", float(pl.train_noise.double().sum()), float(pl.test_noise.double().sum()))
    pl.initialize(5)
    torch.cuda.synchronize()
    print("   again", float(pl.train_noise.double().sum()), float(pl.test_noise.double().sum()))
