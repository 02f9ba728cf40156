import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
from tests.helpers import fixture_model
from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner, JaxStraightLinePlan
from pyrddlgym_jax_b200.core.layout import draw_noise
m = fixture_model("reservoir")
for mode in ("0", "1", "0"):
    os.environ["RK_PARALLEL_BRANCHES"] = mode
    pl = JaxBackpropPlanner(m, JaxStraightLinePlan(), batch_size_train=16, rollout_horizon=6,
                            optimizer="rmsprop", optimizer_kwargs={"learning_rate": 0.05}, pgpe=None,
                            parallel_updates=3)
    lay = pl.train_fwd.program.noise_layout
    print(mode, "layout", [(k, s, n, off, np.asarray(e).reshape(-1)[:4]) for k, s, n, off, e in lay])
    pl._gen.manual_seed(5)
    print("  offset0", pl._gen.get_offset())
    x = draw_noise(lay, 6, 16, pl._gen, pl.device)
    print("  direct", float(x.double().sum()), "offset", pl._gen.get_offset())
    pl._gen.manual_seed(5)
    a = torch.full((6, 16, 3), 2.0, device=pl.device)
    y = torch._standard_gamma(a, generator=pl._gen)
    print("  raw gamma", float(y.double().sum()), "offset", pl._gen.get_offset())
    pl.initialize(5)
    torch.cuda.synchronize()
    print("  init", float(pl.train_noise.double().sum()), "offset", pl._gen.get_offset())
