"""Runs the three rollout programs of the bench workload a few times (for ncu)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner, JaxStraightLinePlan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--horizon", type=int, default=10)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--domain", default="Quadcopter")
a = ap.parse_args()
from pyrddlgym_jax_b200.rddl.model import load_model  # noqa: E402
d = os.path.join(bench.ROOT, "pyrddlgym_jax_b200", "domains", a.domain)
inst = [f for f in sorted(os.listdir(d)) if f.startswith("instance")][0]
model = load_model(os.path.join(d, "domain.rddl"), os.path.join(d, inst))
pl = JaxBackpropPlanner(model, JaxStraightLinePlan(), batch_size_train=a.batch,
                        rollout_horizon=a.horizon, optimizer_kwargs={"learning_rate": 0.03},
                        use_cuda_graph=False)
pl.initialize(42)
for _ in range(a.reps):
    pl.iterate()
torch.cuda.synchronize()
print("ok", float(pl.train_loss[0]), float(pl.test_loss_buf[0]))
