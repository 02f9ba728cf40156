"""Runs a few DRP planner iterations of the bench workload at a short horizon (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner, JaxDeepReactivePolicy

T = int(os.environ.get("T", 4))
wl = bench.WORKLOADS["powergen_drp"]
model = bench.workload_model("powergen_drp")
pl = JaxBackpropPlanner(model, JaxDeepReactivePolicy(**wl["plan_kwargs"]), batch_size_train=wl["B"],
                        rollout_horizon=T, optimizer_kwargs={"learning_rate": 1e-4}, pgpe=None,
                        use_cuda_graph=False)
pl.initialize(42)
for _ in range(2):
    pl.redraw_noise()
    pl.iterate()
torch.cuda.synchronize()
print("ok", float(pl.train_loss[0]), float(pl.test_loss_buf[0]))
