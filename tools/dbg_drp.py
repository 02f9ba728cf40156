import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import planner as OP
from oracle.model_eval import OracleModel
from pyrddlgym_jax_b200.core.drp import JaxDeepReactivePolicy
from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner
from pyrddlgym_jax_b200.core.layout import noise_as_list
from tests.helpers import fixture_model
m = fixture_model("powergen")
T, B = 2, 4
plan = JaxDeepReactivePolicy(topology=[8, 8])
pl = JaxBackpropPlanner(m, plan, batch_size_train=B, rollout_horizon=T, optimizer="sgd",
                        optimizer_kwargs={"learning_rate": 0.0}, pgpe=None, use_cuda_graph=False)
pl.initialize(7)
g = torch.Generator().manual_seed(11)
flat = plan.initial_params(g)
flat[plan.seg["bh"][0]:plan.seg["bh"][0] + 5] = torch.tensor([2., 3., 4., 1., 5.])
flat[plan.seg["bh"][0] + 5:plan.seg["bh"][0] + 10] = -1.0
pl.params[0].copy_(flat)
fw = pl.train_fwd.program
noise = torch.randn(T, fw.N * B, generator=g)
pol_noise = torch.randn(T, 5 * B, generator=g)
pl.train_noise.copy_(noise.reshape(-1)); pl.drp.pol_noise.copy_(pol_noise); pl.drp.ent_coeff.fill_(0.0)
pl.update(); torch.cuda.synchronize()
tree = plan.params_to_pytree(flat)["params"]
om = OracleModel(m, pl.compiled.relax_config())
fls, nfls = om.init_fls_nfls(B)
a, e = OP.drp_actions(m, tree, {k: fls[k] for k in m.state_fluents}, {"curProd": pol_noise[0].reshape(B, 5)}, 1.0, m.bounds, 2)
print("s0 dev", pl.s0_train.cpu().view(6, B))
print("H1 dev", pl.drp.H[0][0].cpu()[:2])
x = torch.cat([fls[v].reshape(B, -1).float() for v in OP.drp_input_order(m)], 1)
print("x oracle", x)
print("H1 ora", torch.tanh(x @ tree["dense_0"]["kernel"] + tree["dense_0"]["bias"])[:2])
print("heads dev", pl.drp.heads[0].cpu()[:2])
print("act dev", pl.drp.act[0].cpu().view(5, B).t())
print("act ora", a["curProd"])
print("rew dev", pl.train_reward[0].view(T, B).cpu())
