"""world_size-2 gloo tests of the multi-GPU path's host logic: each rank runs its
shard of the rollouts (the numpy ISA emulator stands in for the GPU), gradients are
sum-all-reduced, and the result must equal the single-rank run over the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyrddlgym_jax_b200.core import parallel


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 32, 4097):
        for ws in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_slice_rollouts_matches_pack_layout():
    B, widths, T = 5, [3, 1, 2], 4
    vals = [torch.arange(T * B * n, dtype=torch.float32).reshape(T, B, n) + 1000 * i
            for i, n in enumerate(widths)]
    flat = torch.cat([v.reshape(T, -1) for v in vals], dim=1)
    got = parallel.slice_rollouts(flat, widths, B, 1, 4, lead=T)
    want = torch.cat([v[:, 1:4].reshape(T, -1) for v in vals], dim=1)
    assert torch.equal(got, want)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, ws, port, key, B, T, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from tests.parity import gradient_case, run_forward, run_backward
        from tests.helpers import pack_params
        torch.manual_seed(0)
        c = gradient_case("emu", key, B=B, T=T)          # full batch: the single-rank answer
        m, fwd, bwd = c["model"], c["fwd"], c["bwd"]
        b0, b1 = parallel.shard_bounds(B, rank, ws)
        nb = b1 - b0
        s_w = [n for _, n, _ in fwd.state_layout.values()]
        n_w = [n for _, _, n, _, _ in fwd.noise_layout]
        s0 = parallel.slice_rollouts(c["s0"], s_w, B, b0, b1).reshape(-1)
        noise = parallel.slice_rollouts(c["noise"], n_w, B, b0, b1, lead=T)
        pf = pack_params(m, c["params"])
        got = run_forward("emu", fwd, nb, T, s0, pf, noise, c["disc"])
        gret = parallel.mean_cotangent(nb, B)
        grad = torch.from_numpy(run_backward("emu", bwd, nb, T, got["tape"], pf, noise, gret,
                                             c["disc"]).copy())
        parallel.allreduce_gradient(grad)
        loss = parallel.allreduce_losses(torch.tensor([-float(got["returns"].sum())])) / B
        full_ret = c["got"]["returns"]
        ok = (np.allclose(grad.numpy(), c["grad"], rtol=1e-5,
                          atol=1e-6 * np.abs(c["grad"]).max())
              and np.array_equal(got["returns"], full_ret[b0:b1])
              and abs(float(loss) - float(-full_ret.mean())) <= 1e-5 * abs(float(full_ret.mean()))
              and parallel.replicas_identical(grad))
        err = torch.from_numpy(np.bitwise_or.reduce(got["err"], axis=1).astype(np.int32))  # [T]
        parallel.allreduce_error_bits(err)
        ok = ok and np.array_equal(err.numpy().astype(np.uint32),
                                   np.bitwise_or.reduce(c["got"]["err"], axis=1))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("key,B", [("reservoir", 9), ("wildfire", 6)])
def test_two_rank_gradient_equals_single_rank(key, B):
    ws = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_rank_main, args=(ws, _free_port(), key, B, 3, out), nprocs=ws, join=True)
        assert dict(out) == {0: True, 1: True}


def _rank_utility(rank, ws, port, key, B, T, utility, kw, out):
    """Sharded non-mean utility: all-gather the returns, differentiate the utility of the GLOBAL
    vector, keep this rank's slice of the cotangent, sum the shard gradients -- against the single-rank
    gradient of the same utility over the whole batch."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from tests.parity import gradient_case, run_forward, run_backward
        from tests.helpers import pack_params
        from pyrddlgym_jax_b200.core.planner import UTILITY_LOOKUP
        torch.manual_seed(0)
        c = gradient_case("emu", key, B=B, T=T)
        m, fwd, bwd = c["model"], c["fwd"], c["bwd"]
        fn = UTILITY_LOOKUP[utility]

        def cotangent(returns):
            r = torch.as_tensor(np.asarray(returns), dtype=torch.float32).clone().requires_grad_(True)
            loss = -fn(r, **kw)
            (g,) = torch.autograd.grad(loss, r)
            return float(loss), g
        pf = pack_params(m, c["params"])
        # single rank, whole batch
        full = run_forward("emu", fwd, B, T, c["s0"].reshape(-1), pf, c["noise"], c["disc"])
        loss_full, g_full = cotangent(full["returns"])
        want = run_backward("emu", bwd, B, T, full["tape"], pf, c["noise"], g_full, c["disc"])
        # this rank's shard
        b0, b1 = parallel.shard_bounds(B, rank, ws)
        nb = b1 - b0
        assert nb * ws == B
        s_w = [n for _, n, _ in fwd.state_layout.values()]
        n_w = [n for _, _, n, _, _ in fwd.noise_layout]
        s0 = parallel.slice_rollouts(c["s0"], s_w, B, b0, b1).reshape(-1)
        noise = parallel.slice_rollouts(c["noise"], n_w, B, b0, b1, lead=T)
        got = run_forward("emu", fwd, nb, T, s0, pf, noise, c["disc"])
        allr = parallel.allgather_returns(torch.as_tensor(np.asarray(got["returns"]), dtype=torch.float32))
        loss, g = cotangent(allr)
        grad = torch.from_numpy(run_backward("emu", bwd, nb, T, got["tape"], pf, noise,
                                             g[b0:b1].contiguous(), c["disc"]).copy())
        parallel.allreduce_gradient(grad)
        ok = (np.allclose(allr.numpy(), np.asarray(full["returns"]))
              and abs(loss - loss_full) <= 1e-6 * max(1.0, abs(loss_full))
              and np.allclose(grad.numpy(), want, rtol=1e-5, atol=1e-6 * np.abs(want).max())
              and float(np.abs(want).max()) > 0 and parallel.replicas_identical(grad))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("utility,kw", [("cvar", {"alpha": 0.4}), ("mean_var", {"beta": 0.5}),
                                        ("gini", {"gamma": 0.3})])
def test_two_rank_nonmean_utility_gradient(utility, kw):
    ws = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_rank_utility, args=(ws, _free_port(), "reservoir", 10, 3, utility, kw, out), nprocs=ws,
                 join=True)
        assert dict(out) == {0: True, 1: True}
