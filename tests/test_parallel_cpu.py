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
