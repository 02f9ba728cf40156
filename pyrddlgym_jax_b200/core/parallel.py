"""Multi-GPU sharding of the rollout batch: one process per GPU, every rank owns a
contiguous slice of the rollouts, the plan / optimiser state / non-fluents are replicated, and
the only exchange of the path is one sum all-reduce of the policy gradient (+ the scalar
losses) per iteration.  The reference has no counterpart: it runs ``jax.vmap`` over the batch
on one device (compiler.py:575-577) and takes the batch mean inside ``value_and_grad``
(planner.py:2817-2818, 2873).

Everything here is backend-agnostic torch.distributed (NCCL on the GPUs, gloo in the CPU
tests); it never touches rollout data -- there is no data-path collective.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size) of the default process group, (0, 1) when there is none."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_global: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Rollouts [b0, b1) owned by ``rank``: contiguous, sizes differ by at most one."""
    base, rem = divmod(n_global, world_size)
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def slice_rollouts(flat: torch.Tensor, widths: Sequence[int], B: int, b0: int, b1: int,
                   lead: int = 1) -> torch.Tensor:
    """Rows [b0, b1) of every [B, n] block of a fluent-major buffer [lead, sum(n)*B]
    (state, tape, noise or trajectory layout of include/rk.h)."""
    flat = flat.reshape(lead, -1)
    out, off = [], 0
    for n in widths:
        blk = flat[:, off * B:(off + n) * B].reshape(lead, B, n)
        out.append(blk[:, b0:b1].reshape(lead, -1))
        off += n
    return torch.cat(out, dim=1).contiguous() if out else flat[:, :0]


def mean_cotangent(n_local: int, n_global: int, device=None) -> torch.Tensor:
    """d(-mean over the GLOBAL batch)/d return_b for the local rollouts (planner.py:2817)."""
    return torch.full((n_local,), -1.0 / n_global, dtype=torch.float32, device=device)


def allreduce_gradient(grad: torch.Tensor) -> torch.Tensor:
    """The exchange step: grad <- sum over ranks (in place).  Every rank then applies the
    same optimiser step to its replica, so replicas stay bit-identical without a broadcast."""
    if world()[1] > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM)
    return grad


def allgather_returns(local: torch.Tensor) -> torch.Tensor:
    """The GLOBAL return vector [world_size * B] in rank order (equal shards): what a utility other
    than the mean is evaluated on (planner.py:2161-2247)."""
    rank, ws = world()
    if ws == 1:
        return local
    out = torch.empty(ws * local.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


class PeerAllReduce:
    """One-shot sum all-reduce of a short fp32 vector over NVLink peer memory
    (csrc/rk_comm.cu): symmetric buffers and flag words are allocated with torch's symmetric
    memory allocator and their peer addresses exchanged ONCE here; every call is then a single
    kernel launch on the caller's stream (CUDA-graph capturable).  ``create`` returns None --
    and the caller keeps using NCCL -- when symmetric memory is unavailable, the vector is long
    (bandwidth-bound: NCCL's job), or RK_PEER_ALLREDUCE=0."""

    MAX_WORDS = 65536

    def __init__(self, n_max: int, device):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        from .engine import RkError  # noqa: F401  (library must be loadable)
        self.rank, self.world = world()
        self.n_max = int(n_max)
        group = dist.group.WORLD
        self.buf = symm.empty(2 * self.world * self.n_max, dtype=torch.float32, device=device)
        self.flags = symm.empty(2 * 16, dtype=torch.int32, device=device)
        self.buf.zero_()
        self.flags.zero_()
        hb = symm.rendezvous(self.buf, group=group)
        hf = symm.rendezvous(self.flags, group=group)
        U64 = ctypes.c_uint64 * self.world
        self.peer_bufs = U64(*[int(p) for p in hb.buffer_ptrs])
        self.peer_flags = U64(*[int(p) for p in hf.buffer_ptrs])
        self._handles = (hb, hf)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)     # (create()'s all-reduce is the barrier before first use)

    @classmethod
    def create(cls, n_max: int, device):
        import os
        if world()[1] <= 1 or n_max > cls.MAX_WORDS or world()[1] > 16 \
                or os.environ.get("RK_PEER_ALLREDUCE", "1") == "0" \
                or torch.device(device).type != "cuda":
            return None
        ok = torch.ones(1, dtype=torch.int32, device=device)
        obj = None
        try:
            obj = cls(n_max, device)
        except Exception:              # no symmetric memory / no peer access on this box
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)      # all ranks take the same path
        return obj if int(ok.item()) == 1 else None

    def __call__(self, vec: torch.Tensor) -> torch.Tensor:
        from . import engine
        engine.peer_allreduce(self.world, self.rank, self.peer_bufs, self.peer_flags, vec.numel(),
                              self.n_max, vec, self.epoch, self.status)
        return vec


    def gather(self, vec: torch.Tensor, extras, dst: torch.Tensor) -> torch.Tensor:
        """Sum all-reduce of ``vec`` (in place) and of the short vectors ``extras`` (sums -> ``dst``) as
        ONE message and one launch."""
        from . import engine
        engine.peer_allreduce_gather(self.world, self.rank, self.peer_bufs, self.peer_flags, vec.numel(),
                                     self.n_max, vec, self.epoch, self.status, list(extras), dst)
        return vec


def allreduce_losses(local_sums: torch.Tensor) -> torch.Tensor:
    """Sum of per-rank partial sums of returns (callers divide by the global batch)."""
    if world()[1] > 1:
        dist.all_reduce(local_sums, op=dist.ReduceOp.SUM)
    return local_sums


def allreduce_error_bits(err: torch.Tensor) -> torch.Tensor:
    """Bitwise OR over ranks of the per-step error masks (log['error'], compiler.py:539)."""
    if world()[1] > 1:
        dist.all_reduce(err, op=dist.ReduceOp.BOR)
    return err


def replicas_identical(t: torch.Tensor) -> bool:
    """True when every rank holds bit-identical ``t`` (debug check of the replication)."""
    rank, ws = world()
    if ws == 1:
        return True
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([int(torch.equal(ref, t))], device=t.device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    return bool(same.item())
