"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on B200; gloo in the CPU tests).

Only the two places where the hot path shards naturally use a collective:
  * retrieval: the catalogue is split row-wise, every rank scores its shard with the fused kernel (global row ids via
    row_offset), one all-gather of (score, id)[Q, k] per rank, then a k-way select on every rank;
  * data-parallel training: one all-reduce of the flat dense-gradient buffer per step.
The local search / merge callables are injectable so the world_size-2 gloo tests can exercise the partitioning and
the exchange on CPU with the oracle standing in for the CUDA kernels.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) owned by `rank`: the first n_total % world ranks hold one extra row."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


class ShardedFlatIndex:
    """Row-sharded exact inner-product index: `search` returns the GLOBAL top-k on every rank.

    local_search(q, k) -> (scores [Q,k] fp32, ids [Q,k] int64 with GLOBAL row ids, -1 padding)
    merge(scores [P,Q,k], ids [P,Q,k], k) -> (scores [Q,k], ids [Q,k])   order: score desc, id asc
    """

    def __init__(self, local_search: Callable, merge: Callable, group=None):
        self.local_search = local_search
        self.merge = merge
        self.group = group

    @classmethod
    def from_device_index(cls, index, group=None) -> "ShardedFlatIndex":
        from . import kernels as K

        def local(q_op, k):
            return index.search_device(q_op, k)

        def merge(s, i, k):
            return K.topk_merge(s, i, k)

        return cls(local, merge, group)

    def search(self, queries, k: int):
        world, _ = _world()
        s, i = self.local_search(queries, k)
        if world == 1:
            return s, i
        q = s.shape[0]
        gs = torch.empty((world * q, s.shape[1]), dtype=s.dtype, device=s.device)
        gi = torch.empty((world * q, i.shape[1]), dtype=i.dtype, device=i.device)
        dist.all_gather_into_tensor(gs, s.contiguous(), group=self.group)
        dist.all_gather_into_tensor(gi, i.contiguous(), group=self.group)
        return self.merge(gs.view(world, q, -1), gi.view(world, q, -1), k)


def allreduce_mean_(flat: torch.Tensor, group=None) -> None:
    """Data-parallel gradient exchange: one all-reduce over the flat fp32 gradient buffer, then divide by the world."""
    world, _ = _world()
    if world == 1:
        return
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
