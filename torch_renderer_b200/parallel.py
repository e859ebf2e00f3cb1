"""View-batch sharding across the GPUs of one box.

Views are independent in the forward pass, so every rank renders a contiguous slice of the camera
batch against a replicated mesh with no data-path collective.  The backward needs exactly one
exchange: parameters shared by all views (vertices, vertex colours / texture map, a shared pose)
receive per-rank partial gradients that are summed with ONE NCCL all-reduce over a single fused
fp32 buffer (35 KB for the cow .. 12 MB for a 1M-face mesh -- latency regime on NVLink 5 / NVSwitch,
so one launch per step instead of one per tensor).  Per-view camera gradients stay local.

The reference has no distributed path at all (every script pins cuda:0); this is the
B200-native addition BASELINE.json's north_star asks for.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of the view batch owned by ``rank`` (sizes differ by <= 1)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, rem = divmod(n_views, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def chunk_views(n_views: int, max_views_per_chunk: int) -> List[Tuple[int, int]]:
    """Splits a rank's views into chunks that bound Fragments memory (28*K*H*W bytes per view)."""
    if max_views_per_chunk < 1:
        raise ValueError("max_views_per_chunk must be >= 1")
    return [(s, min(s + max_views_per_chunk, n_views)) for s in range(0, n_views, max_views_per_chunk)]


def max_views_for_memory(H: int, W: int, K: int, budget_bytes: int, with_grad: bool = True) -> int:
    """Views per chunk such that Fragments (+ their gradients) fit in ``budget_bytes``."""
    per_view = 28 * K * H * W + 16 * H * W
    if with_grad:
        per_view += 20 * K * H * W + 16 * H * W
    return max(1, budget_bytes // per_view)


def allreduce_shared_grads(tensors: Sequence[Optional[torch.Tensor]], group=None, async_op: bool = False):
    """Sums the gradients of view-shared parameters across ranks with ONE all-reduce.

    ``tensors``: the ``.grad`` tensors (None entries are skipped; every rank must pass the same
    layout).  The result is written back in place.  No-op when torch.distributed is not initialised
    or the world size is 1."""
    grads = [t for t in tensors if t is not None]
    if not grads or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def finish():
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off: off + n].view_as(g))
            off += n

    if async_op:
        return work, finish
    finish()
    return None
