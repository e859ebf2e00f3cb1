"""View-batch sharding across the GPUs of one box.

Views are independent in the forward pass, so every rank renders a contiguous slice of the camera
batch against a replicated mesh with no data-path collective.  The backward needs exactly one
exchange: parameters shared by all views (vertices, vertex colours / texture map, a shared pose)
receive per-rank partial gradients that are summed in ONE kernel over NVLink / NVSwitch peer memory
(``PeerAllReduce`` -> csrc/allreduce.cu: every rank pushes (epoch, value) words straight into all peers'
inboxes and sums what arrives in its own -- one one-way NVLink latency, no barrier; messages up to 256 KB,
e.g. 70 KB for the cow), with a single fused NCCL all-reduce for larger sums (12 MB for a 1M-face mesh:
bandwidth regime), when symmetric memory is not available, and on CPU / gloo.  Per-view camera gradients
stay local.

The reference has no distributed path at all (every script pins cuda:0); this is the
B200-native addition BASELINE.json's north_star asks for.
"""
from __future__ import annotations

import contextlib
import os
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


class PeerAllReduce:
    """One-shot sum all-reduce over NVLink / NVSwitch peer memory (``trb_allreduce_sum_f32``, csrc/allreduce.cu).

    Allocates an inbox in torch symmetric memory (every rank maps every peer's copy) once per (group,
    capacity); each call is ONE kernel: every rank pushes its values, tagged with the call's epoch, into all
    peers' inboxes and sums what arrives in its own.  Raises at construction when symmetric memory is
    unavailable (the caller falls back to NCCL)."""

    MAX_FLOATS = 1 << 16   # 256 KB messages; larger sums are bandwidth- not latency-bound: NCCL

    def __init__(self, capacity_floats: int, device: torch.device, group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._ctypes, self._lib = ctypes, _lib
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device
        self.capacity = (int(capacity_floats) + 1023) // 1024 * 1024
        self.buf = symm_mem.empty(2 * self.world * self.capacity * 2, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group=group)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=group)          # every inbox is zero (epoch 0) before anyone pushes
        self.peer_inbox = (ctypes.c_void_p * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        self.error = torch.zeros(1, dtype=torch.int32, device=device)
        self.epochs = torch.zeros(_lib.lib().trb_allreduce_grid(self.capacity), dtype=torch.int32, device=device)
        self.done = torch.zeros(1, dtype=torch.int32, device=device)   # block counter of the fused backward tail
        self._peer_sum = _lib.PeerSum(ctypes.cast(self.peer_inbox, ctypes.c_void_p), self.capacity, self.rank,
                                      self.world, self.epochs.data_ptr(), self.error.data_ptr(), self.done.data_ptr())
        self.calls = 0     # allreduce_shared_grads reads the error flag every _CHECK_EVERY calls

    def __call__(self, grads: Sequence[torch.Tensor]) -> None:
        ctypes, _lib = self._ctypes, self._lib
        n = len(grads)
        if n > 4 or sum(g.numel() for g in grads) > self.capacity:
            raise ValueError("PeerAllReduce: too many / too large segments for this inbox")
        if any(g.dtype != torch.float32 or not g.is_contiguous() for g in grads):
            raise ValueError("PeerAllReduce: segments must be contiguous float32")
        seg = (ctypes.c_void_p * n)(*[g.data_ptr() for g in grads])
        cnt = (ctypes.c_int64 * n)(*[g.numel() for g in grads])
        _lib.check(_lib.lib().trb_allreduce_sum_f32(
            seg, cnt, n, self.peer_inbox, self.capacity, self.rank, self.world, self.epochs.data_ptr(),
            self.error.data_ptr(), self.device.index, torch._C._cuda_getCurrentRawStream(self.device.index)),
            "peer all-reduce")

    def peer_sum(self, n_floats: int, device: torch.device):
        """The ``trb_peer_sum`` record for ``trb_render_backward_allreduce`` (ops._RenderFn.backward)."""
        if n_floats > self.capacity:
            raise ValueError(f"fused_backward_allreduce: {n_floats} shared gradient values do not fit the "
                             f"{self.capacity}-value inbox; use allreduce_shared_grads after the backward instead")
        if device != self.device:
            raise ValueError("fused_backward_allreduce: the render runs on another device than the inbox")
        return self._peer_sum

    def count_call(self) -> None:
        self.calls += 1
        if self.calls % _CHECK_EVERY == 0 and not torch.cuda.is_current_stream_capturing():
            self.check()

    def check(self) -> None:
        """Raises if a peer's data did not arrive (synchronises; call outside timed regions)."""
        if int(self.error.item()) != 0:
            raise RuntimeError("PeerAllReduce: a peer's contribution did not arrive within the spin limit; "
                               "the affected gradients were left un-reduced")


_peer_allreduce = {}
_CHECK_EVERY = 1024     # allreduce_shared_grads reads the error flag (one device sync) every this many calls


def _get_peer_allreduce(n_floats: int, device: torch.device, group):
    """The cached PeerAllReduce for this (group, device), grown when needed; None when peer memory is not usable."""
    key = (id(group), str(device))
    cur = _peer_allreduce.get(key)
    if cur is False:
        return None
    if n_floats > PeerAllReduce.MAX_FLOATS:
        return None
    if cur is None:
        try:
            cur = PeerAllReduce(PeerAllReduce.MAX_FLOATS, device, group)
        except Exception:  # noqa: BLE001 -- no symmetric memory / no P2P: NCCL does it
            cur = None
        # The route is agreed on COLLECTIVELY: if set-up failed on any rank, every rank takes NCCL (a rank that
        # pushed while a peer sat in dist.all_reduce would spin until the limit).
        ok = torch.tensor([0 if cur is None else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            _peer_allreduce[key] = False
            return None
        _peer_allreduce[key] = cur
    return cur


def check_peer_allreduce() -> None:
    """Raises if any peer-memory all-reduce issued so far timed out (synchronises the device)."""
    for v in _peer_allreduce.values():
        if v not in (None, False):
            v.check()


@contextlib.contextmanager
def fused_backward_allreduce(device=None, group=None):
    """Inside this context the backward of every fused render (``MeshRenderer`` / ``MeshRendererWithFragments`` /
    ``MeshRasterizer`` on the fused path) returns the gradients of the mesh's vertices and vertex colours already
    SUMMED over the ranks: the kernel that finalises them pushes them into the peers' memory from its last blocks
    and a short receive kernel adds what arrives (``trb_render_backward_allreduce``, include/trb.h) -- the step's one
    exchange rides on the tail of the backward instead of following it.  Per-view gradients (R, T, camera and
    light parameters) and UV-map gradients stay local.

        with fused_backward_allreduce() as fused:
            loss.backward()
        if not fused:                                   # no peer memory on this box / group: NCCL afterwards
            allreduce_shared_grads([verts.grad, colors.grad])

    Yields False (and changes nothing) when torch.distributed is not initialised, the world size is 1, the backend
    is not NCCL or peer memory cannot be set up (agreed on collectively).  Every rank must run the same renders in
    the same order; the mesh must be replicated (same vertex count on every rank, 6 V values <= 65,536) and every
    rank's batch non-empty.  Gradients that reach the shared parameters by other routes (regularisers computed
    identically on every rank) are not touched -- each rank ends with sum-over-ranks(render) + its own regulariser,
    the same tensor everywhere."""
    from . import ops
    peer = None
    if (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
            and dist.get_backend(group) == "nccl" and torch.cuda.is_available()
            and not os.environ.get("TRB_NCCL_ALLREDUCE")):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        peer = _get_peer_allreduce(1, dev, group)
    if peer is None:
        yield False
        return
    ops.set_backward_peer_sum(peer)
    try:
        yield True
    finally:
        ops.set_backward_peer_sum(None)


def allreduce_shared_grads(tensors: Sequence[Optional[torch.Tensor]], group=None, async_op: bool = False):
    """Sums the gradients of view-shared parameters across ranks with ONE all-reduce.

    ``tensors``: the ``.grad`` tensors (None entries are skipped; every rank must pass the same
    layout).  The result is written back in place.  No-op when torch.distributed is not initialised
    or the world size is 1."""
    grads = [t for t in tensors if t is not None]
    if not grads or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    # CUDA tensors on an NCCL group: the one-kernel peer-memory path (latency regime: 70 KB .. a few MB)
    if (not async_op and len(grads) <= 4 and all(g.is_cuda and g.dtype == torch.float32 and g.is_contiguous()
                                                 for g in grads)
            and dist.get_backend(group) == "nccl" and not os.environ.get("TRB_NCCL_ALLREDUCE")):
        peer = _get_peer_allreduce(sum(g.numel() for g in grads), grads[0].device, group)
        if peer is not None:
            peer(grads)
            peer.calls += 1
            if peer.calls % _CHECK_EVERY == 0 and not torch.cuda.is_current_stream_capturing():
                peer.check()
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
