"""Autograd wrappers over the C ABI (``include/trb.h``).  Host plumbing only: every tensor is
allocated here through torch's caching allocator and handed to ``libtrb.so`` as a raw pointer with
the current stream.  Mirrors PyTorch3D's ``_RasterizeFaceVerts`` / ``_InterpFaceAttrs`` autograd
boundary (SURVEY.md layer L3); the reference reaches it through ``MeshRasterizer`` /
``MeshRenderer`` (torch_renderer.py:97-108, camera_pose_optimizer.py:130-158).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import _lib
from ._lib import ShadeConfig, check

_launch_count = 0  # kernels launched through libtrb.so (bench.py reports it)


def launch_count() -> int:
    return _launch_count


def _bump(n: int) -> None:
    global _launch_count
    _launch_count += n


# Optional per-call CUDA-event timing (bench.py's roofline leg).  None = off (the default).
_event_log = None


def start_event_log() -> None:
    global _event_log
    _event_log = []


def stop_event_log():
    """Returns {call name: (launches, total milliseconds)}; synchronises the device."""
    global _event_log
    log, _event_log = _event_log or [], None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in log:
        n, ms = out.get(name, (0, 0.0))
        out[name] = (n + 1, ms + e0.elapsed_time(e1))
    return out


class _timed:
    """Brackets one C-ABI call with CUDA events on the launching stream when the log is on."""

    def __init__(self, name: str, device: torch.device):
        self.name, self.device = name, device

    def __enter__(self):
        if _event_log is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if _event_log is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream(self.device))
            _event_log.append((self.name, self.e0, e1))
        return False


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: torch_renderer_b200 runs on CUDA tensors only (got device {t.device}); "
            "there is no CPU fallback")


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    # the raw cudaStream_t of torch's current stream (torch.cuda.current_stream() builds a Stream object: 13 us)
    return torch._C._cuda_getCurrentRawStream(device.index)


_EMPTY_F32 = {}


def _empty_f32(dev: torch.device) -> torch.Tensor:
    t = _EMPTY_F32.get(dev)
    if t is None:
        t = _EMPTY_F32[dev] = torch.empty((0,), dtype=torch.float32, device=dev)
    return t


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# --------------------------------------------------------------------------------------------
@dataclass
class ViewTable:
    """Device copy of the ``trb_view[N]`` table plus the host-side sizes the launches need."""
    views: torch.Tensor          # int32 [N, 8] on device
    host: torch.Tensor           # same on CPU
    N: int
    max_face_count: int
    max_vert_count: int
    total_ndc_verts: int
    total_faces_packed: int      # upper bound (exclusive) of pix_to_face values
    shared_mesh: bool
    pair_capacity: int = 0
    _pending: object = None      # (pinned stats, event) of the last forward
    _plans: dict = field(default_factory=dict)   # _RenderFn: static config / sizes / buffer layout per settings
    _stats_host: object = None   # pinned int32[4], allocated once
    _stats_event: object = None
    _calls: int = 0

    def want_stats(self) -> bool:
        """Whether this forward should report its bin statistics (pair-list demand) back to the host: the first
        calls (capacity is an estimate until then), afterwards one call in 16 -- an overflowing tile only costs
        speed (it falls back to a whole-mesh scan), never faces."""
        if self._pending is not None or torch.cuda.is_current_stream_capturing():
            return False
        self._calls += 1
        return self._calls <= 4 or (self._calls & 15) == 0

    def arm_stats(self, stats_dev: torch.Tensor, device: torch.device) -> None:
        if self._stats_host is None:
            self._stats_host = torch.empty((4,), dtype=torch.int32, pin_memory=True)
            self._stats_event = torch.cuda.Event()
        self._stats_host.copy_(stats_dev, non_blocking=True)
        self._stats_event.record(torch.cuda.current_stream(device))
        self._pending = (self._stats_host, self._stats_event)

    @staticmethod
    def build(face_start, face_count, p2f_base, world_vert_start, vert_count, device, shared_mesh):
        n = len(face_count)
        host = torch.zeros((n, 8), dtype=torch.int32)
        ndc_start = 0
        for i in range(n):
            host[i, 0] = face_start[i]
            host[i, 1] = face_count[i]
            host[i, 2] = ndc_start - world_vert_start[i]
            host[i, 3] = p2f_base[i]
            host[i, 4] = world_vert_start[i]
            host[i, 5] = vert_count[i]
            host[i, 6] = ndc_start
            ndc_start += vert_count[i]
        total_faces = max((p2f_base[i] + face_count[i] for i in range(n)), default=0)
        if ndc_start * 3 >= 2**31 or total_faces >= 2**31:
            raise ValueError("batch too large for int32 indexing; render the views in chunks")
        return ViewTable(views=host.to(device), host=host, N=n,
                         max_face_count=max(face_count, default=0),
                         max_vert_count=max(vert_count, default=0), total_ndc_verts=ndc_start,
                         total_faces_packed=total_faces, shared_mesh=shared_mesh)

    def default_pair_capacity(self, tiles_per_face: int = 8) -> int:
        total_fv = int(self.host[:, 1].sum())
        return int(min(max(max(8, tiles_per_face) * total_fv, 1 << 16), 1 << 28))

    def poll_capacity(self, tiles_per_face: int = 8) -> int:
        """Non-blocking: grows the (tile, face) pair capacity when an earlier forward reported
        that some tiles had to fall back to a whole-mesh scan.  Never synchronises.  ``tiles_per_face``: how many
        tiles the blur-inflated box of a small face touches (the first estimate; a 15-pixel blur band makes every
        face of a fine mesh touch ~16 tiles, and a tile without room scans the whole mesh -- correct, but the
        first render of the 1M-face sphere took 0.8 s that way)."""
        if self.pair_capacity == 0:
            self.pair_capacity = self.default_pair_capacity(tiles_per_face)
        if self._pending is not None and not torch.cuda.is_current_stream_capturing():
            stats, event = self._pending
            if event.query():
                needed = int(stats[0])
                if needed > self.pair_capacity:
                    self.pair_capacity = int(min(needed * 5 // 4 + 1024, (1 << 31) - 1))
                self._pending = None
        return self.pair_capacity


# --------------------------------------------------------------------------------------------
class _TransformFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts_world, R, T, proj, table: ViewTable, perspective: bool):
        _require_cuda(verts_world, "transform")
        verts_world, R, T, proj = _f32c(verts_world), _f32c(R), _f32c(T), _f32c(proj)
        dev = verts_world.device
        out = torch.empty((table.total_ndc_verts, 3), dtype=torch.float32, device=dev)
        with _timed("transform_forward", dev):
            check(_lib.lib().trb_transform_forward(
                _ptr(verts_world), _ptr(R), _ptr(T), _ptr(proj), _ptr(table.views), table.N,
                table.max_vert_count, int(perspective), _ptr(out), dev.index, _stream(dev)), "transform")
        _bump(1)
        ctx.save_for_backward(verts_world, R, T, proj)
        ctx.table, ctx.perspective = table, perspective
        return out

    @staticmethod
    def backward(ctx, grad_out):
        verts_world, R, T, proj = ctx.saved_tensors
        table = ctx.table
        dev = verts_world.device
        need = ctx.needs_input_grad
        g_v = torch.zeros_like(verts_world) if need[0] else None
        g_R = torch.zeros_like(R) if need[1] else None
        g_T = torch.zeros_like(T) if need[2] else None
        g_p = torch.zeros_like(proj) if need[3] else None
        grad_out = _f32c(grad_out)
        with _timed("transform_backward", dev):
            check(_lib.lib().trb_transform_backward(
                _ptr(verts_world), _ptr(R), _ptr(T), _ptr(proj), _ptr(table.views), table.N,
                table.max_vert_count, int(ctx.perspective), _ptr(grad_out), _ptr(g_v), _ptr(g_R), _ptr(g_T),
                _ptr(g_p), dev.index, _stream(dev)), "transform backward")
        _bump(1)
        return g_v, g_R, g_T, g_p, None, None


def transform_verts(verts_world, R, T, proj, table: ViewTable, perspective: bool = True):
    """world -> NDC for every (view, vertex); returns f32 [table.total_ndc_verts, 3]."""
    return _TransformFn.apply(verts_world, R, T, proj, table, perspective)


# --------------------------------------------------------------------------------------------
class _RasterizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts_ndc, faces, table: ViewTable, H, W, K, blur_radius, flags, neighbor=None):
        _require_cuda(verts_ndc, "rasterize_meshes")
        verts_ndc = _f32c(verts_ndc)
        dev = verts_ndc.device
        L = _lib.lib()
        N = table.N
        cap = table.poll_capacity()
        nbytes = ctypes.c_size_t(0)
        check(L.trb_raster_workspace_bytes(N, H, W, K, cap, ctypes.byref(nbytes)), "rasterize_meshes")
        ws = torch.empty((nbytes.value,), dtype=torch.uint8, device=dev)
        p2f = torch.empty((N, H, W, K), dtype=torch.int64, device=dev)
        zbuf = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
        bary = torch.empty((N, H, W, K, 3), dtype=torch.float32, device=dev)
        dists = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
        stats = torch.empty((4,), dtype=torch.int32, device=dev)
        with _timed("raster_forward", dev):
            check(L.trb_raster_forward(
                _ptr(verts_ndc), _ptr(faces), _ptr(table.views), N, table.max_face_count, H, W, K,
                float(blur_radius), int(flags), cap, _ptr(ws), nbytes.value, _ptr(p2f), _ptr(zbuf),
                _ptr(bary), _ptr(dists), _ptr(stats), dev.index, _stream(dev)), "rasterize_meshes")
        _bump(6 if table.max_face_count > 0 else 3)
        if neighbor is not None:
            # faces cut into quadrilaterals by clip_faces: at most one half per pixel (trb_clip_resequence)
            nbr = neighbor.to(torch.int32).contiguous()
            rows = torch.arange(nbr.shape[0], dtype=torch.int32, device=dev)
            pair_face = (nbr == rows + 1).nonzero(as_tuple=True)[0].to(torch.int32)
            if pair_face.numel() > 0:
                starts = table.views[:, 0].contiguous()
                pair_view = (torch.searchsorted(starts, pair_face, right=True) - 1).to(torch.int32)
                with _timed("clip_resequence", dev):
                    check(L.trb_clip_resequence(
                        _ptr(verts_ndc), _ptr(table.views), _ptr(pair_face), _ptr(pair_view), pair_face.numel(),
                        _ptr(nbr), N, H, W, K, float(blur_radius), int(flags), _ptr(p2f), _ptr(zbuf), _ptr(bary),
                        _ptr(dists), 0, dev.index, _stream(dev)), "rasterize_meshes (clipped faces)")
                _bump(1)
        if table.want_stats():
            table.arm_stats(stats, dev)
        ctx.save_for_backward(verts_ndc, faces, p2f)
        ctx.table, ctx.dims, ctx.flags = table, (H, W, K), flags
        ctx.mark_non_differentiable(p2f)
        ctx.set_materialize_grads(False)
        return p2f, zbuf, bary, dists

    @staticmethod
    def backward(ctx, _gp2f, g_zbuf, g_bary, g_dists):
        verts_ndc, faces, p2f = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return (None,) * 9
        table = ctx.table
        H, W, K = ctx.dims
        dev = verts_ndc.device
        g_verts = torch.zeros_like(verts_ndc)
        if (g_zbuf is None and g_bary is None and g_dists is None) or verts_ndc.numel() == 0:
            return (g_verts,) + (None,) * 8
        g_zbuf = None if g_zbuf is None else _f32c(g_zbuf)
        g_bary = None if g_bary is None else _f32c(g_bary)
        g_dists = None if g_dists is None else _f32c(g_dists)
        with _timed("raster_backward", dev):
            check(_lib.lib().trb_raster_backward(
                _ptr(verts_ndc), _ptr(faces), _ptr(table.views), table.N, H, W, K, int(ctx.flags), _ptr(p2f),
                _ptr(g_zbuf), _ptr(g_bary), _ptr(g_dists), _ptr(g_verts), dev.index, _stream(dev)),
                "rasterize_meshes backward")
        _bump(1)
        return (g_verts,) + (None,) * 8


class _BehindState:
    """Per-device state of the near-plane question: the persistent device flag, a pinned host word the library
    copies it into, the event it records after that copy, and the epoch counter."""

    def __init__(self, dev: torch.device):
        self.flag = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.host = torch.zeros((1,), dtype=torch.int32).pin_memory()
        self.host_word = ctypes.c_int32.from_address(self.host.data_ptr())
        self.event = torch.cuda.Event()
        with torch.cuda.device(dev):
            self.event.record()          # materialises the cudaEvent_t handle
        self.event.synchronize()
        self.epoch = 0


_behind_states = {}   # device index -> _BehindState


def any_vertex_behind_async(verts_world, R, T, table: ViewTable, z_plane: float):
    """Asks the device whether some (view, vertex) has view-space depth < ``z_plane``: one small kernel, a 4-byte
    copy into pinned memory and an event record, all enqueued by ONE C-ABI call.  Returns ``answer()``, which blocks
    until that copy has landed -- NOT until later work on the stream has run -- so a caller can enqueue the render
    it expects to keep first and read the answer afterwards without leaving the GPU idle.  One question per device
    may be outstanding (ask, enqueue, answer)."""
    _require_cuda(verts_world, "near-plane test")
    dev = verts_world.device
    st = _behind_states.get(dev.index)
    if st is None or st.epoch >= 2**31 - 2:
        st = _behind_states[dev.index] = _BehindState(dev)
    st.epoch += 1
    epoch = st.epoch
    verts_world, R, T = _f32c(verts_world), _f32c(R), _f32c(T)
    check(_lib.lib().trb_any_vertex_behind(_ptr(verts_world), _ptr(R), _ptr(T), _ptr(table.views), table.N,
                                           table.max_vert_count, float(z_plane), epoch, _ptr(st.flag),
                                           st.host.data_ptr(), st.event.cuda_event, dev.index, _stream(dev)),
          "near-plane test")
    _bump(1)

    def answer() -> bool:
        st.event.synchronize()
        # epochs grow: a later query on another stream may have raised the flag past ours before the copy ran --
        # then the answer errs towards True, which only sends the caller to clip_faces (exact) for nothing
        return st.host_word.value >= epoch
    return answer


class NearPlaneWatch:
    """The near-plane question for renders that cannot wait for the answer (CUDA-graph capture, ``capture.py``):
    the same test kernel raises a STICKY device flag (epoch 1, never lowered) and the stream -- or the captured
    graph -- copies it into a pinned host word after every test; ``tripped()`` reads that word without
    synchronising."""

    def __init__(self, dev: torch.device):
        self.device = dev
        self.flag = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.host = torch.zeros((1,), dtype=torch.int32).pin_memory()
        self.host_word = ctypes.c_int32.from_address(self.host.data_ptr())

    def enqueue(self, verts_world, R, T, table: ViewTable, z_plane: float) -> None:
        _require_cuda(verts_world, "near-plane test")
        dev = verts_world.device
        verts_world, R, T = _f32c(verts_world), _f32c(R), _f32c(T)
        check(_lib.lib().trb_any_vertex_behind(_ptr(verts_world), _ptr(R), _ptr(T), _ptr(table.views), table.N,
                                               table.max_vert_count, float(z_plane), 1, _ptr(self.flag),
                                               self.host.data_ptr(), None, dev.index, _stream(dev)),
              "near-plane test")
        _bump(1)

    def tripped(self) -> bool:
        return self.host_word.value >= 1


_active_watch = None     # set by capture.CapturedStep while it warms up / captures a step


class near_plane_watch:
    """Context manager: renders inside ask the near-plane question through ``watch`` (no host wait) instead of
    the blocking / skipped forms."""

    def __init__(self, watch: NearPlaneWatch):
        self.watch = watch

    def __enter__(self):
        global _active_watch
        self.prev, _active_watch = _active_watch, self.watch
        return self.watch

    def __exit__(self, *exc):
        global _active_watch
        _active_watch = self.prev
        return False


def active_near_plane_watch():
    return _active_watch


def any_vertex_behind(verts_world, R, T, table: ViewTable, z_plane: float) -> bool:
    """Blocking form of ``any_vertex_behind_async`` (PyTorch3D's ``clip_faces`` reads two sums the same way)."""
    return any_vertex_behind_async(verts_world, R, T, table, z_plane)()


def rasterize_face_verts(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, image_size, blur_radius=0.0,
                         faces_per_pixel=1, perspective_correct=False, clip_barycentric_coords=False,
                         cull_backfaces=False, clipped_faces_neighbor_idx=None):
    """The ``_C.rasterize_meshes`` signature itself: ``face_verts`` f32 (F, 3, 3) in NDC x, y + view z, per-mesh
    ranges i64 (N,), optionally the neighbour table of ``clip.clip_faces``.  ``pix_to_face`` indexes ``face_verts``.
    Reads the two range tensors on the host (this is the clipped-faces path, not the hot one)."""
    H, W = image_size
    if faces_per_pixel > _lib.MAX_FACES_PER_PIXEL:
        raise ValueError(f"faces_per_pixel must be <= {_lib.MAX_FACES_PER_PIXEL}")
    if face_verts.dim() != 3 or tuple(face_verts.shape[1:]) != (3, 3):
        raise ValueError("face_verts must have shape (F, 3, 3)")
    first = [int(v) for v in mesh_to_face_first_idx.tolist()]
    count = [int(v) for v in num_faces_per_mesh.tolist()]
    table = ViewTable.build(face_start=first, face_count=count, p2f_base=first, world_vert_start=[3 * f for f in first],
                            vert_count=[3 * c for c in count], device=face_verts.device, shared_mesh=False)
    flags = ((_lib.PERSPECTIVE_CORRECT if perspective_correct else 0)
             | (_lib.CLIP_BARYCENTRIC if clip_barycentric_coords else 0)
             | (_lib.CULL_BACKFACES if cull_backfaces else 0))
    return _RasterizeFn.apply(face_verts, None, table, int(H), int(W), int(faces_per_pixel), float(blur_radius),
                              flags, clipped_faces_neighbor_idx)


def rasterize(verts_ndc, faces, table: ViewTable, image_size, blur_radius=0.0, faces_per_pixel=1,
              perspective_correct=False, clip_barycentric_coords=False, cull_backfaces=False):
    """Returns (pix_to_face i64 [N,H,W,K], zbuf, bary_coords [N,H,W,K,3], dists)."""
    H, W = image_size
    if faces_per_pixel > _lib.MAX_FACES_PER_PIXEL:
        raise ValueError(f"faces_per_pixel must be <= {_lib.MAX_FACES_PER_PIXEL}")
    flags = ((_lib.PERSPECTIVE_CORRECT if perspective_correct else 0)
             | (_lib.CLIP_BARYCENTRIC if clip_barycentric_coords else 0)
             | (_lib.CULL_BACKFACES if cull_backfaces else 0))
    return _RasterizeFn.apply(verts_ndc, faces, table, int(H), int(W), int(faces_per_pixel),
                              float(blur_radius), flags)


# --------------------------------------------------------------------------------------------
class _InterpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pix_to_face, bary, face_attrs):
        _require_cuda(bary, "interpolate_face_attributes")
        bary, face_attrs = _f32c(bary), _f32c(face_attrs)
        pix_to_face = pix_to_face.contiguous()
        dev = bary.device
        P, F, D = pix_to_face.numel(), face_attrs.shape[0], face_attrs.shape[2]
        out = torch.empty(tuple(pix_to_face.shape) + (D,), dtype=torch.float32, device=dev)
        with _timed("interp_forward", dev):
            check(_lib.lib().trb_interp_forward(_ptr(pix_to_face), _ptr(bary), _ptr(face_attrs), P, F, D,
                                                _ptr(out), dev.index, _stream(dev)),
                  "interpolate_face_attributes")
        _bump(1)
        ctx.save_for_backward(pix_to_face, bary, face_attrs)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        pix_to_face, bary, face_attrs = ctx.saved_tensors
        dev = bary.device
        P, F, D = pix_to_face.numel(), face_attrs.shape[0], face_attrs.shape[2]
        g_bary = torch.empty_like(bary)
        g_attrs = torch.zeros_like(face_attrs)
        grad_out = _f32c(grad_out)
        with _timed("interp_backward", dev):
            check(_lib.lib().trb_interp_backward(_ptr(pix_to_face), _ptr(bary), _ptr(face_attrs),
                                                 _ptr(grad_out), P, F, D, _ptr(g_bary), _ptr(g_attrs),
                                                 dev.index, _stream(dev)),
                  "interpolate_face_attributes backward")
        _bump(1)
        return None, g_bary, g_attrs


def interpolate_face_attributes(pix_to_face, barycentric_coords, face_attributes):
    """PyTorch3D ``pytorch3d.ops.interpolate_face_attributes`` twin: (N,H,W,K), (N,H,W,K,3),
    (F,3,D) -> (N,H,W,K,D)."""
    F, FV, D = face_attributes.shape
    if FV != 3:
        raise ValueError("Faces can only have three vertices; got %r" % FV)
    N, H, W, K, _ = barycentric_coords.shape
    if pix_to_face.shape != (N, H, W, K):
        raise ValueError("pix_to_face must have shape (batch_size, H, W, K); got %r" % (pix_to_face.shape,))
    return _InterpFn.apply(pix_to_face, barycentric_coords, face_attributes)


# --------------------------------------------------------------------------------------------
class _VertexNormalsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, faces):
        _require_cuda(verts, "verts_normals")
        verts = _f32c(verts)
        dev = verts.device
        V, F = verts.shape[0], faces.shape[0]
        raw = torch.empty_like(verts)
        normals = torch.empty_like(verts)
        with _timed("vertex_normals_forward", dev):
            check(_lib.lib().trb_vertex_normals_forward(_ptr(verts), _ptr(faces), V, F, _ptr(raw),
                                                        _ptr(normals), dev.index, _stream(dev)),
                  "verts_normals")
        _bump(2)
        ctx.save_for_backward(verts, faces, raw)
        return normals

    @staticmethod
    def backward(ctx, grad_normals):
        verts, faces, raw = ctx.saved_tensors
        dev = verts.device
        V, F = verts.shape[0], faces.shape[0]
        g_raw = torch.empty_like(verts)
        g_verts = torch.zeros_like(verts)
        grad_normals = _f32c(grad_normals)
        with _timed("vertex_normals_backward", dev):
            check(_lib.lib().trb_vertex_normals_backward(_ptr(verts), _ptr(faces), V, F, _ptr(raw),
                                                         _ptr(grad_normals), _ptr(g_raw), _ptr(g_verts),
                                                         dev.index, _stream(dev)), "verts_normals backward")
        _bump(2)
        return g_verts, None


def vertex_normals(verts, faces_i32):
    """Area-weighted unit vertex normals, differentiable w.r.t. ``verts``."""
    return _VertexNormalsFn.apply(verts, faces_i32)


# --------------------------------------------------------------------------------------------
class _ShadeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bary, zbuf, dists, verts, normals, colors, texels, view_params, pix_to_face, faces,
                table: ViewTable, cfg: ShadeConfig):
        _require_cuda(dists, "shader")
        dev = dists.device
        bary, zbuf, dists = _f32c(bary), _f32c(zbuf), _f32c(dists)
        verts = None if verts is None else _f32c(verts)
        normals = None if normals is None else _f32c(normals)
        colors = None if colors is None else _f32c(colors)
        texels = None if texels is None else _f32c(texels)
        view_params = None if view_params is None else _f32c(view_params)
        images = torch.empty((cfg.N, cfg.H, cfg.W, 4), dtype=torch.float32, device=dev)
        with _timed("shade_forward", dev):
            check(_lib.lib().trb_shade_forward(
                ctypes.byref(cfg), _ptr(None if table is None else table.views), _ptr(view_params),
                _ptr(pix_to_face), _ptr(bary), _ptr(zbuf), _ptr(dists), _ptr(faces), _ptr(verts),
                _ptr(normals), _ptr(colors), _ptr(texels), _ptr(images), dev.index, _stream(dev)), "shader")
        _bump(1)
        ctx.save_for_backward(bary, zbuf, dists, verts, normals, colors, texels, view_params, pix_to_face,
                              faces)
        ctx.table, ctx.cfg = table, cfg
        return images

    @staticmethod
    def backward(ctx, grad_images):
        bary, zbuf, dists, verts, normals, colors, texels, view_params, pix_to_face, faces = ctx.saved_tensors
        table, cfg = ctx.table, ctx.cfg
        dev = dists.device
        need = ctx.needs_input_grad
        geom = need[0] or need[1] or need[2]
        g_bary = torch.empty_like(bary) if geom else None
        g_zbuf = torch.empty_like(zbuf) if geom else None
        g_dists = torch.empty_like(dists) if geom else None
        g_verts = torch.zeros_like(verts) if (need[3] and verts is not None) else None
        g_normals = torch.zeros_like(normals) if (need[4] and normals is not None) else None
        g_colors = torch.zeros_like(colors) if (need[5] and colors is not None) else None
        g_texels = torch.empty_like(texels) if (need[6] and texels is not None) else None
        g_vp = torch.zeros_like(view_params) if (need[7] and view_params is not None) else None
        grad_images = _f32c(grad_images)
        with _timed("shade_backward", dev):
            check(_lib.lib().trb_shade_backward(
                ctypes.byref(cfg), _ptr(None if table is None else table.views), _ptr(view_params),
                _ptr(pix_to_face), _ptr(bary), _ptr(zbuf), _ptr(dists), _ptr(faces), _ptr(verts),
                _ptr(normals), _ptr(colors), _ptr(texels), _ptr(grad_images), _ptr(g_bary), _ptr(g_zbuf),
                _ptr(g_dists), _ptr(g_verts), _ptr(g_normals), _ptr(g_colors), _ptr(g_texels), _ptr(g_vp),
                dev.index, _stream(dev)), "shader backward")
        _bump(1)
        return (g_bary if need[0] else None, g_zbuf if need[1] else None, g_dists if need[2] else None,
                g_verts, g_normals, g_colors, g_texels, g_vp, None, None, None, None)


def shade(bary, zbuf, dists, verts, normals, colors, texels, view_params, pix_to_face, faces, table, cfg):
    return _ShadeFn.apply(bary, zbuf, dists, verts, normals, colors, texels, view_params, pix_to_face,
                          faces, table, cfg)


# --------------------------------------------------------------------------------------------
class _RenderFn(torch.autograd.Function):
    """The fused pipeline: one C-ABI call forward (normals, camera centres, world->NDC, binning, fine
    rasterisation + shading epilogue) and one backward.  ``spec`` carries the static configuration."""

    @staticmethod
    def forward(ctx, verts, colors, tex_map, R, T, proj, view_params, faces, table: ViewTable, spec: dict):
        _require_cuda(verts, "render")
        dev = verts.device
        L = _lib.lib()
        verts, R, T, proj = _f32c(verts), _f32c(R), _f32c(T), _f32c(proj)
        colors = None if colors is None else _f32c(colors)
        tex_map = None if tex_map is None else _f32c(tex_map)
        # the kernel may write the camera centres into the block: it works on a copy that its first kernel makes
        # (trb_render_extras.view_params_src) inside the internal buffer -- no clone node per render
        vp_src = None if view_params is None else _f32c(view_params)
        N, (H, W), K = table.N, spec["image_size"], spec["K"]
        shader = spec["shader"]
        phong = shader in (_lib.SHADER_SOFT_PHONG, _lib.SHADER_HARD_PHONG)
        want_light_grad = int(view_params is not None and view_params.requires_grad)
        sparse = int(bool(spec.get("sparse", False)) and shader != _lib.SHADER_NONE)
        tile = 16 if K <= 24 else 8
        blur_px = (spec["blur_radius"] ** 0.5) * min(H, W) / 2.0
        cap = table.poll_capacity((int(2.0 * blur_px / tile) + 3) ** 2 if blur_px > 0 else 8)
        # The config record, the workspace sizes and the layout of the internal buffer depend only on static things:
        # memoised on the view table (one render call used to spend ~25 us of host time rebuilding them).
        plan_key = (shader, spec["light_kind"], tex_map is not None, H, W, K, spec["sigma"], spec["gamma"],
                    spec["background"], spec["blur_radius"], spec["flags"], spec["perspective"],
                    spec["camera_center_from_rt"], want_light_grad, spec.get("z_clip", 0.0) or 0.0, verts.shape[0],
                    0 if faces is None else faces.shape[0], sparse, cap)
        plan = table._plans.get(plan_key)
        if plan is None:
            cfg = _lib.RenderConfig()
            sc = cfg.shade
            sc.N, sc.H, sc.W, sc.K = N, H, W, K
            sc.shader, sc.light_kind = shader, spec["light_kind"]
            sc.texture_mode = _lib.TEX_UV if tex_map is not None else _lib.TEX_VERTEX
            sc.sigma, sc.gamma = spec["sigma"], spec["gamma"]
            sc.background[0], sc.background[1], sc.background[2] = spec["background"]
            cfg.blur_radius, cfg.raster_flags = spec["blur_radius"], spec["flags"]
            cfg.perspective = int(spec["perspective"])
            cfg.max_face_count, cfg.max_vert_count = table.max_face_count, table.max_vert_count
            cfg.camera_center_from_rt = int(spec["camera_center_from_rt"])
            # light / explicit camera-centre gradients are only reduced when the parameter block wants them
            cfg.want_light_grad = want_light_grad
            cfg.z_clip_value = float(spec.get("z_clip", 0.0) or 0.0)
            # the caller returns the image only: Fragments are written for covered pixels only (include/trb.h)
            cfg.sparse_fragments = sparse
            cfg.num_world_verts, cfg.num_faces = verts.shape[0], (0 if faces is None else faces.shape[0])
            cfg.num_ndc_verts = table.total_ndc_verts
            cfg.pair_capacity = cap
            cfg.scratch_is_zeroed = 1     # the backward zero-fills scratch with the gradient outputs (one torch.zeros)
            ws_bytes, n_tiles, n_scratch = ctypes.c_size_t(0), ctypes.c_int64(0), ctypes.c_int64(0)
            check(L.trb_render_sizes(ctypes.byref(cfg), ctypes.byref(ws_bytes), ctypes.byref(n_tiles),
                                     ctypes.byref(n_scratch)), "render")
            # One allocation for everything the caller never sees -- workspace, NDC vertices, vertex normals, the
            # covered-pixel list, bin statistics, and (sparse mode) the Fragments themselves -- addressed by offset
            # (each torch.empty costs ~4 us of host time).
            V3 = verts.shape[0] * 12
            P = N * H * W * K
            sizes = (ws_bytes.value, table.total_ndc_verts * 12, V3 if phong else 0, V3 if phong else 0,
                     max(n_tiles.value, 1) * 4, 16,
                     8 * P if sparse else 0, 4 * P if sparse else 0, 12 * P if sparse else 0, 4 * P if sparse else 0,
                     N * _lib.VIEW_PARAM_STRIDE * 4)
            offs, total = [], 0
            for nb in sizes:
                offs.append(total)
                total += (nb + 255) & ~255
            plan = table._plans[plan_key] = (cfg, ws_bytes.value, n_scratch.value, tuple(offs), max(total, 256),
                                             max(n_tiles.value, 1))
            if len(table._plans) > 64:      # settings that change every call must not grow the memo without bound
                table._plans.pop(next(iter(table._plans)))
        cfg, ws_nbytes, n_scratch, offs, total, n_hit_words = plan
        uv = None
        if tex_map is not None:
            verts_uvs, faces_uvs = spec["uv"]
            uv = _lib.UvTexture(tex_map.data_ptr(), verts_uvs.data_ptr(), faces_uvs.data_ptr(), 0,
                                tex_map.shape[0], tex_map.shape[1])
        want_stats = table.want_stats()
        aux = torch.empty((total,), dtype=torch.uint8, device=dev)
        base = aux.data_ptr()
        p_ws, p_ndc, p_nraw, p_nrm, p_hit, p_stats = (base + o for o in offs[:6])
        if not phong:
            p_nraw = p_nrm = 0
        if not want_stats:
            p_stats = 0
        if sparse:
            # internal Fragments (covered pixels only): views of the same buffer, never shown to the caller
            p2f = zbuf = bary = dists = None
            p_p2f, p_zbuf, p_bary, p_dists = (base + o for o in offs[6:10])
        else:
            p2f = torch.empty((N, H, W, K), dtype=torch.int64, device=dev)
            zbuf = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
            bary = torch.empty((N, H, W, K, 3), dtype=torch.float32, device=dev)
            dists = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
            p_p2f, p_zbuf, p_bary, p_dists = p2f.data_ptr(), zbuf.data_ptr(), bary.data_ptr(), dists.data_ptr()
        images = (torch.empty((N, H, W, 4), dtype=torch.float32, device=dev) if shader != _lib.SHADER_NONE
                  else _empty_f32(dev))
        vp = None
        extras = _lib.RenderExtras(0, 0, 0)
        if vp_src is not None:
            vp = aux[offs[10]:offs[10] + N * _lib.VIEW_PARAM_STRIDE * 4].view(torch.float32).view(N, _lib.VIEW_PARAM_STRIDE)
            extras.view_params_src = vp_src.data_ptr()
        # The backward's one zero-filled allocation (gradient outputs + the kernels' float4 accumulators) is made
        # here when a backward can follow, and zeroed by the forward's first kernel: no fill node in front of the
        # backward.  `backward` takes it once; a second backward over the same graph allocates its own.
        ctx.flat = None
        if N > 0 and any(ctx.needs_input_grad[:7]):
            ctx.flat = torch.empty((sum(_RenderFn._grad_sizes(ctx.needs_input_grad, n_scratch, N, verts.shape[0],
                                                              tex_map)),), dtype=torch.float32, device=dev)
            extras.zero_buffer, extras.zero_count = ctx.flat.data_ptr(), ctx.flat.numel()
        with _timed("render_forward", dev):
            check(L.trb_render_forward(
                ctypes.byref(cfg), _ptr(table.views), _ptr(verts), _ptr(faces), _ptr(colors), _ptr(R), _ptr(T),
                _ptr(proj), _ptr(vp), p_ndc, p_nraw, p_nrm, p_p2f, p_zbuf,
                p_bary, p_dists, _ptr(images if shader != _lib.SHADER_NONE else None), p_hit,
                p_ws, ws_nbytes, p_stats, None if uv is None else ctypes.byref(uv), ctypes.byref(extras), dev.index,
                _stream(dev)), "render")
        _bump(5 + (1 if want_stats else 0))  # prep, count, alloc, fill, fine (+ stats)
        if want_stats:
            table.arm_stats(aux[offs[5]:offs[5] + 16].view(torch.int32), dev)
        ctx.save_for_backward(verts, colors, R, T, proj, vp, faces, aux, p2f, zbuf, bary, dists, tex_map,
                              *(spec["uv"] if tex_map is not None else ()))
        ctx.aux_offsets = (offs[1], offs[2], offs[3], offs[4], phong)
        ctx.frag_offsets = tuple(offs[6:10]) if sparse else None
        ctx.table, ctx.cfg, ctx.n_scratch = table, cfg, n_scratch
        ctx.token = spec.get("_token")   # Fragments cache: set once a backward has consumed this graph
        ctx.set_materialize_grads(False)
        # per-view sums of the alpha channel, accumulated by the fine kernel behind the covered-pixel list
        alpha_sum = None
        if shader != _lib.SHADER_NONE:
            o_alpha = offs[4] + (n_hit_words - N) * 4
            alpha_sum = aux[o_alpha:o_alpha + 4 * N].view(torch.float32)
        # (one call: a second mark_non_differentiable replaces the first)
        ctx.mark_non_differentiable(*[t for t in (p2f, alpha_sum) if t is not None])
        if sparse:
            return images, None, None, None, None, alpha_sum
        return images, p2f, zbuf, bary, dists, alpha_sum

    @staticmethod
    def _grad_sizes(needs_input_grad, n_scratch, N, V, tex_map):
        """Float counts of the parts of the backward's single allocation: the kernels' scratch (first: 16-byte
        aligned), then grad verts, colours, R, T, projection, view parameters, texture map."""
        n_tex = tex_map.numel() if (tex_map is not None and needs_input_grad[2]) else 0
        return [(max(n_scratch, 1) + 3) // 4 * 4, V * 3, V * 3, N * 9, N * 3, N * 4, N * _lib.VIEW_PARAM_STRIDE, n_tex]

    @staticmethod
    def backward(ctx, g_images, _g_p2f, g_zbuf, g_bary, g_dists, _g_alpha_sum=None):
        verts, colors, R, T, proj, vp, faces, aux, p2f, zbuf, bary, dists, tex_map = ctx.saved_tensors[:13]
        o_ndc, o_nraw, o_nrm, o_hit, phong = ctx.aux_offsets
        base = aux.data_ptr()
        p_ndc, p_hit = base + o_ndc, base + o_hit
        p_nraw, p_nrm = (base + o_nraw, base + o_nrm) if phong else (0, 0)
        if ctx.frag_offsets is not None:
            p_p2f, p_zbuf, p_bary, p_dists = (base + o for o in ctx.frag_offsets)
        else:
            p_p2f, p_zbuf, p_bary, p_dists = p2f.data_ptr(), zbuf.data_ptr(), bary.data_ptr(), dists.data_ptr()
        table, cfg = ctx.table, ctx.cfg
        if ctx.token is not None:
            ctx.token["consumed"] = True
        dev = verts.device
        need = list(ctx.needs_input_grad)  # verts, colors, tex_map, R, T, proj, view_params
        need_tex = need.pop(2)
        N, V = table.N, verts.shape[0]
        shader = cfg.shade.shader
        if shader != _lib.SHADER_NONE and g_images is None:
            g_images = torch.zeros((N, cfg.shade.H, cfg.shade.W, 4), dtype=torch.float32, device=dev)
        # one zero-filled buffer for every accumulated gradient and for the kernels' float4 accumulators: the one
        # the forward allocated and its first kernel zeroed, or (second backward over the same graph) a fresh one
        sizes = _RenderFn._grad_sizes(ctx.needs_input_grad, ctx.n_scratch, N, V, tex_map)
        n_tex = sizes[-1]
        flat, ctx.flat = ctx.flat, None
        if flat is None or flat.numel() != sum(sizes):
            flat = torch.zeros((sum(sizes),), dtype=torch.float32, device=dev)
        parts = list(flat.split(sizes))
        scratch, g_verts, g_cols, g_R, g_T, g_proj, g_vp, g_tex = parts
        uv = None
        if tex_map is not None:
            verts_uvs, faces_uvs = ctx.saved_tensors[13:15]
            uv = _lib.UvTexture(tex_map.data_ptr(), verts_uvs.data_ptr(), faces_uvs.data_ptr(),
                                g_tex.data_ptr() if n_tex else 0, tex_map.shape[0], tex_map.shape[1])
        f32 = lambda t: None if t is None else _f32c(t)
        want_vp = vp is not None
        want_cols = need[1] and colors is not None
        args = (ctypes.byref(cfg), _ptr(table.views), _ptr(verts), _ptr(faces), _ptr(colors), _ptr(R), _ptr(T),
                _ptr(proj), _ptr(vp), p_ndc, p_nraw, p_nrm, p_p2f, p_zbuf,
                p_bary, p_dists, p_hit, _ptr(f32(g_images) if shader != _lib.SHADER_NONE else None),
                _ptr(f32(g_zbuf)), _ptr(f32(g_bary)), _ptr(f32(g_dists)),
                _ptr(g_verts if need[0] else None), _ptr(g_cols if want_cols else None),
                _ptr(g_R if need[2] else None), _ptr(g_T if need[3] else None), _ptr(g_proj if need[4] else None),
                _ptr(g_vp if want_vp else None), _ptr(scratch), None if uv is None else ctypes.byref(uv))
        # multi-GPU (parallel.fused_backward_allreduce): the sum of the view-shared gradients over the ranks is part
        # of this backward -- pushed to the peers by the tail kernel, summed by a short receive kernel
        peer = _backward_peer_sum
        fused = peer is not None and (need[0] or want_cols)
        with _timed("render_backward", dev):
            if fused:
                n_shared = V * 3 * (int(bool(need[0])) + int(bool(want_cols) and shader in (
                    _lib.SHADER_SOFT_PHONG, _lib.SHADER_HARD_PHONG) and tex_map is None))
                check(_lib.lib().trb_render_backward_allreduce(*args, ctypes.byref(peer.peer_sum(n_shared, dev)),
                                                               dev.index, _stream(dev)), "render backward")
                peer.count_call()
            else:
                check(_lib.lib().trb_render_backward(*args, dev.index, _stream(dev)), "render backward")
        _bump(3 if fused else 2)  # fused backward, post (+ receive)
        return (g_verts.view(V, 3) if need[0] else None,
                g_cols.view(V, 3) if (need[1] and colors is not None) else None,
                g_tex.view(tex_map.shape) if n_tex else None,
                g_R.view(N, 3, 3) if need[2] else None, g_T.view(N, 3) if need[3] else None,
                g_proj.view(N, 4) if need[4] else None,
                g_vp.view(N, _lib.VIEW_PARAM_STRIDE) if (need[5] and want_vp) else None, None, None, None)


# The PeerAllReduce (parallel.py) the next fused backwards push their shared gradients through; None = off.  A plain
# module global: autograd runs backward on its own thread.
_backward_peer_sum = None


def set_backward_peer_sum(peer) -> None:
    global _backward_peer_sum
    _backward_peer_sum = peer


def render(verts, colors, R, T, proj, view_params, faces, table: ViewTable, spec: dict, tex_map=None):
    """Returns (images [N,H,W,4] or empty, pix_to_face, zbuf, bary, dists).  With ``tex_map`` (f32 [Ht,Wt,3])
    the texture is a UV map sampled in the kernels; ``spec["uv"]`` = (verts_uvs f32 [Vt,2], faces_uvs i32 [F,3])."""
    if spec["K"] > _lib.MAX_FACES_PER_PIXEL:
        raise ValueError(f"faces_per_pixel must be <= {_lib.MAX_FACES_PER_PIXEL}")
    images, p2f, zbuf, bary, dists, alpha_sum = _RenderFn.apply(verts, colors, tex_map, R, T, proj, view_params, faces,
                                                                table, spec)
    if alpha_sum is not None:
        # f32 [N]: images[n, :, :, 3].sum(), accumulated by the kernel that wrote the pixels (no second pass over the
        # image; float atomics, so the last bits vary from run to run).  Not differentiable: a statistic, not a loss.
        images.alpha_sum = alpha_sum
    return images, p2f, zbuf, bary, dists


# --------------------------------------------------------------------------------------------
# Point clouds (csrc/points_render.cu)
class _RasterizePointsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points_ndc, radius, table: ViewTable, H, W, K, tiles_per_point):
        _require_cuda(points_ndc, "rasterize_points")
        points_ndc, radius = _f32c(points_ndc), _f32c(radius)
        dev = points_ndc.device
        N = table.N
        L = _lib.lib()
        idx = torch.empty((N, H, W, K), dtype=torch.int32, device=dev)
        zbuf = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
        dists = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
        # per-tile point lists: every point enters the tiles its disc can reach (a known scalar radius bounds that
        # number; per-point radii get 8 entries per point -- tiles that do not fit scan the whole cloud instead)
        total_points = points_ndc.shape[0]
        # (capped at 64 M entries = 256 MB: huge discs then overflow into the whole-cloud scan, which is what they need)
        capacity = int(min(total_points * max(int(tiles_per_point), 1) + 1024, 1 << 26))
        nbytes = ctypes.c_size_t(0)
        check(L.trb_points_raster_workspace_bytes(N, H, W, capacity, ctypes.byref(nbytes)), "rasterize_points")
        ws = torch.empty((nbytes.value,), dtype=torch.uint8, device=dev)
        with _timed("points_raster_forward", dev):
            check(L.trb_points_raster_forward_binned(_ptr(points_ndc), _ptr(radius), _ptr(table.views), N,
                                                     table.max_face_count, H, W, K, capacity, _ptr(ws), nbytes.value,
                                                     _ptr(idx), _ptr(zbuf), _ptr(dists), dev.index, _stream(dev)),
                  "rasterize_points")
        _bump(5)
        ctx.save_for_backward(points_ndc, idx)
        ctx.dims = (N, H, W, K)
        ctx.mark_non_differentiable(idx)
        ctx.set_materialize_grads(False)
        return idx, zbuf, dists

    @staticmethod
    def backward(ctx, _g_idx, g_zbuf, g_dists):
        points_ndc, idx = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return (None,) * 7
        N, H, W, K = ctx.dims
        dev = points_ndc.device
        g_points = torch.zeros_like(points_ndc)
        if (g_zbuf is None and g_dists is None) or points_ndc.numel() == 0:
            return (g_points,) + (None,) * 6
        g_zbuf = None if g_zbuf is None else _f32c(g_zbuf)
        g_dists = None if g_dists is None else _f32c(g_dists)
        with _timed("points_raster_backward", dev):
            check(_lib.lib().trb_points_raster_backward(_ptr(points_ndc), _ptr(idx), _ptr(g_zbuf), _ptr(g_dists), N, H,
                                                        W, K, _ptr(g_points), dev.index, _stream(dev)),
                  "rasterize_points backward")
        _bump(1)
        return (g_points,) + (None,) * 6


def rasterize_points_ndc(points_ndc, radius, table: ViewTable, image_size, points_per_pixel, max_radius=None):
    """points_ndc f32 (P, 3) packed, radius f32 (P,), one view per cloud -> (idx i32 (N,H,W,K), zbuf, dists).
    ``max_radius`` (a host float, when the settings hold a scalar radius) sizes the per-tile lists exactly."""
    H, W = image_size
    if points_per_pixel > _lib.MAX_FACES_PER_PIXEL:
        raise ValueError(f"Must have points_per_pixel <= {_lib.MAX_FACES_PER_PIXEL}")
    tiles_per_point = 8
    if max_radius is not None and max_radius == max_radius and max_radius >= 0:
        r_px = float(max_radius) * 1.0001 * min(H, W) / 2.0 + 0.6      # NDC -> pixels (the short side spans 2)
        tiles_per_point = min((int(2.0 * r_px / 16.0) + 2) ** 2, 4096)
    return _RasterizePointsFn.apply(points_ndc, radius, table, int(H), int(W), int(points_per_pixel), tiles_per_point)


COMPOSITE_ALPHA, COMPOSITE_NORM_WEIGHTED = 0, 1


class _CompositeFn(torch.autograd.Function):
    """idx i32 (N,H,W,K), alphas f32 (N,H,W,K), features f32 (P,C) -> images f32 (N,H,W,C)."""

    @staticmethod
    def forward(ctx, idx, alphas, features, background, mode: int):
        _require_cuda(alphas, "compositor")
        idx = idx.to(torch.int32).contiguous()
        alphas, features = _f32c(alphas), _f32c(features)
        background = None if background is None else _f32c(background)
        dev = alphas.device
        N, H, W, K = alphas.shape
        C = features.shape[1]
        images = torch.empty((N, H, W, C), dtype=torch.float32, device=dev)
        with _timed("points_composite_forward", dev):
            check(_lib.lib().trb_points_composite_forward(mode, _ptr(idx), _ptr(alphas), _ptr(features), N * H * W, K,
                                                          C, _ptr(background), _ptr(images), dev.index, _stream(dev)),
                  "compositor")
        _bump(1)
        ctx.save_for_backward(idx, alphas, features)
        ctx.mode, ctx.has_background = mode, background is not None
        return images

    @staticmethod
    def backward(ctx, g_images):
        idx, alphas, features = ctx.saved_tensors
        dev = alphas.device
        N, H, W, K = alphas.shape
        C = features.shape[1]
        need = ctx.needs_input_grad
        g_alphas = torch.empty_like(alphas) if need[1] else None
        g_features = torch.zeros_like(features) if need[2] else None
        with _timed("points_composite_backward", dev):
            check(_lib.lib().trb_points_composite_backward(
                ctx.mode, _ptr(idx), _ptr(alphas), _ptr(features), _ptr(_f32c(g_images)), N * H * W, K, C,
                int(ctx.has_background), _ptr(g_alphas), _ptr(g_features), dev.index, _stream(dev)),
                "compositor backward")
        _bump(1)
        return None, g_alphas, g_features, None, None


def composite(idx, alphas, features, mode: int, background=None):
    if idx.shape != alphas.shape or alphas.dim() != 4:
        raise ValueError("idx and alphas must both have shape (N, H, W, K)")
    if features.dim() != 2:
        raise ValueError("features must have shape (P, C)")
    if alphas.shape[-1] > _lib.MAX_FACES_PER_PIXEL:
        raise ValueError(f"points_per_pixel must be <= {_lib.MAX_FACES_PER_PIXEL}")
    return _CompositeFn.apply(idx, alphas, features, background, mode)
