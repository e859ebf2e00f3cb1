"""``RasterizationSettings``, ``Fragments``, ``MeshRasterizer`` and ``rasterize_meshes`` with the
PyTorch3D call surface.  Reference usage:
``MeshRasterizer(cameras=..., raster_settings=...)(meshes, R=Rs, T=ts).zbuf[..., 0]``
(torch_renderer.py:97-113, camera_pose_optimizer.py:139-142,244, batch_rendering_test.py:196-274,
myrenderer.py:84-103).

Host side only: resolves settings, turns the camera into ``(R, T, fx, fy, px, py)`` and launches the
CUDA transform + rasteriser through ``ops``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import NamedTuple, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from . import _lib, clip, ops
from .common import named_tensors
from .structures import Meshes

kMaxBinsPerDim = 22  # PyTorch3D refuses bin grids this large; kept for error parity


class Fragments(NamedTuple):
    """pix_to_face i64 (N,H,W,K) packed face index or -1; zbuf, dists f32 (N,H,W,K); bary_coords
    f32 (N,H,W,K,3); background is -1 in all four."""
    pix_to_face: torch.Tensor
    zbuf: torch.Tensor
    bary_coords: torch.Tensor
    dists: Optional[torch.Tensor]

    def detach(self) -> "Fragments":
        return Fragments(self.pix_to_face, self.zbuf.detach(), self.bary_coords.detach(),
                         None if self.dists is None else self.dists.detach())


@dataclass
class RasterizationSettings:
    image_size: Union[int, Tuple[int, int]] = 256
    blur_radius: float = 0.0
    faces_per_pixel: int = 1
    bin_size: Optional[int] = None
    max_faces_per_bin: Optional[int] = None
    perspective_correct: Optional[bool] = None
    clip_barycentric_coords: Optional[bool] = None
    cull_backfaces: bool = False
    z_clip_value: Optional[float] = None
    cull_to_frustum: bool = False


def _parse_image_size(image_size) -> Tuple[int, int]:
    if isinstance(image_size, int):
        if image_size < 1:
            raise ValueError("image size must be positive")
        return image_size, image_size
    if torch.is_tensor(image_size):
        image_size = [int(v) for v in image_size.flatten().tolist()]
    if not isinstance(image_size, (tuple, list)):
        raise ValueError("Image size can only be a tuple/list of (H, W)")
    if len(image_size) != 2:
        raise ValueError("Image size can only be a tuple/list of (H, W)")
    if not all(isinstance(i, int) for i in image_size):
        raise ValueError("Image sizes must be integers; got %r" % (image_size,))
    if not all(i > 0 for i in image_size):
        raise ValueError("Image sizes must be greater than 0; got %d, %d" % tuple(image_size))
    return int(image_size[0]), int(image_size[1])


def _check_bin_size(bin_size, H: int, W: int) -> None:
    """Our tiling is fixed by the kernel, but invalid requests fail exactly like upstream."""
    if bin_size is None or bin_size == 0:
        return
    if bin_size < 0:
        raise ValueError("bin_size must be >= 0")
    bins = 1 + (max(H, W) - 1) // bin_size
    if bins >= kMaxBinsPerDim:
        raise ValueError("bin_size too small, number of bins per dimension must be less than %d; got %d"
                         % (kMaxBinsPerDim, bins))


def rasterize_meshes(meshes: Meshes, image_size=256, blur_radius: float = 0.0, faces_per_pixel: int = 8,
                     bin_size: Optional[int] = None, max_faces_per_bin: Optional[int] = None,
                     perspective_correct: bool = False, clip_barycentric_coords: bool = False,
                     cull_backfaces: bool = False, z_clip_value: Optional[float] = None,
                     cull_to_frustum: bool = False):
    """PyTorch3D ``rasterize_meshes`` twin: ``meshes`` hold vertices already in NDC (x, y) + view z.
    Returns (pix_to_face, zbuf, bary_coords, dists)."""
    H, W = _parse_image_size(image_size)
    _check_bin_size(bin_size, H, W)
    table = meshes.view_table()
    verts = meshes._unique_verts()
    if table.shared_mesh:
        # every replica has the same NDC vertices: expand (the kernels index per view)
        verts = verts.repeat(table.N, 1)
    if z_clip_value is not None or cull_to_frustum:
        clipped = _rasterize_clipped(verts, meshes.faces_packed_i32(), table, (H, W), blur_radius, faces_per_pixel,
                                     perspective_correct, clip_barycentric_coords, cull_backfaces, z_clip_value,
                                     cull_to_frustum)
        if clipped is not None:
            return clipped
    return ops.rasterize(verts, meshes.faces_packed_i32(), table, (H, W), blur_radius, faces_per_pixel,
                         perspective_correct, clip_barycentric_coords, cull_backfaces)


_clipping_mode = "exact"


def set_near_plane_clipping(mode: str = "exact") -> None:
    """How ``MeshRasterizer`` treats faces that cross the near clipping plane ``z_clip_value`` (``znear / 2`` for
    cameras that define ``znear``, as upstream) or, with ``cull_to_frustum``, leave the view frustum.

    ``"exact"`` (default): PyTorch3D's behaviour.  Every rasterisation with an active plane asks the device whether
    any vertex lies behind it (one small kernel and a 4-byte read -- upstream's ``clip_faces`` reads two sums the
    same way), enqueues the fused render, and reads the answer afterwards, so the GPU is never left waiting for the
    host; only on "yes" is that render dropped and the faces cut (``clip.clip_faces``) and drawn by the stand-alone
    rasteriser.  The question cannot be asked while a CUDA graph is being captured: captured renders behave
    like ``"off"``.

    ``"off"``: never ask.  Faces entirely behind the plane are still culled inside the kernels; a face crossing it
    is drawn whole (or dropped when a vertex is at / behind the camera plane).  For loops that are known to keep
    the camera away from the mesh and must not synchronise."""
    global _clipping_mode
    if mode not in ("exact", "off"):
        raise ValueError("mode must be 'exact' or 'off'")
    _clipping_mode = mode


def _face_vertex_rows(faces_i32: torch.Tensor, table) -> torch.Tensor:
    """i64 (F_packed, 3): rows of the view-major NDC vertex array for every packed face of the batch."""
    faces = faces_i32.long()
    if not table.shared_mesh:
        return faces
    V = table.max_vert_count
    shift = torch.arange(table.N, device=faces.device, dtype=torch.long) * V
    return (faces[None] + shift[:, None, None]).reshape(-1, 3)


def _rasterize_clipped(verts_ndc, faces_i32, table, image_size, blur_radius, K, perspective_correct,
                       clip_barycentric_coords, cull_backfaces, z_clip_value, cull_to_frustum):
    """``rasterize_meshes`` with ``clip_faces`` in front, as upstream: (pix_to_face, zbuf, bary, dists) indexed by
    the faces of the batch, or ``None`` when ``clip_faces`` leaves the face list untouched."""
    host = table.host
    dev = verts_ndc.device
    first = host[:, 3].long().to(dev)
    count = host[:, 1].long().to(dev)
    face_verts = verts_ndc[_face_vertex_rows(faces_i32, table)]
    frustum = clip.rasterizer_frustum(perspective_correct, z_clip_value, cull_to_frustum)
    cf = clip.clip_faces(face_verts, first, count, frustum)
    if cf.faces_clipped_to_unclipped_idx is None:
        return None
    p2f, zbuf, bary, dists = ops.rasterize_face_verts(
        cf.face_verts, cf.mesh_to_face_first_idx, cf.num_faces_per_mesh, image_size, blur_radius, K,
        perspective_correct, clip_barycentric_coords, cull_backfaces, cf.clipped_faces_neighbor_idx)
    p2f, bary = clip.convert_clipped_rasterization_to_original_faces(p2f, bary, cf)
    return p2f, zbuf, bary, dists


def _cached_projection(cameras, proj_kwargs):
    """``cameras.ndc_projection_params`` memoised on the camera object while its intrinsics are
    untouched constants (keyed on tensor identity + version) -- a handful of tiny launches saved per call."""
    if proj_kwargs:
        return cameras.ndc_projection_params(**proj_kwargs)
    tensors = [(k, v) for k, v in named_tensors(cameras).items() if k not in ("R", "T")]
    if any(v.requires_grad for _, v in tensors):
        return cameras.ndc_projection_params()
    key = tuple((k, id(v), v._version) for k, v in tensors)
    cache = cameras.__dict__.get("_trb_proj_cache")
    if cache is None or cache[0] != key:
        cache = (key, cameras.ndc_projection_params())
        cameras.__dict__["_trb_proj_cache"] = cache
    return cache[1]


def _cached_half_znear(cameras):
    """znear / 2 of a camera that defines znear (FoV cameras), else None; the host read of the tensor is
    done once per (tensor, version)."""
    znear = cameras.get_znear()
    if znear is None:
        return None
    if not torch.is_tensor(znear):
        return float(znear) / 2.0
    key = (id(znear), znear._version)
    cache = cameras.__dict__.get("_trb_znear_cache")
    if cache is None or cache[0] != key:
        cache = (key, float(znear.min()) / 2.0)
        cameras.__dict__["_trb_znear_cache"] = cache
    return cache[1]


class _FragmentCache:
    """The Fragments of the most recent rasterisation, keyed on the identity + version of every input tensor
    and on the raster settings.  The reference rasterises the same scene two or three times
    per step -- ``rasterizer(meshes, R=R, T=T).zbuf``, then the silhouette renderer, then the Phong renderer
    with the same meshes, R, T and settings (camera_pose_optimizer.py:244-250, torch_renderer.py:113-120,
    deform_mesh_with_color.py:373-381); with the cache the second and third call only run the shading kernel
    on the stored Fragments, and the backward rasterises once (autograd sums the three gradient streams).

    One entry; entries above ``max_bytes`` are not kept (a chunk of a 1024-view job must not stay alive past
    its iteration); nothing is stored while a CUDA graph is being captured.

    OPT-IN (``set_fragment_cache(True)``).  The key sees tensor identity, ``_version`` and ``requires_grad``;
    writes that do not bump ``_version`` -- through ``.data``, a numpy view of a CPU tensor, a raw pointer in a
    custom kernel -- are invisible to it and would return the previous Fragments.  A loop that updates its inputs
    only through autograd-visible in-place ops (every torch optimiser) or by building new tensors is safe."""

    def __init__(self):
        self.enabled = False
        self.max_bytes = 2 << 30
        self.key = None
        self.refs = ()
        self.value = None
        self.token = None
        self.hits = 0

    def clear(self):
        self.key, self.refs, self.value, self.token = None, (), None, None

    @staticmethod
    def make_key(tensors, spec, table):
        ids = tuple((id(t), t._version, t.requires_grad) for t in tensors)
        raster = (spec["image_size"], spec["K"], spec["blur_radius"], spec["flags"], spec["z_clip"],
                  spec["perspective"], spec["cull_to_frustum"])
        return (ids, raster, id(table), torch.is_grad_enabled(), _clipping_mode)

    def lookup(self, key, tensors):
        if not self.enabled or self.key != key or self.value is None:
            return None
        # a backward pass has already run through the stored Fragments: their autograd graph is spent
        if self.token is not None and self.token.get("consumed"):
            return None
        # ids are only unique among live objects: every input must still be the tensor the entry was made from
        if len(self.refs) != len(tensors) or any(r() is not t for r, t in zip(self.refs, tensors)):
            return None
        self.hits += 1
        return self.value

    def store(self, key, tensors, fragments, token=None):
        if not self.enabled or torch.cuda.is_current_stream_capturing():
            return
        nbytes = sum(t.numel() * t.element_size() for t in fragments if t is not None)
        if nbytes > self.max_bytes:
            self.clear()
            return
        import weakref
        self.key, self.refs, self.value = key, tuple(weakref.ref(t) for t in tensors), fragments
        self.token = token


_fragment_cache = _FragmentCache()


def set_fragment_cache(enabled: bool = True, max_bytes: Optional[int] = None) -> None:
    """Switches the reuse of Fragments between back-to-back renders of identical inputs on or off (default: off;
    see ``_FragmentCache`` for what "identical" can and cannot see)."""
    _fragment_cache.enabled = bool(enabled)
    if max_bytes is not None:
        _fragment_cache.max_bytes = int(max_bytes)
    _fragment_cache.clear()


def _expand_views(t: torch.Tensor, N: int, what: str) -> torch.Tensor:
    if t.shape[0] == N:
        return t
    if t.shape[0] == 1:
        return t.expand((N,) + tuple(t.shape[1:]))
    raise ValueError(f"Wrong number of cameras: {what} has batch {t.shape[0]} but the mesh batch has {N} "
                     "(must be 1 or equal)")


class MeshRasterizer(nn.Module):
    def __init__(self, cameras=None, raster_settings: Optional[RasterizationSettings] = None) -> None:
        super().__init__()
        if raster_settings is None:
            raster_settings = RasterizationSettings()
        self.cameras = cameras
        self.raster_settings = raster_settings

    def to(self, device):
        if self.cameras is not None:
            self.cameras = self.cameras.to(device)
        return self

    def _camera_inputs(self, meshes_world: Meshes, kwargs):
        cameras = kwargs.get("cameras", self.cameras)
        if cameras is None:
            raise ValueError("Cameras must be specified either at initialization or in the forward pass "
                             "of MeshRasterizer")
        N = len(meshes_world)
        dev = meshes_world.device
        R = kwargs.get("R", None)
        T = kwargs.get("T", None)
        R = cameras.R if R is None else R
        T = cameras.T if T is None else T
        # PyTorch3D's get_world_to_view_transform stores per-call overrides on the camera object;
        # the shaders' specular term (get_camera_center() without kwargs) relies on it.
        cameras.__dict__["R"], cameras.__dict__["T"] = R, T   # plain attributes (nn.Module.__setattr__ is slow)
        proj_kwargs = {k: v for k, v in kwargs.items()
                       if k not in ("R", "T", "cameras", "lights", "materials", "blend_params", "raster_settings")}
        proj, perspective = _cached_projection(cameras, proj_kwargs)
        self.__dict__["_identity"] = (R, T, proj)   # the caller's tensor objects (expand() below makes new ones per call)
        R = _expand_views(R.to(dev), N, "R")
        T = _expand_views(T.to(dev), N, "T")
        proj = _expand_views(proj.to(dev), N, "projection")
        return cameras, R, T, proj, perspective

    def _resolve(self, meshes_world: Meshes, kwargs):
        """Everything the kernels need for this call: camera tensors + the static raster spec."""
        raster_settings = kwargs.get("raster_settings", self.raster_settings)
        H, W = _parse_image_size(raster_settings.image_size)
        _check_bin_size(raster_settings.bin_size, H, W)
        cameras, R, T, proj, perspective = self._camera_inputs(meshes_world, kwargs)
        clip_bary = raster_settings.clip_barycentric_coords
        if clip_bary is None:
            clip_bary = raster_settings.blur_radius > 0.0
        persp_correct = raster_settings.perspective_correct
        if persp_correct is None:
            persp_correct = cameras.is_perspective()
        # z_clip_value: like upstream, a perspective-correct render with a camera that defines znear clips at
        # znear / 2 unless the settings say otherwise.  Faces entirely nearer than the plane are culled in the
        # kernels; when a vertex lies behind the plane the call goes through clip_faces (``_clipped_fragments``).
        z_clip = raster_settings.z_clip_value
        if z_clip is None and persp_correct:
            z_clip = _cached_half_znear(cameras)
        K = int(raster_settings.faces_per_pixel)
        if K > _lib.MAX_FACES_PER_PIXEL:
            raise ValueError(f"faces_per_pixel must be <= {_lib.MAX_FACES_PER_PIXEL}")
        flags = ((_lib.PERSPECTIVE_CORRECT if persp_correct else 0) | (_lib.CLIP_BARYCENTRIC if clip_bary else 0)
                 | (_lib.CULL_BACKFACES if raster_settings.cull_backfaces else 0))
        spec = dict(image_size=(H, W), K=K, blur_radius=float(raster_settings.blur_radius), flags=flags,
                    z_clip=float(z_clip or 0.0), z_clip_value=(None if z_clip is None else float(z_clip)),
                    cull_to_frustum=bool(raster_settings.cull_to_frustum),
                    perspective=bool(perspective), shader=_lib.SHADER_NONE, light_kind=0, sigma=1.0, gamma=1.0,
                    background=(0.0, 0.0, 0.0), camera_center_from_rt=False)
        return cameras, R, T, proj, spec

    def transform(self, meshes_world: Meshes, **kwargs) -> torch.Tensor:
        """World -> NDC (x, y) + view z for every (view, vertex): f32 [sum_n V_n, 3], view-major."""
        _, R, T, proj, perspective = self._camera_inputs(meshes_world, kwargs)
        return ops.transform_verts(meshes_world._unique_verts(), R, T, proj, meshes_world.view_table(),
                                   perspective)

    def _cache_key(self, meshes_world: Meshes, spec):
        """(key, tensors) of this call for the Fragments cache; valid right after ``_resolve``."""
        tensors = (meshes_world._unique_verts(), meshes_world.faces_packed_i32()) + tuple(self._identity)
        return _fragment_cache.make_key(tensors, spec, meshes_world.view_table()), tensors

    def _near_plane_question(self, meshes_world: Meshes, R, T, spec):
        """``None`` when this call has no clipping decision to make (no active plane, ``"off"``, graph capture);
        else a callable that blocks until the device has said whether any vertex lies behind the plane.  The
        question is ENQUEUED here; callers enqueue the fused render they expect to keep and only then read the
        answer, so the host read never leaves the GPU idle (``ops.any_vertex_behind_async``)."""
        z_clip, cull = spec["z_clip_value"], spec["cull_to_frustum"]
        if cull and _clipping_mode == "off":
            raise ValueError("cull_to_frustum needs set_near_plane_clipping('exact')")
        watch = ops.active_near_plane_watch()
        if watch is not None:
            # capture_step: the step is (about to be) replayed from a CUDA graph and cannot read an answer; the
            # test raises a sticky flag the CapturedStep looks at between replays (capture.py)
            if cull:
                raise ValueError("cull_to_frustum cuts faces on the host's say-so and cannot run inside capture_step")
            if z_clip is not None and _clipping_mode != "off":
                watch.enqueue(meshes_world._unique_verts(), R, T, meshes_world.view_table(), z_clip)
            return None
        if (z_clip is None and not cull) or _clipping_mode == "off" or torch.cuda.is_current_stream_capturing():
            return None
        if cull:
            return lambda: True      # clip_faces decides (it culls against x, y planes too)
        return ops.any_vertex_behind_async(meshes_world._unique_verts(), R, T, meshes_world.view_table(), z_clip)

    def _clipped_fragments(self, meshes_world: Meshes, R, T, proj, spec) -> Optional[Fragments]:
        """The upstream ``clip_faces`` route: transform, cut, draw with the stand-alone rasteriser, map back.
        ``None`` when ``clip_faces`` leaves the face list untouched."""
        table = meshes_world.view_table()
        verts_ndc = ops.transform_verts(meshes_world._unique_verts(), R, T, proj, table, spec["perspective"])
        flags = spec["flags"]
        out = _rasterize_clipped(verts_ndc, meshes_world.faces_packed_i32(), table, spec["image_size"],
                                 spec["blur_radius"], spec["K"], bool(flags & _lib.PERSPECTIVE_CORRECT),
                                 bool(flags & _lib.CLIP_BARYCENTRIC), bool(flags & _lib.CULL_BACKFACES),
                                 spec["z_clip_value"], spec["cull_to_frustum"])
        if out is None:
            return None
        return Fragments(pix_to_face=out[0], zbuf=out[1], bary_coords=out[2], dists=out[3])

    def forward(self, meshes_world: Meshes, **kwargs) -> Fragments:
        _, R, T, proj, spec = self._resolve(meshes_world, kwargs)
        key, tensors = self._cache_key(meshes_world, spec)
        cached = _fragment_cache.lookup(key, tensors)
        if cached is not None:
            return cached
        behind = self._near_plane_question(meshes_world, R, T, spec)
        if behind is not None and spec["cull_to_frustum"]:
            clipped = self._clipped_fragments(meshes_world, R, T, proj, spec)
            if clipped is not None:
                return clipped
            behind = None
        token = {"consumed": False}
        spec["_token"] = token
        _, p2f, zbuf, bary, dists = ops.render(meshes_world._unique_verts(), None, R, T, proj, None,
                                                  meshes_world.faces_packed_i32(), meshes_world.view_table(), spec)
        if behind is not None and behind():
            # a vertex behind the near plane: the render above is discarded, faces are cut by clip_faces
            clipped = self._clipped_fragments(meshes_world, R, T, proj, spec)
            if clipped is not None:
                return clipped
        fragments = Fragments(pix_to_face=p2f, zbuf=zbuf, bary_coords=bary, dists=dists)
        _fragment_cache.store(key, tensors, fragments, token)
        return fragments
