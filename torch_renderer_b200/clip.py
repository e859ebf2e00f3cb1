"""Near-plane clipping and frustum culling of ``face_verts`` before rasterisation: ``ClipFrustum``, ``ClippedFaces``,
``clip_faces`` and ``convert_clipped_rasterization_to_original_faces`` with the call surface of
``pytorch3d.renderer.mesh.clip``.  Upstream runs this on every render whose settings carry a
``z_clip_value`` -- which ``MeshRasterizer`` sets to ``znear / 2`` for the ``FoVPerspectiveCameras`` of
camera_pose_optimizer.py:105, mesh_deformer.py:119 and batch_rendering_test.py:225 -- or ``cull_to_frustum``.

Host side: device-agnostic torch tensor code, differentiable w.r.t. ``face_verts`` (the cut weights are constants,
as upstream detaches them).  One prefix sum places every surviving face; faces with one or two vertices behind the
plane are rebuilt from a "pivot" vertex (the lone vertex on its side of the plane) and the two cut points on the
edges leaving it.  Semantics and operator order: ``oracle/clip_ref.py``.

``MeshRasterizer`` only comes here when a vertex of the batch lies behind the plane (``rasterizer._needs_clipping``);
the cut faces are then drawn by the stand-alone rasteriser (``ops.rasterize_face_verts``) -- the fused render kernels
never see them.
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Tuple

import torch


class ClipFrustum:
    """Planes of the view frustum in the space of ``face_verts`` (NDC x, y; view-space z); ``None`` = no plane."""
    __slots__ = ("left", "right", "top", "bottom", "znear", "zfar", "perspective_correct", "cull", "z_clip_value")

    def __init__(self, left: Optional[float] = None, right: Optional[float] = None, top: Optional[float] = None,
                 bottom: Optional[float] = None, znear: Optional[float] = None, zfar: Optional[float] = None,
                 perspective_correct: bool = False, cull: bool = True, z_clip_value: Optional[float] = None) -> None:
        self.left, self.right, self.top, self.bottom = left, right, top, bottom
        self.znear, self.zfar = znear, zfar
        self.perspective_correct, self.cull, self.z_clip_value = perspective_correct, cull, z_clip_value


class ClippedFaces(NamedTuple):
    """``face_verts`` f32 (F', 3, 3) after clipping, its per-mesh ranges, and -- unless nothing changed -- the maps
    back: ``faces_clipped_to_unclipped_idx`` i64 (F',), ``barycentric_conversion`` f32 (T, 3, 3) with
    ``bary_original = M @ bary_clipped``, ``faces_clipped_to_conversion_idx`` i64 (F',) (-1 = face not cut) and
    ``clipped_faces_neighbor_idx`` i64 (F',): the other half of a face cut into a quadrilateral, else -1."""
    face_verts: torch.Tensor
    mesh_to_face_first_idx: torch.Tensor
    num_faces_per_mesh: torch.Tensor
    faces_clipped_to_unclipped_idx: Optional[torch.Tensor] = None
    barycentric_conversion: Optional[torch.Tensor] = None
    faces_clipped_to_conversion_idx: Optional[torch.Tensor] = None
    clipped_faces_neighbor_idx: Optional[torch.Tensor] = None


def rasterizer_frustum(perspective_correct: bool, z_clip_value: Optional[float], cull_to_frustum: bool) -> ClipFrustum:
    """The frustum ``rasterize_meshes`` clips against: the NDC square [-1, 1]^2 (culling only) and the z plane."""
    return ClipFrustum(left=-1.0, right=1.0, top=-1.0, bottom=1.0, perspective_correct=bool(perspective_correct),
                       cull=bool(cull_to_frustum), z_clip_value=z_clip_value)


def _faces_outside_frustum(face_verts: torch.Tensor, frustum: ClipFrustum) -> torch.Tensor:
    """bool (F,): all three vertices beyond one of the frustum planes."""
    gone = torch.zeros(face_verts.shape[0], dtype=torch.bool, device=face_verts.device)
    if not frustum.cull:
        return gone
    for value, axis, below in ((frustum.left, 0, True), (frustum.right, 0, False), (frustum.top, 1, True),
                               (frustum.bottom, 1, False), (frustum.znear, 2, True), (frustum.zfar, 2, False)):
        if value is None:
            continue
        col = face_verts[:, :, axis]
        gone |= ((col < value) if below else (col > value)).all(dim=1)
    return gone


def _cut_point(p1: torch.Tensor, p: torch.Tensor, c: float, perspective_correct: bool):
    """Where the segments p1 -> p (T, 3) cross depth ``c``: points (T, 3) and their constant weights on p (T,)."""
    w = ((p1[:, 2] - c) / (p1[:, 2] - p[:, 2])).detach()
    one_w = 1.0 - w
    q = p1 * one_w[:, None] + p * w[:, None]
    if perspective_correct:
        # x, y are NDC: interpolate them in view space (multiply by z), then project at the new depth c
        xy = (p1[:, :2] * p1[:, 2:3] * one_w[:, None] + p[:, :2] * p[:, 2:3] * w[:, None]) / c
        q = torch.cat([xy, q[:, 2:3]], dim=1)
    return q, w


def clip_faces(face_verts_unclipped: torch.Tensor, mesh_to_face_first_idx: torch.Tensor,
               num_faces_per_mesh: torch.Tensor, frustum: ClipFrustum) -> ClippedFaces:
    fv = face_verts_unclipped
    F, dev = fv.shape[0], fv.device
    zc = frustum.z_clip_value
    behind = (fv[:, :, 2] < zc) if zc is not None else torch.zeros((F, 3), dtype=torch.bool, device=dev)
    gone = _faces_outside_frustum(fv, frustum)
    # one host read decides everything (upstream reads two sums the same way)
    n_behind, n_gone = torch.stack([behind.sum(), gone.sum()]).tolist()
    if n_behind == 0 and n_gone == 0:
        return ClippedFaces(fv, mesh_to_face_first_idx, num_faces_per_mesh)

    nb = behind.sum(dim=1)
    gone = gone | (nb == 3)
    one_cut = (nb == 2) & ~gone          # lone vertex in front  -> 1 triangle
    two_cut = (nb == 1) & ~gone          # lone vertex behind    -> 2 triangles
    whole = (nb == 0) & ~gone
    emitted = whole.long() + one_cut.long() + 2 * two_cut.long()
    ends = emitted.cumsum(0)
    slot = ends - emitted                # first output row of every input face
    Fc = int(ends[-1]) if F > 0 else 0
    slot_ext = torch.cat([slot, ends[-1:]]) if F > 0 else torch.zeros(1, dtype=torch.long, device=dev)
    first_c = slot_ext[mesh_to_face_first_idx.clamp(max=F)]
    count_c = torch.cat([first_c[1:], first_c.new_full((1,), Fc)]) - first_c

    idx_whole = whole.nonzero(as_tuple=True)[0]
    idx_one = one_cut.nonzero(as_tuple=True)[0]
    idx_two = two_cut.nonzero(as_tuple=True)[0]
    rows = [slot[idx_whole]]
    tris = [fv[idx_whole]]
    origin = [idx_whole]
    if idx_one.numel() + idx_two.numel() == 0:
        out = fv.new_zeros((Fc, 3, 3)).index_copy(0, rows[0], tris[0])
        to_unclipped = torch.zeros(Fc, dtype=torch.long, device=dev).index_copy(0, rows[0], origin[0])
        return ClippedFaces(out, first_c, count_c, to_unclipped)

    eye = torch.eye(3, dtype=fv.dtype, device=dev)
    conv = []
    conv_rows = []

    def rebuild(idx, pivot_mask):
        """(p1..p5, b1..b5) of the faces ``idx``; the pivot is the vertex where ``pivot_mask`` is set."""
        i1 = pivot_mask.long().argmax(dim=1)
        i2, i3 = (i1 + 1) % 3, (i1 + 2) % 3
        tri = fv[idx]
        ar = torch.arange(idx.numel(), device=dev)
        p1, p2, p3 = tri[ar, i1], tri[ar, i2], tri[ar, i3]
        p4, w2 = _cut_point(p1, p2, zc, frustum.perspective_correct)
        p5, w3 = _cut_point(p1, p3, zc, frustum.perspective_correct)
        b1, b2, b3 = eye[i1], eye[i2], eye[i3]
        b4 = b1 * (1.0 - w2)[:, None] + b2 * w2[:, None]
        b5 = b1 * (1.0 - w3)[:, None] + b3 * w3[:, None]
        return (p1, p2, p3, p4, p5), (b1, b2, b3, b4, b5)

    if idx_one.numel():
        (p1, _, _, p4, p5), (b1, _, _, b4, b5) = rebuild(idx_one, ~behind[idx_one])
        rows.append(slot[idx_one]); tris.append(torch.stack([p4, p5, p1], dim=1)); origin.append(idx_one)
        conv.append(torch.stack([b4, b5, b1], dim=2)); conv_rows.append(slot[idx_one])
    if idx_two.numel():
        (_, p2, p3, p4, p5), (_, b2, b3, b4, b5) = rebuild(idx_two, behind[idx_two])
        rows += [slot[idx_two], slot[idx_two] + 1]
        tris += [torch.stack([p4, p2, p5], dim=1), torch.stack([p5, p2, p3], dim=1)]
        origin += [idx_two, idx_two]
        conv += [torch.stack([b4, b2, b5], dim=2), torch.stack([b5, b2, b3], dim=2)]
        conv_rows += [slot[idx_two], slot[idx_two] + 1]
    rows_all = torch.cat(rows)
    out = fv.new_zeros((Fc, 3, 3)).index_copy(0, rows_all, torch.cat(tris))
    to_unclipped = torch.zeros(Fc, dtype=torch.long, device=dev).index_copy(0, rows_all, torch.cat(origin))
    conversion = torch.cat(conv)
    conv_rows_all = torch.cat(conv_rows)
    to_conv = torch.full((Fc,), -1, dtype=torch.long, device=dev)
    to_conv[conv_rows_all] = torch.arange(conversion.shape[0], device=dev)
    neighbor = torch.full((Fc,), -1, dtype=torch.long, device=dev)
    if idx_two.numel():
        t1 = slot[idx_two]
        neighbor[t1] = t1 + 1
        neighbor[t1 + 1] = t1
    return ClippedFaces(out, first_c, count_c, to_unclipped, conversion, to_conv, neighbor)


def convert_clipped_rasterization_to_original_faces(
        pix_to_face_clipped: torch.Tensor, bary_coords_clipped: torch.Tensor,
        clipped_faces: ClippedFaces) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fragments indexed by clipped faces -> Fragments indexed by the faces of the mesh: ids through
    ``faces_clipped_to_unclipped_idx``, barycentrics of cut faces through their 3x3 conversion."""
    to_unclipped = clipped_faces.faces_clipped_to_unclipped_idx
    if to_unclipped is None or to_unclipped.numel() == 0:
        return pix_to_face_clipped, bary_coords_clipped
    hit = pix_to_face_clipped >= 0
    safe = pix_to_face_clipped.clamp(min=0)
    pix_to_face = torch.where(hit, to_unclipped[safe], pix_to_face_clipped)
    conversion = clipped_faces.barycentric_conversion
    if conversion is None:
        return pix_to_face, bary_coords_clipped
    which = torch.where(hit, clipped_faces.faces_clipped_to_conversion_idx[safe], torch.full_like(safe, -1))
    sel = (which >= 0).nonzero(as_tuple=True)
    if sel[0].numel() == 0:
        return pix_to_face, bary_coords_clipped
    M = conversion[which[sel]]                       # (S, 3, 3)
    b = bary_coords_clipped[sel]                     # (S, 3)
    mapped = M[:, :, 0] * b[:, 0:1] + M[:, :, 1] * b[:, 1:2] + M[:, :, 2] * b[:, 2:3]
    bary = bary_coords_clipped.index_put(sel, mapped)
    return pix_to_face, bary
