"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement (numpy fp32, one face at a time) of PyTorch3D's near-plane clipping and frustum culling,
``pytorch3d/renderer/mesh/clip.py`` (``clip_faces``, ``_find_verts_intersecting_clipping_plane``,
``_get_culled_faces``, ``convert_clipped_rasterization_to_original_faces``), which ``rasterize_meshes`` runs on
``face_verts`` whenever ``z_clip_value is not None or cull_to_frustum`` -- i.e. on every render of the reference's
``FoVPerspectiveCameras`` scripts (camera_pose_optimizer.py:105, mesh_deformer.py:119, batch_rendering_test.py:225),
where ``MeshRasterizer`` sets ``z_clip_value = znear / 2``.  PyTorch3D is an un-vendored, un-pinned dependency of the
reference and not installable here: **parity unpinned** (SURVEY.md 8c, 8f rank 3); the algorithm is restated from its
published source, and pinned by the hand-checkable scenes of tests/test_clip.py.

Semantics restated (F faces in, F' faces out, ascending order kept):
  * a vertex is "behind" when ``z < z_clip_value``; a face with 3 vertices behind, or (``cull``) with all three
    vertices outside one of the planes x < left, x > right, y < top, y > bottom, z < znear, z > zfar, is removed;
  * 2 behind (vertex p1 in front): the face becomes the triangle (p4, p5, p1);
  * 1 behind (vertex p1 behind): the face becomes the two triangles (p4, p2, p5), (p5, p2, p3) stored next to each
    other and naming each other in ``clipped_faces_neighbor_idx``;
    with p2, p3 the vertices after p1 in cyclic order, p4 on p1p2 and p5 on p1p3 where z = z_clip_value:
    ``w = (p1.z - c) / (p1.z - p.z)``, ``p4 = p1 * (1 - w) + p * w``; with ``perspective_correct`` x, y of the new
    vertex are interpolated in view space: ``(p1.xy * p1.z * (1 - w) + p.xy * p.z * w) / c``;
  * ``barycentric_conversion[t]`` (3x3) has as COLUMNS the barycentric coordinates of the clipped triangle's three
    vertices in the original triangle: ``bary_original = M @ bary_clipped``.
  * when nothing is behind the plane and nothing is culled, the input comes back untouched (all extras ``None``).
Every operator is one fp32 operation, in the order written above.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np

f32 = np.float32


class ClipFrustum(NamedTuple):
    left: Optional[float] = None
    right: Optional[float] = None
    top: Optional[float] = None
    bottom: Optional[float] = None
    znear: Optional[float] = None
    zfar: Optional[float] = None
    perspective_correct: bool = False
    cull: bool = True
    z_clip_value: Optional[float] = None


class ClippedFaces(NamedTuple):
    face_verts: np.ndarray
    mesh_to_face_first_idx: np.ndarray
    num_faces_per_mesh: np.ndarray
    faces_clipped_to_unclipped_idx: Optional[np.ndarray] = None
    barycentric_conversion: Optional[np.ndarray] = None
    faces_clipped_to_conversion_idx: Optional[np.ndarray] = None
    clipped_faces_neighbor_idx: Optional[np.ndarray] = None


def rasterizer_frustum(perspective_correct: bool, z_clip_value, cull_to_frustum: bool) -> ClipFrustum:
    """The frustum ``rasterize_meshes`` builds: the NDC square [-1, 1]^2, no znear / zfar planes."""
    return ClipFrustum(left=-1.0, right=1.0, top=-1.0, bottom=1.0, perspective_correct=bool(perspective_correct),
                       cull=bool(cull_to_frustum), z_clip_value=z_clip_value)


def _outside_frustum(tri: np.ndarray, fr: ClipFrustum) -> bool:
    if not fr.cull:
        return False
    planes = ((fr.left, 0, -1), (fr.right, 0, +1), (fr.top, 1, -1), (fr.bottom, 1, +1), (fr.znear, 2, -1),
              (fr.zfar, 2, +1))
    for value, axis, side in planes:
        if value is None:
            continue
        col = tri[:, axis]
        out = (col < f32(value)) if side < 0 else (col > f32(value))
        if out.all():
            return True
    return False


def _cut(p1: np.ndarray, p: np.ndarray, c: np.float32, perspective_correct: bool):
    """Point of the segment p1 -> p at depth c, and its weight on p."""
    w = f32(f32(p1[2] - c) / f32(p1[2] - p[2]))
    one_w = f32(f32(1.0) - w)
    q = (p1 * one_w + p * w).astype(f32)
    if perspective_correct:
        a = (p1[:2] * p1[2]).astype(f32)
        b = (p[:2] * p[2]).astype(f32)
        q[:2] = ((a * one_w + b * w) / c).astype(f32)
    return q, w


def clip_faces(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, frustum: ClipFrustum) -> ClippedFaces:
    fv = np.ascontiguousarray(face_verts, dtype=f32).reshape(-1, 3, 3)
    first = np.asarray(mesh_to_face_first_idx, dtype=np.int64)
    count = np.asarray(num_faces_per_mesh, dtype=np.int64)
    F = fv.shape[0]
    zc = frustum.z_clip_value
    behind = (fv[:, :, 2] < f32(zc)) if zc is not None else np.zeros((F, 3), bool)
    culled = np.array([_outside_frustum(fv[i], frustum) for i in range(F)], dtype=bool).reshape(F)
    if behind.sum() == 0 and culled.sum() == 0:
        return ClippedFaces(fv, first, count)

    out_verts, to_unclipped, conv, to_conv, neighbor = [], [], [], [], []
    new_first_of_face = np.zeros(F + 1, np.int64)
    any_cut = False
    eye = np.eye(3, dtype=f32)
    for i in range(F):
        new_first_of_face[i] = len(out_verts)
        nb = int(behind[i].sum())
        if culled[i] or nb == 3:
            continue
        if nb == 0:
            out_verts.append(fv[i]); to_unclipped.append(i); to_conv.append(-1); neighbor.append(-1)
            continue
        any_cut = True
        c = f32(zc)
        # pivot: the lone vertex in front (2 behind) or the lone vertex behind (1 behind)
        i1 = int(np.argmin(behind[i])) if nb == 2 else int(np.argmax(behind[i]))
        i2, i3 = (i1 + 1) % 3, (i1 + 2) % 3
        p1, p2, p3 = fv[i, i1], fv[i, i2], fv[i, i3]
        p4, w2 = _cut(p1, p2, c, frustum.perspective_correct)
        p5, w3 = _cut(p1, p3, c, frustum.perspective_correct)
        b1, b2, b3 = eye[i1], eye[i2], eye[i3]
        b4 = (b1 * f32(f32(1.0) - w2) + b2 * w2).astype(f32)
        b5 = (b1 * f32(f32(1.0) - w3) + b3 * w3).astype(f32)
        if nb == 2:
            tris = [((p4, p5, p1), (b4, b5, b1))]
        else:
            tris = [((p4, p2, p5), (b4, b2, b5)), ((p5, p2, p3), (b5, b2, b3))]
        base = len(out_verts)
        for j, (pts, bs) in enumerate(tris):
            out_verts.append(np.stack(pts).astype(f32))
            to_unclipped.append(i)
            to_conv.append(-2)          # patched below (PyTorch3D orders the matrices: all 2-behind faces,
            conv.append((nb, j, np.stack(bs, axis=1).astype(f32)))  # then first halves, then second halves)
            neighbor.append(-1 if len(tris) == 1 else base + 1 - j)
    new_first_of_face[F] = len(out_verts)
    Fc = len(out_verts)
    face_verts_c = np.stack(out_verts).astype(f32) if Fc else np.zeros((0, 3, 3), f32)
    first_c = new_first_of_face[first]
    count_c = np.concatenate([first_c[1:], [Fc]]) - first_c
    if not any_cut:
        return ClippedFaces(face_verts_c, first_c, count_c, np.asarray(to_unclipped, np.int64))
    # conversion matrices in PyTorch3D's order; to_conv points each clipped face at its matrix
    order = sorted(range(len(conv)), key=lambda t: (0 if conv[t][0] == 2 else 1 + conv[t][1], t))
    rank = {t: r for r, t in enumerate(order)}
    mats = np.stack([conv[t][2] for t in order]).astype(f32)
    to_conv_arr = np.asarray(to_conv, np.int64)
    cut_rows = np.nonzero(to_conv_arr == -2)[0]
    for t, row in enumerate(cut_rows):
        to_conv_arr[row] = rank[t]
    return ClippedFaces(face_verts_c, first_c, count_c, np.asarray(to_unclipped, np.int64), mats, to_conv_arr,
                        np.asarray(neighbor, np.int64))


def convert_clipped_rasterization_to_original_faces(pix_to_face_clipped, bary_coords_clipped, clipped: ClippedFaces):
    """Clipped face ids -> original face ids; barycentrics of cut faces -> barycentrics of the original face."""
    p2f = np.asarray(pix_to_face_clipped, np.int64)
    bary = np.asarray(bary_coords_clipped, f32)
    if clipped.faces_clipped_to_unclipped_idx is None or clipped.faces_clipped_to_unclipped_idx.size == 0:
        return p2f, bary
    hit = p2f >= 0
    out_p2f = np.where(hit, clipped.faces_clipped_to_unclipped_idx[np.where(hit, p2f, 0)], -1)
    if clipped.barycentric_conversion is None:
        return out_p2f, bary
    out_bary = bary.copy()
    which = np.where(hit, clipped.faces_clipped_to_conversion_idx[np.where(hit, p2f, 0)], -1)
    for idx in zip(*np.nonzero(which >= 0)):
        M = clipped.barycentric_conversion[which[idx]]
        b = bary[idx]
        # batched matrix-vector product, accumulated left to right in fp32 (torch.bmm on a 3x3 @ 3x1)
        out_bary[idx] = np.array([f32(f32(f32(M[r, 0] * b[0]) + f32(M[r, 1] * b[1])) + f32(M[r, 2] * b[2]))
                                  for r in range(3)], f32)
    return out_p2f, out_bary
