"""ORACLE -- TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity unpinned.

Plain-torch (CPU, fp32 or fp64) restatement of the *differentiable* part of the path:
the camera transform, the barycentric/zbuf/dists recomputation for a fixed ``pix_to_face``,
attribute interpolation, vertex normals, Phong lighting and the three blend functions, as
written down in SURVEY.md Appendix A (A1, A2, A4, A6, A7, A8).  These restate PyTorch3D's
``renderer/mesh/shading.py``, ``renderer/lighting.py``, ``renderer/blending.py``,
``structures/meshes.py::verts_normals_packed`` and ``renderer/cameras.py`` -- the un-vendored
dependency every reference script calls (renderer.py:87-101, torch_renderer.py:102-158,
camera_pose_optimizer.py:130-158, mesh_deformer.py:142-145).

Run in fp64 with ``requires_grad`` inputs it is the autograd truth the CUDA backward kernels
are compared with; run in fp32 it doubles as the CPU shading baseline in ``bench.py``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

K_EPS = 1e-8


# ----------------------------------------------------------------------------- cameras (A1, A2)
def world_to_ndc(verts, R, T, fx, fy, px, py, perspective=True):
    """verts [V,3] (shared) or [N,V,3]; R [N,3,3]; T [N,3]; fx..py [N] -> [N,V,3].

    X_view = X_world @ R + T;  x_ndc = fx*X/Z + px (perspective) or fx*X + px; z = Z_view."""
    if verts.dim() == 2:
        verts = verts[None].expand(R.shape[0], -1, -1)
    view = torch.matmul(verts, R) + T[:, None, :]
    z = view[..., 2]
    den = z if perspective else torch.ones_like(z)
    x = fx[:, None] * view[..., 0] / den + px[:, None]
    y = fy[:, None] * view[..., 1] / den + py[:, None]
    return torch.stack([x, y, z], dim=-1)


# ----------------------------------------------------------------------------- raster (A3, A4)
def pixel_centers(H, W, dtype=torch.float32):
    def ndc(i, S1, S2):
        rng = 2.0 if S1 <= S2 else (S1 * 2.0) / S2
        off = rng / 2.0
        return -off + (rng * i + off) / S1
    ys = torch.tensor([ndc(H - 1 - yi, H, W) for yi in range(H)], dtype=dtype)
    xs = torch.tensor([ndc(W - 1 - xi, W, H) for xi in range(W)], dtype=dtype)
    return ys, xs


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def _seg_d2(px, py, ax, ay, bx, by):
    bax, bay = bx - ax, by - ay
    l2 = bax * bax + bay * bay
    degenerate = l2 <= K_EPS
    l2s = torch.where(degenerate, torch.ones_like(l2), l2)
    # t is treated as a constant in the backward.
    t = ((bax * (px - ax) + bay * (py - ay)) / l2s).clamp(0.0, 1.0).detach()
    qx, qy = ax + t * bax, ay + t * bay
    d = (qx - px) ** 2 + (qy - py) ** 2
    d_deg = (px - bx) ** 2 + (py - by) ** 2
    return torch.where(degenerate, d_deg, d)


def raster_recompute(face_verts, pix_to_face, perspective_correct, clip_barycentric_coords):
    """Differentiable zbuf / bary / dists for a FIXED pix_to_face [N,H,W,K] (A4 steps 3-8)."""
    N, H, W, K = pix_to_face.shape
    dt = face_verts.dtype
    ys, xs = pixel_centers(H, W, dt)
    py = ys.view(1, H, 1, 1).expand(N, H, W, K)
    px = xs.view(1, 1, W, 1).expand(N, H, W, K)
    mask = pix_to_face >= 0
    f = pix_to_face.clamp(min=0)
    v = face_verts.reshape(-1, 3, 3)[f]  # [N,H,W,K,3,3]
    x0, y0, z0 = v[..., 0, 0], v[..., 0, 1], v[..., 0, 2]
    x1, y1, z1 = v[..., 1, 0], v[..., 1, 1], v[..., 1, 2]
    x2, y2, z2 = v[..., 2, 0], v[..., 2, 1], v[..., 2, 2]
    area = _edge(x2, y2, x0, y0, x1, y1) + K_EPS
    w0 = _edge(px, py, x1, y1, x2, y2) / area
    w1 = _edge(px, py, x2, y2, x0, y0) / area
    w2 = _edge(px, py, x0, y0, x1, y1) / area
    b0, b1, b2 = w0, w1, w2
    if perspective_correct:
        t0, t1, t2 = w0 * z1 * z2, w1 * z0 * z2, w2 * z0 * z1
        den = (t0 + t1 + t2).clamp(min=K_EPS)
        b0, b1, b2 = t0 / den, t1 / den, t2 / den
    c0, c1, c2 = b0, b1, b2
    if clip_barycentric_coords:
        c0, c1, c2 = b0.clamp(min=0), b1.clamp(min=0), b2.clamp(min=0)
        s = (c0 + c1 + c2).clamp(min=1e-5)
        c0, c1, c2 = c0 / s, c1 / s, c2 / s
    pz = c0 * z0 + c1 * z1 + c2 * z2
    e01 = _seg_d2(px, py, x0, y0, x1, y1)
    e02 = _seg_d2(px, py, x0, y0, x2, y2)
    e12 = _seg_d2(px, py, x1, y1, x2, y2)
    # arg-min edge with the e01 <= e02 <= e12 tie order of the C oracle
    use01 = (e01 <= e02) & (e01 <= e12)
    use02 = (~use01) & (e02 <= e01) & (e02 <= e12)
    dist = torch.where(use01, e01, torch.where(use02, e02, e12))
    inside = (b0 > 0) & (b1 > 0) & (b2 > 0)
    signed = torch.where(inside, -dist, dist)
    neg = torch.full_like(pz, -1.0)
    zbuf = torch.where(mask, pz, neg)
    dists = torch.where(mask, signed, neg)
    bary = torch.where(mask[..., None], torch.stack([c0, c1, c2], dim=-1), neg[..., None])
    return zbuf, bary, dists


# ----------------------------------------------------------------------------- interpolation (A6)
def interpolate_face_attributes(pix_to_face, bary, face_attrs):
    """face_attrs [F,3,D] -> [N,H,W,K,D]; zero where pix_to_face < 0."""
    mask = pix_to_face >= 0
    a = face_attrs[pix_to_face.clamp(min=0)]  # [...,3,D]
    out = (bary[..., None] * a).sum(dim=-2)
    return torch.where(mask[..., None], out, torch.zeros_like(out))


def vertex_normals(verts, faces):
    """Area-weighted vertex normals (restates Meshes.verts_normals_packed, A6)."""
    v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    fn = torch.cross(v2 - v1, v0 - v1, dim=1)  # |fn| = 2 * area  => area weighting
    n = torch.zeros_like(verts)
    n = n.index_add(0, faces[:, 0], fn)
    n = n.index_add(0, faces[:, 1], fn)
    n = n.index_add(0, faces[:, 2], fn)
    return F.normalize(n, eps=1e-6, dim=1)


# ----------------------------------------------------------------------------- lighting (A6)
def phong_colors(points, normals, texels, view_index_shape, light_kind, light_vec, light_ambient,
                 light_diffuse, light_specular, mat_ambient, mat_diffuse, mat_specular, shininess,
                 camera_center):
    """points/normals/texels [N,H,W,K,3]; per-view params [N,3] (shininess [N]).

    light_kind: 'point' (light_vec = location), 'directional' (light_vec = direction) or
    'ambient' (diffuse = specular = 0)."""
    def bc(t):  # [N,3] -> [N,1,1,1,3]
        return t[:, None, None, None, :]
    ambient = bc(mat_ambient) * bc(light_ambient)
    if light_kind == "ambient":
        return ambient * texels
    direction = bc(light_vec) - points if light_kind == "point" else bc(light_vec).expand_as(points)
    n = F.normalize(normals, p=2, dim=-1, eps=1e-6)
    d = F.normalize(direction, p=2, dim=-1, eps=1e-6)
    cos = (n * d).sum(-1)
    diffuse = bc(mat_diffuse) * bc(light_diffuse) * F.relu(cos)[..., None]
    mask = (cos > 0).to(points.dtype)
    view_dir = F.normalize(bc(camera_center) - points, p=2, dim=-1, eps=1e-6)
    reflect = -d + 2 * (cos[..., None] * n)
    alpha = F.relu((view_dir * reflect).sum(-1)) * mask
    spec = bc(mat_specular) * bc(light_specular) * torch.pow(alpha, shininess[:, None, None, None])[..., None]
    return (ambient + diffuse) * texels + spec


# ----------------------------------------------------------------------------- blending (A8)
def softmax_rgb_blend(colors, pix_to_face, zbuf, dists, sigma, gamma, background, znear, zfar):
    eps = 1e-10
    mask = (pix_to_face >= 0).to(colors.dtype)
    prob = torch.sigmoid(-dists / sigma) * mask
    alpha = torch.prod(1.0 - prob, dim=-1)
    z_inv = (zfar - zbuf) / (zfar - znear) * mask
    z_inv_max = torch.max(z_inv, dim=-1).values[..., None].clamp(min=eps)
    w = prob * torch.exp((z_inv - z_inv_max) / gamma)
    delta = torch.exp((eps - z_inv_max) / gamma).clamp(min=eps)
    denom = w.sum(dim=-1)[..., None] + delta
    bg = torch.as_tensor(background, dtype=colors.dtype)
    rgb = ((w[..., None] * colors).sum(dim=-2) + delta * bg) / denom
    return torch.cat([rgb, (1.0 - alpha)[..., None]], dim=-1)


def sigmoid_alpha_blend(colors, pix_to_face, dists, sigma):
    mask = (pix_to_face >= 0).to(colors.dtype)
    prob = torch.sigmoid(-dists / sigma) * mask
    alpha = torch.prod(1.0 - prob, dim=-1)
    return torch.cat([colors[..., 0, :], (1.0 - alpha)[..., None]], dim=-1)


def hard_rgb_blend(colors, pix_to_face, background):
    is_bg = pix_to_face[..., 0] < 0
    bg = torch.as_tensor(background, dtype=colors.dtype)
    rgb = torch.where(is_bg[..., None], bg.expand_as(colors[..., 0, :]), colors[..., 0, :])
    return torch.cat([rgb, (~is_bg).to(colors.dtype)[..., None]], dim=-1)


# ----------------------------------------------------------------------------- whole shader
def shade(pix_to_face, bary, zbuf, dists, faces_rows, verts_world, vert_normals, vert_colors, *,
          shader="soft_phong", light_kind="point", light_vec=None, light_ambient=None,
          light_diffuse=None, light_specular=None, mat_ambient=None, mat_diffuse=None,
          mat_specular=None, shininess=None, camera_center=None, sigma=1e-4, gamma=1e-4,
          background=(1.0, 1.0, 1.0), znear=1.0, zfar=100.0, texels=None):
    """Restates SoftPhongShader / HardPhongShader / SoftSilhouetteShader forward.

    faces_rows i64 [F_total,3]: vertex ids (into verts_world / vert_normals / vert_colors) of
    every packed face pix_to_face can name."""
    if shader == "soft_silhouette":
        colors = torch.ones_like(bary)
        return sigmoid_alpha_blend(colors, pix_to_face, dists, sigma)
    pts = interpolate_face_attributes(pix_to_face, bary, verts_world[faces_rows])
    nrm = interpolate_face_attributes(pix_to_face, bary, vert_normals[faces_rows])
    if texels is None:
        texels = interpolate_face_attributes(pix_to_face, bary, vert_colors[faces_rows])
    colors = phong_colors(pts, nrm, texels, None, light_kind, light_vec, light_ambient, light_diffuse,
                          light_specular, mat_ambient, mat_diffuse, mat_specular, shininess,
                          camera_center)
    if shader == "hard_phong":
        return hard_rgb_blend(colors, pix_to_face, background)
    return softmax_rgb_blend(colors, pix_to_face, zbuf, dists, sigma, gamma, background, znear, zfar)
