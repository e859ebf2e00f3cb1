"""Shared scene builders and the oracle pipeline used by the tests (and by smoke()/bench.py)."""
import math
import os

import numpy as np
import torch

import oracle
from oracle import shading_ref as sref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_MESHES = None


def load_mesh(name):
    """'teapot' | 'cow' -> (verts f32 [V,3], faces i64 [F,3]) CPU tensors (tests/golden/meshes.npz)."""
    global _MESHES
    if _MESHES is None:
        _MESHES = np.load(os.path.join(GOLDEN, "meshes.npz"))
    return (torch.from_numpy(_MESHES[f"{name}_verts"].copy()),
            torch.from_numpy(_MESHES[f"{name}_faces"].astype(np.int64)))


def cow_uvs():
    load_mesh("cow")
    return (torch.from_numpy(_MESHES["cow_verts_uvs"].copy()),
            torch.from_numpy(_MESHES["cow_faces_uvs"].astype(np.int64)))


def uv_sphere(rings=12, segments=16, radius=1.0, noise=0.0, seed=0):
    """Closed lat-long sphere: V = rings*segments + 2, F = 2*segments*rings."""
    g = torch.Generator().manual_seed(seed)
    verts = [[0.0, radius, 0.0]]
    for r in range(1, rings + 1):
        th = math.pi * r / (rings + 1)
        for s in range(segments):
            ph = 2 * math.pi * s / segments
            verts.append([radius * math.sin(th) * math.cos(ph), radius * math.cos(th),
                          radius * math.sin(th) * math.sin(ph)])
    verts.append([0.0, -radius, 0.0])
    faces = []
    idx = lambda r, s: 1 + (r - 1) * segments + (s % segments)
    for s in range(segments):
        faces.append([0, idx(1, s + 1), idx(1, s)])
        faces.append([len(verts) - 1, idx(rings, s), idx(rings, s + 1)])
    for r in range(1, rings):
        for s in range(segments):
            a, b, c, d = idx(r, s), idx(r, s + 1), idx(r + 1, s), idx(r + 1, s + 1)
            faces.append([a, b, c])
            faces.append([b, d, c])
    v = torch.tensor(verts, dtype=torch.float32)
    if noise > 0:
        v = v * (1 + noise * torch.randn(v.shape[0], 1, generator=g))
    return v, torch.tensor(faces, dtype=torch.int64)


def normalize_mesh(verts):
    c = (verts.max(0)[0] + verts.min(0)[0]) / 2
    s = (verts - c).abs().max()
    return (verts - c) / s


def fov_proj(n, fov_deg=60.0, aspect=1.0):
    t = math.tan(math.radians(fov_deg) / 2)
    return torch.tensor([[1 / (t * aspect), 1 / t, 0.0, 0.0]], dtype=torch.float32).repeat(n, 1)


def oracle_rasterize(verts_ndc, faces, image_size, blur_radius=0.0, K=1, persp=False, clip=False,
                     cull=False, threads=0):
    """verts_ndc [N,V,3] (tensor), faces [F,3] shared by all views -> oracle fragments (numpy).
    pix_to_face indexes the packed list: view n's faces start at n*F."""
    N, V, _ = verts_ndc.shape
    F = faces.shape[0]
    fv = verts_ndc.detach().cpu().float()[:, faces.cpu()]  # [N,F,3,3]
    first = np.arange(N, dtype=np.int64) * F
    count = np.full((N,), F, dtype=np.int64)
    return oracle.rasterize_forward(fv.reshape(-1, 3, 3).numpy(), first, count, image_size, blur_radius, K,
                                    persp, clip, cull, threads)


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def oracle_rasterize_clipped(verts_ndc, faces, image_size, blur_radius=0.0, K=1, persp=False, clip=False,
                             cull=False, z_clip_value=None, cull_to_frustum=False, threads=0):
    """The upstream route for a batch with an active near plane: clip_faces -> rasterise (with the neighbour
    table) -> convert back to the faces of the batch.  Returns (fragments, ClippedFaces)."""
    from oracle import clip_ref
    N, V, _ = verts_ndc.shape
    F = faces.shape[0]
    fv = verts_ndc.detach().cpu().float()[:, faces.cpu()].reshape(-1, 3, 3).numpy()
    first = np.arange(N, dtype=np.int64) * F
    count = np.full((N,), F, dtype=np.int64)
    cf = clip_ref.clip_faces(fv, first, count, clip_ref.rasterizer_frustum(persp, z_clip_value, cull_to_frustum))
    p2f, zbuf, bary, dists = oracle.rasterize_forward(
        cf.face_verts, cf.mesh_to_face_first_idx, cf.num_faces_per_mesh, image_size, blur_radius, K, persp, clip,
        cull, threads, clipped_faces_neighbor_idx=cf.clipped_faces_neighbor_idx)
    p2f_u, bary_u = clip_ref.convert_clipped_rasterization_to_original_faces(p2f, bary, cf)
    return (p2f_u, zbuf, bary_u, dists), cf, p2f
