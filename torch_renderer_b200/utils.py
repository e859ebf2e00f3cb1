"""``ico_sphere`` (mesh_deformer.py:15, deform_mesh_with_color.py:7) and the OpenCV camera helper
(renderer.py:10) -- the ``pytorch3d.utils`` names the reference imports.  Host-side only."""
from __future__ import annotations

import math

import torch

from .cameras import cameras_from_opencv_projection  # noqa: F401  (re-export)
from .structures import Meshes


def _icosahedron():
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = torch.tensor([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t],
                      [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=torch.float32)
    v = v / v.norm(dim=1, keepdim=True)
    f = torch.tensor([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                      [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                      [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=torch.int64)
    return v, f


def _subdivide(verts: torch.Tensor, faces: torch.Tensor):
    """Loop-style 1->4 split: one new vertex per unique edge, projected back to the unit sphere."""
    V = verts.shape[0]
    e = torch.cat([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], dim=0)
    e_sorted, _ = e.sort(dim=1)
    uniq, inv = torch.unique(e_sorted, dim=0, return_inverse=True)
    mid = verts[uniq].mean(dim=1)
    mid = mid / mid.norm(dim=1, keepdim=True)
    F = faces.shape[0]
    m01, m12, m20 = inv[:F] + V, inv[F:2 * F] + V, inv[2 * F:] + V
    f0, f1, f2 = faces[:, 0], faces[:, 1], faces[:, 2]
    new_faces = torch.cat([
        torch.stack([f0, m01, m20], dim=1), torch.stack([f1, m12, m01], dim=1),
        torch.stack([f2, m20, m12], dim=1), torch.stack([m01, m12, m20], dim=1)], dim=0)
    return torch.cat([verts, mid], dim=0), new_faces


def ico_sphere(level: int = 0, device=None) -> Meshes:
    """Unit icosphere; level L has 10*4^L + 2 vertices and 20*4^L faces (level 4: 2,562 / 5,120)."""
    if level < 0:
        raise ValueError("level must be >= 0.")
    v, f = _icosahedron()
    for _ in range(level):
        v, f = _subdivide(v, f)
    if device is not None:
        v, f = v.to(device), f.to(device)
    return Meshes(verts=[v], faces=[f])
