"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement of PyTorch3D's point-cloud rendering path, which the reference's ``AlphaPointRender`` /
``NormPointRender`` wrap (torch_renderer.py:163-208; marked "not tested" there): ``rasterize_points`` (naive CPU
rasteriser, C twin in ``trb_oracle.c``), ``PointsRenderer.forward``'s weights ``1 - dist^2 / r^2``, and the two
compositors ``alpha_composite`` / ``norm_weighted_sum`` (pytorch3d/renderer/compositing.py,
csrc/compositing/alpha_composite_cpu.cpp, norm_weighted_sum_cpu.cpp) with ``_add_background_color_to_images``.
PyTorch3D is an un-vendored, un-pinned dependency of the reference and is not installable here: **parity unpinned**.

The compositors are written in plain torch so that fp64 autograd through them is the gradient truth.
Layouts here are channels-last: idx / alphas (N, H, W, K), features (P, C) -> images (N, H, W, C).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _p, lib

K_EPS_NORM = 1e-4   # norm_weighted_sum clamps the sum of weights here


def rasterize_points(points_packed, first, count, radius, image_size, points_per_pixel):
    """numpy in / out: points f32 (P, 3) in NDC x, y + view z; per-cloud ranges; radius float or (P,).
    Returns idx i32 (N, H, W, K) into the packed points, zbuf, dists (squared NDC distance)."""
    pts = np.ascontiguousarray(points_packed, np.float32).reshape(-1, 3)
    first = np.ascontiguousarray(first, np.int64)
    count = np.ascontiguousarray(count, np.int64)
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (pts.shape[0],)))
    H, W = (image_size, image_size) if isinstance(image_size, int) else image_size
    N, K = first.shape[0], int(points_per_pixel)
    idx = np.empty((N, H, W, K), np.int32)
    zbuf = np.empty((N, H, W, K), np.float32)
    dists = np.empty((N, H, W, K), np.float32)
    fn = lib().trb_oracle_rasterize_points_forward
    fn.restype = ctypes.c_int
    rc = fn(_p(pts), _p(first), _p(count), _p(r), N, H, W, K, _p(idx), _p(zbuf), _p(dists))
    if rc == 2:
        raise ValueError("points_per_pixel must be in [1, 150]")
    if rc != 0:
        raise RuntimeError(f"oracle rasterize_points failed rc={rc}")
    return idx, zbuf, dists


def rasterize_points_backward(points_packed, idx, grad_zbuf, grad_dists):
    pts = np.ascontiguousarray(points_packed, np.float32).reshape(-1, 3)
    idx = np.ascontiguousarray(idx, np.int32)
    N, H, W, K = idx.shape
    out = np.zeros(pts.shape, np.float64)
    fn = lib().trb_oracle_rasterize_points_backward
    fn.restype = ctypes.c_int
    fn(_p(pts), _p(idx), _p(np.ascontiguousarray(grad_zbuf, np.float32)),
       _p(np.ascontiguousarray(grad_dists, np.float32)), N, H, W, K, _p(out))
    return out


def _gather(features, idx):
    """features (P, C), idx (N, H, W, K) -> (N, H, W, K, C), zero where idx < 0."""
    f = features[idx.clamp(min=0).long()]
    return f * (idx >= 0)[..., None].to(f.dtype)


def alpha_composite(idx, alphas, features):
    """out[c] = sum_k f[idx_k, c] * a_k * prod_{j<k, idx_j >= 0} (1 - a_j); empty slots are skipped."""
    hit = (idx >= 0).to(alphas.dtype)
    a = alphas * hit
    keep = 1.0 - a
    cum = torch.cumprod(torch.cat([torch.ones_like(keep[..., :1]), keep[..., :-1]], dim=-1), dim=-1)
    return (_gather(features, idx) * (a * cum)[..., None]).sum(dim=-2)


def norm_weighted_sum(idx, alphas, features):
    """out[c] = sum_k f[idx_k, c] * a_k / max(sum_k a_k, 1e-4) over the filled slots."""
    hit = (idx >= 0).to(alphas.dtype)
    a = alphas * hit
    total = a.sum(dim=-1, keepdim=True).clamp(min=K_EPS_NORM)
    return (_gather(features, idx) * (a / total)[..., None]).sum(dim=-2)


def add_background(images, idx, background_color):
    """Pixels without any point (first slot empty) take the background colour; a C-1 colour gets alpha = 1."""
    if background_color is None:
        return images
    bg = torch.as_tensor(background_color, dtype=images.dtype)
    if bg.ndim == 0:
        bg = bg.expand(images.shape[-1])
    if bg.shape[0] + 1 == images.shape[-1]:
        bg = torch.cat([bg, bg.new_ones(1)])
    if bg.shape[0] != images.shape[-1]:
        raise ValueError("background color has %d channels, images have %d" % (bg.shape[0], images.shape[-1]))
    return torch.where((idx[..., 0] < 0)[..., None], bg, images)


def render_points(idx, dists, features, radius, mode="alpha", background_color=None):
    """PointsRenderer.forward after rasterisation: weights = 1 - dist^2 / r^2, compositor, background."""
    weights = 1.0 - dists / (radius * radius)
    comp = alpha_composite if mode == "alpha" else norm_weighted_sum
    return add_background(comp(idx, weights, features), idx, background_color)
