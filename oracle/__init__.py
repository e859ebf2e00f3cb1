"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the rasterise + shade path the reference reaches through PyTorch3D
(see ``trb_oracle.c`` and ``shading_ref.py`` headers).  **Parity unpinned**: the reference
owns no golden vector for this path and PyTorch3D is not installable here.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package.  The product package ``torch_renderer_b200``
never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtrb_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile ``trb_oracle.c`` with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "trb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libtrb_oracle.so"])
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        for name in ("trb_oracle_rasterize_forward", "trb_oracle_rasterize_forward_clipped",
                     "trb_oracle_rasterize_backward",
                     "trb_oracle_interp_forward", "trb_oracle_interp_backward",
                     "trb_oracle_num_threads"):
            getattr(_lib, name).restype = ctypes.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads() -> int:
    return int(lib().trb_oracle_num_threads())


def rasterize_forward(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, image_size,
                      blur_radius=0.0, faces_per_pixel=1, perspective_correct=False,
                      clip_barycentric_coords=False, cull_backfaces=False, num_threads=0,
                      clipped_faces_neighbor_idx=None):
    """numpy in / numpy out twin of ``_C.rasterize_meshes`` with ``bin_size=0`` on CPU tensors.
    ``clipped_faces_neighbor_idx`` i64[F] (-1 = none) is the output of ``clip_ref.clip_faces``.

    Returns (pix_to_face i64[N,H,W,K], zbuf, bary[N,H,W,K,3], dists)."""
    fv = np.ascontiguousarray(face_verts, dtype=np.float32).reshape(-1, 3, 3)
    first = np.ascontiguousarray(mesh_to_face_first_idx, dtype=np.int64)
    count = np.ascontiguousarray(num_faces_per_mesh, dtype=np.int64)
    H, W = (image_size, image_size) if isinstance(image_size, int) else image_size
    N, K = int(first.shape[0]), int(faces_per_pixel)
    p2f = np.empty((N, H, W, K), np.int64)
    zbuf = np.empty((N, H, W, K), np.float32)
    bary = np.empty((N, H, W, K, 3), np.float32)
    dists = np.empty((N, H, W, K), np.float32)
    nbr = None
    if clipped_faces_neighbor_idx is not None:
        nbr = np.ascontiguousarray(clipped_faces_neighbor_idx, dtype=np.int64)
        if nbr.shape != (fv.shape[0],):
            raise ValueError("clipped_faces_neighbor_idx must have one entry per face")
    rc = lib().trb_oracle_rasterize_forward_clipped(
        _p(fv), _p(first), _p(count), None if nbr is None else _p(nbr), N, H, W, K, ctypes.c_float(blur_radius),
        int(perspective_correct), int(clip_barycentric_coords), int(cull_backfaces),
        _p(p2f), _p(zbuf), _p(bary), _p(dists), int(num_threads))
    if rc == 2:
        raise ValueError("faces_per_pixel must be in [1, 150]")
    if rc != 0:
        raise RuntimeError(f"oracle rasterize_forward failed rc={rc}")
    return p2f, zbuf, bary, dists


def rasterize_backward(face_verts, pix_to_face, grad_zbuf, grad_bary, grad_dists,
                       perspective_correct=False, clip_barycentric_coords=False):
    """Twin of ``_C.rasterize_meshes_backward`` -> grad_face_verts f32[F,3,3]."""
    fv = np.ascontiguousarray(face_verts, dtype=np.float32).reshape(-1, 3, 3)
    p2f = np.ascontiguousarray(pix_to_face, dtype=np.int64)
    N, H, W, K = p2f.shape
    gz = np.ascontiguousarray(grad_zbuf, dtype=np.float32)
    gb = np.ascontiguousarray(grad_bary, dtype=np.float32)
    gd = np.ascontiguousarray(grad_dists, dtype=np.float32)
    out = np.zeros_like(fv)
    rc = lib().trb_oracle_rasterize_backward(
        _p(fv), _p(p2f), _p(gz), _p(gb), _p(gd), N, H, W, K, ctypes.c_int64(fv.shape[0]),
        int(perspective_correct), int(clip_barycentric_coords), _p(out))
    if rc != 0:
        raise RuntimeError(f"oracle rasterize_backward failed rc={rc}")
    return out


def interp_forward(pix_to_face, bary, face_attrs):
    p2f = np.ascontiguousarray(pix_to_face, dtype=np.int64)
    b = np.ascontiguousarray(bary, dtype=np.float32)
    fa = np.ascontiguousarray(face_attrs, dtype=np.float32)
    P, D = p2f.size, fa.shape[-1]
    out = np.empty(p2f.shape + (D,), np.float32)
    lib().trb_oracle_interp_forward(_p(p2f), _p(b), _p(fa), ctypes.c_int64(P), D, _p(out))
    return out


def interp_backward(pix_to_face, bary, face_attrs, grad_out):
    p2f = np.ascontiguousarray(pix_to_face, dtype=np.int64)
    b = np.ascontiguousarray(bary, dtype=np.float32)
    fa = np.ascontiguousarray(face_attrs, dtype=np.float32)
    go = np.ascontiguousarray(grad_out, dtype=np.float32)
    P, D = p2f.size, fa.shape[-1]
    gb = np.empty_like(b)
    gfa = np.zeros_like(fa)
    lib().trb_oracle_interp_backward(_p(p2f), _p(b), _p(fa), _p(go), ctypes.c_int64(P),
                                     ctypes.c_int64(fa.shape[0]), D, _p(gb), _p(gfa))
    return gb, gfa
