"""ctypes binding of ``libtrb.so`` (the C ABI declared in ``include/trb.h``).

There is no fallback: if the shared library is missing or a symbol is absent this module raises,
and every op in the package raises with it.  Pointers are passed as raw ``data_ptr()`` integers;
the caller (the autograd Functions in ``ops.py``) owns all memory.
"""
from __future__ import annotations

import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
# TRB_LIB_PATH selects another build of the same ABI (same-box A/B of kernel variants: profiles/ab_run.sh)
LIB_PATH = os.environ.get("TRB_LIB_PATH") or os.path.join(_PKG, "libtrb.so")

ABI_VERSION = 5   # include/trb.h TRB_ABI_VERSION
TRB_OK, TRB_ERR_BAD_ARG, TRB_ERR_K_TOO_LARGE, TRB_ERR_WORKSPACE, TRB_ERR_CUDA = range(5)

PERSPECTIVE_CORRECT, CLIP_BARYCENTRIC, CULL_BACKFACES = 1, 2, 4
SHADER_NONE, SHADER_SOFT_PHONG, SHADER_HARD_PHONG, SHADER_SOFT_SILHOUETTE = -1, 0, 1, 2
LIGHT_AMBIENT, LIGHT_POINT, LIGHT_DIRECTIONAL = 0, 1, 2
TEX_VERTEX, TEX_TEXELS, TEX_UV = 0, 1, 2
VIEW_PARAM_STRIDE = 20
MAX_FACES_PER_PIXEL = 150

_c = ctypes
_vp, _i, _i64, _f, _u32, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float, _c.c_uint32, _c.c_size_t


class ShadeConfig(ctypes.Structure):
    _fields_ = [("N", _c.c_int32), ("H", _c.c_int32), ("W", _c.c_int32), ("K", _c.c_int32),
                ("shader", _c.c_int32), ("light_kind", _c.c_int32), ("texture_mode", _c.c_int32),
                ("sigma", _f), ("gamma", _f), ("background", _f * 3)]


class RenderConfig(ctypes.Structure):
    _fields_ = [("shade", ShadeConfig), ("blur_radius", _f), ("raster_flags", _u32),
                ("perspective", _c.c_int32), ("max_face_count", _c.c_int32),
                ("max_vert_count", _c.c_int32), ("camera_center_from_rt", _c.c_int32),
                ("want_light_grad", _c.c_int32), ("z_clip_value", _f),
                ("num_world_verts", _c.c_int64), ("num_faces", _c.c_int64),
                ("num_ndc_verts", _c.c_int64), ("pair_capacity", _c.c_int64),
                ("scratch_is_zeroed", _c.c_int32), ("sparse_fragments", _c.c_int32)]


class UvTexture(ctypes.Structure):
    _fields_ = [("map", _vp), ("verts_uvs", _vp), ("faces_uvs", _vp), ("grad_map", _vp),
                ("map_h", _c.c_int32), ("map_w", _c.c_int32)]


class RenderExtras(ctypes.Structure):
    """trb_render_extras: work folded into the forward's first kernel (parameter-block copy, zero fill)."""
    _fields_ = [("view_params_src", _vp), ("zero_buffer", _vp), ("zero_count", _i64)]


class PeerSum(ctypes.Structure):
    """trb_peer_sum: what trb_render_backward_allreduce needs to push the shared gradients to the peers."""
    _fields_ = [("host_peer_inbox", _vp), ("capacity_floats", _i64), ("rank", _c.c_int32), ("world", _c.c_int32),
                ("epochs", _vp), ("error_flag", _vp), ("done_counter", _vp)]


# name -> argtypes; every function returns int (trb_status) unless noted
_SIGNATURES = {
    "trb_abi_version": [],
    "trb_last_cuda_error": [],
    "trb_abi_struct_size": [_i],
    "trb_transform_forward": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp],
    "trb_transform_backward": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "trb_raster_workspace_bytes": [_i, _i, _i, _i, _i64, _c.POINTER(_sz)],
    "trb_raster_forward": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _u32, _i64, _vp, _sz, _vp, _vp, _vp, _vp,
                           _vp, _i, _vp],
    "trb_raster_backward": [_vp, _vp, _vp, _i, _i, _i, _i, _u32, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "trb_any_vertex_behind": [_vp, _vp, _vp, _vp, _i, _i, _f, _c.c_int32, _vp, _vp, _vp, _i, _vp],
    "trb_clip_resequence": [_vp, _vp, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _f, _u32, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "trb_interp_forward": [_vp, _vp, _vp, _i64, _i64, _i, _vp, _i, _vp],
    "trb_interp_backward": [_vp, _vp, _vp, _vp, _i64, _i64, _i, _vp, _vp, _i, _vp],
    "trb_vertex_normals_forward": [_vp, _vp, _i64, _i64, _vp, _vp, _i, _vp],
    "trb_vertex_normals_backward": [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i, _vp],
    "trb_shade_forward": [_c.POINTER(ShadeConfig)] + [_vp] * 12 + [_i, _vp],
    "trb_shade_backward": [_c.POINTER(ShadeConfig)] + [_vp] * 20 + [_i, _vp],
    "trb_render_sizes": [_c.POINTER(RenderConfig), _c.POINTER(_sz), _c.POINTER(_i64), _c.POINTER(_i64)],
    "trb_render_forward": [_c.POINTER(RenderConfig)] + [_vp] * 18 + [_sz, _vp, _c.POINTER(UvTexture),
                                                                    _c.POINTER(RenderExtras), _i, _vp],
    "trb_render_backward": [_c.POINTER(RenderConfig)] + [_vp] * 27 + [_c.POINTER(UvTexture), _i, _vp],
    "trb_render_backward_allreduce": [_c.POINTER(RenderConfig)] + [_vp] * 27 + [_c.POINTER(UvTexture),
                                                                                _c.POINTER(PeerSum), _i, _vp],
    "trb_debug_set_events": [_vp, _vp, _vp, _vp],
    "trb_points_raster_forward": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp],
    "trb_points_raster_workspace_bytes": [_i, _i, _i, _i64, _c.POINTER(_sz)],
    "trb_points_raster_forward_binned": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _vp, _sz, _vp, _vp, _vp, _i, _vp],
    "trb_points_raster_backward": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp],
    "trb_points_composite_forward": [_i, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _i, _vp],
    "trb_points_composite_backward": [_i, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _i, _vp],
    "trb_nn_forward": [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp],
    "trb_nn_backward": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp],
    "trb_allreduce_grid": [_i64],
    "trb_allreduce_set_timing": [_vp],
    "trb_allreduce_sum_f32": [_vp, _vp, _i, _vp, _i64, _i, _i, _vp, _vp, _i, _vp],
}

_lib = None
_lock = threading.Lock()


class TrbLibraryError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Loads libtrb.so once.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise TrbLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m torch_renderer_b200.build` "
                "(nvcc, sm_100a).  torch_renderer_b200 has no CPU or eager fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name, None)
            if fn is None:
                raise TrbLibraryError(f"libtrb.so does not export {name}; rebuild it")
            fn.argtypes = argtypes
            fn.restype = _c.c_int
        handle.trb_status_string.argtypes = [_i]
        handle.trb_status_string.restype = _c.c_char_p
        if handle.trb_abi_version() != ABI_VERSION:
            raise TrbLibraryError("libtrb.so ABI version mismatch; rebuild it")
        for which, (name, size) in enumerate((("trb_view", 32), ("trb_shade_config", _c.sizeof(ShadeConfig)),
                                              ("trb_render_config", _c.sizeof(RenderConfig)),
                                              ("trb_uv_texture", _c.sizeof(UvTexture)),
                                              ("trb_peer_sum", _c.sizeof(PeerSum)),
                                              ("trb_render_extras", _c.sizeof(RenderExtras)))):
            if handle.trb_abi_struct_size(which) != size:
                raise TrbLibraryError(f"libtrb.so is stale: sizeof({name}) is {handle.trb_abi_struct_size(which)} in "
                                      f"the library but {size} in the binding; run python -m torch_renderer_b200.build")
        _lib = handle
    return _lib


def declared_symbols():
    return sorted(list(_SIGNATURES) + ["trb_status_string"])


def check(status: int, what: str) -> None:
    """Maps a trb_status to the exception types PyTorch3D raises at the same boundary."""
    if status == TRB_OK:
        return
    msg = lib().trb_status_string(status).decode()
    if status == TRB_ERR_K_TOO_LARGE:
        raise ValueError(f"{what}: faces_per_pixel must be <= {MAX_FACES_PER_PIXEL}")
    if status == TRB_ERR_BAD_ARG:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg} (status {status}, cuda error {lib().trb_last_cuda_error()})")
