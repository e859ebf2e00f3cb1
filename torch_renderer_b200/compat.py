"""Lets the reference scripts keep their ``from pytorch3d... import ...`` lines.

``install()`` registers alias modules ``pytorch3d``, ``pytorch3d.renderer``, ``pytorch3d.structures``,
``pytorch3d.io``, ``pytorch3d.transforms``, ``pytorch3d.utils``, ``pytorch3d.ops`` and ``pytorch3d.loss`` in
``sys.modules`` that resolve to this package (reference imports: renderer.py:7-26, torch_renderer.py:8-36,
camera_pose_optimizer.py:13-43, mesh_deformer.py:12-40, myrenderer.py:36-49).  Names the package does
not cover (``PulsarPointsRenderer``, Gouraud / flat shaders, ICP, ``knn_points``) resolve to stubs that raise ``NotImplementedError`` when *used*, so
that module-level imports of the scripts still succeed.
"""
from __future__ import annotations

import sys
import types

_OUT_OF_SCOPE = {
    "pytorch3d.renderer": ["PulsarPointsRenderer", "TexturesAtlas", "SoftGouraudShader", "HardGouraudShader",
                           "HardFlatShader"],
    "pytorch3d.structures": [],
    "pytorch3d.ops": ["iterative_closest_point", "knn_points"],
    "pytorch3d.io": ["IO"],
}


def _stub(qualname: str):
    class _OutOfScope:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{qualname} is outside the mesh-rendering hot path this package builds")

    _OutOfScope.__name__ = qualname.rsplit(".", 1)[-1]

    def fn(*a, **k):
        raise NotImplementedError(f"{qualname} is outside the mesh-rendering hot path this package builds")

    fn.__name__ = _OutOfScope.__name__
    return _OutOfScope if _OutOfScope.__name__[0].isupper() else fn


def install(force: bool = False) -> None:
    if "pytorch3d" in sys.modules and not force and not getattr(sys.modules["pytorch3d"], "_trb_alias", False):
        raise RuntimeError("a real pytorch3d is already imported; pass force=True to shadow it")
    import torch_renderer_b200 as trb
    from torch_renderer_b200 import io, loss, ops, renderer, structures, transforms, utils

    def alias(name, src, extra=None):
        mod = types.ModuleType(name)
        mod._trb_alias = True
        for k in dir(src):
            if not k.startswith("_"):
                setattr(mod, k, getattr(src, k))
        for k in _OUT_OF_SCOPE.get(name, []):
            if not hasattr(mod, k):
                setattr(mod, k, _stub(f"{name}.{k}"))
        for k, v in (extra or {}).items():
            setattr(mod, k, v)
        sys.modules[name] = mod
        return mod

    root = alias("pytorch3d", trb, {"__version__": "0.7.x (torch_renderer_b200 alias)"})
    root.renderer = alias("pytorch3d.renderer", renderer)
    root.structures = alias("pytorch3d.structures", structures)
    root.io = alias("pytorch3d.io", io)
    root.transforms = alias("pytorch3d.transforms", transforms)
    root.utils = alias("pytorch3d.utils", utils)
    root.ops = alias("pytorch3d.ops", types.SimpleNamespace(
        interpolate_face_attributes=ops.interpolate_face_attributes,
        sample_points_from_meshes=loss.sample_points_from_meshes))
    root.loss = alias("pytorch3d.loss", types.SimpleNamespace(
        chamfer_distance=loss.chamfer_distance, mesh_edge_loss=loss.mesh_edge_loss,
        mesh_laplacian_smoothing=loss.mesh_laplacian_smoothing, mesh_normal_consistency=loss.mesh_normal_consistency))
    # sub-modules some scripts import from directly
    sys.modules["pytorch3d.renderer.mesh"] = root.renderer
    sys.modules["pytorch3d.renderer.cameras"] = root.renderer
    from torch_renderer_b200 import clip
    root.renderer.mesh = root.renderer
    root.renderer.clip = sys.modules["pytorch3d.renderer.mesh.clip"] = alias("pytorch3d.renderer.mesh.clip", clip)
