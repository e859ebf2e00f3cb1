"""Lights and materials with PyTorch3D's defaults.

Reference usage: ``PointLights(device=..., location=[[0,0,-3]])`` (renderer.py:76,
torch_renderer.py:132, camera_pose_optimizer.py:144), ``lights.location = tensor`` (renderer.py:82-83),
``AmbientLights(device=...)`` (mesh_deformer.py:113); ``DirectionalLights`` and ``Materials`` are
imported by several scripts.  The lighting arithmetic itself lives in the fused CUDA shade kernel
(csrc/shade.cu); these classes only carry the parameters.
"""
from __future__ import annotations

import torch

from .common import Device, TensorProperties


class Materials(TensorProperties):
    def __init__(self, ambient_color=((1, 1, 1),), diffuse_color=((1, 1, 1),), specular_color=((1, 1, 1),),
                 shininess=64, device: Device = "cpu") -> None:
        super().__init__(device=device, diffuse_color=diffuse_color, ambient_color=ambient_color,
                         specular_color=specular_color, shininess=shininess)
        for n in ("ambient_color", "diffuse_color", "specular_color"):
            if getattr(self, n).shape[-1] != 3:
                raise ValueError("Expected %s to have shape (N, 3); got %r" % (n, getattr(self, n).shape))
        if self.shininess.shape != torch.Size([self._N]):
            raise ValueError("shininess should have shape (N); got %r" % repr(self.shininess.shape))


class _Lights(TensorProperties):
    kind = "ambient"

    def _check(self, names):
        for n in names:
            if getattr(self, n).shape[-1] != 3:
                raise ValueError("Expected %s to have shape (N, 3); got %r" % (n, getattr(self, n).shape))


class PointLights(_Lights):
    kind = "point"

    def __init__(self, ambient_color=((0.5, 0.5, 0.5),), diffuse_color=((0.3, 0.3, 0.3),),
                 specular_color=((0.2, 0.2, 0.2),), location=((0, 1, 0),), device: Device = "cpu") -> None:
        super().__init__(device=device, ambient_color=ambient_color, diffuse_color=diffuse_color,
                         specular_color=specular_color, location=location)
        self._check(("ambient_color", "diffuse_color", "specular_color", "location"))


class DirectionalLights(_Lights):
    kind = "directional"

    def __init__(self, ambient_color=((0.5, 0.5, 0.5),), diffuse_color=((0.3, 0.3, 0.3),),
                 specular_color=((0.2, 0.2, 0.2),), direction=((0, 1, 0),), device: Device = "cpu") -> None:
        super().__init__(device=device, ambient_color=ambient_color, diffuse_color=diffuse_color,
                         specular_color=specular_color, direction=direction)
        self._check(("ambient_color", "diffuse_color", "specular_color", "direction"))


class AmbientLights(_Lights):
    """Ambient term only: colour = ambient_color * texel (mesh_deformer.py:113)."""
    kind = "ambient"

    def __init__(self, *, ambient_color=None, device: Device = "cpu") -> None:
        if ambient_color is None:
            ambient_color = ((1.0, 1.0, 1.0),)
        super().__init__(ambient_color=ambient_color, device=device)
        self._check(("ambient_color",))
