"""Point-cloud rendering with the PyTorch3D call surface: ``PointsRasterizationSettings``, ``PointFragments``,
``PointsRasterizer``, ``rasterize_points``, ``AlphaCompositor``, ``NormWeightedCompositor``, ``alpha_composite``,
``norm_weighted_sum`` and ``PointsRenderer`` (last item).  Reference usage:
``AlphaPointRender`` / ``NormPointRender`` (torch_renderer.py:163-208) build
``PointsRenderer(rasterizer=PointsRasterizer(cameras, PointsRasterizationSettings(image_size, radius,
points_per_pixel)), compositor=AlphaCompositor(background_color))`` and call it with ``R=, T=``.

Host side only; the arithmetic is in csrc/points_render.cu (rasteriser forward / backward, both compositors forward /
backward) and the camera transform is the meshes' (csrc/transform.cu).  Semantics: oracle/points_render_ref.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import NamedTuple, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import ops
from .pointclouds import Pointclouds
from .rasterizer import _cached_projection, _check_bin_size, _expand_views, _parse_image_size


class PointFragments(NamedTuple):
    """idx i32 (N,H,W,K) index into the packed points or -1; zbuf f32; dists f32 squared NDC distance; -1 fill."""
    idx: torch.Tensor
    zbuf: torch.Tensor
    dists: torch.Tensor


@dataclass
class PointsRasterizationSettings:
    image_size: Union[int, Tuple[int, int]] = 256
    radius: Union[float, torch.Tensor] = 0.01
    points_per_pixel: int = 8
    bin_size: Optional[int] = None
    max_points_per_bin: Optional[int] = None


def _packed_radius(radius, pointclouds: Pointclouds) -> torch.Tensor:
    """float, (N, P_max) padded tensor or (P,) packed tensor -> f32 (P,) on the clouds' device."""
    P = pointclouds.points_packed().shape[0]
    dev = pointclouds.device
    if not torch.is_tensor(radius):
        return torch.full((P,), float(radius), dtype=torch.float32, device=dev)
    radius = radius.to(device=dev, dtype=torch.float32)
    if radius.dim() == 0:
        return radius.expand(P).contiguous()
    if radius.dim() == 1 and radius.shape[0] == P:
        return radius
    n = [p.shape[0] for p in pointclouds.points_list()]
    if radius.dim() == 2 and radius.shape[0] == len(n) and radius.shape[1] >= max(n, default=0):
        return torch.cat([radius[i, :k] for i, k in enumerate(n)])
    raise ValueError("radius must be of shape (N, P): got %s" % repr(tuple(radius.shape)))


def rasterize_points(pointclouds: Pointclouds, image_size=256, radius=0.01, points_per_pixel: int = 8,
                     bin_size: Optional[int] = None, max_points_per_bin: Optional[int] = None):
    """PyTorch3D ``rasterize_points`` twin: ``pointclouds`` hold points already in NDC (x, y) + view z.
    Returns (idx i32 (N,H,W,K), zbuf, dists2)."""
    H, W = _parse_image_size(image_size)
    _check_bin_size(bin_size, H, W)
    return ops.rasterize_points_ndc(pointclouds.points_packed(), _packed_radius(radius, pointclouds),
                                    pointclouds.view_table(), (H, W), points_per_pixel,
                                    max_radius=None if torch.is_tensor(radius) else float(radius))


class PointsRasterizer(nn.Module):
    def __init__(self, cameras=None, raster_settings: Optional[PointsRasterizationSettings] = None) -> None:
        super().__init__()
        self.cameras = cameras
        self.raster_settings = raster_settings if raster_settings is not None else PointsRasterizationSettings()

    def to(self, device):
        if self.cameras is not None:
            self.cameras = self.cameras.to(device)
        return self

    def transform(self, point_clouds: Pointclouds, **kwargs) -> torch.Tensor:
        """World -> NDC (x, y) + view z of every point: f32 (P, 3), packed."""
        cameras = kwargs.get("cameras", self.cameras)
        if cameras is None:
            raise ValueError("Cameras must be specified either at initialization or in the forward pass of "
                             "PointsRasterizer")
        N, dev = len(point_clouds), point_clouds.device
        R = kwargs.get("R", None)
        T = kwargs.get("T", None)
        R = cameras.R if R is None else R
        T = cameras.T if T is None else T
        cameras.__dict__["R"], cameras.__dict__["T"] = R, T
        proj_kwargs = {k: v for k, v in kwargs.items() if k not in ("R", "T", "cameras", "raster_settings")}
        proj, perspective = _cached_projection(cameras, proj_kwargs)
        R = _expand_views(R.to(dev), N, "R")
        T = _expand_views(T.to(dev), N, "T")
        proj = _expand_views(proj.to(dev), N, "projection")
        return ops.transform_verts(point_clouds.points_packed(), R, T, proj, point_clouds.view_table(), perspective)

    def forward(self, point_clouds: Pointclouds, **kwargs) -> PointFragments:
        settings = kwargs.get("raster_settings", self.raster_settings)
        H, W = _parse_image_size(settings.image_size)
        _check_bin_size(settings.bin_size, H, W)
        points_ndc = self.transform(point_clouds, **kwargs)
        idx, zbuf, dists = ops.rasterize_points_ndc(
            points_ndc, _packed_radius(settings.radius, point_clouds), point_clouds.view_table(), (H, W),
            settings.points_per_pixel,
            max_radius=None if torch.is_tensor(settings.radius) else float(settings.radius))
        return PointFragments(idx=idx, zbuf=zbuf, dists=dists)


# ------------------------------------------------------------------------------------------ compositors
def _composite_nkhw(fragments, alphas, ptclds, mode: int, background=None) -> torch.Tensor:
    """Upstream layouts: fragments / alphas (N, K, H, W), ptclds (C, P) -> (N, C, H, W).  The kernels are
    channels-last; the permutes are views when the inputs came from ``PointsRenderer`` (which permuted them from
    channels-last in the first place)."""
    if fragments.dim() != 4 or alphas.shape != fragments.shape:
        raise ValueError("fragments and alphas must both have shape (N, points_per_pixel, H, W)")
    if ptclds.dim() != 2:
        raise ValueError("ptclds must have shape (C, P)")
    images = ops.composite(fragments.permute(0, 2, 3, 1), alphas.permute(0, 2, 3, 1), ptclds.t(), mode, background)
    return images.permute(0, 3, 1, 2)


def alpha_composite(pointsidx, alphas, pt_clds) -> torch.Tensor:
    """``pytorch3d.renderer.compositing.alpha_composite``: (N,K,H,W), (N,K,H,W), (C,P) -> (N,C,H,W)."""
    return _composite_nkhw(pointsidx, alphas, pt_clds, ops.COMPOSITE_ALPHA)


def norm_weighted_sum(pointsidx, alphas, pt_clds) -> torch.Tensor:
    """``pytorch3d.renderer.compositing.norm_weighted_sum``: same layouts."""
    return _composite_nkhw(pointsidx, alphas, pt_clds, ops.COMPOSITE_NORM_WEIGHTED)


def _background_tensor(background_color, C: int, device) -> Optional[torch.Tensor]:
    """``_add_background_color_to_images``: a scalar is broadcast, a (C-1) colour gets alpha = 1."""
    if background_color is None:
        return None
    bg = background_color if torch.is_tensor(background_color) else torch.tensor(background_color, dtype=torch.float32)
    bg = bg.to(device=device, dtype=torch.float32)
    if bg.dim() == 0:
        bg = bg.expand(C)
    if bg.dim() > 1:
        raise ValueError("Wrong shape of background_color")
    if bg.shape[0] + 1 == C:
        bg = torch.cat([bg, bg.new_ones(1)])
    if bg.shape[0] != C:
        raise ValueError("Background color has %s channels not %s" % (bg.shape[0], C))
    return bg.contiguous()


class _Compositor(nn.Module):
    mode = ops.COMPOSITE_ALPHA

    def __init__(self, background_color=None) -> None:
        super().__init__()
        self.background_color = background_color

    def forward(self, fragments, alphas, ptclds, **kwargs) -> torch.Tensor:
        background_color = kwargs.get("background_color", self.background_color)
        bg = _background_tensor(background_color, ptclds.shape[0], alphas.device)
        return _composite_nkhw(fragments, alphas, ptclds, self.mode, bg)


class AlphaCompositor(_Compositor):
    """Front-to-back alpha compositing of the K nearest points of every pixel."""
    mode = ops.COMPOSITE_ALPHA


class NormWeightedCompositor(_Compositor):
    """Weighted mean of the K nearest points of every pixel (weights normalised to sum to 1)."""
    mode = ops.COMPOSITE_NORM_WEIGHTED


class PointsRenderer(nn.Module):
    """``images = compositor(fragments.idx, 1 - dists / r^2, features)``, returned channels-last (N, H, W, C)."""

    def __init__(self, rasterizer, compositor) -> None:
        super().__init__()
        self.rasterizer = rasterizer
        self.compositor = compositor

    def to(self, device):
        self.rasterizer = self.rasterizer.to(device)
        self.compositor = self.compositor.to(device)
        return self

    def forward(self, point_clouds: Pointclouds, **kwargs) -> torch.Tensor:
        fragments = self.rasterizer(point_clouds, **kwargs)
        r = self.rasterizer.raster_settings.radius
        dists2 = fragments.dists.permute(0, 3, 1, 2)
        weights = 1 - dists2 / (r * r)
        features = point_clouds.features_packed()
        if features is None:
            raise ValueError("Pointclouds must carry features to be rendered")
        # (upstream widens idx to int64 here; the kernels read the rasteriser's int32 directly)
        images = self.compositor(fragments.idx.permute(0, 3, 1, 2), weights, features.permute(1, 0), **kwargs)
        return images.permute(0, 2, 3, 1)
