"""Camera models and view helpers with the PyTorch3D call surface the reference uses
(SURVEY.md A0-A2, Appendix B): ``FoVPerspectiveCameras`` (camera_pose_optimizer.py:105),
``PerspectiveCameras`` in NDC (mesh_deformer.py:120-124, myrenderer.py:81) and in screen space with
``in_ndc=False`` / a 4x4 ``K`` (torch_renderer.py:67-71, renderer.py:47-69,
batch_rendering_test.py:225-229), ``look_at_view_transform`` (camera_pose_optimizer.py:167,
mesh_deformer.py:119; pinned by gradient.log:1-6).

Conventions: row vectors, ``X_view = X_world @ R + T``; +X left, +Y up, +Z into the screen; NDC
keeps view-space z.  For the CUDA path every camera reduces to ``(R, T, fx, fy, px, py,
perspective)`` in NDC units -- see ``ndc_projection_params``.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F

from .common import Device, TensorProperties, convert_to_tensors_and_broadcast, format_tensor, make_device
from .transforms import Rotate, Transform3d, Translate

_R = torch.eye(3)[None]
_T = torch.zeros(1, 3)


# --------------------------------------------------------------------------------------------
def camera_position_from_spherical_angles(distance, elevation, azimuth, degrees: bool = True,
                                          device: Device = "cpu") -> torch.Tensor:
    dist, elev, azim = convert_to_tensors_and_broadcast(distance, elevation, azimuth, device=device)
    if degrees:
        elev = math.pi / 180.0 * elev
        azim = math.pi / 180.0 * azim
    x = dist * torch.cos(elev) * torch.sin(azim)
    y = dist * torch.sin(elev)
    z = dist * torch.cos(elev) * torch.cos(azim)
    camera_position = torch.stack([x, y, z], dim=1)
    if camera_position.dim() == 0:
        camera_position = camera_position.view(1, -1)
    return camera_position.view(-1, 3)


def look_at_rotation(camera_position, at=((0, 0, 0),), up=((0, 1, 0),), device: Device = "cpu") -> torch.Tensor:
    camera_position, at, up = convert_to_tensors_and_broadcast(camera_position, at, up, device=device)
    for t, n in zip([camera_position, at, up], ["camera_position", "at", "up"]):
        if t.shape[-1] != 3:
            raise ValueError("Expected arg %s to have shape (N, 3); got %r" % (n, t.shape))
    z_axis = F.normalize(at - camera_position, eps=1e-5)
    x_axis = F.normalize(torch.cross(up, z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    is_close = torch.isclose(x_axis, torch.tensor(0.0, device=x_axis.device), atol=5e-3).all(dim=1, keepdim=True)
    if is_close.any():
        replacement = F.normalize(torch.cross(y_axis, z_axis, dim=1), eps=1e-5)
        x_axis = torch.where(is_close, replacement, x_axis)
    R = torch.cat((x_axis[:, None, :], y_axis[:, None, :], z_axis[:, None, :]), dim=1)
    return R.transpose(1, 2)


def look_at_view_transform(dist=1.0, elev=0.0, azim=0.0, degrees: bool = True, eye=None,
                           at=((0, 0, 0),), up=((0, 1, 0),), device: Device = "cpu"):
    """Returns (R [N,3,3], T [N,3]) of a camera looking at ``at`` from spherical angles or ``eye``."""
    if eye is not None:
        eye, at, up = convert_to_tensors_and_broadcast(eye, at, up, device=device)
        C = eye
    else:
        dist, elev, azim, at, up = convert_to_tensors_and_broadcast(dist, elev, azim, at, up, device=device)
        C = camera_position_from_spherical_angles(dist, elev, azim, degrees=degrees, device=device) + at
    R = look_at_rotation(C, at, up, device=device)
    T = -torch.bmm(R.transpose(1, 2), C[:, :, None])[:, :, 0]
    return R, T


def get_world_to_view_transform(R=_R, T=_T) -> Transform3d:
    if T.shape[0] != R.shape[0]:
        raise ValueError("Expected R, T to have the same batch dimension; got %r, %r" % (R.shape[0], T.shape[0]))
    if T.dim() != 2 or T.shape[1:] != (3,):
        raise ValueError("Expected T to have shape (N, 3); got %r" % repr(T.shape))
    if R.dim() != 3 or R.shape[1:] != (3, 3):
        raise ValueError("Expected R to have shape (N, 3, 3); got %r" % repr(R.shape))
    return Rotate(R, device=R.device).compose(Translate(T, device=T.device))


# --------------------------------------------------------------------------------------------
class CamerasBase(TensorProperties):
    """Common behaviour; subclasses define the projection."""

    _FIELDS: Tuple[str, ...] = ()

    def get_projection_transform(self, **kwargs) -> Transform3d:
        raise NotImplementedError()

    def is_perspective(self) -> bool:
        raise NotImplementedError()

    def in_ndc(self) -> bool:
        raise NotImplementedError()

    def get_znear(self):
        return getattr(self, "znear", None)

    def get_image_size(self):
        return getattr(self, "image_size", None)

    def _override(self, name: str, kwargs):
        v = kwargs.get(name, None)
        return getattr(self, name) if v is None else v

    def _rt(self, kwargs):
        """R, T for this call: kwargs override the stored ones (and are remembered, as upstream does)."""
        R = kwargs.get("R", None)
        T = kwargs.get("T", None)
        R = self.R if R is None else R
        T = self.T if T is None else T
        return R, T

    def _store_rt(self, R, T) -> None:
        """Upstream's ``get_world_to_view_transform`` stores per-call overrides on the camera object."""
        if R is not self.R:
            self._set_tensor("R", R)
        if T is not self.T:
            self._set_tensor("T", T)

    def get_world_to_view_transform(self, **kwargs) -> Transform3d:
        R, T = self._rt(kwargs)
        self._store_rt(R, T)
        return get_world_to_view_transform(R=R, T=T)

    def get_camera_center(self, **kwargs) -> torch.Tensor:
        """Camera centre in world coordinates: ``-T @ inv(R)``.  Like upstream (which goes through
        ``get_world_to_view_transform(**kwargs)``), ``R=`` / ``T=`` overrides are remembered on the camera."""
        R, T = self._rt(kwargs)
        self._store_rt(R, T)
        return -torch.matmul(T[:, None, :], torch.linalg.inv_ex(R)[0])[:, 0, :]  # inv_ex: no host sync

    def get_full_projection_transform(self, **kwargs) -> Transform3d:
        w2v = self.get_world_to_view_transform(**kwargs)
        return w2v.compose(self.get_projection_transform(**kwargs))

    def get_ndc_camera_transform(self, **kwargs) -> Transform3d:
        """Transform from the camera's projection space to NDC (identity for NDC cameras)."""
        if self.in_ndc():
            return Transform3d(device=self.device, dtype=torch.float32)
        # screen-space cameras give the principal point in image coordinates (+X right, +Y down)
        # while points live in the +X left, +Y up system: x' = x_proj - 2*px, then screen -> NDC.
        N = max(len(self), 1)
        fix = torch.eye(4, dtype=torch.float32, device=self.device).view(1, 4, 4).repeat(N, 1, 1)
        _, _, px, py = self._focal_pp(kwargs)
        fix[:, 3, 0] = -2.0 * px
        fix[:, 3, 1] = -2.0 * py
        image_size = kwargs.get("image_size", self.get_image_size())
        return Transform3d(matrix=fix).compose(
            _ndc_to_screen_transform(self, with_xyflip=False, image_size=image_size).inverse())

    def transform_points(self, points, eps: Optional[float] = None, **kwargs) -> torch.Tensor:
        return self.get_full_projection_transform(**kwargs).transform_points(points, eps=eps)

    def transform_points_ndc(self, points, eps: Optional[float] = None, **kwargs) -> torch.Tensor:
        t = self.get_full_projection_transform(**kwargs)
        if not self.in_ndc():
            t = t.compose(self.get_ndc_camera_transform(**kwargs))
        return t.transform_points(points, eps=eps)

    def transform_points_screen(self, points, eps: Optional[float] = None, with_xyflip: bool = True,
                                **kwargs) -> torch.Tensor:
        points_ndc = self.transform_points_ndc(points, eps=eps, **kwargs)
        image_size = kwargs.get("image_size", self.get_image_size())
        return _ndc_to_screen_transform(self, with_xyflip=with_xyflip,
                                        image_size=image_size).transform_points(points_ndc, eps=eps)

    def ndc_projection_params(self, **kwargs):
        """(proj f32[N,4] = fx,fy,px,py in NDC units, perspective: bool) for the CUDA transform."""
        raise NotImplementedError()

    def unproject_points(self, xy_depth, world_coordinates: bool = True, **kwargs):
        t = self.get_full_projection_transform(**kwargs) if world_coordinates else self.get_projection_transform(**kwargs)
        return t.inverse().transform_points(xy_depth)


def _image_size_hw(cameras, image_size, N: int, device) -> torch.Tensor:
    if image_size is None:
        raise ValueError("For screen-space cameras image_size=(height, width) is required")
    if not torch.is_tensor(image_size):
        image_size = torch.tensor(image_size, device=device)
    image_size = image_size.to(device=device, dtype=torch.float32)
    if image_size.dim() == 1:
        image_size = image_size[None]
    return image_size.expand(N, 2) if image_size.shape[0] != N else image_size


def _ndc_to_screen_transform(cameras, with_xyflip: bool, image_size) -> Transform3d:
    """x_s = scale*x_ndc - W/2 (then negated when ``with_xyflip``), scale = min(H, W)/2."""
    N = max(len(cameras), 1)
    hw = _image_size_hw(cameras, image_size, N, cameras.device)
    height, width = hw[:, 0], hw[:, 1]
    scale = torch.minimum(height, width) / 2.0
    sgn = -1.0 if with_xyflip else 1.0
    M = torch.zeros((N, 4, 4), device=cameras.device, dtype=torch.float32)  # row-vector form
    M[:, 0, 0] = sgn * scale
    M[:, 1, 1] = sgn * scale
    M[:, 3, 0] = -sgn * width / 2.0
    M[:, 3, 1] = -sgn * height / 2.0
    M[:, 2, 2] = 1.0
    M[:, 3, 3] = 1.0
    return Transform3d(matrix=M)


# --------------------------------------------------------------------------------------------
class FoVPerspectiveCameras(CamerasBase):
    """OpenGL-style perspective camera defined by a field of view; projection lands in NDC."""

    def __init__(self, znear=1.0, zfar=100.0, aspect_ratio=1.0, fov=60.0, degrees: bool = True,
                 R=_R, T=_T, K=None, device: Device = "cpu"):
        super().__init__(device=device, znear=znear, zfar=zfar, aspect_ratio=aspect_ratio, fov=fov,
                         R=R, T=T, K=K)
        self.degrees = degrees

    def is_perspective(self) -> bool:
        return True

    def in_ndc(self) -> bool:
        return True

    def _tan_half_fov(self, fov):
        if self.degrees:
            fov = (math.pi / 180.0) * fov
        return torch.tan(fov / 2.0)

    def compute_projection_matrix(self, znear, zfar, fov, aspect_ratio, degrees: bool) -> torch.Tensor:
        N = max(len(self), 1)
        K = torch.zeros((N, 4, 4), dtype=torch.float32, device=self.device)
        if not torch.is_tensor(fov):
            fov = torch.tensor(fov, device=self.device)
        if degrees:
            fov = (math.pi / 180.0) * fov
        tan_half = torch.tan(fov / 2.0)
        max_y = tan_half * znear
        min_y = -max_y
        max_x = max_y * aspect_ratio
        min_x = -max_x
        z_sign = 1.0
        K[:, 0, 0] = 2.0 * znear / (max_x - min_x)
        K[:, 1, 1] = 2.0 * znear / (max_y - min_y)
        K[:, 0, 2] = (max_x + min_x) / (max_x - min_x)
        K[:, 1, 2] = (max_y + min_y) / (max_y - min_y)
        K[:, 3, 2] = z_sign
        K[:, 2, 2] = z_sign * zfar / (zfar - znear)
        K[:, 2, 3] = -(zfar * znear) / (zfar - znear)
        return K

    def get_projection_transform(self, **kwargs) -> Transform3d:
        K = kwargs.get("K", self.K)
        if K is not None:
            if K.shape != (len(self), 4, 4):
                raise ValueError("Expected K to have shape of (%r, 4, 4)" % len(self))
        else:
            K = self.compute_projection_matrix(self._override("znear", kwargs), self._override("zfar", kwargs),
                                               self._override("fov", kwargs),
                                               self._override("aspect_ratio", kwargs),
                                               kwargs.get("degrees", self.degrees))
        return Transform3d(matrix=K.transpose(1, 2).contiguous())

    def ndc_projection_params(self, **kwargs):
        if kwargs.get("K", self.K) is not None:
            raise NotImplementedError("FoVPerspectiveCameras with an explicit K matrix is not supported "
                                      "by the CUDA transform")
        fov = self._override("fov", kwargs)
        aspect = self._override("aspect_ratio", kwargs)
        tan_half = self._tan_half_fov(fov)
        fy = 1.0 / tan_half
        fx = fy / aspect
        zero = torch.zeros_like(fx)
        return torch.stack([fx, fy, zero, zero], dim=-1), True


class FoVOrthographicCameras(CamerasBase):
    def __init__(self, znear=1.0, zfar=100.0, max_y=1.0, min_y=-1.0, max_x=1.0, min_x=-1.0,
                 scale_xyz=((1.0, 1.0, 1.0),), R=_R, T=_T, K=None, device: Device = "cpu"):
        super().__init__(device=device, znear=znear, zfar=zfar, max_y=max_y, min_y=min_y, max_x=max_x,
                         min_x=min_x, scale_xyz=scale_xyz, R=R, T=T, K=K)

    def is_perspective(self) -> bool:
        return False

    def in_ndc(self) -> bool:
        return True

    def get_projection_transform(self, **kwargs) -> Transform3d:
        N = max(len(self), 1)
        g = lambda n: self._override(n, kwargs)
        znear, zfar, max_x, min_x, max_y, min_y, s = (g("znear"), g("zfar"), g("max_x"), g("min_x"),
                                                      g("max_y"), g("min_y"), g("scale_xyz"))
        K = torch.zeros((N, 4, 4), dtype=torch.float32, device=self.device)
        K[:, 0, 0] = (2.0 / (max_x - min_x)) * s[:, 0]
        K[:, 1, 1] = (2.0 / (max_y - min_y)) * s[:, 1]
        K[:, 0, 3] = -(max_x + min_x) / (max_x - min_x)
        K[:, 1, 3] = -(max_y + min_y) / (max_y - min_y)
        K[:, 3, 3] = 1.0
        K[:, 2, 2] = (1.0 / (zfar - znear)) * s[:, 2]
        K[:, 2, 3] = -znear / (zfar - znear)
        return Transform3d(matrix=K.transpose(1, 2).contiguous())

    def ndc_projection_params(self, **kwargs):
        g = lambda n: self._override(n, kwargs)
        max_x, min_x, max_y, min_y, s = g("max_x"), g("min_x"), g("max_y"), g("min_y"), g("scale_xyz")
        fx = (2.0 / (max_x - min_x)) * s[:, 0]
        fy = (2.0 / (max_y - min_y)) * s[:, 1]
        px = -(max_x + min_x) / (max_x - min_x)
        py = -(max_y + min_y) / (max_y - min_y)
        return torch.stack([fx, fy, px, py], dim=-1), False


class _SfMCameras(CamerasBase):
    """Shared implementation of Perspective/Orthographic cameras (focal length + principal point)."""

    _perspective = True

    def __init__(self, focal_length=1.0, principal_point=((0.0, 0.0),), R=_R, T=_T, K=None,
                 device: Device = "cpu", in_ndc: bool = True, image_size=None):
        kw = {"image_size": image_size} if image_size is not None else {}
        super().__init__(device=device, focal_length=focal_length, principal_point=principal_point,
                         R=R, T=T, K=K, **kw)
        self._in_ndc = in_ndc
        if image_size is not None:
            if (self.image_size < 1).any():
                raise ValueError("Image_size provided has invalid values")
        else:
            self.image_size = None
        if self.focal_length.dim() == 1:  # (N,) -> (N, 1)
            self.focal_length = self.focal_length[:, None]

    def is_perspective(self) -> bool:
        return self._perspective

    def in_ndc(self) -> bool:
        return self._in_ndc

    def _focal_pp(self, kwargs):
        K = kwargs.get("K", self.K)
        if K is not None:
            fx, fy, px, py = K[:, 0, 0], K[:, 1, 1], K[:, 0, 2 if self._perspective else 3], \
                K[:, 1, 2 if self._perspective else 3]
            return fx, fy, px, py
        f = self._override("focal_length", kwargs)
        p = self._override("principal_point", kwargs)
        if not torch.is_tensor(f):
            f = format_tensor(f, device=self.device)
        if not torch.is_tensor(p):
            p = format_tensor(p, device=self.device)
        if f.dim() == 1:
            f = f[:, None]
        fx, fy = (f[:, 0], f[:, 0]) if f.shape[1] == 1 else (f[:, 0], f[:, 1])
        return fx, fy, p[:, 0], p[:, 1]

    def get_projection_transform(self, **kwargs) -> Transform3d:
        K = kwargs.get("K", self.K)
        if K is None:
            fx, fy, px, py = self._focal_pp(kwargs)
            N = max(len(self), fx.shape[0])
            K = torch.zeros((N, 4, 4), dtype=torch.float32, device=self.device)
            K[:, 0, 0], K[:, 1, 1] = fx, fy
            if self._perspective:
                K[:, 0, 2], K[:, 1, 2] = px, py
                K[:, 3, 2] = 1.0
                K[:, 2, 3] = 1.0
            else:
                K[:, 0, 3], K[:, 1, 3] = px, py
                K[:, 2, 2] = 1.0
                K[:, 3, 3] = 1.0
        elif K.shape[-2:] != (4, 4):
            raise ValueError("Expected K to have shape of (N, 4, 4)")
        return Transform3d(matrix=K.transpose(1, 2).contiguous())

    def ndc_projection_params(self, **kwargs):
        fx, fy, px, py = self._focal_pp(kwargs)
        if not self._in_ndc:
            image_size = kwargs.get("image_size", self.image_size)
            hw = _image_size_hw(self, image_size, fx.shape[0] if fx.shape[0] > 1 else max(len(self), 1),
                                self.device)
            h, w = hw[:, 0], hw[:, 1]
            s = torch.minimum(h, w) / 2.0
            fx, fy = fx / s, fy / s
            px = -(px - w / 2.0) / s
            py = -(py - h / 2.0) / s
        n = max(fx.shape[0], px.shape[0])
        ex = lambda t: t.expand(n) if t.shape[0] != n else t
        return torch.stack([ex(fx), ex(fy), ex(px), ex(py)], dim=-1), self._perspective


class PerspectiveCameras(_SfMCameras):
    """Multi-view-geometry perspective camera: ``x = fx*X/Z + px`` (NDC or, with ``in_ndc=False``
    and ``image_size=(H, W)``, pixels)."""
    _perspective = True


class OrthographicCameras(_SfMCameras):
    _perspective = False


# legacy aliases PyTorch3D still exports
OpenGLPerspectiveCameras = FoVPerspectiveCameras
OpenGLOrthographicCameras = FoVOrthographicCameras
SfMPerspectiveCameras = PerspectiveCameras
SfMOrthographicCameras = OrthographicCameras


def cameras_from_opencv_projection(R: torch.Tensor, tvec: torch.Tensor, camera_matrix: torch.Tensor,
                                   image_size: torch.Tensor) -> PerspectiveCameras:
    """OpenCV (R, t, K, image_size=(H,W)) -> NDC ``PerspectiveCameras`` (``pytorch3d.utils`` twin;
    imported by renderer.py:10, torch_renderer.py:10)."""
    focal_length = torch.stack([camera_matrix[:, 0, 0], camera_matrix[:, 1, 1]], dim=-1)
    principal_point = camera_matrix[:, :2, 2]
    image_size_wh = image_size.to(R).flip(dims=(1,))
    scale = image_size_wh.to(R).min(dim=1, keepdim=True)[0] / 2.0
    scale = scale.expand(-1, 2)
    c0 = image_size_wh / 2.0
    focal_pytorch3d = focal_length / scale
    p0_pytorch3d = -(principal_point - c0) / scale
    R_pytorch3d = R.clone().permute(0, 2, 1)
    T_pytorch3d = tvec.clone()
    R_pytorch3d[:, :, :2] *= -1
    T_pytorch3d[:, :2] *= -1
    return PerspectiveCameras(R=R_pytorch3d, T=T_pytorch3d, focal_length=focal_pytorch3d,
                              principal_point=p0_pytorch3d, image_size=image_size, device=R.device)
