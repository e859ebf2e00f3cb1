"""Rotation conversions and a small ``Transform3d`` (row-vector convention, points @ M).

These are the ``pytorch3d.transforms`` names the reference imports: ``quaternion_to_matrix``,
``matrix_to_quaternion``, ``quaternion_apply`` (torch_renderer.py:32-36, camera_pose_optimizer.py:18-24,
where the pose is stored as ``[T(3), quaternion(4)]`` real part first), ``Rotate``, ``Translate``,
``axis_angle_to_matrix`` (myrenderer.py:42,98).  ``matrix_to_quaternion`` is pinned by the
reference's own log: gradient.log:1-6.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


_Q2M = {}


def _q2m_table(dtype, device):
    """[16, 9] constant: column c holds the +-1 weights of the products q_a q_b (row 4a+b) in entry c of the
    rotation matrix divided by two_s (the diagonal's leading 1 is added separately)."""
    key = (dtype, str(device))
    t = _Q2M.get(key)
    if t is None:
        c = torch.zeros(16, 9, dtype=dtype)
        r, i, j, k = 0, 1, 2, 3
        for col, terms in enumerate((
                ((j, j, -1), (k, k, -1)), ((i, j, 1), (k, r, -1)), ((i, k, 1), (j, r, 1)),
                ((i, j, 1), (k, r, 1)), ((i, i, -1), (k, k, -1)), ((j, k, 1), (i, r, -1)),
                ((i, k, 1), (j, r, -1)), ((j, k, 1), (i, r, 1)), ((i, i, -1), (j, j, -1)))):
            for a, b, sgn in terms:
                c[4 * a + b, col] = sgn
        eye = torch.eye(3, dtype=dtype).reshape(9)
        t = (c.to(device), eye.to(device))
        _Q2M[key] = t
    return t


def quaternion_to_matrix(quaternions: torch.Tensor) -> torch.Tensor:
    """(..., 4) real-first quaternions -> (..., 3, 3) rotation matrices (need not be unit).

    Same polynomial as ``pytorch3d.transforms.quaternion_to_matrix`` (R = I + two_s * L(q q^T)), written as one
    outer product and one [16, 9] contraction: 7 small kernels forward instead of ~40 (and ~80 fewer in the
    backward) -- the pose -> matrix conversion was 60% of the host time of a camera_pose_optimizer.py step."""
    table, eye = _q2m_table(quaternions.dtype, quaternions.device)
    two_s = 2.0 / (quaternions * quaternions).sum(-1, keepdim=True)
    qq = (quaternions[..., :, None] * quaternions[..., None, :]).reshape(quaternions.shape[:-1] + (16,))
    o = eye + two_s * (qq @ table)
    return o.reshape(quaternions.shape[:-1] + (3, 3))


def _sqrt_positive_part(x: torch.Tensor) -> torch.Tensor:
    ret = torch.zeros_like(x)
    positive = x > 0
    ret[positive] = torch.sqrt(x[positive])
    return ret


def matrix_to_quaternion(matrix: torch.Tensor) -> torch.Tensor:
    """(..., 3, 3) -> (..., 4) real-first quaternions (numerically stable branch selection)."""
    if matrix.size(-1) != 3 or matrix.size(-2) != 3:
        raise ValueError(f"Invalid rotation matrix shape {matrix.shape}.")
    batch_dim = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(matrix.reshape(batch_dim + (9,)), dim=-1)
    q_abs = _sqrt_positive_part(
        torch.stack(
            [1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22],
            dim=-1,
        )
    )
    quat_by_rijk = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1),
        ],
        dim=-2,
    )
    flr = torch.tensor(0.1).to(dtype=q_abs.dtype, device=q_abs.device)
    quat_candidates = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    best = F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5
    return quat_candidates[best, :].reshape(batch_dim + (4,))


def standardize_quaternion(quaternions: torch.Tensor) -> torch.Tensor:
    return torch.where(quaternions[..., 0:1] < 0, -quaternions, quaternions)


def quaternion_raw_multiply(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    ow = aw * bw - ax * bx - ay * by - az * bz
    ox = aw * bx + ax * bw + ay * bz - az * by
    oy = aw * by - ax * bz + ay * bw + az * bx
    oz = aw * bz + ax * by - ay * bx + az * bw
    return torch.stack((ow, ox, oy, oz), -1)


def quaternion_multiply(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return standardize_quaternion(quaternion_raw_multiply(a, b))


def quaternion_invert(quaternion: torch.Tensor) -> torch.Tensor:
    scaling = torch.tensor([1, -1, -1, -1], device=quaternion.device)
    return quaternion * scaling


def quaternion_apply(quaternion: torch.Tensor, point: torch.Tensor) -> torch.Tensor:
    """Rotate (..., 3) points by (..., 4) quaternions."""
    if point.size(-1) != 3:
        raise ValueError(f"Points are not in 3D, {point.shape}.")
    real_parts = point.new_zeros(point.shape[:-1] + (1,))
    point_as_quaternion = torch.cat((real_parts, point), -1)
    out = quaternion_raw_multiply(
        quaternion_raw_multiply(quaternion, point_as_quaternion), quaternion_invert(quaternion)
    )
    return out[..., 1:]


def axis_angle_to_quaternion(axis_angle: torch.Tensor) -> torch.Tensor:
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half_angles = angles * 0.5
    eps = 1e-6
    small_angles = angles.abs() < eps
    sin_half_angles_over_angles = torch.empty_like(angles)
    sin_half_angles_over_angles[~small_angles] = torch.sin(half_angles[~small_angles]) / angles[~small_angles]
    # sin(x/2)/x ~ 1/2 - x^2/48 near 0
    sin_half_angles_over_angles[small_angles] = 0.5 - (angles[small_angles] * angles[small_angles]) / 48
    return torch.cat([torch.cos(half_angles), axis_angle * sin_half_angles_over_angles], dim=-1)


def axis_angle_to_matrix(axis_angle: torch.Tensor) -> torch.Tensor:
    return quaternion_to_matrix(axis_angle_to_quaternion(axis_angle))


def quaternion_to_axis_angle(quaternions: torch.Tensor) -> torch.Tensor:
    norms = torch.norm(quaternions[..., 1:], p=2, dim=-1, keepdim=True)
    half_angles = torch.atan2(norms, quaternions[..., :1])
    angles = 2 * half_angles
    eps = 1e-6
    small_angles = angles.abs() < eps
    sin_half_angles_over_angles = torch.empty_like(angles)
    sin_half_angles_over_angles[~small_angles] = torch.sin(half_angles[~small_angles]) / angles[~small_angles]
    sin_half_angles_over_angles[small_angles] = 0.5 - (angles[small_angles] * angles[small_angles]) / 48
    return quaternions[..., 1:] / sin_half_angles_over_angles


def matrix_to_axis_angle(matrix: torch.Tensor) -> torch.Tensor:
    return quaternion_to_axis_angle(matrix_to_quaternion(matrix))


def _axis_angle_rotation(axis: str, angle: torch.Tensor) -> torch.Tensor:
    cos, sin = torch.cos(angle), torch.sin(angle)
    one, zero = torch.ones_like(angle), torch.zeros_like(angle)
    if axis == "X":
        flat = (one, zero, zero, zero, cos, -sin, zero, sin, cos)
    elif axis == "Y":
        flat = (cos, zero, sin, zero, one, zero, -sin, zero, cos)
    elif axis == "Z":
        flat = (cos, -sin, zero, sin, cos, zero, zero, zero, one)
    else:
        raise ValueError("letter must be either X, Y or Z.")
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(euler_angles: torch.Tensor, convention: str) -> torch.Tensor:
    if euler_angles.dim() == 0 or euler_angles.shape[-1] != 3:
        raise ValueError("Invalid input euler angles.")
    if len(convention) != 3:
        raise ValueError("Convention must have 3 letters.")
    matrices = [_axis_angle_rotation(c, e) for c, e in zip(convention, torch.unbind(euler_angles, -1))]
    return torch.matmul(torch.matmul(matrices[0], matrices[1]), matrices[2])


# --------------------------------------------------------------------------------------------
class Transform3d:
    """Batch of 4x4 affine/projective transforms applied to row vectors: ``p' = [p, 1] @ M``."""

    def __init__(self, dtype=torch.float32, device="cpu", matrix: Optional[torch.Tensor] = None):
        if matrix is None:
            self._matrix = torch.eye(4, dtype=dtype, device=device).view(1, 4, 4)
        else:
            if matrix.ndim not in (2, 3) or matrix.shape[-2:] != (4, 4):
                raise ValueError('"matrix" has to be a tensor of shape (minibatch, 4, 4) or (4, 4).')
            self._matrix = matrix.view(-1, 4, 4)
        self._transforms = []
        self.device = self._matrix.device
        self.dtype = self._matrix.dtype

    def __len__(self) -> int:
        return self.get_matrix().shape[0]

    def compose(self, *others: "Transform3d") -> "Transform3d":
        out = Transform3d(dtype=self.dtype, device=self.device)
        out._matrix = self._matrix
        out._transforms = self._transforms + list(others)
        return out

    def get_matrix(self) -> torch.Tensor:
        m = self._matrix
        for other in self._transforms:
            m = torch.matmul(m, other.get_matrix())  # broadcasts batch 1 vs N
        return m

    def inverse(self) -> "Transform3d":
        return Transform3d(matrix=torch.linalg.inv(self.get_matrix()))

    def stack(self, *others: "Transform3d") -> "Transform3d":
        mats = [self.get_matrix()] + [o.get_matrix() for o in others]
        return Transform3d(matrix=torch.cat(mats, dim=0))

    def transform_points(self, points: torch.Tensor, eps: Optional[float] = None) -> torch.Tensor:
        pts = points if points.dim() == 3 else points[None]
        if pts.dim() != 3:
            raise ValueError("Expected points to have dim = 2 or dim = 3: got shape %r" % (points.shape,))
        ones = torch.ones(pts.shape[:2] + (1,), dtype=pts.dtype, device=pts.device)
        out = torch.matmul(torch.cat([pts, ones], dim=2), self.get_matrix())
        denom = out[..., 3:]
        if eps is not None:
            sign = denom.sign() + (denom == 0.0).type_as(denom)
            denom = sign * torch.clamp(denom.abs(), eps)
        out = out[..., :3] / denom
        if out.shape[0] == 1 and points.dim() == 2:
            out = out.reshape(points.shape)
        return out

    def transform_normals(self, normals: torch.Tensor) -> torch.Tensor:
        nrm = normals if normals.dim() == 3 else normals[None]
        mat = self.get_matrix()[:, :3, :3]
        out = torch.matmul(nrm, mat.transpose(1, 2).inverse())
        if out.shape[0] == 1 and normals.dim() == 2:
            out = out.reshape(normals.shape)
        return out

    def translate(self, *args, **kwargs) -> "Transform3d":
        return self.compose(Translate(*args, device=self.device, dtype=self.dtype, **kwargs))

    def rotate(self, *args, **kwargs) -> "Transform3d":
        return self.compose(Rotate(*args, device=self.device, dtype=self.dtype, **kwargs))

    def scale(self, *args, **kwargs) -> "Transform3d":
        return self.compose(Scale(*args, device=self.device, dtype=self.dtype, **kwargs))

    def clone(self) -> "Transform3d":
        return Transform3d(matrix=self.get_matrix().clone())

    def to(self, device, copy: bool = False, dtype=None) -> "Transform3d":
        return Transform3d(matrix=self.get_matrix().to(device=device, dtype=dtype or self.dtype))

    def cpu(self):
        return self.to("cpu")

    def cuda(self, idx=None):
        return self.to(torch.device("cuda" if idx is None else f"cuda:{idx}"))


def _xyz_to_tensor(x, y, z, dtype, device) -> torch.Tensor:
    if torch.is_tensor(x) and x.dim() == 2 and y is None and z is None:
        if x.shape[1] != 3:
            raise ValueError("Expected tensor of shape (N, 3); got %r" % (x.shape,))
        return x.to(device=device, dtype=dtype)
    if y is None and z is None:
        y = z = x
    from .common import convert_to_tensors_and_broadcast
    xyz = convert_to_tensors_and_broadcast(x, y, z, dtype=dtype, device=device)
    return torch.stack([t.reshape(-1) for t in xyz], dim=1)


class Translate(Transform3d):
    def __init__(self, x, y=None, z=None, dtype=torch.float32, device=None):
        if device is None:
            device = x.device if torch.is_tensor(x) else "cpu"
        xyz = _xyz_to_tensor(x, y, z, dtype, device)
        N = xyz.shape[0]
        mat = torch.eye(4, dtype=dtype, device=xyz.device).view(1, 4, 4).repeat(N, 1, 1)
        mat[:, 3, :3] = xyz
        super().__init__(matrix=mat)


class Scale(Transform3d):
    def __init__(self, x, y=None, z=None, dtype=torch.float32, device=None):
        if device is None:
            device = x.device if torch.is_tensor(x) else "cpu"
        xyz = _xyz_to_tensor(x, y, z, dtype, device)
        N = xyz.shape[0]
        mat = torch.eye(4, dtype=dtype, device=xyz.device).view(1, 4, 4).repeat(N, 1, 1)
        mat[:, 0, 0], mat[:, 1, 1], mat[:, 2, 2] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
        super().__init__(matrix=mat)


class Rotate(Transform3d):
    def __init__(self, R: torch.Tensor, dtype=torch.float32, device=None, orthogonal_tol: float = 1e-5):
        device = R.device if device is None else device
        if R.dim() == 2:
            R = R[None]
        if R.shape[-2:] != (3, 3):
            raise ValueError("R must have shape (3, 3) or (N, 3, 3); got %s" % repr(R.shape))
        R = R.to(device=device, dtype=dtype)
        N = R.shape[0]
        mat = torch.eye(4, dtype=dtype, device=R.device).view(1, 4, 4).repeat(N, 1, 1)
        mat[:, :3, :3] = R
        super().__init__(matrix=mat)
