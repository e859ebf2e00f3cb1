"""``Pointclouds``: a batch of point clouds with optional per-point normals and features, with the part of
``pytorch3d.structures.Pointclouds`` the point renderers need (reference usage:
torch_renderer.py:181,204 ``isinstance(points, Pointclouds)``, pytorch3d_icp_registeration.py:93,142,169
``Pointclouds(list)``, ``.points_padded()``).  Host-side container only: list / packed / padded views of the same
tensors; the arithmetic of rendering lives in csrc/points_render.cu."""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch

from . import ops
from .common import unbind_batch

_Input = Union[Sequence[torch.Tensor], torch.Tensor, None]


def _as_list(x: _Input, what: str, trailing: Optional[int] = None) -> Optional[List[torch.Tensor]]:
    if x is None:
        return None
    if torch.is_tensor(x):
        if x.dim() != 3:
            raise ValueError(f"{what} tensor has incorrect dimensions.")
        x = unbind_batch(x)
    out = []
    for t in x:
        if t.dim() != 2 or (trailing is not None and t.shape[1] != trailing):
            raise ValueError(f"Clouds in list must be of shape Px{trailing or 'C'} or empty")
        out.append(t)
    return out


class Pointclouds:
    def __init__(self, points: _Input, normals: _Input = None, features: _Input = None) -> None:
        pts = _as_list(points, "Points", 3)
        if pts is None:
            raise ValueError("Points must be either a list or a tensor with shape (batch_size, P, 3)")
        self._points_list = pts
        self._normals_list = _as_list(normals, "Normals", 3)
        self._features_list = _as_list(features, "Features")
        self._N = len(pts)
        self.device = pts[0].device if pts else torch.device("cpu")
        for name, aux in (("normals", self._normals_list), ("features", self._features_list)):
            if aux is None:
                continue
            if len(aux) != self._N or any(a.shape[0] != p.shape[0] for a, p in zip(aux, pts)):
                raise ValueError(f"Points and {name} must have the same number of clouds and points per cloud")
        if any(p.device != self.device for p in pts):
            raise ValueError("All points must be on the same device")
        self._packed = {}
        self._view_table = None

    # ------------------------------------------------------------------ sizes
    def __len__(self) -> int:
        return self._N

    def isempty(self) -> bool:
        return self._N == 0 or all(p.shape[0] == 0 for p in self._points_list)

    def num_points_per_cloud(self) -> torch.Tensor:
        return torch.tensor([p.shape[0] for p in self._points_list], dtype=torch.int64, device=self.device)

    def cloud_to_packed_first_idx(self) -> torch.Tensor:
        n = [p.shape[0] for p in self._points_list]
        return torch.tensor([sum(n[:i]) for i in range(len(n))], dtype=torch.int64, device=self.device)

    def packed_to_cloud_idx(self) -> torch.Tensor:
        n = self.num_points_per_cloud()
        return torch.repeat_interleave(torch.arange(self._N, device=self.device), n)

    # ------------------------------------------------------------------ views of the data
    def points_list(self) -> List[torch.Tensor]:
        return list(self._points_list)

    def normals_list(self) -> Optional[List[torch.Tensor]]:
        return None if self._normals_list is None else list(self._normals_list)

    def features_list(self) -> Optional[List[torch.Tensor]]:
        return None if self._features_list is None else list(self._features_list)

    def _pack(self, key: str, lst: Optional[List[torch.Tensor]]) -> Optional[torch.Tensor]:
        if lst is None:
            return None
        if key not in self._packed:
            if len(lst) == 1:
                self._packed[key] = lst[0]
            elif len(lst) == 0:
                self._packed[key] = torch.zeros((0, 3), dtype=torch.float32, device=self.device)
            else:
                self._packed[key] = torch.cat(lst, dim=0)
        return self._packed[key]

    def points_packed(self) -> torch.Tensor:
        return self._pack("points", self._points_list)

    def normals_packed(self) -> Optional[torch.Tensor]:
        return self._pack("normals", self._normals_list)

    def features_packed(self) -> Optional[torch.Tensor]:
        return self._pack("features", self._features_list)

    @staticmethod
    def _pad(lst: Optional[List[torch.Tensor]]) -> Optional[torch.Tensor]:
        if lst is None:
            return None
        P = max((t.shape[0] for t in lst), default=0)
        if all(t.shape[0] == P for t in lst) and lst:
            return torch.stack(lst)
        out = lst[0].new_zeros((len(lst), P, lst[0].shape[1])) if lst else torch.zeros((0, 0, 3))
        for i, t in enumerate(lst):
            out[i, : t.shape[0]] = t
        return out

    def points_padded(self) -> torch.Tensor:
        return self._pad(self._points_list)

    def normals_padded(self) -> Optional[torch.Tensor]:
        return self._pad(self._normals_list)

    def features_padded(self) -> Optional[torch.Tensor]:
        return self._pad(self._features_list)

    def get_cloud(self, index: int):
        if not isinstance(index, int):
            raise ValueError("Cloud index must be an integer.")
        if index < 0 or index >= self._N:
            raise ValueError("Cloud index must be in the range [0, N) where N is the number of clouds in the batch.")
        return (self._points_list[index], None if self._normals_list is None else self._normals_list[index],
                None if self._features_list is None else self._features_list[index])

    def __getitem__(self, index) -> "Pointclouds":
        if isinstance(index, int):
            index = [index]
        elif isinstance(index, slice):
            index = list(range(self._N))[index]
        elif torch.is_tensor(index):
            index = index.nonzero().flatten().tolist() if index.dtype == torch.bool else index.tolist()
        pick = lambda lst: None if lst is None else [lst[i] for i in index]
        return Pointclouds(pick(self._points_list), pick(self._normals_list), pick(self._features_list))

    # ------------------------------------------------------------------ new clouds
    def _map(self, fn) -> "Pointclouds":
        app = lambda lst: None if lst is None else [fn(t) for t in lst]
        return Pointclouds(app(self._points_list), app(self._normals_list), app(self._features_list))

    def clone(self) -> "Pointclouds":
        return self._map(lambda t: t.clone())

    def detach(self) -> "Pointclouds":
        return self._map(lambda t: t.detach())

    def to(self, device, copy: bool = False) -> "Pointclouds":
        device = torch.device(device) if not isinstance(device, torch.device) else device
        if device == self.device and not copy:
            return self
        return self._map(lambda t: t.to(device))

    def cpu(self) -> "Pointclouds":
        return self.to("cpu")

    def cuda(self, device=None) -> "Pointclouds":
        return self.to(torch.device("cuda" if device is None else f"cuda:{device}"))

    def extend(self, N: int) -> "Pointclouds":
        if not isinstance(N, int):
            raise ValueError("N must be an integer.")
        if N <= 0:
            raise ValueError("N must be > 0.")
        rep = lambda lst: None if lst is None else [t for t in lst for _ in range(N)]
        return Pointclouds(rep(self._points_list), rep(self._normals_list), rep(self._features_list))

    def offset(self, offsets_packed: torch.Tensor) -> "Pointclouds":
        pts = self.points_packed()
        if offsets_packed.shape == (3,):
            offsets_packed = offsets_packed.expand_as(pts)
        if offsets_packed.shape != pts.shape:
            raise ValueError("Offsets must have dimension (all_p, 3).")
        sizes = [p.shape[0] for p in self._points_list]
        return Pointclouds(list((pts + offsets_packed).split(sizes)), self._normals_list, self._features_list)

    def scale(self, scale) -> "Pointclouds":
        if not torch.is_tensor(scale):
            scale = torch.full((self._N,), float(scale), device=self.device)
        if scale.shape != (self._N,):
            raise ValueError("New scale must have dimension (num_clouds,).")
        return Pointclouds([p * scale[i] for i, p in enumerate(self._points_list)], self._normals_list,
                           self._features_list)

    def update_padded(self, new_points_padded, new_normals_padded=None, new_features_padded=None) -> "Pointclouds":
        sizes = [p.shape[0] for p in self._points_list]
        cut = lambda padded, old: old if padded is None else [padded[i, :n] for i, n in enumerate(sizes)]
        if new_points_padded.shape[0] != self._N or new_points_padded.shape[2] != 3:
            raise ValueError("new values must have the same batch dimension and 3 coordinates")
        return Pointclouds(cut(new_points_padded, None), cut(new_normals_padded, self._normals_list),
                           cut(new_features_padded, self._features_list))

    def get_bounding_boxes(self) -> torch.Tensor:
        return torch.stack([torch.stack([p.min(0)[0], p.max(0)[0]], dim=1) for p in self._points_list])

    # ------------------------------------------------------------------ kernels' view of the batch
    def view_table(self) -> "ops.ViewTable":
        """One ``trb_view`` per cloud: its range in the packed points (for the transform and the rasteriser)."""
        if self._view_table is None:
            n = [p.shape[0] for p in self._points_list]
            first = [sum(n[:i]) for i in range(len(n))]
            self._view_table = ops.ViewTable.build(face_start=first, face_count=n, p2f_base=first,
                                                   world_vert_start=first, vert_count=n, device=self.device,
                                                   shared_mesh=False)
        return self._view_table


def join_pointclouds_as_batch(pointclouds: Sequence[Pointclouds]) -> Pointclouds:
    if isinstance(pointclouds, Pointclouds) or not pointclouds:
        raise ValueError("Wrong first argument to join_points_as_batch.")
    cat = lambda name: (None if any(getattr(p, name) is None for p in pointclouds)
                        else [t for p in pointclouds for t in getattr(p, name)])
    return Pointclouds(cat("_points_list"), cat("_normals_list"), cat("_features_list"))
