"""``Meshes``: batched triangle-mesh container with the PyTorch3D surface the reference scripts use
(SURVEY.md 8a row a17, Appendix B): list / padded construction, ``verts_packed``, ``faces_packed``,
``verts_padded``, ``extend(N)`` (batch_rendering_test.py:326, mesh_deformer.py:150,235),
``offset_verts`` (pose_optimizer.py:103, mesh_deformer.py:167), ``scale_verts_``, ``update_padded``,
``get_mesh_verts_faces`` (mesh_deformer.py:241), ``.textures =`` (mesh_deformer.py:190), ``clone``,
``to``, ``verts_normals_packed``.

B200-specific behaviour: ``extend(N)`` is *lazy*.  The reference pattern "one mesh, N cameras"
replicates vertices and faces N times upstream (N*V*12 + N*F*24 bytes and an N*F*36-byte face_verts
gather per render); here the N views share one vertex / face array and the kernels get a per-view
table (``trb_view``) instead.  Any accessor that needs the replicated tensors materialises them on
demand, so the container still behaves like the upstream one.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch

from . import ops
from .common import unbind_batch


# Topology-only device data is shared between Meshes objects: optimisation loops build a new Meshes every
# step (mesh_deformer.py:190, `src_mesh.offset_verts(deform)`), and rebuilding the int32 face table and the
# per-view records each time costs a conversion kernel, a host loop and an H2D copy per step.
_FACES_I32_CACHE: "dict[int, tuple]" = {}
_VIEW_TABLE_CACHE: "dict[tuple, object]" = {}


def _faces_i32_cached(src: torch.Tensor) -> torch.Tensor:
    import weakref
    key = id(src)
    hit = _FACES_I32_CACHE.get(key)
    if hit is not None:
        ref, version, out = hit
        if ref() is src and version == src._version:
            return out
    if len(_FACES_I32_CACHE) > 256:
        _FACES_I32_CACHE.clear()
    out = src.to(torch.int32).contiguous()
    _FACES_I32_CACHE[key] = (weakref.ref(src), src._version, out)
    return out


def _view_table_cached(key, build):
    table = _VIEW_TABLE_CACHE.get(key)
    if table is None:
        if len(_VIEW_TABLE_CACHE) > 256:
            _VIEW_TABLE_CACHE.clear()
        table = build()
        _VIEW_TABLE_CACHE[key] = table
    return table


def _list_to_padded(xs: Sequence[torch.Tensor], pad_value, dtype, device, trailing) -> torch.Tensor:
    n = len(xs)
    m = max((x.shape[0] for x in xs), default=0)
    out = torch.full((n, m) + trailing, pad_value, dtype=dtype, device=device)
    for i, x in enumerate(xs):
        if x.shape[0] > 0:
            out[i, : x.shape[0]] = x
    return out


class Meshes:
    def __init__(self, verts=None, faces=None, textures=None, *, verts_normals=None) -> None:
        self.device = torch.device("cpu")
        self.textures = textures
        self._verts_list: Optional[List[torch.Tensor]] = None
        self._faces_list: Optional[List[torch.Tensor]] = None
        self._replicas = 1          # >1: lazily extended single mesh (see module docstring)
        self._reset_caches()

        if isinstance(verts, (list, tuple)) and isinstance(faces, (list, tuple)):
            if len(verts) != len(faces):
                raise ValueError("verts and faces lists must have the same length")
            self._verts_list = [v for v in verts]
            # list input is taken as is (like upstream: only padded tensors carry -1 filler rows);
            # no boolean indexing here -- it would cost a host sync per Meshes construction
            self._faces_list = [(f if f.dtype == torch.int64 else f.to(torch.int64)) if f.numel() > 0 else
                                f.reshape(0, 3).to(torch.int64) for f in faces]
            if len(self._verts_list) > 0:
                self.device = self._verts_list[0].device
                for v, f in zip(self._verts_list, self._faces_list):
                    if v.device != self.device or f.device != self.device:
                        raise ValueError("All Verts and Faces tensors should be on same device.")
        elif torch.is_tensor(verts) and torch.is_tensor(faces):
            if verts.dim() != 3 or verts.shape[2] != 3:
                raise ValueError("Verts tensor has incorrect dimensions.")
            if faces.dim() != 3 or faces.shape[2] != 3:
                raise ValueError("Faces tensor has incorrect dimensions.")
            if verts.shape[0] != faces.shape[0]:
                raise ValueError("verts and faces must have the same batch size")
            self.device = verts.device
            if faces.device != self.device:
                raise ValueError("Verts and Faces tensors should be on same device.")
            self._verts_list = unbind_batch(verts)
            self._faces_list = []
            for i in range(faces.shape[0]):
                f = faces[i]
                self._faces_list.append(f[(f >= 0).all(dim=1)].to(torch.int64))
        else:
            raise ValueError("Verts and Faces must be either a list or a tensor with "
                             "shape (batch_size, N, 3) where N is either the maximum number of "
                             "verts or faces respectively.")
        self._N = len(self._verts_list)
        if textures is not None:
            self._check_textures(textures)
        if verts_normals is not None:
            raise NotImplementedError("explicit verts_normals are not supported; normals are "
                                      "recomputed from geometry as PyTorch3D's shaders do")

    # ------------------------------------------------------------------ bookkeeping
    def _reset_caches(self) -> None:
        self._verts_packed = None
        self._faces_packed = None
        self._faces_i32 = None
        self._verts_normals_packed = None
        self._verts_normals_key = None
        self._view_table = None
        self._num_verts = None
        self._num_faces = None

    def _check_textures(self, textures) -> None:
        n = getattr(textures, "_N", None)
        if n is not None and n != len(self):
            raise ValueError("Textures do not match the dimensions of Meshes.")

    def __len__(self) -> int:
        return self._N * self._replicas

    def isempty(self) -> bool:
        return len(self) == 0 or all(v.shape[0] == 0 for v in self._verts_list)

    def _materialize(self) -> None:
        """Turns a lazily extended mesh into an ordinary batch (replicated lists)."""
        if self._replicas > 1:
            r = self._replicas
            self._verts_list = [v for v in self._verts_list for _ in range(r)]
            self._faces_list = [f for f in self._faces_list for _ in range(r)]
            self._N *= r
            self._replicas = 1
            self._reset_caches()

    @property
    def is_shared_replica(self) -> bool:
        return self._replicas > 1

    # ------------------------------------------------------------------ accessors
    def verts_list(self) -> List[torch.Tensor]:
        self._materialize()
        return self._verts_list

    def faces_list(self) -> List[torch.Tensor]:
        self._materialize()
        return self._faces_list

    def num_verts_per_mesh(self) -> torch.Tensor:
        self._materialize()
        return torch.tensor([v.shape[0] for v in self._verts_list], dtype=torch.int64, device=self.device)

    def num_faces_per_mesh(self) -> torch.Tensor:
        self._materialize()
        return torch.tensor([f.shape[0] for f in self._faces_list], dtype=torch.int64, device=self.device)

    def mesh_to_verts_packed_first_idx(self) -> torch.Tensor:
        n = self.num_verts_per_mesh()
        return torch.cumsum(n, 0) - n

    def mesh_to_faces_packed_first_idx(self) -> torch.Tensor:
        n = self.num_faces_per_mesh()
        return torch.cumsum(n, 0) - n

    def verts_packed(self) -> torch.Tensor:
        self._materialize()
        if self._verts_packed is None:
            self._verts_packed = (torch.cat(self._verts_list, dim=0) if self._N > 0 else
                                  torch.zeros((0, 3), dtype=torch.float32, device=self.device))
        return self._verts_packed

    def faces_packed(self) -> torch.Tensor:
        self._materialize()
        if self._faces_packed is None:
            off, parts = 0, []
            for v, f in zip(self._verts_list, self._faces_list):
                parts.append(f + off)
                off += v.shape[0]
            self._faces_packed = (torch.cat(parts, dim=0) if parts else
                                  torch.zeros((0, 3), dtype=torch.int64, device=self.device))
        return self._faces_packed

    def verts_padded(self) -> torch.Tensor:
        self._materialize()
        if self._N > 0 and all(v.shape[0] == self._verts_list[0].shape[0] for v in self._verts_list):
            return torch.stack(self._verts_list, dim=0)
        return _list_to_padded(self._verts_list, 0.0, torch.float32, self.device, (3,))

    def faces_padded(self) -> torch.Tensor:
        self._materialize()
        return _list_to_padded(self._faces_list, -1, torch.int64, self.device, (3,))

    def get_mesh_verts_faces(self, index: int):
        if not isinstance(index, int):
            raise ValueError("Mesh index must be an integer.")
        if index < 0 or index >= len(self):
            raise ValueError("Mesh index must be in the range [0, N) where N is the number of meshes.")
        if self._replicas > 1:
            return self._verts_list[0], self._faces_list[0]
        return self._verts_list[index], self._faces_list[index]

    def __getitem__(self, index) -> "Meshes":
        self._materialize()
        if isinstance(index, int):
            idx = [index]
        elif isinstance(index, slice):
            idx = list(range(len(self)))[index]
        elif isinstance(index, (list, tuple)):
            idx = list(index)
        elif torch.is_tensor(index):
            idx = index.nonzero().flatten().tolist() if index.dtype == torch.bool else index.tolist()
        else:
            raise IndexError(index)
        tex = None if self.textures is None else self.textures[idx]
        return Meshes([self._verts_list[i] for i in idx], [self._faces_list[i] for i in idx], textures=tex)

    def get_bounding_boxes(self) -> torch.Tensor:
        boxes = []
        for v in self.verts_list():
            boxes.append(torch.stack([v.min(dim=0)[0], v.max(dim=0)[0]], dim=1))
        return torch.stack(boxes, dim=0)

    # ------------------------------------------------------------------ geometry derived data
    def faces_packed_i32(self) -> torch.Tensor:
        """int32 copy of the (unique) face table the kernels index; cached because topology is static."""
        if self._faces_i32 is None:
            src = self._faces_list[0] if (self._replicas > 1 or self._N == 1) else self.faces_packed()
            self._faces_i32 = _faces_i32_cached(src)
        return self._faces_i32

    def _unique_verts(self) -> torch.Tensor:
        """World-space vertices without replication: [V,3] for a shared replica, packed otherwise."""
        return self._verts_list[0] if self._replicas > 1 else self.verts_packed()

    def _unique_verts_normals(self) -> torch.Tensor:
        verts = self._unique_verts()
        if not verts.is_cuda:
            raise RuntimeError("vertex normals are computed by the CUDA extension; move the mesh "
                               "to a CUDA device (no CPU fallback)")
        if verts.requires_grad:
            # a cached result would carry an autograd graph that the previous backward() freed
            return ops.vertex_normals(verts, self.faces_packed_i32())
        key = (verts.data_ptr(), verts._version)
        if self._verts_normals_packed is None or self._verts_normals_key != key:
            self._verts_normals_packed = ops.vertex_normals(verts, self.faces_packed_i32())
            self._verts_normals_key = key
        return self._verts_normals_packed

    def verts_normals_packed(self) -> torch.Tensor:
        n = self._unique_verts_normals()
        return n.repeat(self._replicas, 1) if self._replicas > 1 else n

    def verts_normals_list(self) -> List[torch.Tensor]:
        n = self.verts_normals_packed()
        sizes = [v.shape[0] for v in self.verts_list()]
        return list(n.split(sizes, 0))

    def verts_normals_padded(self) -> torch.Tensor:
        return _list_to_padded(self.verts_normals_list(), 0.0, torch.float32, self.device, (3,))

    def view_table(self) -> "ops.ViewTable":
        """Per-view ``trb_view`` records for the CUDA kernels (cached; depends on topology only)."""
        if self._view_table is None:
            if self._replicas > 1:
                V, F, r = self._verts_list[0].shape[0], self._faces_list[0].shape[0], self._replicas
                self._view_table = _view_table_cached(
                    (str(self.device), "shared", V, F, r),
                    lambda: ops.ViewTable.build(
                        face_start=[0] * r, face_count=[F] * r, p2f_base=[i * F for i in range(r)],
                        world_vert_start=[0] * r, vert_count=[V] * r, device=self.device, shared_mesh=True))
            else:
                nv = [v.shape[0] for v in self._verts_list]
                nf = [f.shape[0] for f in self._faces_list]

                def build():
                    vs = [sum(nv[:i]) for i in range(len(nv))]
                    fs = [sum(nf[:i]) for i in range(len(nf))]
                    return ops.ViewTable.build(face_start=fs, face_count=nf, p2f_base=fs, world_vert_start=vs,
                                               vert_count=nv, device=self.device, shared_mesh=False)
                self._view_table = _view_table_cached((str(self.device), "packed", tuple(nv), tuple(nf)), build)
        return self._view_table

    # ------------------------------------------------------------------ construction of new meshes
    def _new_like(self, verts_list, textures="same") -> "Meshes":
        other = Meshes.__new__(Meshes)
        other.device = self.device
        other.textures = self.textures if textures == "same" else textures
        other._verts_list = verts_list
        other._faces_list = self._faces_list
        other._replicas = self._replicas
        other._N = self._N
        other._reset_caches()
        # topology-only caches can be shared
        other._faces_i32 = self._faces_i32
        other._view_table = self._view_table
        other._faces_packed = self._faces_packed
        return other

    def clone(self) -> "Meshes":
        other = self._new_like([v.clone() for v in self._verts_list],
                               textures=None if self.textures is None else self.textures.clone())
        other._faces_list = [f.clone() for f in self._faces_list]
        other._faces_i32 = other._faces_packed = other._view_table = None
        return other

    def detach(self) -> "Meshes":
        return self._new_like([v.detach() for v in self._verts_list],
                              textures=None if self.textures is None else self.textures.detach())

    def to(self, device, copy: bool = False) -> "Meshes":
        device = torch.device(device) if not isinstance(device, torch.device) else device
        if not copy and self.device == device:
            return self
        other = Meshes.__new__(Meshes)
        other.device = device
        other.textures = None if self.textures is None else self.textures.to(device)
        other._verts_list = [v.to(device) for v in self._verts_list]
        other._faces_list = [f.to(device) for f in self._faces_list]
        other._replicas = self._replicas
        other._N = self._N
        other._reset_caches()
        return other

    def cpu(self) -> "Meshes":
        return self.to("cpu")

    def cuda(self, device=None) -> "Meshes":
        return self.to(torch.device("cuda" if device is None else f"cuda:{device}"))

    def extend(self, N: int) -> "Meshes":
        """N copies of every mesh in the batch.  For a single mesh the copies share storage."""
        if not isinstance(N, int):
            raise ValueError("N must be an integer.")
        if N <= 0:
            raise ValueError("N must be > 0.")
        tex = None if self.textures is None else self.textures.extend(N)
        if self._N == 1:
            other = self._new_like(self._verts_list, textures=tex)
            other._replicas = self._replicas * N
            other._view_table = None
            other._faces_packed = None
            return other
        self._materialize()
        verts = [v for v in self._verts_list for _ in range(N)]
        faces = [f for f in self._faces_list for _ in range(N)]
        return Meshes(verts, faces, textures=tex)

    def offset_verts(self, vert_offsets_packed: torch.Tensor) -> "Meshes":
        """Out of place: adds a (sum V, 3) or (3,) offset to the vertices."""
        if vert_offsets_packed.shape == (3,):
            return self._new_like([v + vert_offsets_packed for v in self._verts_list])
        total_unique = sum(v.shape[0] for v in self._verts_list)
        if self._replicas > 1 and vert_offsets_packed.shape[0] != total_unique:
            self._materialize()
        expected = sum(v.shape[0] for v in self._verts_list)
        if vert_offsets_packed.shape != (expected, 3):
            raise ValueError("Verts offset must have dimension (all_v, 3).")
        out, cur = [], 0
        for v in self._verts_list:
            out.append(v + vert_offsets_packed[cur: cur + v.shape[0]])
            cur += v.shape[0]
        return self._new_like(out)

    def offset_verts_(self, vert_offsets_packed: torch.Tensor) -> "Meshes":
        new = self.offset_verts(vert_offsets_packed)
        self._verts_list = new._verts_list
        self._verts_packed = None
        self._verts_normals_packed = None
        return self

    def scale_verts(self, scale) -> "Meshes":
        if not torch.is_tensor(scale):
            scale = torch.full((self._N,), float(scale), device=self.device)
        if scale.numel() == 1:
            scale = scale.reshape(1).expand(self._N)
        if scale.shape[0] != self._N:
            self._materialize()
        return self._new_like([v * scale[i] for i, v in enumerate(self._verts_list)])

    def scale_verts_(self, scale) -> "Meshes":
        new = self.scale_verts(scale)
        self._verts_list = new._verts_list
        self._verts_packed = None
        self._verts_normals_packed = None
        return self

    def update_padded(self, new_verts_padded: torch.Tensor) -> "Meshes":
        """New Meshes with the same topology and textures but different vertex positions."""
        self._materialize()
        if new_verts_padded.dim() != 3 or new_verts_padded.shape[0] != self._N or new_verts_padded.shape[2] != 3:
            raise ValueError("new values must have the same batch dimension / be of shape (N, V, 3).")
        verts = [new_verts_padded[i, : v.shape[0]] for i, v in enumerate(self._verts_list)]
        return self._new_like(verts)

    def sample_textures(self, fragments) -> torch.Tensor:
        if self.textures is None:
            raise ValueError("Meshes does not have textures")
        return self.textures.sample_textures(fragments, faces_packed=self.faces_packed())

    # losses in the reference's deformation loops need these
    def edges_packed(self) -> torch.Tensor:
        f = self.faces_packed()
        e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], dim=0)
        e, _ = e.sort(dim=1)
        return torch.unique(e, dim=0)


def join_meshes_as_batch(meshes: Sequence[Meshes], include_textures: bool = True) -> Meshes:
    if isinstance(meshes, Meshes):
        raise ValueError("Wrong first argument to join_meshes_as_batch.")
    verts = [v for m in meshes for v in m.verts_list()]
    faces = [f for m in meshes for f in m.faces_list()]
    tex = None
    if include_textures and all(m.textures is not None for m in meshes) and len(meshes) > 0:
        tex = meshes[0].textures.join_batch([m.textures for m in meshes[1:]])
    return Meshes(verts, faces, textures=tex)


from .pointclouds import Pointclouds, join_pointclouds_as_batch  # noqa: E402,F401
