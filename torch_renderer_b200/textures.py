"""``TexturesVertex`` and ``TexturesUV``.

Reference usage: per-vertex RGB (myrenderer.py:77, camera_pose_optimizer.py:92-93,
mesh_deformer.py:190 -- optimised with ``requires_grad``), UV maps loaded by
``load_objs_as_meshes`` for the cow (camera_pose_optimizer.py:102) and optimised in
deform_mesh_with_color.py:266-271,329.

The Phong shaders consume ``TexturesVertex`` inside the fused CUDA shade kernel (the per-vertex
colours are interpolated there, no (N,H,W,K,3) texel tensor exists).  ``TexturesUV`` produces texels
with the CUDA interpolation kernel + ``grid_sample`` and hands them to the same shade kernel.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch
import torch.nn.functional as F

from . import ops
from .common import unbind_batch


def _to_list(x, n_expected=None):
    if torch.is_tensor(x):
        return unbind_batch(x)
    return list(x)


class TexturesVertex:
    def __init__(self, verts_features) -> None:
        """verts_features: list of (V_i, C) tensors or a padded (N, V, C) tensor."""
        if isinstance(verts_features, (list, tuple)):
            self._feats = [f for f in verts_features]
        elif torch.is_tensor(verts_features):
            if verts_features.dim() != 3:
                raise ValueError("Expected verts_features to be of shape (N, V, C)")
            self._feats = unbind_batch(verts_features)
        else:
            raise ValueError("verts_features must be a tensor or list of tensors")
        self._N = len(self._feats)
        self._replicas = 1
        self.device = self._feats[0].device if self._N > 0 else torch.device("cpu")

    def __len__(self):
        return self._N * self._replicas

    def _clone_with(self, feats) -> "TexturesVertex":
        other = TexturesVertex.__new__(TexturesVertex)
        other._feats, other._N, other._replicas = feats, len(feats), self._replicas
        other.device = feats[0].device if feats else self.device
        return other

    def clone(self):
        return self._clone_with([f.clone() for f in self._feats])

    def detach(self):
        return self._clone_with([f.detach() for f in self._feats])

    def to(self, device):
        return self._clone_with([f.to(device) for f in self._feats])

    def extend(self, N: int) -> "TexturesVertex":
        if self._N == 1:
            other = self._clone_with(self._feats)
            other._replicas = self._replicas * N
            return other
        self._materialize()
        return TexturesVertex([f for f in self._feats for _ in range(N)])

    def _materialize(self):
        if self._replicas > 1:
            self._feats = [f for f in self._feats for _ in range(self._replicas)]
            self._N, self._replicas = len(self._feats), 1

    def __getitem__(self, index):
        self._materialize()
        idx = [index] if isinstance(index, int) else list(index)
        return TexturesVertex([self._feats[i] for i in idx])

    def verts_features_list(self) -> List[torch.Tensor]:
        self._materialize()
        return self._feats

    def verts_features_packed(self) -> torch.Tensor:
        self._materialize()
        if len(self._feats) == 1:
            return self._feats[0]          # one mesh: the packed tensor IS its feature tensor (no copy kernel)
        return torch.cat(self._feats, dim=0)

    def verts_features_padded(self) -> torch.Tensor:
        self._materialize()
        return torch.stack(self._feats, dim=0)

    def _unique_features(self, shared: bool) -> torch.Tensor:
        """Colours indexed like ``Meshes._unique_verts``."""
        if shared and self._replicas > 1:
            return self._feats[0]
        return self.verts_features_packed()

    def join_batch(self, textures: Sequence["TexturesVertex"]) -> "TexturesVertex":
        feats = list(self.verts_features_list())
        for t in textures:
            feats += t.verts_features_list()
        return TexturesVertex(feats)

    def sample_textures(self, fragments, faces_packed=None) -> torch.Tensor:
        """(N,H,W,K,C) texels by barycentric interpolation of the per-vertex features."""
        feats = self.verts_features_packed()
        faces_feats = feats[faces_packed]
        return ops.interpolate_face_attributes(fragments.pix_to_face, fragments.bary_coords, faces_feats)


class TexturesUV:
    def __init__(self, maps, faces_uvs, verts_uvs, padding_mode: str = "border",
                 align_corners: bool = True, sampling_mode: str = "bilinear") -> None:
        """maps (N,Ht,Wt,C) or list; faces_uvs (N,F,3) i64 or list; verts_uvs (N,Vt,2) or list."""
        self._maps = _to_list(maps)
        self._faces_uvs = _to_list(faces_uvs)
        self._verts_uvs = _to_list(verts_uvs)
        if not (len(self._maps) == len(self._faces_uvs) == len(self._verts_uvs)):
            raise ValueError("maps, faces_uvs and verts_uvs must have the same batch dimension")
        self._N = len(self._maps)
        self.padding_mode, self.align_corners, self.sampling_mode = padding_mode, align_corners, sampling_mode
        self.device = self._maps[0].device if self._N > 0 else torch.device("cpu")

    def __len__(self):
        return self._N

    def _clone_with(self, maps, faces_uvs, verts_uvs) -> "TexturesUV":
        return TexturesUV(maps, faces_uvs, verts_uvs, self.padding_mode, self.align_corners, self.sampling_mode)

    def clone(self):
        return self._clone_with([m.clone() for m in self._maps], [f.clone() for f in self._faces_uvs],
                                [v.clone() for v in self._verts_uvs])

    def detach(self):
        return self._clone_with([m.detach() for m in self._maps], [f.detach() for f in self._faces_uvs],
                                [v.detach() for v in self._verts_uvs])

    def to(self, device):
        return self._clone_with([m.to(device) for m in self._maps], [f.to(device) for f in self._faces_uvs],
                                [v.to(device) for v in self._verts_uvs])

    def extend(self, N: int) -> "TexturesUV":
        return self._clone_with([m for m in self._maps for _ in range(N)],
                                [f for f in self._faces_uvs for _ in range(N)],
                                [v for v in self._verts_uvs for _ in range(N)])

    def __getitem__(self, index):
        idx = [index] if isinstance(index, int) else list(index)
        return self._clone_with([self._maps[i] for i in idx], [self._faces_uvs[i] for i in idx],
                                [self._verts_uvs[i] for i in idx])

    def maps_padded(self) -> torch.Tensor:
        return torch.stack(self._maps, dim=0)

    def maps_list(self):
        return self._maps

    def faces_uvs_list(self):
        return self._faces_uvs

    def verts_uvs_list(self):
        return self._verts_uvs

    def faces_uvs_padded(self):
        return torch.stack(self._faces_uvs, dim=0)

    def verts_uvs_padded(self):
        return torch.stack(self._verts_uvs, dim=0)

    def join_batch(self, textures: Sequence["TexturesUV"]) -> "TexturesUV":
        maps, fu, vu = list(self._maps), list(self._faces_uvs), list(self._verts_uvs)
        for t in textures:
            maps += t._maps; fu += t._faces_uvs; vu += t._verts_uvs
        return self._clone_with(maps, fu, vu)

    def _fused_inputs(self, table, faces_i32):
        """(map f32 [Ht,Wt,3], (verts_uvs f32 [Vt,2], faces_uvs i32 [F,3])) when the fused kernels can sample this
        texture themselves -- one mesh, or one mesh extended to N views, with a single RGB map, bilinear /
        align_corners / border (the defaults the reference uses) -- else None (-> sample_textures + texels)."""
        if (self.sampling_mode != "bilinear" or not self.align_corners or self.padding_mode != "border"
                or self._N == 0):
            return None
        m0, f0, v0 = self._maps[0], self._faces_uvs[0], self._verts_uvs[0]
        if self._N > 1 and not (table.shared_mesh and all(m is m0 for m in self._maps)
                                and all(f is f0 for f in self._faces_uvs) and all(v is v0 for v in self._verts_uvs)):
            return None
        if self._N != table.N and self._N != 1:
            return None
        if self._N == 1 and table.N != 1:
            return None
        if m0.dim() != 3 or m0.shape[-1] != 3 or not m0.is_cuda or f0.shape[0] != faces_i32.shape[0]:
            return None
        cache = self.__dict__.get("_fused_cache")
        key = (id(f0), id(v0), v0._version)
        if cache is None or cache[0] != key:
            cache = (key, f0.to(torch.int32).contiguous(), v0.detach().float().contiguous())
            self.__dict__["_fused_cache"] = cache
        if v0.requires_grad:
            return None   # gradients w.r.t. the UV coordinates go through the composed path
        return m0, (cache[2], cache[1])

    def sample_textures(self, fragments, faces_packed=None) -> torch.Tensor:
        """(N,H,W,K,C): interpolate UVs (CUDA kernel), then bilinear lookup in the y-flipped map
        with ``align_corners=True`` and border padding (A7)."""
        packing = [v[fu] for v, fu in zip(self._verts_uvs, self._faces_uvs)]
        faces_verts_uvs = torch.cat(packing, dim=0)  # (sum F, 3, 2)
        pixel_uvs = ops.interpolate_face_attributes(fragments.pix_to_face, fragments.bary_coords,
                                                    faces_verts_uvs)  # (N,H,W,K,2)
        N, H_out, W_out, K = fragments.pix_to_face.shape
        maps = self.maps_padded()
        if maps.shape[0] != N:
            raise ValueError("texture batch does not match the fragments batch")
        _, H_in, W_in, C = maps.shape
        pixel_uvs = pixel_uvs.permute(0, 3, 1, 2, 4).reshape(N * K, H_out, W_out, 2)
        texture_maps = maps.permute(0, 3, 1, 2)[None, ...].expand(K, -1, -1, -1, -1).transpose(0, 1)
        texture_maps = texture_maps.reshape(N * K, C, H_in, W_in)
        pixel_uvs = pixel_uvs * 2.0 - 1.0
        texture_maps = torch.flip(texture_maps, [2])
        if texture_maps.device != pixel_uvs.device:
            texture_maps = texture_maps.to(pixel_uvs.device)
        texels = F.grid_sample(texture_maps, pixel_uvs, mode=self.sampling_mode,
                               align_corners=self.align_corners, padding_mode=self.padding_mode)
        return texels.reshape(N, K, C, H_out, W_out).permute(0, 3, 4, 1, 2)
