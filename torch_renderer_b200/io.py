"""OBJ I/O with the ``pytorch3d.io`` surface the reference uses: ``load_obj`` (camera_pose_optimizer.py:87,
myrenderer.py:66, mesh_deformer.py:12), ``load_objs_as_meshes`` (camera_pose_optimizer.py:102,
renderer.py:106, mesh_deformer.py:92), ``save_obj`` (mesh_deformer.py:376; with a UV texture: deform_mesh_with_color.py:460).  Host-side only.
"""
from __future__ import annotations

import os
from collections import namedtuple
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .structures import Meshes, join_meshes_as_batch
from .textures import TexturesUV

Faces = namedtuple("Faces", "verts_idx normals_idx textures_idx materials_idx")
Properties = namedtuple("Properties", "normals verts_uvs material_colors texture_images texture_atlas")


def _parse_mtl(path: str):
    material_colors: Dict[str, Dict[str, torch.Tensor]] = {}
    texture_files: Dict[str, str] = {}
    name = None
    if not os.path.isfile(path):
        return material_colors, texture_files
    with open(path, "r") as f:
        for line in f:
            tok = line.strip().split()
            if not tok:
                continue
            if tok[0] == "newmtl":
                name = tok[1]
                material_colors[name] = {}
            elif name is None:
                continue
            elif tok[0] == "map_Kd":
                texture_files[name] = line.strip()[len("map_Kd"):].strip()
            elif tok[0] in ("Ka", "Kd", "Ks"):
                key = {"Ka": "ambient_color", "Kd": "diffuse_color", "Ks": "specular_color"}[tok[0]]
                material_colors[name][key] = torch.tensor([float(x) for x in tok[1:4]], dtype=torch.float32)
            elif tok[0] == "Ns":
                material_colors[name]["shininess"] = torch.tensor([float(tok[1])], dtype=torch.float32)
    return material_colors, texture_files


def _load_image(path: str) -> torch.Tensor:
    from PIL import Image
    with Image.open(path) as im:
        arr = np.asarray(im.convert("RGB"), dtype=np.float32) / 255.0
    return torch.from_numpy(arr.copy())


def load_obj(f, load_textures: bool = True, create_texture_atlas: bool = False, texture_atlas_size: int = 4,
             texture_wrap: Optional[str] = "repeat", device="cpu", path_manager=None):
    """Returns (verts (V,3) f32, Faces(verts_idx, normals_idx, textures_idx, materials_idx) i64 (F,3)/(F,),
    Properties(normals, verts_uvs, material_colors, texture_images, texture_atlas))."""
    if create_texture_atlas:
        raise NotImplementedError("texture atlases are not supported")
    path = os.fspath(f)
    data_dir = os.path.dirname(path) or "."
    verts, normals, uvs = [], [], []
    f_v, f_t, f_n, f_m = [], [], [], []
    material_names: List[str] = []
    mtl_path = None
    cur_mat = -1
    with open(path, "r") as fh:
        for line in fh:
            if not line or line[0] == "#":
                continue
            tok = line.split()
            if not tok:
                continue
            t0 = tok[0]
            if t0 == "v":
                verts.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif t0 == "vt":
                uvs.append((float(tok[1]), float(tok[2])))
            elif t0 == "vn":
                normals.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif t0 == "f":
                vi, ti, ni = [], [], []
                for c in tok[1:]:
                    parts = c.split("/")
                    vi.append(int(parts[0]))
                    ti.append(int(parts[1]) if len(parts) > 1 and parts[1] != "" else 0)
                    ni.append(int(parts[2]) if len(parts) > 2 and parts[2] != "" else 0)
                nv, nt, nn_ = len(verts), len(uvs), len(normals)
                fix = lambda i, n: (i - 1) if i > 0 else ((n + i) if i < 0 else -1)
                vi = [fix(i, nv) for i in vi]
                ti = [fix(i, nt) for i in ti]
                ni = [fix(i, nn_) for i in ni]
                for k in range(1, len(vi) - 1):  # fan triangulation
                    f_v.append((vi[0], vi[k], vi[k + 1]))
                    f_t.append((ti[0], ti[k], ti[k + 1]))
                    f_n.append((ni[0], ni[k], ni[k + 1]))
                    f_m.append(cur_mat)
            elif t0 == "mtllib":
                mtl_path = os.path.join(data_dir, line.strip()[len("mtllib"):].strip())
            elif t0 == "usemtl":
                nm = tok[1]
                if nm not in material_names:
                    material_names.append(nm)
                cur_mat = material_names.index(nm)
    dev = torch.device(device)
    verts_t = torch.tensor(verts, dtype=torch.float32, device=dev).reshape(-1, 3)
    normals_t = torch.tensor(normals, dtype=torch.float32, device=dev).reshape(-1, 3) if normals else None
    uvs_t = torch.tensor(uvs, dtype=torch.float32, device=dev).reshape(-1, 2) if uvs else None
    as_idx = lambda x, w: torch.tensor(x, dtype=torch.int64, device=dev).reshape((-1, w) if w else (-1,))
    faces = Faces(verts_idx=as_idx(f_v, 3), normals_idx=as_idx(f_n, 3), textures_idx=as_idx(f_t, 3),
                  materials_idx=as_idx(f_m, 0))
    material_colors, texture_images = None, None
    if load_textures and mtl_path is not None:
        colors, tex_files = _parse_mtl(mtl_path)
        material_colors = {k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in colors.items()}
        texture_images = {}
        for name, rel in tex_files.items():
            p = os.path.join(data_dir, rel)
            if os.path.isfile(p):
                texture_images[name] = _load_image(p).to(dev)
    aux = Properties(normals=normals_t, verts_uvs=uvs_t, material_colors=material_colors,
                     texture_images=texture_images, texture_atlas=None)
    return verts_t, faces, aux


def load_objs_as_meshes(files: Sequence, device=None, load_textures: bool = True,
                        create_texture_atlas: bool = False, texture_atlas_size: int = 4,
                        texture_wrap: Optional[str] = "repeat", path_manager=None) -> Meshes:
    mesh_list = []
    for f_obj in files:
        verts, faces, aux = load_obj(f_obj, load_textures=load_textures)
        tex = None
        if load_textures and aux.texture_images is not None and len(aux.texture_images) > 0 \
                and aux.verts_uvs is not None:
            image = list(aux.texture_images.values())[0]
            tex = TexturesUV(maps=[image.to(device) if device is not None else image],
                             faces_uvs=[faces.textures_idx.to(device) if device is not None else faces.textures_idx],
                             verts_uvs=[aux.verts_uvs.to(device) if device is not None else aux.verts_uvs])
        v = verts.to(device) if device is not None else verts
        fi = faces.verts_idx.to(device) if device is not None else faces.verts_idx
        mesh_list.append(Meshes(verts=[v], faces=[fi], textures=tex))
    if len(mesh_list) == 1:
        return mesh_list[0]
    return join_meshes_as_batch(mesh_list)


def save_obj(f, verts: torch.Tensor, faces: torch.Tensor, decimal_places: Optional[int] = None, path_manager=None, *,
             verts_uvs: Optional[torch.Tensor] = None, faces_uvs: Optional[torch.Tensor] = None,
             texture_map: Optional[torch.Tensor] = None) -> None:
    """Writes ``verts`` (V, 3) / ``faces`` (F, 3) as Wavefront OBJ.  With ``verts_uvs`` (Vt, 2), ``faces_uvs`` (F, 3)
    and ``texture_map`` (H, W, 3) in [0, 1] all given (deform_mesh_with_color.py:460) the texture goes along as
    upstream writes it: ``vt`` lines and ``v/vt`` face corners, ``<stem>.mtl`` naming ``<stem>.png``, and the map
    saved as that PNG (``load_obj`` reads the triple back)."""
    if verts.dim() != 2 or verts.shape[1] != 3:
        raise ValueError("Argument 'verts' should either be empty or of shape (num_verts, 3).")
    if faces.numel() and (faces.dim() != 2 or faces.shape[1] != 3):
        raise ValueError("Argument 'faces' should either be empty or of shape (num_faces, 3).")
    if faces_uvs is not None and (faces_uvs.dim() != 2 or faces_uvs.shape[1] != 3):
        raise ValueError("Argument 'faces_uvs' should either be empty or of shape (num_faces, 3).")
    if verts_uvs is not None and (verts_uvs.dim() != 2 or verts_uvs.shape[1] != 2):
        raise ValueError("Argument 'verts_uvs' should either be empty or of shape (num_verts, 2).")
    if texture_map is not None and (texture_map.dim() != 3 or texture_map.shape[2] != 3):
        raise ValueError("Argument 'texture_map' should either be empty or of shape (H, W, 3).")
    textured = verts_uvs is not None and faces_uvs is not None and texture_map is not None
    if textured and faces_uvs.shape[0] != faces.shape[0]:
        raise ValueError("faces_uvs must have one row per face")
    path = os.fspath(f)
    stem = os.path.splitext(os.path.basename(path))[0]
    fmt = "%f" if decimal_places is None else "%." + str(decimal_places) + "f"
    v = verts.detach().cpu().tolist()
    fc = (faces.detach().cpu() + 1).tolist()
    with open(path, "w") as fh:
        if textured:
            fh.write("\nmtllib %s.mtl\nusemtl mesh\n\n" % stem)
        for x in v:
            fh.write("v " + " ".join(fmt % c for c in x) + "\n")
        if textured:
            for uv in verts_uvs.detach().cpu().tolist():
                fh.write("vt " + " ".join(fmt % c for c in uv) + "\n")
            ft = (faces_uvs.detach().cpu() + 1).tolist()
            for a, b in zip(fc, ft):
                fh.write("f %d/%d %d/%d %d/%d\n" % (a[0], b[0], a[1], b[1], a[2], b[2]))
        else:
            for t in fc:
                fh.write("f %d %d %d\n" % tuple(t))
    if textured:
        from PIL import Image
        base = os.path.splitext(path)[0]
        pixels = (texture_map.detach().cpu().float() * 255.0).clamp(0.0, 255.0).numpy().astype(np.uint8)
        Image.fromarray(pixels).save(base + ".png")
        with open(base + ".mtl", "w") as fh:
            fh.write("newmtl mesh\nmap_Kd %s.png\n# Test colors\nKa 1.000 1.000 1.000\nKd 1.000 1.000 1.000\n"
                     "Ks 0.000 0.000 0.000\nNs 10.0\n" % stem)
