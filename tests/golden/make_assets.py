"""Generates tests/golden/meshes.npz from the reference's mesh assets (run in the authoring
container only, where /root/reference exists):

    python tests/golden/make_assets.py

Stores geometry only (float32 vertices, int32 faces, and the cow's UV coordinates) of
data/teapot.obj and data/cow_mesh/cow.obj -- the two assets BASELINE.json's configs name -- so that
the GPU box, which has no /root/reference, can run the same scenes.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from torch_renderer_b200.io import load_obj  # noqa: E402

REF = "/root/reference/data"
out = {}
v, f, _ = load_obj(os.path.join(REF, "teapot.obj"), load_textures=False)
out["teapot_verts"], out["teapot_faces"] = v.numpy().astype(np.float32), f.verts_idx.numpy().astype(np.int32)
v, f, aux = load_obj(os.path.join(REF, "cow_mesh", "cow.obj"), load_textures=False)
out["cow_verts"], out["cow_faces"] = v.numpy().astype(np.float32), f.verts_idx.numpy().astype(np.int32)
out["cow_verts_uvs"] = aux.verts_uvs.numpy().astype(np.float32)
out["cow_faces_uvs"] = f.textures_idx.numpy().astype(np.int32)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "meshes.npz")
np.savez_compressed(dst, **out)
print({k: a.shape for k, a in out.items()}, os.path.getsize(dst))
