"""CPU tests: the C-ABI library loads and exports every symbol include/trb.h declares, the host-side
containers / settings behave like the PyTorch3D surface the reference uses, and the product path
fails loudly without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import torch_renderer_b200 as trb
from torch_renderer_b200 import _lib
from helpers import load_mesh, uv_sphere

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "trb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 14
    if not os.path.exists(_lib.LIB_PATH):
        from torch_renderer_b200 import build
        build.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), f"libtrb.so does not export {s}"
    # the Python binding covers the same set
    assert set(syms) == set(_lib.declared_symbols())
    L = _lib.lib()
    assert L.trb_abi_version() == 5
    assert L.trb_status_string(2).decode().startswith("faces_per_pixel")


def test_workspace_query_and_argument_errors_without_gpu():
    L = _lib.lib()
    n = ctypes.c_size_t(0)
    assert L.trb_raster_workspace_bytes(2, 64, 64, 1, 1000, ctypes.byref(n)) == _lib.TRB_OK
    assert n.value >= 1000 * 4 + 3 * 2 * 16 * 4
    assert L.trb_raster_workspace_bytes(2, 64, 64, 151, 1000, ctypes.byref(n)) == _lib.TRB_ERR_K_TOO_LARGE
    assert L.trb_raster_workspace_bytes(2, 0, 64, 1, 1000, ctypes.byref(n)) == _lib.TRB_ERR_BAD_ARG
    with pytest.raises(ValueError):
        _lib.check(_lib.TRB_ERR_K_TOO_LARGE, "x")
    with pytest.raises(ValueError):
        _lib.check(_lib.TRB_ERR_BAD_ARG, "x")
    with pytest.raises(RuntimeError):
        _lib.check(_lib.TRB_ERR_WORKSPACE, "x")


def test_product_path_has_no_cpu_fallback():
    v, f = uv_sphere(4, 6)
    meshes = trb.Meshes(verts=[v], faces=[f], textures=trb.TexturesVertex(torch.ones(1, v.shape[0], 3)))
    cameras = trb.FoVPerspectiveCameras()
    rasterizer = trb.MeshRasterizer(cameras=cameras, raster_settings=trb.RasterizationSettings(image_size=16))
    with pytest.raises(RuntimeError, match="CUDA"):
        rasterizer(meshes)
    with pytest.raises(RuntimeError, match="CUDA"):
        meshes.verts_normals_packed()
    # and nothing in the package imports, loads or links the oracle
    pkg = os.path.join(ROOT, "torch_renderer_b200")
    pat = re.compile(r"^\s*(import|from)\s+oracle\b|libtrb_oracle|trb_oracle_|#include\s+\"[^\"]*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, fn)).read()), fn


def test_meshes_container():
    v, f = uv_sphere(4, 6)
    v2, f2 = uv_sphere(3, 5)
    m = trb.Meshes(verts=[v, v2], faces=[f, f2])
    assert len(m) == 2
    assert m.verts_packed().shape == (v.shape[0] + v2.shape[0], 3)
    assert torch.equal(m.faces_packed()[f.shape[0]:], f2 + v.shape[0])
    assert m.mesh_to_faces_packed_first_idx().tolist() == [0, f.shape[0]]
    assert m.num_verts_per_mesh().tolist() == [v.shape[0], v2.shape[0]]
    assert m.verts_padded().shape == (2, v.shape[0], 3)
    assert (m.faces_padded()[1, f2.shape[0]:] == -1).all()
    t = m.view_table()
    assert t.N == 2 and not t.shared_mesh and t.max_face_count == f.shape[0]
    assert t.host[1, 0] == f.shape[0] and t.host[1, 3] == f.shape[0] and t.host[1, 6] == v.shape[0]
    # lazy extend shares storage but behaves like a replicated batch
    single = trb.Meshes(verts=[v], faces=[f], textures=trb.TexturesVertex(torch.rand(1, v.shape[0], 3)))
    e = single.extend(5)
    assert len(e) == 5 and e.is_shared_replica and len(e.textures) == 5
    te = e.view_table()
    assert te.shared_mesh and te.N == 5 and te.total_ndc_verts == 5 * v.shape[0]
    assert te.host[3].tolist() == [0, f.shape[0], 3 * v.shape[0], 3 * f.shape[0], 0, v.shape[0], 3 * v.shape[0], 0]
    assert e.verts_packed().shape == (5 * v.shape[0], 3)  # materialises
    assert torch.equal(e.faces_packed()[f.shape[0]: 2 * f.shape[0]], f + v.shape[0])
    assert e.textures.verts_features_packed().shape == (5 * v.shape[0], 3)
    # offsets / scaling / update_padded / clone
    o = single.offset_verts(torch.ones(v.shape[0], 3))
    assert torch.allclose(o.verts_packed(), v + 1)
    o2 = single.offset_verts(torch.tensor([1.0, 0.0, 0.0]))
    assert torch.allclose(o2.verts_packed()[:, 0], v[:, 0] + 1)
    c = single.clone()
    c.scale_verts_(2.0)
    assert torch.allclose(c.verts_packed(), 2 * v) and torch.allclose(single.verts_packed(), v)
    u = m.update_padded(m.verts_padded() * 3)
    assert torch.allclose(u.verts_list()[1], v2 * 3)
    vv, ff = single.get_mesh_verts_faces(0)
    assert vv.shape == v.shape and ff.shape == f.shape
    with pytest.raises(ValueError):
        trb.Meshes(verts=[v], faces=[f, f2])
    with pytest.raises(ValueError):
        trb.Meshes(verts=[v], faces=[f], textures=trb.TexturesVertex(torch.rand(2, v.shape[0], 3)))


def test_settings_validation_matches_upstream_errors():
    from torch_renderer_b200.rasterizer import _check_bin_size, _parse_image_size
    assert _parse_image_size(64) == (64, 64) and _parse_image_size((180, 320)) == (180, 320)
    for bad in ((1, 2, 3), (0, 4), "x", (1.5, 2)):
        with pytest.raises(ValueError):
            _parse_image_size(bad)
    with pytest.raises(ValueError):
        _check_bin_size(8, 512, 512)  # 64 bins per side >= 22
    _check_bin_size(32, 512, 512)
    _check_bin_size(None, 512, 512)
    _check_bin_size(0, 512, 512)
    s = trb.RasterizationSettings()
    assert (s.image_size, s.blur_radius, s.faces_per_pixel, s.perspective_correct) == (256, 0.0, 1, None)
    b = trb.BlendParams()
    assert (b.sigma, b.gamma, tuple(b.background_color)) == (1e-4, 1e-4, (1.0, 1.0, 1.0))


def test_cameras_lights_materials_broadcast_and_defaults():
    cams = trb.FoVPerspectiveCameras(R=torch.eye(3)[None].repeat(4, 1, 1), T=torch.zeros(4, 3))
    assert len(cams) == 4 and cams.fov.shape == (4,) and cams.is_perspective() and cams.in_ndc()
    proj, persp = cams.ndc_projection_params()
    assert persp and proj.shape == (4, 4)
    assert torch.allclose(proj[0], torch.tensor([1.7320508, 1.7320508, 0.0, 0.0]))
    assert torch.allclose(cams[1:3].R, torch.eye(3)[None].repeat(2, 1, 1))
    pc = trb.PerspectiveCameras(focal_length=torch.tensor([[500.0, 510.0]]), principal_point=torch.tensor([[300.0, 200.0]]),
                                in_ndc=False, image_size=torch.tensor([[480, 640]]))
    proj, _ = pc.ndc_projection_params()
    assert torch.allclose(proj[0], torch.tensor([500 / 240, 510 / 240, -(300 - 320) / 240, -(200 - 240) / 240]))
    pts = torch.tensor([[0.3, -0.2, 2.0], [0.0, 0.1, 1.5]])
    ndc = pc.transform_points_ndc(pts)
    assert torch.allclose(ndc[:, 0], proj[0, 0] * pts[:, 0] / pts[:, 2] + proj[0, 2], atol=1e-6)
    assert torch.allclose(ndc[:, 1], proj[0, 1] * pts[:, 1] / pts[:, 2] + proj[0, 3], atol=1e-6)
    # 4x4 K matrix form (reference renderer.py:47-69)
    K = torch.tensor([[500.0, 0, 300, 0], [0, 510, 200, 0], [0, 0, 0, 1], [0, 0, 1, 0]])[None]
    pk = trb.PerspectiveCameras(K=K, in_ndc=False, image_size=torch.tensor([[480, 640]]))
    assert torch.allclose(pk.ndc_projection_params()[0], proj)
    lights = trb.PointLights(location=[[0.0, 0.0, -3.0]])
    assert lights.location.shape == (1, 3) and torch.allclose(lights.ambient_color, torch.full((1, 3), 0.5))
    lights.location = torch.tensor([[1.0, 2.0, 3.0]])
    assert lights.clone().location[0, 1] == 2.0
    assert torch.allclose(trb.AmbientLights().ambient_color, torch.ones(1, 3))
    m = trb.Materials()
    assert m.shininess.shape == (1,) and float(m.shininess[0]) == 64.0
    # camera centre: C = -T R^-1
    R, T = trb.look_at_view_transform(dist=2.0, elev=30.0, azim=40.0)
    c = trb.FoVPerspectiveCameras(R=R, T=T).get_camera_center()
    assert abs(float(c.norm()) - 2.0) < 1e-5


def test_obj_roundtrip_and_ico_sphere(tmp_path):
    m = trb.ico_sphere(2)
    assert m.verts_packed().shape == (162, 3) and m.faces_packed().shape == (320, 3)
    assert torch.allclose(m.verts_packed().norm(dim=1), torch.ones(162), atol=1e-6)
    assert trb.ico_sphere(4).faces_packed().shape == (5120, 3)
    p = tmp_path / "s.obj"
    trb.save_obj(str(p), m.verts_packed(), m.faces_packed())
    v, f, aux = trb.load_obj(str(p))
    assert torch.allclose(v, m.verts_packed(), atol=1e-5) and torch.equal(f.verts_idx, m.faces_packed())
    # polygons are fan-triangulated, v/vt/vn and negative indices parse
    q = tmp_path / "q.obj"
    q.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\n"
                 "f 1/1/1 2/2/1 3/3/1 4/4/1\nf -4//-1 -3//-1 -2//-1\n")
    v, f, aux = trb.load_obj(str(q))
    assert f.verts_idx.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]]
    assert f.textures_idx.tolist()[:2] == [[0, 1, 2], [0, 2, 3]] and f.textures_idx.tolist()[2] == [-1, -1, -1]
    assert aux.verts_uvs.shape == (4, 2) and aux.normals.shape == (1, 3)
    # textured save (deform_mesh_with_color.py:460): .obj + .mtl + .png, read back by load_obj / load_objs_as_meshes
    uvs = torch.tensor([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    tex = (torch.arange(6 * 5 * 3, dtype=torch.float32).reshape(6, 5, 3) % 256) / 255.0
    t = tmp_path / "colored.obj"
    trb.save_obj(str(t), v, f.verts_idx, verts_uvs=uvs, faces_uvs=f.verts_idx[:, [0, 2, 1]], texture_map=tex)
    assert (tmp_path / "colored.mtl").read_text().startswith("newmtl mesh\nmap_Kd colored.png\n")
    v2, f2, aux2 = trb.load_obj(str(t))
    assert torch.allclose(v2, v) and torch.equal(f2.verts_idx, f.verts_idx)
    assert torch.equal(f2.textures_idx, f.verts_idx[:, [0, 2, 1]]) and torch.allclose(aux2.verts_uvs, uvs)
    assert list(aux2.texture_images) == ["mesh"] and torch.allclose(aux2.texture_images["mesh"], tex, atol=1 / 255.0)
    mesh = trb.load_objs_as_meshes([str(t)])
    assert isinstance(mesh.textures, trb.TexturesUV) and torch.allclose(mesh.textures.maps_padded()[0], tex, atol=1 / 255.0)
    with pytest.raises(ValueError):
        trb.save_obj(str(t), v, f.verts_idx, verts_uvs=uvs[:, :1], faces_uvs=f.verts_idx, texture_map=tex)
    with pytest.raises(ValueError):
        trb.save_obj(str(t), v, f.verts_idx, verts_uvs=uvs, faces_uvs=f.verts_idx, texture_map=tex[..., :2])
    plain = tmp_path / "plain.obj"     # a partial texture triple is ignored, as upstream
    trb.save_obj(str(plain), v, f.verts_idx, verts_uvs=uvs)
    assert "vt" not in plain.read_text() and not (tmp_path / "plain.mtl").exists()


def test_golden_meshes_fixture():
    v, f = load_mesh("teapot")
    assert v.shape == (1292, 3) and f.shape == (2464, 3)
    v, f = load_mesh("cow")
    assert v.shape == (2930, 3) and f.shape == (5856, 3) and int(f.max()) == 2929


def test_compat_alias_keeps_reference_import_lines_working():
    import subprocess, sys
    code = (
        "import torch_renderer_b200.compat as c; c.install()\n"
        "from pytorch3d.io import load_objs_as_meshes, load_obj\n"
        "from pytorch3d.utils import cameras_from_opencv_projection, ico_sphere\n"
        "from pytorch3d.structures import Meshes, Pointclouds\n"
        "from pytorch3d.renderer import (PerspectiveCameras, PointLights, DirectionalLights, Materials,"
        " RasterizationSettings, MeshRenderer, MeshRasterizer, SoftPhongShader, SoftSilhouetteShader, BlendParams,"
        " PointsRasterizationSettings, PointsRenderer, PulsarPointsRenderer, PointsRasterizer, AlphaCompositor,"
        " NormWeightedCompositor, FoVPerspectiveCameras, look_at_view_transform, TexturesVertex, HardPhongShader,"
        " AmbientLights)\n"
        "from pytorch3d.transforms import quaternion_to_matrix, quaternion_apply, matrix_to_quaternion, Rotate,"
        " Translate, axis_angle_to_matrix\n"
        "from pytorch3d.loss import chamfer_distance\n"
        "import torch_renderer_b200 as t; assert MeshRenderer is t.MeshRenderer\n"
        "assert Pointclouds is t.Pointclouds and PointsRenderer is t.PointsRenderer\n"
        "try:\n    PulsarPointsRenderer()\nexcept NotImplementedError:\n    print('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


def test_allreduce_and_uv_entry_points_reject_bad_arguments_without_gpu():
    """The argument checks of the newer entry points run before any CUDA call."""
    L = _lib.lib()
    assert L.trb_allreduce_grid(1) == 1 and L.trb_allreduce_grid(1 << 16) == 64
    seg = (ctypes.c_void_p * 1)(8)
    cnt = (ctypes.c_int64 * 1)(4)
    inbox = (ctypes.c_void_p * 2)(16, 32)
    ok_tail = (1 << 16, 0, 2, 64, 128, 0, None)
    assert L.trb_allreduce_sum_f32(seg, cnt, 5, inbox, *ok_tail) == _lib.TRB_ERR_BAD_ARG      # > 4 segments
    assert L.trb_allreduce_sum_f32(seg, cnt, 1, inbox, 1 << 16, 2, 2, 64, 128, 0, None) == _lib.TRB_ERR_BAD_ARG  # rank >= world
    assert L.trb_allreduce_sum_f32(seg, cnt, 1, inbox, 2, 0, 2, 64, 128, 0, None) == _lib.TRB_ERR_BAD_ARG  # total > capacity
    assert L.trb_allreduce_sum_f32(seg, cnt, 1, None, *ok_tail) == _lib.TRB_ERR_BAD_ARG        # no inboxes
    # fused render with texture_mode UV but no trb_uv_texture
    cfg = _lib.RenderConfig()
    cfg.shade.N, cfg.shade.H, cfg.shade.W, cfg.shade.K = 1, 8, 8, 1
    cfg.shade.shader, cfg.shade.light_kind, cfg.shade.texture_mode = _lib.SHADER_SOFT_PHONG, 1, _lib.TEX_UV
    cfg.shade.sigma = cfg.shade.gamma = 1e-4
    cfg.max_face_count = cfg.max_vert_count = 1
    args = [8] * 18
    assert L.trb_render_forward(ctypes.byref(cfg), *args, 1024, 8, None, None, 0, None) == _lib.TRB_ERR_BAD_ARG
    cfg.shade.texture_mode = 7
    assert L.trb_render_forward(ctypes.byref(cfg), *args, 1024, 8, None, None, 0, None) == _lib.TRB_ERR_BAD_ARG
    assert L.trb_abi_struct_size(3) == ctypes.sizeof(_lib.UvTexture)


def test_compat_resolves_every_reference_import():
    """Every `from pytorch3d... import name` of the reference scripts (fixture generated from /root/reference by
    tests/golden/make_reference_imports.py) resolves after compat.install(): the scripts import unchanged."""
    import importlib
    import json
    import torch_renderer_b200.compat as compat
    compat.install(force=True)
    table = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_imports.json")))
    assert sum(len(v) for v in table.values()) > 40
    missing = []
    for module, names in table.items():
        mod = importlib.import_module(module)
        missing += [f"{module}.{n}" for n in names if not hasattr(mod, n)]
    assert not missing, missing
    # the in-path names are the real thing, not stubs
    import pytorch3d.loss as p3l
    import pytorch3d.renderer as p3r
    assert p3r.MeshRenderer is trb.MeshRenderer and p3l.chamfer_distance is trb.chamfer_distance


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/trb.h is the C-ABI boundary: it must compile as strict C99 and a C program must link against
    libtrb.so and call the bookkeeping entry points (what a cgo / JNI / N-API binding would do first)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "trb.h"\n'
                   'int main(void) { size_t n = 0; trb_render_config c; (void)c;\n'
                   '  if (trb_abi_version() != TRB_ABI_VERSION) return 1;\n'
                   '  if (trb_abi_struct_size(0) != (int)sizeof(trb_view)) return 2;\n'
                   '  if (trb_abi_struct_size(1) != (int)sizeof(trb_shade_config)) return 3;\n'
                   '  if (trb_abi_struct_size(2) != (int)sizeof(trb_render_config)) return 4;\n'
                   '  if (trb_abi_struct_size(3) != (int)sizeof(trb_uv_texture)) return 5;\n'
                   '  if (trb_raster_workspace_bytes(1, 32, 32, 2, 100, &n) != TRB_OK || n == 0) return 6;\n'
                   '  puts(trb_status_string(TRB_ERR_K_TOO_LARGE)); return 0; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        str(src), "-o", str(exe), "-L", libdir, "-l:libtrb.so", f"-Wl,-rpath,{libdir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    assert "150" in r.stdout


def test_clip_entry_points_and_host_logic_without_gpu():
    """Argument checks of the clipping entry points (before any CUDA call), the face -> NDC-row map of the clipped
    route, the mode switch, and the upstream module path of clip_faces through compat."""
    L = _lib.lib()
    p = 8  # any non-null pointer value: the checks below fail before it is dereferenced
    ok = (p, p, p, p, 4, p, 1, 16, 16)
    assert L.trb_clip_resequence(*ok, 151, 0.0, 0, p, p, p, p, None, 0, None) == _lib.TRB_ERR_K_TOO_LARGE
    assert L.trb_clip_resequence(*ok, 0, 0.0, 0, p, p, p, p, None, 0, None) == _lib.TRB_ERR_BAD_ARG
    assert L.trb_clip_resequence(*ok, 2, -1.0, 0, p, p, p, p, None, 0, None) == _lib.TRB_ERR_BAD_ARG
    assert L.trb_clip_resequence(p, p, p, p, -1, p, 1, 16, 16, 2, 0.0, 0, p, p, p, p, None, 0, None) == _lib.TRB_ERR_BAD_ARG
    assert L.trb_clip_resequence(p, p, p, p, 0, p, 1, 16, 16, 2, 0.0, 0, p, p, p, p, None, 0, None) == _lib.TRB_OK  # no pairs
    assert L.trb_clip_resequence(p, p, None, p, 4, p, 1, 16, 16, 2, 0.0, 0, p, p, p, p, None, 0, None) == _lib.TRB_ERR_BAD_ARG
    assert L.trb_any_vertex_behind(p, p, p, p, 1, 10, 0.5, 1, None, None, None, 0, None) == _lib.TRB_ERR_BAD_ARG   # no flag
    assert L.trb_any_vertex_behind(p, p, p, p, 1, 10, float("nan"), 1, p, None, None, 0, None) == _lib.TRB_ERR_BAD_ARG


    from torch_renderer_b200 import rasterizer
    with pytest.raises(ValueError):
        trb.set_near_plane_clipping("sometimes")
    trb.set_near_plane_clipping("off")
    assert rasterizer._clipping_mode == "off"
    trb.set_near_plane_clipping()
    assert rasterizer._clipping_mode == "exact"

    v, f = uv_sphere(4, 6)
    shared = trb.Meshes(verts=[v], faces=[f]).extend(3)
    rows = rasterizer._face_vertex_rows(shared.faces_packed_i32(), shared.view_table())
    assert rows.shape == (3 * f.shape[0], 3)
    assert torch.equal(rows.reshape(3, -1, 3)[2], f + 2 * v.shape[0])
    packed = trb.Meshes(verts=[v, v[:10]], faces=[f, f[:3].clamp(max=9)])
    assert torch.equal(rasterizer._face_vertex_rows(packed.faces_packed_i32(), packed.view_table()),
                       packed.faces_packed())
    # z_clip_value resolution: znear / 2 only for perspective-correct renders of cameras that define znear
    rast = trb.MeshRasterizer(trb.FoVPerspectiveCameras(znear=0.4), trb.RasterizationSettings(image_size=8))
    one = trb.Meshes(verts=[v], faces=[f])
    assert rast._resolve(one, {})[4]["z_clip_value"] == pytest.approx(0.2)
    flat = trb.RasterizationSettings(image_size=8, perspective_correct=False)
    assert rast._resolve(one, {"raster_settings": flat})[4]["z_clip_value"] is None
    assert trb.MeshRasterizer(trb.PerspectiveCameras(), trb.RasterizationSettings(image_size=8))._resolve(one, {})[4]["z_clip_value"] is None
    explicit = trb.RasterizationSettings(image_size=8, z_clip_value=0.05, cull_to_frustum=True)
    spec = rast._resolve(one, {"raster_settings": explicit})[4]
    assert spec["z_clip_value"] == pytest.approx(0.05) and spec["cull_to_frustum"] is True

    import subprocess, sys
    code = ("import torch_renderer_b200.compat as c; c.install()\n"
            "from pytorch3d.renderer.mesh.clip import ClipFrustum, ClippedFaces, clip_faces, "
            "convert_clipped_rasterization_to_original_faces\n"
            "import torch_renderer_b200 as t; assert clip_faces is t.clip.clip_faces; print('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


def test_memoised_parameter_blocks_see_parameters_and_grad_flags():
    """a light / camera attribute assigned as ``nn.Parameter`` lands in ``_parameters``; the memoised
    view-parameter block and projection must see it (no stale values, no stale autograd graph), and the Fragments
    cache key carries ``requires_grad``."""
    import torch.nn as nn
    from torch_renderer_b200 import shader as sh
    from torch_renderer_b200.rasterizer import _cached_projection, _FragmentCache, _fragment_cache
    from torch_renderer_b200.common import named_tensors

    lights = trb.PointLights(location=[[0.0, 0.0, -3.0]])
    materials = trb.Materials()
    cams = trb.FoVPerspectiveCameras()
    owner = nn.Module()
    a = sh._cached_view_params(owner, 1, torch.device("cpu"), lights, materials, cams, 1.0, 100.0, False)
    b = sh._cached_view_params(owner, 1, torch.device("cpu"), lights, materials, cams, 1.0, 100.0, False)
    assert a is b                                              # constants: memoised
    lights.location = nn.Parameter(torch.tensor([[1.0, 2.0, 3.0]]))
    assert "location" in named_tensors(lights) and "location" not in lights.__dict__
    c = sh._cached_view_params(owner, 1, torch.device("cpu"), lights, materials, cams, 1.0, 100.0, False)
    assert c is not a and c.requires_grad and torch.equal(c[0, :3].detach(), torch.tensor([1.0, 2.0, 3.0]))
    with torch.no_grad():
        lights.location.mul_(2.0)
    d = sh._cached_view_params(owner, 1, torch.device("cpu"), lights, materials, cams, 1.0, 100.0, False)
    assert d is not c and torch.equal(d[0, :3].detach(), torch.tensor([2.0, 4.0, 6.0]))
    d.sum().backward()                                         # a fresh graph every call
    assert lights.location.grad is not None

    pc = trb.PerspectiveCameras(focal_length=((1.0, 1.0),))
    p0, _ = _cached_projection(pc, {})
    pc.focal_length = nn.Parameter(torch.tensor([[1.0, 1.0]]))
    with torch.no_grad():
        pc.focal_length.mul_(2.0)
    p1, _ = _cached_projection(pc, {})
    assert p1.requires_grad and torch.allclose(p1[0, :2].detach(), 2.0 * p0[0, :2])
    # .to() / clone() / indexing carry Parameter-valued attributes too
    assert torch.equal(lights.clone().location, lights.location.detach())
    assert lights[0].location.shape == (1, 3)

    # the Fragments cache is opt-in, and its key sees requires_grad
    assert _fragment_cache.enabled is False
    t = torch.zeros(3)
    spec = dict(image_size=(4, 4), K=1, blur_radius=0.0, flags=0, z_clip=0.0, perspective=True, cull_to_frustum=False)
    k0 = _FragmentCache.make_key((t,), spec, None)
    t.requires_grad_(True)
    assert _FragmentCache.make_key((t,), spec, None) != k0

    # get_camera_center remembers R / T overrides on the camera, like upstream
    R = torch.eye(3)[None] * torch.tensor([1.0, -1.0, -1.0])
    T = torch.tensor([[0.0, 0.0, 5.0]])
    cams.get_camera_center(R=R, T=T)
    assert cams.R is R and cams.T is T


def test_round2_host_pieces_without_gpu():
    """capture_step needs CUDA and says so; the bench workload builders import without the oracle and build the
    BASELINE configs[4] mesh at the stated size; the C5-at-spec sharding covers every view exactly once."""
    import importlib
    import sys as _sys
    from torch_renderer_b200.parallel import chunk_views, shard_views
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            trb.capture_step(lambda: None)
    assert issubclass(trb.NearPlaneCrossed, RuntimeError)
    _sys.path.insert(0, ROOT)
    wl = importlib.import_module("bench_workloads")
    assert "oracle" not in wl.__dict__ and set(wl.BUILDERS) >= {"C1", "C3", "C3cow", "C4", "C5", "pose_step", "clipped"}
    v, f = wl.grid_sphere(11, 20)
    assert f.shape[0] == 2 * 11 * 20 - 2 * 20 and v.shape[0] == (11 - 1) * 20 + 2 and int(f.max()) == v.shape[0] - 1
    # 501 x 1000 quads -> exactly 1,000,000 triangles and 500,002 vertices (the C5 mesh), by the same closed form
    assert 2 * 501 * 1000 - 2 * 1000 == 1_000_000 and (501 - 1) * 1000 + 2 == 500_002
    assert wl.bview(8, 1024, 1024, 500_002, 1_000_000) == 599_316_672
    seen = []
    for r in range(8):
        lo, hi = shard_views(1024, r, 8)
        for s0, s1 in chunk_views(hi - lo, 32):
            seen.extend(range(lo + s0, lo + s1))
    assert seen == list(range(1024))
    # tile-list capacity estimate of the point rasteriser: a scalar radius bounds the tiles a disc can reach
    from torch_renderer_b200 import ops
    import inspect
    assert "max_radius" in inspect.signature(ops.rasterize_points_ndc).parameters


def test_padded_inputs_are_split_without_select_nodes():
    """`TexturesVertex(rgb[None])` / `Meshes(verts=padded, faces=padded)` -- the batch-of-one pattern of every reference
    script -- must not put a SelectBackward node (a zeros + copy_ kernel pair per backward) between the leaf and the
    kernels: one row is squeezed, several are unbound; gradients still reach the leaf."""
    import torch_renderer_b200 as trb
    from torch_renderer_b200.common import unbind_batch
    rgb = torch.rand(1, 5, 3, requires_grad=True)
    tex = trb.TexturesVertex(rgb)
    feats = tex.verts_features_packed()
    assert type(feats.grad_fn).__name__.startswith("Squeeze")
    feats.sum().backward()
    assert torch.equal(rgb.grad, torch.ones_like(rgb))
    verts = torch.rand(1, 4, 3, requires_grad=True)
    faces = torch.tensor([[[0, 1, 2], [1, 2, 3]]])
    mesh = trb.Meshes(verts=verts, faces=faces)
    assert type(mesh.verts_list()[0].grad_fn).__name__.startswith("Squeeze")
    (mesh.verts_packed() * 2).sum().backward()
    assert torch.equal(verts.grad, torch.full_like(verts, 2.0))
    many = torch.rand(3, 4, 3, requires_grad=True)
    rows = unbind_batch(many)
    assert len(rows) == 3 and all(type(r.grad_fn).__name__.startswith("Unbind") for r in rows)
    (rows[0].sum() + 3 * rows[2].sum()).backward()
    assert torch.equal(many.grad[0], torch.ones(4, 3)) and torch.equal(many.grad[1], torch.zeros(4, 3))
    assert torch.equal(many.grad[2], torch.full((4, 3), 3.0))
