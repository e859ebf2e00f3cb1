"""GPU parity tests proper: the CUDA path (through the C ABI in libtrb.so) against the CPU oracle on
the same seeded inputs.  Bars: pix_to_face bit-exact; zbuf / bary / dists
within 1e-5 + 1e-5*|x| (they are in fact produced by the same IEEE sequence); images within 1e-4;
gradients within 1e-3 relative L2 of the fp64 autograd model."""
import math

import numpy as np
import pytest
import torch

import oracle
from oracle import shading_ref as sref
from helpers import (cow_uvs, fov_proj, load_mesh, normalize_mesh, oracle_rasterize, rel_l2, uv_sphere)

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0")
TOL = dict(atol=1e-5, rtol=1e-5)


def _trb():
    import torch_renderer_b200 as trb
    return trb


def _views(n, dist=2.7, seed=0):
    from torch_renderer_b200.cameras import look_at_view_transform
    g = torch.Generator().manual_seed(seed)
    elev = torch.rand(n, generator=g) * 140 - 70
    azim = torch.rand(n, generator=g) * 360 - 180
    return look_at_view_transform(dist=dist, elev=elev, azim=azim)


def _ndc(verts, R, T, proj, perspective=True):
    return sref.world_to_ndc(verts, R, T, proj[:, 0], proj[:, 1], proj[:, 2], proj[:, 3], perspective)


def _cuda_raster_from_ndc(ndc, faces, image_size, blur, K, persp, clip, cull=False, capacity=None):
    """ndc [N,V,3] CPU tensor -> CUDA fragments via Meshes of NDC vertices (rasterize_meshes twin)."""
    trb = _trb()
    N = ndc.shape[0]
    meshes = trb.Meshes(verts=[ndc[i].to(DEV) for i in range(N)], faces=[faces.to(DEV)] * N)
    if capacity is not None:
        meshes.view_table().pair_capacity = capacity
    return trb.renderer.rasterize_meshes(meshes, image_size, blur, K, perspective_correct=persp,
                                         clip_barycentric_coords=clip, cull_backfaces=cull)


def _assert_fragments_equal(got, want, exact_floats=True):
    p2f, zbuf, bary, dists = [t.cpu().numpy() for t in got]
    o_p2f, o_z, o_b, o_d = want
    assert p2f.dtype == np.int64
    mism = int((p2f != o_p2f).sum())
    assert mism == 0, f"pix_to_face differs at {mism} of {p2f.size} samples"
    for name, a, b in (("zbuf", zbuf, o_z), ("bary", bary, o_b), ("dists", dists, o_d)):
        assert np.allclose(a, b, **TOL), f"{name} max abs diff {np.abs(a - b).max()}"
    return all(np.array_equal(a, b) for a, b in ((zbuf, o_z), (bary, o_b), (dists, o_d)))


SCENES = [
    # name, image_size, K, blur, persp, clip
    ("teapot", (64, 64), 1, 0.0, True, False),
    ("teapot", (256, 256), 1, 0.0, True, False),
    ("teapot", (96, 160), 4, 0.0, False, False),
    ("teapot", (128, 128), 8, 9.21024e-4, True, True),
    ("cow", (128, 128), 1, 0.0, True, False),
    ("cow", (200, 120), 3, 1e-3, False, True),
    ("teapot", (45, 61), 1, 0.0, True, False),       # rows not 16-byte aligned, partial edge tiles
    ("teapot", (45, 61), 3, 5e-4, True, True),
    ("cow", (130, 258), 1, 0.0, False, False),       # W % 4 == 2: scalar fill path of the K=1 strips
    ("sphere", (33, 47), 30, 2e-3, True, True),      # 8x8 tiles, odd sizes
    ("sphere", (64, 64), 50, 9.21024e-4, True, True),
    ("sphere", (40, 40), 150, 4e-3, False, True),
]


def _scene(name):
    if name == "sphere":
        v, f = uv_sphere(20, 24, 1.0, noise=0.03, seed=3)
    else:
        v, f = load_mesh(name)
        v = normalize_mesh(v)
    return v, f


@pytest.mark.parametrize("name,image_size,K,blur,persp,clip", SCENES)
def test_raster_forward_bit_exact(name, image_size, K, blur, persp, clip):
    v, f = _scene(name)
    R, T = _views(3, seed=sum(map(ord, name)) % 100)
    ndc = _ndc(v, R, T, fov_proj(3))
    want = oracle_rasterize(ndc, f, image_size, blur, K, persp, clip)
    got = _cuda_raster_from_ndc(ndc, f, image_size, blur, K, persp, clip)
    _assert_fragments_equal(got, want)
    assert (want[0] >= 0).sum() > 100


def test_raster_hand_scene_and_edge_cases():
    trb = _trb()
    big = lambda z: [[-0.9, -0.9, z], [0.9, -0.9, z], [0.0, 0.9, z]]
    tris = torch.tensor([big(2.0), big(1.0), big(2.0),
                         [[-0.5, -0.5, 1.0], [0.0, 0.0, 1.0], [0.5, 0.5, 1.0]],   # degenerate
                         [[-0.9, -0.9, -1.0], [0.9, -0.9, -1.0], [0.0, 0.9, -1.0]],  # behind
                         [[-0.9, -0.9, 1.0], [0.9, -0.9, 1.0], [0.0, 0.9, 0.0]],   # touches plane
                         [[5.0, 5.0, 1.0], [6.0, 5.0, 1.0], [5.0, 6.0, 1.0]]])      # off screen
    verts = tris.reshape(-1, 3)
    faces = torch.arange(verts.shape[0]).reshape(-1, 3)
    for K in (1, 2, 3, 5):
        for cull in (False, True):
            want = oracle_rasterize(verts[None], faces, (16, 24), 0.0, K, False, False, cull)
            got = _cuda_raster_from_ndc(verts[None], faces, (16, 24), 0.0, K, False, False, cull)
            _assert_fragments_equal(got, want)
    # empty mesh: all -1
    meshes = trb.Meshes(verts=[torch.zeros(0, 3, device=DEV)], faces=[torch.zeros(0, 3, dtype=torch.int64, device=DEV)])
    p2f, zbuf, bary, dists = trb.renderer.rasterize_meshes(meshes, 8, 0.0, 2)
    assert (p2f == -1).all() and (zbuf == -1).all() and (bary == -1).all() and (dists == -1).all()
    with pytest.raises(ValueError):
        trb.renderer.rasterize_meshes(meshes, 8, 0.0, 151)


def test_raster_heterogeneous_batch():
    trb = _trb()
    v1, f1 = _scene("teapot")
    v2, f2 = _scene("sphere")
    R, T = _views(2, seed=5)
    n1 = _ndc(v1, R[:1], T[:1], fov_proj(1))[0]
    n2 = _ndc(v2, R[1:], T[1:], fov_proj(1))[0]
    meshes = trb.Meshes(verts=[n1.to(DEV), n2.to(DEV)], faces=[f1.to(DEV), f2.to(DEV)])
    got = trb.renderer.rasterize_meshes(meshes, (72, 56), 1e-4, 3, perspective_correct=True,
                                        clip_barycentric_coords=True)
    fv = torch.cat([n1[f1], n2[f2]], 0).numpy()
    first = np.array([0, f1.shape[0]], np.int64)
    count = np.array([f1.shape[0], f2.shape[0]], np.int64)
    want = oracle.rasterize_forward(fv, first, count, (72, 56), 1e-4, 3, True, True, False)
    _assert_fragments_equal(got, want)
    assert got[0][1].max() >= f1.shape[0]


def test_raster_bin_overflow_falls_back_to_full_scan():
    v, f = _scene("teapot")
    R, T = _views(2, seed=9)
    ndc = _ndc(v, R, T, fov_proj(2))
    want = oracle_rasterize(ndc, f, (96, 96), 1e-4, 2, True, True)
    got = _cuda_raster_from_ndc(ndc, f, (96, 96), 1e-4, 2, True, True, capacity=64)
    _assert_fragments_equal(got, want)


@pytest.mark.parametrize("persp,clip,blur,K", [(False, False, 0.0, 1), (True, False, 0.0, 2), (True, True, 2e-3, 4)])
def test_raster_backward(persp, clip, blur, K):
    trb = _trb()
    torch.manual_seed(1)
    v, f = uv_sphere(10, 12, 1.0, noise=0.05, seed=2)
    R, T = _views(2, seed=11)
    ndc = _ndc(v, R, T, fov_proj(2))
    H, W = 48, 40
    verts_dev = ndc.reshape(-1, 3).to(DEV).requires_grad_(True)
    meshes = trb.Meshes(verts=[verts_dev[: v.shape[0]], verts_dev[v.shape[0]:]], faces=[f.to(DEV)] * 2)
    p2f, zbuf, bary, dists = trb.renderer.rasterize_meshes(meshes, (H, W), blur, K, perspective_correct=persp,
                                                           clip_barycentric_coords=clip)
    gz, gb, gd = torch.randn(2, H, W, K), torch.randn(2, H, W, K, 3), torch.randn(2, H, W, K)
    m = (p2f >= 0).cpu()
    loss = (zbuf * (gz * m).to(DEV)).sum() + (bary * (gb * m[..., None]).to(DEV)).sum() + (dists * (gd * m).to(DEV)).sum()
    loss.backward()
    got = verts_dev.grad.cpu()
    # oracle 1: C backward (fp32 contributions, fp64 accumulation)
    fv = ndc[:, f].reshape(-1, 3, 3)
    g_fv = torch.from_numpy(oracle.rasterize_backward(fv.numpy(), p2f.cpu().numpy(), (gz * m).numpy(),
                                                      (gb * m[..., None]).numpy(), (gd * m).numpy(), persp, clip))
    want = torch.zeros(2, v.shape[0], 3, dtype=torch.float64)
    for n in range(2):
        want[n].index_add_(0, f.reshape(-1), g_fv.reshape(2, -1, 3, 3)[n].reshape(-1, 3).double())
    assert rel_l2(got, want.reshape(-1, 3)) < 1e-3
    # oracle 2: fp64 autograd
    fv64 = fv.double().requires_grad_(True)
    z64, b64, d64 = sref.raster_recompute(fv64, p2f.cpu(), persp, clip)
    ((z64 * gz * m).sum() + (b64 * gb * m[..., None]).sum() + (d64 * gd * m).sum()).backward()
    want2 = torch.zeros(2, v.shape[0], 3, dtype=torch.float64)
    for n in range(2):
        want2[n].index_add_(0, f.reshape(-1), fv64.grad.reshape(2, -1, 3, 3)[n].reshape(-1, 3))
    assert rel_l2(got, want2.reshape(-1, 3)) < 1e-3


def test_transform_and_normals_forward_backward():
    from torch_renderer_b200 import ops
    trb = _trb()
    torch.manual_seed(0)
    v, f = uv_sphere(8, 10, 0.8, noise=0.1, seed=4)
    N = 3
    R, T = _views(N, seed=2)
    proj = fov_proj(N) + torch.tensor([0.0, 0.0, 0.05, -0.02])
    for perspective in (True, False):
        vd = v.to(DEV).requires_grad_(True)
        Rd, Td, pd = R.to(DEV).requires_grad_(True), T.to(DEV).requires_grad_(True), proj.to(DEV).requires_grad_(True)
        mesh = trb.Meshes(verts=[vd], faces=[f.to(DEV)]).extend(N)
        out = ops.transform_verts(vd, Rd, Td, pd, mesh.view_table(), perspective)
        v64, R64, T64, p64 = (t.double().requires_grad_(True) for t in (v, R, T, proj))
        ref = _ndc(v64, R64, T64, p64, perspective)
        assert torch.allclose(out.cpu().reshape(N, -1, 3), ref.float(), atol=2e-6, rtol=2e-6)
        g = torch.randn(N, v.shape[0], 3)
        (out.reshape(N, -1, 3) * g.to(DEV)).sum().backward()
        (ref * g.double()).sum().backward()
        for a, b in ((vd, v64), (Rd, R64), (Td, T64), (pd, p64)):
            assert rel_l2(a.grad.cpu(), b.grad) < 1e-4
    # vertex normals
    vd = v.to(DEV).requires_grad_(True)
    n = ops.vertex_normals(vd, f.to(DEV).to(torch.int32))
    v64 = v.double().requires_grad_(True)
    nref = sref.vertex_normals(v64, f)
    assert torch.allclose(n.cpu(), nref.float(), atol=1e-5)
    g = torch.randn_like(v)
    (n * g.to(DEV)).sum().backward()
    (nref * g.double()).sum().backward()
    assert rel_l2(vd.grad.cpu(), v64.grad) < 1e-4


def test_interpolate_face_attributes():
    trb = _trb()
    torch.manual_seed(0)
    for D in (2, 3, 5):
        p2f = torch.randint(-1, 7, (2, 5, 6, 3))
        bary = torch.rand(2, 5, 6, 3, 3)
        attrs = torch.randn(7, 3, D)
        bd, ad = bary.to(DEV).requires_grad_(True), attrs.to(DEV).requires_grad_(True)
        out = trb.interpolate_face_attributes(p2f.to(DEV), bd, ad)
        want = oracle.interp_forward(p2f.numpy(), bary.numpy(), attrs.numpy())
        assert np.allclose(out.detach().cpu().numpy(), want, atol=1e-6)
        g = torch.randn(2, 5, 6, 3, D)
        (out * g.to(DEV)).sum().backward()
        gb, ga = oracle.interp_backward(p2f.numpy(), bary.numpy(), attrs.numpy(), g.numpy())
        assert np.allclose(bd.grad.cpu().numpy(), gb, atol=1e-5)
        assert np.allclose(ad.grad.cpu().numpy(), ga, atol=1e-4)


def _render_setup(name, N, image_size, K, blur, shader, lights_kind, seed=0, background=(1.0, 1.0, 1.0),
                  persp=None, dist=2.7):
    """Builds the same scene for the CUDA API and for the oracle pipeline."""
    trb = _trb()
    torch.manual_seed(seed)
    v, f = _scene(name)
    colors = torch.rand(v.shape[0], 3)
    R, T = _views(N, dist=dist, seed=seed)
    return dict(v=v, f=f, colors=colors, R=R, T=T, N=N, image_size=image_size, K=K, blur=blur, shader=shader,
                lights_kind=lights_kind, background=background, persp=persp)


def _cuda_render(s, requires_grad=False, fused=False):
    trb = _trb()
    v = s["v"].to(DEV).requires_grad_(requires_grad)
    c = s["colors"].to(DEV).requires_grad_(requires_grad)
    R = s["R"].to(DEV).requires_grad_(requires_grad)
    T = s["T"].to(DEV).requires_grad_(requires_grad)
    mesh = trb.Meshes(verts=[v], faces=[s["f"].to(DEV)], textures=trb.TexturesVertex(verts_features=c[None]))
    meshes = mesh.extend(s["N"])
    cameras = trb.FoVPerspectiveCameras(device=DEV, R=R, T=T)
    settings = trb.RasterizationSettings(image_size=s["image_size"], blur_radius=s["blur"], faces_per_pixel=s["K"],
                                         perspective_correct=s["persp"])
    blend = trb.BlendParams(1e-4, 1e-4, s["background"])
    if s["lights_kind"] == "point":
        lights = trb.PointLights(device=DEV, location=[[0.5, 1.0, -2.5]])
    elif s["lights_kind"] == "directional":
        lights = trb.DirectionalLights(device=DEV, direction=[[0.3, 1.0, -0.5]])
    else:
        lights = trb.AmbientLights(device=DEV)
    rasterizer = trb.MeshRasterizer(cameras=cameras, raster_settings=settings)
    if s["shader"] == "soft_phong":
        shader = trb.SoftPhongShader(device=DEV, cameras=cameras, lights=lights, blend_params=blend)
    elif s["shader"] == "hard_phong":
        shader = trb.HardPhongShader(device=DEV, cameras=cameras, lights=lights, blend_params=blend)
    else:
        shader = trb.SoftSilhouetteShader(blend_params=blend)
    verts_ndc = rasterizer.transform(meshes)
    if fused:   # ONE C-ABI call each way (trb_render_forward / trb_render_backward)
        images, fragments = trb.MeshRendererWithFragments(rasterizer, shader)(meshes)
    else:       # rasteriser and shader composed as separate modules
        fragments = rasterizer(meshes)
        images = shader(fragments, meshes)
    return dict(images=images, fragments=fragments, verts_ndc=verts_ndc, v=v, c=c, R=R, T=T)


def _oracle_render(s, p2f, dtype=torch.float64, verts_ndc=None):
    """Oracle shading on top of a fixed pix_to_face (differentiable in fp64)."""
    v, c, R, T = (s[k].to(dtype).clone().requires_grad_(True) for k in ("v", "colors", "R", "T"))
    N, f = s["N"], s["f"]
    proj = fov_proj(N).to(dtype)
    ndc = _ndc(v, R, T, proj) if verts_ndc is None else verts_ndc
    persp = True if s["persp"] is None else s["persp"]
    clip = s["blur"] > 0
    fv = ndc[:, f].reshape(-1, 3, 3)
    zbuf, bary, dists = sref.raster_recompute(fv, p2f, persp, clip)
    F = f.shape[0]
    faces_rows = f.repeat(N, 1)  # every view indexes the shared vertex arrays
    normals = sref.vertex_normals(v, f)
    cam_center = -torch.matmul(T[:, None, :], torch.linalg.inv(R))[:, 0, :]
    ones = lambda *x: torch.tensor([list(x)], dtype=dtype).repeat(N, 1)
    kind = s["lights_kind"]
    lv = ones(0.5, 1.0, -2.5) if kind == "point" else ones(0.3, 1.0, -0.5)
    amb = ones(1, 1, 1) if kind == "ambient" else ones(0.5, 0.5, 0.5)
    img = sref.shade(p2f, bary, zbuf, dists, faces_rows, v, normals, c, shader=s["shader"], light_kind=kind,
                     light_vec=lv, light_ambient=amb, light_diffuse=ones(0.3, 0.3, 0.3),
                     light_specular=ones(0.2, 0.2, 0.2), mat_ambient=ones(1, 1, 1), mat_diffuse=ones(1, 1, 1),
                     mat_specular=ones(1, 1, 1), shininess=torch.full((N,), 64.0, dtype=dtype),
                     camera_center=cam_center, sigma=1e-4, gamma=1e-4, background=s["background"],
                     znear=1.0, zfar=100.0)
    return dict(images=img, v=v, c=c, R=R, T=T, zbuf=zbuf, bary=bary, dists=dists)


RENDER_CASES = [
    ("teapot", 2, (64, 64), 1, 0.0, "soft_phong", "point"),
    ("cow", 2, (96, 80), 1, 0.0, "soft_phong", "point"),
    ("teapot", 2, (64, 64), 4, 9.21024e-4, "soft_phong", "point"),
    ("sphere", 2, (48, 48), 3, 1e-3, "soft_phong", "directional"),
    ("teapot", 1, (64, 64), 1, 0.0, "soft_phong", "ambient"),
    ("teapot", 2, (64, 64), 2, 0.0, "hard_phong", "point"),
    ("teapot", 2, (64, 64), 10, 9.21024e-4, "soft_silhouette", "point"),
    ("sphere", 1, (48, 48), 50, 9.21024e-4, "soft_silhouette", "point"),
    ("sphere", 1, (40, 40), 30, 2e-3, "soft_phong", "point"),
]


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name,N,image_size,K,blur,shader,lights", RENDER_CASES)
def test_render_end_to_end_forward(name, N, image_size, K, blur, shader, lights, fused):
    """Public API (Meshes + cameras + lights + settings -> images / Fragments) vs the oracle pipeline
    fed with the NDC vertices the CUDA transform emitted (the bit-exact contract starts there)."""
    s = _render_setup(name, N, image_size, K, blur, shader, lights, seed=3, background=(0.1, 0.2, 0.3))
    out = _cuda_render(s, fused=fused)
    ndc = out["verts_ndc"].detach().cpu().reshape(N, -1, 3)
    # the CUDA transform itself against the torch restatement
    assert torch.allclose(ndc, _ndc(s["v"], s["R"], s["T"], fov_proj(N)), atol=2e-6, rtol=2e-6)
    want = oracle_rasterize(ndc, s["f"], image_size, blur, K, True, blur > 0)
    fr = out["fragments"]
    _assert_fragments_equal((fr.pix_to_face, fr.zbuf, fr.bary_coords, fr.dists), want)
    ref = _oracle_render(s, torch.from_numpy(want[0]), torch.float64, verts_ndc=ndc.double())
    img = out["images"].detach().cpu()
    assert img.shape == (N, image_size[0], image_size[1], 4)
    err = (img.double() - ref["images"]).abs().max().item()
    assert err < 1e-4, f"image max abs err {err}"


GRAD_CASES = [
    ("teapot", 2, (48, 48), 1, 0.0, "soft_phong", "point"),
    ("sphere", 2, (40, 40), 4, 2e-3, "soft_phong", "point"),
    ("sphere", 2, (40, 40), 3, 2e-3, "soft_phong", "directional"),
    ("teapot", 1, (48, 48), 1, 0.0, "soft_phong", "ambient"),
    ("teapot", 2, (48, 48), 1, 0.0, "hard_phong", "point"),
    ("sphere", 2, (40, 40), 12, 2e-3, "soft_silhouette", "point"),
    ("sphere", 1, (32, 32), 30, 2e-3, "soft_phong", "point"),
    ("sphere", 2, (40, 40), 3, 1e-3, "hard_phong", "directional"),
]


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name,N,image_size,K,blur,shader,lights", GRAD_CASES)
def test_render_end_to_end_gradients(name, N, image_size, K, blur, shader, lights, fused):
    """loss.backward() through the public API: gradients w.r.t. vertices, vertex colours and camera
    R / T against fp64 autograd of the oracle model (pix_to_face held fixed)."""
    s = _render_setup(name, N, image_size, K, blur, shader, lights, seed=5, background=(0.0, 0.0, 0.0))
    out = _cuda_render(s, requires_grad=True, fused=fused)
    torch.manual_seed(7)
    target = torch.rand(N, image_size[0], image_size[1], 4)
    w = torch.rand(N, image_size[0], image_size[1])
    loss = ((out["images"] - target.to(DEV)) ** 2).mean() + (torch.relu(out["fragments"].zbuf[..., 0]) * w.to(DEV)).mean()
    loss.backward()
    p2f = out["fragments"].pix_to_face.cpu()
    ref = _oracle_render(s, p2f, torch.float64)
    loss_ref = ((ref["images"] - target.double()) ** 2).mean() + (torch.relu(ref["zbuf"][..., 0]) * w.double()).mean()
    loss_ref.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-5 * max(1.0, abs(loss_ref.item()))
    names = ["v", "R", "T"] + (["c"] if shader != "soft_silhouette" else [])
    for k in names:
        g, gr = out[k].grad, ref[k].grad
        assert g is not None, k
        if gr.abs().max() == 0:
            assert g.abs().max() < 1e-8
            continue
        e = rel_l2(g.cpu(), gr)
        assert e < 1e-3, f"grad {k}: rel L2 err {e}"


def test_textures_uv_path():
    trb = _trb()
    torch.manual_seed(0)
    v, f = load_mesh("cow")
    v = normalize_mesh(v)
    vt, ft = cow_uvs()
    tex = torch.rand(1, 64, 64, 3)
    texd = tex.to(DEV).requires_grad_(True)
    mesh = trb.Meshes(verts=[v.to(DEV)], faces=[f.to(DEV)],
                      textures=trb.TexturesUV(maps=texd, faces_uvs=[ft.to(DEV)], verts_uvs=[vt.to(DEV)]))
    R, T = _views(1, seed=1)
    cameras = trb.FoVPerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV))
    renderer = trb.MeshRenderer(
        rasterizer=trb.MeshRasterizer(cameras=cameras, raster_settings=trb.RasterizationSettings(image_size=64)),
        shader=trb.SoftPhongShader(device=DEV, cameras=cameras, lights=trb.AmbientLights(device=DEV)))
    img = renderer(mesh)
    frag = renderer.rasterizer(mesh)
    # oracle: interpolate uvs, bilinear sample of the flipped map (A7)
    p2f, bary = frag.pix_to_face.cpu(), frag.bary_coords.cpu()
    uv = sref.interpolate_face_attributes(p2f, bary, vt[ft])
    grid = (uv[:, :, :, 0] * 2 - 1)
    maps = torch.flip(tex.permute(0, 3, 1, 2), [2])
    texels = torch.nn.functional.grid_sample(maps, grid, mode="bilinear", align_corners=True, padding_mode="border")
    texels = texels.permute(0, 2, 3, 1)
    covered = p2f[..., 0] >= 0
    assert covered.sum() > 100
    assert torch.allclose(img.detach().cpu()[..., :3][covered], texels[covered], atol=1e-4)
    img[..., :3].sum().backward()
    assert texd.grad is not None and texd.grad.abs().sum() > 0


def test_camera_models_and_kwargs_override():
    """PerspectiveCameras NDC / screen-space (in_ndc=False) and per-call R=, T= overrides reach the
    kernels: depth images agree with the oracle run on the torch-restated NDC vertices."""
    trb = _trb()
    v, f = _scene("teapot")
    H, W = 60, 80
    R, T = _views(2, seed=4)
    K = torch.tensor([[90.0, 0.0, 38.0], [0.0, 95.0, 31.0], [0.0, 0.0, 1.0]])[None]
    focal = torch.stack([K[:, 0, 0], K[:, 1, 1]], dim=-1)
    pp = K[:, :2, 2]
    cams = trb.PerspectiveCameras(focal_length=focal, principal_point=pp, device=DEV, in_ndc=False,
                                  image_size=torch.tensor([[H, W]]))
    rasterizer = trb.MeshRasterizer(cameras=cams, raster_settings=trb.RasterizationSettings(image_size=(H, W)))
    meshes = trb.Meshes(verts=[v.to(DEV)], faces=[f.to(DEV)]).extend(2)
    frag = rasterizer(meshes, R=R.to(DEV), T=T.to(DEV))
    s = min(H, W) / 2.0
    proj = torch.tensor([[90.0 / s, 95.0 / s, -(38.0 - W / 2) / s, -(31.0 - H / 2) / s]]).repeat(2, 1)
    ndc = _ndc(v, R, T, proj)
    got_ndc = rasterizer.transform(meshes, R=R.to(DEV), T=T.to(DEV)).cpu().reshape(2, -1, 3)
    assert torch.allclose(got_ndc, ndc, atol=2e-6, rtol=2e-6)
    want = oracle_rasterize(got_ndc, f, (H, W), 0.0, 1, True, False)
    _assert_fragments_equal((frag.pix_to_face, frag.zbuf, frag.bary_coords, frag.dists), want)
    assert (want[0] >= 0).sum() > 50
    # wrong camera count -> ValueError like upstream
    with pytest.raises(ValueError):
        rasterizer(meshes, R=torch.eye(3, device=DEV)[None].repeat(3, 1, 1), T=torch.zeros(3, 3, device=DEV))


def test_cpu_tensors_fail_loudly():
    trb = _trb()
    v, f = _scene("teapot")
    meshes = trb.Meshes(verts=[v], faces=[f])
    with pytest.raises(RuntimeError):
        trb.renderer.rasterize_meshes(meshes, 16, 0.0, 1)


def test_full_size_properties():
    """BASELINE config C2 size (cow, 64 views, 512^2, K=1) -- size-independent properties: determinism,
    zbuf == bary . z of the named face, K=1 result is layer 0 of K=2, images in range, silhouette alpha."""
    trb = _trb()
    from torch_renderer_b200.cameras import look_at_view_transform
    v, f = load_mesh("cow")
    N = 64
    elev = torch.linspace(0, 360, N)
    azim = torch.linspace(-180, 180, N)
    R, T = look_at_view_transform(dist=0.7, elev=elev, azim=azim)
    vd = v.to(DEV)
    mesh = trb.Meshes(verts=[vd], faces=[f.to(DEV)], textures=trb.TexturesVertex(torch.ones(1, v.shape[0], 3, device=DEV)))
    meshes = mesh.extend(N)
    cameras = trb.FoVPerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV))
    rast1 = trb.MeshRasterizer(cameras, trb.RasterizationSettings(image_size=512, faces_per_pixel=1))
    rast2 = trb.MeshRasterizer(cameras, trb.RasterizationSettings(image_size=512, faces_per_pixel=2))
    a, b, c = rast1(meshes), rast1(meshes), rast2(meshes)
    assert torch.equal(a.pix_to_face, b.pix_to_face) and torch.equal(a.zbuf, b.zbuf)
    assert torch.equal(a.bary_coords, b.bary_coords) and torch.equal(a.dists, b.dists)
    assert torch.equal(a.pix_to_face[..., 0], c.pix_to_face[..., 0]) and torch.equal(a.zbuf[..., 0], c.zbuf[..., 0])
    covered = a.pix_to_face[..., 0] >= 0
    frac = covered.float().mean().item()
    assert 0.01 < frac < 0.9
    # pix_to_face of view n lies in [n*F, (n+1)*F)
    F = f.shape[0]
    view_of = torch.arange(N, device=DEV).view(N, 1, 1).expand(N, 512, 512)
    assert torch.equal((a.pix_to_face[..., 0] // F)[covered], view_of[covered])
    # zbuf = sum_i bary_i * z_i
    ndc = rast1.transform(meshes).reshape(N, -1, 3)
    local = (a.pix_to_face[..., 0] % F)[covered]
    zs = ndc[view_of[covered][:, None], f.to(DEV)[local], 2]
    z = (a.bary_coords[..., 0, :][covered] * zs).sum(-1)
    assert torch.allclose(z, a.zbuf[..., 0][covered], atol=1e-5, rtol=1e-5)
    assert (a.bary_coords[..., 0, :][covered] > 0).all()
    assert (a.dists[..., 0][covered] <= 0).all()
    shader = trb.SoftPhongShader(device=DEV, cameras=cameras, lights=trb.PointLights(device=DEV, location=[[0, 0, -3.0]]))
    img = shader(a, meshes)
    assert torch.isfinite(img).all() and img.min() >= 0 and img[..., :3].max() <= 1.0 + 1e-5
    assert torch.equal(img[..., 3] > 0, covered)
    sil = trb.SoftSilhouetteShader()(a, meshes)
    assert torch.equal(sil[..., 3] >= 0.5, covered)


def test_fused_renderer_kwargs_cameras_and_separate_shader_camera():
    """camera_pose_optimizer.py pattern: one camera object shared by rasteriser and shader, pose passed
    as R=, T= kwargs with requires_grad; and the mesh_deformer.py pattern: cameras= / lights= kwargs.
    Fused and modular paths must agree on images and gradients."""
    trb = _trb()
    torch.manual_seed(0)
    v, f = _scene("teapot")
    colors = torch.rand(v.shape[0], 3)
    R0, T0 = _views(2, seed=12)
    res = {}
    for fused in (False, True):
        R = R0.to(DEV).requires_grad_(True)
        T = T0.to(DEV).requires_grad_(True)
        vd = v.to(DEV).requires_grad_(True)
        cams = trb.FoVPerspectiveCameras(device=DEV)
        lights = trb.PointLights(device=DEV, location=[[0.0, 0.0, -3.0]])
        mesh = trb.Meshes(verts=[vd], faces=[f.to(DEV)], textures=trb.TexturesVertex(colors.to(DEV)[None])).extend(2)
        rast = trb.MeshRasterizer(cameras=cams, raster_settings=trb.RasterizationSettings(image_size=64))
        shader = trb.SoftPhongShader(device=DEV, cameras=cams, lights=lights,
                                     blend_params=trb.BlendParams(1e-4, 1e-4, (0, 0, 0)))
        if fused:
            img = trb.MeshRenderer(rast, shader)(mesh, R=R, T=T)
        else:
            frag = rast(mesh, R=R, T=T)
            img = shader(frag, mesh, R=R, T=T)
        (img[..., :3] ** 2).sum().backward()
        res[fused] = (img.detach(), R.grad.clone(), T.grad.clone(), vd.grad.clone())
    assert torch.allclose(res[False][0], res[True][0], atol=1e-6)
    for a, b in zip(res[False][1:], res[True][1:]):
        assert rel_l2(a, b) < 1e-4
    # cameras= and lights= kwargs with a different camera object in the shader
    cams_a = trb.PerspectiveCameras(device=DEV, R=R0.to(DEV), T=T0.to(DEV))
    cams_b = trb.PerspectiveCameras(device=DEV, R=R0[:1].to(DEV), T=T0[:1].to(DEV))
    mesh = trb.Meshes(verts=[v.to(DEV)], faces=[f.to(DEV)], textures=trb.TexturesVertex(colors.to(DEV)[None])).extend(2)
    rast = trb.MeshRasterizer(cameras=cams_b, raster_settings=trb.RasterizationSettings(image_size=48, perspective_correct=False))
    shader = trb.SoftPhongShader(device=DEV, cameras=cams_b, lights=trb.AmbientLights(device=DEV))
    lights = trb.PointLights(device=DEV, location=[[2.0, 2.0, -2.0]])
    img_f = trb.MeshRenderer(rast, shader)(mesh, cameras=cams_a, lights=lights)
    frag = rast(mesh, cameras=cams_a, lights=lights)
    img_m = shader(frag, mesh, cameras=cams_a, lights=lights)
    assert torch.allclose(img_f, img_m, atol=1e-6)
    assert (frag.pix_to_face >= 0).sum() > 100


def test_fused_rasterizer_depth_gradient_and_fragment_grads():
    """torch_renderer.DepthRender pattern: loss on relu(zbuf[...,0]) from the bare rasteriser, plus a
    renderer-with-fragments loss that mixes image and fragment gradients."""
    trb = _trb()
    s = _render_setup("teapot", 2, (48, 48), 2, 1e-3, "soft_phong", "point", seed=8, background=(0, 0, 0))
    out = _cuda_render(s, requires_grad=True, fused=True)
    fr = out["fragments"]
    torch.manual_seed(3)
    wz, wb, wd = torch.rand(2, 48, 48, 2), torch.rand(2, 48, 48, 2, 3), torch.rand(2, 48, 48, 2)
    m = (fr.pix_to_face >= 0).cpu()
    loss = (out["images"] ** 2).sum() + (fr.zbuf * (wz * m).to(DEV)).sum() + \
        (fr.bary_coords * (wb * m[..., None]).to(DEV)).sum() + (fr.dists * (wd * m).to(DEV)).sum() * 100
    loss.backward()
    ref = _oracle_render(s, fr.pix_to_face.cpu(), torch.float64)
    loss_ref = (ref["images"] ** 2).sum() + (ref["zbuf"] * wz * m).sum() + (ref["bary"] * wb * m[..., None]).sum() + \
        (ref["dists"] * wd * m).sum() * 100
    loss_ref.backward()
    for k in ("v", "R", "T", "c"):
        e = rel_l2(out[k].grad.cpu(), ref[k].grad)
        assert e < 1e-3, f"grad {k}: rel L2 err {e}"


def test_z_clip_culls_faces_entirely_nearer_than_the_plane():
    """FoV cameras clip at znear/2 by default (upstream semantics); an explicit z_clip_value overrides.
    Faces with all three vertices nearer than the plane disappear; the oracle sees the same face list
    with those faces removed."""
    trb = _trb()
    v, f = uv_sphere(16, 20, 1.0, noise=0.02, seed=6)
    R, T = _views(2, dist=1.6, seed=21)
    z_clip = 1.0
    mesh = trb.Meshes(verts=[v.to(DEV)], faces=[f.to(DEV)]).extend(2)
    cams = trb.FoVPerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV))
    rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=96, faces_per_pixel=2, z_clip_value=z_clip))
    # "off": no clip_faces route -- this test pins what the fused kernels themselves do with the plane
    # (tests/test_gpu_clip.py covers the default, upstream-exact route that cuts the crossing faces)
    trb.set_near_plane_clipping("off")
    try:
        frag = rast(mesh)
    finally:
        trb.set_near_plane_clipping("exact")
    ndc = rast.transform(mesh).cpu().reshape(2, -1, 3)
    F = f.shape[0]
    fz = ndc[:, f][..., 2]                       # [2, F, 3] view depths of the face corners
    keep = ~(fz < z_clip).all(dim=-1)            # [2, F]
    assert (~keep).sum() > 10 and keep.sum() > 100
    fv, first, count, remap = [], [], [], []
    for n in range(2):
        idx = keep[n].nonzero().flatten()
        first.append(sum(count)); count.append(len(idx))
        fv.append(ndc[n][f[idx]])
        remap.append(idx + n * F)
    want = oracle.rasterize_forward(torch.cat(fv).numpy(), np.array(first, np.int64), np.array(count, np.int64),
                                    (96, 96), 0.0, 2, True, False, False)
    remap = torch.cat(remap).numpy()
    want_p2f = np.where(want[0] >= 0, remap[np.clip(want[0], 0, None)], -1)
    _assert_fragments_equal((frag.pix_to_face, frag.zbuf, frag.bary_coords, frag.dists),
                            (want_p2f, want[1], want[2], want[3]))
    # without clipping the nearer cap is visible instead
    frag0 = trb.MeshRasterizer(trb.PerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV), focal_length=1.7320508),
                               trb.RasterizationSettings(image_size=96, faces_per_pixel=2))(mesh)
    assert (frag0.pix_to_face != frag.pix_to_face).any()


def test_fragment_cache_reference_step_pattern():
    """camera_pose_optimizer.py:237-254: rasterizer(meshes, R=R, T=T).zbuf, then the silhouette renderer, then the
    Phong renderer on the same meshes / R / T / settings.  With the Fragments cache the scene is rasterised once
    per step; values and pose gradients must equal the uncached run, and nothing may leak into the next step."""
    trb = _trb()
    from torch_renderer_b200.rasterizer import _fragment_cache, set_fragment_cache
    v, f = _scene("teapot")
    cols = torch.rand(1, v.shape[0], 3, generator=torch.Generator().manual_seed(1))
    mesh = trb.Meshes([v.to(DEV)], [f.to(DEV)], textures=trb.TexturesVertex(cols.to(DEV)))
    cams = trb.FoVPerspectiveCameras(device=DEV)
    settings = trb.RasterizationSettings(image_size=128, blur_radius=0.0, faces_per_pixel=1)
    blend = trb.BlendParams(1e-4, 1e-4, (0, 0, 0))
    rast = trb.MeshRasterizer(cams, settings)
    sil = trb.MeshRenderer(trb.MeshRasterizer(cams, settings), trb.SoftSilhouetteShader(blend))
    phong = trb.MeshRenderer(trb.MeshRasterizer(cams, settings),
                             trb.SoftPhongShader(device=DEV, cameras=cams, blend_params=blend,
                                                 lights=trb.PointLights(device=DEV, location=[[0.0, 0.0, -3.0]])))
    R0, T0 = trb.look_at_view_transform(2.7, 30, 60)
    q0 = torch.cat([T0, trb.transforms.matrix_to_quaternion(R0)], -1).to(DEV)

    def step(pose):
        R = trb.transforms.quaternion_to_matrix(pose[:, 3:]); T = pose[:, :3]
        depth = torch.relu(rast(meshes_world=mesh, R=R, T=T).zbuf[..., 0])
        alpha = sil(mesh, R=R, T=T)[..., 3]
        rgb = phong(mesh, R=R, T=T)[..., :3]
        loss = depth.mean() + alpha.mean() + (rgb ** 2).mean()
        loss.backward()
        return depth.detach(), alpha.detach(), rgb.detach(), pose.grad.clone()

    set_fragment_cache(False)
    want = step(q0.clone().requires_grad_(True))
    set_fragment_cache(True)
    pose = q0.clone().requires_grad_(True)
    h0 = _fragment_cache.hits
    got = step(pose)
    assert _fragment_cache.hits - h0 == 2          # silhouette + Phong reused the rasteriser's Fragments
    for a, b in zip(got[:3], want[:3]):
        assert torch.allclose(a, b, atol=1e-4, rtol=0)
    assert rel_l2(got[3].cpu(), want[3].cpu()) < 1e-3
    # next step, same tensors: the stored graph is spent, so the scene is rasterised again (no stale reuse)
    pose.grad = None
    h1 = _fragment_cache.hits
    got2 = step(pose)
    assert _fragment_cache.hits - h1 == 2
    assert rel_l2(got2[3].cpu(), want[3].cpu()) < 1e-3
    # an in-place update of an input is a different scene
    with torch.no_grad():
        frag_a = rast(mesh, R=R0.to(DEV), T=T0.to(DEV))
        Rm, Tm = R0.to(DEV), T0.to(DEV)
        fa = rast(mesh, R=Rm, T=Tm)
        fb = rast(mesh, R=Rm, T=Tm)
        assert fb.zbuf is fa.zbuf                   # identical inputs: reused
        Tm.add_(0.2)
        fc = rast(mesh, R=Rm, T=Tm)
        assert fc.zbuf is not fa.zbuf and not torch.equal(fc.zbuf, fa.zbuf)
    set_fragment_cache(False)


@pytest.mark.parametrize("K,blur,lights", [(1, 0.0, "point"), (4, 1e-3, "point"), (1, 0.0, "ambient")])
def test_textures_uv_fused_in_kernels(K, blur, lights):
    """TexturesUV sampled inside the fused kernels (the cow asset of camera_pose_optimizer.py /
    deform_mesh_with_color.py:266-271,329) against the composed path -- rasterise, CUDA UV interpolation,
    torch grid_sample(flip(map), 2uv-1, bilinear, align_corners=True, border), shade with texels: same images,
    same gradients w.r.t. the texture map, the vertices and the per-view R / T."""
    trb = _trb()
    torch.manual_seed(0)
    v, f = load_mesh("cow")
    v = normalize_mesh(v)
    vt, ft = cow_uvs()
    tex = torch.rand(1, 48, 80, 3)
    R, T = _views(2, seed=4)

    def run(fused):
        texd = tex.to(DEV).requires_grad_(True)
        vd = v.to(DEV).requires_grad_(True)
        Rd, Td = R.to(DEV).requires_grad_(True), T.to(DEV).requires_grad_(True)
        mesh = trb.Meshes(verts=[vd], faces=[f.to(DEV)],
                          textures=trb.TexturesUV(maps=texd, faces_uvs=[ft.to(DEV)], verts_uvs=[vt.to(DEV)])).extend(2)
        cameras = trb.FoVPerspectiveCameras(device=DEV, R=Rd, T=Td)
        lt = (trb.AmbientLights(device=DEV) if lights == "ambient"
              else trb.PointLights(device=DEV, location=[[0.5, 1.0, -2.5]]))
        rast = trb.MeshRasterizer(cameras=cameras, raster_settings=trb.RasterizationSettings(
            image_size=(72, 96), blur_radius=blur, faces_per_pixel=K))
        shader = trb.SoftPhongShader(device=DEV, cameras=cameras, lights=lt)
        if fused:
            from torch_renderer_b200 import ops
            n0 = ops.launch_count()
            img = trb.MeshRenderer(rast, shader)(mesh)
            # 5 fused stages + the capacity statistics + the near-plane question (FoV camera: z_clip = znear / 2)
            assert ops.launch_count() - n0 <= 7, "TexturesUV did not take the fused path"
        else:
            img = shader(rast(mesh), mesh)
        w = torch.linspace(0.5, 1.5, 72 * 96 * 4, device=DEV).reshape(1, 72, 96, 4)
        (img * w).sum().backward()
        return img.detach(), texd.grad, vd.grad, Rd.grad, Td.grad

    a, b = run(True), run(False)
    assert (a[0][..., 3] > 0).sum() > 500
    assert torch.allclose(a[0], b[0], atol=1e-4, rtol=0), (a[0] - b[0]).abs().max()
    for name, x, y in zip(("texture map", "verts", "R", "T"), a[1:], b[1:]):
        assert x is not None and y is not None, name
        assert rel_l2(x.cpu(), y.cpu()) < 1e-3, (name, rel_l2(x.cpu(), y.cpu()))
    assert a[1].abs().sum() > 0


def _triangle_soup(seed, n_small=300, n_big=6):
    """Random triangle soup in NDC: many small faces, a few screen-sized ones, exact depth ties (duplicated faces
    and faces quantised to a coarse depth grid), degenerate and zero-area faces, faces partly / wholly outside the
    image, faces at and behind the camera plane."""
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(n_small, 1, 2, generator=g) * 2.6 - 1.3
    small = torch.cat([c + 0.12 * torch.randn(n_small, 3, 2, generator=g),
                       torch.round(torch.rand(n_small, 1, 1, generator=g).expand(-1, 3, -1) * 8) / 4 + 0.5
                       + 0.05 * torch.randn(n_small, 3, 1, generator=g) * (torch.rand(n_small, 1, 1, generator=g) > 0.5)], -1)
    big = torch.cat([torch.rand(n_big, 3, 2, generator=g) * 3.0 - 1.5, 1.0 + torch.rand(n_big, 3, 1, generator=g) * 2], -1)
    dup = small[:20].clone()                                  # exact ties: same geometry, later face index
    flat = small[20:30].clone(); flat[:, 2] = flat[:, 1]      # zero area
    behind = small[30:40].clone(); behind[:, :, 2] = -behind[:, :, 2]
    touch = small[40:50].clone(); touch[:, 0, 2] = 0.0        # one vertex on the camera plane
    tris = torch.cat([small, big, dup, flat, behind, touch], 0)
    perm = torch.randperm(tris.shape[0], generator=g)
    verts = tris[perm].reshape(-1, 3).float()
    return verts, torch.arange(verts.shape[0]).reshape(-1, 3)


@pytest.mark.parametrize("seed,image_size,K,blur,persp,clip,cull", [
    (0, (48, 64), 1, 0.0, False, False, False), (1, (48, 64), 1, 0.0, True, False, True),
    (2, (37, 53), 1, 2e-3, True, True, False), (3, (64, 48), 2, 0.0, True, False, False),
    (4, (40, 40), 5, 1e-3, False, True, False), (5, (33, 70), 5, 1e-3, True, True, True),
    (6, (32, 32), 30, 3e-3, True, True, False), (7, (56, 56), 8, 0.0, False, False, False),
    (8, (24, 100), 3, 5e-4, True, False, False), (9, (50, 50), 150, 2e-3, False, True, False),
])
def test_random_triangle_soups_bit_exact(seed, image_size, K, blur, persp, clip, cull):
    """Adversarial soups (exact depth ties, degenerate faces, big + tiny faces mixed, faces across the image border
    and the camera plane) against the oracle: pix_to_face bit-exact for every kernel variant."""
    verts, faces = _triangle_soup(seed)
    want = oracle_rasterize(verts[None], faces, image_size, blur, K, persp, clip, cull)
    got = _cuda_raster_from_ndc(verts[None], faces, image_size, blur, K, persp, clip, cull)
    _assert_fragments_equal(got, want)
    assert (want[0] >= 0).sum() > 200
    if K > 1:
        assert (want[0][..., 1] >= 0).sum() > 50


def test_integration_md_ctypes_stub_matches_oracle(monkeypatch):
    """The drop-in `rasterize_meshes` ctypes stub printed in INTEGRATION.md (what a maintainer of the reference
    would paste in place of pytorch3d._C.rasterize_meshes) is executed as written and checked against the oracle."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    code = next(b for b in blocks if "def rasterize_meshes(" in b and "ctypes.CDLL" in b)
    monkeypatch.chdir(root)
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    v, f = _scene("teapot")
    R, T = _views(2, seed=3)
    ndc = _ndc(v, R, T, fov_proj(2))
    F = f.shape[0]
    face_verts = ndc[:, f].reshape(-1, 3, 3).contiguous().to(DEV)
    first = torch.arange(2, dtype=torch.int64) * F
    count = torch.full((2,), F, dtype=torch.int64)
    neighbours = torch.full((2 * F,), -1, dtype=torch.int64)
    for K, blur, persp, clip in ((1, 0.0, True, False), (4, 1e-3, True, True)):
        got = ns["rasterize_meshes"](face_verts, first.to(DEV), count.to(DEV), neighbours.to(DEV), (48, 64), blur, K,
                                     0, 0, persp, clip, False)
        want = oracle_rasterize(ndc, f, (48, 64), blur, K, persp, clip)
        _assert_fragments_equal(got, want)
    # ... and with faces cut by the near plane: the stub forwards clipped_faces_neighbor_idx to trb_clip_resequence
    from oracle import clip_ref
    g = torch.Generator().manual_seed(5)
    soup = torch.cat([torch.rand(90, 3, 2, generator=g) * 1.6 - 0.8 + (torch.rand(90, 1, 2, generator=g) - 0.5),
                      torch.rand(90, 3, 1, generator=g) * 1.6 + 0.05], dim=2)
    cf = clip_ref.clip_faces(soup.numpy(), [0, 40], [40, 50], clip_ref.rasterizer_frustum(True, 0.5, False))
    assert (cf.clipped_faces_neighbor_idx >= 0).sum() > 20
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    got = ns["rasterize_meshes"](t(cf.face_verts), t(cf.mesh_to_face_first_idx), t(cf.num_faces_per_mesh),
                                 t(cf.clipped_faces_neighbor_idx), (40, 56), 2e-3, 3, 0, 0, True, True, False)
    want = oracle.rasterize_forward(cf.face_verts, cf.mesh_to_face_first_idx, cf.num_faces_per_mesh, (40, 56), 2e-3, 3,
                                    True, True, False, 0, clipped_faces_neighbor_idx=cf.clipped_faces_neighbor_idx)
    _assert_fragments_equal(got, want)


@pytest.mark.parametrize("K,blur", [(1, 0.0), (6, 1e-3)])
def test_forward_is_bit_reproducible_under_repetition(K, blur):
    """Thirty repetitions of a render whose CTAs each rasterise several busy tiles (K = 1: a screen-filling mesh, 8
    tiles per CTA) must give bit-identical Fragments and images every time: shared-memory reuse between the tiles of
    one CTA, the compacted epilogue and the depth-ordered walk leave no room for a race to hide."""
    trb = _trb()
    v, f = uv_sphere(40, 60, 1.0, noise=0.04, seed=5)
    R, T = trb.look_at_view_transform(1.9, torch.tensor([15.0, -35.0]), torch.tensor([30.0, 190.0]))
    mesh = trb.Meshes([v.to(DEV)], [f.to(DEV)], textures=trb.TexturesVertex(torch.rand(1, v.shape[0], 3, device=DEV))).extend(2)
    cams = trb.FoVPerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV))
    rend = trb.MeshRendererWithFragments(
        trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=(208, 256), blur_radius=blur, faces_per_pixel=K)),
        trb.SoftPhongShader(device=DEV, cameras=cams, lights=trb.PointLights(device=DEV, location=[[0.0, 2.0, -3.0]])))
    img0, fr0 = rend(mesh)
    assert (fr0.pix_to_face[..., 0] >= 0).float().mean() > 0.5
    for _ in range(30):
        img, fr = rend(mesh)
        assert torch.equal(fr.pix_to_face, fr0.pix_to_face) and torch.equal(fr.zbuf, fr0.zbuf)
        assert torch.equal(fr.bary_coords, fr0.bary_coords) and torch.equal(fr.dists, fr0.dists)
        assert torch.allclose(img, img0, atol=1e-5, rtol=0)  # vertex normals are accumulated with fp32 atomics: order noise ~2e-6


@pytest.mark.parametrize("shader_kind,K,blur,size", [
    ("soft_phong", 1, 0.0, (64, 64)), ("soft_phong", 1, 0.0, (45, 61)), ("hard_phong", 1, 0.0, (96, 96)),
    ("soft_silhouette", 1, 0.0, (64, 64)), ("soft_phong", 8, 9.21024e-4, (128, 128)), ("soft_phong", 5, 2e-3, (45, 61)),
    ("soft_silhouette", 50, 9.21024e-4, (96, 96)), ("soft_silhouette", 30, 2e-3, (33, 47)), ("hard_phong", 4, 0.0, (64, 64)),
    ("soft_silhouette", 10, 9.21024e-4, (128, 128)), ("soft_silhouette", 3, 1e-3, (45, 61)),
])
def test_image_only_renderer_sparse_fragments_equal_dense(shader_kind, K, blur, size):
    """``MeshRenderer`` returns the image only, so its kernels write Fragments for covered pixels only
    (trb_render_config.sparse_fragments); ``MeshRendererWithFragments`` writes PyTorch3D's dense layout.  Same
    image (bit for bit where no atomically accumulated vertex normal enters), same gradients (the scatter order of
    the atomics is the only difference).  The allocator is poisoned with NaNs before the sparse render, so a read of
    an unwritten background sample would show."""
    trb = _trb()
    torch.manual_seed(1)
    v, f = _scene("teapot")
    colors = torch.rand(v.shape[0], 3)
    N = 3
    R0, T0 = _views(N, seed=5)
    out = {}
    for dense in (True, False):
        vd, cd = v.to(DEV).requires_grad_(True), colors.to(DEV).requires_grad_(True)
        R, T = R0.to(DEV).requires_grad_(True), T0.to(DEV).requires_grad_(True)
        cams = trb.FoVPerspectiveCameras(device=DEV)
        mesh = trb.Meshes(verts=[vd], faces=[f.to(DEV)], textures=trb.TexturesVertex(cd[None])).extend(N)
        rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=size, blur_radius=blur, faces_per_pixel=K))
        if shader_kind == "soft_silhouette":
            shader = trb.SoftSilhouetteShader(trb.BlendParams(1e-4, 1e-4, (0, 0, 0)))
        else:
            cls = trb.SoftPhongShader if shader_kind == "soft_phong" else trb.HardPhongShader
            shader = cls(device=DEV, cameras=cams, lights=trb.PointLights(device=DEV, location=[[0.0, 0.0, -3.0]]))
        # a poisoned allocator would show any read of an unwritten background sample as NaN / garbage
        if dense:
            img, frag = trb.MeshRendererWithFragments(rast, shader)(mesh, R=R, T=T)
            assert (frag.pix_to_face >= -1).all()
        else:
            junk = torch.full((N * size[0] * size[1] * K * 8,), float("nan"), device=DEV)
            del junk
            img = trb.MeshRenderer(rast, shader)(mesh, R=R, T=T)
        # the fine kernel's own per-view sums of the alpha channel (ops.render: images.alpha_sum), both layouts
        want_alpha = img.detach()[..., 3].double().sum((1, 2))
        assert img.alpha_sum.shape == (N,) and not img.alpha_sum.requires_grad
        assert ((img.alpha_sum.double() - want_alpha).abs() <= 1e-5 * want_alpha.abs() + 1e-4).all()
        torch.manual_seed(2)
        w = torch.rand(img.shape, device=DEV)
        (img * w).sum().backward()
        out[dense] = (img.detach().clone(), vd.grad.clone(), R.grad.clone(), T.grad.clone(),
                      cd.grad.clone() if cd.grad is not None else torch.zeros(1, device=DEV))
    if shader_kind == "soft_silhouette":
        assert torch.equal(out[True][0], out[False][0])
    else:
        # vertex normals are accumulated with atomics (order of the float sums differs from run to run): ulps
        assert torch.equal(out[True][0][..., 3], out[False][0][..., 3])
        assert (out[True][0] - out[False][0]).abs().max() < 1e-5
    for a, b in zip(out[True][1:], out[False][1:]):
        assert rel_l2(a, b) < 1e-4      # two runs of the same kernels: atomics order + the normals' ulps through the blend


def test_k1_many_screen_sized_faces_item_table_overflow():
    """K=1 tile raster with more (face, pixel) work items per staging chunk than its per-item face table holds
    (40 screen-sized faces over every 16x16 tile = 10,240 items > 4,096): the binary-search path, against the oracle;
    and with a few hundred small faces mixed in (both paths in one image)."""
    g = torch.Generator().manual_seed(21)
    big = torch.cat([torch.rand(40, 3, 2, generator=g) * 3.0 - 1.5, 1.0 + torch.rand(40, 3, 1, generator=g) * 2], -1)
    c = torch.rand(400, 1, 2, generator=g) * 2.0 - 1.0
    small = torch.cat([c + 0.05 * torch.randn(400, 3, 2, generator=g), 0.8 + torch.rand(400, 3, 1, generator=g)], -1)
    for tris in (big, torch.cat([big[:25], small])):
        verts = tris.reshape(-1, 3).float()
        faces = torch.arange(verts.shape[0]).reshape(-1, 3)
        for persp in (False, True):
            want = oracle_rasterize(verts[None], faces, (64, 80), 0.0, 1, persp, False)
            got = _cuda_raster_from_ndc(verts[None], faces, (64, 80), 0.0, 1, persp, False)
            _assert_fragments_equal(got, want)
            assert (want[0] >= 0).mean() > 0.5
