"""GPU parity at the shapes BASELINE.json's configs name: C1 teapot 256^2 SoftPhong forward,
C3 soft silhouette K=50 sigma=1e-4 with a 7-vector pose, C4 ico-sphere deformation (perspective_correct=False,
ambient and point lights), C5 1M-face sphere K=8 with the soft blur.  Where the oracle finishes in seconds the
comparison is exact; at full C5 size it is size-independent properties."""
import math

import numpy as np
import pytest
import torch

from oracle import shading_ref as sref
from helpers import fov_proj, load_mesh, normalize_mesh, oracle_rasterize, rel_l2

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
SIGMA = 1e-4
BLUR = math.log(1.0 / 1e-4 - 1.0) * SIGMA  # 9.21024e-4, the reference's soft-raster setting


def _trb():
    import torch_renderer_b200 as trb
    return trb


def _grid_sphere(nlat, nlon, seed=0):
    g = torch.Generator().manual_seed(seed)
    th = torch.linspace(0, math.pi, nlat + 1)[1:-1]
    ph = torch.linspace(0, 2 * math.pi, nlon + 1)[:-1]
    T, P = torch.meshgrid(th, ph, indexing="ij")
    ring = torch.stack([torch.sin(T) * torch.cos(P), torch.cos(T), torch.sin(T) * torch.sin(P)], -1).reshape(-1, 3)
    v = torch.cat([torch.tensor([[0.0, 1.0, 0.0]]), ring, torch.tensor([[0.0, -1.0, 0.0]])])
    v = v * (1 + 0.05 * torch.randn(v.shape[0], 1, generator=g))
    idx = lambda r, s: 1 + r * nlon + (s % nlon)
    r = torch.arange(nlat - 2)[:, None]
    s_ = torch.arange(nlon)[None, :]
    a, b, c, d = idx(r, s_), idx(r, s_ + 1), idx(r + 1, s_), idx(r + 1, s_ + 1)
    quads = torch.cat([torch.stack([a, b, c], -1).reshape(-1, 3), torch.stack([b, d, c], -1).reshape(-1, 3)])
    s1 = torch.arange(nlon)
    top = torch.stack([torch.zeros_like(s1), idx(0, s1 + 1), idx(0, s1)], -1)
    bot = torch.stack([torch.full_like(s1, v.shape[0] - 1), idx(nlat - 2, s1), idx(nlat - 2, s1 + 1)], -1)
    return v.float(), torch.cat([top, quads, bot]).long()


def _check_fragments(frag, want):
    p2f = frag.pix_to_face.cpu().numpy()
    mism = int((p2f != want[0]).sum())
    assert mism == 0, f"pix_to_face differs at {mism} of {p2f.size} samples"
    for name, a, b in (("zbuf", frag.zbuf, want[1]), ("bary", frag.bary_coords, want[2]), ("dists", frag.dists, want[3])):
        a = a.detach().cpu().numpy()
        assert np.allclose(a, b, atol=1e-5, rtol=1e-5), f"{name} max abs diff {np.abs(a - b).max()}"


def test_c1_teapot_256_softphong_forward():
    """renderer_comparison_with_pyrender.py:166-220 pattern on the in-repo teapot (the reference's CPU config)."""
    trb = _trb()
    v, f = load_mesh("teapot")
    R, T = trb.look_at_view_transform(2.7, 10, 20)
    mesh = trb.Meshes([v.to(DEV)], [f.to(DEV)], textures=trb.TexturesVertex(torch.ones(1, v.shape[0], 3, device=DEV)))
    cams = trb.FoVPerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV))
    rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=256, blur_radius=0.0, faces_per_pixel=1))
    shader = trb.SoftPhongShader(device=DEV, cameras=cams, lights=trb.PointLights(device=DEV, location=[[0.0, 0.0, -3.0]]))
    images, frag = trb.MeshRendererWithFragments(rast, shader)(mesh)
    ndc = rast.transform(mesh).cpu().reshape(1, -1, 3)
    want = oracle_rasterize(ndc, f, (256, 256), 0.0, 1, True, False)
    _check_fragments(frag, want)
    assert (want[0] >= 0).mean() > 0.05
    ones = lambda *x: torch.tensor([list(x)], dtype=torch.float64)
    p2f = torch.from_numpy(want[0])
    v64 = v.double()
    ref = sref.shade(p2f, torch.from_numpy(want[2]).double(), torch.from_numpy(want[1]).double(),
                     torch.from_numpy(want[3]).double(), f, v64, sref.vertex_normals(v64, f), torch.ones_like(v64),
                     shader="soft_phong", light_kind="point", light_vec=ones(0, 0, -3.0), light_ambient=ones(.5, .5, .5),
                     light_diffuse=ones(.3, .3, .3), light_specular=ones(.2, .2, .2), mat_ambient=ones(1, 1, 1),
                     mat_diffuse=ones(1, 1, 1), mat_specular=ones(1, 1, 1), shininess=torch.tensor([64.0], dtype=torch.float64),
                     camera_center=-torch.matmul(T[:, None, :].double(), torch.linalg.inv(R.double()))[:, 0, :])
    assert (images.cpu().double() - ref).abs().max() < 1e-4


@pytest.mark.parametrize("name", ["teapot", "cow"])
def test_c3_soft_silhouette_k50_pose_gradient(name):
    """camera_pose_optimizer.py:116-121 settings: blur = log(1/1e-4 - 1) * sigma, K=50, SoftSilhouette, pose stored
    as (T, quaternion).  Forward exact at 512^2; the pose gradient against fp64 autograd at 128^2."""
    trb = _trb()
    v, f = load_mesh(name)
    v = normalize_mesh(v)
    R0, T0 = trb.look_at_view_transform(2.7, 30, 60)
    pose0 = torch.cat([T0, trb.transforms.matrix_to_quaternion(R0)], -1)
    for size, check_grad in ((512, False), (128, True)):
        pose = pose0.to(DEV).requires_grad_(True)
        Rm = trb.transforms.quaternion_to_matrix(pose[:, 3:])
        Tm = pose[:, :3]
        mesh = trb.Meshes([v.to(DEV)], [f.to(DEV)])
        cams = trb.FoVPerspectiveCameras(device=DEV)
        rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=size, blur_radius=BLUR, faces_per_pixel=50))
        shader = trb.SoftSilhouetteShader(trb.BlendParams(SIGMA, 1e-4, (0, 0, 0)))
        images, frag = trb.MeshRendererWithFragments(rast, shader)(mesh, R=Rm, T=Tm)
        ndc = rast.transform(mesh, R=Rm, T=Tm).detach().cpu().reshape(1, -1, 3)
        want = oracle_rasterize(ndc, f, (size, size), BLUR, 50, True, True)
        _check_fragments(frag, want)
        assert (want[0][..., 1] >= 0).sum() > 100  # several layers really are in play
        if not check_grad:
            continue
        torch.manual_seed(0)
        target = (torch.rand(1, size, size) > 0.5).float()
        depth_w = torch.rand(1, size, size)
        loss = (images[..., 3] - target.to(DEV)).abs().mean() + (torch.relu(frag.zbuf[..., 0]) * depth_w.to(DEV)).mean()
        loss.backward()
        p64 = pose0.double().requires_grad_(True)
        R64 = trb.transforms.quaternion_to_matrix(p64[:, 3:])
        proj = fov_proj(1).double()
        ndc64 = sref.world_to_ndc(v.double(), R64, p64[:, :3], proj[:, 0], proj[:, 1], proj[:, 2], proj[:, 3], True)
        p2f = torch.from_numpy(want[0])
        z64, b64, d64 = sref.raster_recompute(ndc64[:, f].reshape(-1, 3, 3), p2f, True, True)
        img64 = sref.sigmoid_alpha_blend(torch.ones_like(b64), p2f, d64, SIGMA)
        loss64 = (img64[..., 3] - target.double()).abs().mean() + (torch.relu(z64[..., 0]) * depth_w.double()).mean()
        loss64.backward()
        assert abs(loss.item() - loss64.item()) < 1e-5
        assert rel_l2(pose.grad.cpu(), p64.grad) < 1e-3


@pytest.mark.parametrize("lights_kind", ["ambient", "point"])
def test_c4_ico_sphere_deformation_gradients(lights_kind):
    """mesh_deformer.py:135-145,181-222: ico-sphere + per-vertex offsets and colours with requires_grad, 5 views,
    perspective_correct=False, MSE to target images."""
    trb = _trb()
    ico = trb.ico_sphere(4, device=DEV)
    v0, f0 = ico.get_mesh_verts_faces(0)
    torch.manual_seed(0)
    deform = (0.02 * torch.randn_like(v0)).requires_grad_(True)
    rgb = torch.rand(1, v0.shape[0], 3, device=DEV).requires_grad_(True)
    NV, S = 5, 128
    R, T = trb.look_at_view_transform(dist=2.7, elev=torch.linspace(0, 360, NV), azim=torch.linspace(-180, 180, NV))
    cams = trb.PerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV))
    lights = trb.AmbientLights(device=DEV) if lights_kind == "ambient" else trb.PointLights(device=DEV, location=[[0.0, 0.0, 2.0]])
    rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=S, blur_radius=0.0, faces_per_pixel=1,
                                                              perspective_correct=False))
    renderer = trb.MeshRendererWithFragments(rast, trb.SoftPhongShader(device=DEV, cameras=cams, lights=lights))
    mesh = trb.Meshes([v0 + deform], [f0], textures=trb.TexturesVertex(rgb)).extend(NV)
    images, frag = renderer(mesh)
    target = torch.rand(NV, S, S, 3)
    loss = ((images[..., :3] - target.to(DEV)) ** 2).mean()
    loss.backward()
    # oracle
    f = f0.cpu()
    ndc = rast.transform(mesh).detach().cpu().reshape(NV, -1, 3)
    want = oracle_rasterize(ndc, f, (S, S), 0.0, 1, False, False)
    _check_fragments(frag, want)
    d64 = deform.detach().cpu().double().requires_grad_(True)
    c64 = rgb.detach().cpu().double()[0].requires_grad_(True)
    v64 = v0.cpu().double() + d64
    ones = lambda *x: torch.tensor([list(x)], dtype=torch.float64).repeat(NV, 1)
    proj = ones(1, 1, 0, 0)
    ndc64 = sref.world_to_ndc(v64, R.double(), T.double(), proj[:, 0], proj[:, 1], proj[:, 2], proj[:, 3], True)
    p2f = torch.from_numpy(want[0])
    z, b, d = sref.raster_recompute(ndc64[:, f].reshape(-1, 3, 3), p2f, False, False)
    cam = -torch.matmul(T[:, None, :].double(), torch.linalg.inv(R.double()))[:, 0, :]
    amb = ones(1, 1, 1) if lights_kind == "ambient" else ones(.5, .5, .5)
    img64 = sref.shade(p2f, b, z, d, f.repeat(NV, 1), v64, sref.vertex_normals(v64, f), c64, shader="soft_phong",
                       light_kind=lights_kind, light_vec=ones(0, 0, 2.0), light_ambient=amb, light_diffuse=ones(.3, .3, .3),
                       light_specular=ones(.2, .2, .2), mat_ambient=ones(1, 1, 1), mat_diffuse=ones(1, 1, 1),
                       mat_specular=ones(1, 1, 1), shininess=torch.full((NV,), 64.0, dtype=torch.float64),
                       camera_center=cam)
    assert (images.detach().cpu().double() - img64).abs().max() < 1e-4
    ((img64[..., :3] - target.double()) ** 2).mean().backward()
    assert rel_l2(rgb.grad.cpu()[0], c64.grad) < 1e-3
    if lights_kind == "ambient":
        # perspective_correct=False + ambient light: colours only depend on screen-space barycentrics
        assert rel_l2(deform.grad.cpu(), d64.grad) < 1e-3
    else:
        assert rel_l2(deform.grad.cpu(), d64.grad) < 1e-3


def test_c5_scaled_down_exact_and_full_size_properties():
    """C5: lat-long sphere with noisy radius, K=8, blur 9.21e-4, SoftPhong.  40k faces at 192^2 against the oracle;
    the full 1M faces at 1024^2 (2 views) through properties: determinism, depth-sorted layers, layer 0 == K=1
    result, every named face really belongs to the view, images finite and in range, gradients finite."""
    trb = _trb()
    v, f = _grid_sphere(101, 200)
    assert f.shape[0] > 39000
    eye = 2.7 * torch.nn.functional.normalize(torch.tensor([[0.3, 0.5, -0.8], [-0.6, -0.2, 0.7]]), dim=1)
    R, T = trb.look_at_view_transform(eye=eye)
    mesh = trb.Meshes([v.to(DEV)], [f.to(DEV)], textures=trb.TexturesVertex(torch.rand(1, v.shape[0], 3, device=DEV))).extend(2)
    cams = trb.FoVPerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV))
    rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=192, blur_radius=BLUR, faces_per_pixel=8))
    frag = rast(mesh)
    ndc = rast.transform(mesh).cpu().reshape(2, -1, 3)
    want = oracle_rasterize(ndc, f, (192, 192), BLUR, 8, True, True)
    _check_fragments(frag, want)
    assert (want[0][..., 7] >= 0).sum() > 1000  # the K-th layer is in use: the top-K cut really happens

    # full size
    v, f = _grid_sphere(501, 1000)
    assert f.shape[0] == 1_000_000
    vd = v.to(DEV).requires_grad_(True)
    mesh = trb.Meshes([vd], [f.to(DEV)], textures=trb.TexturesVertex(torch.rand(1, v.shape[0], 3, device=DEV))).extend(2)
    settings8 = trb.RasterizationSettings(image_size=1024, blur_radius=BLUR, faces_per_pixel=8)
    settings1 = trb.RasterizationSettings(image_size=1024, blur_radius=BLUR, faces_per_pixel=1)
    shader = trb.SoftPhongShader(device=DEV, cameras=cams, lights=trb.PointLights(device=DEV, location=[[0.0, 0.0, -3.0]]))
    images, fa = trb.MeshRendererWithFragments(trb.MeshRasterizer(cams, settings8), shader)(mesh)
    fb = trb.MeshRasterizer(cams, settings8)(mesh)
    f1 = trb.MeshRasterizer(cams, settings1)(mesh)
    assert torch.equal(fa.pix_to_face, fb.pix_to_face) and torch.equal(fa.zbuf, fb.zbuf)
    assert torch.equal(fa.pix_to_face[..., 0], f1.pix_to_face[..., 0]) and torch.equal(fa.zbuf[..., 0], f1.zbuf[..., 0])
    valid = fa.pix_to_face >= 0
    z = torch.where(valid, fa.zbuf, torch.full_like(fa.zbuf, float("inf")))
    assert (z[..., 1:] >= z[..., :-1]).all()                       # front to back
    assert (valid[..., 1:] <= valid[..., :-1]).all()               # -1 only at the tail
    F = f.shape[0]
    view_of = torch.arange(2, device=DEV).view(2, 1, 1, 1).expand_as(fa.pix_to_face)
    assert torch.equal((fa.pix_to_face // F)[valid], view_of[valid])
    assert 0.2 < valid[..., 0].float().mean().item() < 0.9
    assert torch.isfinite(images).all() and images.min() >= 0 and images[..., :3].max() <= 1.0 + 1e-4
    (images ** 2).mean().backward()
    assert torch.isfinite(vd.grad).all() and vd.grad.abs().sum() > 0


@pytest.mark.parametrize("nlat,nlon,size,K", [(101, 200, 64, 50), (246, 245, 96, 8)])
def test_dense_tile_lists_ordered_in_super_chunks(nlat, nlon, size, K):
    """Tile lists longer than one ordered super-chunk of the K > 1 fine kernel (2,048 entries for 8x8 tiles,
    8,192 for 16x16): 40k faces on a 64^2 image at K=50 and 120k faces on a 96^2 image at K=8, soft blur with
    clipped, perspective-correct barycentrics.  Every early-exit path (depth-bucket order, CTA stop, one-positive-
    weight shortcut, approximate-depth pre-reject) must leave the result bit-identical to the oracle's full walk."""
    trb = _trb()
    v, f = _grid_sphere(nlat, nlon, seed=4)
    eye = 2.7 * torch.nn.functional.normalize(torch.tensor([[0.5, -0.4, -0.7]]), dim=1)
    R, T = trb.look_at_view_transform(eye=eye)
    mesh = trb.Meshes([v.to(DEV)], [f.to(DEV)])
    cams = trb.FoVPerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV))
    rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=size, blur_radius=BLUR, faces_per_pixel=K))
    frag = rast(mesh)
    ndc = rast.transform(mesh).cpu().reshape(1, -1, 3)
    want = oracle_rasterize(ndc, f, (size, size), BLUR, K, True, True)
    _check_fragments(frag, want)
    assert (want[0][..., K - 1] >= 0).sum() > 500


def test_c2_headline_workload_exact_at_its_own_size():
    """The bench.py workload itself: the UN-normalised cow (+-0.1 units) from the dist-0.7 orbit
    of 64 cameras at 512^2, K=1 -- ~3.5 px per face, >90 % empty tiles, the strip-fill fast path.  The whole 64-view
    batch goes through the fused renderer as in the bench; views 5 and 37 of it are checked against the oracle:
    pix_to_face bit-exact, zbuf / bary / dists 1e-5, image 1e-4, gradients 1e-3 vs fp64 autograd."""
    trb = _trb()
    v, f = load_mesh("cow")
    torch.manual_seed(0)
    colors = torch.rand(v.shape[0], 3)
    N, H, W = 64, 512, 512
    R, T = trb.look_at_view_transform(dist=0.7, elev=torch.linspace(0, 360, N), azim=torch.linspace(-180, 180, N))
    pick = [5, 37]
    vd, cd = v.to(DEV).requires_grad_(True), colors.to(DEV).requires_grad_(True)
    Rd, Td = R.to(DEV).requires_grad_(True), T.to(DEV).requires_grad_(True)
    meshes = trb.Meshes(verts=[vd], faces=[f.to(DEV)], textures=trb.TexturesVertex(cd[None])).extend(N)
    cams = trb.FoVPerspectiveCameras(device=DEV)
    rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=H, blur_radius=0.0, faces_per_pixel=1))
    shader = trb.SoftPhongShader(device=DEV, cameras=cams, lights=trb.PointLights(device=DEV, location=[[0.0, 0.0, -3.0]]))
    images, frag = trb.MeshRendererWithFragments(rast, shader)(meshes, R=Rd, T=Td)
    # loss over the two checked views only, so that the fp64 model needs only them
    gsel = torch.zeros(N, 1, 1, 1, device=DEV)
    gsel[pick] = 1.0
    ((images ** 2) * gsel).mean().backward()
    ndc = rast.transform(meshes, R=Rd, T=Td).detach().cpu().reshape(N, -1, 3)[pick]
    want = oracle_rasterize(ndc, f, (H, W), 0.0, 1, True, False)
    F = f.shape[0]
    got_p2f = frag.pix_to_face[pick].cpu().numpy()
    base = np.array(pick, dtype=np.int64).reshape(2, 1, 1, 1) * F
    want_p2f = np.where(want[0] >= 0, want[0] - (np.arange(2).reshape(2, 1, 1, 1) * F) + base, -1)
    mism = int((got_p2f != want_p2f).sum())
    assert mism == 0, f"pix_to_face differs at {mism} samples"
    cover = (want[0] >= 0).mean()
    assert 0.005 < cover < 0.2, cover          # the headline regime: a few percent of the image is covered
    for name, a, b in (("zbuf", frag.zbuf, want[1]), ("bary", frag.bary_coords, want[2]), ("dists", frag.dists, want[3])):
        a = a[pick].detach().cpu().numpy()
        assert np.allclose(a, b, atol=1e-5, rtol=1e-5), f"{name} max abs diff {np.abs(a - b).max()}"
    # fp64 model of the two views
    n = len(pick)
    v64, c64 = v.double().requires_grad_(True), colors.double().requires_grad_(True)
    R64, T64 = R[pick].double().requires_grad_(True), T[pick].double().requires_grad_(True)
    proj = fov_proj(n).double()
    ndc64 = sref.world_to_ndc(v64, R64, T64, proj[:, 0], proj[:, 1], proj[:, 2], proj[:, 3], True)
    p2f = torch.from_numpy(want[0])
    z64, b64, d64 = sref.raster_recompute(ndc64[:, f].reshape(-1, 3, 3), p2f, True, False)
    ones = lambda *x: torch.tensor([list(x)], dtype=torch.float64).repeat(n, 1)
    cam = -torch.matmul(T64[:, None, :], torch.linalg.inv(R64))[:, 0, :]
    ref = sref.shade(p2f, b64, z64, d64, f.repeat(n, 1), v64, sref.vertex_normals(v64, f), c64, shader="soft_phong",
                     light_kind="point", light_vec=ones(0, 0, -3.0), light_ambient=ones(.5, .5, .5),
                     light_diffuse=ones(.3, .3, .3), light_specular=ones(.2, .2, .2), mat_ambient=ones(1, 1, 1),
                     mat_diffuse=ones(1, 1, 1), mat_specular=ones(1, 1, 1),
                     shininess=torch.full((n,), 64.0, dtype=torch.float64), camera_center=cam)
    err = (images[pick].detach().cpu().double() - ref).abs().max().item()
    assert err < 1e-4, f"image error {err}"
    ((ref ** 2).sum() / (N * H * W * 4)).backward()
    assert rel_l2(vd.grad.cpu(), v64.grad) < 1e-3
    assert rel_l2(cd.grad.cpu(), c64.grad) < 1e-3
    assert rel_l2(Rd.grad[pick].cpu(), R64.grad) < 1e-3
    assert rel_l2(Td.grad[pick].cpu(), T64.grad) < 1e-3
    others = [i for i in range(N) if i not in pick]
    assert float(Rd.grad[others].abs().max()) == 0.0 and float(Td.grad[others].abs().max()) == 0.0
