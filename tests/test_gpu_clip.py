"""GPU parity of the near-plane clipping route: clip_faces (torch, on the device) ->
trb_raster_forward on the cut faces -> trb_clip_resequence (upstream's one-of-two-neighbours rule) -> conversion back,
against oracle/clip_ref.py + the C oracle; backward against fp64 autograd through the same route."""
import math

import numpy as np
import pytest
import torch

import oracle
from oracle import clip_ref
from oracle import shading_ref as sref
from helpers import fov_proj, oracle_rasterize_clipped, rel_l2, uv_sphere

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0")
TOL = dict(atol=1e-5, rtol=1e-5)


def _trb():
    import torch_renderer_b200 as trb
    return trb


def _assert_close_fragments(got, want):
    p2f, zbuf, bary, dists = [t.detach().cpu().numpy() for t in got]
    mism = int((p2f != want[0]).sum())
    assert mism == 0, f"pix_to_face differs at {mism} of {p2f.size} samples"
    for name, a, b in (("zbuf", zbuf, want[1]), ("bary", bary, want[2]), ("dists", dists, want[3])):
        assert np.allclose(a, b, **TOL), f"{name} max abs diff {np.abs(a - b).max()}"


def _soup(seed, n_faces, big=0.9):
    """Random triangles in NDC whose depths straddle z = 0.5 (many cut into quadrilaterals), a few large ones."""
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(n_faces, 1, 2, generator=g) * 2.4 - 1.2
    size = torch.where(torch.rand(n_faces, 1, 1, generator=g) < 0.15, big, 0.25)
    xy = c + (torch.rand(n_faces, 3, 2, generator=g) - 0.5) * size * 2
    z = torch.rand(n_faces, 3, 1, generator=g) * 1.6 + 0.05
    verts = torch.cat([xy, z], dim=2).reshape(-1, 3)
    faces = torch.arange(3 * n_faces).reshape(-1, 3)
    return verts, faces


def _ground_scene():
    h = 0.4
    quad = torch.tensor([[-3.0, -h, -1.0], [3.0, -h, -1.0], [3.0, -h, 6.0], [-3.0, -h, 6.0]])
    faces = torch.tensor([[0, 2, 1], [0, 3, 2]])
    t = math.tan(math.radians(30.0))
    ndc = quad.clone()
    ndc[:, 0] = quad[:, 0] / (quad[:, 2] * t)
    ndc[:, 1] = quad[:, 1] / (quad[:, 2] * t)
    return ndc, faces


@pytest.mark.parametrize("image_size,K,blur,persp,clip,cull", [
    ((48, 48), 1, 0.0, True, False, False),
    ((48, 48), 3, 2e-4, True, True, False),
    ((33, 61), 8, 3e-3, True, True, False),
])
def test_ground_plane_through_the_near_plane(image_size, K, blur, persp, clip, cull):
    trb = _trb()
    ndc, faces = _ground_scene()
    want, cf, _ = oracle_rasterize_clipped(ndc[None], faces, image_size, blur, K, persp, clip, cull, z_clip_value=0.5)
    meshes = trb.Meshes(verts=[ndc.to(DEV)], faces=[faces.to(DEV)])
    got = trb.renderer.rasterize_meshes(meshes, image_size, blur, K, perspective_correct=persp,
                                        clip_barycentric_coords=clip, cull_backfaces=cull, z_clip_value=0.5)
    assert (want[0] >= 0).sum() > 100
    _assert_close_fragments(got, want)


@pytest.mark.parametrize("seed,n_faces,image_size,K,blur,persp,clip,cull,frustum", [
    (0, 60, (40, 40), 1, 0.0, True, False, False, False),
    (1, 60, (40, 40), 1, 2e-3, True, True, False, False),      # K = 1 with blur: upstream's order-dependent replacements
    (2, 80, (37, 53), 2, 4e-3, False, True, False, False),
    (3, 80, (64, 64), 4, 1e-3, True, True, True, False),
    (4, 50, (32, 32), 50, 8e-3, True, True, False, False),     # 8x8 tiles
    (5, 120, (48, 48), 3, 2e-3, True, False, False, True),     # cull_to_frustum as well
    (6, 40, (24, 24), 150, 1e-2, False, True, False, False),
])
def test_cut_triangle_soups_bit_exact(seed, n_faces, image_size, K, blur, persp, clip, cull, frustum):
    """Two views with different soups: pix_to_face equals the oracle's sample for sample -- including the pixels
    where upstream's neighbour rule makes the answer depend on the order faces arrive in."""
    trb = _trb()
    scenes = [_soup(seed * 10 + i, n_faces + 7 * i) for i in range(2)]
    fv, first, count = [], [], []
    for v, f in scenes:
        first.append(sum(count)); count.append(f.shape[0]); fv.append(v[f])
    fv = torch.cat(fv).numpy()
    cf = clip_ref.clip_faces(fv, np.array(first), np.array(count), clip_ref.rasterizer_frustum(persp, 0.5, frustum))
    assert (cf.clipped_faces_neighbor_idx >= 0).sum() >= 20
    raw = oracle.rasterize_forward(cf.face_verts, cf.mesh_to_face_first_idx, cf.num_faces_per_mesh, image_size, blur, K,
                                   persp, clip, cull, 0, clipped_faces_neighbor_idx=cf.clipped_faces_neighbor_idx)
    plain = oracle.rasterize_forward(cf.face_verts, cf.mesh_to_face_first_idx, cf.num_faces_per_mesh, image_size, blur, K,
                                     persp, clip, cull, 0)
    if blur > 0:
        assert (raw[0] != plain[0]).any()      # the rule matters in this scene
    p2f_u, bary_u = clip_ref.convert_clipped_rasterization_to_original_faces(raw[0], raw[2], cf)
    meshes = trb.Meshes(verts=[v.to(DEV) for v, _ in scenes], faces=[f.to(DEV) for _, f in scenes])
    got = trb.renderer.rasterize_meshes(meshes, image_size, blur, K, perspective_correct=persp,
                                        clip_barycentric_coords=clip, cull_backfaces=cull, z_clip_value=0.5,
                                        cull_to_frustum=frustum)
    _assert_close_fragments(got, (p2f_u, raw[1], bary_u, raw[3]))


def test_nothing_behind_the_plane_is_the_ordinary_route():
    trb = _trb()
    v, f = _soup(3, 40)
    v[:, 2] += 0.6
    meshes = trb.Meshes(verts=[v.to(DEV)], faces=[f.to(DEV)])
    a = trb.renderer.rasterize_meshes(meshes, 32, 1e-3, 3, perspective_correct=True, clip_barycentric_coords=True, z_clip_value=0.5)
    b = trb.renderer.rasterize_meshes(meshes, 32, 1e-3, 3, perspective_correct=True, clip_barycentric_coords=True)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("K,blur", [(1, 0.0), (4, 1.5e-3)])
def test_mesh_rasterizer_with_fov_camera_inside_the_clip_distance(K, blur):
    """FoVPerspectiveCameras define znear = 1: upstream clips at 0.5.  A camera 1.3 from the centre of a unit sphere
    has the near cap closer than that: the cap's faces are removed, the ring crossing the plane is cut."""
    trb = _trb()
    v, f = uv_sphere(16, 20, 1.0, noise=0.02, seed=6)
    R, T = trb.look_at_view_transform(dist=1.3, elev=torch.tensor([20.0, -35.0]), azim=torch.tensor([40.0, 170.0]))
    mesh = trb.Meshes(verts=[v.to(DEV)], faces=[f.to(DEV)]).extend(2)
    cams = trb.FoVPerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV), fov=120.0)   # wide: the cut ring is in view
    rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=(80, 96), blur_radius=blur, faces_per_pixel=K))
    frag = rast(mesh)
    ndc = rast.transform(mesh).cpu().reshape(2, -1, 3)
    assert (ndc[..., 2] < 0.5).any()
    want, cf, raw = oracle_rasterize_clipped(ndc, f, (80, 96), blur, K, True, blur > 0, False, z_clip_value=0.5)
    assert cf.barycentric_conversion is not None and cf.barycentric_conversion.shape[0] > 20
    on_cut_faces = np.where(raw >= 0, cf.faces_clipped_to_conversion_idx[np.clip(raw, 0, None)], -1) >= 0
    assert on_cut_faces.sum() > 200
    _assert_close_fragments((frag.pix_to_face, frag.zbuf, frag.bary_coords, frag.dists), want)
    assert (frag.zbuf[frag.pix_to_face >= 0] >= 0.5 - 1e-5).all()
    # "off": the pre-clipping behaviour -- whole faces, only those entirely behind the plane removed
    trb.set_near_plane_clipping("off")
    try:
        frag_off = rast(mesh)
    finally:
        trb.set_near_plane_clipping("exact")
    assert (frag_off.pix_to_face != frag.pix_to_face).any()
    # the renderer takes the same route and shades the clipped Fragments
    cols = torch.rand(1, v.shape[0], 3, device=DEV)
    mesh_c = trb.Meshes([v.to(DEV)], [f.to(DEV)], textures=trb.TexturesVertex(cols)).extend(2)
    shader = trb.SoftPhongShader(device=DEV, cameras=cams, lights=trb.PointLights(device=DEV, location=[[0.0, 1.0, -2.0]]))
    trb.set_fragment_cache(False)
    try:
        images, frag_r = trb.MeshRendererWithFragments(rast, shader)(mesh_c)
    finally:
        trb.set_fragment_cache(True)
    assert torch.equal(frag_r.pix_to_face, frag.pix_to_face)
    assert torch.allclose(images, shader(frag, mesh_c), atol=1e-5)


def test_clipped_route_backward_matches_fp64_autograd():
    """d loss / d NDC vertices through convert <- rasterise <- clip_faces, cut weights constant as upstream."""
    trb = _trb()
    from torch_renderer_b200 import clip
    torch.manual_seed(4)
    v, f = _soup(11, 70)
    H, W, K, blur, persp, clipb = 40, 44, 3, 2e-3, True, True
    want_frag, cf, raw_p2f = oracle_rasterize_clipped(v[None], f, (H, W), blur, K, persp, clipb, False, z_clip_value=0.5)
    verts_dev = v.to(DEV).requires_grad_(True)
    meshes = trb.Meshes(verts=[verts_dev], faces=[f.to(DEV)])
    p2f, zbuf, bary, dists = trb.renderer.rasterize_meshes(meshes, (H, W), blur, K, perspective_correct=persp,
                                                           clip_barycentric_coords=clipb, z_clip_value=0.5)
    _assert_close_fragments((p2f, zbuf, bary, dists), want_frag)
    gz, gb, gd = torch.randn(1, H, W, K), torch.randn(1, H, W, K, 3), torch.randn(1, H, W, K)
    m = (p2f >= 0).cpu()
    loss = (zbuf * (gz * m).to(DEV)).sum() + (bary * (gb * m[..., None]).to(DEV)).sum() + (dists * (gd * m).to(DEV)).sum()
    loss.backward()
    got = verts_dev.grad.cpu()
    # fp64: the same route in torch on the CPU with pix_to_face (clipped indexing) fixed
    v64 = v.double().requires_grad_(True)
    fv64 = v64[f]
    cf64 = clip.clip_faces(fv64, torch.tensor([0]), torch.tensor([f.shape[0]]), clip.rasterizer_frustum(persp, 0.5, False))
    assert np.array_equal(cf64.clipped_faces_neighbor_idx.numpy(), cf.clipped_faces_neighbor_idx)
    raw = torch.from_numpy(raw_p2f)
    z64, b64, d64 = sref.raster_recompute(cf64.face_verts, raw, persp, clipb)
    _, b64u = clip.convert_clipped_rasterization_to_original_faces(raw, b64, cf64)
    ((z64 * gz * m).sum() + (b64u * gb * m[..., None]).sum() + (d64 * gd * m).sum()).backward()
    assert v64.grad.abs().sum() > 0
    assert rel_l2(got, v64.grad) < 1e-3


def _plane(nx=6, nz=10):
    """A strip under the camera from z = -1 (behind it) to z = 4, in view coordinates."""
    xs, zs = torch.linspace(-0.6, 0.6, nx + 1), torch.linspace(-1.0, 4.0, nz + 1)
    v = torch.stack([xs[None, :].expand(nz + 1, -1), torch.full((nz + 1, nx + 1), -0.4),
                     zs[:, None].expand(-1, nx + 1)], -1).reshape(-1, 3)
    f = []
    for j in range(nz):
        for i in range(nx):
            a = j * (nx + 1) + i
            f += [[a, a + nx + 1, a + 1], [a + 1, a + nx + 1, a + nx + 2]]
    return v.contiguous(), torch.tensor(f)


def test_camera_translation_gradient_through_the_whole_clipped_route():
    """MeshRenderer with a FoV camera standing on a strip that runs through the near plane: d loss / d T through
    shade <- convert <- rasterise <- clip_faces <- transform equals fp64 autograd of the same route."""
    trb = _trb()
    from torch_renderer_b200 import clip
    torch.manual_seed(2)
    v, f = _plane()
    H, W, K, blur, sigma = 48, 64, 8, 4e-3, 1e-3
    mesh = trb.Meshes(verts=[v.to(DEV)], faces=[f.to(DEV)])
    cams = trb.FoVPerspectiveCameras(device=DEV)
    rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=(H, W), blur_radius=blur, faces_per_pixel=K))
    rend = trb.MeshRendererWithFragments(rast, trb.SoftSilhouetteShader(trb.BlendParams(sigma=sigma, gamma=1e-3)))
    R = torch.eye(3)[None].to(DEV)
    T = torch.tensor([[0.05, 0.0, 0.1]], device=DEV, requires_grad=True)
    weights = torch.rand(1, H, W)
    trb.set_fragment_cache(False)
    try:
        images, frag = rend(mesh, R=R, T=T)
        (images[..., 3] * weights.to(DEV)).sum().backward()
        ndc = rast.transform(mesh, R=R, T=T.detach()).cpu().reshape(1, -1, 3)
    finally:
        trb.set_fragment_cache(True)
    assert (ndc[..., 2] < 0.5).any() and (frag.pix_to_face >= 0).sum() > 500
    want_frag, cf, raw_p2f = oracle_rasterize_clipped(ndc, f, (H, W), blur, K, True, True, False, z_clip_value=0.5)
    _assert_close_fragments((frag.pix_to_face, frag.zbuf, frag.bary_coords, frag.dists), want_frag)
    # fp64: the same route with pix_to_face (clipped indexing) fixed
    T64 = T.detach().cpu().double().requires_grad_(True)
    vv = v.double() + T64
    t = math.tan(math.radians(30.0))
    ndc64 = torch.stack([vv[:, 0] / (vv[:, 2] * t), vv[:, 1] / (vv[:, 2] * t), vv[:, 2]], -1)
    assert torch.allclose(ndc64.float(), ndc[0], atol=1e-4, rtol=1e-4)
    cf64 = clip.clip_faces(ndc64[f], torch.tensor([0]), torch.tensor([f.shape[0]]), clip.rasterizer_frustum(True, 0.5, False))
    raw = torch.from_numpy(raw_p2f)
    _, _, d64 = sref.raster_recompute(cf64.face_verts, raw, True, True)
    alpha = 1 - torch.prod(1 - torch.sigmoid(-d64 / sigma) * (raw >= 0), dim=-1)
    assert torch.allclose(alpha.float(), images[..., 3].detach().cpu(), atol=2e-4)
    (alpha * weights.double()).sum().backward()
    assert T64.grad.abs().max() > 1e-2
    assert rel_l2(T.grad.cpu(), T64.grad) < 5e-3


def test_everything_behind_the_plane_renders_background():
    """Camera in front of the whole mesh by less than the clip distance: clip_faces removes every face."""
    trb = _trb()
    v, f = uv_sphere(6, 8, 0.1)
    mesh = trb.Meshes(verts=[v.to(DEV)], faces=[f.to(DEV)], textures=trb.TexturesVertex(torch.ones(1, v.shape[0], 3, device=DEV)))
    cams = trb.FoVPerspectiveCameras(device=DEV, T=torch.tensor([[0.0, 0.0, 0.3]], device=DEV))   # depths 0.2 .. 0.4 < 0.5
    rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=32, blur_radius=1e-3, faces_per_pixel=3))
    trb.set_fragment_cache(False)
    try:
        frag = rast(mesh)
        img = trb.MeshRenderer(rast, trb.SoftPhongShader(device=DEV, cameras=cams, blend_params=trb.BlendParams(
            background_color=(0.25, 0.5, 0.75))))(mesh)
    finally:
        trb.set_fragment_cache(True)
    assert (frag.pix_to_face == -1).all() and (frag.zbuf == -1).all() and (frag.bary_coords == -1).all()
    # ... and an optimisation loop that wanders there gets zero gradients, not an exception
    Tg = torch.tensor([[0.0, 0.0, 0.3]], device=DEV, requires_grad=True)
    trb.set_fragment_cache(False)
    try:
        rast(mesh, T=Tg).zbuf.sum().backward()
    finally:
        trb.set_fragment_cache(True)
    assert Tg.grad is not None and float(Tg.grad.abs().max()) == 0.0
    assert torch.allclose(img[..., :3], torch.tensor([0.25, 0.5, 0.75], device=DEV).expand(1, 32, 32, 3), atol=1e-6)
    assert float(img[..., 3].abs().max()) < 1e-6


def test_mesh_rasterizer_cull_to_frustum_matches_oracle():
    """cull_to_frustum through RasterizationSettings: faces wholly outside [-1, 1]^2 disappear -- with blur that
    changes the band they would have cast into the frame -- on a heterogeneous (packed) batch."""
    trb = _trb()
    v0, f0 = uv_sphere(10, 12, 1.0, noise=0.03, seed=1)
    v1, f1 = uv_sphere(8, 10, 0.8, noise=0.0, seed=2)
    R, T = trb.look_at_view_transform(dist=torch.tensor([1.45, 1.2]), elev=torch.tensor([10.0, 40.0]), azim=torch.tensor([30.0, -100.0]))
    mesh = trb.Meshes(verts=[v0.to(DEV), v1.to(DEV)], faces=[f0.to(DEV), f1.to(DEV)])
    cams = trb.FoVPerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV), fov=50.0)
    H, W, K, blur = 48, 56, 4, 3e-3
    frags = {}
    for cull in (False, True):
        rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=(H, W), blur_radius=blur, faces_per_pixel=K,
                                                                 cull_to_frustum=cull))
        frags[cull] = rast(mesh)
        ndc = rast.transform(mesh).cpu()
    fv = torch.cat([ndc[:v0.shape[0]][f0], ndc[v0.shape[0]:][f1]]).numpy()
    first, count = np.array([0, f0.shape[0]]), np.array([f0.shape[0], f1.shape[0]])
    for cull in (False, True):
        cf = clip_ref.clip_faces(fv, first, count, clip_ref.rasterizer_frustum(True, 0.5, cull))
        raw = oracle.rasterize_forward(cf.face_verts, cf.mesh_to_face_first_idx, cf.num_faces_per_mesh, (H, W), blur, K,
                                       True, True, False, 0, clipped_faces_neighbor_idx=cf.clipped_faces_neighbor_idx)
        p2f, bary = clip_ref.convert_clipped_rasterization_to_original_faces(raw[0], raw[2], cf)
        fr = frags[cull]
        _assert_close_fragments((fr.pix_to_face, fr.zbuf, fr.bary_coords, fr.dists), (p2f, raw[1], bary, raw[3]))
    assert (frags[True].pix_to_face != frags[False].pix_to_face).any()
