"""GPU parity of the point-cloud rendering path (last item) through the C ABI: rasteriser idx
bit-exact against the C oracle, zbuf / dists / images within fp32 tolerance, every gradient against fp64 autograd."""
import numpy as np
import pytest
import torch

from oracle import points_render_ref as pr
from helpers import fov_proj, rel_l2
from oracle import shading_ref as sref

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
TOL = dict(atol=1e-5, rtol=1e-5)


def _trb():
    import torch_renderer_b200 as trb
    return trb


def _clouds(seed, sizes, spread=1.1):
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in sizes:
        xy = (torch.rand(n, 2, generator=g) * 2 - 1) * spread
        z = torch.rand(n, 1, generator=g) * 3 - 0.3          # a few behind the camera
        out.append(torch.cat([xy, z], dim=1))
    # exact depth ties: duplicate a few depths
    for c in out:
        if c.shape[0] > 8:
            c[5:8, 2] = c[4, 2]
    return out


@pytest.mark.parametrize("seed,sizes,image_size,K,radius", [
    (0, (400, 250), (64, 64), 1, 0.05),
    (1, (600, 0, 300), (45, 61), 5, 0.12),         # an empty cloud in the batch; partial tiles
    (2, (300,), (32, 48), 20, 0.3),
    (3, (200,), (16, 16), 150, 1.5),               # every point reaches every pixel; K at the limit
    (4, (500, 400), (40, 40), 4, "per-point"),
])
def test_rasterize_points_bit_exact(seed, sizes, image_size, K, radius):
    trb = _trb()
    clouds = _clouds(seed, sizes)
    P = sum(sizes)
    if radius == "per-point":
        g = torch.Generator().manual_seed(seed)
        r_packed = torch.rand(P, generator=g) * 0.15 + 0.01
        r_arg = torch.zeros(len(sizes), max(sizes))
        o = 0
        for i, n in enumerate(sizes):
            r_arg[i, :n] = r_packed[o:o + n]
            o += n
    else:
        r_packed, r_arg = torch.full((P,), radius), radius
    first = np.cumsum([0] + list(sizes[:-1]))
    want = pr.rasterize_points(torch.cat(clouds).numpy(), first, np.array(sizes), r_packed.numpy(), image_size, K)
    pc = trb.Pointclouds([c.to(DEV) for c in clouds])
    idx, zbuf, dists = trb.rasterize_points(pc, image_size, r_arg if not torch.is_tensor(r_arg) else r_arg.to(DEV), K)
    assert idx.dtype == torch.int32 and idx.shape == (len(sizes), image_size[0], image_size[1], K)
    assert (want[0] >= 0).sum() > 100
    mism = int((idx.cpu().numpy() != want[0]).sum())
    assert mism == 0, f"idx differs at {mism} samples"
    assert np.allclose(zbuf.cpu().numpy(), want[1], **TOL) and np.allclose(dists.cpu().numpy(), want[2], **TOL)


def test_rasterize_points_backward():
    trb = _trb()
    torch.manual_seed(3)
    clouds = _clouds(7, (300, 200))
    H, W, K = 40, 36, 4
    pts = [c.to(DEV).requires_grad_(True) for c in clouds]
    idx, zbuf, dists = trb.rasterize_points(trb.Pointclouds(pts), (H, W), 0.15, K)
    gz, gd = torch.randn(2, H, W, K), torch.randn(2, H, W, K)
    m = (idx >= 0).cpu()
    ((zbuf * (gz * m).to(DEV)).sum() + (dists * (gd * m).to(DEV)).sum()).backward()
    got = torch.cat([p.grad for p in pts]).cpu()
    want = pr.rasterize_points_backward(torch.cat(clouds).numpy(), idx.cpu().numpy(), (gz * m).numpy(), (gd * m).numpy())
    assert rel_l2(got, torch.from_numpy(want)) < 1e-5


@pytest.mark.parametrize("mode", ["alpha", "norm"])
@pytest.mark.parametrize("C,background", [(3, None), (4, (0.2, 0.4, 0.6)), (3, (0.1, 0.9, 0.5))])
def test_compositors_forward_and_gradients(mode, C, background):
    trb = _trb()
    torch.manual_seed(11)
    N, H, W, K, P = 2, 24, 20, 6, 500
    idx = torch.randint(0, P, (N, H, W, K))
    fill = torch.randint(0, K + 1, (N, H, W, 1))                      # slots beyond `fill` are empty; some pixels have none
    idx = torch.where(torch.arange(K).view(1, 1, 1, K) < fill, idx, torch.full_like(idx, -1))
    alphas = torch.rand(N, H, W, K)
    alphas[0, 0, 0] = 1e-6                                             # the clamp of the normaliser
    feats = torch.rand(P, C)
    a_d = alphas.to(DEV).requires_grad_(True)
    f_d = feats.to(DEV).requires_grad_(True)
    comp = (trb.AlphaCompositor if mode == "alpha" else trb.NormWeightedCompositor)(background_color=background)
    # upstream layouts: (N, K, H, W), (C, P) -> (N, C, H, W)
    img = comp(idx.to(DEV).permute(0, 3, 1, 2), a_d.permute(0, 3, 1, 2), f_d.permute(1, 0))
    assert img.shape == (N, C, H, W)
    a64, f64 = alphas.double().requires_grad_(True), feats.double().requires_grad_(True)
    ref = (pr.alpha_composite if mode == "alpha" else pr.norm_weighted_sum)(idx, a64, f64)
    ref = pr.add_background(ref, idx, background)
    assert torch.allclose(img.permute(0, 2, 3, 1).cpu().double(), ref, atol=2e-6)
    w = torch.randn(N, H, W, C)
    (img.permute(0, 2, 3, 1) * w.to(DEV)).sum().backward()
    (ref * w.double()).sum().backward()
    assert rel_l2(a_d.grad.cpu(), a64.grad) < 1e-5
    assert rel_l2(f_d.grad.cpu(), f64.grad) < 1e-5
    # the functional forms (no background)
    fn = trb.renderer.alpha_composite if mode == "alpha" else trb.renderer.norm_weighted_sum
    plain = fn(idx.to(DEV).permute(0, 3, 1, 2), alphas.to(DEV).permute(0, 3, 1, 2), feats.to(DEV).t())
    ref0 = (pr.alpha_composite if mode == "alpha" else pr.norm_weighted_sum)(idx, alphas.double(), feats.double())
    assert torch.allclose(plain.permute(0, 2, 3, 1).cpu().double(), ref0, atol=2e-6)


@pytest.mark.parametrize("mode", ["alpha", "norm"])
def test_points_renderer_end_to_end(mode):
    """The reference's AlphaPointRender / NormPointRender call (torch_renderer.py:163-208): PointsRenderer with a FoV
    camera and per-call R, T; images and the gradients w.r.t. points, features and T against the oracle route."""
    trb = _trb()
    torch.manual_seed(5)
    g = torch.Generator().manual_seed(9)
    n_pts = (700, 500)
    clouds = [torch.randn(n, 3, generator=g) * 0.45 for n in n_pts]
    feats = [torch.rand(n, 3, generator=g) for n in n_pts]
    R, T = trb.look_at_view_transform(dist=2.5, elev=torch.tensor([10.0, -25.0]), azim=torch.tensor([30.0, 200.0]))
    H, W, K, radius, bg = 48, 56, 8, 0.06, (0.0, 0.3, 0.1)
    pts_d = [c.to(DEV).requires_grad_(True) for c in clouds]
    f_d = [f.to(DEV).requires_grad_(True) for f in feats]
    T_d = T.to(DEV).requires_grad_(True)
    pc = trb.Pointclouds(pts_d, features=f_d)
    cams = trb.FoVPerspectiveCameras(device=DEV)
    rast = trb.PointsRasterizer(cams, trb.PointsRasterizationSettings(image_size=(H, W), radius=radius, points_per_pixel=K))
    comp = (trb.AlphaCompositor if mode == "alpha" else trb.NormWeightedCompositor)(background_color=bg)
    images = trb.PointsRenderer(rasterizer=rast, compositor=comp)(pc, R=R.to(DEV), T=T_d)
    assert images.shape == (2, H, W, 3)
    w = torch.rand(2, H, W, 3)
    (images * w.to(DEV)).sum().backward()
    frag = rast(pc, R=R.to(DEV), T=T_d.detach())
    ndc = rast.transform(pc, R=R.to(DEV), T=T_d.detach()).detach().cpu()
    first = np.array([0, n_pts[0]])
    want = pr.rasterize_points(ndc.numpy(), first, np.array(n_pts), radius, (H, W), K)
    assert (want[0][..., -1] >= 0).sum() > 50 and (want[0][..., 0] < 0).sum() > 50      # full and empty pixels both occur
    assert np.array_equal(frag.idx.cpu().numpy(), want[0])
    assert np.allclose(frag.zbuf.detach().cpu().numpy(), want[1], **TOL) and np.allclose(frag.dists.detach().cpu().numpy(), want[2], **TOL)
    # fp64 route with idx fixed: transform -> dist^2 to the pixel centre -> weights -> compositor -> background
    p64 = [c.double().requires_grad_(True) for c in clouds]
    f64 = [f.double().requires_grad_(True) for f in feats]
    T64 = T.double().requires_grad_(True)
    proj = fov_proj(2).double()
    ndc64 = torch.cat([sref.world_to_ndc(p64[i], R[i:i + 1].double(), T64[i:i + 1], proj[i:i + 1, 0], proj[i:i + 1, 1],
                                         proj[i:i + 1, 2], proj[i:i + 1, 3], True)[0] for i in range(2)])
    assert torch.allclose(ndc64.float(), ndc, atol=1e-5, rtol=1e-5)
    idx = torch.from_numpy(want[0]).long()
    ys, xs = sref.pixel_centers(H, W, torch.float64)
    sel = ndc64[idx.clamp(min=0)]                                                       # (N, H, W, K, 3)
    d2 = (xs.view(1, 1, W, 1) - sel[..., 0]) ** 2 + (ys.view(1, H, 1, 1) - sel[..., 1]) ** 2
    ref = pr.render_points(idx, d2, torch.cat(f64), radius, mode, bg)
    assert torch.allclose(images.detach().cpu().double(), ref, atol=5e-5)
    (ref * w.double()).sum().backward()
    assert rel_l2(torch.cat([p.grad for p in pts_d]).cpu(), torch.cat([p.grad for p in p64])) < 1e-3
    assert rel_l2(torch.cat([f.grad for f in f_d]).cpu(), torch.cat([f.grad for f in f64])) < 1e-4
    assert rel_l2(T_d.grad.cpu(), T64.grad) < 1e-3


def test_binned_point_rasteriser_reference_settings_and_list_overflow():
    """The per-tile point lists (count -> allocate -> fill): (a) the reference's AlphaPointRender settings
    (torch_renderer.py:163-208: radius 0.003, 10 points per pixel) on a 30 k-point cloud at 256^2 against the C
    oracle; (b) lists that do NOT fit their capacity (radius hint far too small) fall back to the whole-cloud scan
    tile by tile -- same result, bit for bit; (c) the un-binned entry point gives the same tensors."""
    import ctypes
    trb = _trb()
    from torch_renderer_b200 import _lib, ops
    g = torch.Generator().manual_seed(11)
    n = 30000
    u = torch.randn(n, 3, generator=g)
    u = u / u.norm(dim=1, keepdim=True) * 0.8
    cloud = torch.stack([u[:, 0], u[:, 1], u[:, 2] + 2.0], dim=1)          # a sphere shell in NDC x, y + depth
    want = pr.rasterize_points(cloud.numpy(), np.array([0]), np.array([n]), np.full((n,), 0.003, np.float32), (256, 256), 10)
    pc = trb.Pointclouds([cloud.to(DEV)])
    idx, zbuf, dists = trb.rasterize_points(pc, 256, 0.003, 10)
    assert (want[0] >= 0).sum() > 5000
    assert int((idx.cpu().numpy() != want[0]).sum()) == 0
    assert np.allclose(zbuf.cpu().numpy(), want[1], **TOL) and np.allclose(dists.cpu().numpy(), want[2], **TOL)
    # (b) big discs, capacity sized for tiny ones: most tiles overflow
    clouds = _clouds(7, (3000,))
    pc = trb.Pointclouds([c.to(DEV) for c in clouds])
    radius = torch.full((3000,), 0.2, device=DEV)
    table = pc.view_table()
    a = ops.rasterize_points_ndc(pc.points_packed(), radius, table, (96, 96), 6, max_radius=0.2)
    b = ops.rasterize_points_ndc(pc.points_packed(), radius, table, (96, 96), 6, max_radius=0.0)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    # (c) the streaming kernel behind trb_points_raster_forward
    c = [torch.empty_like(t) for t in a]
    _lib.check(_lib.lib().trb_points_raster_forward(
        pc.points_packed().data_ptr(), radius.data_ptr(), table.views.data_ptr(), 1, 96, 96, 6, c[0].data_ptr(),
        c[1].data_ptr(), c[2].data_ptr(), 0, torch.cuda.current_stream().cuda_stream), "points")
    torch.cuda.synchronize()
    for x, y in zip(a, c):
        assert torch.equal(x, y)
