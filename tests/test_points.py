"""Point-cloud rendering (last item): the oracle on hand-checkable scenes, the ``Pointclouds``
container, the call surface and the argument checks of the new entry points.  CPU only; parity on the GPU is in
test_gpu_points.py."""
import ctypes

import numpy as np
import pytest
import torch

import torch_renderer_b200 as trb
from oracle import points_render_ref as pr
from torch_renderer_b200 import _lib


def test_oracle_point_rasteriser_hand_scene():
    # 2 x 2 image: pixel centres at +-0.5; image row 0 / col 0 is NDC (+0.5, +0.5)
    pts = np.array([[0.5, 0.5, 2.0], [0.45, 0.5, 1.0], [-0.5, -0.5, 3.0], [0.5, -0.5, -0.1], [0.5, 0.5, 1.0]], np.float32)
    idx, z, d = pr.rasterize_points(pts, [0], [5], 0.1, (2, 2), 3)
    assert idx[0, 0, 0].tolist() == [1, 4, 0]            # depth 1 (index 1 before 4 on the tie), then depth 2
    assert np.allclose(z[0, 0, 0], [1.0, 1.0, 2.0]) and np.allclose(d[0, 0, 0], [0.0025, 0.0, 0.0], atol=1e-7)
    assert idx[0, 1, 1].tolist() == [2, -1, -1] and z[0, 1, 1, 1] == -1 and d[0, 1, 1, 2] == -1
    assert (idx[0, 1, 0] == -1).all()                    # the point there is behind the camera (z < 0)
    assert (idx[0, 0, 1] == -1).all()
    # K = 1 keeps the nearest; the radius test is strict
    assert pr.rasterize_points(pts, [0], [5], 0.1, (2, 2), 1)[0][0, 0, 0, 0] == 1
    assert (pr.rasterize_points(pts, [0], [5], 0.05, (2, 2), 3)[0][0, 0, 0] == [4, 0, -1]).all()   # 0.05^2 !< 0.05^2
    # two clouds: indices are into the packed points
    idx2 = pr.rasterize_points(pts, [0, 2], [2, 3], 0.1, (2, 2), 2)[0]
    assert idx2[0, 0, 0].tolist() == [1, 0] and idx2[1, 0, 0].tolist() == [4, -1] and idx2[1, 1, 1].tolist() == [2, -1]
    with pytest.raises(ValueError):
        pr.rasterize_points(pts, [0], [5], 0.1, (2, 2), 151)


def test_oracle_compositors_hand_values():
    idx = torch.tensor([[[[1, 0, -1], [-1, -1, -1]]]])                       # (1, 1, 2, 3)
    alphas = torch.tensor([[[[0.75, 1.0, 9.0], [5.0, 5.0, 5.0]]]])
    feats = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    a = pr.alpha_composite(idx, alphas, feats)
    assert torch.allclose(a[0, 0, 0], torch.tensor([0.25, 0.75, 0.0])) and (a[0, 0, 1] == 0).all()
    n = pr.norm_weighted_sum(idx, alphas, feats)
    assert torch.allclose(n[0, 0, 0], torch.tensor([1.0 / 1.75, 0.75 / 1.75, 0.0]))
    bg = pr.add_background(a, idx, (0.1, 0.2, 0.3))
    assert torch.allclose(bg[0, 0, 1], torch.tensor([0.1, 0.2, 0.3])) and torch.equal(bg[0, 0, 0], a[0, 0, 0])
    rgba = pr.add_background(torch.zeros(1, 1, 2, 4), idx, (0.1, 0.2, 0.3))
    assert rgba[0, 0, 1].tolist() == pytest.approx([0.1, 0.2, 0.3, 1.0])
    # tiny weights: the normaliser is clamped at 1e-4
    tiny = pr.norm_weighted_sum(idx[..., :1], torch.full((1, 1, 2, 1), 1e-6), feats)
    assert tiny[0, 0, 0, 1] == pytest.approx(1e-2)


def test_pointclouds_container():
    a, b = torch.rand(5, 3), torch.rand(3, 3)
    fa, fb = torch.rand(5, 4), torch.rand(3, 4)
    pc = trb.Pointclouds([a, b], features=[fa, fb])
    assert len(pc) == 2 and not pc.isempty()
    assert pc.points_packed().shape == (8, 3) and pc.features_packed().shape == (8, 4) and pc.normals_packed() is None
    assert pc.num_points_per_cloud().tolist() == [5, 3] and pc.cloud_to_packed_first_idx().tolist() == [0, 5]
    assert pc.packed_to_cloud_idx().tolist() == [0] * 5 + [1] * 3
    padded = pc.points_padded()
    assert padded.shape == (2, 5, 3) and torch.equal(padded[1, :3], b) and (padded[1, 3:] == 0).all()
    assert torch.equal(trb.Pointclouds(padded[:1]).points_list()[0], a)
    ext = pc.extend(2)
    assert len(ext) == 4 and torch.equal(ext.points_list()[1], a) and torch.equal(ext.points_list()[2], b)
    assert torch.equal(pc[1].points_packed(), b) and len(pc[:1]) == 1
    off = pc.offset(torch.ones(8, 3))
    assert torch.allclose(off.points_packed(), pc.points_packed() + 1) and torch.equal(off.features_packed(), pc.features_packed())
    assert torch.allclose(pc.scale(2.0).points_list()[1], 2 * b)
    upd = pc.update_padded(padded * 0 + 7)
    assert (upd.points_packed() == 7).all() and upd.points_list()[1].shape == (3, 3)
    p, nrm, f = pc.get_cloud(1)
    assert torch.equal(p, b) and nrm is None and torch.equal(f, fb)
    table = pc.view_table()
    assert table.N == 2 and table.host[:, 0].tolist() == [0, 5] and table.host[:, 1].tolist() == [5, 3]
    assert table.host[:, 3].tolist() == [0, 5] and table.total_ndc_verts == 8
    joined = trb.join_pointclouds_as_batch([pc, pc[0]])
    assert len(joined) == 3 and joined.features_packed().shape == (13, 4)
    with pytest.raises(ValueError):
        trb.Pointclouds([a, b], features=[fa])
    with pytest.raises(ValueError):
        trb.Pointclouds(torch.rand(5, 3))
    with pytest.raises(ValueError):
        pc.extend(0)


def test_points_call_surface_and_errors_without_gpu():
    from torch_renderer_b200.points_renderer import _background_tensor, _packed_radius
    pc = trb.Pointclouds([torch.rand(5, 3), torch.rand(3, 3)], features=[torch.rand(5, 3), torch.rand(3, 3)])
    assert _packed_radius(0.02, pc).tolist() == pytest.approx([0.02] * 8)
    per_point = torch.rand(2, 5)
    assert torch.equal(_packed_radius(per_point, pc), torch.cat([per_point[0], per_point[1, :3]]))
    with pytest.raises(ValueError):
        _packed_radius(torch.rand(3, 5), pc)
    assert _background_tensor((0.1, 0.2, 0.3), 4, "cpu").tolist() == pytest.approx([0.1, 0.2, 0.3, 1.0])
    assert _background_tensor(0.5, 3, "cpu").tolist() == [0.5] * 3 and _background_tensor(None, 3, "cpu") is None
    with pytest.raises(ValueError):
        _background_tensor((0.1, 0.2), 4, "cpu")
    settings = trb.PointsRasterizationSettings()
    assert (settings.image_size, settings.radius, settings.points_per_pixel) == (256, 0.01, 8)
    renderer = trb.PointsRenderer(trb.PointsRasterizer(trb.FoVPerspectiveCameras(), settings), trb.AlphaCompositor())
    with pytest.raises(RuntimeError, match="CUDA"):      # no CPU fallback
        renderer(pc)
    with pytest.raises(ValueError):
        trb.PointsRasterizer(None, settings)(pc)
    with pytest.raises(ValueError):
        trb.rasterize_points(pc, 16, 0.1, 151)
    # C ABI argument checks run before any CUDA call
    L, p = _lib.lib(), 8
    assert L.trb_points_raster_forward(p, p, p, 1, 16, 16, 151, p, p, p, 0, None) == _lib.TRB_ERR_K_TOO_LARGE
    assert L.trb_points_raster_forward(p, p, p, 1, 0, 16, 1, p, p, p, 0, None) == _lib.TRB_ERR_BAD_ARG
    assert L.trb_points_raster_forward(p, p, p, 1, 16, 16, 1, None, p, p, 0, None) == _lib.TRB_ERR_BAD_ARG
    assert L.trb_points_raster_forward(p, p, p, 0, 16, 16, 1, p, p, p, 0, None) == _lib.TRB_OK
    assert L.trb_points_raster_backward(None, p, p, p, 1, 16, 16, 1, p, 0, None) == _lib.TRB_ERR_BAD_ARG
    assert L.trb_points_composite_forward(2, p, p, p, 10, 1, 3, None, p, 0, None) == _lib.TRB_ERR_BAD_ARG
    assert L.trb_points_composite_forward(0, p, p, p, 10, 151, 3, None, p, 0, None) == _lib.TRB_ERR_K_TOO_LARGE
    assert L.trb_points_composite_forward(1, p, p, p, 0, 1, 3, None, p, 0, None) == _lib.TRB_OK
    assert L.trb_points_composite_backward(0, p, p, p, None, 10, 1, 3, 0, p, p, 0, None) == _lib.TRB_ERR_BAD_ARG


def test_oracle_point_backward_matches_fp64_autograd():
    torch.manual_seed(0)
    pts = torch.rand(120, 3) * torch.tensor([2.0, 2.0, 2.0]) - torch.tensor([1.0, 1.0, 0.1])
    H, W, K, r = 12, 10, 3, 0.25
    idx, z, d = pr.rasterize_points(pts.numpy(), [0, 70], [70, 50], r, (H, W), K)
    gz, gd = torch.randn(2, H, W, K), torch.randn(2, H, W, K)
    m = torch.from_numpy(idx >= 0)
    got = pr.rasterize_points_backward(pts.numpy(), idx, (gz * m).numpy(), (gd * m).numpy())
    from oracle import shading_ref as sref
    p64 = pts.double().requires_grad_(True)
    ys, xs = sref.pixel_centers(H, W, torch.float64)
    sel = p64[torch.from_numpy(idx).long().clamp(min=0)]
    d2 = (xs.view(1, 1, W, 1) - sel[..., 0]) ** 2 + (ys.view(1, H, 1, 1) - sel[..., 1]) ** 2
    assert np.allclose(d2.detach().numpy()[idx >= 0], d[idx >= 0], atol=1e-6)
    ((d2 * gd * m).sum() + (sel[..., 2] * gz * m).sum()).backward()
    assert np.allclose(got, p64.grad.numpy(), rtol=1e-5, atol=1e-6)
