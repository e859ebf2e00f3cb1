"""CPU tests of the oracle itself: hand-checkable scenes pinning the conventions of SURVEY.md
Appendix A (pixel centres, x/y flip, tie-break, blur band, K ordering, -1 fill), the C backward
against fp64 autograd, and the one reference-owned golden vector (gradient.log:1-6)."""
import math

import numpy as np
import pytest
import torch

import oracle
from oracle import shading_ref as sref
from helpers import load_mesh, normalize_mesh, oracle_rasterize, rel_l2, uv_sphere


def _raster(fv, H=8, W=8, K=1, blur=0.0, persp=False, clip=False, cull=False, first=None, count=None):
    fv = np.asarray(fv, np.float32).reshape(-1, 3, 3)
    first = np.array([0], np.int64) if first is None else first
    count = np.array([fv.shape[0]], np.int64) if count is None else count
    return oracle.rasterize_forward(fv, first, count, (H, W), blur, K, persp, clip, cull, 1)


def test_pixel_convention_top_left_is_plus_x_plus_y():
    # a triangle entirely in the +x,+y NDC quadrant lands in the TOP-LEFT quadrant of the image (A3)
    tri = [[[0.1, 0.1, 1.0], [0.9, 0.1, 1.0], [0.1, 0.9, 1.0]]]
    p2f, zbuf, bary, dists = _raster(tri, 16, 16)
    ys, xs = np.nonzero(p2f[0, :, :, 0] >= 0)
    assert len(ys) > 0
    assert ys.max() < 8 and xs.max() < 8
    # pixel centres: 16 px span [-1,1]; col 0 centre is +0.9375
    assert math.isclose(float(sref.pixel_centers(16, 16)[1][0]), 0.9375)
    assert (zbuf[p2f >= 0] == 1.0).all()
    assert (dists[p2f >= 0] <= 0).all()


def test_background_is_minus_one_everywhere():
    tri = [[[0.1, 0.1, 1.0], [0.9, 0.1, 1.0], [0.1, 0.9, 1.0]]]
    p2f, zbuf, bary, dists = _raster(tri, 8, 8, K=3)
    bg = p2f < 0
    assert bg.any()
    assert (p2f[bg] == -1).all() and (zbuf[bg] == -1).all() and (dists[bg] == -1).all()
    assert (bary[bg] == -1).all()
    # only one face: layers 1,2 are always empty
    assert (p2f[..., 1:] == -1).all()


def test_depth_order_and_index_tiebreak():
    big = lambda z: [[-0.9, -0.9, z], [0.9, -0.9, z], [0.0, 0.9, z]]
    # faces 0 and 2 at the same depth, face 1 nearer
    p2f, zbuf, _, _ = _raster([big(2.0), big(1.0), big(2.0)], 8, 8, K=3)
    covered = p2f[0, :, :, 0] >= 0
    assert covered.sum() > 4
    assert (p2f[0][covered][:, 0] == 1).all()
    assert (p2f[0][covered][:, 1] == 0).all()  # tie on z -> lower face index first
    assert (p2f[0][covered][:, 2] == 2).all()
    assert (np.diff(zbuf[0][covered], axis=1) >= 0).all()
    # K=1 keeps the nearest only; K=2 drops the larger index of the tie
    p2f2, _, _, _ = _raster([big(2.0), big(1.0), big(2.0)], 8, 8, K=2)
    assert (p2f2[0][covered] == np.array([1, 0])).all()


def test_faces_behind_camera_are_dropped():
    tri = lambda z0, z1, z2: [[-0.9, -0.9, z0], [0.9, -0.9, z1], [0.0, 0.9, z2]]
    p2f, *_ = _raster([tri(-1, -1, -1)], 8, 8)
    assert (p2f == -1).all()
    # ANY vertex at/behind the plane kills the whole face (A4.2)
    p2f, *_ = _raster([tri(1.0, 1.0, 0.0)], 8, 8)
    assert (p2f == -1).all()
    p2f, *_ = _raster([tri(1.0, 1.0, 1e-3)], 8, 8)
    assert (p2f >= 0).any()


def test_degenerate_and_backface_culling():
    line = [[-0.5, -0.5, 1.0], [0.0, 0.0, 1.0], [0.5, 0.5, 1.0]]
    p2f, *_ = _raster([line], 8, 8)
    assert (p2f == -1).all()
    ccw = [[-0.9, -0.9, 1.0], [0.9, -0.9, 1.0], [0.0, 0.9, 1.0]]
    cw = [ccw[0], ccw[2], ccw[1]]
    a, *_ = _raster([ccw], 8, 8, cull=True)
    b, *_ = _raster([cw], 8, 8, cull=True)
    assert ((a >= 0).any()) != ((b >= 0).any())  # exactly one winding survives culling
    c, *_ = _raster([cw], 8, 8, cull=False)
    assert (c >= 0).sum() == max((a >= 0).sum(), (b >= 0).sum())


def test_blur_band_uses_squared_distance():
    tri = [[[-0.5, -0.5, 1.0], [0.5, -0.5, 1.0], [0.0, 0.5, 1.0]]]
    hard, _, _, d0 = _raster(tri, 32, 32, blur=0.0)
    r = 0.2
    soft, _, _, d1 = _raster(tri, 32, 32, blur=r * r, clip=True)
    inside = hard[0, :, :, 0] >= 0
    band = (soft[0, :, :, 0] >= 0) & ~inside
    assert band.sum() > 0
    assert (d1[0, :, :, 0][band] > 0).all() and (d1[0, :, :, 0][band] < r * r).all()
    assert (d1[0, :, :, 0][inside] <= 0).all()
    # a pixel just below the bottom edge (y=-0.5): distance is vertical offset squared
    ys, xs = sref.pixel_centers(32, 32)
    yi = int(np.argmin(np.abs(ys.numpy() + 0.5 + 0.03125)))
    xi = 16
    py, px = float(ys[yi]), float(xs[xi])
    assert band[yi, xi]
    assert math.isclose(float(d1[0, yi, xi, 0]), (py + 0.5) ** 2, rel_tol=1e-5)


def test_perspective_correct_zbuf():
    # vertex depths 1 and 3: at the screen-space midpoint perspective-correct z is the harmonic mix
    tri = [[[-0.8, -0.8, 1.0], [0.8, -0.8, 3.0], [0.0, 0.9, 3.0]]]
    _, z_lin, b_lin, _ = _raster(tri, 64, 64, persp=False)
    p2f, z_pc, b_pc, _ = _raster(tri, 64, 64, persp=True)
    m = p2f[0, :, :, 0] >= 0
    w = b_lin[0, :, :, 0][m].astype(np.float64)  # screen-space barycentrics
    zs = np.array([1.0, 3.0, 3.0])
    expect = 1.0 / (w[:, 0] / zs[0] + w[:, 1] / zs[1] + w[:, 2] / zs[2])
    assert np.allclose(z_pc[0, :, :, 0][m], expect, rtol=1e-4)
    assert np.allclose(b_pc[0, :, :, 0][m].sum(-1), 1.0, atol=1e-5)


def test_non_square_ranges():
    ys, xs = sref.pixel_centers(180, 320)
    assert math.isclose(float(xs[0]), 1.77222, abs_tol=1e-4) and math.isclose(float(ys[0]), 0.99444, abs_tol=1e-4)
    tri = [[[1.2, -0.5, 1.0], [1.7, -0.5, 1.0], [1.45, 0.5, 1.0]]]  # beyond |x|=1 but inside the wide image
    p2f, *_ = _raster(tri, 18, 32)
    ys_, xs_ = np.nonzero(p2f[0, :, :, 0] >= 0)
    assert len(xs_) > 0 and xs_.max() < 6


def test_faces_per_pixel_limit():
    tri = [[[0.1, 0.1, 1.0], [0.9, 0.1, 1.0], [0.1, 0.9, 1.0]]]
    with pytest.raises(ValueError):
        _raster(tri, 4, 4, K=151)
    _raster(tri, 4, 4, K=150)


def test_batched_meshes_only_see_their_own_faces():
    a = [[0.1, 0.1, 1.0], [0.9, 0.1, 1.0], [0.1, 0.9, 1.0]]
    b = [[-0.1, -0.1, 1.0], [-0.9, -0.1, 1.0], [-0.1, -0.9, 1.0]]
    p2f, *_ = _raster([a, b], 8, 8, first=np.array([0, 1], np.int64), count=np.array([1, 1], np.int64))
    assert set(np.unique(p2f[0])) == {-1, 0}
    assert set(np.unique(p2f[1])) == {-1, 1}


def test_threads_do_not_change_the_result():
    v, f = load_mesh("teapot")
    v = normalize_mesh(v) * 0.8
    v = v + torch.tensor([0, 0, 3.0])
    ndc = torch.stack([v[:, 0] / v[:, 2] * 1.7, v[:, 1] / v[:, 2] * 1.7, v[:, 2]], -1)[None]
    r1 = oracle_rasterize(ndc, f, (48, 48), 1e-4, 4, True, True, False, threads=1)
    r8 = oracle_rasterize(ndc, f, (48, 48), 1e-4, 4, True, True, False, threads=0)
    for a, b in zip(r1, r8):
        assert np.array_equal(a, b)
    assert (r1[0] >= 0).sum() > 200


@pytest.mark.parametrize("persp,clip,blur", [(False, False, 0.0), (True, False, 0.0), (True, True, 2e-3)])
def test_c_backward_matches_fp64_autograd(persp, clip, blur):
    torch.manual_seed(0)
    v, f = uv_sphere(6, 8, 0.7, noise=0.05, seed=1)
    v = v + torch.tensor([0.1, -0.05, 2.5])
    ndc = torch.stack([v[:, 0] / v[:, 2] * 1.7, v[:, 1] / v[:, 2] * 1.7, v[:, 2]], -1)
    K, H, W = 3, 24, 20
    p2f, zbuf, bary, dists = oracle_rasterize(ndc[None], f, (H, W), blur, K, persp, clip)
    fv32 = ndc[f]
    gz, gb, gd = (torch.randn(1, H, W, K), torch.randn(1, H, W, K, 3), torch.randn(1, H, W, K))
    got = oracle.rasterize_backward(fv32.numpy(), p2f, gz.numpy(), gb.numpy(), gd.numpy(), persp, clip)
    fv64 = fv32.double().requires_grad_(True)
    z64, b64, d64 = sref.raster_recompute(fv64, torch.from_numpy(p2f), persp, clip)
    # forward parity of the torch model with the C oracle first
    m = torch.from_numpy(p2f >= 0)
    assert torch.allclose(z64[m].float(), torch.from_numpy(zbuf)[m], atol=1e-5, rtol=1e-5)
    assert torch.allclose(b64[m].float(), torch.from_numpy(bary)[m], atol=1e-5, rtol=1e-4)
    assert torch.allclose(d64[m].float(), torch.from_numpy(dists)[m], atol=1e-6, rtol=1e-4)
    loss = (z64 * gz.double() * m).sum() + (b64 * gb.double() * m[..., None]).sum() + (d64 * gd.double() * m).sum()
    loss.backward()
    assert rel_l2(torch.from_numpy(got), fv64.grad) < 1e-4


def test_interp_oracle_matches_torch():
    torch.manual_seed(0)
    p2f = torch.randint(-1, 5, (2, 3, 4, 2))
    bary = torch.rand(2, 3, 4, 2, 3)
    attrs = torch.randn(5, 3, 4)
    out = oracle.interp_forward(p2f.numpy(), bary.numpy(), attrs.numpy())
    ref = sref.interpolate_face_attributes(p2f, bary, attrs)
    assert np.allclose(out, ref.numpy(), atol=1e-6)
    g = torch.randn_like(ref)
    gb, ga = oracle.interp_backward(p2f.numpy(), bary.numpy(), attrs.numpy(), g.numpy())
    b64, a64 = bary.double().requires_grad_(True), attrs.double().requires_grad_(True)
    (sref.interpolate_face_attributes(p2f, b64, a64) * g.double()).sum().backward()
    m = (p2f >= 0)
    assert np.allclose(gb[m.numpy()], b64.grad[m].numpy(), atol=1e-5)
    assert np.allclose(ga, a64.grad.numpy(), atol=1e-5)


def test_camera_helper_matches_reference_log():
    """gradient.log:1-6 of the reference: look_at_view_transform(0.7, 50, 30) and matrix_to_quaternion."""
    from torch_renderer_b200.cameras import look_at_view_transform
    from torch_renderer_b200.transforms import matrix_to_quaternion, quaternion_to_matrix
    R, T = look_at_view_transform(dist=0.7, elev=50.0, azim=30.0)
    R_ref = torch.tensor([[-0.8660, -0.3830, -0.3214], [0.0000, 0.6428, -0.7660], [0.5000, -0.6634, -0.5567]])
    assert torch.allclose(R[0], R_ref, atol=1e-4)
    assert torch.allclose(T[0], torch.tensor([0.0, 0.0, 0.7]), atol=1e-6)
    q = matrix_to_quaternion(R)[0]
    assert torch.allclose(q, torch.tensor([-0.2346, -0.1094, 0.8754, -0.4082]), atol=1e-4)
    assert torch.allclose(quaternion_to_matrix(q[None])[0], R[0], atol=1e-5)


def test_shading_ref_blend_identities():
    torch.manual_seed(0)
    N, H, W, K = 1, 4, 4, 3
    p2f = torch.randint(-1, 3, (N, H, W, K)).sort(dim=-1, descending=True)[0]
    colors = torch.rand(N, H, W, K, 3)
    zbuf = torch.rand(N, H, W, K) + 1.0
    dists = -torch.rand(N, H, W, K) * 1e-2
    img = sref.softmax_rgb_blend(colors, p2f, zbuf, dists, 1e-4, 1e-4, (0.2, 0.3, 0.4), 1.0, 100.0)
    bg = (p2f < 0).all(-1)
    assert torch.allclose(img[bg][:, :3], torch.tensor([0.2, 0.3, 0.4]).expand(int(bg.sum()), 3), atol=1e-6)
    assert torch.allclose(img[bg][:, 3], torch.zeros(int(bg.sum())))
    sil = sref.sigmoid_alpha_blend(torch.ones(N, H, W, K, 3), p2f, dists, 1e-4)
    assert torch.allclose(sil[..., 3], img[..., 3])
    hard = sref.hard_rgb_blend(colors, p2f, (0.2, 0.3, 0.4))
    assert torch.equal(hard[..., 3] > 0, p2f[..., 0] >= 0)


def test_phong_lighting_hand_values():
    """Known answers worked by hand for the lighting restatement.  Surface point at the origin, normal +z.
    (a) Point light at (0, 0, 2), camera at (0, 0, 5): n.l = 1, the reflection is +z = the view direction, so
        colour = (ambient + diffuse) * texel + specular.
    (b) Light at (2, 0, 2): l = (1, 0, 1)/sqrt2, n.l = 1/sqrt2; r = -l + 2 (n.l) n = (-1, 0, 1)/sqrt2; camera at
        (0, 0, 5): v = +z, r.v = 1/sqrt2, specular = (1/sqrt2)^shininess.
    (c) Light behind the surface: diffuse and specular vanish (the specular mask is n.l > 0).
    (d) Directional light (0, 0, 1) equals (a) without the position dependence; ambient-only lights ignore both."""
    pts = torch.zeros(1, 1, 1, 1, 3)
    nrm = torch.tensor([0.0, 0.0, 1.0]).view(1, 1, 1, 1, 3)
    tex = torch.tensor([0.5, 0.25, 1.0]).view(1, 1, 1, 1, 3)
    one3 = lambda x: torch.tensor([[x, x, x]])
    cam = torch.tensor([[0.0, 0.0, 5.0]])
    shin = torch.tensor([4.0])

    def colour(kind, vec, amb=0.3, dif=0.5, spec=0.2):
        return sref.phong_colors(pts, nrm, tex, None, kind, torch.tensor([vec]), one3(amb), one3(dif), one3(spec),
                                 one3(1.0), one3(1.0), one3(1.0), shin, cam).view(3)

    a = colour("point", [0.0, 0.0, 2.0])
    assert torch.allclose(a, (0.3 + 0.5) * tex.view(3) + 0.2, atol=1e-6)
    b = colour("point", [2.0, 0.0, 2.0])
    c = 2 ** -0.5
    assert torch.allclose(b, (0.3 + 0.5 * c) * tex.view(3) + 0.2 * c ** 4, atol=1e-6)
    behind = colour("point", [0.0, 0.0, -2.0])
    assert torch.allclose(behind, 0.3 * tex.view(3), atol=1e-6)
    d = colour("directional", [0.0, 0.0, 1.0])
    assert torch.allclose(d, a, atol=1e-6)
    assert torch.allclose(colour("ambient", [9.0, 9.0, 9.0]), 0.3 * tex.view(3), atol=1e-6)


def test_vertex_normals_hand_values():
    """Area-weighted vertex normals: a unit right triangle in the z = 0 plane and a twice-as-large one in
    the x = 0 plane sharing the vertex at the origin.  Face normals (v2 - v1) x (v0 - v1): (0, 0, 1) * 1 for the
    first, (1, 0, 0) * 4 for the second (|cross| = 2 * area), so the shared vertex gets normalize((4, 0, 1))."""
    verts = torch.tensor([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 2, 0], [0, 0, 2]])
    faces = torch.tensor([[0, 1, 2], [0, 3, 4]])
    n = sref.vertex_normals(verts, faces)
    assert torch.allclose(n[1], torch.tensor([0.0, 0.0, 1.0])) and torch.allclose(n[2], torch.tensor([0.0, 0.0, 1.0]))
    assert torch.allclose(n[3], torch.tensor([1.0, 0.0, 0.0])) and torch.allclose(n[4], torch.tensor([1.0, 0.0, 0.0]))
    assert torch.allclose(n[0], torch.tensor([4.0, 0.0, 1.0]) / 17 ** 0.5, atol=1e-6)


def test_softmax_blend_hand_values():
    """softmax_rgb_blend worked by hand with sigma = gamma = 1, znear = 1, zfar = 11, two layers:
    layer 0: z = 3, signed distance -ln 3  => p0 = sigmoid(ln 3) = 3/4, z_inv = 0.8 (the maximum), weight 3/4;
    layer 1: z = 6, signed distance  ln 3  => p1 = 1/4,              z_inv = 0.5, weight exp(-0.3) / 4;
    background weight delta = exp(-0.8); alpha = 1 - (1/4)(3/4) = 13/16."""
    import math
    p2f = torch.tensor([[[[5, 9]]]])
    z = torch.tensor([[[[3.0, 6.0]]]])
    d = torch.tensor([[[[-math.log(3.0), math.log(3.0)]]]])
    cols = torch.tensor([[[[[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]]]])
    out = sref.softmax_rgb_blend(cols, p2f, z, d, 1.0, 1.0, (0.0, 0.0, 1.0), 1.0, 11.0).view(4)
    w0, w1, delta = 0.75, 0.25 * math.exp(-0.3), math.exp(-0.8)
    den = w0 + w1 + delta
    assert torch.allclose(out, torch.tensor([w0 / den, w1 / den, delta / den, 13.0 / 16.0]), atol=1e-6)
    # one empty slot (-1) contributes nothing: same pixel with the second layer masked out
    out1 = sref.softmax_rgb_blend(cols, torch.tensor([[[[5, -1]]]]), z, d, 1.0, 1.0, (0.0, 0.0, 1.0), 1.0, 11.0).view(4)
    assert torch.allclose(out1, torch.tensor([w0 / (w0 + delta), 0.0, delta / (w0 + delta), 0.75]), atol=1e-6)
    # the silhouette shader's alpha is the same product
    sil = sref.sigmoid_alpha_blend(cols, p2f, d, 1.0).view(4)
    assert abs(float(sil[3]) - 13.0 / 16.0) < 1e-6
