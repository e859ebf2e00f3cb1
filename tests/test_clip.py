"""Near-plane clipping / frustum culling: the numpy + C oracle on hand-checkable scenes, and the
product's torch ``clip_faces`` against it.  CPU only; the CUDA side is in test_gpu_clip.py."""
import math

import numpy as np
import pytest
import torch

import oracle
from oracle import clip_ref
from helpers import oracle_rasterize_clipped, uv_sphere

f32 = np.float32


def _frustum(persp, zc=0.5, cull=False):
    return clip_ref.rasterizer_frustum(persp, zc, cull)


def test_untouched_when_nothing_is_behind_or_culled():
    fv = np.array([[[0, 0, 1], [1, 0, 2], [0, 1, 3]]], f32)
    cf = clip_ref.clip_faces(fv, [0], [1], _frustum(True))
    assert cf.face_verts is not None and cf.faces_clipped_to_unclipped_idx is None
    assert cf.barycentric_conversion is None and cf.clipped_faces_neighbor_idx is None
    assert np.array_equal(cf.face_verts, fv)


@pytest.mark.parametrize("persp", [False, True])
def test_one_vertex_in_front_gives_one_triangle(persp):
    # vertices 0 and 2 behind z = 0.5; vertex 1 (z = 1.5) in front: pivot = vertex 1, p2 = vertex 2, p3 = vertex 0
    tri = np.array([[[-1.0, 0.0, 0.25], [0.0, 1.0, 1.5], [1.0, 0.0, 0.0]]], f32)
    cf = clip_ref.clip_faces(tri, [0], [1], _frustum(persp))
    assert cf.face_verts.shape == (1, 3, 3) and list(cf.num_faces_per_mesh) == [1]
    p4, p5, p1 = cf.face_verts[0]
    assert np.array_equal(p1, tri[0, 1])
    w2 = (1.5 - 0.5) / (1.5 - 0.0)      # towards vertex 2
    w3 = (1.5 - 0.5) / (1.5 - 0.25)     # towards vertex 0
    assert p4[2] == pytest.approx(0.5, abs=1e-6) and p5[2] == pytest.approx(0.5, abs=1e-6)
    if not persp:
        assert np.allclose(p4[:2], tri[0, 1, :2] * (1 - w2) + tri[0, 2, :2] * w2, atol=1e-6)
        assert np.allclose(p5[:2], tri[0, 1, :2] * (1 - w3) + tri[0, 0, :2] * w3, atol=1e-6)
    else:
        # x, y are NDC = view / z: the cut point is linear in VIEW space
        view = tri[0].copy(); view[:, :2] *= view[:, 2:3]
        q4 = view[1] * (1 - w2) + view[2] * w2
        q5 = view[1] * (1 - w3) + view[0] * w3
        assert np.allclose(p4[:2], q4[:2] / 0.5, atol=1e-6) and np.allclose(p5[:2], q5[:2] / 0.5, atol=1e-6)
    M = cf.barycentric_conversion[0]
    assert np.allclose(M[:, 0], [0, 1 - w2, w2], atol=1e-6)      # p4 in terms of (v0, v1, v2)
    assert np.allclose(M[:, 1], [w3, 1 - w3, 0], atol=1e-6)      # p5
    assert np.array_equal(M[:, 2], np.array([0, 1, 0], f32))     # p1
    assert list(cf.faces_clipped_to_conversion_idx) == [0] and list(cf.clipped_faces_neighbor_idx) == [-1]
    # same winding as the original face
    e = lambda t: (t[1, 0] - t[0, 0]) * (t[2, 1] - t[0, 1]) - (t[1, 1] - t[0, 1]) * (t[2, 0] - t[0, 0])
    if not persp:
        assert np.sign(e(cf.face_verts[0])) == np.sign(e(tri[0]))


def test_one_vertex_behind_gives_two_neighbouring_triangles_and_index_maps():
    # mesh 0: face 0 whole, face 1 entirely behind (removed); mesh 1: face 2 cut into a quadrilateral, face 3 whole
    fv = np.array([
        [[0, 0, 1], [1, 0, 1], [0, 1, 1]],
        [[0, 0, .1], [1, 0, .2], [0, 1, .3]],
        [[0.0, -1.0, 0.25], [1.0, 1.0, 1.0], [-1.0, 1.0, 1.5]],
        [[0, 0, 2], [1, 0, 2], [0, 1, 2]]], f32)
    cf = clip_ref.clip_faces(fv, [0, 2], [2, 2], _frustum(False))
    assert cf.face_verts.shape == (4, 3, 3)
    assert list(cf.mesh_to_face_first_idx) == [0, 1] and list(cf.num_faces_per_mesh) == [1, 3]
    assert list(cf.faces_clipped_to_unclipped_idx) == [0, 2, 2, 3]
    assert list(cf.clipped_faces_neighbor_idx) == [-1, 2, 1, -1]
    assert list(cf.faces_clipped_to_conversion_idx) == [-1, 0, 1, -1]
    t1, t2 = cf.face_verts[1], cf.face_verts[2]
    p2, p3 = fv[2, 1], fv[2, 2]
    assert np.array_equal(t1[1], p2) and np.array_equal(t2[1], p2) and np.array_equal(t2[2], p3)
    assert np.array_equal(t1[2], t2[0])                     # p5 shared
    assert t1[0][2] == pytest.approx(0.5, abs=1e-6) and t1[2][2] == pytest.approx(0.5, abs=1e-6)
    # the two halves tile the part of the face in front of the plane: areas add up
    area = lambda t: 0.5 * abs((t[1, 0] - t[0, 0]) * (t[2, 1] - t[0, 1]) - (t[1, 1] - t[0, 1]) * (t[2, 0] - t[0, 0]))
    w2, w3 = (0.25 - 0.5) / (0.25 - 1.0), (0.25 - 0.5) / (0.25 - 1.5)
    assert area(t1) + area(t2) == pytest.approx(area(fv[2]) * (1 - w2 * w3), rel=1e-5)
    # conversion: clipped barycentrics -> barycentrics of the original face reproduce the same point
    for t, tri in ((0, t1), (1, t2)):
        b = np.array([0.2, 0.3, 0.5], f32)
        point = b @ tri
        assert np.allclose((cf.barycentric_conversion[t] @ b) @ fv[2], point, atol=1e-6)


def test_frustum_culling_only_removes_faces_entirely_beyond_one_plane():
    fv = np.array([
        [[1.5, 0, 1], [2.0, 0, 1], [1.7, 1, 1]],        # all x > 1: gone
        [[0.5, 0, 1], [2.0, 0, 1], [1.7, 1, 1]],        # straddles x = 1: kept whole
        [[1.5, -2, 1], [-2.0, 1.5, 1], [1.5, 1.5, 1]],  # every vertex outside, but no single plane separates: kept
        [[0, -1.5, 1], [1, -1.2, 1], [0, -3, 1]]], f32)  # all y < -1: gone
    cf = clip_ref.clip_faces(fv, [0], [4], clip_ref.rasterizer_frustum(True, None, True))
    assert list(cf.faces_clipped_to_unclipped_idx) == [1, 2] and cf.barycentric_conversion is None
    assert list(cf.num_faces_per_mesh) == [2]
    off = clip_ref.clip_faces(fv, [0], [4], clip_ref.rasterizer_frustum(True, None, False))
    assert off.faces_clipped_to_unclipped_idx is None


def test_product_clip_faces_matches_the_oracle_bit_for_bit():
    from torch_renderer_b200 import clip
    rng = np.random.default_rng(7)
    for trial in range(120):
        F = int(rng.integers(1, 40))
        fv = rng.normal(size=(F, 3, 3)).astype(f32)
        fv[:, :, 2] = rng.uniform(0.0 if trial % 3 else 0.6, 2.0, size=(F, 3)).astype(f32)
        nm = int(rng.integers(1, 4))
        first = np.concatenate([[0], np.sort(rng.integers(0, F + 1, size=nm - 1))]).astype(np.int64)
        count = np.diff(np.concatenate([first, [F]])).astype(np.int64)
        persp, cull = bool(trial % 2), bool((trial // 2) % 2)
        zc = 0.5 if trial % 5 else None
        a = clip_ref.clip_faces(fv, first, count, clip_ref.rasterizer_frustum(persp, zc, cull))
        b = clip.clip_faces(torch.from_numpy(fv), torch.from_numpy(first), torch.from_numpy(count),
                            clip.rasterizer_frustum(persp, zc, cull))
        for name in a._fields:
            x, y = getattr(a, name), getattr(b, name)
            assert (x is None) == (y is None), (trial, name)
            if x is not None:
                assert np.array_equal(np.asarray(x), y.numpy()), (trial, name)
        if a.faces_clipped_to_unclipped_idx is not None and a.face_verts.shape[0] > 0:
            Fc = a.face_verts.shape[0]
            p2f = rng.integers(-1, Fc, size=(1, 4, 4, 3))
            bary = rng.uniform(size=(1, 4, 4, 3, 3)).astype(f32)
            pa, ba = clip_ref.convert_clipped_rasterization_to_original_faces(p2f, bary, a)
            pb, bb = clip.convert_clipped_rasterization_to_original_faces(torch.from_numpy(p2f), torch.from_numpy(bary), b)
            assert np.array_equal(pa, pb.numpy()) and np.array_equal(ba, bb.numpy())


def test_product_clip_faces_gradients_flow_with_constant_cut_weights():
    from torch_renderer_b200 import clip
    fv = torch.tensor([[[0.0, -1.0, 0.25], [1.0, 1.0, 1.0], [-1.0, 1.0, 1.5]]], dtype=torch.float64, requires_grad=True)
    cf = clip.clip_faces(fv, torch.tensor([0]), torch.tensor([1]), clip.rasterizer_frustum(False, 0.5, False))
    w = torch.arange(18, dtype=torch.float64).reshape(2, 3, 3)
    (cf.face_verts * w).sum().backward()
    # t1 = (p4, p2, p5), t2 = (p5, p2, p3); p4 = p1 (1 - w2) + p2 w2, p5 = p1 (1 - w3) + p3 w3 with w2, w3 constants
    w2, w3 = (0.25 - 0.5) / (0.25 - 1.0), (0.25 - 0.5) / (0.25 - 1.5)
    g4, g2a, g5a, g5b, g2b, g3 = w[0, 0], w[0, 1], w[0, 2], w[1, 0], w[1, 1], w[1, 2]
    want = torch.stack([g4 * (1 - w2) + (g5a + g5b) * (1 - w3), g4 * w2 + g2a + g2b, (g5a + g5b) * w3 + g3])
    assert torch.allclose(fv.grad[0], want, atol=1e-12)


# ------------------------------------------------------------------------------------------ the neighbour rule
def _quad_scene():
    """One face cut into a quadrilateral (t1 = row 0, t2 = row 1), drawn orthographically."""
    fv = np.array([[[0.0, -1.2, 0.2], [0.9, 0.8, 1.0], [-0.9, 0.8, 1.4]]], f32)
    return clip_ref.clip_faces(fv, [0], [1], _frustum(False))


def test_neighbour_rule_keeps_at_most_one_half_per_pixel():
    cf = _quad_scene()
    assert list(cf.clipped_faces_neighbor_idx) == [1, 0]
    args = (cf.face_verts, cf.mesh_to_face_first_idx, cf.num_faces_per_mesh, (32, 32))
    plain = oracle.rasterize_forward(*args, 4e-3, 2, False, True, False, 1)
    ruled = oracle.rasterize_forward(*args, 4e-3, 2, False, True, False, 1,
                                     clipped_faces_neighbor_idx=cf.clipped_faces_neighbor_idx)
    both = (plain[0][..., 0] >= 0) & (plain[0][..., 1] >= 0)
    assert both.sum() > 10                                         # the blur band around the shared diagonal
    assert (ruled[0][..., 1] == -1).all()                          # never both
    assert np.array_equal(ruled[0][..., 0] >= 0, plain[0][..., 0] >= 0)   # coverage unchanged
    # where both were candidates the survivor is the one with the smaller unsigned distance (ties: the first)
    d = np.abs(plain[3])
    f_by_slot = plain[0]
    d0 = np.where(f_by_slot[..., 0] == 0, d[..., 0], d[..., 1])    # distance to t1
    d1 = np.where(f_by_slot[..., 0] == 1, d[..., 0], d[..., 1])    # distance to t2
    want = np.where(d1 < d0, 1, 0)
    assert np.array_equal(ruled[0][..., 0][both], want[both])
    # hard edges: interiors are disjoint, the rule never fires
    hard_a = oracle.rasterize_forward(*args, 0.0, 2, False, False, False, 1)
    hard_b = oracle.rasterize_forward(*args, 0.0, 2, False, False, False, 1,
                                      clipped_faces_neighbor_idx=cf.clipped_faces_neighbor_idx)
    assert all(np.array_equal(x, y) for x, y in zip(hard_a, hard_b))


def test_neighbour_rule_is_order_dependent_like_upstream():
    """K = 1: face a (z = 5) arrives first, then t1 (z = 3) pushes it out, then t2 (z = 7, nearer in the image
    plane) REPLACES t1 -- the pixel ends at depth 7 although a face at depth 5 covers it.  An order-free 'drop the
    farther half, then take the top K' rule would answer 5; upstream answers 7, and so does the oracle."""
    a = [[-1, -1, 5], [1, -1, 5], [0, 1, 5]]
    t1 = [[-0.5, -0.6, 3], [0.5, -0.6, 3], [0.0, 0.4, 3]]          # pixel (0, 0) inside, distance to edges ~0.3
    t2 = [[-0.9, -0.9, 7], [0.9, -0.9, 7], [0.0, 0.9, 7]]          # inside too, but deeper inside: |d| larger
    fv = np.array([a, t1, t2], f32)
    nbr = np.array([-1, 2, 1], np.int64)
    out = oracle.rasterize_forward(fv, [0], [3], (1, 1), 1e-4, 1, False, False, False, 1, clipped_faces_neighbor_idx=nbr)
    assert out[0].item() == 1 and out[1].item() == 3.0            # t2's |d| is larger: it is dropped, t1 stays
    # make t2 the one closer to an edge: now it replaces t1 and the pixel reports depth 7
    t2b = [[-0.05, -0.1, 7], [0.9, -0.1, 7], [0.4, 0.9, 7]]       # pixel centre (0,0) just inside, near an edge
    fv2 = np.array([a, t1, t2b], f32)
    out2 = oracle.rasterize_forward(fv2, [0], [3], (1, 1), 1e-4, 1, False, False, False, 1, clipped_faces_neighbor_idx=nbr)
    assert out2[0].item() == 2 and out2[1].item() == pytest.approx(7.0, rel=1e-6)
    plain = oracle.rasterize_forward(fv2, [0], [3], (1, 1), 1e-4, 1, False, False, False, 1)
    assert plain[0].item() == 1 and plain[1].item() == 3.0


# ------------------------------------------------------------------------------------------ physical check
@pytest.mark.parametrize("K,blur", [(1, 0.0), (3, 2e-4)])
def test_ground_plane_through_the_near_plane_has_the_analytic_depth(K, blur):
    """A big quad (two faces) under the camera runs from behind the camera to far in front of it.  With clipping
    every pixel whose viewing ray meets the plane at depth >= z_clip sees it at exactly the ray depth; pixels whose
    ray meets it nearer than the plane see nothing.  Without clipping the faces (a vertex behind the camera) are
    dropped altogether."""
    h = 0.4                                                        # camera height above the plane y_view = -h ... +Y up
    quad = torch.tensor([[-3.0, -h, -1.0], [3.0, -h, -1.0], [3.0, -h, 6.0], [-3.0, -h, 6.0]])
    faces = torch.tensor([[0, 2, 1], [0, 3, 2]])
    t = math.tan(math.radians(30.0))
    ndc = quad.clone()
    ndc[:, 0] = quad[:, 0] / (quad[:, 2] * t)
    ndc[:, 1] = quad[:, 1] / (quad[:, 2] * t)
    H = W = 48
    z_clip = 0.5
    (p2f, zbuf, bary, dists), cf, _ = oracle_rasterize_clipped(ndc[None], faces, (H, W), blur, K, True, blur > 0, False,
                                                               z_clip_value=z_clip, threads=1)
    assert cf.face_verts.shape[0] == 3 and cf.clipped_faces_neighbor_idx.tolist() == [-1, 2, 1]
    assert (cf.face_verts[:, :, 2] >= z_clip - 1e-6).all()
    # pixel rays: y_ndc = Y / (Z t)  =>  the plane Y = -h is met at Z = -h / (y_ndc t)
    ys = np.array([-1 + (2 * (H - 1 - i) + 1) / H for i in range(H)])
    z_ray = np.where(ys < 0, -h / (np.minimum(ys, -1e-9) * t), np.inf)           # [H]
    xs = np.array([-1 + (2 * (W - 1 - i) + 1) / W for i in range(W)])
    x_hit = xs[None, :] * z_ray[:, None] * t                                      # view-space X of the hit
    on_quad = (z_ray[:, None] <= 6.0) & (np.abs(x_hit) <= 3.0)
    margin = 0.02
    sure_in = on_quad & (z_ray[:, None] > z_clip * (1 + margin)) & (z_ray[:, None] < 6.0 * (1 - margin)) & (np.abs(x_hit) < 3.0 * (1 - margin))
    sure_out = ~on_quad | (z_ray[:, None] < z_clip * (1 - margin))
    hit = p2f[0, :, :, 0] >= 0
    if blur == 0.0:
        assert hit[sure_in].all() and not hit[sure_out & (ys[:, None] > -0.97)].any()
    assert sure_in.sum() > 300
    # (with blur the half that survives the neighbour rule next to the shared diagonal may be the one the pixel is
    # just outside of; its clipped barycentrics then give the depth of the diagonal, not of the ray)
    hit = hit & (dists[0, :, :, 0] < 0)
    assert (sure_in & hit).sum() > 300
    assert np.allclose(zbuf[0, :, :, 0][sure_in & hit], np.broadcast_to(z_ray[:, None], (H, W))[sure_in & hit], rtol=2e-4)
    # converted barycentrics refer to the ORIGINAL faces: they reproduce the depth from the original corner depths
    fz = quad[faces][..., 2].numpy()                                              # [2, 3]
    sel = sure_in & hit
    f = p2f[0, :, :, 0][sel]
    b = bary[0, :, :, 0][sel]
    assert set(np.unique(f)) <= {0, 1}
    assert np.allclose((b * fz[f]).sum(-1), zbuf[0, :, :, 0][sel], rtol=2e-4)
    assert np.allclose(b.sum(-1), 1.0, atol=1e-4)
    if K > 1:
        # never both halves of a cut face in one pixel
        raw = oracle_rasterize_clipped(ndc[None], faces, (H, W), blur, K, True, True, False, z_clip_value=z_clip, threads=1)[2]
        nbr = cf.clipped_faces_neighbor_idx
        for k in range(K):
            fk = raw[0, :, :, k]
            partner = np.where(fk >= 0, nbr[np.clip(fk, 0, None)], -2)
            partner = np.where(partner < 0, -2, partner)
            assert not (raw[0] == partner[..., None]).any()
    # no clipping: a vertex behind the camera kills both faces (A4.2)
    unclipped = oracle.rasterize_forward(ndc[faces].numpy(), [0], [2], (H, W), blur, K, True, blur > 0, False, 1)
    assert (unclipped[0] == -1).all()


def test_neighbour_rule_only_matters_where_both_halves_are_candidates():
    """The premise of the GPU design (csrc/clip.cu): upstream's order-dependent queue can differ from the plain
    top-K only at pixels where BOTH halves of some cut face pass the candidate test -- everywhere else the fast
    kernels' result stands.  Checked on random cut soups with the oracle alone (K = 150 lists every candidate)."""
    rng = np.random.default_rng(3)
    for trial, (K, blur, persp) in enumerate([(1, 2e-3, True), (2, 4e-3, False), (3, 1e-3, True), (1, 0.0, True)]):
        F = 70
        c = rng.uniform(-1.1, 1.1, size=(F, 1, 2))
        fv = np.concatenate([c + rng.uniform(-0.35, 0.35, size=(F, 3, 2)), rng.uniform(0.05, 1.6, size=(F, 3, 1))], axis=2).astype(f32)
        cf = clip_ref.clip_faces(fv, [0], [F], _frustum(persp))
        nbr = cf.clipped_faces_neighbor_idx
        assert (nbr >= 0).sum() >= 20
        args = (cf.face_verts, cf.mesh_to_face_first_idx, cf.num_faces_per_mesh, (36, 36), blur)
        ruled = oracle.rasterize_forward(*args, K, persp, blur > 0, False, 1, clipped_faces_neighbor_idx=nbr)
        plain = oracle.rasterize_forward(*args, K, persp, blur > 0, False, 1)
        every = oracle.rasterize_forward(*args, 150, persp, blur > 0, False, 1)[0]      # all candidates of every pixel
        assert (every[..., -1] == -1).all()
        partner = np.where(every >= 0, nbr[np.clip(every, 0, None)], -2)
        partner = np.where(partner < 0, -2, partner)
        both = (every[..., :, None] == partner[..., None, :]).any(axis=(-1, -2))         # some pair fully present
        differs = (ruled[0] != plain[0]).any(axis=-1)
        assert not (differs & ~both).any()
        if blur > 0:
            assert differs.any()
        else:
            assert not differs.any()
