"""Point-set and mesh-regulariser ops next to the renderer: the torch-composed regularisers on
CPU against the numpy oracle and hand-checkable meshes; the CUDA nearest-neighbour kernel of the chamfer distance
on the GPU against a float64 restatement (values and gradients)."""
import numpy as np
import pytest
import torch

import torch_renderer_b200 as trb
from oracle import points_ref
from helpers import uv_sphere

DEV = torch.device("cuda:0")


def _two_triangles(angle_deg):
    a = np.radians(angle_deg)
    v = torch.tensor([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.5, 1.0, 0.0],
                      [0.5, -np.cos(a), np.sin(a)]], dtype=torch.float32)
    f = torch.tensor([[0, 1, 2], [1, 0, 3]])
    return v, f


def test_regularisers_match_oracle_and_hand_cases():
    v, f = uv_sphere(6, 8, noise=0.1, seed=2)
    m = trb.Meshes([v], [f])
    vn, fn = v.double().numpy(), f.numpy()
    assert abs(float(trb.mesh_edge_loss(m)) - points_ref.edge_loss(vn, fn)) < 1e-5
    assert abs(float(trb.mesh_edge_loss(m, target_length=0.3)) - points_ref.edge_loss(vn, fn, 0.3)) < 1e-5
    assert abs(float(trb.mesh_laplacian_smoothing(m, method="uniform")) - points_ref.laplacian_uniform(vn, fn)) < 1e-5
    assert abs(float(trb.mesh_normal_consistency(m)) - points_ref.normal_consistency(vn, fn)) < 1e-5
    # flat pair of triangles: consistent normals; folded by 90 degrees: 1 - cos(90) = 1
    for ang, want in ((0.0, 0.0), (90.0, 1.0), (180.0, 2.0)):
        vv, ff = _two_triangles(ang)
        assert abs(float(trb.mesh_normal_consistency(trb.Meshes([vv], [ff]))) - want) < 1e-5, ang
    # batch of two meshes: mean of the per-mesh values
    v2, f2 = uv_sphere(4, 5, noise=0.05, seed=7)
    mb = trb.Meshes([v, v2], [f, f2])
    for fn_, ref in ((trb.mesh_edge_loss, points_ref.edge_loss), (trb.mesh_laplacian_smoothing, points_ref.laplacian_uniform),
                     (trb.mesh_normal_consistency, points_ref.normal_consistency)):
        want = 0.5 * (ref(vn, fn) + ref(v2.double().numpy(), f2.numpy()))
        assert abs(float(fn_(mb)) - want) < 1e-5
    # gradients flow to the vertices
    vg = v.clone().requires_grad_(True)
    mg = trb.Meshes([vg], [f])
    (trb.mesh_edge_loss(mg) + trb.mesh_laplacian_smoothing(mg) + trb.mesh_normal_consistency(mg)).backward()
    assert vg.grad is not None and torch.isfinite(vg.grad).all() and vg.grad.abs().sum() > 0
    with pytest.raises(NotImplementedError):
        trb.mesh_laplacian_smoothing(m, method="cot")


def test_sample_points_from_meshes_on_surface_and_by_area():
    torch.manual_seed(0)
    # two triangles of area 0.5 and 4.5 in the z = 0 plane: samples stay in-plane and follow the area ratio
    v = torch.tensor([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [10, 0, 0], [13, 0, 0], [10, 3, 0]])
    f = torch.tensor([[0, 1, 2], [3, 4, 5]])
    pts, nrm = trb.sample_points_from_meshes(trb.Meshes([v], [f]), 4000, return_normals=True)
    assert pts.shape == (1, 4000, 3) and (pts[..., 2].abs() < 1e-6).all()
    big = (pts[0, :, 0] > 5).float().mean().item()
    assert abs(big - 0.9) < 0.03
    assert torch.allclose(nrm[0, :, 2].abs(), torch.ones(4000), atol=1e-5)
    small = pts[0][pts[0, :, 0] < 5]
    assert (small[:, 0] >= 0).all() and (small[:, 1] >= 0).all() and (small[:, 0] + small[:, 1] <= 1 + 1e-5).all()
    vg = v.clone().requires_grad_(True)
    trb.sample_points_from_meshes(trb.Meshes([vg], [f]), 100).sum().backward()
    assert vg.grad.abs().sum() > 0


@pytest.mark.gpu
def test_chamfer_distance_cuda_matches_oracle():
    torch.manual_seed(0)
    for N, P1, P2 in ((1, 1000, 1000), (3, 257, 700), (2, 5, 1)):
        x = torch.randn(N, P1, 3)
        y = torch.randn(N, P2, 3) * 1.2 + 0.1
        xd, yd = x.to(DEV).requires_grad_(True), y.to(DEV).requires_grad_(True)
        loss, nrm = trb.chamfer_distance(xd, yd)
        assert nrm is None
        x64, y64 = x.double().requires_grad_(True), y.double().requires_grad_(True)
        want = points_ref.chamfer(x64, y64)
        assert abs(float(loss.detach()) - float(want.detach())) < 1e-5 * max(1.0, float(want.detach()))
        loss.backward(); want.backward()
        assert torch.allclose(xd.grad.cpu().double(), x64.grad, atol=1e-6, rtol=1e-4)
        assert torch.allclose(yd.grad.cpu().double(), y64.grad, atol=1e-6, rtol=1e-4)
    d, idx = trb.loss.nearest_points(x.to(DEV), y.to(DEV))
    ref = ((x[:, :, None] - y[:, None]) ** 2).sum(-1).min(2)
    assert torch.equal(idx.cpu().long(), ref[1]) and torch.allclose(d.cpu(), ref[0], atol=1e-6)
    per = trb.chamfer_distance(x.to(DEV), y.to(DEV), batch_reduction=None)[0]
    assert per.shape == (2,)
    with pytest.raises(RuntimeError, match="CUDA"):
        trb.chamfer_distance(x, y)


@pytest.mark.gpu
def test_deformation_losses_step_like_the_reference():
    """One optimisation step of mesh_deformer.py:300-330: sample both meshes, chamfer + edge + normal + laplacian
    with the script's weights, backward to the vertex offsets."""
    torch.manual_seed(0)
    src = trb.ico_sphere(3, device=DEV)
    v, f = src.get_mesh_verts_faces(0)
    trg = trb.Meshes([v * torch.tensor([1.3, 0.8, 1.0], device=DEV)], [f])
    deform = torch.zeros_like(v, requires_grad=True)
    opt = torch.optim.SGD([deform], lr=1.0, momentum=0.9)
    losses = []
    for _ in range(30):
        opt.zero_grad()
        new = src.offset_verts(deform)
        l_ch, _ = trb.chamfer_distance(trb.sample_points_from_meshes(trg, 1000), trb.sample_points_from_meshes(new, 1000))
        loss = l_ch + 1.0 * trb.mesh_edge_loss(new) + 0.01 * trb.mesh_normal_consistency(new) \
            + 0.1 * trb.mesh_laplacian_smoothing(new, method="uniform")
        loss.backward()
        opt.step()
        losses.append(float(l_ch.detach()))
    assert losses[-1] < 0.5 * losses[0], losses[::6]


def test_regularisers_isolated_vertices_and_non_manifold_edges():
    """Upstream keeps L_ii = -1 for a vertex no face uses (it contributes |v|), and pairs EVERY two faces that
    share an edge, however many there are (a fan of 6 faces around one edge: 15 pairs)."""
    verts = torch.tensor([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [3.0, 4.0, 0.0]])     # vertex 3 is isolated
    faces = torch.tensor([[0, 1, 2]])
    got = float(trb.mesh_laplacian_smoothing(trb.Meshes([verts], [faces])))
    assert abs(got - points_ref.laplacian_uniform(verts.numpy(), faces.numpy())) < 1e-6
    assert got > 5.0 / 4 - 1e-6                                                   # |(3,4,0)| / V alone is 1.25
    g = torch.Generator().manual_seed(3)
    fan_v = torch.cat([torch.tensor([[0.0, 0, 0], [0, 0, 1.0]]), torch.randn(6, 3, generator=g)])
    fan_f = torch.tensor([[0, 1, 2 + i] for i in range(6)])
    got = float(trb.mesh_normal_consistency(trb.Meshes([fan_v], [fan_f])))
    assert abs(got - points_ref.normal_consistency(fan_v.numpy().astype(np.float64), fan_f.numpy())) < 1e-5
