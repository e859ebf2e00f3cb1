"""Application-level GPU tests: the optimisation loops the reference scripts run, end to end through the public
API, judged by whether they converge (camera_pose_optimizer.py:237-330 -- pose from silhouette + depth + colour;
mesh_deformer.py:181-222 -- per-vertex colours from multi-view images; deform_mesh_from_pcd.py / mesh_deformer.py
-- vertex offsets from silhouettes).  They exercise forward + backward of every fused kernel together with
autograd, the Fragments cache, TexturesUV and per-call camera overrides the way a user of the reference would."""
import math

import numpy as np
import pytest
import torch

from helpers import cow_uvs, load_mesh, normalize_mesh

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
SIGMA = 1e-4
BLUR = math.log(1.0 / 1e-4 - 1.0) * SIGMA


def _trb():
    import torch_renderer_b200 as trb
    return trb


def test_camera_pose_optimisation_converges():
    """Pose = (T, quaternion) optimised against silhouette + depth + colour references of the UV-textured cow."""
    trb = _trb()
    trb.set_fragment_cache(True)   # the three renders of a step share one rasterisation, as in the reference
    torch.manual_seed(0)
    v, f = load_mesh("cow")
    v = normalize_mesh(v)
    vt, ft = cow_uvs()
    yy, xx = torch.meshgrid(torch.linspace(0, 1, 64), torch.linspace(0, 1, 64), indexing="ij")
    tex = torch.stack([xx, yy, 0.5 + 0.5 * torch.sin(12 * xx) * torch.cos(9 * yy)], -1)[None]
    mesh = trb.Meshes([v.to(DEV)], [f.to(DEV)],
                      textures=trb.TexturesUV(maps=tex.to(DEV), faces_uvs=[ft.to(DEV)], verts_uvs=[vt.to(DEV)]))
    cams = trb.FoVPerspectiveCameras(device=DEV)
    blend = trb.BlendParams(SIGMA, 1e-4, (0.0, 0.0, 0.0))
    soft = trb.RasterizationSettings(image_size=128, blur_radius=BLUR, faces_per_pixel=20)
    hard = trb.RasterizationSettings(image_size=128, blur_radius=0.0, faces_per_pixel=1)
    sil = trb.MeshRenderer(trb.MeshRasterizer(cams, soft), trb.SoftSilhouetteShader(blend))
    rast = trb.MeshRasterizer(cams, hard)
    phong = trb.MeshRenderer(trb.MeshRasterizer(cams, hard),
                             trb.SoftPhongShader(device=DEV, cameras=cams, blend_params=blend,
                                                 lights=trb.PointLights(device=DEV, location=[[0.0, 0.0, -3.0]])))
    R_ref, T_ref = trb.look_at_view_transform(2.7, 30.0, 60.0)
    q_ref = torch.cat([T_ref, trb.transforms.matrix_to_quaternion(R_ref)], -1).to(DEV)

    def render(pose):
        R = trb.transforms.quaternion_to_matrix(pose[:, 3:]); T = pose[:, :3]
        depth = torch.relu(rast(meshes_world=mesh, R=R, T=T).zbuf[..., 0])
        alpha = sil(mesh, R=R, T=T)[..., 3]
        rgb = phong(mesh, R=R, T=T)[..., :3]
        return depth, alpha, rgb

    with torch.no_grad():
        d_ref, a_ref, c_ref = render(q_ref)
    mask = d_ref > 0
    R0, T0 = trb.look_at_view_transform(2.9, 22.0, 48.0)
    pose = torch.cat([T0, trb.transforms.matrix_to_quaternion(R0)], -1).to(DEV).requires_grad_(True)
    opt = torch.optim.Adam([pose], lr=0.01)

    def loss_fn():
        depth, alpha, rgb = render(pose)
        both = mask & (depth > 0)
        return ((alpha - a_ref).abs().mean() + torch.nn.functional.huber_loss(depth[both], d_ref[both])
                + 0.1 * ((rgb - c_ref) ** 2).mean())

    def pose_error():
        q = pose.detach()
        qn = q[:, 3:] / q[:, 3:].norm()
        qr = q_ref[:, 3:] / q_ref[:, 3:].norm()
        return float((q[:, :3] - q_ref[:, :3]).norm() + torch.minimum((qn - qr).norm(), (qn + qr).norm()))

    l0, e0 = float(loss_fn().detach()), pose_error()
    for _ in range(150):
        opt.zero_grad()
        loss = loss_fn()
        loss.backward()
        opt.step()
    l1, e1 = float(loss_fn().detach()), pose_error()
    trb.set_fragment_cache(False)
    assert l1 < 0.2 * l0, (l0, l1)
    assert e1 < 0.35 * e0, (e0, e1)


def test_vertex_colour_fitting_converges():
    """mesh_deformer.py:181-222: per-vertex colours fitted to multi-view target images, two random views per
    iteration through per-call `cameras=` / `lights=` overrides, SGD with momentum."""
    trb = _trb()
    torch.manual_seed(0)
    ico = trb.ico_sphere(3, device=DEV)
    v, f = ico.get_mesh_verts_faces(0)
    target_rgb_v = (0.5 + 0.5 * torch.sin(3.0 * v)).clamp(0, 1)
    nv = 8
    R, T = trb.look_at_view_transform(dist=2.7, elev=torch.linspace(-40, 40, nv), azim=torch.linspace(-180, 180, nv))
    cams = [trb.FoVPerspectiveCameras(device=DEV, R=R[i:i + 1].to(DEV), T=T[i:i + 1].to(DEV)) for i in range(nv)]
    lights = trb.AmbientLights(device=DEV)
    rend = trb.MeshRenderer(trb.MeshRasterizer(cams[0], trb.RasterizationSettings(image_size=96, perspective_correct=False)),
                            trb.SoftPhongShader(device=DEV, cameras=cams[0], lights=lights))
    with torch.no_grad():
        tgt_mesh = trb.Meshes([v], [f], textures=trb.TexturesVertex(target_rgb_v[None]))
        targets = [rend(tgt_mesh, cameras=c, lights=lights)[0, ..., :3] for c in cams]
    verts_rgb = torch.full((1, v.shape[0], 3), 0.5, device=DEV, requires_grad=True)
    opt = torch.optim.SGD([verts_rgb], lr=1.0, momentum=0.9)
    mesh = trb.Meshes([v], [f])
    err0 = float((verts_rgb.detach()[0] - target_rgb_v).abs().mean())
    g = torch.Generator().manual_seed(1)
    for _ in range(150):
        opt.zero_grad()
        rgb = torch.nn.functional.hardtanh(verts_rgb, min_val=0.0, max_val=1.0)
        mesh.textures = trb.TexturesVertex(verts_features=rgb)
        loss = 0.0
        for j in torch.randperm(nv, generator=g)[:2].tolist():
            pred = rend(mesh, cameras=cams[j], lights=lights)[0, ..., :3]
            loss = loss + ((pred - targets[j]) ** 2).mean()
        (loss * v.shape[0] / 40.0 + ((rgb - verts_rgb) ** 2).sum()).backward()
        opt.step()
    err1 = float((verts_rgb.detach()[0] - target_rgb_v).abs().mean())
    assert err1 < 0.3 * err0, (err0, err1)


def test_vertex_offsets_from_silhouettes_converge():
    """Vertex offsets of an ico-sphere optimised so that its soft silhouettes from six views match those of an
    ellipsoid (the silhouette term of the reference's deformation scripts), Adam on `deform_verts`."""
    trb = _trb()
    ico = trb.ico_sphere(3, device=DEV)
    v, f = ico.get_mesh_verts_faces(0)
    scale = torch.tensor([1.25, 0.8, 1.0], device=DEV)
    nv = 6
    R, T = trb.look_at_view_transform(dist=3.0, elev=torch.tensor([0.0, 0.0, 0.0, 0.0, 80.0, -80.0]),
                                      azim=torch.tensor([0.0, 90.0, 180.0, 270.0, 0.0, 0.0]))
    cams = trb.FoVPerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV))
    sil = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=96, blur_radius=BLUR,
                                                                              faces_per_pixel=12)),
                           trb.SoftSilhouetteShader(trb.BlendParams(SIGMA, 1e-4, (0.0, 0.0, 0.0))))
    with torch.no_grad():
        target = sil(trb.Meshes([v * scale], [f]).extend(nv))[..., 3]
    deform = torch.zeros_like(v, requires_grad=True)
    opt = torch.optim.Adam([deform], lr=5e-3)

    def loss_fn():
        pred = sil(trb.Meshes([v + deform], [f]).extend(nv))[..., 3]
        return ((pred - target) ** 2).mean()

    l0 = float(loss_fn().detach())
    for _ in range(200):
        opt.zero_grad()
        loss = loss_fn() + 1e-3 * (deform ** 2).mean()
        loss.backward()
        opt.step()
    l1 = float(loss_fn().detach())
    assert l1 < 0.25 * l0, (l0, l1)
    assert torch.isfinite(deform).all()


@pytest.mark.parametrize("kind,K,blur,shader_name", [("point", 1, 0.0, "soft"), ("directional", 4, 1e-3, "soft"),
                                                     ("ambient", 1, 0.0, "soft"), ("point", 2, 0.0, "hard")])
def test_light_and_material_colour_gradients(kind, K, blur, shader_name):
    """d loss / d (light colours, material colours) -- SURVEY 8a row a11 -- against fp64 autograd of the oracle
    shading on the same Fragments; MeshRenderer and the stand-alone shader give the same numbers."""
    import torch_renderer_b200 as trb
    from oracle import shading_ref as sref
    from helpers import fov_proj, oracle_rasterize, rel_l2, uv_sphere
    torch.manual_seed(5)
    v, f = uv_sphere(14, 18, 1.0, noise=0.03, seed=2)
    N, H, W = 2, 56, 64
    R, T = trb.look_at_view_transform(dist=2.6, elev=torch.tensor([15.0, -30.0]), azim=torch.tensor([25.0, 140.0]))
    cols = torch.rand(v.shape[0], 3)
    names = ("ambient_color",) if kind == "ambient" else ("ambient_color", "diffuse_color", "specular_color")
    light_vals = {"ambient_color": [[0.4, 0.5, 0.6]], "diffuse_color": [[0.35, 0.25, 0.3]], "specular_color": [[0.2, 0.3, 0.1]]}
    mat_vals = {"ambient_color": [[0.9, 0.8, 1.0]], "diffuse_color": [[0.7, 1.0, 0.8]], "specular_color": [[1.0, 0.6, 0.9]]}
    lt = {n: torch.tensor(light_vals[n], device=DEV, requires_grad=True) for n in names}
    mt = {n: torch.tensor(mat_vals[n], device=DEV, requires_grad=True) for n in ("ambient_color", "diffuse_color", "specular_color")}
    vec = [[0.5, 1.0, -2.5]]
    if kind == "point":
        lights = trb.PointLights(device=DEV, location=vec, **lt)
    elif kind == "directional":
        lights = trb.DirectionalLights(device=DEV, direction=vec, **lt)
    else:
        lights = trb.AmbientLights(device=DEV, **lt)
    materials = trb.Materials(device=DEV, shininess=20.0, **mt)
    mesh = trb.Meshes([v.to(DEV)], [f.to(DEV)], textures=trb.TexturesVertex(cols.to(DEV)[None])).extend(N)
    cams = trb.FoVPerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV))
    rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=(H, W), blur_radius=blur, faces_per_pixel=K))
    blend = trb.BlendParams(1e-4, 1e-4, (0.2, 0.1, 0.3))
    cls = trb.SoftPhongShader if shader_name == "soft" else trb.HardPhongShader
    shader = cls(device=DEV, cameras=cams, lights=lights, materials=materials, blend_params=blend)
    weights = torch.rand(N, H, W, 4)
    params = list(lt.values()) + list(mt.values())
    trb.set_fragment_cache(False)
    try:
        images, frag = trb.MeshRendererWithFragments(rast, shader)(mesh)
        (images * weights.to(DEV)).sum().backward()
        got = [None if p.grad is None else p.grad.clone().cpu() for p in params]
        for p in params:
            p.grad = None
        images2 = shader(rast(mesh), mesh)
        (images2 * weights.to(DEV)).sum().backward()
        got2 = [None if p.grad is None else p.grad.clone().cpu() for p in params]
        ndc = rast.transform(mesh).cpu().reshape(N, -1, 3)
    finally:
        trb.set_fragment_cache(True)
    assert torch.allclose(images, images2, atol=1e-6)
    for a, b in zip(got, got2):
        assert (a is None) == (b is None) and (a is None or torch.allclose(a, b, rtol=1e-4, atol=1e-5))
    # fp64 oracle on the same Fragments
    want = oracle_rasterize(ndc, f, (H, W), blur, K, True, blur > 0)
    assert np.array_equal(frag.pix_to_face.cpu().numpy(), want[0])
    p2f = torch.from_numpy(want[0])
    v64 = v.double()
    zbuf, bary, dists = sref.raster_recompute(ndc.double()[:, f].reshape(-1, 3, 3), p2f, True, blur > 0)
    l64 = {n: torch.tensor(light_vals[n], dtype=torch.float64, requires_grad=True) for n in names}
    m64 = {n: torch.tensor(mat_vals[n], dtype=torch.float64, requires_grad=True) for n in mt}
    rep = lambda t: t.expand(N, 3)
    zero3 = torch.zeros(N, 3, dtype=torch.float64)
    cam = -torch.matmul(T.double()[:, None, :], torch.linalg.inv(R.double()))[:, 0, :]
    ref = sref.shade(p2f, bary, zbuf, dists, f.repeat(N, 1), v64, sref.vertex_normals(v64, f), cols.double(),
                     shader="soft_phong" if shader_name == "soft" else "hard_phong", light_kind=kind,
                     light_vec=torch.tensor(vec, dtype=torch.float64).expand(N, 3),
                     light_ambient=rep(l64["ambient_color"]),
                     light_diffuse=rep(l64["diffuse_color"]) if kind != "ambient" else zero3,
                     light_specular=rep(l64["specular_color"]) if kind != "ambient" else zero3,
                     mat_ambient=rep(m64["ambient_color"]), mat_diffuse=rep(m64["diffuse_color"]),
                     mat_specular=rep(m64["specular_color"]), shininess=torch.full((N,), 20.0, dtype=torch.float64),
                     camera_center=cam, sigma=1e-4, gamma=1e-4, background=(0.2, 0.1, 0.3), znear=1.0, zfar=100.0)
    assert (images.detach().cpu().double() - ref).abs().max() < 1e-4
    (ref * weights.double()).sum().backward()
    want_grads = [l64[n].grad for n in names] + [m64[n].grad for n in mt]
    for name, a, b in zip(list(names) + ["m_" + n for n in mt], got, want_grads):
        if b is None or (kind == "ambient" and name in ("m_diffuse_color", "m_specular_color")):
            assert a is None or float(a.abs().max()) == 0.0
            continue
        assert rel_l2(a, b) < 1e-3, (name, a, b)


def test_capture_step_replays_equal_eager_and_flag_the_near_plane():
    """``trb.capture_step``: a camera_pose_optimizer-style step (quaternion pose -> silhouette + Phong renders ->
    loss.backward()) replayed from ONE CUDA graph gives the eager step's loss and gradient; moving the camera INTO
    the mesh between replays trips the asynchronous near-plane flag (``NearPlaneCrossed``)."""
    import torch_renderer_b200 as trb
    from helpers import load_mesh, normalize_mesh
    dev = torch.device("cuda:0")
    # torch's rule for whole-step capture: leaves that take part must not have been used on the legacy default
    # stream (their gradient accumulation would be bound to it) -- the step lives on a side stream from the start
    with torch.cuda.stream(torch.cuda.Stream(device=dev)):
        _capture_step_body(trb, dev, load_mesh, normalize_mesh)


def _capture_step_body(trb, dev, load_mesh, normalize_mesh):
    v, f = load_mesh("teapot")
    v = normalize_mesh(v)
    torch.manual_seed(0)
    mesh = trb.Meshes([v.to(dev)], [f.to(dev)], textures=trb.TexturesVertex(torch.rand(1, v.shape[0], 3, device=dev)))
    cams = trb.FoVPerspectiveCameras(device=dev)
    settings = trb.RasterizationSettings(image_size=96, blur_radius=9.21024e-4, faces_per_pixel=10)
    sil = trb.MeshRenderer(trb.MeshRasterizer(cams, settings), trb.SoftSilhouetteShader(trb.BlendParams(1e-4, 1e-4, (0, 0, 0))))
    phong = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=96)),
                             trb.SoftPhongShader(device=dev, cameras=cams, lights=trb.PointLights(device=dev, location=[[0.0, 0.0, -3.0]])))
    R, T = trb.look_at_view_transform(2.7, 20, 40)
    pose = torch.cat([T, trb.transforms.matrix_to_quaternion(R)], -1).to(dev).requires_grad_(True)
    target = torch.rand(1, 96, 96, device=dev)
    loss_out = torch.zeros((), device=dev)

    def step():
        pose.grad = None
        Rm = trb.transforms.quaternion_to_matrix(pose[:, 3:])
        Tm = pose[:, :3]
        loss = (sil(mesh, R=Rm, T=Tm)[..., 3] - target).abs().mean() + (phong(mesh, R=Rm, T=Tm)[..., :3] ** 2).mean()
        loss.backward()
        loss_out.copy_(loss.detach())
        return loss_out

    step()
    want_loss, want_grad = float(loss_out), pose.grad.clone()
    cap = trb.capture_step(step)
    for _ in range(3):
        got = cap()
    cap.check()
    assert abs(float(got) - want_loss) < 1e-6
    assert (pose.grad - want_grad).abs().max() < 1e-5 * want_grad.abs().max()
    # inputs are updated in place between replays: a different pose gives the eager result for that pose
    with torch.no_grad():
        pose[:, 2] += 0.3
    cap()
    cap.check()
    g_replay, l_replay = pose.grad.clone(), float(loss_out)
    step()
    assert abs(float(loss_out) - l_replay) < 1e-6
    assert (pose.grad - g_replay).abs().max() < 1e-5 * g_replay.abs().max()
    # camera moved into the mesh: vertices behind z_clip = znear / 2 -> the flag trips, check() raises
    with torch.no_grad():
        pose[:, :3] = torch.tensor([[0.0, 0.0, 0.2]], device=dev)
    cap()
    with pytest.raises(trb.NearPlaneCrossed):
        cap.check()
    with pytest.raises(trb.NearPlaneCrossed):
        cap()


def test_captured_pose_optimisation_with_adam_converges():
    """The whole optimisation step of the reference's camera_pose_optimizer loop (:299-329) -- zero_grad, silhouette
    + Phong renders of the pose, loss.backward(), Adam.step() -- captured ONCE with trb.capture_step and replayed:
    the pose converges like the eager loop, and the near-plane flag stays down."""
    trb = _trb()
    with torch.cuda.stream(torch.cuda.Stream(device=DEV)):
        torch.manual_seed(0)
        v, f = load_mesh("teapot")
        v = normalize_mesh(v)
        mesh = trb.Meshes([v.to(DEV)], [f.to(DEV)],
                          textures=trb.TexturesVertex(torch.rand(1, v.shape[0], 3, device=DEV)))
        cams = trb.FoVPerspectiveCameras(device=DEV)
        blend = trb.BlendParams(SIGMA, 1e-4, (0.0, 0.0, 0.0))
        sil = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=96, blur_radius=BLUR,
                                                                                  faces_per_pixel=20)),
                               trb.SoftSilhouetteShader(blend))
        phong = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=96)),
                                 trb.SoftPhongShader(device=DEV, cameras=cams, blend_params=blend,
                                                     lights=trb.PointLights(device=DEV, location=[[0.0, 0.0, -3.0]])))
        R_ref, T_ref = trb.look_at_view_transform(2.7, 30.0, 60.0)
        q_ref = torch.cat([T_ref, trb.transforms.matrix_to_quaternion(R_ref)], -1).to(DEV)

        def render(pose):
            R = trb.transforms.quaternion_to_matrix(pose[:, 3:]); T = pose[:, :3]
            return sil(mesh, R=R, T=T)[..., 3], phong(mesh, R=R, T=T)[..., :3]

        with torch.no_grad():
            a_ref, c_ref = render(q_ref)
        R0, T0 = trb.look_at_view_transform(2.9, 22.0, 48.0)
        pose = torch.cat([T0, trb.transforms.matrix_to_quaternion(R0)], -1).to(DEV).requires_grad_(True)
        opt = torch.optim.Adam([pose], lr=0.01, capturable=True)
        loss_out = torch.zeros((), device=DEV)

        def step():
            opt.zero_grad(set_to_none=False) if pose.grad is not None else None
            alpha, rgb = render(pose)
            loss = (alpha - a_ref).abs().mean() + 0.1 * ((rgb - c_ref) ** 2).mean()
            loss.backward()
            opt.step()
            loss_out.copy_(loss.detach())
            return loss_out

        def pose_error():
            q = pose.detach()
            qn = q[:, 3:] / q[:, 3:].norm()
            qr = q_ref[:, 3:] / q_ref[:, 3:].norm()
            return float((q[:, :3] - q_ref[:, :3]).norm() + torch.minimum((qn - qr).norm(), (qn + qr).norm()))

        e0 = pose_error()
        cap = trb.capture_step(step, warmup=3)        # 3 warm-up steps + the captured one already move the pose
        l0 = float(cap())
        for _ in range(150):
            cap()
        cap.check()
        l1, e1 = float(loss_out), pose_error()
        assert l1 < 0.3 * l0, (l0, l1)
        assert e1 < 0.5 * e0, (e0, e1)
