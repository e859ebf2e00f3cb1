"""Step builders for every BASELINE.json config (C1, C3, C4, C5 and the reference's own
``camera_pose_optimizer`` step), through the public API.  Used by ``bench.py`` (the ``other_configs`` / ``c5`` blocks
of its JSON line), ``profiles/bench_configs.py`` and the ncu targets in ``profiles/``.

Product-side only: meshes come straight from ``tests/golden/meshes.npz`` (geometry of the reference's
``data/teapot.obj`` / ``data/cow_mesh/cow.obj``, written by ``tests/golden/make_assets.py``); nothing here imports
``oracle/``.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

import torch_renderer_b200 as trb

ROOT = os.path.dirname(os.path.abspath(__file__))
SIGMA = 1e-4
BLUR = math.log(1.0 / 1e-4 - 1.0) * SIGMA          # 9.21024e-4: the blur radius the reference's soft settings use
_MESHES = None


def load_mesh(name):
    """'teapot' | 'cow' -> (verts f32 [V,3], faces i64 [F,3]) CPU tensors."""
    global _MESHES
    if _MESHES is None:
        _MESHES = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    return (torch.from_numpy(_MESHES[f"{name}_verts"].copy()),
            torch.from_numpy(_MESHES[f"{name}_faces"].astype(np.int64)))


def normalize_mesh(verts):
    c = (verts.max(0)[0] + verts.min(0)[0]) / 2
    s = (verts - c).abs().max()
    return (verts - c) / s


def bview(K, H, W, V, F):
    """algorithmic bytes per view, forward + backward."""
    return 56 * K * H * W + 32 * H * W + 96 * V + 48 * F


def bview_forward(K, H, W, V, F):
    return 28 * K * H * W + 16 * H * W + 36 * V + 24 * F


def grid_sphere(nlat, nlon, seed=0):
    """Closed-form lat-long sphere: nlat x nlon quads -> 2*nlat*nlon - 2*nlon triangles."""
    g = torch.Generator().manual_seed(seed)
    th = torch.linspace(0, math.pi, nlat + 1)[1:-1]
    ph = torch.linspace(0, 2 * math.pi, nlon + 1)[:-1]
    T, P = torch.meshgrid(th, ph, indexing="ij")
    ring = torch.stack([torch.sin(T) * torch.cos(P), torch.cos(T), torch.sin(T) * torch.sin(P)], -1).reshape(-1, 3)
    v = torch.cat([torch.tensor([[0.0, 1.0, 0.0]]), ring, torch.tensor([[0.0, -1.0, 0.0]])])
    v = v * (1 + 0.05 * torch.randn(v.shape[0], 1, generator=g))
    idx = lambda r, s: 1 + r * nlon + (s % nlon)
    r = torch.arange(nlat - 2)[:, None]; s_ = torch.arange(nlon)[None, :]
    a, b, c, d = idx(r, s_), idx(r, s_ + 1), idx(r + 1, s_), idx(r + 1, s_ + 1)
    quads = torch.cat([torch.stack([a, b, c], -1).reshape(-1, 3), torch.stack([b, d, c], -1).reshape(-1, 3)])
    s1 = torch.arange(nlon)
    top = torch.stack([torch.zeros_like(s1), idx(0, s1 + 1), idx(0, s1)], -1)
    bot = torch.stack([torch.full_like(s1, v.shape[0] - 1), idx(nlat - 2, s1), idx(nlat - 2, s1 + 1)], -1)
    return v.float(), torch.cat([top, quads, bot]).long()


def fibonacci_eyes(n, dist=2.7):
    i = torch.arange(n) + 0.5
    phi = torch.acos(1 - 2 * i / n); theta = math.pi * (1 + 5 ** 0.5) * i
    return dist * torch.stack([torch.cos(theta) * torch.sin(phi), torch.cos(phi), torch.sin(theta) * torch.sin(phi)], -1)


def c1(dev):
    """C1: teapot, single view 256^2, SoftPhong + PointLights, forward only (renderer_comparison_with_pyrender.py)."""
    v, f = load_mesh("teapot"); v = normalize_mesh(v)
    mesh = trb.Meshes([v.to(dev)], [f.to(dev)], textures=trb.TexturesVertex(torch.ones(1, v.shape[0], 3, device=dev)))
    R, T = trb.look_at_view_transform(2.7, 10, 20)
    cams = trb.FoVPerspectiveCameras(device=dev, R=R.to(dev), T=T.to(dev))
    rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=256)),
                            trb.SoftPhongShader(device=dev, cameras=cams,
                                                lights=trb.PointLights(device=dev, location=[[0, 0, -3.0]])))
    return (lambda: rend(mesh)), dict(views=1, bytes=bview_forward(1, 256, 256, v.shape[0], f.shape[0]),
                                      what="teapot 256^2 K=1 SoftPhong forward")


def c3(dev, name="teapot", K=50, size=512):
    """C3: single view 512^2, K=50 soft silhouette, pose (T, quaternion) requires grad, fwd+bwd
    (camera_pose_optimizer.py:116-121)."""
    v, f = load_mesh(name); v = normalize_mesh(v)
    mesh = trb.Meshes([v.to(dev)], [f.to(dev)])
    R, T = trb.look_at_view_transform(2.7, 30, 60)
    pose = torch.cat([T, trb.transforms.matrix_to_quaternion(R)], -1).to(dev).requires_grad_(True)
    cams = trb.FoVPerspectiveCameras(device=dev)
    sil = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=size, blur_radius=BLUR,
                                                                              faces_per_pixel=K)),
                           trb.SoftSilhouetteShader(trb.BlendParams(SIGMA, 1e-4, (0, 0, 0))))
    target = torch.rand(1, size, size, device=dev)

    def step():
        pose.grad = None
        Rm = trb.transforms.quaternion_to_matrix(pose[:, 3:]); Tm = pose[:, :3]
        img = sil(mesh, R=Rm, T=Tm)
        (img[..., 3] - target).abs().mean().backward()
    return step, dict(views=1, bytes=bview(K, size, size, v.shape[0], f.shape[0]),
                      what=f"{name} {size}^2 K={K} soft silhouette fwd+bwd to the pose")


def c4(dev, light="point", nv=5, level=6):
    """C4: ico_sphere(level), nv views 512^2 K=1, perspective_correct=False, verts + colours require grad
    (mesh_deformer.py:135-145,181-222)."""
    ico = trb.ico_sphere(level, device=dev)
    v0, f0 = ico.get_mesh_verts_faces(0)
    deform = torch.zeros_like(v0, requires_grad=True)
    rgb = torch.full((1, v0.shape[0], 3), 0.5, device=dev, requires_grad=True)
    R, T = trb.look_at_view_transform(dist=2.7, elev=torch.linspace(0, 360, nv), azim=torch.linspace(-180, 180, nv))
    cams = trb.PerspectiveCameras(device=dev, R=R.to(dev), T=T.to(dev))
    lights = trb.AmbientLights(device=dev) if light == "ambient" else trb.PointLights(device=dev, location=[[0.0, 0.0, 2.0]])
    rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=512, perspective_correct=False)),
                            trb.SoftPhongShader(device=dev, cameras=cams, lights=lights))
    target = torch.rand(nv, 512, 512, 3, device=dev)

    def step():
        deform.grad = None; rgb.grad = None
        m = trb.Meshes([v0 + deform], [f0], textures=trb.TexturesVertex(rgb)).extend(nv)
        img = rend(m)
        ((img[..., :3] - target) ** 2).mean().backward()
    return step, dict(views=nv, bytes=nv * bview(1, 512, 512, v0.shape[0], f0.shape[0]),
                      what=f"ico_sphere({level}) x {nv} views 512^2 K=1 SoftPhong+{light} fwd+bwd to verts/colours")


def c5(dev, nv=4, nlat=501, nlon=1000, size=1024, K=8):
    """C5 (one chunk of it): 1M-face grid sphere, 1024^2, K=8, blur, SoftPhong + PointLights, verts require grad."""
    v, f = grid_sphere(nlat, nlon)
    vd = v.to(dev).requires_grad_(True)
    cols = torch.rand(1, v.shape[0], 3, generator=torch.Generator().manual_seed(0)).to(dev)
    R, T = trb.look_at_view_transform(eye=fibonacci_eyes(nv))
    cams = trb.FoVPerspectiveCameras(device=dev, R=R.to(dev), T=T.to(dev))
    rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=size, blur_radius=BLUR,
                                                                               faces_per_pixel=K)),
                            trb.SoftPhongShader(device=dev, cameras=cams,
                                                lights=trb.PointLights(device=dev, location=[[0, 0, -3.0]])))
    fd = f.to(dev)

    def step():
        vd.grad = None
        m = trb.Meshes([vd], [fd], textures=trb.TexturesVertex(cols)).extend(nv)
        (rend(m) ** 2).mean().backward()
    return step, dict(views=nv, bytes=nv * bview(K, size, size, v.shape[0], f.shape[0]), faces=int(f.shape[0]),
                      verts=int(v.shape[0]),
                      what=f"{f.shape[0]}-face sphere x {nv} views {size}^2 K={K} blur SoftPhong fwd+bwd to verts")


def pose_step(dev, cache):
    """One step of camera_pose_optimizer.Model.forward (:237-254) with the script's ACTIVE settings (:123-128):
    cow, 512^2, K=1, blur 0; rasterizer -> zbuf, silhouette renderer, Phong renderer on the same R, T; pose =
    (T, quaternion) requires grad.  `cache` switches the reuse of Fragments between the three calls."""
    trb.set_fragment_cache(cache)
    v, f = load_mesh("cow")
    cols = torch.rand(1, v.shape[0], 3)
    mesh = trb.Meshes([v.to(dev)], [f.to(dev)], textures=trb.TexturesVertex(cols.to(dev)))
    cams = trb.FoVPerspectiveCameras(device=dev)
    settings = trb.RasterizationSettings(image_size=512, blur_radius=0.0, faces_per_pixel=1)
    blend = trb.BlendParams(SIGMA, 1e-4, (0, 0, 0))
    rast = trb.MeshRasterizer(cams, settings)
    sil = trb.MeshRenderer(trb.MeshRasterizer(cams, settings), trb.SoftSilhouetteShader(blend))
    phong = trb.MeshRenderer(trb.MeshRasterizer(cams, settings),
                             trb.SoftPhongShader(device=dev, cameras=cams, blend_params=blend,
                                                 lights=trb.PointLights(device=dev, location=[[0.0, 0.0, -3.0]])))
    R, T = trb.look_at_view_transform(0.7, 30, 60)
    pose = torch.cat([T, trb.transforms.matrix_to_quaternion(R)], -1).to(dev).requires_grad_(True)
    tgt_d = torch.rand(1, 512, 512, device=dev); tgt_a = torch.rand(1, 512, 512, device=dev)
    tgt_c = torch.rand(1, 512, 512, 3, device=dev)

    def step():
        pose.grad = None
        Rm = trb.transforms.quaternion_to_matrix(pose[:, 3:]); Tm = pose[:, :3]
        depth = torch.relu(rast(meshes_world=mesh, R=Rm, T=Tm).zbuf[..., 0])
        alpha = sil(mesh, R=Rm, T=Tm)[..., 3]
        rgb = phong(mesh, R=Rm, T=Tm)[..., :3]
        ((depth - tgt_d).abs().mean() + (alpha - tgt_a).abs().mean() + ((rgb - tgt_c) ** 2).mean()).backward()
    return step, dict(views=1, bytes=3 * bview(1, 512, 512, v.shape[0], f.shape[0]),
                      what="camera_pose_optimizer step: cow 512^2 K=1, rasterizer + silhouette + Phong renders, fwd+bwd"
                           + (" (Fragments cache on)" if cache else ""))


def clipped(dev, size=256):
    """The near-plane route: teapot with the FoV camera 1.25 units from its centre, so part of the mesh
    lies behind z_clip = znear / 2 = 0.5 -- every render goes through clip_faces (faces cut, stand-alone
    rasteriser, clip_resequence, stand-alone shader), forward + backward to the vertices."""
    v, f = load_mesh("teapot"); v = normalize_mesh(v)
    vd = v.to(dev).requires_grad_(True)
    cols = torch.rand(1, v.shape[0], 3, generator=torch.Generator().manual_seed(0)).to(dev)
    R, T = trb.look_at_view_transform(1.25, 10, 20)
    cams = trb.FoVPerspectiveCameras(device=dev, R=R.to(dev), T=T.to(dev))
    rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=size)),
                            trb.SoftPhongShader(device=dev, cameras=cams,
                                                lights=trb.PointLights(device=dev, location=[[0, 0, -3.0]])))
    fd = f.to(dev)

    def step():
        vd.grad = None
        m = trb.Meshes([vd], [fd], textures=trb.TexturesVertex(cols))
        (rend(m)[..., :3] ** 2).mean().backward()
    return step, dict(views=1, bytes=bview(1, size, size, v.shape[0], f.shape[0]),
                      what=f"teapot {size}^2 K=1 SoftPhong fwd+bwd with vertices behind the near plane (clip_faces route)")


BUILDERS = {
    "clipped": clipped,
    "pose_step": lambda dev: pose_step(dev, False),
    "pose_step_cached": lambda dev: pose_step(dev, True),
    "C1": c1,
    "C3": c3,
    "C3cow": lambda dev: c3(dev, "cow"),
    "C4": c4,
    "C4ambient": lambda dev: c4(dev, "ambient"),
    "C5": c5,
}
