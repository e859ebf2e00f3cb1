"""Timings of the other BASELINE.json configs (C1, C3, C4, C5-like) through the public API on one B200.
Not the bench line (that is C2, bench.py); these show how the same kernels behave off the headline shape."""
import json, math, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch_renderer_b200 as trb
from helpers import load_mesh, normalize_mesh

dev = torch.device("cuda:0")
torch.cuda.set_stream(torch.cuda.Stream(device=dev))


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def bview(K, H, W, V, F):
    return 56 * K * H * W + 32 * H * W + 96 * V + 48 * F


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


out = {}
sigma = 1e-4
blur = math.log(1.0 / 1e-4 - 1.0) * sigma

# C1: teapot, single view 256^2, SoftPhong forward
v, f = load_mesh("teapot"); v = normalize_mesh(v)
mesh = trb.Meshes([v.to(dev)], [f.to(dev)], textures=trb.TexturesVertex(torch.ones(1, v.shape[0], 3, device=dev)))
R, T = trb.look_at_view_transform(2.7, 10, 20)
cams = trb.FoVPerspectiveCameras(device=dev, R=R.to(dev), T=T.to(dev))
rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=256)),
                        trb.SoftPhongShader(device=dev, cameras=cams, lights=trb.PointLights(device=dev, location=[[0, 0, -3.0]])))
out["C1_teapot_256_fwd_ms"] = timeit(lambda: rend(mesh))

# C3: teapot / cow, single view 512^2, K=50 soft silhouette, pose (T, quaternion) requires grad, fwd+bwd
for name in ("teapot", "cow"):
    v, f = load_mesh(name); v = normalize_mesh(v)
    mesh = trb.Meshes([v.to(dev)], [f.to(dev)])
    R, T = trb.look_at_view_transform(2.7, 30, 60)
    pose = torch.cat([T, trb.transforms.matrix_to_quaternion(R)], -1).to(dev).requires_grad_(True)
    cams = trb.FoVPerspectiveCameras(device=dev)
    sil = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=512, blur_radius=blur, faces_per_pixel=50)),
                           trb.SoftSilhouetteShader(trb.BlendParams(sigma, 1e-4, (0, 0, 0))))
    target = torch.rand(1, 512, 512, device=dev)
    def step():
        pose.grad = None
        Rm = trb.transforms.quaternion_to_matrix(pose[:, 3:]); Tm = pose[:, :3]
        img = sil(mesh, R=Rm, T=Tm)
        (img[..., 3] - target).abs().mean().backward()
    ms = timeit(step)
    V, F = v.shape[0], f.shape[0]
    out[f"C3_{name}_512_K50_silhouette_fwdbwd_ms"] = ms
    out[f"C3_{name}_GBps_algorithmic"] = bview(50, 512, 512, V, F) / ms / 1e6

# C4: ico_sphere(6) (V=40,962 F=81,920), 5 views 512^2 K=1, perspective_correct=False, verts + colours require grad
ico = trb.ico_sphere(6, device=dev)
v0, f0 = ico.get_mesh_verts_faces(0)
deform = torch.zeros_like(v0, requires_grad=True)
rgb = torch.full((1, v0.shape[0], 3), 0.5, device=dev, requires_grad=True)
R, T = trb.look_at_view_transform(dist=2.7, elev=torch.linspace(0, 360, 5), azim=torch.linspace(-180, 180, 5))
cams = trb.PerspectiveCameras(device=dev, R=R.to(dev), T=T.to(dev))
for lk, lights in (("ambient", trb.AmbientLights(device=dev)), ("point", trb.PointLights(device=dev, location=[[0.0, 0.0, 2.0]]))):
    rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=512, perspective_correct=False)),
                            trb.SoftPhongShader(device=dev, cameras=cams, lights=lights))
    target = torch.rand(5, 512, 512, 3, device=dev)
    def step():
        deform.grad = None; rgb.grad = None
        m = trb.Meshes([v0 + deform], [f0], textures=trb.TexturesVertex(rgb)).extend(5)
        img = rend(m)
        ((img[..., :3] - target) ** 2).mean().backward()
    ms = timeit(step)
    out[f"C4_ico6_5views_512_{lk}_fwdbwd_ms"] = ms
    out[f"C4_ico6_{lk}_views_per_s"] = 5 / ms * 1e3
    out[f"C4_ico6_{lk}_GBps_algorithmic"] = 5 * bview(1, 512, 512, 40962, 81920) / ms / 1e6

# C5-like: 1M-face grid sphere, 1024^2, K=8, blur, SoftPhong + PointLights, verts require grad; 4 views
v, f = grid_sphere(501, 1000)
out["C5_faces"] = int(f.shape[0]); out["C5_verts"] = int(v.shape[0])
vd = v.to(dev).requires_grad_(True)
cols = torch.rand(1, v.shape[0], 3, device=dev)
NV = 4
i = torch.arange(NV) + 0.5
phi = torch.acos(1 - 2 * i / NV); theta = math.pi * (1 + 5 ** 0.5) * i
eye = 2.7 * torch.stack([torch.cos(theta) * torch.sin(phi), torch.cos(phi), torch.sin(theta) * torch.sin(phi)], -1)
R, T = trb.look_at_view_transform(eye=eye)
cams = trb.FoVPerspectiveCameras(device=dev, R=R.to(dev), T=T.to(dev))
rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=1024, blur_radius=blur, faces_per_pixel=8)),
                        trb.SoftPhongShader(device=dev, cameras=cams, lights=trb.PointLights(device=dev, location=[[0, 0, -3.0]])))
fd = f.to(dev)
def step():
    vd.grad = None
    m = trb.Meshes([vd], [fd], textures=trb.TexturesVertex(cols)).extend(NV)
    (rend(m) ** 2).mean().backward()
ms = timeit(step, n=3, warm=2)
out["C5_1Mfaces_4views_1024_K8_fwdbwd_ms"] = ms
out["C5_views_per_s"] = NV / ms * 1e3
out["C5_GBps_algorithmic"] = NV * bview(8, 1024, 1024, v.shape[0], f.shape[0]) / ms / 1e6
print(json.dumps(out, indent=1))
