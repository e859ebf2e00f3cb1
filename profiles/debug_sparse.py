"""Debug aid: dense vs sparse Fragments renders of the same scene, with / without a NaN-poisoned allocator."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import configs
import torch_renderer_b200 as trb
from bench_workloads import load_mesh, normalize_mesh
DEV = torch.device("cuda:0")
v, f = load_mesh("teapot"); v = normalize_mesh(v)
torch.manual_seed(1)
colors = torch.rand(v.shape[0], 3)
N = 3
g = torch.Generator().manual_seed(5)
R0, T0 = trb.look_at_view_transform(dist=2.7, elev=torch.rand(N, generator=g) * 140 - 70, azim=torch.rand(N, generator=g) * 360 - 180)

def render(kind, K, blur, size, dense, poison):
    cams = trb.FoVPerspectiveCameras(device=DEV)
    mesh = trb.Meshes(verts=[v.to(DEV)], faces=[f.to(DEV)], textures=trb.TexturesVertex(colors.to(DEV)[None])).extend(N)
    rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=size, blur_radius=blur, faces_per_pixel=K))
    if kind == "sil":
        shader = trb.SoftSilhouetteShader(trb.BlendParams(1e-4, 1e-4, (0, 0, 0)))
    else:
        shader = trb.SoftPhongShader(device=DEV, cameras=cams, lights=trb.PointLights(device=DEV, location=[[0.0, 0.0, -3.0]]))
    if poison:
        junk = torch.full((64 << 20,), float("nan"), device=DEV); del junk
    cls = trb.MeshRendererWithFragments if dense else trb.MeshRenderer
    out = cls(rast, shader)(mesh, R=R0.to(DEV), T=T0.to(DEV))
    return (out[0] if dense else out).clone()

for kind, K, blur, size in (("phong", 1, 0.0, (64, 64)), ("phong", 8, 9.21024e-4, (128, 128)), ("sil", 50, 9.21024e-4, (96, 96))):
    ref = render(kind, K, blur, size, True, False)
    for dense, poison in ((True, True), (False, False), (False, True), (True, False)):
        img = render(kind, K, blur, size, dense, poison)
        bad = (img != ref) & ~(torch.isnan(img) & torch.isnan(ref))
        idx = bad.nonzero()
        print(kind, K, size, "dense" if dense else "sparse", "poison" if poison else "clean", "differing values:", int(bad.sum()),
              "nan:", int(torch.isnan(img).sum()), idx[:4].tolist(), img[bad][:4].tolist(), ref[bad][:4].tolist())
