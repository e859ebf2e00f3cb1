"""Tiny forward+backward renders that touch every fused kernel (K=1 strips + busy tiles, K>1 with depth-ordered
lists, UV textures, silhouette, the backward with run merging, the post kernel, chamfer) -- the target of
    compute-sanitizer --tool memcheck|racecheck python profiles/sanitize_target.py"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch_renderer_b200 as trb
from helpers import cow_uvs, load_mesh, normalize_mesh, uv_sphere
dev = torch.device("cuda:0")
trb.set_fragment_cache(False)
v, f = load_mesh("cow"); v = normalize_mesh(v)
vt, ft = cow_uvs()
R, T = trb.look_at_view_transform(2.7, torch.tensor([10.0, 40.0, -20.0]), torch.tensor([20.0, 200.0, 110.0]))
blur = math.log(1.0 / 1e-4 - 1.0) * 1e-4
for K, br, size, shader, tex in ((1, 0.0, (80, 112), "phong", "vertex"), (1, 0.0, (45, 61), "phong", "uv"),
                                 (6, blur, (64, 64), "phong", "vertex"), (40, blur, (40, 48), "sil", None),
                                 (3, 1e-3, (48, 48), "phong", "uv")):
    vd = v.to(dev).requires_grad_(True)
    Rd, Td = R.to(dev).requires_grad_(True), T.to(dev).requires_grad_(True)
    if tex == "uv":
        texmap = torch.rand(1, 32, 40, 3, device=dev, requires_grad=True)
        textures = trb.TexturesUV(maps=texmap, faces_uvs=[ft.to(dev)], verts_uvs=[vt.to(dev)])
    else:
        textures = trb.TexturesVertex(torch.rand(1, v.shape[0], 3, device=dev, requires_grad=True))
    mesh = trb.Meshes([vd], [f.to(dev)], textures=textures).extend(3)
    cams = trb.FoVPerspectiveCameras(device=dev, R=Rd, T=Td)
    rast = trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=size, blur_radius=br, faces_per_pixel=K))
    sh = (trb.SoftSilhouetteShader(trb.BlendParams(1e-4, 1e-4, (0, 0, 0))) if shader == "sil" else
          trb.SoftPhongShader(device=dev, cameras=cams, lights=trb.PointLights(device=dev, location=[[0.0, 1.0, -3.0]])))
    img = trb.MeshRenderer(rast, sh)(mesh)
    (img ** 2).sum().backward()
    torch.cuda.synchronize()
    print("ok", K, size, shader, tex, float(img.sum()))
sv, sf = uv_sphere(40, 60, noise=0.05)   # dense lists for 16x16 tiles at a small image
m = trb.Meshes([sv.to(dev)], [sf.to(dev)])
fr = trb.MeshRasterizer(trb.FoVPerspectiveCameras(device=dev, R=R[:1].to(dev), T=T[:1].to(dev)),
                        trb.RasterizationSettings(image_size=32, blur_radius=blur, faces_per_pixel=8))(m)
torch.cuda.synchronize()
x = torch.randn(2, 300, 3, device=dev, requires_grad=True); y = torch.randn(2, 500, 3, device=dev)
trb.chamfer_distance(x, y)[0].backward()
torch.cuda.synchronize()
print("done")
