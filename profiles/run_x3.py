"""Forward-only render of the cow scaled x3 (18% coverage) -- used as an ncu target."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch_renderer_b200 as trb
from helpers import load_mesh
dev = torch.device("cuda:0"); N, H = 64, 512
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
v, f = load_mesh("cow")
R, T = trb.look_at_view_transform(dist=0.7, elev=torch.linspace(0, 360, N), azim=torch.linspace(-180, 180, N))
mesh = trb.Meshes(verts=[(v * scale).to(dev)], faces=[f.to(dev)], textures=trb.TexturesVertex(torch.rand(1, v.shape[0], 3, device=dev))).extend(N)
cams = trb.FoVPerspectiveCameras(device=dev, R=R.to(dev), T=T.to(dev))
renderer = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=H)),
                            trb.SoftPhongShader(device=dev, cameras=cams, lights=trb.PointLights(device=dev, location=[[0, 0, -3.0]])))
for _ in range(4):
    img = renderer(mesh)
torch.cuda.synchronize()
print("ok", float(img.mean()))
