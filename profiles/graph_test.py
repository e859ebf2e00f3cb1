import os, sys, torch, traceback
sys.path[:0] = ["/root/repo", "/root/repo/tests"]
import torch_renderer_b200 as trb
from helpers import load_mesh
dev = torch.device("cuda:0"); N, H = 8, 256
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
v, f = load_mesh("cow")
R, T = trb.look_at_view_transform(dist=0.7, elev=torch.linspace(0, 360, N), azim=torch.linspace(-180, 180, N))
verts = v.to(dev).requires_grad_(True); cols = torch.rand(v.shape[0], 3, device=dev).requires_grad_(True)
Rd = R.to(dev).requires_grad_(True); Td = T.to(dev).requires_grad_(True)
mesh = trb.Meshes(verts=[verts], faces=[f.to(dev)], textures=trb.TexturesVertex(cols[None])).extend(N)
cams = trb.FoVPerspectiveCameras(device=dev)
renderer = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=H)),
                            trb.SoftPhongShader(device=dev, cameras=cams, lights=trb.PointLights(device=dev, location=[[0, 0, -3.0]])))
g_img = torch.randn(N, H, H, 4, device=dev)
params = [verts, cols, Rd, Td]
def step():
    for p in params: p.grad = None
    img = renderer(mesh, R=Rd, T=Td)
    img.backward(g_img)
for _ in range(3): step()
torch.cuda.synchronize()
ref = [p.grad.clone() for p in params]
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        step()
    torch.cuda.synchronize()
    for p in params: p.grad.zero_()
    g.replay(); torch.cuda.synchronize()
    print("graph ok; grads match:", [float((p.grad - r).abs().max()) for p, r in zip(params, ref)])
except Exception:
    traceback.print_exc()
