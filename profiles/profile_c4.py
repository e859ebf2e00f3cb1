"""C4 step (ico_sphere(6), 5 views, 512^2, K=1, verts+colours grads): eager wall time per step and, with
--events, per C-ABI call time."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch_renderer_b200 as trb
from torch_renderer_b200 import ops
dev = torch.device("cuda:0")
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
NV = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ico = trb.ico_sphere(6, device=dev)
v0, f0 = ico.get_mesh_verts_faces(0)
deform = torch.zeros_like(v0, requires_grad=True)
rgb = torch.full((1, v0.shape[0], 3), 0.5, device=dev, requires_grad=True)
R, T = trb.look_at_view_transform(dist=2.7, elev=torch.linspace(0, 360, NV), azim=torch.linspace(-180, 180, NV))
cams = trb.PerspectiveCameras(device=dev, R=R.to(dev), T=T.to(dev))
rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=512, perspective_correct=False)),
                        trb.SoftPhongShader(device=dev, cameras=cams, lights=trb.PointLights(device=dev, location=[[0.0, 0.0, 2.0]])))
target = torch.rand(NV, 512, 512, 3, device=dev)
def step():
    deform.grad = None; rgb.grad = None
    m = trb.Meshes([v0 + deform], [f0], textures=trb.TexturesVertex(rgb)).extend(NV)
    img = rend(m)
    ((img[..., :3] - target) ** 2).mean().backward()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
print("ms/step wall", (time.perf_counter() - t0) / 20 * 1e3)
t0 = time.perf_counter()
for _ in range(20): step()
print("ms/step host only (no sync)", (time.perf_counter() - t0) / 20 * 1e3)
torch.cuda.synchronize()
ops.start_event_log()
for _ in range(10): step()
log = ops.stop_event_log()
print({k: round(ms / n, 4) for k, (n, ms) in log.items()})
if len(sys.argv) > 2 and sys.argv[2] == "cprofile":
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200): step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(35)
