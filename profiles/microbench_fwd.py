"""GPU micro-measurements behind the forward-kernel design (run with gpurun):
  a) streaming-store floor: torch fill_ of the same bytes the fine kernel writes
  b) fused fine kernel on the C2 scene, c) same with the mesh shrunk so that (almost) every tile is empty."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch_renderer_b200 as trb  # noqa: E402
from torch_renderer_b200 import _lib  # noqa: E402
from helpers import load_mesh  # noqa: E402

dev = torch.device("cuda:0")
N, H, W = 64, 512, 512
out = {}


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


nbytes = N * H * W * 44
buf = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
ms = timeit(lambda: buf.fill_(-1.0))
out["fill_738MB_ms"] = ms
out["fill_GBps"] = nbytes / ms / 1e6
p2f = torch.empty((N, H, W, 1), dtype=torch.int64, device=dev)
zb = torch.empty((N, H, W, 1), device=dev); bary = torch.empty((N, H, W, 1, 3), device=dev)
di = torch.empty((N, H, W, 1), device=dev); img = torch.empty((N, H, W, 4), device=dev)
def five():
    p2f.fill_(-1); zb.fill_(-1); bary.fill_(-1); di.fill_(-1); img.fill_(1.0)
out["five_fills_ms"] = timeit(five)
src = torch.empty_like(buf)
ms = timeit(lambda: buf.copy_(src))
out["copy_738MB_ms"] = ms
out["copy_GBps_rw"] = 2 * nbytes / ms / 1e6

v, f = load_mesh("cow")
R, T = trb.look_at_view_transform(dist=0.7, elev=torch.linspace(0, 360, N), azim=torch.linspace(-180, 180, N))
evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for e in evs:
    e.record()
torch.cuda.synchronize()
_lib.lib().trb_debug_set_events(*[e.cuda_event for e in evs])
for name, scale in (("cow", 1.0), ("cow_x0.001", 1e-3), ("cow_x3", 3.0)):
    verts = (v * scale).to(dev)
    mesh = trb.Meshes(verts=[verts], faces=[f.to(dev)], textures=trb.TexturesVertex(torch.rand(1, v.shape[0], 3, device=dev))).extend(N)
    cams = trb.FoVPerspectiveCameras(device=dev, R=R.to(dev), T=T.to(dev))
    renderer = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=H)),
                                trb.SoftPhongShader(device=dev, cameras=cams, lights=trb.PointLights(device=dev, location=[[0, 0, -3.0]])))
    ts = []
    for _ in range(10):
        images, frag = trb.MeshRendererWithFragments(renderer.rasterizer, renderer.shader)(mesh)
        torch.cuda.synchronize()
        ts.append(evs[0].elapsed_time(evs[1]))
    out[f"fine_ms_{name}"] = sum(ts[3:]) / len(ts[3:])
    out[f"covered_{name}"] = float((frag.pix_to_face >= 0).float().mean())
print(json.dumps(out, indent=1))
