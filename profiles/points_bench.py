"""Timing of the point-cloud path on one B200 with the reference's AlphaPointRender settings (torch_renderer.py:163-179:
radius 0.003, points_per_pixel 10): 100k points x 4 views at 512^2, forward + backward to points and features.
  python profiles/points_bench.py > gpurun_out/points_bench.json"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch_renderer_b200 as trb
from torch_renderer_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
P, N, S, K, radius = 100_000, 4, 512, 10, 0.003
pts = torch.nn.functional.normalize(torch.randn(P, 3, device=dev), dim=1) * (1 + 0.05 * torch.randn(P, 1, device=dev))
pts.requires_grad_(True)
feats = torch.rand(P, 3, device=dev, requires_grad=True)
R, T = trb.look_at_view_transform(dist=2.7, elev=torch.linspace(0, 60, N), azim=torch.linspace(-90, 90, N))
R, T = R.to(dev), T.to(dev)
pc = trb.Pointclouds([pts], features=[feats]).extend(N)
renderer = trb.PointsRenderer(
    trb.PointsRasterizer(trb.FoVPerspectiveCameras(device=dev),
                         trb.PointsRasterizationSettings(image_size=S, radius=radius, points_per_pixel=K)),
    trb.AlphaCompositor(background_color=(0, 0, 0)))
grad = torch.randn(N, S, S, 3, device=dev) / (N * S * S)


def step():
    pts.grad = feats.grad = None
    renderer(pc, R=R, T=T).backward(grad)


for _ in range(3):
    step()
torch.cuda.synchronize()
steps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
ops.start_event_log()
for _ in range(steps):
    step()
calls = ops.stop_event_log()
with torch.no_grad():
    frag = renderer.rasterizer(pc, R=R, T=T)
    covered = float((frag.idx[..., 0] >= 0).float().mean())
    filled = float((frag.idx >= 0).float().mean())
print(json.dumps({"workload": f"{P} points x {N} views at {S}^2, radius {radius}, points_per_pixel {K}, AlphaCompositor, fwd+bwd",
                  "ms_per_step": round(ms, 4), "views_per_s": round(N / ms * 1e3, 1),
                  "calls_ms": {k: round(t / n, 4) for k, (n, t) in calls.items()},
                  "covered_pixel_fraction": round(covered, 4), "filled_slot_fraction": round(filled, 4)}))
