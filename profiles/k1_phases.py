"""Where a busy K=1 tile spends its time (diagnostic build: TRB_BUILD_TAG=stats TRB_EXTRA_NVCC_FLAGS=-DTRB_KN_STATS
python -m torch_renderer_b200.build; run with TRB_LIB_PATH=.../libtrb_stats.so).  One forward of the bench workload
(cow x 64 views, 512^2, K=1, SoftPhong): clock64 cycles of thread 0 per phase, averaged over the busy tiles."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch_renderer_b200 as trb
from torch_renderer_b200 import _lib
from bench_workloads import load_mesh
dev = torch.device("cuda:0")
v, f = load_mesh("cow")
N = 64
R, T = trb.look_at_view_transform(dist=0.7, elev=torch.linspace(0, 360, N), azim=torch.linspace(-180, 180, N))
mesh = trb.Meshes([v.to(dev)], [f.to(dev)], textures=trb.TexturesVertex(torch.rand(1, v.shape[0], 3, device=dev))).extend(N)
cams = trb.FoVPerspectiveCameras(device=dev, R=R.to(dev), T=T.to(dev))
rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=512)),
                        trb.SoftPhongShader(device=dev, cameras=cams, lights=trb.PointLights(device=dev, location=[[0, 0, -3.0]])))
trb.set_near_plane_clipping("off")
L = _lib.lib()
buf = (ctypes.c_ulonglong * 16)()
for _ in range(3):
    rend(mesh)
torch.cuda.synchronize()
L.trb_debug_k1_phases(buf)
rend(mesh); torch.cuda.synchronize()
L.trb_debug_k1_phases(buf)
tiles = max(int(buf[0]), 1)
names = ["tiles", "header+staging", "prefix sums", "items", "hit compaction + list atomic", "finish covered pixels"]
out = {"busy_tiles": tiles}
for i in range(1, 6):
    out[names[i] + " (cycles/tile)"] = round(int(buf[i]) / tiles, 1)
out["(face, pixel) items per tile"] = round(int(buf[6]) / tiles, 1)
out["staged faces per tile"] = round(int(buf[7]) / tiles, 1)
out["items: slowest thread (cycles/tile)"] = round(int(buf[8]) / tiles, 1)
out["items: mean thread (cycles/tile)"] = round(int(buf[9]) / tiles, 1)
out["sum (cycles/tile)"] = round(sum(int(buf[i]) for i in range(1, 6)) / tiles, 1)
print(json.dumps(out, indent=1))
