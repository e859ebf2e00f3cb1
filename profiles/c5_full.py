"""BASELINE config 5 at full size: the 1M-face sphere x `--views` cameras (default 1024) at 1024^2, K=8, soft blur,
SoftPhong + PointLights, loss = mean(image^2), gradient w.r.t. the vertices.  Views are sharded contiguously over the
ranks (torchrun) and rendered in memory-bounded chunks per rank; the vertex gradient is summed once at the end
(peer-memory kernel below 256 KB, NCCL above: 6 MB here).  Prints one JSON line with whole-job views/s.
    python profiles/c5_full.py [--views 1024] [--chunk 32]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/c5_full.py"""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist
import configs
import torch_renderer_b200 as trb
from torch_renderer_b200.parallel import allreduce_shared_grads, chunk_views, max_views_for_memory, shard_views

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=1024)
ap.add_argument("--chunk", type=int, default=0, help="views per chunk (0: from a 40 GB Fragments budget)")
ap.add_argument("--steps", type=int, default=1)
args = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device(f"cuda:{lr}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
trb.set_fragment_cache(False)
H = W = 1024; K = 8
v, f = configs.grid_sphere(501, 1000)
verts = v.to(dev).requires_grad_(True); faces = f.to(dev)
cols = torch.rand(1, v.shape[0], 3, generator=torch.Generator().manual_seed(0)).to(dev)   # same colours on every rank
R, T = trb.look_at_view_transform(eye=configs.fibonacci_eyes(args.views))
lo, hi = shard_views(args.views, rank, world)
R, T = R[lo:hi].to(dev), T[lo:hi].to(dev)
chunk = args.chunk or min(64, max_views_for_memory(H, W, K, 40 * 10**9))
cams = trb.FoVPerspectiveCameras(device=dev)
rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=H, blur_radius=configs.BLUR, faces_per_pixel=K)),
                        trb.SoftPhongShader(device=dev, cameras=cams, lights=trb.PointLights(device=dev, location=[[0, 0, -3.0]])))

def step():
    verts.grad = None
    total = 0.0
    for s, e in chunk_views(hi - lo, chunk):
        m = trb.Meshes([verts], [faces], textures=trb.TexturesVertex(cols)).extend(e - s)
        img = rend(m, R=R[s:e], T=T[s:e])
        loss = (img ** 2).sum() / (args.views * H * W * 4)
        loss.backward()
        total += float(loss.detach())
    allreduce_shared_grads([verts.grad])
    return total

step()   # warm-up (also grows the pair capacity)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
for _ in range(args.steps):
    loss = step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = (time.perf_counter() - t0) / args.steps
if rank == 0:
    print(json.dumps({"config": "C5 full", "views": args.views, "gpus": world, "chunk": chunk, "seconds_per_step": round(dt, 3),
                      "views_per_s": round(args.views / dt, 1), "algorithmic_GBps_per_gpu": round(args.views / world * configs.bview(K, H, W, v.shape[0], f.shape[0]) / dt / 1e9, 1),
                      "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1), "grad_norm": float(verts.grad.norm())}))
if world > 1: dist.destroy_process_group()
