"""Latency of the shared-gradient all-reduce on N GPUs: NCCL vs the peer-memory kernel (run under torchrun)."""
import os, sys, json, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from torch_renderer_b200 import parallel
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device(f"cuda:{lr}")
dist.init_process_group("nccl", device_id=dev)
out = {}
for n in (2930 * 3, 500_002 * 3):
    a, b = torch.randn(n, device=dev), torch.randn(n, device=dev)
    for mode in ("nccl", "peer"):
        if mode == "nccl": os.environ["TRB_NCCL_ALLREDUCE"] = "1"
        else: os.environ.pop("TRB_NCCL_ALLREDUCE", None)
        for _ in range(20): parallel.allreduce_shared_grads([a, b])
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(500): parallel.allreduce_shared_grads([a, b])
        e1.record(); torch.cuda.synchronize()
        out[f"{mode}_{2 * n * 4 // 1024}KB_us"] = round(e0.elapsed_time(e1) / 500 * 1e3, 2)
        a.normal_(); b.normal_()
for v in parallel._peer_allreduce.values():
    if v not in (None, False): v.check()
if rank == 0: print(json.dumps(out))
dist.destroy_process_group()
