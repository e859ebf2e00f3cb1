"""Walk statistics of the K > 1 fine kernel (diagnostic build: TRB_EXTRA_NVCC_FLAGS=-DTRB_KN_STATS python -m
torch_renderer_b200.build --force).  python profiles/kn_stats.py C5"""
import ctypes, json, sys
import torch
import configs
from torch_renderer_b200 import _lib
name = sys.argv[1]
dev = torch.device("cuda:0")
step, info = configs.BUILDERS[name](dev)
L = _lib.lib()
buf = (ctypes.c_ulonglong * 16)()
step(); torch.cuda.synchronize()
L.trb_debug_kn_stats(buf)
step(); torch.cuda.synchronize()
L.trb_debug_kn_stats(buf)
names = ["busy_tiles", "list_entries", "faces_staged", "walk_iters", "pass_zlo", "pass_bbox", "npos1_shortcut", "npos2_pass",
         "full_evals", "pass_depth", "insertions", "shift_steps", "early_stops", "pixels_not_full"]
out = {n: int(buf[i]) for i, n in enumerate(names)}
out["views"] = info["views"]
print(json.dumps(out, indent=1))
