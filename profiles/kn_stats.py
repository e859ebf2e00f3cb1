"""Walk statistics of the K > 1 fine kernel (diagnostic build: TRB_EXTRA_NVCC_FLAGS=-DTRB_KN_STATS python -m
torch_renderer_b200.build --force).  python profiles/kn_stats.py C5"""
import ctypes, json, sys
import torch
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
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
out["max_list_entries"] = int(buf[14]); out["max_walk_cycles"] = int(buf[15])
out["views"] = info["views"]
if hasattr(L, "trb_debug_kn_phases"):
    ph = (ctypes.c_ulonglong * 8)()
    L.trb_debug_kn_phases(ph)
    t = max(int(ph[0]), 1)
    out["phase_cycles_per_busy_tile_thread0"] = {"tiles": t, "list ordering": round(int(ph[1]) / t), "staging": round(int(ph[2]) / t),
                                                 "walk": round(int(ph[3]) / t), "epilogue": round(int(ph[4]) / t)}
if hasattr(L, "trb_debug_bw_stats"):
    bw = (ctypes.c_ulonglong * 8)()
    L.trb_debug_bw_stats(bw)
    out["backward_scatter"] = {"warp_layer_rounds": int(bw[0]), "lanes_with_sample": int(bw[1]),
                               "distinct_faces": int(bw[2]), "shuffle_aggregated_rounds": int(bw[3])}
print(json.dumps(out, indent=1))
