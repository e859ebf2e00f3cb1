"""Runs one BASELINE config (profiles/configs.py) for a few steps: prints wall / device ms per step, the per
C-ABI-call event times and the two fused kernels' own durations.  Also the ncu target for that config:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv python profiles/run_config.py C5 2"""
import json, sys, time
import torch
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import configs
from torch_renderer_b200 import _lib, ops

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
step, info = configs.BUILDERS[name](dev)
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(steps):
    step()
host_ms = (time.perf_counter() - t0) / steps * 1e3
e1.record(); torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / steps
evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for e in evs:
    e.record()
torch.cuda.synchronize()
_lib.lib().trb_debug_set_events(*[e.cuda_event for e in evs])
ops.start_event_log()
fine, bwd = [], []
for _ in range(steps):
    step(); torch.cuda.synchronize()
    fine.append(evs[0].elapsed_time(evs[1]))
    try:
        bwd.append(evs[2].elapsed_time(evs[3]))
    except Exception:
        pass
_lib.lib().trb_debug_set_events(None, None, None, None)
calls = ops.stop_event_log()
out = {"config": name, "ms_per_step_device": round(dev_ms, 4), "ms_per_step_host_issue": round(host_ms, 4),
       "views_per_s": round(info["views"] / dev_ms * 1e3, 2), "algorithmic_GBps": round(info["bytes"] / dev_ms / 1e6, 1),
       "fine_kernel_ms": round(sum(fine) / len(fine), 4), "backward_kernel_ms": round(sum(bwd) / max(len(bwd), 1), 4),
       "calls_ms": {k: round(ms / n, 4) for k, (n, ms) in calls.items()}, **{k: v for k, v in info.items()}}
print(json.dumps(out))
