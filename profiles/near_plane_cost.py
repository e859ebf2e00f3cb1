"""What the default near-plane question (rasterizer.set_near_plane_clipping("exact")) costs per eager step on the
single-view BASELINE configs: the same steps with the question asked ("exact") and not asked ("off"), one process.
  python profiles/near_plane_cost.py > gpurun_out/near_plane_cost.json"""
import json, sys, time
import torch
import configs
import torch_renderer_b200 as trb

dev = torch.device("cuda:0")
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
out = []
for name, steps in (("C1", 200), ("C3", 100), ("C4", 200)):
    step, info = configs.BUILDERS[name](dev)
    row = {"config": name, "steps": steps}
    for rep in range(2):
        for mode in ("exact", "off"):
            trb.set_near_plane_clipping(mode)
            for _ in range(10):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record()
            for _ in range(steps):
                step()
            host_ms = (time.perf_counter() - t0) / steps * 1e3
            e1.record(); torch.cuda.synchronize()
            row[f"{mode}_ms_device_{rep}"] = round(e0.elapsed_time(e1) / steps, 4)
            row[f"{mode}_ms_host_issue_{rep}"] = round(host_ms, 4)
    trb.set_near_plane_clipping("exact")
    out.append(row)
print(json.dumps(out, indent=1))
