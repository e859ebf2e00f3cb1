"""What the end-to-end leg's metric (mean alpha of a 64 x 512 x 512 x 4 image batch) costs on the device, by form."""
import json
import torch

dev = torch.device("cuda:0")
img = torch.rand(64, 512, 512, 4, device=dev)
forms = {
    "images[..., 3].mean()": lambda: img[..., 3].mean(),
    "images.view(-1, 4).sum(0)[3] / n": lambda: img.view(-1, 4).sum(0)[3] / (img.numel() // 4),
    "images.view(-1, 4).mean(0)": lambda: img.view(-1, 4).mean(0),
    "images.sum() (all channels, contiguous)": lambda: img.sum(),
    "images.view(-1, 2048).sum(0) (wide column sums)": lambda: img.view(-1, 2048).sum(0),
}
out = {}
big = torch.empty(64 << 20, device=dev)
for name, fn in forms.items():
    for _ in range(5):
        fn()
    ts = []
    for _ in range(20):
        big.zero_()          # flush L2 (256 MB written)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    out[name] = round(ts[len(ts) // 2] * 1e3, 1)
print(json.dumps({"us_per_call_median_cold_L2": out, "bytes": img.numel() * 4}))
