"""ptxas -v of every kernel at HEAD (registers, spills, barriers, static shared memory), no GPU needed:
python profiles/ptxas_report.py > profiles/r02_ptxas_v.txt"""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "torch_renderer_b200", "csrc")
print("# ptxas -v at HEAD (sm_100a, -O3 -lineinfo): registers, spills, static shared memory per kernel")
for f in ("render", "render_kn", "render_stages", "allreduce", "points_render", "clip", "raster", "shade", "transform", "points"):
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                            "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", os.path.join(ROOT, "include"),
                            "-I", CSRC, "-c", os.path.join(CSRC, f + ".cu"), "-o", os.path.join(tmp, f + ".o")],
                           capture_output=True, text=True)
    print(f"\n## csrc/{f}.cu")
    cur, spill = None, ""
    for ln in r.stderr.splitlines():
        m = re.search(r"Function properties for (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            spill = ""
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m and (int(m.group(2)) or int(m.group(3))):
            spill = f"  [spills: {m.group(2)} B stores, {m.group(3)} B loads]"
        m = re.search(r"Used (\d+) registers.*", ln)
        if m and cur:
            print(f"{cur:78s} {m.group(0)}{spill}")
            cur = None
