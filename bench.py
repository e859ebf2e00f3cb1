#!/usr/bin/env python
"""bench.py -- views/sec of the mesh-rendering hot path (rasterise + SoftPhong, forward + backward).

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): the cow mesh (V=2,930, F=5,856) rendered from
64 cameras at 512x512, faces_per_pixel=1, blur 0, FoV perspective cameras on the
look_at_view_transform(dist=0.7, elev=linspace(0,360,64), azim=linspace(-180,180,64)) orbit,
SoftPhongShader + PointLights; gradients flow to vertices, vertex colours and the per-view R / T.
One "step" = one forward + backward over the 64-view batch (per GPU; weak scaling over GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` goes through the public
API from pinned host buffers (H2D of vertices / colours / R / T and D2H of loss + gradients inside the
timed region); `roofline` is the dominant kernel against the measured HBM peak; `cpu_baseline` is the
CPU oracle (a port: PyTorch3D itself is not installable) timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "views/sec fwd+bwd (raster+SoftPhong) at 512^2"
UNIT = "views/s"
H = W = 512
K = 1
VIEWS_PER_GPU = 64
WORKLOAD = ("cow.obj (V=2930,F=5856) x 64 cameras at 512x512, faces_per_pixel=1, blur 0, SoftPhong+PointLights, "
            "fwd+bwd to verts/colours/R/T (BASELINE configs[1])")
FALLBACK_HBM_GBS = 6650.0


def _peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def _bytes_per_view(V, F):
    """SURVEY 8d / BASELINE.md section 3: B_view = 56*K*H*W + 32*H*W + 96*V + 48*F."""
    return 56 * K * H * W + 32 * H * W + 96 * V + 48 * F


# per-kernel share of B_view (DESIGN.md "kernels and their algorithmic bytes")
def _kernel_bytes(name, V, F):
    hw = H * W
    return {
        # fused fine pass: Fragments (28 B/sample) + RGBA written; NDC verts, faces, world verts, normals,
        # colours read (all L2 resident after first touch)
        "render_fine_kernel": 28 * K * hw + 16 * hw + 12 * V + 12 * F + 36 * V,
        # fused backward: Fragments + image gradient read, mesh re-read, vertex gradients written
        "render_backward_kernel": 28 * K * hw + 16 * hw + 36 * V + 12 * V + 12 * F + 24 * V + 12 * V,
    }.get(name, 0)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()  # exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _scene(device):
    import torch_renderer_b200 as trb
    from bench_workloads import load_mesh      # reads tests/golden/meshes.npz itself; no oracle import on this arm
    v, f = load_mesh("cow")
    torch.manual_seed(0)
    colors = torch.rand(v.shape[0], 3)
    N = VIEWS_PER_GPU
    elev = torch.linspace(0, 360, N)
    azim = torch.linspace(-180, 180, N)
    R, T = trb.look_at_view_transform(dist=0.7, elev=elev, azim=azim)
    return v, f, colors, R, T


def run_ours(args):
    import torch.distributed as dist
    import torch_renderer_b200 as trb
    from torch_renderer_b200 import ops
    from torch_renderer_b200.parallel import allreduce_shared_grads, fused_backward_allreduce

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: torch_renderer_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world == 1 and os.environ.get("TRB_BENCH_INIT_NCCL"):
        # control experiment: a single-GPU run that merely CREATES an NCCL communicator first (profiles/r02_scaling.md)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        dist.all_reduce(torch.ones(8, device=dev))
        torch.cuda.synchronize()
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None) if os.environ.get("NCCL_DEBUG") in ("VERSION", "WARN") else None  # "NCCL version ..." goes to stdout at those levels: keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    # Everything runs on a non-default stream: autograd binds each leaf's gradient accumulator to the
    # stream that was current when the leaf was first used, and work bound to the legacy default stream
    # cannot take part in a CUDA-graph capture.
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    # warm-up, capture and replay deliberately run on different (non-default) streams
    _quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
    if _quiet is not None:
        _quiet(False)

    # every step rasterises: no reuse of Fragments between identical renders (the inputs of the device-resident
    # leg do not change from step to step)
    trb.set_fragment_cache(False)
    v, f, colors, R, T = _scene(dev)
    V, F, N = v.shape[0], f.shape[0], VIEWS_PER_GPU
    # weak scaling: every rank renders its own 64 views of the orbit (rotated per rank), mesh replicated
    if world > 1:
        elev = torch.linspace(0, 360, N) + 360.0 * rank / world / N
        azim = torch.linspace(-180, 180, N) + 360.0 * rank / world / N
        R, T = trb.look_at_view_transform(dist=0.7, elev=elev, azim=azim)

    # the four parameter tensors live in ONE device buffer (leaf views of it), so that the end-to-end leg brings a
    # step's inputs in with a single host->device copy that lands directly in the parameters
    host_params = [t.contiguous().float() for t in (v, colors, R, T)]
    sizes = [t.numel() for t in host_params]
    dev_in = torch.cat([t.reshape(-1) for t in host_params]).to(dev)
    offs = [sum(sizes[:i]) for i in range(len(sizes))]
    verts, cols, Rd, Td = (dev_in[o:o + n].view(t.shape).detach().requires_grad_(True)
                           for o, n, t in zip(offs, sizes, host_params))
    faces = f.to(dev)
    mesh = trb.Meshes(verts=[verts], faces=[faces], textures=trb.TexturesVertex(cols[None]))
    meshes = mesh.extend(N)
    cameras = trb.FoVPerspectiveCameras(device=dev)
    lights = trb.PointLights(device=dev, location=[[0.0, 0.0, -3.0]])
    settings = trb.RasterizationSettings(image_size=H, blur_radius=0.0, faces_per_pixel=K)
    renderer = trb.MeshRenderer(rasterizer=trb.MeshRasterizer(cameras=cameras, raster_settings=settings),
                                shader=trb.SoftPhongShader(device=dev, cameras=cameras, lights=lights))
    grad_img = torch.randn(N, H, W, 4, device=dev) / (N * H * W)
    params = [verts, cols, Rd, Td]
    # Near plane (FoV camera: z_clip = znear / 2 = 0.5, as upstream): whether any vertex lies behind it is a host
    # read and cannot be asked inside a captured step.  The workload's poses are fixed, so it is asked once, here,
    # and the steps run without the question (the eager warm-up and the captured replay then launch the same kernels).
    if ops.any_vertex_behind(verts.detach(), Rd.detach(), Td.detach(), meshes.view_table(), 0.5):
        raise SystemExit("bench workload has vertices behind the near plane: the clipped route is not the benchmark")
    trb.set_near_plane_clipping("off")

    def core_device():          # zero grads + forward + backward: the graph-captured part
        for p in params:
            p.grad = None
        images = renderer(meshes, R=Rd, T=Td)
        images.backward(grad_img)
        return images

    def step_separate():        # the all-reduce as its own kernel after the backward (round-2 baseline, kept for the A/B)
        core_device()
        allreduce_shared_grads([verts.grad, cols.grad])

    def step_fused():           # the all-reduce's push half inside the backward's tail kernel + a receive kernel
        for p in params:
            p.grad = None
        images = renderer(meshes, R=Rd, T=Td)
        with fused_backward_allreduce(dev) as fused:
            images.backward(grad_img)
        if not fused:           # no peer memory: NCCL after the backward
            allreduce_shared_grads([verts.grad, cols.grad])
        return images

    step_device = step_fused if (world > 1 and args.collective == "fused") else step_separate
    e2e_fused = world > 1 and args.collective == "fused"

    def _fused_available():     # a live PeerAllReduce exists once the first multi-GPU step has run
        from torch_renderer_b200 import parallel as _p
        return any(v not in (None, False) for v in _p._peer_allreduce.values())

    # pinned host buffers for the end-to-end leg: ONE packed buffer each way (inputs in; gradients + metric out)
    host_in = torch.cat([t.reshape(-1) for t in host_params]).pin_memory()
    host_out = torch.empty(sum(sizes) + 1, dtype=torch.float32).pin_memory()
    dev_out = torch.empty(sum(sizes) + 1, dtype=torch.float32, device=dev)

    def core_e2e():             # H2D of this step's inputs + forward + backward (+ fused all-reduce) + metric
        for p in params:
            p.grad = None
        dev_in.copy_(host_in, non_blocking=True)   # lands in verts / cols / Rd / Td (views of dev_in)
        images = renderer(meshes, R=Rd, T=Td)
        if e2e_fused:
            with fused_backward_allreduce(dev):
                images.backward(grad_img)
        else:
            images.backward(grad_img)
        # the step's result read back by the host: mean alpha (silhouette coverage), from the per-view sums the fine
        # kernel accumulates while it writes the pixels (images.alpha_sum; a second pass over the 268 MB image batch
        # -- images[..., 3].mean() -- costs 63 us on its own, profiles/r02_mean_alpha_cost.json) ...
        # (one kernel: the sum lands in the packed output buffer; the host divides by the pixel count)
        torch.sum(images.alpha_sum.detach(), dim=0, keepdim=True, out=dev_out[-1:])

    def readback_e2e():         # ... and every gradient (after the all-reduce of the shared ones)
        torch.cat([p.grad.reshape(-1) for p in params], out=dev_out[:-1])
        host_out.copy_(dev_out, non_blocking=True)

    def graphed(fn):
        """Captures one step (forward + backward [+ copies]) into a CUDA graph; eager on failure."""
        if args.no_graph:
            return fn, "eager"
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            torch.cuda.synchronize()
            return g.replay, "cuda-graph"
        except Exception as e:  # noqa: BLE001
            torch.cuda.synchronize()
            return fn, f"eager (graph capture failed: {type(e).__name__})"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, repeats=None):
        """Median over `repeats` repeats of: EXACTLY `steps` steps between CUDA events on the launching stream,
        bracketed by barrier + synchronize on both sides, max over ranks.  Returns (median ms, all repeats)."""
        out = []
        for _ in range(repeats or args.repeats):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            out.append(ms)
        return statistics.median(out), out

    for _ in range(args.warmup):
        step_device()
    l0 = ops.launch_count()
    step_device()
    launches_per_step = ops.launch_count() - l0
    # world > 1: the all-reduce of the shared gradients is part of the captured step (the peer-memory kernel keeps
    # its epochs in device memory, so it replays); its block 0 accumulates push / wait nanoseconds (diagnostics)
    ar_timing = None
    fused_in_use = False
    e2e_fused = e2e_fused and _fused_available()
    if world > 1:
        ar_timing = torch.zeros(3, dtype=torch.int64, device=dev)
        from torch_renderer_b200 import _lib as _l
        _l.lib().trb_allreduce_set_timing(ar_timing.data_ptr())
        step_run, mode_device = graphed(step_device)
        if mode_device != "cuda-graph":
            core_run, mode_device = graphed(core_device)

            def step_run():
                core_run()
                allreduce_shared_grads([verts.grad, cols.grad])
            mode_device += " + eager all-reduce"
            fused_in_use = False
        else:
            fused_in_use = step_device is step_fused and _fused_available()
            mode_device = "cuda-graph (all-reduce captured" + (", push fused into the backward tail)" if fused_in_use else ")")
        run_device = step_run
    else:
        run_device, mode_device = graphed(core_device)

    for _ in range(args.warmup):
        run_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # keep the GPU under the same load until nvidia-smi has a few samples (its period is 100 ms);
    # a fixed step count so that every rank issues the same number of all-reduces
    for _ in range(1500):
        run_device()
    if ar_timing is not None:
        torch.cuda.synchronize()
        ar_timing.zero_()
    # diagnostic builds only (libtrb built with -DTRB_STEP_STAMPS, selected through TRB_LIB_PATH): %globaltimer stamps
    # of the step's first kernel, of the backward's tail kernel and of the all-reduce kernel, per step, in a ring
    stamps = None
    if os.environ.get("TRB_STEP_STAMPS_OUT"):
        import ctypes
        from torch_renderer_b200 import _lib as _l2
        h = ctypes.CDLL(_l2.LIB_PATH)
        if hasattr(h, "trb_debug_set_step_stamps"):
            stamps = (torch.zeros(256 * 8, dtype=torch.int64, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
            torch.cuda.synchronize()
            h.trb_debug_set_step_stamps(ctypes.c_void_p(stamps[0].data_ptr()), ctypes.c_void_p(stamps[1].data_ptr()))
    ms_total, ms_repeats = timed(run_device, args.steps)
    if stamps is not None:
        torch.cuda.synchronize()
        h.trb_debug_set_step_stamps(None, None)
        ring = stamps[0].cpu().view(256, 8).numpy().astype(np.int64)
        last = int(stamps[1].item())
        rows = [(k & 255) for k in range(max(last - 200, 1), last)]      # the newest 200 steps, oldest first
        seg = {"fwd_bwd_until_tail_starts": [], "tail_kernel": [], "gap_tail_end_to_allreduce_start": [],
               "allreduce_kernel": [], "gap_last_kernel_end_to_next_step": [], "step": []}
        for a_, b_ in zip(rows[:-1], rows[1:]):
            r0, r1 = ring[a_], ring[b_]
            if r1[0] <= r0[0] or r1[0] - r0[0] > 2_000_000:   # a repeat boundary (barrier) between the two steps
                continue
            has_ar = r0[4] > 0
            seg["fwd_bwd_until_tail_starts"].append(r0[1] - r0[0])
            seg["tail_kernel"].append(r0[2] - r0[1])
            if has_ar:
                seg["gap_tail_end_to_allreduce_start"].append(r0[3] - r0[2])
                seg["allreduce_kernel"].append(r0[4] - r0[3])
            seg["gap_last_kernel_end_to_next_step"].append(r1[0] - (r0[4] if has_ar else r0[2]))
            seg["step"].append(r1[0] - r0[0])
        rep = {k: round(float(np.median(v)) / 1e3, 2) for k, v in seg.items() if v}
        rep["unit"] = "us, median over %d consecutive steps of the timed region" % len(seg["step"])
        with open(f"{os.environ['TRB_STEP_STAMPS_OUT']}_n{world}_rank{rank}.json", "w") as fh:
            json.dump(rep, fh)
    ar_report = None
    if ar_timing is not None:
        t = ar_timing.clone()
        per_rank = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(per_rank, t)
        calls = max(int(t[2].item()), 1)
        ar_report = {"what": ("block 0 of the receive kernel (the push runs inside post_backward_kernel's last blocks: push_us is 0 here)"
                              if fused_in_use else "block 0 of the peer all-reduce kernel") +
                             ", %globaltimer ns averaged over the timed steps, per rank",
                     "push_us": [round(float(r[0]) / max(int(r[2]), 1) / 1e3, 2) for r in per_rank],
                     "wait_and_sum_us": [round(float(r[1]) / max(int(r[2]), 1) / 1e3, 2) for r in per_rank],
                     "calls": calls,
                     "reading": "wait_and_sum = rank skew (the slowest rank's backward) + one NVLink one-way latency; "
                                "push = issuing the remote stores"}
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    contexts = None
    if world > 1 and rank == 0:
        # which processes hold a CUDA context on which GPU (a rank that also opened a context on a neighbour's GPU
        # would make that GPU switch between two contexts)
        try:
            out = subprocess.run(["nvidia-smi", "--query-compute-apps=pid,gpu_bus_id,used_memory", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=10).stdout
            per_gpu = {}
            for ln in out.strip().splitlines():
                parts = [x.strip() for x in ln.split(",")]
                if len(parts) >= 2:
                    per_gpu.setdefault(parts[1], []).append(parts[0])
            contexts = {"processes_per_gpu": {k: len(v) for k, v in per_gpu.items()}}
        except Exception:  # noqa: BLE001
            contexts = None
    ms_step = ms_total / args.steps
    value = world * N / (ms_step / 1e3)
    # same box, same process: the step with the all-reduce as its own kernel after the backward (the other variant)
    collective_ab = None
    if world > 1 and fused_in_use:
        sep_run, sep_mode = graphed(step_separate)
        if sep_mode == "cuda-graph":
            for _ in range(200):
                sep_run()
            ms_sep = timed(sep_run, args.steps)[0] / args.steps
            for _ in range(200):
                run_device()
            ms_fused_again = timed(run_device, args.steps)[0] / args.steps
            collective_ab = {"fused_into_backward_tail_ms_per_step": round(ms_fused_again, 4),
                             "separate_kernel_ms_per_step": round(ms_sep, 4),
                             "what": "same process, both captured in the step graph, timed back to back after the headline"}

    # Where the N-GPU step's extra time goes: every rank times its OWN step without the all-reduce (same views, same
    # graph minus the collective, all ranks running at once, no synchronisation between them).  A step with the
    # all-reduce ends on the slowest rank, so max - mean of these is the rank skew the collective exposes; what is
    # left of (N-GPU step - slowest rank's own step) is the all-reduce itself (launch + push + NVLink + sum).
    skew_report = None
    if world > 1:
        core_run, core_mode = graphed(core_device)
        if core_mode == "cuda-graph":
            for _ in range(200):
                core_run()
            own = []
            for _ in range(args.repeats):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    core_run()
                e1.record()
                torch.cuda.synchronize()
                own.append(e0.elapsed_time(e1) / args.steps)
            mine = torch.tensor([statistics.median(own)], device=dev)
            every = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
            per_rank = [round(float(t.item()), 4) for t in every]
            skew_report = {"own_step_without_allreduce_ms_per_rank": per_rank,
                           "slowest_minus_mean_us": round((max(per_rank) - sum(per_rank) / world) * 1e3, 2),
                           "step_with_allreduce_minus_slowest_own_us": round((ms_step - max(per_rank)) * 1e3, 2)}

    # per-kernel durations: the same steps again, with CUDA events recorded by libtrb on the launching
    # stream right around its two dominant kernels (the fused fine pass and the fused backward), and
    # around every C-ABI call (which also covers the small kernels of each call)
    from torch_renderer_b200 import _lib
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for e in evs:
        e.record()  # materialise the handles
    torch.cuda.synchronize()
    _lib.lib().trb_debug_set_events(*[e.cuda_event for e in evs])
    fine_ms, bwd_ms = [], []
    ops.start_event_log()
    for _ in range(args.steps):
        step_device()
        torch.cuda.synchronize()
        fine_ms.append(evs[0].elapsed_time(evs[1]))
        bwd_ms.append(evs[2].elapsed_time(evs[3]))
    _lib.lib().trb_debug_set_events(None, None, None, None)
    calls = ops.stop_event_log()
    per_launch_ms = {"render_fine_kernel": statistics.mean(fine_ms), "render_backward_kernel": statistics.mean(bwd_ms)}
    call_ms = {k: ms / n for k, (n, ms) in calls.items()}
    dominant = max(per_launch_ms, key=per_launch_ms.get)
    peak, peak_src = _peak()
    dom_bytes = _kernel_bytes(dominant, V, F) * N
    achieved = dom_bytes / (per_launch_ms[dominant] * 1e-3) / 1e9
    step_bytes = _bytes_per_view(V, F) * N
    step_achieved = step_bytes / (ms_step * 1e-3) / 1e9
    traffic = TRAFFIC_BYTES_PER_LAUNCH.get(dominant)
    # Bytes the kernels really have to move now that the image-only renderer writes Fragments for COVERED pixels
    # only (trb_render_config.sparse_fragments): the RGBA image in full, 28 B + a 4-byte list entry per covered
    # pixel, the mesh once.  `frac` above keeps SURVEY 8d's model (dense Fragments) as the contract asks -- it can
    # exceed 1 for that reason -- and `moved` says what fraction of the HBM peak the kernel reaches on the bytes it
    # does move.
    with torch.no_grad():
        covered = int((renderer(meshes, R=Rd, T=Td)[..., 3] > 0).sum().item())
    moved_bytes = {
        "render_fine_kernel": 16 * H * W * N + 32 * covered + N * (12 * V + 12 * F) + 36 * V,
        "render_backward_kernel": (28 + 4 + 16) * covered + N * (12 * V + 12 * F) + 36 * V + 16 * (N * V + 3 * V),
    }[dominant]
    moved = {"bytes_per_launch": int(moved_bytes), "covered_pixels": covered,
             "covered_fraction": round(covered / (N * H * W), 4),
             "GBps": round(moved_bytes / (per_launch_ms[dominant] * 1e-3) / 1e9, 1),
             "frac": round(moved_bytes / (per_launch_ms[dominant] * 1e-3) / 1e9 / peak, 4),
             "what": "image in full + Fragments of covered pixels only (sparse) + mesh; the kernel is bound by the "
                     "latency of rasterising its busy tiles, not by these bytes"}

    # end-to-end through the public API with host buffers
    core_e2e_run, mode_e2e = graphed(core_e2e)
    if world == 1:
        def whole_e2e():
            core_e2e()
            readback_e2e()
        whole_run, mode_e2e = graphed(whole_e2e)
        run_e2e = whole_run
    else:
        readback_run, _ = graphed(readback_e2e)

        def run_e2e():
            core_e2e_run()
            if not e2e_fused:
                allreduce_shared_grads([verts.grad, cols.grad])
            readback_run()

    for _ in range(max(3, args.warmup)):
        run_e2e()
    ms_e2e = timed(run_e2e, args.steps)[0] / args.steps
    e2e_value = world * N / (ms_e2e / 1e3)
    # the metric the host read back (fused per-view alpha sums) against a second pass over the image batch
    torch.cuda.synchronize()
    e2e_metric = float(host_out[-1].item()) / float(N * H * W)
    with torch.no_grad():
        metric_ref = float(renderer(meshes, R=Rd, T=Td)[..., 3].double().mean().item())
    if not abs(e2e_metric - metric_ref) <= 1e-4 * abs(metric_ref) + 1e-7:
        raise SystemExit(f"bench: the fused mean-alpha metric {e2e_metric} disagrees with images[..., 3].mean() = {metric_ref}")
    h2d = host_in.numel() * 4
    d2h = host_out.numel() * 4

    from torch_renderer_b200 import parallel as _par
    peer = [v for v in _par._peer_allreduce.values() if v not in (None, False)]
    for v in peer:
        v.check()   # no peer's contribution timed out
    collective = "none" if world == 1 else ("peer-memory one-shot kernel (csrc/allreduce.cu)" if peer else "nccl all-reduce")

    # Is the reduced gradient RIGHT?  One more step: the local (pre-reduction) gradients are kept, reduced by the
    # product path, and compared with (a) the rank-ordered sum of an all_gather of the local copies -- bit-exact,
    # the peer kernel sums in rank order -- and (b) NCCL's own all-reduce of the copies (its order is its own:
    # tolerance).  Every rank checks its own result; the verdicts are AND-ed.
    collective_check = None
    fused_check = None
    if world > 1 and fused_in_use:
        # the fused form sums in place inside the backward, so its inputs are not observable: the result must be
        # bit-identical on every rank (same values, same rank order) and agree with NCCL's all-reduce of the local
        # gradients of a separate, un-reduced backward (whose atomics settle in another order: tolerance)
        step_fused()
        got = [verts.grad.detach().clone(), cols.grad.detach().clone()]
        core_device()
        same = close = True
        for g, loc in zip(got, (verts.grad, cols.grad)):
            every = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(every, g)
            same = same and all(bool(torch.equal(every[0], e)) for e in every[1:])
            nccl = loc.detach().clone()
            dist.all_reduce(nccl)
            close = close and bool(torch.allclose(g, nccl, rtol=1e-4, atol=1e-5 * float(nccl.abs().max())))
        flags = torch.tensor([int(same), int(close)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        same, close = bool(flags[0].item()), bool(flags[1].item())
        fused_check = ("ok" if (same and close) else "MISMATCH") + \
            f" (bit-identical on all {world} ranks: {same}; vs NCCL all-reduce of a separate backward within 1e-4: {close})"
    if world > 1:
        core_device()
        local = [verts.grad.detach().clone(), cols.grad.detach().clone()]
        allreduce_shared_grads([verts.grad, cols.grad])
        ok_exact = ok_close = True
        for mine, loc in zip((verts.grad, cols.grad), local):
            gathered = [torch.empty_like(loc) for _ in range(world)]
            dist.all_gather(gathered, loc)
            want = gathered[0].clone()
            for g in gathered[1:]:
                want += g
            ok_exact = ok_exact and bool(torch.equal(mine, want))
            nccl = loc.clone()
            dist.all_reduce(nccl)
            ok_close = ok_close and bool(torch.allclose(mine, nccl, rtol=1e-4, atol=1e-6 * float(nccl.abs().max())))
        flags = torch.tensor([int(ok_exact), int(ok_close)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        ok_exact, ok_close = bool(flags[0].item()), bool(flags[1].item())
        collective_check = ("ok" if (ok_close and (ok_exact or not peer)) else "MISMATCH") + \
            f" (rank-ordered sum bit-exact: {ok_exact}; vs NCCL all-reduce within 1e-4: {ok_close}; all {world} ranks)"

    # the same step launched EAGERLY with the default near-plane handling ("exact": every render asks the device
    # whether a vertex lies behind z_clip and reads the answer after enqueueing) -- what a reference script that
    # only switched its imports gets without graph capture
    trb.set_near_plane_clipping("exact")
    for _ in range(args.warmup):
        step_device()
    ms_eager = timed(step_device, args.steps)[0] / args.steps
    trb.set_near_plane_clipping("off")
    eager = {"value": round(world * N / (ms_eager / 1e3), 2), "unit": UNIT, "ms_per_step": round(ms_eager, 4),
             "launch_mode": "eager", "near_plane": "exact (asked every step)"}

    # every other BASELINE config (C1, C3, C4, the reference's pose step, one chunk of C5) on this GPU, and C5 as
    # specified (1024 views in total, strong-scaled over the ranks)
    other = None
    if world == 1 and not args.no_configs:
        other = measure_other_configs(dev, args, peak)
    c5 = None
    if not args.no_c5:
        c5 = run_c5_spec(dev, world, rank, args, peak, barrier)

    cpu = cpu_baseline(sample_views=args.cpu_views) if (rank == 0 and world == 1 and not args.no_cpu) else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic cameras on the reference cow mesh geometry "
            "(tests/golden/meshes.npz), random vertex colours, seed 0",
            "config": {"workload": WORKLOAD, "views_per_gpu": N, "image": [H, W], "faces_per_pixel": K,
                       "parallelism": f"view-sharded x{world}, mesh replicated, 1 fused allreduce of shared grads",
                       "collective": collective,
                       "l2_policy": "inputs larger than L2: Fragments + images + their grads = 1.2 GB per step vs 126 MB L2",
                       "near_plane": "z_clip 0.5 checked once before the timed region (no vertex behind it); not re-asked per step",
                       "launch_mode": mode_device, "e2e_launch_mode": mode_e2e},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(ms_e2e, 4),
                    "result_read_back": {"what": "every gradient + mean alpha of the image batch; the mean comes from the "
                                                 "per-view alpha sums the fine kernel accumulates (images.alpha_sum), "
                                                 "checked here against a second pass over the images",
                                         "mean_alpha": e2e_metric, "images_mean_alpha": metric_ref}},
            "gpu_launches": launches,
            "timing": {"protocol": f"median of {args.repeats} repeats of {args.steps} steps, CUDA events on the launching "
                                   "stream, barrier + synchronize around every repeat, max over ranks",
                       "ms_per_step_repeats": [round(m / args.steps, 4) for m in ms_repeats]},
            "eager_exact": eager,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": round(achieved, 1), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                         "peak_source": peak_src, "kernel_ms": round(per_launch_ms[dominant], 4),
                         "algorithmic_bytes_per_launch": dom_bytes, "moved": moved,
                         "frac_note": "frac = SURVEY 8d's algorithmic bytes (dense Fragments) / kernel time / peak; the "
                                      "image-only renderer no longer writes background Fragments, so frac can exceed 1 "
                                      "-- see `moved` and `traffic` (ncu DRAM bytes of one launch)",
                         "step_achieved": round(step_achieved, 1), "step_frac": round(step_achieved / peak, 4),
                         "step_note": "step_* uses SURVEY 8d's byte model (Fragments re-read in full by the backward); "
                                      "the backward really re-reads only the covered pixels, so step_frac can exceed 1",
                         "kernels_ms_per_launch": {k: round(x, 4) for k, x in sorted(per_launch_ms.items())},
                         "calls_ms": {k: round(x, 4) for k, x in sorted(call_ms.items())}},
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if collective_check is not None:
            line["collective_check"] = collective_check
        if fused_check is not None:
            line["collective_check_fused"] = fused_check
        if collective_ab is not None:
            line["collective_ab"] = collective_ab
        if ar_report is not None:
            if skew_report is not None:
                ar_report["rank_skew"] = skew_report
            if contexts is not None:
                ar_report["cuda_contexts"] = contexts
            line["collective_timing"] = ar_report
        if other is not None:
            line["other_configs"] = other
        if c5 is not None:
            line["c5"] = c5
        print(json.dumps(line), flush=True)
    if world > 1:
        # The result is out: teardown must not outlive it.  (With the NCCL fallback captured in the step graph,
        # destroy_process_group has been seen to stall for minutes; the peer-memory route exits cleanly.)
        sys.stdout.flush()
        sys.stderr.flush()
        bail = threading.Timer(20.0, lambda: os._exit(0))
        bail.daemon = True
        bail.start()
        try:
            torch.cuda.synchronize()
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass
        os._exit(0)


# Warp instructions per launch of the K>1 fine kernel on the C5 chunk / C3, from the committed ncu captures
# (profiles/r02_ncu_kn_C5_after.txt, r02_ncu_kn_C3_after.txt: smsp__inst_executed.sum; the fine kernels' instruction
# streams did not change afterwards).  C5 is bound by instruction issue, not by HBM, so its fine kernel is also
# reported against the issue-slot roofline: warp instructions / (148 SMs x 4 schedulers x SM clock x kernel time).
KN_FINE_WARP_INSTRUCTIONS = {"C5": 3_241_201_101, "C3": 55_032_257}
KN_FINE_LANES_PER_INSTRUCTION = {"C5": 19.44, "C3": 23.20}


def measure_other_configs(dev, args, peak):
    """C1, C3 (teapot, cow), C4, the reference's camera_pose_optimizer step and one 4-view chunk of C5 through the
    public API, eager: ms/step = median of `repeats` repeats of `steps` steps (CUDA events), the two fused kernels'
    own durations (events recorded by libtrb around them), per C-ABI-call times, algorithmic GB/s (SURVEY 8d byte
    model) and its fraction of the measured HBM peak."""
    import torch_renderer_b200 as trb
    import bench_workloads as wl
    from torch_renderer_b200 import _lib, ops
    out = {}
    plan = (("C1", 20), ("C3", 20), ("C3cow", 20), ("C4", 20), ("pose_step", 20), ("pose_step_cached", 20),
            ("clipped", 20), ("C5", 3))
    for name, steps in plan:
        trb.set_fragment_cache(False)
        trb.set_near_plane_clipping("exact")     # the defaults a drop-in script gets
        try:
            step, info = wl.BUILDERS[name](dev)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            reps, host = [], []
            for _ in range(args.repeats):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record()
                for _ in range(steps):
                    step()
                host.append((time.perf_counter() - t0) / steps * 1e3)
                e1.record()
                torch.cuda.synchronize()
                reps.append(e0.elapsed_time(e1) / steps)
            ms = statistics.median(reps)
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            for e in evs:
                e.record()
            torch.cuda.synchronize()
            _lib.lib().trb_debug_set_events(*[e.cuda_event for e in evs])
            ops.start_event_log()
            fine, bwd = [], []
            for _ in range(min(steps, 10)):
                step()
                torch.cuda.synchronize()
                if name != "clipped":      # the clipped route runs the stand-alone kernels (see calls_ms)
                    fine.append(evs[0].elapsed_time(evs[1]))
                if name not in ("C1", "clipped"):
                    bwd.append(evs[2].elapsed_time(evs[3]))
            _lib.lib().trb_debug_set_events(None, None, None, None)
            calls = ops.stop_event_log()
            gbs = info["bytes"] / ms / 1e6
            # the same step replayed from ONE CUDA graph (trb.capture_step: the near-plane flag is checked
            # asynchronously between replays instead of being waited for)
            graph = None
            if name not in ("C5", "clipped") and not args.no_graph:   # clipped: cutting faces needs the host's answer
                try:
                    cap = trb.capture_step(step)
                    for _ in range(3):
                        cap()
                    greps = []
                    for _ in range(args.repeats):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(steps):
                            cap()
                        e1.record()
                        torch.cuda.synchronize()
                        greps.append(e0.elapsed_time(e1) / steps)
                    cap.check()
                    gms = statistics.median(greps)
                    graph = {"ms_per_step": round(gms, 4), "views_per_s": round(info["views"] / gms * 1e3, 2),
                             "hbm_frac": round(info["bytes"] / gms / 1e6 / peak, 4),
                             "launch_mode": "cuda-graph (trb.capture_step), near-plane flag checked asynchronously"}
                    del cap
                except Exception as e:  # noqa: BLE001
                    graph = {"error": f"{type(e).__name__}: {e}"[:300]}
            issue = None
            if name in KN_FINE_WARP_INSTRUCTIONS and fine:
                slots = 148 * 4 * 1.965e9 * statistics.median(fine) * 1e-3     # issue slots at the max SM clock
                issue = {"kernel": "render_fine_kn_kernel", "warp_instructions_per_launch": KN_FINE_WARP_INSTRUCTIONS[name],
                         "issue_slot_frac": round(KN_FINE_WARP_INSTRUCTIONS[name] / slots, 4),
                         "lanes_per_instruction": KN_FINE_LANES_PER_INSTRUCTION[name],
                         "thread_slot_frac": round(KN_FINE_WARP_INSTRUCTIONS[name] * KN_FINE_LANES_PER_INSTRUCTION[name]
                                                   / 32 / slots, 4),
                         "source": "ncu smsp__inst_executed.sum (profiles/r02_ncu_kn_*_after.txt) / (148 SMs x 4 "
                                   "schedulers x 1965 MHz x this run's kernel time)"}
            out[name] = {"what": info["what"], "views_per_s": round(info["views"] / ms * 1e3, 2), "captured": graph,
                         "issue_roofline": issue,
                         "ms_per_step": round(ms, 4), "ms_per_step_repeats": [round(r, 4) for r in reps],
                         "host_issue_ms_per_step": round(statistics.median(host), 4), "steps": steps,
                         "fine_kernel_ms": round(statistics.median(fine), 4) if fine else None,
                         "backward_kernel_ms": round(statistics.median(bwd), 4) if bwd else None,
                         "calls_ms": {k: round(t / n, 4) for k, (n, t) in sorted(calls.items())},
                         "algorithmic_bytes_per_step": int(info["bytes"]), "algorithmic_GBps": round(gbs, 1),
                         "hbm_frac": round(gbs / peak, 4), "launch_mode": "eager", "near_plane": "exact"}
        except Exception as e:  # noqa: BLE001 -- one failing config must not lose the headline line
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        finally:
            step = None
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
    trb.set_fragment_cache(False)
    trb.set_near_plane_clipping("off")
    return out


def run_c5_spec(dev, world, rank, args, peak, barrier):
    """BASELINE configs[4] as written: the 1M-face sphere x 1024 views at 1024^2, K=8, blur, SoftPhong, loss =
    mean(image^2), gradient to the vertices.  STRONG scaling: the 1024 views are sharded contiguously over the
    ranks (`parallel.shard_views`), each rank renders its slice in memory-bounded chunks (`parallel.chunk_views`),
    and the 6 MB vertex gradient is summed once per pass (`allreduce_shared_grads`: above 256 KB that is the fused
    NCCL all-reduce).  One warm-up pass, then `--c5-passes` timed passes (CUDA events, max over ranks)."""
    import torch.distributed as dist
    import torch_renderer_b200 as trb
    import bench_workloads as wl
    from torch_renderer_b200.parallel import allreduce_shared_grads, chunk_views, shard_views
    views = args.c5_views
    Hc = Wc = 1024
    Kc = 8
    trb.set_fragment_cache(False)
    trb.set_near_plane_clipping("exact")
    v, f = wl.grid_sphere(501, 1000)
    verts = v.to(dev).requires_grad_(True)
    faces = f.to(dev)
    cols = torch.rand(1, v.shape[0], 3, generator=torch.Generator().manual_seed(0)).to(dev)
    Rall, Tall = trb.look_at_view_transform(eye=wl.fibonacci_eyes(views))
    lo, hi = shard_views(views, rank, world)
    R, T = Rall[lo:hi].to(dev), Tall[lo:hi].to(dev)
    chunk = args.c5_chunk
    cams = trb.FoVPerspectiveCameras(device=dev)
    rend = trb.MeshRenderer(
        trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=Hc, blur_radius=wl.BLUR, faces_per_pixel=Kc)),
        trb.SoftPhongShader(device=dev, cameras=cams, lights=trb.PointLights(device=dev, location=[[0, 0, -3.0]])))
    loss_acc = torch.zeros((), device=dev)

    def one_pass():
        verts.grad = None
        loss_acc.zero_()
        for s0, s1 in chunk_views(hi - lo, chunk):
            m = trb.Meshes([verts], [faces], textures=trb.TexturesVertex(cols)).extend(s1 - s0)
            img = rend(m, R=R[s0:s1], T=T[s0:s1])
            loss = (img ** 2).sum() / (views * Hc * Wc * 4)
            loss.backward()
            loss_acc.add_(loss.detach())
        allreduce_shared_grads([verts.grad])

    torch.cuda.reset_peak_memory_stats(dev)
    one_pass()          # warm-up (also grows the (tile, face) pair capacity)
    times = []
    for _ in range(args.c5_passes):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_pass()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        times.append(ms)
    sec = statistics.median(times) / 1e3
    loss_total = loss_acc.clone()
    if world > 1:
        dist.all_reduce(loss_total)
    bytes_view = wl.bview(Kc, Hc, Wc, v.shape[0], f.shape[0])
    gbs_per_gpu = views / world * bytes_view / sec / 1e9
    trb.set_near_plane_clipping("off")
    out = {"workload": f"{f.shape[0]}-face sphere (V={v.shape[0]}) x {views} views at {Hc}x{Wc}, K={Kc}, blur {wl.BLUR:.3e}, "
                       "SoftPhong+PointLights, fwd+bwd to verts (BASELINE configs[4])",
           "scaling": "strong", "n_gpus": world, "views_total": views, "views_per_gpu": hi - lo, "chunk_views": chunk,
           "views_per_s": round(views / sec, 2), "views_per_s_per_gpu": round(views / sec / world, 2),
           "seconds_per_pass": round(sec, 4), "passes_ms": [round(t, 1) for t in times],
           "algorithmic_bytes_per_view": int(bytes_view), "algorithmic_GBps_per_gpu": round(gbs_per_gpu, 1),
           "hbm_frac": round(gbs_per_gpu / peak, 4),
           "collective": "none" if world == 1 else "fused NCCL all-reduce of the 6 MB vertex gradient, once per pass",
           "peak_mem_GB": round(torch.cuda.max_memory_allocated(dev) / 1e9, 1),
           "loss": float(loss_total), "grad_norm": float(verts.grad.norm()), "launch_mode": "eager", "near_plane": "exact"}
    del rend, verts, faces, cols
    torch.cuda.empty_cache()
    return out


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (profiles/);
# filled in after each profiling pass, None until a capture of that kernel exists.
TRAFFIC_BYTES_PER_LAUNCH = {
    # profiles/r02_ncu_c2_head.txt (ncu --set full at HEAD, one launch each, 64 views of C2, sparse Fragments)
    "render_fine_kernel": 228_833_024,      # 6.82 MB read + 222.02 MB written (the 268 MB image, tail still in L2)
    "render_backward_kernel": 22_642_432,   # only covered pixels (1.9% of the image) are re-read
}


def cpu_baseline(sample_views=2, threads=0):
    """The oracle (kind 'port': C restatement of PyTorch3D's naive CPU rasteriser + torch-CPU shading)
    on this box's host cores, forward + backward, on `sample_views` views of the bench workload."""
    import oracle
    from helpers import fov_proj, load_mesh
    from oracle import shading_ref as sref
    import torch_renderer_b200 as trb
    v, f = load_mesh("cow")
    torch.manual_seed(0)
    colors = torch.rand(v.shape[0], 3)
    n = sample_views
    elev = torch.linspace(0, 360, VIEWS_PER_GPU)[:n]
    azim = torch.linspace(-180, 180, VIEWS_PER_GPU)[:n]
    R, T = trb.look_at_view_transform(dist=0.7, elev=elev, azim=azim)
    cores = oracle.num_threads() if threads <= 0 else threads
    torch.set_num_threads(cores)
    proj = fov_proj(n)
    ones = lambda *x: torch.tensor([list(x)], dtype=torch.float32).repeat(n, 1)
    grad_img = torch.randn(n, H, W, 4) / (n * H * W)
    t0 = time.perf_counter()
    vr, cr, Rr, Tr = (t.clone().requires_grad_(True) for t in (v, colors, R, T))
    ndc = sref.world_to_ndc(vr, Rr, Tr, proj[:, 0], proj[:, 1], proj[:, 2], proj[:, 3], True)
    fv = ndc[:, f].reshape(-1, 3, 3)
    first = np.arange(n, dtype=np.int64) * f.shape[0]
    count = np.full((n,), f.shape[0], dtype=np.int64)
    p2f, zbuf, bary, dists = oracle.rasterize_forward(fv.detach().numpy(), first, count, (H, W), 0.0, K, True,
                                                      False, False, cores)
    p2f_t = torch.from_numpy(p2f)
    zb, ba, di = (torch.from_numpy(a).requires_grad_(True) for a in (zbuf, bary, dists))
    cam = -torch.matmul(Tr[:, None, :], torch.linalg.inv(Rr))[:, 0, :]
    img = sref.shade(p2f_t, ba, zb, di, f.repeat(n, 1), vr, sref.vertex_normals(vr, f), cr, shader="soft_phong",
                     light_kind="point", light_vec=ones(0, 0, -3.0), light_ambient=ones(.5, .5, .5),
                     light_diffuse=ones(.3, .3, .3), light_specular=ones(.2, .2, .2), mat_ambient=ones(1, 1, 1),
                     mat_diffuse=ones(1, 1, 1), mat_specular=ones(1, 1, 1), shininess=torch.full((n,), 64.0),
                     camera_center=cam)
    img.backward(grad_img)
    g_fv = oracle.rasterize_backward(fv.detach().numpy(), p2f, zb.grad.numpy(), ba.grad.numpy(), di.grad.numpy(),
                                     True, False)
    fv.backward(torch.from_numpy(g_fv))
    dt = time.perf_counter() - t0
    return {"value": round(n / dt, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} of the 64 views of the same workload, forward+backward, {dt:.1f} s wall",
            "seconds": round(dt, 2)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  PyTorch3D (which holds it) is
    not installable, so this is the oracle port with all host threads; each step = 1 view fwd+bwd."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(sample_views=1)
    times = []
    for _ in range(max(1, min(args.steps, 5))):
        r = cpu_baseline(sample_views=1)
        times.append(r["seconds"])
    sec = statistics.mean(times)
    value = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world,
            "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": round(sec * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic cameras on the reference cow mesh geometry, seed 0",
            "config": {"workload": WORKLOAD, "sample": "each step = 1 view of the 64-view batch, forward+backward"},
            "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": "1 view per step, forward+backward, all host threads"},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-views", type=int, default=8, help="views in the cpu_baseline sample (~1.5 s each)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--repeats", type=int, default=5, help="timed repeats of --steps steps; the median is reported")
    ap.add_argument("--no-configs", action="store_true", help="skip the other_configs block (C1, C3, C4, pose step, C5 chunk)")
    ap.add_argument("--collective", choices=("fused", "separate"), default="separate",
                    help="N > 1: the all-reduce as its own kernel behind the backward (default: it measured 1.5-4.5 us "
                         "per step faster, profiles/r02_collective_ab.md) or pushed from the backward's tail kernel "
                         "(parallel.fused_backward_allreduce; adds a same-process A/B of both to the line)")
    ap.add_argument("--no-c5", action="store_true", help="skip BASELINE configs[4] at spec (1M faces x 1024 views)")
    ap.add_argument("--c5-views", type=int, default=1024)
    ap.add_argument("--c5-chunk", type=int, default=32, help="views per chunk of a rank's C5 slice (bounds Fragments memory)")
    ap.add_argument("--c5-passes", type=int, default=2)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
