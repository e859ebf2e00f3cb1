"""Multi-rank host logic on CPU with the gloo backend, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torch_renderer_b200.parallel import allreduce_shared_grads, chunk_views, max_views_for_memory, shard_views


def test_shard_and_chunk_views():
    for n in (1, 7, 64, 1024):
        for w in (1, 2, 3, 8):
            spans = [shard_views(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)
    assert chunk_views(10, 4) == [(0, 4), (4, 8), (8, 10)]
    assert chunk_views(0, 4) == []
    # C5: 1024^2, K=8: 235 MB of Fragments per view -> a 30 GB budget holds ~70 views with grads
    n = max_views_for_memory(1024, 1024, 8, 30 * 10**9)
    assert 50 < n < 128


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        n_views, V = 6, 11
        per_view = torch.randn(n_views, V, 3)          # every rank knows the full problem
        s, e = shard_views(n_views, rank, world)
        g_verts = per_view[s:e].sum(0)                 # partial gradient from this rank's views
        g_cols = (per_view[s:e] ** 2).sum(0)
        # the fused form (push from the backward's tail kernel) needs NCCL + peer memory: on gloo the context yields
        # False and changes nothing, and the caller reduces afterwards -- the pattern of its docstring
        from torch_renderer_b200 import ops, parallel
        with parallel.fused_backward_allreduce() as fused:
            armed = ops._backward_peer_sum is not None
        if not fused:
            allreduce_shared_grads([g_verts, None, g_cols])
        ok = torch.allclose(g_verts, per_view.sum(0), atol=1e-5) and torch.allclose(g_cols, (per_view ** 2).sum(0), atol=1e-5)
        q.put((rank, bool(ok and not fused and not armed and ops._backward_peer_sum is None)))
    finally:
        dist.destroy_process_group()


def test_allreduce_shared_grads_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert res == [(0, True), (1, True)]


def test_allreduce_is_noop_without_process_group():
    g = torch.ones(3)
    assert allreduce_shared_grads([g]) is None and torch.equal(g, torch.ones(3))
    from torch_renderer_b200 import ops, parallel
    with parallel.fused_backward_allreduce() as fused:
        assert fused is False and ops._backward_peer_sum is None


def _peer_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from torch_renderer_b200 import parallel
        ok = True
        for trial, sizes in enumerate(([2930 * 3, 2930 * 3], [7], [500_002 * 3, 11, 64 * 9], [1, 2, 3, 4])):
            g = torch.Generator().manual_seed(100 + trial)
            full = [torch.randn(world, n, generator=g) for n in sizes]          # every rank knows all partials
            mine = [f[rank].clone().to(dev) for f in full]
            for _ in range(3):                                                   # repeated calls reuse the flags
                cur = [m.clone() for m in mine]
                allreduce_shared_grads(cur)
                torch.cuda.synchronize()
                for c, f in zip(cur, full):
                    want = f[0].clone()
                    for r in range(1, world):
                        want += f[r]                                             # rank order, like the kernel
                    ok = ok and torch.equal(c.cpu(), want)
        used_peer = any(v not in (None, False) for v in parallel._peer_allreduce.values())
        for v in parallel._peer_allreduce.values():
            if v not in (None, False):
                v.check()
        q.put((rank, bool(ok), bool(used_peer)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_memory_allreduce_two_gpus():
    """The one-kernel all-reduce over peer memory (csrc/allreduce.cu) on 2 GPUs: bit-identical to the rank-ordered
    sum, across segment layouts and repeated calls.  Skipped on a single-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert [r[:2] for r in res] == [(0, True), (1, True)]
    assert all(r[2] for r in res), "fell back to NCCL: symmetric memory unavailable"


def _shard_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import torch_renderer_b200 as trb
        from helpers import uv_sphere
        torch.manual_seed(0)
        v, f = uv_sphere(10, 14, noise=0.05, seed=1)
        cols = torch.rand(1, v.shape[0], 3)
        nv = 6
        R, T = trb.look_at_view_transform(dist=2.7, elev=torch.linspace(-30, 40, nv), azim=torch.linspace(0, 300, nv))

        def grads(lo, hi):
            vd = v.to(dev).requires_grad_(True)
            cd = cols.to(dev).requires_grad_(True)
            mesh = trb.Meshes([vd], [f.to(dev)], textures=trb.TexturesVertex(cd)).extend(hi - lo)
            cams = trb.FoVPerspectiveCameras(device=dev, R=R[lo:hi].to(dev), T=T[lo:hi].to(dev))
            rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=64)),
                                    trb.SoftPhongShader(device=dev, cameras=cams,
                                                        lights=trb.PointLights(device=dev, location=[[0.0, 1.0, -3.0]])))
            (rend(mesh)[..., :3] ** 2).sum().backward()
            return vd.grad, cd.grad

        lo, hi = shard_views(nv, rank, world)
        gv, gc = grads(lo, hi)
        allreduce_shared_grads([gv, gc])
        wv, wc = grads(0, nv)                      # the whole batch on this GPU alone
        torch.cuda.synchronize()
        err = max(float((gv - wv).norm() / wv.norm()), float((gc - wc).norm() / wc.norm()))
        q.put((rank, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_view_sharded_gradients_equal_single_gpu():
    """Views sharded over 2 GPUs + the shared-gradient all-reduce == all views on one GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert all(e < 1e-4 for _, e in res), res


def _fused_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import torch_renderer_b200 as trb
        from torch_renderer_b200 import parallel
        from helpers import uv_sphere
        torch.manual_seed(0)
        v, f = uv_sphere(10, 14, noise=0.05, seed=1)
        cols = torch.rand(1, v.shape[0], 3)
        nv = 6
        R, T = trb.look_at_view_transform(dist=2.7, elev=torch.linspace(-30, 40, nv), azim=torch.linspace(0, 300, nv))

        def grads(lo, hi, kind):
            vd = v.to(dev).requires_grad_(True)
            cd = cols.to(dev).requires_grad_(True)
            mesh = trb.Meshes([vd], [f.to(dev)], textures=trb.TexturesVertex(cd)).extend(hi - lo)
            cams = trb.FoVPerspectiveCameras(device=dev, R=R[lo:hi].to(dev), T=T[lo:hi].to(dev))
            if kind == "phong":
                rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=64)),
                                        trb.SoftPhongShader(device=dev, cameras=cams, lights=trb.PointLights(
                                            device=dev, location=[[0.0, 1.0, -3.0]])))
                (rend(mesh)[..., :3] ** 2).sum().backward()
                return [vd.grad, cd.grad]
            rend = trb.MeshRenderer(trb.MeshRasterizer(cams, trb.RasterizationSettings(
                image_size=64, blur_radius=2e-3, faces_per_pixel=6)), trb.SoftSilhouetteShader())
            (rend(mesh)[..., 3] ** 2).sum().backward()
            return [vd.grad]

        lo, hi = shard_views(nv, rank, world)
        worst, same, fused_all = 0.0, True, True
        for trial in range(3):
            for kind in ("phong", "silhouette"):
                with parallel.fused_backward_allreduce() as fused:
                    got = grads(lo, hi, kind)               # already summed over the ranks
                fused_all = fused_all and bool(fused)
                whole = grads(0, nv, kind)                  # the whole batch on this GPU alone, no exchange
                if trial == 1:                              # stand-alone calls share the inbox and the epochs
                    extra = [torch.full((5,), float(rank + 1), device=dev)]
                    allreduce_shared_grads(extra)
                    same = same and bool((extra[0] == sum(range(1, world + 1))).all())
                torch.cuda.synchronize()
                for g, w in zip(got, whole):
                    worst = max(worst, float((g - w).norm() / w.norm()))
                    both = [torch.empty_like(g) for _ in range(world)]
                    dist.all_gather(both, g.contiguous())
                    same = same and all(torch.equal(both[0], b) for b in both[1:])   # bit-identical on every rank
        parallel.check_peer_allreduce()
        q.put((rank, worst, bool(same), bool(fused_all)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_backward_with_fused_allreduce_two_gpus():
    """`fused_backward_allreduce`: the render backward's tail kernel pushes the shared gradients to the peer and a
    receive kernel sums them (trb_render_backward_allreduce).  Views sharded over 2 GPUs give the gradients of the
    whole batch on one GPU, bit-identical on both ranks, for two segments (vertices + colours) and one (vertices),
    over repeated calls and interleaved with the stand-alone all-reduce.  Skipped on a single-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fused_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert all(r[3] for r in res), "fell back: symmetric memory unavailable"
    assert all(r[2] for r in res), res
    assert all(r[1] < 1e-4 for r in res), res
