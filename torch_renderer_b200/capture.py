"""``capture_step``: one whole optimisation step (renders, loss, backward, optimiser update) as ONE CUDA-graph replay.

The reference's own loops render ONE view per step (``camera_pose_optimizer.py:299-329``,
``mesh_deformer.py:181-222``): ~0.1 ms of kernels per render behind ~0.2 ms of Python, ctypes and launch latency,
plus the autograd engine's own time for the small torch ops around them (pose composition, losses).  Replaying the
step from a graph removes all of it; the step then costs what its kernels cost.

    step = trb.capture_step(lambda: one_step())     # one_step(): zero grads, render, loss.backward(), opt.step()
    for i in range(iters):
        step()                                      # replays; returns what one_step() returned (static tensors)
    step.check()                                    # raises if a vertex crossed the near plane in some replay

What a captured step cannot do is ask the host a question.  The one question the eager path asks -- does any
vertex lie behind the near clipping plane (``rasterizer.set_near_plane_clipping``)? -- is therefore asked
ASYNCHRONOUSLY: every render inside the step still runs the small test kernel, which raises a sticky device flag
that is copied to pinned memory by the graph itself; the flag is looked at before the next replay (without
waiting) and by ``check()`` (waiting).  A raised flag means some replay drew faces that cross the plane unclipped
(faces entirely behind it are culled in the kernels either way); ``on_near_plane`` selects what happens then:
``"raise"`` (default) raises ``NearPlaneCrossed``, ``"ignore"`` carries on.

Rules for ``fn`` (those of ``torch.cuda.graphs``): fixed shapes, no host reads (``.item()``, prints of tensors),
optimisers constructed with ``capturable=True`` where torch requires it, inputs updated IN PLACE between replays,
and the program's tensors living on a side stream from the start (``torch.cuda.set_stream(torch.cuda.Stream())``:
a leaf whose gradient was first accumulated on the legacy default stream cannot take part in a capture).
"""
from __future__ import annotations

from typing import Any, Callable, Optional

import torch

from . import ops


class NearPlaneCrossed(RuntimeError):
    """A captured step rendered a batch in which some vertex lay behind the near clipping plane."""


class CapturedStep:
    def __init__(self, fn: Callable[[], Any], warmup: int = 3, on_near_plane: str = "raise",
                 device: Optional[torch.device] = None):
        if on_near_plane not in ("raise", "ignore"):
            raise ValueError("on_near_plane must be 'raise' or 'ignore'")
        if not torch.cuda.is_available():
            raise RuntimeError("capture_step needs a CUDA device")
        self.fn = fn
        self.on_near_plane = on_near_plane
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.watch = ops.NearPlaneWatch(self.device)
        self.replays = 0
        self._done = torch.cuda.Event()
        # warm-up on a side stream (allocator warm-up, lazy initialisations, pair-capacity estimates), then capture
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with ops.near_plane_watch(self.watch):
            with torch.cuda.stream(side):
                for _ in range(max(int(warmup), 1)):
                    fn()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self._raise_if_tripped("during warm-up")
            self.graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(self.graph):
                    self.outputs = fn()
            except Exception as e:  # noqa: BLE001
                if "legacy stream" in str(e) or "StreamCapture" in str(e):
                    raise RuntimeError(
                        "capture_step: CUDA refused the capture because part of the step is bound to the legacy "
                        "default stream -- typically a leaf tensor whose gradient was first accumulated there.  "
                        "Create the parameters and run the loop under a side stream "
                        "(`torch.cuda.set_stream(torch.cuda.Stream())` at start-up), as torch's whole-network "
                        "capture recipe asks.") from e
                raise
        torch.cuda.synchronize(self.device)

    def _raise_if_tripped(self, when: str) -> None:
        if self.on_near_plane == "raise" and self.watch.tripped():
            raise NearPlaneCrossed(
                f"capture_step: a vertex lay behind the near clipping plane {when}; captured renders draw faces "
                "that cross the plane unclipped.  Run this step eagerly (the default 'exact' handling cuts such "
                "faces like PyTorch3D's clip_faces) or pass on_near_plane='ignore'.")

    def __call__(self):
        # non-blocking look at the flag the PREVIOUS replays left (the graph copies it to pinned memory itself)
        if self.replays and self._done.query():
            self._raise_if_tripped(f"in one of the first {self.replays} replays")
        self.graph.replay()
        self._done.record(torch.cuda.current_stream(self.device))
        self.replays += 1
        return self.outputs

    def check(self) -> None:
        """Waits for the replays issued so far and raises ``NearPlaneCrossed`` if one of them tripped the flag."""
        self._done.synchronize()
        self._raise_if_tripped(f"in one of the {self.replays} replays")


def capture_step(fn: Callable[[], Any], warmup: int = 3, on_near_plane: str = "raise") -> CapturedStep:
    """Captures ``fn`` (one full step) into a CUDA graph; see the module docstring."""
    return CapturedStep(fn, warmup=warmup, on_near_plane=on_near_plane)
