"""``BlendParams``.  Reference usage: ``BlendParams(sigma=1e-4, gamma=1e-4,
background_color=(0,0,0))`` (camera_pose_optimizer.py:109, torch_renderer.py:88).  The blend
arithmetic (softmax_rgb_blend / sigmoid_alpha_blend / hard_rgb_blend) is fused into the CUDA shade
kernel (csrc/shade.cu)."""
from __future__ import annotations

from typing import NamedTuple, Sequence, Union

import torch


class BlendParams(NamedTuple):
    sigma: float = 1e-4
    gamma: float = 1e-4
    background_color: Union[torch.Tensor, Sequence[float]] = (1.0, 1.0, 1.0)
