"""The BASELINE config step builders now live in ``bench_workloads.py`` at the repo root (bench.py reports them in
its ``other_configs`` block); this module keeps the old import path of the scripts in ``profiles/`` working."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import torch_renderer_b200 as trb  # noqa: E402,F401
from bench_workloads import *  # noqa: E402,F401,F403
from bench_workloads import BUILDERS, BLUR, SIGMA, bview, c1, c3, c4, c5, fibonacci_eyes, grid_sphere, pose_step  # noqa: E402,F401
