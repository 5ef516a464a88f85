"""LayerNorm / GroupNorm on the 64x64-level shapes at batch 2F (HBM-bound passes): time and achieved bytes/s.
F=30 python tools/norm_bench.py          ONLY=ln|gn ITERS=2 gives an ncu target."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stable_diffusion_videos_b200 import _native as n  # noqa: E402

F = int(os.environ.get("F", "30"))
B = 2 * F
iters = int(os.environ.get("ITERS", "20"))
only = os.environ.get("ONLY")


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for (hw, C) in ((64, 320), (32, 640)):
    rows = B * hw * hw
    x = torch.randn(rows, C, device="cuda").half()
    y = torch.empty_like(x)
    g, b = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
    if only in (None, "ln"):
        us = timed(lambda: n.layernorm(x, rows, C, g, b, 1e-5, y))
        print(f"layernorm C{C} rows{rows}: {us:7.1f} us  {2 * rows * C * 2 / us / 1e6:6.2f} TB/s (read + write)", flush=True)
    if only in (None, "gn"):
        x4, y4 = x.view(B, hw * hw, C), y.view(B, hw * hw, C)
        us = timed(lambda: n.groupnorm(x4, B, hw * hw, C, 32, g, b, 1e-5, 1, y4))
        print(f"groupnorm+silu C{C} {B}x{hw}x{hw}: {us:7.1f} us  {3 * rows * C * 2 / us / 1e6:6.2f} TB/s (2 reads + write)", flush=True)
