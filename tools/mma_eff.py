"""tcgen05 2-CTA GEMM efficiency vs BLOCK_N on a large compute-bound problem (M=65536, N=3840, K=2048)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_bench import bench  # noqa: E402

if __name__ == "__main__":
    for (B, H, W, C, N, conv) in [(1, 1, 65536, 2048, 3840, 0), (16, 64, 64, 320, 3840, 1)]:
        for bn in (128, 160, 192, 256):
            ms, tf = bench(B, H, W, C, N, conv, bn=bn, ver=2, iters=5)
            print(f"M={B*H*W} N={N} C={C} conv={conv} bn={bn}: {ms*1e3:9.1f} us {tf:8.1f} TFLOP/s", flush=True)
