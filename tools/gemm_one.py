"""one GEMM shape, a few launches — target for `ncu --set full -k regex:gemm`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_bench import bench  # noqa: E402

if __name__ == "__main__":
    B, H, W, C, N, conv, bn = [int(v) for v in os.environ.get("SHAPE", "16,64,64,320,320,1,0").split(",")]
    ms, tf = bench(B, H, W, C, N, conv, iters=3, bn=bn, epi=bool(int(os.environ.get("EPI", "1"))))
    print(f"{ms*1e3:.1f} us {tf:.1f} TFLOP/s")
