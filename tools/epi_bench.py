"""Short-K linear layers of the SD-1.4 transformer blocks (token lattice, batch 2F): these are epilogue / HBM bound.
Prints time, TFLOP/s and the achieved fraction of the HBM floor.  ONLY=<idx> ITERS=3 gives an ncu target."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stable_diffusion_videos_b200 import _native as n  # noqa: E402


def bench(T, C, N, mode=0, bias=True, resid=False, iters=20, bn=0, et=0, ew=0):
    x = torch.randn(T, C, device="cuda").half()
    w = (torch.randn(N, C, 1, 1, device="cuda") * C ** -0.5).half()
    wp = n.pack_weight(w, geglu=(mode == 1))
    ncols = N // 2 if mode == 1 else N
    out = torch.empty(T, ncols, device="cuda", dtype=torch.float16)
    d = n.GemmDesc()
    d.A = x.data_ptr(); d.C, d.W, d.H, d.B = C, T, 1, 1
    d.sW = C
    d.Wt = wp.data_ptr(); d.N = N
    d.out = out.data_ptr(); d.ldc = ncols
    d.alpha = 1.0; d.mode = mode; d.bn = bn; d.et = et; d.ew = ew
    keep = []
    if bias:
        b = torch.randn(N, device="cuda"); keep.append(b); d.bias = b.data_ptr()
    if resid:
        r = torch.randn(T, ncols, device="cuda").half(); keep.append(r); d.resid = r.data_ptr(); d.ldr = ncols
    for _ in range(3):
        n.gemm(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        n.gemm(d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    byt = 2.0 * T * (C + ncols * (2 if resid else 1))
    return ms, 2.0 * T * C * N / ms / 1e9, byt / ms / 1e6


if __name__ == "__main__":
    F = int(os.environ.get("F", "16"))
    T = 2 * F * 4096
    shapes = [
        ("ff.geglu 64x64 320->2560", T, 320, 2560, 1, True, False),
        ("attn.out 64x64 320->320 +res", T, 320, 320, 0, True, True),
        ("ff.out 64x64 1280->320 +res", T, 1280, 320, 0, True, True),
        ("qkv-like 64x64 320->960", T, 320, 960, 0, False, False),
        ("ff.geglu 32x32 640->5120", T // 4, 640, 5120, 1, True, False),
        ("attn.out 32x32 640->640 +res", T // 4, 640, 640, 0, True, True),
        ("ff.geglu 16x16 1280->10240", T // 16, 1280, 10240, 1, True, False),
    ]
    only = os.environ.get("ONLY")
    iters = int(os.environ.get("ITERS", "20"))
    for i, (name, t, c, nn, mode, bias, resid) in enumerate(shapes):
        if only is not None and int(only) != i:
            continue
        et_env = int(os.environ.get("ET", "0"))
        ew_env = int(os.environ.get("EW", "0"))
        # ONLY=<i>: one variant (ncu target); otherwise the epilogue-width A/B (8 vs 16 epilogue warps) at auto BLOCK_N
        variants = ((0, et_env, ew_env),) if only is not None else ((0, 2, 2), (0, 2, 4))
        for bn, et, ew in variants:
            ms, tf, gbs = bench(t, c, nn, mode, bias, resid, iters=iters, bn=bn, et=et, ew=ew)
            print(f"{name:34s} bn={bn or 'auto':>4} et={et} ew={ew} {ms*1e3:8.1f} us {tf:7.1f} TFLOP/s {gbs:7.1f} GB/s algorithmic", flush=True)
