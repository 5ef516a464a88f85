"""Per-tile timeline of the two-query-tile attention kernel (attn_pp_kernel): CTA 0 stamps clock64 at five points of
every KV tile for both softmax warpgroups (sdw_debug_attention_trace).  Prints the mean cycle budget per tile."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stable_diffusion_videos_b200 import _native as n  # noqa: E402

B, heads, Nq, Nk, d = (int(x) for x in os.environ.get("SHAPE", "16,8,4096,4096,40").split(","))
Cc = heads * d
q = torch.randn(B, Nq, Cc, device="cuda").half()
k = torch.randn(B, Nk, Cc, device="cuda").half()
vt = torch.randn(B, heads, d, Nk, device="cuda").half()
out = torch.empty(B, Nq, Cc, device="cuda", dtype=torch.float16)
buf = torch.zeros(2, 4096, 8, dtype=torch.int64, device="cuda")


def run():
    n.check(n.lib().sdw_attention(n.ptr(q), C.c_int64(Cc), n.ptr(k), C.c_int64(Cc), n.ptr(vt), C.c_int64(Nk), B, Nq, Nk,
                                  heads, d, n.ptr(out), C.c_int64(Cc), n.stream_ptr()))


for _ in range(2):
    run()
torch.cuda.synchronize()
n.lib().sdw_debug_attention_trace.restype = None
n.lib().sdw_debug_attention_trace(C.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
n.lib().sdw_debug_attention_trace(C.c_void_p(0))
t = buf.cpu()
nt = (Nk + 127) // 128
for X in range(2):
    rows = t[X]
    rows = rows[rows[:, 6] > 0]
    rows = rows[rows[:, 0].argsort()]
    if rows.shape[0] < 8:
        print("tile", "AB"[X], "no samples")
        continue
    body = rows[4:-4].double()
    d = lambda a, b: float((body[:, b] - body[:, a]).mean())  # noqa: E731
    period = float((body[1:, 0] - body[:-1, 0]).mean())
    print(f"query tile {'AB'[X]}: {rows.shape[0]} tiles; mean cycles between the stamps of a KV tile: "
          + "wait S %.0f, TMEM->regs %.0f, row max %.0f, exp + P store %.0f" % (d(0, 1), d(1, 2), d(2, 3), d(3, 6)) + f"; period {period:.0f} cycles per KV tile of {os.environ.get('BKV', '?')} keys")
