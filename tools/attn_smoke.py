"""Smallest possible run of the two-tile attention kernel (used under a short timeout before longer GPU jobs)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stable_diffusion_videos_b200 import _native as n  # noqa: E402

B, h, Nq, Nk, d = 1, 2, 512, 512, 40
Cc = h * d
q = torch.randn(B, Nq, Cc, device="cuda").half()
k = torch.randn(B, Nk, Cc, device="cuda").half()
v = torch.randn(B, Nk, Cc, device="cuda").half()
vt = v.reshape(B, Nk, h, d).permute(0, 2, 3, 1).contiguous()
out = torch.empty(B, Nq, Cc, device="cuda", dtype=torch.float16)
n.check(n.lib().sdw_attention(n.ptr(q), C.c_int64(Cc), n.ptr(k), C.c_int64(Cc), n.ptr(vt), C.c_int64(Nk), B, Nq, Nk, h, d,
                              n.ptr(out), C.c_int64(Cc), n.stream_ptr()))
torch.cuda.synchronize()
qf, kf, vf = (t.float().reshape(B, -1, h, d).transpose(1, 2) for t in (q, k, v))
ref = (torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, -1) @ vf).transpose(1, 2).reshape(B, Nq, Cc)
print("attn smoke max err", float((out.float() - ref).abs().max()))
