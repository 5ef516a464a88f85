import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stable_diffusion_videos_b200 import _native as n
torch.manual_seed(0)
Bn, Ntok, Cc, heads = 2, 300, 320, 8
d = Cc // heads; ld = 304
x = torch.randn(Bn, 1, Ntok, Cc, device="cuda").half()
w = (torch.randn(3 * Cc, Cc, 1, 1, device="cuda") * Cc ** -0.5).half()
wp = n.pack_weight(w)
qk = torch.full((Bn, Ntok, 2 * Cc), 3.0, dtype=torch.float16, device="cuda")
vt = torch.full((Bn, heads, d, ld), 5.0, dtype=torch.float16, device="cuda")
g = n.GemmDesc()
g.A = x.data_ptr(); g.C, g.W, g.H, g.B = Cc, Ntok, 1, Bn
g.sW, g.sH, g.sB = Cc, Ntok * Cc, Ntok * Cc
g.Wt = wp.data_ptr(); g.N = 3 * Cc
g.out = qk.data_ptr(); g.ldc = 2 * Cc
g.mode = 2; g.alpha = 1.0; g.ver = 2; g.et = 2
g.vt_col0, g.vt_d, g.vt_heads, g.vt_ntok = 2 * Cc, d, heads, Ntok
g.vt = vt.data_ptr(); g.vt_ld = ld
n.gemm(g); torch.cuda.synchronize()
pad = vt[..., Ntok:]
bad = (pad != 5.0).nonzero()
print("bad count", bad.shape[0], "of", pad.numel())
print(bad[:10].tolist())
print(pad[pad != 5.0][:10].tolist())
# guard test for plain ragged output
T, C, N = 5000, 320, 320
x = torch.randn(T, C, device="cuda").half()
w = (torch.randn(N, C, 1, 1, device="cuda") * C ** -0.5).half()
wp = n.pack_weight(w)
buf = torch.full((T + 256, N), 9.0, dtype=torch.float16, device="cuda")
g = n.GemmDesc()
g.A = x.data_ptr(); g.C, g.W, g.H, g.B = C, T, 1, 1
g.sW = C; g.Wt = wp.data_ptr(); g.N = N; g.out = buf.data_ptr(); g.ldc = N; g.alpha = 1.0; g.ver = 2; g.et = 2
n.gemm(g); torch.cuda.synchronize()
print("guard rows untouched:", bool((buf[T:] == 9.0).all()), "valid ok:", float((buf[:T].float() - x.float() @ w.float().reshape(N, C).t()).abs().max()))
