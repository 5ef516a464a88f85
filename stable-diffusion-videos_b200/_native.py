"""ctypes binding of libsdwalk.so (the C ABI in include/sdwalk.h).

torch tensors are the only host container: every call passes `tensor.data_ptr()` plus explicit
sizes and the current CUDA stream.  There is NO CPU fallback: if the shared library is missing
or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdwalk.so")
_lib = None


class SdwError(RuntimeError):
    pass


class StepCoef(C.Structure):
    _fields_ = [
        ("guidance", C.c_float), ("c_x", C.c_float), ("c_e", C.c_float * 5),
        ("hist_slot", C.c_int32 * 4), ("use_x_base", C.c_int32), ("save_x_base", C.c_int32),
        ("push_slot", C.c_int32), ("next_in_scale", C.c_float), ("push_e", C.c_float), ("push_x", C.c_float),
    ]


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("C", C.c_int32), ("W", C.c_int32), ("H", C.c_int32), ("B", C.c_int32),
        ("sW", C.c_int64), ("sH", C.c_int64), ("sB", C.c_int64),
        ("conv", C.c_int32), ("up_px", C.c_int32), ("up_py", C.c_int32),
        ("Wt", C.c_void_p), ("N", C.c_int32), ("ldb", C.c_int64), ("Kb", C.c_int64),
        ("b_batched", C.c_int32), ("sBh", C.c_int64), ("sBb", C.c_int64),
        ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("rowvec_ld", C.c_int32),
        ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("out", C.c_void_p), ("ldc", C.c_int64),
        ("o_sW", C.c_int64), ("o_sH", C.c_int64), ("o_sB", C.c_int64),
        ("mode", C.c_int32), ("act", C.c_int32), ("alpha", C.c_float),
        ("vt_col0", C.c_int32), ("vt_d", C.c_int32), ("vt_heads", C.c_int32), ("vt_ntok", C.c_int32),
        ("vt", C.c_void_p), ("vt_ld", C.c_int64), ("bn", C.c_int32), ("ver", C.c_int32), ("nsub", C.c_int32), ("ew", C.c_int32), ("tr", C.c_int32), ("et", C.c_int32), ("reserved0", C.c_int32),
    ]


def lib():
    """Load libsdwalk.so (once).  Fails loudly: the CUDA extension IS the product path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdwError(
                f"{LIB_PATH} not built: run `python __graft_entry__.py build` "
                "(the native sm_100a library is required; there is no fallback path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.sdw_last_error.restype = C.c_char_p
    return _lib


def check(rc):
    if rc != 0:
        raise SdwError(f"libsdwalk error {rc}: {lib().sdw_last_error().decode()}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SdwError("libsdwalk operates on CUDA tensors only (no CPU fallback)")


# ----------------------------------------------------------------------------------------------
def slerp_lerp_batch(lat_a, lat_b, emb_a, emb_b, t, dot_threshold=0.9995):
    """Batched generate_inputs math (stable_diffusion_pipeline.py:466-468): returns (latents[n], embeds[n])."""
    require_cuda(lat_a, lat_b, emb_a, emb_b, t)
    assert lat_a.dtype == lat_b.dtype == emb_a.dtype == emb_b.dtype and lat_a.dtype in (torch.float16, torch.float32)
    n = t.numel()
    t = t.to(torch.float32).contiguous()
    out_lat = torch.empty((n,) + tuple(lat_a.shape[1:] if lat_a.shape[0] == 1 else lat_a.shape),
                          dtype=lat_a.dtype, device=lat_a.device)
    out_emb = torch.empty((n,) + tuple(emb_a.shape[1:] if emb_a.shape[0] == 1 else emb_a.shape),
                          dtype=emb_a.dtype, device=emb_a.device)
    check(lib().sdw_slerp_lerp_batch(ptr(lat_a.contiguous()), ptr(lat_b.contiguous()), ptr(emb_a.contiguous()),
                                     ptr(emb_b.contiguous()), ptr(t), C.c_int(n), C.c_int64(lat_a.numel()),
                                     C.c_int64(emb_a.numel()), C.c_int(lat_a.dtype == torch.float16),
                                     C.c_float(dot_threshold), ptr(out_lat), ptr(out_emb), stream_ptr()))
    return out_lat, out_emb


def pack_weight(w, geglu=False):
    """OIHW / [N,K] fp16 weight -> K-major [N][taps][ceil64(C)] layout of the tcgen05 kernel."""
    require_cuda(w)
    w = w.to(torch.float16).contiguous()
    if w.dim() == 2:
        w = w[:, :, None, None]
    N, Cc, kh, kw = w.shape
    cp = (Cc + 63) // 64 * 64
    out = torch.empty((N, kh * kw * cp), dtype=torch.float16, device=w.device)
    check(lib().sdw_pack_weight(ptr(w), N, Cc, kh, kw, int(geglu), ptr(out), stream_ptr()))
    return out


def pack_weight_up4(w):
    """[N, C, 3, 3] upsampler weight -> [4 parities][N][4 taps][ceil64(C)] (taps pre-summed in fp32)."""
    require_cuda(w)
    w = w.to(torch.float16).contiguous()
    N, Cc = w.shape[0], w.shape[1]
    cp = (Cc + 63) // 64 * 64
    out = torch.empty((4, N, 4 * cp), dtype=torch.float16, device=w.device)
    check(lib().sdw_pack_weight_up4(ptr(w), N, Cc, ptr(out), stream_ptr()))
    return out


def groupnorm(x, B, P, Cc, G, gamma, beta, eps, silu, y):
    """x, y: fp16 [B][P][ld] views whose last dim may be a channel slice of a wider buffer (ld = stride of dim -2)."""
    require_cuda(x, y, gamma, beta)
    check(lib().sdw_groupnorm(ptr(x), C.c_int64(x.stride(-2)), C.c_int(B), C.c_int64(P), C.c_int(Cc), C.c_int(G), ptr(gamma),
                              ptr(beta), C.c_float(eps), C.c_int(int(silu)), ptr(y), C.c_int64(y.stride(-2)), stream_ptr()))


def layernorm(x, rows, Cc, gamma, beta, eps, y):
    require_cuda(x, y, gamma, beta)
    check(lib().sdw_layernorm(ptr(x), C.c_int64(x.stride(-2)), C.c_int64(rows), C.c_int(Cc), ptr(gamma), ptr(beta),
                              C.c_float(eps), ptr(y), C.c_int64(y.stride(-2)), stream_ptr()))


def gemm(desc):
    check(lib().sdw_gemm(C.byref(desc), stream_ptr()))
