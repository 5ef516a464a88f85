// sdw_gemm_epi.cuh — the fused epilogue shared by the 1-CTA and 2-CTA tcgen05 implicit-GEMM kernels:
// TMEM accumulator (thread = output row) -> alpha / bias / time-embedding row / SiLU / residual / GEGLU /
// per-head V^T scatter -> fp16 global stores.
#pragma once
#include "sdw_internal.h"
#include "sdw_ptx.cuh"

namespace sdw {

__device__ __forceinline__ void store8(__half* dst, const float* v, bool vec_ok, int nvalid) {
  if (vec_ok && nvalid >= 8) {
    uint4 u;
    u.x = pack_h2(v[0], v[1]);
    u.y = pack_h2(v[2], v[3]);
    u.z = pack_h2(v[4], v[5]);
    u.w = pack_h2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst) = u;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nvalid) dst[j] = __float2half_rn(v[j]);
  }
}
__device__ __forceinline__ void load8(const __half* src, float* v, bool vec_ok, int nvalid) {
  if (vec_ok && nvalid >= 8) {
    uint4 u = *reinterpret_cast<const uint4*>(src);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h[j]);
      v[2 * j] = f.x;
      v[2 * j + 1] = f.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (j < nvalid) ? __half2float(src[j]) : 0.f;
  }
}


// warp = absolute warp index of an epilogue warp (its TMEM lane quarter is warp & 3); tmem_acc = accumulator base.
template <int BN>
__device__ __forceinline__ void gemm_epilogue(const GemmKParams& p, uint32_t tmem_acc, int warp, int lane, int x0,
                                              int y0, int b0, int n0, uint64_t* tmem_full_bar) {
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int r = quarter * 32 + lane;
    const int lx = r % p.bw;
    const int ly = (r / p.bw) % p.bh;
    const int lb = r / (p.bw * p.bh);
    const int x = x0 + lx, y = y0 + ly, b = b0 + lb;
    const bool row_ok = (x < p.W) && (y < p.H) && (b < p.B);
    const int64_t pix_in = (static_cast<int64_t>(b) * p.H + y) * p.W + x;  // lattice-linear index
    const int oyy = y * p.os + p.oy, oxx = x * p.os + p.ox;
    const int64_t out_off = static_cast<int64_t>(b) * p.o_sB + oyy * p.o_sH + oxx * p.o_sW;
    const int64_t res_off = static_cast<int64_t>(b) * p.r_sB + oyy * p.r_sH + oxx * p.r_sW;
    const bool vec_out = (((p.o_sW | p.o_sH | p.o_sB) & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const bool vec_res = p.resid && (((p.r_sW | p.r_sH | p.r_sB) & 7) == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.resid) & 15) == 0);
    const float* rowvec = p.rowvec ? p.rowvec + static_cast<int64_t>(b) * p.rowvec_ld : nullptr;

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16);

    if (p.mode == GEMM_GEGLU) {
      // packed columns: [32 value | 32 gate] pairs -> 32 outputs
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int n = n0 + c * 64;
        if (n >= p.N) break;
        uint32_t va[32], vg[32];
        tmem_ld_32x32(taddr + c * 64, va);
        tmem_ld_32x32(taddr + c * 64 + 32, vg);
        tmem_ld_wait();
        if (row_ok) {
          __half* dst = p.out + out_off + n / 2;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int jj = j8 * 8 + j;
              float a = __uint_as_float(va[jj]) * p.alpha;
              float g = __uint_as_float(vg[jj]) * p.alpha;
              if (p.bias) {
                a += __ldg(&p.bias[n + jj]);
                g += __ldg(&p.bias[n + 32 + jj]);
              }
              o[j] = a * gelu_erf_f(g);
            }
            store8(dst + j8 * 8, o, vec_out, 8);
          }
        }
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int n = n0 + c * 32;
        if (n >= p.N) break;
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        if (!row_ok) continue;
        const bool to_vt = (p.mode == GEMM_QKV_VT) && (n >= p.vt_col0);
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          const int nn = n + j8 * 8;
          const int nvalid = min(8, p.N - nn);
          if (nvalid <= 0) break;
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a = __uint_as_float(v[j8 * 8 + j]) * p.alpha;
            if (j < nvalid) {
              if (p.bias) a += __ldg(&p.bias[nn + j]);
              if (rowvec) a += __ldg(&rowvec[nn + j]);
            }
            if (p.act == 1) a = silu_f(a);
            o[j] = a;
          }
          if (to_vt) {
            const int64_t bq = pix_in / p.vt_ntok;
            const int tok = static_cast<int>(pix_in - bq * p.vt_ntok);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (j < nvalid) {
                const int cc = nn + j - p.vt_col0;
                const int head = cc / p.vt_d;
                const int dd = cc - head * p.vt_d;
                p.vt[((bq * p.vt_heads + head) * p.vt_d + dd) * p.vt_ld + tok] = __float2half_rn(o[j]);
              }
            }
          } else {
            if (p.resid) {
              float rr[8];
              load8(p.resid + res_off + nn, rr, vec_res, nvalid);
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] += rr[j];
            }
            store8(p.out + out_off + nn, o, vec_out, nvalid);
          }
        }
      }
    }
}

}  // namespace sdw
