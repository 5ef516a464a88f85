// sdw_attn.cu — fused (flash) attention on tcgen05 for the UNet's self- and cross-attention
// (the SDPA inside `BasicTransformerBlock`, reached from stable_diffusion_pipeline.py:418).
//
//   O[b, q, h*d:(h+1)*d] = softmax(Q_h K_h^T * d^-1/2) V_h          per (batch b, head h), fp16 in / fp16 out
//
// One CTA = one 128-query tile of one (b, h) (several tiles in turn when all keys fit one KV tile: cross attention).
// Nothing but Q, K, V^T tiles and the O tile touches HBM:
//   warp 0    : TMA producer.  Q tile once; per KV tile a K box [BKV x dk] and a V^T box [dv x BKV]
//               (SWIZZLE_128B, head dim zero-filled up to 64*DKA by TMA OOB) into a STAGES-deep ring.
//   warp 1    : MMA issuer.  S_j = Q K_j^T  (M=128, N=BKV, fp32 in TMEM) and O += P_j V_j (N = DVP, fp32 in TMEM).
//               Shipped variants for head dims <= 80 (PT = 1): A = P_j read from TENSOR MEMORY (TS-mode tcgen05.mma);
//               legacy / head dim 160: A = P_j from shared memory.
//   warps 2-5 : online softmax, thread = query row.  tcgen05.ld S_j, reference-max ("lazy") exponentials in fp32 with
//               packed FFMA2 / FADD2, exact slow path when a row max moves by more than 2^8 (O rescaled in TMEM),
//               P_j written back to tensor memory as fp16 pairs chunk by chunk (tcgen05.st), finally O / l -> fp16.
// Ordering is carried by mbarriers only (s_full, p_ready, pv_done, kv_full/empty, q_empty/o_free for the tile loop).
// Opt-in experiment variants kept for the record (all measured, none faster): split-S pipeline, two threads per query
// row (attn_pair_kernel), BKV = 64 with double-buffered S, exponentials on the FMA pipe — see variant_for().
#include "sdw_internal.h"
#include "sdw_ptx.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace sdw {

static constexpr int ATT_THREADS = 192;
static constexpr int ATT_BQ = 128;

struct alignas(64) AttnKParams {
  CUtensorMap mapQ, mapK, mapV;
  int Nq, Nk, d, heads;
  int dk_steps;          // ceil(d / 16)
  float scale_log2e;     // d^-1/2 * log2(e)
  __half* out;
  int64_t out_ld;
  int qt_per_cta;        // query tiles per CTA: > 1 only when all keys fit one KV tile (cross attention): K / V^T are
                         // then loaded once and the TMEM / barrier set-up is amortised over the tiles
};

// SB = number of S accumulator buffers: 2 (double-buffered, one CTA per SM) or 1 (TMEM 256 columns and <= 113 KB of
// shared memory, so TWO CTAs share an SM: one CTA's softmax overlaps the other's MMAs and both keep the MUFU pipe fed)
// PT = 1: P stays in tensor memory (BKV / 2 extra columns, fp16 pairs) and PV runs as a TS-mode MMA: no P tile in
// shared memory, no generic->async proxy fence, half the shared-memory traffic per KV tile.
template <int DKA, int DVP, int BKV, int ST, int SB, int PT = 0>
struct AttnCfg {
  static constexpr int Q_BYTES = DKA * ATT_BQ * 128;
  static constexpr int K_STAGE = DKA * BKV * 128;
  static constexpr int V_STAGE = (BKV / 64) * DVP * 128;
  static constexpr int P_BYTES = PT ? 0 : (BKV / 64) * ATT_BQ * 128;
  static constexpr int SMEM = Q_BYTES + P_BYTES + ST * (K_STAGE + V_STAGE) + 1024 + 256;
  static constexpr int P_COL = SB * BKV;                       // PT: P_j as packed fp16 pairs
  static constexpr int NEED = SB * BKV + DVP + (PT ? BKV / 2 : 0);
  static constexpr int TMEM_COLS = NEED <= 128 ? 128 : (NEED <= 256 ? 256 : 512);
  static constexpr int NCTA = SMEM <= 64 * 1024 && TMEM_COLS <= 128 ? 3 : (SMEM <= 114 * 1024 && TMEM_COLS <= 256 ? 2 : 1);  // CTAs per SM
  static constexpr int O_COL = SB * BKV + (PT ? BKV / 2 : 0);
  static_assert(NEED <= TMEM_COLS, "TMEM budget");
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^t for t <= 0 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax, max rel. error 7.5e-5 — below the fp16
// rounding of P): the MUFU pipe (16 ex2/clk/SM) is the attention bottleneck at head dim 40, so every 4th score is
// exponentiated here instead (the split FlashAttention-4 uses).
__device__ __forceinline__ float ex2_poly(float t) {
  t = fmaxf(t, -126.f);
  const float xr = __fadd_rd(t, 12582912.f);  // 1.5 * 2^23: the low mantissa bits now hold floor(t)
  const float f = t - (xr - 12582912.f);       // fractional part in [0, 1)
  float q = fmaf(f, 0.0780244991f, 0.2260671854f);
  q = fmaf(q, f, 0.6958335042f);
  q = fmaf(q, f, 0.9999251962f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(xr) << 23));
}

// SPLIT = 1 (BKV = 128, one S buffer, two CTAs per SM): S_j is produced and consumed as two 64-column halves with their
// own full / free barriers.  With a single S buffer the softmax warps used to idle from their last read of S_j until
// S_{j+1} had been issued (after ALL rows were done) and completed — a third of their stall samples in
// profiles/r01_ncu_attn_lazy.md.  Now S_{j+1}[0] is issued as soon as every row has read S_j[0], i.e. half a tile
// before it is needed, and the softmax treats each half as its own online-softmax step (P of half 0 is rescaled in
// shared memory in the rare case half 1 raises the row maximum past the lazy threshold).
template <int DKA, int DVP, int BKV, int ST, int SB, int POLY = 4, int SPLIT = 0, int PT = 0>  // POLY: every POLY-th exp on the FMA pipe (0 = none)
__global__ void __launch_bounds__(ATT_THREADS, AttnCfg<DKA, DVP, BKV, ST, SB, PT>::NCTA)
    attn_fwd_kernel(const __grid_constant__ AttnKParams p) {
  static_assert(!SPLIT || (SB == 1 && BKV == 128 && DKA == 1), "the split-S pipeline is the BKV = 128, single-buffer variant");
  static_assert(!PT || !SPLIT, "P-in-TMEM is a variant of the whole-tile pipeline");
  using Cfg = AttnCfg<DKA, DVP, BKV, ST, SB, PT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* p_smem = q_smem + Cfg::Q_BYTES;
  uint8_t* k_smem = p_smem + Cfg::P_BYTES;
  uint8_t* v_smem = k_smem + ST * Cfg::K_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + ST * Cfg::V_STAGE);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + ST;
  uint64_t* s_full = kv_empty + ST;  // [2]
  uint64_t* p_ready = s_full + 2;
  uint64_t* pv_done = p_ready + 1;
  uint64_t* s_free = pv_done + 1;  // [2]  (SPLIT)
  uint64_t* q_empty = s_free + 2;  // query-tile loop: Q smem free (S of this tile issued and complete)
  uint64_t* o_free = q_empty + 1;  //                  O accumulator read out by the epilogue
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int ntiles = (p.Nk + BKV - 1) / BKV;
  // query tiles of this CTA; tile t of the CTA is tile (g0 + j) of every barrier's phase sequence (nqt > 1 => ntiles == 1)
  const int qt_first = blockIdx.x * p.qt_per_cta;
  const int nqt = SPLIT ? 1 : min(p.qt_per_cta, (p.Nq + ATT_BQ - 1) / ATT_BQ - qt_first);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.mapQ);
    tma_prefetch_desc(&p.mapK);
    tma_prefetch_desc(&p.mapV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_ready, 128);
    mbar_init(pv_done, 1);
    mbar_init(&s_free[0], 128);
    mbar_init(&s_free[1], 128);
    mbar_init(q_empty, 1);
    mbar_init(o_free, 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ============================ TMA producer ============================================
    if (lane == 0) {
      for (int t = 0; t < nqt; ++t) {
      const int q0 = (qt_first + t) * ATT_BQ;
      if (t > 0) mbar_wait(q_empty, (t - 1) & 1);
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int a = 0; a < DKA; ++a) tma_load_4d(&p.mapQ, q_full, q_smem + a * (ATT_BQ * 128), a * 64, q0, head, b);
      if (t > 0) continue;  // single KV tile: K / V^T stay resident
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % ST;
        mbar_wait(&kv_empty[s], ((j / ST) & 1) ^ 1);
        mbar_expect_tx(&kv_full[s], Cfg::K_STAGE + Cfg::V_STAGE);
        const int kv0 = j * BKV;
#pragma unroll
        for (int a = 0; a < DKA; ++a)
          tma_load_4d(&p.mapK, &kv_full[s], k_smem + s * Cfg::K_STAGE + a * (BKV * 128), a * 64, kv0, head, b);
#pragma unroll
        for (int a = 0; a < BKV / 64; ++a)
          tma_load_4d(&p.mapV, &kv_full[s], v_smem + s * Cfg::V_STAGE + a * (DVP * 128), kv0 + a * 64, 0, head, b);
      }
      }  // query tiles
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ================================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(ATT_BQ, BKV);
      constexpr uint32_t idesc_o = make_idesc_f16(ATT_BQ, DVP);
      const uint32_t q_addr = smem_u32(q_smem);
      const uint32_t p_addr = smem_u32(p_smem);
      int g0 = 0;  // tiles issued by earlier query tiles of this CTA
      auto issue_s = [&](int j) {
        const int s = j % ST;
        mbar_wait(&kv_full[s], (j / ST) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(k_smem + s * Cfg::K_STAGE);
        const uint32_t ts = tmem + ((g0 + j) % SB) * BKV;
        for (int ks = 0; ks < p.dk_steps; ++ks) {
          const uint64_t da = make_desc_k_sw128(q_addr + (ks >> 2) * (ATT_BQ * 128) + (ks & 3) * 32);
          const uint64_t db = make_desc_k_sw128(k_addr + (ks >> 2) * (BKV * 128) + (ks & 3) * 32);
          umma_f16_ss(ts, da, db, idesc_s, ks != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[(g0 + j) % SB]);
      };
      constexpr uint32_t idesc_h = make_idesc_f16(ATT_BQ, 64);
      auto issue_half = [&](int j, int h) {  // S_j[:, 64h .. 64h+63] = Q K_j[64h .. 64h+63]^T
        const int s = j % ST;
        if (h == 0) {
          mbar_wait(&kv_full[s], (j / ST) & 1);
          tc_fence_after();
        }
        const uint32_t k_addr = smem_u32(k_smem + s * Cfg::K_STAGE) + h * (64 * 128);
        for (int ks = 0; ks < p.dk_steps; ++ks) {
          const uint64_t da = make_desc_k_sw128(q_addr + (ks & 3) * 32);
          const uint64_t db = make_desc_k_sw128(k_addr + (ks & 3) * 32);
          umma_f16_ss(tmem + h * 64, da, db, idesc_h, ks != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[h]);
      };
      for (int t = 0; t < nqt; ++t) {
      g0 = t * ntiles;
      mbar_wait(q_full, t & 1);
      tc_fence_after();
      if (SPLIT) {
        issue_half(0, 0);
        issue_half(0, 1);
      } else {
        issue_s(0);
        if (SB == 2 && ntiles > 1) issue_s(1);
        if (nqt > 1) umma_commit(q_empty);  // Q smem may be refilled once S of this (only) KV tile has completed
      }
      if (t > 0) {
        mbar_wait(o_free, (t - 1) & 1);  // the previous query tile's epilogue has read O out
        tc_fence_after();
      }
      for (int j = 0; j < ntiles; ++j) {
        if (SPLIT && j + 1 < ntiles) {
          // each half of S_{j+1} goes out as soon as every row has read that half of S_j
          mbar_wait(&s_free[0], j & 1);
          tc_fence_after();
          issue_half(j + 1, 0);
          mbar_wait(&s_free[1], j & 1);
          tc_fence_after();
          issue_half(j + 1, 1);
        }
        mbar_wait(p_ready, (g0 + j) & 1);
        tc_fence_after();
        // single S buffer: the softmax has consumed S_j (p_ready), so S_{j+1} goes first and overlaps PV_j
        if (!SPLIT && SB == 1 && j + 1 < ntiles) issue_s(j + 1);
        const int s = j % ST;
        const uint32_t v_addr = smem_u32(v_smem + s * Cfg::V_STAGE);
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          const uint64_t db = make_desc_k_sw128(v_addr + (ks >> 2) * (DVP * 128) + (ks & 3) * 32);
          if (PT) {
            umma_f16_ts(tmem + Cfg::O_COL, tmem + Cfg::P_COL + ks * 8, db, idesc_o, (j | ks) != 0 ? 1u : 0u);
          } else {
            const uint64_t da = make_desc_k_sw128(p_addr + (ks >> 2) * (ATT_BQ * 128) + (ks & 3) * 32);
            umma_f16_ss(tmem + Cfg::O_COL, da, db, idesc_o, (j | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(pv_done);
        umma_commit(&kv_empty[s]);
        if (SB == 2 && j + 2 < ntiles) issue_s(j + 2);
      }
      }  // query tiles
    }
  } else {
    // ============================ softmax / correction / epilogue ============================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_o = tmem + lane_base + Cfg::O_COL;
    const float sl2 = p.scale_log2e;
    uint8_t* p_row = p_smem + r * 128;
    const int sw = r & 7;
    for (int t = 0; t < nqt; ++t) {
    const int q0 = (qt_first + t) * ATT_BQ;
    const int g0 = t * ntiles;
    float m_run = -INFINITY, l_run = 0.f;

    // Two softmax paths per KV tile (the FlashAttention-4 "lazy rescale" idea):
    //  fast : exponentials are taken against the row's REFERENCE max m_run (the true running max as of the last slow
    //         tile) instead of this tile's max, so nothing depends on a whole-row reduction: 32-column TMEM loads are
    //         software-pipelined against the MUFU / FMA work of the previous chunk, and O is never rescaled.  P values
    //         may exceed 1, by at most 2^LAZY_LOG2 — harmless in fp16 P / fp32 sums since every term shares m_run.
    //  slow : the exact online-softmax update (true max, O rescale in TMEM).  Taken for tile 0, for ragged tiles, and
    //         — warp-uniformly — whenever any row's tile max exceeds its reference by more than 2^LAZY_LOG2
    //         (S_j is still intact in TMEM, so the tile is simply re-read).
    constexpr float LAZY_LOG2 = 8.f;
    if constexpr (SPLIT != 0) {
      for (int j = 0; j < ntiles; ++j) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&s_full[h], j & 1);
          tc_fence_after();
          const uint32_t t_s = tmem + lane_base + h * 64;
          const int c0 = j * BKV + h * 64;
          const bool ragged = c0 + 64 > p.Nk;
          bool need_slow = (j == 0 && h == 0) || ragged;
          float alpha = 1.f;
          uint32_t pkh[32];  // P of this row and half, packed fp16 pairs
          if (!need_slow) {
            // ---------------- fast path: reference max, 32-column chunks pipelined against the MUFU work ----------------
            const float mb = m_run * sl2;
            const uint64_t sl2_2 = pk2(sl2, sl2), nmb_2 = pk2(-mb, -mb);
            uint64_t sm2[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
            float mx[2] = {-INFINITY, -INFINITY};
            uint32_t va[32], vb[32];
            tmem_ld_32x32(t_s, va);
            tmem_ld_wait();
            tmem_ld_32x32(t_s + 32, vb);  // in flight while chunk 0 is exponentiated
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t(&cur)[32] = c ? vb : va;
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float x0 = __uint_as_float(cur[i]), x1 = __uint_as_float(cur[i + 1]);
                mx[0] = fmaxf(mx[0], x0);
                mx[1] = fmaxf(mx[1], x1);
                float t0, t1;
                upk2(fma2(pk2(x0, x1), sl2_2, nmb_2), t0, t1);
                const float e0 = ex2f(t0), e1 = ex2f(t1);
                sm2[(i >> 1) & 1] = add2(sm2[(i >> 1) & 1], pk2(e0, e1));
                pkh[c * 16 + (i >> 1)] = pack_h2(e0, e1);
              }
              if (c == 0) tmem_ld_wait();
            }
            const float m_t = fmaxf(mx[0], mx[1]);
            need_slow = __any_sync(0xffffffffu, (m_t - m_run) * sl2 > LAZY_LOG2);
            if (!need_slow) {
              float s0, s1, s2, s3;
              upk2(sm2[0], s0, s1);
              upk2(sm2[1], s2, s3);
              l_run += (s0 + s1) + (s2 + s3);
            }
          }
          if (need_slow) {
            // ---------------- slow path: exact online softmax on this half (S is still intact in TMEM) ----------------
            float v[64];
            {
              uint32_t vu[2][32];
              tmem_ld_32x32(t_s, vu[0]);
              tmem_ld_32x32(t_s + 32, vu[1]);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(vu[i / 32][i % 32]);
            }
            if (ragged) {
#pragma unroll
              for (int i = 0; i < 64; ++i)
                if (c0 + i >= p.Nk) v[i] = -INFINITY;
            }
            float mx[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
            for (int i = 4; i < 64; ++i) mx[i & 3] = fmaxf(mx[i & 3], v[i]);
            const float m_t = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
            const float m_new = fmaxf(m_run, m_t);  // finite: half 0 of tile 0 always holds a valid key
            alpha = ex2f((m_run - m_new) * sl2);
            const float mb = m_new * sl2;
            const uint64_t sl2_2 = pk2(sl2, sl2), nmb_2 = pk2(-mb, -mb);
            uint64_t sm2[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
#pragma unroll
            for (int i = 0; i < 64; i += 2) {
              float t0, t1;
              upk2(fma2(pk2(v[i], v[i + 1]), sl2_2, nmb_2), t0, t1);
              const float e0 = ex2f(t0), e1 = ex2f(t1);
              sm2[(i >> 1) & 1] = add2(sm2[(i >> 1) & 1], pk2(e0, e1));
              pkh[i >> 1] = pack_h2(e0, e1);
            }
            float s0, s1, s2, s3;
            upk2(sm2[0], s0, s1);
            upk2(sm2[1], s2, s3);
            l_run = fmaf(l_run, alpha, (s0 + s1) + (s2 + s3));
            m_run = m_new;
          }
          // every tcgen05.ld of this half has completed: the MMA warp may overwrite it with S_{j+1}
          tc_fence_before();
          mbar_arrive(&s_free[h]);
          if (h == 0 && j > 0) {
            // P smem and the O accumulator are free once PV_{j-1} has completed
            mbar_wait(pv_done, (j - 1) & 1);
            tc_fence_after();
          }
          if (need_slow && __any_sync(0xffffffffu, alpha != 1.f)) {
            if (j > 0) {
#pragma unroll
              for (int c = 0; c < DVP / 16; ++c) {
                uint32_t o[16];
                tmem_ld_32x16(t_o + c * 16, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st_32x16(t_o + c * 16, o);
              }
              tmem_st_wait();
            }
            if (h == 1) {
              // half 0 of this tile was exponentiated against the old maximum: bring it to the new one
              const __half2 a2 = __float2half2_rn(alpha);
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8) {
                uint4* q = reinterpret_cast<uint4*>(p_row + ((c8 ^ sw) << 4));
                uint4 u = *q;
                __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int i = 0; i < 4; ++i) hh[i] = __hmul2(hh[i], a2);
                *q = u;
              }
            }
          }
          // P half -> shared memory, K-major SWIZZLE_128B: 16-byte chunk c of row r lives at chunk (c ^ (r & 7))
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            uint4 u;
            u.x = pkh[c8 * 4 + 0];
            u.y = pkh[c8 * 4 + 1];
            u.z = pkh[c8 * 4 + 2];
            u.w = pkh[c8 * 4 + 3];
            *reinterpret_cast<uint4*>(p_row + h * (ATT_BQ * 128) + ((c8 ^ sw) << 4)) = u;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(p_ready);
      }
    } else {
    uint32_t pk[BKV / 2];  // P_j of this row, packed fp16 pairs

    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(&s_full[(g0 + j) % SB], ((g0 + j) / SB) & 1);
      tc_fence_after();
      const uint32_t t_s = tmem + lane_base + ((g0 + j) % SB) * BKV;
      const int kv0 = j * BKV;
      const bool ragged = kv0 + BKV > p.Nk;
      bool need_slow = (j == 0) || ragged;
      bool p_stored = false;  // PT: the fast path has already put P_j into tensor memory
      float alpha = 1.f;

      if (!need_slow) {
        // ---------------- fast path: chunk-pipelined, reference max ----------------
        const float mb = m_run * sl2;
        const uint64_t sl2_2 = pk2(sl2, sl2), nmb_2 = pk2(-mb, -mb);
        uint64_t sm2[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
        float mx[2] = {-INFINITY, -INFINITY};
        uint32_t va[32], vb[32];
        tmem_ld_32x32(t_s, va);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) {
          uint32_t(&cur)[32] = (c & 1) ? vb : va;
          uint32_t(&nxt)[32] = (c & 1) ? va : vb;
          if (c + 1 < BKV / 32) tmem_ld_32x32(t_s + (c + 1) * 32, nxt);  // in flight while this chunk is exponentiated
          uint32_t pkc[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float x0 = __uint_as_float(cur[i]), x1 = __uint_as_float(cur[i + 1]);
            mx[0] = fmaxf(mx[0], x0);
            mx[1] = fmaxf(mx[1], x1);
            float t0, t1;
            upk2(fma2(pk2(x0, x1), sl2_2, nmb_2), t0, t1);
            // POLY (with PT): one exponential of every POLY-th pair runs on the FMA pipe (1 / (2 POLY) of all scores)
            const float e0 = ex2f(t0);
            const float e1 = (PT && POLY > 0 && ((i >> 1) % (POLY > 0 ? POLY : 1)) == 0) ? ex2_poly(t1) : ex2f(t1);
            sm2[(i >> 1) & 1] = add2(sm2[(i >> 1) & 1], pk2(e0, e1));
            if (PT) pkc[i >> 1] = pack_h2(e0, e1);
            else pk[c * 16 + (i >> 1)] = pack_h2(e0, e1);
          }
          if (PT) {
            // P chunks go to tensor memory as they are produced (the store overlaps the next chunk's exponentials and
            // the row never holds all 64 packed registers); the P columns are free once PV_{j-1} has completed
            if (c == 0 && j > 0) {
              mbar_wait(pv_done, (g0 + j - 1) & 1);
              tc_fence_after();
            }
            tmem_st_32x16(tmem + lane_base + Cfg::P_COL + c * 16, pkc);
          }
          if (c + 1 < BKV / 32) tmem_ld_wait();
        }
        const float m_t = fmaxf(mx[0], mx[1]);
        need_slow = __any_sync(0xffffffffu, (m_t - m_run) * sl2 > LAZY_LOG2);
        p_stored = PT && !need_slow;
        if (!need_slow) {
          float s0, s1, s2, s3;
          upk2(sm2[0], s0, s1);
          upk2(sm2[1], s2, s3);
          l_run += (s0 + s1) + (s2 + s3);
        }
      }
      if (need_slow) {
        // ---------------- slow path: exact online softmax ----------------
        float v[BKV];
        {
          uint32_t vu[BKV / 32][32];
#pragma unroll
          for (int c = 0; c < BKV / 32; ++c) tmem_ld_32x32(t_s + c * 32, vu[c]);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < BKV; ++i) v[i] = __uint_as_float(vu[i / 32][i % 32]);
        }
        if (ragged) {
#pragma unroll
          for (int i = 0; i < BKV; ++i)
            if (kv0 + i >= p.Nk) v[i] = -INFINITY;
        }
        float mx[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
        for (int i = 4; i < BKV; ++i) mx[i & 3] = fmaxf(mx[i & 3], v[i]);
        const float m_t = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        const float m_new = fmaxf(m_run, m_t);
        alpha = ex2f((m_run - m_new) * sl2);
        const float mb = m_new * sl2;
        const uint64_t sl2_2 = pk2(sl2, sl2), nmb_2 = pk2(-mb, -mb);
        uint64_t sm2[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
#pragma unroll
        for (int i = 0; i < BKV; i += 2) {
          float t0, t1;
          upk2(fma2(pk2(v[i], v[i + 1]), sl2_2, nmb_2), t0, t1);
          const float e0 = ex2f(t0);
          const float e1 = (POLY > 0 && (i / 2) % (POLY > 1 ? POLY / 2 : 1) == (POLY > 1 ? POLY / 2 : 1) - 1)
                               ? ex2_poly(t1)
                               : ex2f(t1);
          sm2[(i >> 1) & 1] = add2(sm2[(i >> 1) & 1], pk2(e0, e1));
          pk[i >> 1] = pack_h2(e0, e1);
        }
        float s0, s1, s2, s3;
        upk2(sm2[0], s0, s1);
        upk2(sm2[1], s2, s3);
        l_run = fmaf(l_run, alpha, (s0 + s1) + (s2 + s3));
        m_run = m_new;
      }
      if (j > 0) {
        // P smem and the O accumulator are free once PV_{j-1} has completed
        mbar_wait(pv_done, (g0 + j - 1) & 1);
        tc_fence_after();
        if (need_slow && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
          for (int c = 0; c < DVP / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x16(t_o + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x16(t_o + c * 16, o);
          }
          tmem_st_wait();
        }
      }
      if (PT) {
        // P_j -> tensor memory: row = this thread's lane, column c = keys (2c, 2c+1) as an fp16 pair (the TS-mode A layout)
        if (!p_stored) {
          const uint32_t t_p = tmem + lane_base + Cfg::P_COL;
#pragma unroll
          for (int c = 0; c < BKV / 32; ++c) {
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = pk[c * 16 + i];
            tmem_st_32x16(t_p + c * 16, w);
          }
        }
        tmem_st_wait();
      } else {
        // P_j -> shared memory, K-major SWIZZLE_128B: 16-byte chunk c of row r lives at chunk (c ^ (r & 7))
#pragma unroll
        for (int c8 = 0; c8 < BKV / 8; ++c8) {
          uint4 u;
          u.x = pk[c8 * 4 + 0];
          u.y = pk[c8 * 4 + 1];
          u.z = pk[c8 * 4 + 2];
          u.w = pk[c8 * 4 + 3];
          *reinterpret_cast<uint4*>(p_row + (c8 >> 3) * (ATT_BQ * 128) + (((c8 & 7) ^ sw) << 4)) = u;
        }
        fence_proxy_async_smem();
      }
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    }  // !SPLIT
    // ---- epilogue: O / l -> fp16 ------------------------------------------------------------
    mbar_wait(pv_done, (g0 + ntiles - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const int row = q0 + r;
    __half* orow = p.out + (static_cast<int64_t>(b) * p.Nq + row) * p.out_ld + head * p.d;
    const bool vec_ok = ((p.out_ld & 7) == 0) && ((p.d & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
#pragma unroll
    for (int c = 0; c < DVP / 16; ++c) {
      uint32_t o[16];
      tmem_ld_32x16(t_o + c * 16, o);
      tmem_ld_wait();
      if (row < p.Nq) {
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          const int dd = c * 16 + h8 * 8;
          if (dd >= p.d) break;
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[h8 * 8 + i]) * inv_l;
          if (vec_ok && dd + 8 <= p.d) {
            uint4 u;
            u.x = pack_h2(f[0], f[1]);
            u.y = pack_h2(f[2], f[3]);
            u.z = pack_h2(f[4], f[5]);
            u.w = pack_h2(f[6], f[7]);
            *reinterpret_cast<uint4*>(orow + dd) = u;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (dd + i < p.d) orow[dd + i] = __float2half_rn(f[i]);
          }
        }
      }
    }
    if (nqt > 1) {  // O has been read out: the MMA warp may start the next query tile's PV
      tc_fence_before();
      mbar_arrive(o_free);
    }
    }  // query tiles
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::TMEM_COLS);
  }
}

// =============================================================================================
// attn_pair_kernel — head dim <= 64, BKV = 128, two threads per query row.
//
// attn_fwd_kernel keeps the MUFU pipe only 57 % busy at head dim 40 with one softmax warp per scheduler and CTA
// (profiles/r01_ncu_attn_lazy.md).  Hypothesis tested here: the per-row chain TMEM load -> FFMA2 -> MUFU -> pack is
// latency bound and more warps would fill the pipe.  Result: same speed (see variant_for) — kept as an opt-in variant.  EIGHT
// softmax warps per CTA (two CTAs per SM -> four per scheduler) split every row in two 64-column halves.  Both
// threads of a row share the reference maximum m_run, so the fast path needs no communication at all; the slow-path
// decision and the exact row maximum are agreed through shared memory and one 256-thread named barrier; the row sums
// stay per thread and meet in the epilogue.  P, the PV product and the O accumulator are unchanged (one 128 x 128 P
// tile, one accumulator), so the MMA warp is the same as in attn_fwd_kernel with a single S buffer.
// =============================================================================================
static constexpr int ATTP_THREADS = 320;  // warp 0 producer, warp 1 MMA, warps 2-9 softmax

__device__ __forceinline__ void bar_sync_softmax() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int DVP>
__global__ void __launch_bounds__(ATTP_THREADS, 2) attn_pair_kernel(const __grid_constant__ AttnKParams p) {
  constexpr int BKV = 128, ST = 2;
  using Cfg = AttnCfg<1, DVP, BKV, ST, 1>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* p_smem = q_smem + Cfg::Q_BYTES;
  uint8_t* k_smem = p_smem + Cfg::P_BYTES;
  uint8_t* v_smem = k_smem + ST * Cfg::K_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + ST * Cfg::V_STAGE);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + ST;
  uint64_t* s_full = kv_empty + ST;
  uint64_t* p_ready = s_full + 1;
  uint64_t* pv_done = p_ready + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pv_done + 1);
  int* s_vote = reinterpret_cast<int*>(tmem_ptr_smem + 2);  // [8]
  // [128][2] exchange of partial row maxima (slow path) / row sums (epilogue).  It overlays the head of the P tile,
  // which is idle whenever PV_{j-1} has completed and this tile's P rows have not been written yet.
  float* s_x = reinterpret_cast<float*>(p_smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int head = blockIdx.y, b = blockIdx.z;
  const int ntiles = (p.Nk + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.mapQ);
    tma_prefetch_desc(&p.mapK);
    tma_prefetch_desc(&p.mapV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
      tma_load_4d(&p.mapQ, q_full, q_smem, 0, q0, head, b);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % ST;
        mbar_wait(&kv_empty[s], ((j / ST) & 1) ^ 1);
        mbar_expect_tx(&kv_full[s], Cfg::K_STAGE + Cfg::V_STAGE);
        const int kv0 = j * BKV;
        tma_load_4d(&p.mapK, &kv_full[s], k_smem + s * Cfg::K_STAGE, 0, kv0, head, b);
#pragma unroll
        for (int a = 0; a < BKV / 64; ++a)
          tma_load_4d(&p.mapV, &kv_full[s], v_smem + s * Cfg::V_STAGE + a * (DVP * 128), kv0 + a * 64, 0, head, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(ATT_BQ, BKV);
      constexpr uint32_t idesc_o = make_idesc_f16(ATT_BQ, DVP);
      const uint32_t q_addr = smem_u32(q_smem);
      const uint32_t p_addr = smem_u32(p_smem);
      auto issue_s = [&](int j) {
        const int s = j % ST;
        mbar_wait(&kv_full[s], (j / ST) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(k_smem + s * Cfg::K_STAGE);
        for (int ks = 0; ks < p.dk_steps; ++ks)
          umma_f16_ss(tmem, make_desc_k_sw128(q_addr + ks * 32), make_desc_k_sw128(k_addr + ks * 32), idesc_s,
                      ks != 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        if (j + 1 < ntiles) issue_s(j + 1);  // every row has consumed S_j: S_{j+1} goes first and overlaps PV_j
        const int s = j % ST;
        const uint32_t v_addr = smem_u32(v_smem + s * Cfg::V_STAGE);
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          const uint64_t da = make_desc_k_sw128(p_addr + (ks >> 2) * (ATT_BQ * 128) + (ks & 3) * 32);
          const uint64_t db = make_desc_k_sw128(v_addr + (ks >> 2) * (DVP * 128) + (ks & 3) * 32);
          umma_f16_ss(tmem + Cfg::O_COL, da, db, idesc_o, (j | ks) != 0 ? 1u : 0u);
        }
        umma_commit(pv_done);
        umma_commit(&kv_empty[s]);
      }
    }
  } else {
    // ============================ softmax: thread = (row, 64-column half) ============================
    const int quarter = warp & 3;
    const int hh = warp >= 6 ? 1 : 0;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s = tmem + lane_base + hh * 64;
    const uint32_t t_o = tmem + lane_base + Cfg::O_COL;
    float m_run = -INFINITY, l_part = 0.f;
    const float sl2 = p.scale_log2e;
    uint8_t* p_row = p_smem + hh * (ATT_BQ * 128) + r * 128;
    const int sw = r & 7;
    constexpr float LAZY_LOG2 = 8.f;

    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int c0 = j * BKV + hh * 64;  // first key of this thread's half
      const bool ragged = j * BKV + BKV > p.Nk;
      bool need_slow = (j == 0) || ragged;  // CTA-uniform
      float alpha = 1.f;
      uint32_t pk[32];
      if (!need_slow) {
        // ---- fast path: reference max, 16-column TMEM chunks pipelined against the exponentials ----
        const float mb = m_run * sl2;
        const uint64_t sl2_2 = pk2(sl2, sl2), nmb_2 = pk2(-mb, -mb);
        uint64_t sm2 = pk2(0.f, 0.f);
        float mx0 = -INFINITY, mx1 = -INFINITY;
        uint32_t va[16], vb[16];
        tmem_ld_32x16(t_s, va);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t(&cur)[16] = (c & 1) ? vb : va;
          uint32_t(&nxt)[16] = (c & 1) ? va : vb;
          if (c + 1 < 4) tmem_ld_32x16(t_s + (c + 1) * 16, nxt);
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float x0 = __uint_as_float(cur[i]), x1 = __uint_as_float(cur[i + 1]);
            mx0 = fmaxf(mx0, x0);
            mx1 = fmaxf(mx1, x1);
            float t0, t1;
            upk2(fma2(pk2(x0, x1), sl2_2, nmb_2), t0, t1);
            const float e0 = ex2f(t0), e1 = ex2f(t1);
            sm2 = add2(sm2, pk2(e0, e1));
            pk[c * 8 + (i >> 1)] = pack_h2(e0, e1);
          }
          if (c + 1 < 4) tmem_ld_wait();
        }
        const bool vote = __any_sync(0xffffffffu, (fmaxf(mx0, mx1) - m_run) * sl2 > LAZY_LOG2);
        if (lane == 0) s_vote[warp - 2] = vote ? 1 : 0;
        bar_sync_softmax();
        int any = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) any |= s_vote[w];
        need_slow = any != 0;
        if (!need_slow) {
          float s0, s1;
          upk2(sm2, s0, s1);
          l_part += s0 + s1;
        }
      }
      if (need_slow) {
        // ---- slow path (tile 0, ragged tiles, a row max moved past the lazy threshold): exact, two TMEM passes ----
        if (j > 0) {
          mbar_wait(pv_done, (j - 1) & 1);  // the P tile doubles as the exchange buffer
          tc_fence_after();
        }
        float pm = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[16];
          tmem_ld_32x16(t_s + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (!ragged || c0 + c * 16 + i < p.Nk) pm = fmaxf(pm, __uint_as_float(v[i]));
        }
        s_x[r * 2 + hh] = pm;
        bar_sync_softmax();
        const float m_t = fmaxf(s_x[r * 2], s_x[r * 2 + 1]);
        bar_sync_softmax();  // all exchange reads precede the P rows that overwrite them
        const float m_new = fmaxf(m_run, m_t);  // finite: tile 0 always holds a valid key
        alpha = ex2f((m_run - m_new) * sl2);
        const float mb = m_new * sl2;
        const uint64_t sl2_2 = pk2(sl2, sl2), nmb_2 = pk2(-mb, -mb);
        uint64_t sm2 = pk2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[16];
          tmem_ld_32x16(t_s + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float x0 = __uint_as_float(v[i]), x1 = __uint_as_float(v[i + 1]);
            if (ragged) {
              if (c0 + c * 16 + i >= p.Nk) x0 = -INFINITY;
              if (c0 + c * 16 + i + 1 >= p.Nk) x1 = -INFINITY;
            }
            float t0, t1;
            upk2(fma2(pk2(x0, x1), sl2_2, nmb_2), t0, t1);
            const float e0 = ex2f(t0), e1 = ex2f(t1);
            sm2 = add2(sm2, pk2(e0, e1));
            pk[c * 8 + (i >> 1)] = pack_h2(e0, e1);
          }
        }
        float s0, s1;
        upk2(sm2, s0, s1);
        l_part = fmaf(l_part, alpha, s0 + s1);
        m_run = m_new;
      }
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // P smem and the O accumulator are free once PV_{j-1} has completed
        tc_fence_after();
        if (need_slow && __any_sync(0xffffffffu, alpha != 1.f)) {
          // the two threads of a row share the O columns: 16-column chunk c belongs to half (c & 1)
#pragma unroll
          for (int c = 0; c < DVP / 16; ++c) {
            if ((c & 1) != hh) continue;
            uint32_t o[16];
            tmem_ld_32x16(t_o + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x16(t_o + c * 16, o);
          }
          tmem_st_wait();
        }
      }
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        *reinterpret_cast<uint4*>(p_row + ((c8 ^ sw) << 4)) =
            make_uint4(pk[c8 * 4 + 0], pk[c8 * 4 + 1], pk[c8 * 4 + 2], pk[c8 * 4 + 3]);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // ---- epilogue: O / (l_a + l_b) -> fp16, the pair splits the 16-column chunks ----
    mbar_wait(pv_done, (ntiles - 1) & 1);
    tc_fence_after();
    s_x[r * 2 + hh] = l_part;
    bar_sync_softmax();
    const float inv_l = 1.f / (s_x[r * 2] + s_x[r * 2 + 1]);
    const int row = q0 + r;
    __half* orow = p.out + (static_cast<int64_t>(b) * p.Nq + row) * p.out_ld + head * p.d;
    const bool vec_ok = ((p.out_ld & 7) == 0) && ((p.d & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
#pragma unroll
    for (int c = 0; c < DVP / 16; ++c) {
      if ((c & 1) != hh) continue;  // warp-uniform
      uint32_t o[16];
      tmem_ld_32x16(t_o + c * 16, o);
      tmem_ld_wait();
      if (row < p.Nq) {
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          const int dd = c * 16 + h8 * 8;
          if (dd >= p.d) break;
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[h8 * 8 + i]) * inv_l;
          if (vec_ok && dd + 8 <= p.d) {
            *reinterpret_cast<uint4*>(orow + dd) =
                make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (dd + i < p.d) orow[dd + i] = __float2half_rn(f[i]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::TMEM_COLS);
  }
}

// =============================================================================================
// host
// =============================================================================================
struct AttnLaunchImpl {
  AttnKParams p;
  dim3 grid;
  int variant;
};

template <int DKA, int DVP, int BKV, int ST, int SB, int POLY = 4, int SPLIT = 0, int PT = 0>
static int attn_set_attr() {
  SDW_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<DKA, DVP, BKV, ST, SB, POLY, SPLIT, PT>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<DKA, DVP, BKV, ST, SB, PT>::SMEM));
  return 0;
}

static bool g_attn_init = false;
static int attn_init() {
  if (g_attn_init) return 0;
  if (int e = attn_set_attr<1, 16, 128, 2, 1>()) return e;
  if (int e = attn_set_attr<1, 32, 128, 2, 1>()) return e;
  if (int e = attn_set_attr<1, 48, 128, 2, 1>()) return e;
  if (int e = attn_set_attr<1, 64, 128, 2, 1>()) return e;
  if (int e = attn_set_attr<2, 80, 128, 2, 2>()) return e;
  if (int e = attn_set_attr<3, 160, 64, 3, 2>()) return e;
  if (int e = attn_set_attr<1, 48, 64, 2, 1>()) return e;
  if (int e = attn_set_attr<1, 48, 128, 2, 1, 0>()) return e;
  if (int e = attn_set_attr<1, 48, 128, 2, 1, 8>()) return e;
  if (int e = attn_set_attr<2, 80, 64, 2, 1, 0>()) return e;
  SDW_CUDA_OK(cudaFuncSetAttribute(attn_pair_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<1, 16, 128, 2, 1>::SMEM));
  SDW_CUDA_OK(cudaFuncSetAttribute(attn_pair_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<1, 32, 128, 2, 1>::SMEM));
  SDW_CUDA_OK(cudaFuncSetAttribute(attn_pair_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<1, 48, 128, 2, 1>::SMEM));
  SDW_CUDA_OK(cudaFuncSetAttribute(attn_pair_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<1, 64, 128, 2, 1>::SMEM));
  if (int e = attn_set_attr<1, 48, 64, 2, 2, 0>()) return e;
  if (int e = attn_set_attr<1, 16, 128, 2, 1, 0, 0, 1>()) return e;
  if (int e = attn_set_attr<1, 32, 128, 2, 1, 0, 0, 1>()) return e;
  if (int e = attn_set_attr<1, 48, 128, 2, 1, 0, 0, 1>()) return e;
  if (int e = attn_set_attr<1, 64, 128, 2, 1, 0, 0, 1>()) return e;
  if (int e = attn_set_attr<2, 80, 64, 2, 1, 0, 0, 1>()) return e;
  if (int e = attn_set_attr<1, 48, 128, 3, 1, 0, 0, 1>()) return e;
  if (int e = attn_set_attr<1, 48, 128, 2, 1, 1, 0, 1>()) return e;
  if (int e = attn_set_attr<1, 48, 128, 2, 1, 2, 0, 1>()) return e;
  if (int e = attn_set_attr<1, 48, 128, 2, 1, 3, 0, 1>()) return e;
  if (int e = attn_set_attr<1, 16, 128, 2, 1, 0, 1>()) return e;
  if (int e = attn_set_attr<1, 32, 128, 2, 1, 0, 1>()) return e;
  if (int e = attn_set_attr<1, 48, 128, 2, 1, 0, 1>()) return e;
  if (int e = attn_set_attr<1, 64, 128, 2, 1, 0, 1>()) return e;
  g_attn_init = true;
  return 0;
}

bool attn_supported(int d) { return d % 8 == 0 && d >= 8 && d <= 160; }

static int variant_for(int d) {
  // experiments: SDW_ATTN_VARIANT=18 -> head dims 33..48 on the BKV = 64, double-buffered-S tile (2 CTAs per SM)
  static const int forced = [] { const char* e = std::getenv("SDW_ATTN_VARIANT"); return e ? std::atoi(e) : -1; }();
  if (forced == 18 && d > 32 && d <= 48) return 18;
  if (forced == 24 && d > 32 && d <= 48) return 24;  // P in TMEM with a three-stage K/V ring
  if (forced >= 25 && forced <= 27 && d > 32 && d <= 48) return forced;  // P in TMEM + 1/2, 1/4, 1/6 of the exps on the FMA pipe
  // P in tensor memory + TS-mode PV (variants 19-22) for head dims <= 64: SDW_ATTN_PT=0 reverts to P in shared memory
  static const bool pt = [] { const char* e = std::getenv("SDW_ATTN_PT"); return !(e && e[0] == '0'); }();
  static const bool other = [] {
    return std::getenv("SDW_ATTN_BKV64") || std::getenv("SDW_ATTN_POLY") || std::getenv("SDW_ATTN_SPLIT") || std::getenv("SDW_ATTN_PAIR");
  }();
  if (pt && !other && d <= 64) return d <= 16 ? 19 : (d <= 32 ? 20 : (d <= 48 ? 21 : 22));
  if (pt && !other && d <= 80 && !std::getenv("SDW_ATTN_D80_BKV64")) return 23;
  // split-S pipeline (variants 10-13): correct, but measured 5 % SLOWER than the whole-tile variants (self-attention
  // 64x64, d = 40, batch 32: 1838 vs 1746 us, profiles/r01_attn_bench_split_s.txt) — the softmax warps' S waits were a
  // symptom, the MUFU + TMEM-read floor is what binds — so it is opt-in: SDW_ATTN_SPLIT=1
  static const bool split = [] { const char* e = std::getenv("SDW_ATTN_SPLIT"); return e && e[0] == '1'; }();
  static const bool legacy = [] {
    return std::getenv("SDW_ATTN_BKV64") != nullptr || std::getenv("SDW_ATTN_POLY") != nullptr;
  }();
  if (split && !legacy && d <= 64) return d <= 16 ? 10 : (d <= 32 ? 11 : (d <= 48 ? 12 : 13));
  // two threads per query row (variants 14-17): correct, but not faster than one thread per row (self-attention 64x64,
  // d = 40, batch 32: 1786 vs 1746 us, profiles/r01_attn_bench_pair.txt) — more softmax warps do not help either — so it
  // is opt-in: SDW_ATTN_PAIR=1
  static const bool pair = [] { const char* e = std::getenv("SDW_ATTN_PAIR"); return e && e[0] == '1'; }();
  if (pair && !legacy && d <= 64) return d <= 16 ? 14 : (d <= 32 ? 15 : (d <= 48 ? 16 : 17));
  if (d <= 16) return 0;
  if (d <= 32) return 1;
  static const bool bkv64 = [] { const char* e = std::getenv("SDW_ATTN_BKV64"); return e && e[0] == '1'; }();
  static const int poly = [] { const char* e = std::getenv("SDW_ATTN_POLY"); return e ? std::atoi(e) : 0; }();
  // measured equal within noise (profiles/r01_attn_bench_packed_poly.txt): the softmax warps are latency-, not
  // MUFU-bound, so the default is the exact MUFU path (POLY = 0); SDW_ATTN_POLY=4|8 selects the offload variants
  if (d <= 48 && d > 32 && !bkv64 && poly == 0) return 7;
  if (d <= 48 && d > 32 && !bkv64 && poly == 8) return 8;
  if (d <= 48 && d > 32 && poly != 4 && !bkv64) return 7;
  if (d <= 48) return bkv64 && d > 32 ? 6 : 2;
  if (d <= 64) return 3;
  // 64 < d <= 80: the BKV = 64 tile (2 CTAs per SM) beats the double-buffered BKV = 128 one (1 CTA per SM):
  // 109 vs 129 us self, 28 vs 44 us cross at batch 16 (profiles/r01_attn_bench_lazy.txt); SDW_ATTN_D80_BKV64=0 reverts
  static const bool d80_bkv128 = [] { const char* e = std::getenv("SDW_ATTN_D80_BKV64"); return e && e[0] == '0'; }();
  if (d <= 80) return d80_bkv128 ? 4 : 9;
  return 5;
}

int plan_attention(const AttnDesc& a, AttnLaunch* L) {
  SDW_REQUIRE(attn_supported(a.d), "flash attention supports head dims 8..160 (multiples of 8)");
  SDW_REQUIRE(a.q && a.k && a.vt && a.out, "null operand");
  SDW_REQUIRE(a.Nq > 0 && a.Nk > 0 && a.heads > 0 && a.B > 0, "empty attention");
  static_assert(sizeof(AttnLaunchImpl) <= sizeof(AttnLaunch::storage), "AttnLaunch storage too small");
  AttnLaunchImpl* I = reinterpret_cast<AttnLaunchImpl*>(L->storage);
  std::memset(I, 0, sizeof(*I));
  I->variant = variant_for(a.d);
  const int bkv = (I->variant == 5 || I->variant == 6 || I->variant == 9 || I->variant == 18 || I->variant == 23) ? 64 : 128;
  const int dvp_tab[28] = {16, 32, 48, 64, 80, 160, 48, 48, 48, 80, 16, 32, 48, 64, 16, 32, 48, 64, 48, 16, 32, 48, 64, 80, 48, 48, 48, 48};
  const int dvp = dvp_tab[I->variant];
  AttnKParams& p = I->p;
  p.Nq = a.Nq; p.Nk = a.Nk; p.d = a.d; p.heads = a.heads;
  p.dk_steps = (a.d + 15) / 16;
  p.scale_log2e = (1.f / std::sqrt(static_cast<float>(a.d))) * 1.4426950408889634f;
  p.out = a.out; p.out_ld = a.out_ld;
  {
    // cross attention (all keys in one KV tile): several query tiles per CTA, as long as >= ~3 waves of CTAs remain
    const int bkv_v = (I->variant == 5 || I->variant == 6 || I->variant == 9 || I->variant == 18 || I->variant == 23) ? 64 : 128;
    const int qtiles = (a.Nq + ATT_BQ - 1) / ATT_BQ;
    int qt = 1;
    static const int qt_env = [] { const char* e = std::getenv("SDW_ATTN_QT"); return e ? std::atoi(e) : 0; }();
    const bool pair_or_split = (I->variant >= 10 && I->variant <= 17);
    if (a.Nk <= bkv_v && !pair_or_split) {
      qt = qt_env > 0 ? qt_env : 8;  // cross attention 64x64, d = 40: 153 -> 119 us (profiles/r01_attn_bench_qtile_loop.txt)
      while (qt > 1 && static_cast<int64_t>((qtiles + qt - 1) / qt) * a.heads * a.B < 148 * 2 * 3) qt >>= 1;
      qt = std::max(1, std::min(qt, qtiles));
    }
    p.qt_per_cta = qt;
    I->grid = dim3((qtiles + qt - 1) / qt, a.heads, a.B);
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(a.d), static_cast<uint64_t>(a.Nq), static_cast<uint64_t>(a.heads),
                        static_cast<uint64_t>(a.B)};
    uint64_t str[4] = {1, static_cast<uint64_t>(a.q_ld), static_cast<uint64_t>(a.d),
                       static_cast<uint64_t>(a.Nq) * a.q_ld};
    uint32_t box[4] = {64, ATT_BQ, 1, 1};
    if (int e = encode_map(&p.mapQ, a.q, 4, dims, str, box)) return e;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(a.d), static_cast<uint64_t>(a.Nk), static_cast<uint64_t>(a.heads),
                        static_cast<uint64_t>(a.B)};
    uint64_t str[4] = {1, static_cast<uint64_t>(a.k_ld), static_cast<uint64_t>(a.d),
                       static_cast<uint64_t>(a.Nk) * a.k_ld};
    uint32_t box[4] = {64, static_cast<uint32_t>(bkv), 1, 1};
    if (int e = encode_map(&p.mapK, a.k, 4, dims, str, box)) return e;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(a.Nk), static_cast<uint64_t>(a.d), static_cast<uint64_t>(a.heads),
                        static_cast<uint64_t>(a.B)};
    uint64_t str[4] = {1, static_cast<uint64_t>(a.vt_ld), static_cast<uint64_t>(a.d) * a.vt_ld,
                       static_cast<uint64_t>(a.heads) * a.d * a.vt_ld};
    uint32_t box[4] = {64, static_cast<uint32_t>(dvp), 1, 1};
    if (int e = encode_map(&p.mapV, a.vt, 4, dims, str, box)) return e;
  }
  return 0;
}

// planner introspection (host only): {variant, query tiles per CTA, grid.x, grid.y, grid.z}
void attention_plan_info(const AttnLaunch& L, int out[5]) {
  const AttnLaunchImpl* I = reinterpret_cast<const AttnLaunchImpl*>(L.storage);
  out[0] = I->variant;
  out[1] = I->p.qt_per_cta;
  out[2] = static_cast<int>(I->grid.x);
  out[3] = static_cast<int>(I->grid.y);
  out[4] = static_cast<int>(I->grid.z);
}

int launch_attention(const AttnLaunch& L, cudaStream_t stream) {
  if (int e = attn_init()) return e;
  const AttnLaunchImpl* I = reinterpret_cast<const AttnLaunchImpl*>(L.storage);
  switch (I->variant) {
    case 0: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 16, 128, 2, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 16, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 1: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 32, 128, 2, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 32, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 2: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 48, 128, 2, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 48, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 3: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 64, 128, 2, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 64, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 4: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<2, 80, 128, 2, 2>, I->grid, dim3(ATT_THREADS), AttnCfg<2, 80, 128, 2, 2>::SMEM, stream, I->p)); break;
    case 5: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<3, 160, 64, 3, 2>, I->grid, dim3(ATT_THREADS), AttnCfg<3, 160, 64, 3, 2>::SMEM, stream, I->p)); break;
    case 6: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 48, 64, 2, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 48, 64, 2, 1>::SMEM, stream, I->p)); break;
    case 7: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 48, 128, 2, 1, 0>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 48, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 8: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 48, 128, 2, 1, 8>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 48, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 9: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<2, 80, 64, 2, 1, 0>, I->grid, dim3(ATT_THREADS), AttnCfg<2, 80, 64, 2, 1>::SMEM, stream, I->p)); break;
    case 10: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 16, 128, 2, 1, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 16, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 11: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 32, 128, 2, 1, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 32, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 12: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 48, 128, 2, 1, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 48, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 13: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 64, 128, 2, 1, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 64, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 19: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 16, 128, 2, 1, 0, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 16, 128, 2, 1, 1>::SMEM, stream, I->p)); break;
    case 20: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 32, 128, 2, 1, 0, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 32, 128, 2, 1, 1>::SMEM, stream, I->p)); break;
    case 21: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 48, 128, 2, 1, 0, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 48, 128, 2, 1, 1>::SMEM, stream, I->p)); break;
    case 22: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 64, 128, 2, 1, 0, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 64, 128, 2, 1, 1>::SMEM, stream, I->p)); break;
    case 25: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 48, 128, 2, 1, 1, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 48, 128, 2, 1, 1>::SMEM, stream, I->p)); break;
    case 26: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 48, 128, 2, 1, 2, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 48, 128, 2, 1, 1>::SMEM, stream, I->p)); break;
    case 27: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 48, 128, 2, 1, 3, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 48, 128, 2, 1, 1>::SMEM, stream, I->p)); break;
    case 24: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 48, 128, 3, 1, 0, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 48, 128, 3, 1, 1>::SMEM, stream, I->p)); break;
    case 23: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<2, 80, 64, 2, 1, 0, 0, 1>, I->grid, dim3(ATT_THREADS), AttnCfg<2, 80, 64, 2, 1, 1>::SMEM, stream, I->p)); break;
    case 18: SDW_CUDA_OK(launch_pdl(attn_fwd_kernel<1, 48, 64, 2, 2, 0>, I->grid, dim3(ATT_THREADS), AttnCfg<1, 48, 64, 2, 2>::SMEM, stream, I->p)); break;
    case 14: SDW_CUDA_OK(launch_pdl(attn_pair_kernel<16>, I->grid, dim3(ATTP_THREADS), AttnCfg<1, 16, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 15: SDW_CUDA_OK(launch_pdl(attn_pair_kernel<32>, I->grid, dim3(ATTP_THREADS), AttnCfg<1, 32, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 16: SDW_CUDA_OK(launch_pdl(attn_pair_kernel<48>, I->grid, dim3(ATTP_THREADS), AttnCfg<1, 48, 128, 2, 1>::SMEM, stream, I->p)); break;
    case 17: SDW_CUDA_OK(launch_pdl(attn_pair_kernel<64>, I->grid, dim3(ATTP_THREADS), AttnCfg<1, 64, 128, 2, 1>::SMEM, stream, I->p)); break;
    default: set_error("bad attention variant"); return 1;
  }
  SDW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sdw
