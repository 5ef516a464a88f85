// sdw_gemm2.cu — persistent 2-CTA (cta_group::2) variant of the implicit-GEMM kernel.
//
// Why: a 128 x BN tile moves (128 + BN) * 128 B of operands per 64-deep K block for 2 * 128 * BN * 64 FLOP —
// 64-71 FLOP per L2 byte, which caps the whole chip near 40 % of tensor peak (profiles/r01_ncu_gemm_v1.md: LTS 33 % at
// tensor 34 %).  A CTA pair computes a 256 x BN tile with ONE tcgen05.mma.cta_group::2 (M = 256): each CTA stages its
// own 128 activation rows and only HALF of the weight tile, so operand bytes per FLOP drop by ~1.6-2x, and the
// accumulators of two consecutive tiles double-buffer in TMEM so the epilogue of tile i overlaps the mainloop of
// tile i+1 (persistent grid: one cluster per SM pair, static round-robin over tiles, N fastest so an activation tile leaves DRAM once).
//
//   warp 0 (both CTAs)  : TMA producer — own A tile [128 x 64] + own half of B [BN/2 x 64]; completion bytes are
//                         signalled on the LEADER's full barrier (cta_group::2 TMA, mapa'd barrier address).
//   warp 1 (leader)     : MMA issuer, 4 x UMMA(M=256, N=BN, K=16) per K block; tcgen05.commit multicast frees the
//                         smem stage in both CTAs / publishes the accumulator to both epilogues.
//   warp 2 (both CTAs)  : residual producer of the TMA epilogue (one thread; a 4-deep ring of [128 x 32] chunks).
//   warps 4-11 (both)   : epilogue on the CTA's own 128 rows, two warps per TMEM lane quarter (sdw_gemm_epi.cuh), then a
//                         remote arrive on the leader's tmem_empty barrier.
//
// Mainloop flavours (all in this kernel; chosen per GEMM by plan_gemm, sdw_gemm.cu):
//   per-tap      one activation box per (tap, channel chunk)              every conv / linear (the original form)
//   TR = 1       tap reuse: one (8+2)-row box per (channel chunk, kx) feeds the three ky taps      3x3 stride-1 convs
//   a_stationary the activation rows of an M pair stay resident across its N tiles (opt-in, no gain measured)
//   CL = 4       activation tile TMA-multicast to two CTA pairs (opt-in, slower)
// Shared memory is carved at run time: [barriers 1 KB | operand ring | epilogue buffers]; the planner sizes the ring
// from what the chosen epilogue (classic: 16 KB, TMA: 40-72 KB) leaves of the 227 KB.
#include "sdw_gemm_epi.cuh"
#include "sdw_internal.h"
#include "sdw_ptx.cuh"

namespace sdw {

static constexpr int G2_THREADS = 384;  // warpgroup 0: producer, MMA (+2 idle warps); warpgroups 1-2: 8 epilogue warps
static constexpr int G2_A_STAGE = 128 * 64 * 2;

// NSUB = accumulators per activation tile: NSUB = 2 computes a 256 x (2*BN) tile per CTA pair — the A tile is pulled
// from L2 once for twice the columns (the kernel is L2->SM bandwidth bound: profiles/r01_mma_eff_vs_blockN.txt) at the
// price of single-buffered TMEM (2 * 2 * 160 > 512 columns), so it is used for long-K problems (3x3 convs) only.
//
// TR = 1 (tap reuse, 3x3 stride-1 convs on 16 x 8-pixel tiles): the kernel is bound by the operand bytes an SM ingests
// per K block (~41 B/clk/SM: profiles/r01_gemm_shapes_cluster4_multicast.txt), and the three ky taps of one kx read
// the same pixels shifted by whole image rows.  One TMA box of (8 + 2) rows x 16 px x 64 ch per (channel chunk, kx)
// therefore serves three taps: tap ky's A operand is the same shared-memory tile entered 16 rows (2 KB, swizzle-atom
// aligned) further down, so a pipeline stage is 20 KB of A + 3 weight tiles instead of 3 x 16 KB of A + 3 weight tiles.
static constexpr int G2_TR_BW = 16, G2_TR_BH = 8;
static constexpr int G2_A_STAGE_TR = (G2_TR_BH + 2) * G2_TR_BW * 128;
template <int BN, int NSUB, int TR = 0>
struct Gemm2Cfg {
  static constexpr int BH = BN / 2;
  static constexpr int B_SUB = BH * 128;           // one CTA's half of one BN-wide weight tile
  static constexpr int A_STAGE = TR ? G2_A_STAGE_TR : G2_A_STAGE;
  static constexpr int TAPS = TR ? 3 : 1;          // taps per pipeline stage
  static constexpr int B_STAGE = TAPS * NSUB * B_SUB;
  static constexpr int NBUF = (2 * NSUB * BN <= 512) ? 2 : 1;  // accumulator buffers in TMEM
  static constexpr int ACC_COLS = NSUB * BN;
  static constexpr int TMEM_COLS = NBUF * ACC_COLS <= 256 ? 256 : 512;
  // shared memory: [barriers 1 KB][nstages x A][nstages x B][epilogue buffers]; the planner sizes nstages from what the
  // chosen epilogue leaves (sdw_internal.h: G2_*), so the carve-up below is a run-time one
  static constexpr int MAX_STAGES = 8;
  static constexpr int SMEM_BYTES = G2_SMEM_DYN;
};

// CL = cluster size.  CL = 4: two CTA pairs of one cluster compute the two neighbouring N tiles of the same M pair; the
// activation tile is loaded ONCE per cluster (TMA multicast from pair 0 into both pairs' shared memory), which removes
// a third of the L2->SM operand traffic that bounds this kernel (profiles/r01_mma_eff_vs_blockN.txt).
template <int BN, int NSUB, int CL, int TR = 0>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(G2_THREADS, 1)
    gemm2_tc_kernel(const __grid_constant__ GemmKParams p) {
  static_assert(!(TR && CL != 2), "tap reuse is a CTA-pair variant");
  using Cfg = Gemm2Cfg<BN, NSUB, TR>;
  constexpr int A_STAGE = Cfg::A_STAGE;
  constexpr int NBUF = Cfg::NBUF;
  const int STAGES = p.nstages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + Cfg::MAX_STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::MAX_STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;               // [2]  (leader's copy is the one in use)
  uint64_t* res_full = tmem_empty + 2;                // [G2_RES_STAGES]  residual ring (TMA epilogue)
  uint64_t* res_empty = res_full + G2_RES_STAGES;
  uint64_t* a_full = res_empty + G2_RES_STAGES;  // [8]  A-stationary: resident activation slots
  uint64_t* a_empty = a_full + 8;                // [8]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(a_empty + 8);
  const bool AS = !TR && NSUB == 1 && CL == 2 && p.a_stationary != 0;
  const int NA = p.a_slots;
  uint8_t* smem_a = smem + G2_BAR_BYTES;
  uint8_t* smem_b = smem_a + (AS ? NA : STAGES) * A_STAGE;
  uint8_t* epi_stage = smem_b + STAGES * Cfg::B_STAGE;  // classic: 8 x 2 KB; TMA: 8 x 4 KB slabs, 8 x 1 KB bias, ring
  uint8_t* epi_bias = epi_stage + G2_EPI_OUT;
  uint8_t* res_ring = epi_bias + G2_EPI_BIAS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();  // rank in the cluster
  const uint32_t rank = crank & 1;           // rank in the CTA pair
  const uint32_t pair_id = crank >> 1;       // 0 for CL = 2
  const uint32_t lead_rank = crank & ~1u;    // cluster rank of this pair's leader
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x / CL;
  const int nclusters = gridDim.x / CL;
  constexpr int NPAIR = CL / 2;
  const int n_groups = (p.n_tiles + NPAIR - 1) / NPAIR;  // N tiles are handed out NPAIR at a time
  const int total_tiles = p.m_pairs * n_groups;
  const int num_kb = TR ? 3 * p.kchunks : p.ntaps * p.kchunks;  // pipeline stages per tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.mapA[0]);
    tma_prefetch_desc(&p.mapB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      // pair 0 also fills pair 1's activation stage, so its slot is free only when BOTH pairs' MMAs have drained it
      mbar_init(&empty_bar[s], (CL == 4 && pair_id == 0) ? 2 : 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 16);  // 8 epilogue warps x 2 CTAs
    }
    for (int r = 0; r < G2_RES_STAGES; ++r) {
      mbar_init(&res_full[r], 1);
      mbar_init(&res_empty[r], 4);  // the four warps (one per TMEM lane quarter) that own the chunk's parity
    }
    for (int r = 0; r < 8; ++r) {
      mbar_init(&a_full[r], 1);
      mbar_init(&a_empty[r], 1);
    }
    if (p.epi_tma) {
      tma_prefetch_desc(&p.mapOut);
      if (p.resid) tma_prefetch_desc(&p.mapRes);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();  // CTA-scope barrier between tcgen05.alloc's shared-memory write and its readers: the cluster barrier
                    // below already orders them, but compute-sanitizer racecheck only models bar.sync (108 false hazards)
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();               // the set-up above overlapped the previous kernel's tail; its outputs are visible from here
  pdl_launch_dependents();  // let the next kernel's CTAs start their own set-up as soon as SMs free up
  auto tile_coords = [&](int t, int& x0, int& y0, int& b0, int& n0) {
    // N fastest: the n_tiles column tiles of one M pair run on neighbouring clusters at the same time, so the
    // activation tile is fetched from DRAM once and re-read from L2 (the whole weight matrix is L2-resident anyway)
    const int m_pair = t / n_groups;
    const int n_tile = (t - m_pair * n_groups) * NPAIR + static_cast<int>(pair_id);
    const int m_tile = m_pair * 2 + static_cast<int>(rank);
    const int tw = m_tile % p.tiles_w;
    const int th = (m_tile / p.tiles_w) % p.tiles_h;
    const int tb = m_tile / (p.tiles_w * p.tiles_h);
    x0 = tw * p.bw;
    y0 = th * p.bh;
    b0 = tb * p.bb;
    n0 = n_tile * (NSUB * BN);
  };

  // i-th tile of this cluster.  Default: t = cluster_id + i * nclusters (N fastest across clusters).  A-stationary: the
  // cluster owns M pairs cluster_id, cluster_id + nclusters, ... and walks all N tiles of each in turn.
  auto tile_index = [&](int i, int& t) {
    if (AS) {
      const int mp = cluster_id + (i / n_groups) * nclusters;
      t = mp * n_groups + i % n_groups;
      return mp < p.m_pairs;
    }
    t = cluster_id + i * nclusters;
    return t < total_tiles;
  };

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) ==========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t ga = 0;  // A-stationary: activation chunks loaded so far
      int t;
      for (int i = 0; tile_index(i, t); ++i) {
        int x0, y0, b0, n0;
        tile_coords(t, x0, y0, b0, n0);
        if (AS) {
          const bool first_n = (t % n_groups) == 0;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            if (first_n) {
              const uint32_t slot = ga % NA;
              mbar_wait(&a_empty[slot], ((ga / NA) & 1) ^ 1);
              if (leader) mbar_expect_tx(&a_full[slot], 2 * A_STAGE);
              tma_load_4d_2sm(&p.mapA[0], mapa_rank(smem_u32(&a_full[slot]), lead_rank), smem_a + slot * A_STAGE, kc * 64, x0,
                              y0, b0);
              ++ga;
            }
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::B_STAGE);
            tma_load_4d_2sm(&p.mapB, mapa_rank(smem_u32(&full_bar[stage]), lead_rank), smem_b + stage * Cfg::B_STAGE, kc * 64,
                            n0 + static_cast<int>(rank) * Cfg::BH, 0, 0);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          continue;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * (A_STAGE + Cfg::B_STAGE));
          const uint32_t bar = mapa_rank(smem_u32(&full_bar[stage]), lead_rank);
          if (TR) {
            // stage = (channel chunk kc, column tap kx): one (bh+2)-row box + the weight tiles of taps (ky, kx), ky = 0..2
            const int kc = kb / 3, kx = kb - kc * 3;
            tma_load_4d_2sm(&p.mapA[0], bar, smem_a + stage * A_STAGE, kc * 64, x0 + kx - 1, y0 - 1, b0);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int sub = 0; sub < NSUB; ++sub)
                tma_load_4d_2sm(&p.mapB, bar, smem_b + stage * Cfg::B_STAGE + (ky * NSUB + sub) * Cfg::B_SUB,
                                ((ky * 3 + kx) * p.kchunks + kc) * 64, n0 + sub * BN + static_cast<int>(rank) * Cfg::BH, 0,
                                0);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          const int tap = kb / p.kchunks;
          const int kc = kb - tap * p.kchunks;
          if (CL == 2) {
            tma_load_4d_2sm(&p.mapA[p.tap_map[tap]], bar, smem_a + stage * A_STAGE, kc * 64, x0 + p.tap_dx[tap],
                            y0 + p.tap_dy[tap], b0);
          } else if (pair_id == 0) {
            // one L2 read feeds CTA `rank` of both pairs; each pair's leader barrier gets the bytes
            tma_load_4d_2sm_mc(&p.mapA[p.tap_map[tap]], bar, smem_a + stage * A_STAGE, kc * 64, x0 + p.tap_dx[tap],
                               y0 + p.tap_dy[tap], b0, static_cast<uint16_t>(0x5u << rank));
          }
#pragma unroll
          for (int sub = 0; sub < NSUB; ++sub)
            tma_load_4d_2sm(&p.mapB, bar, smem_b + stage * Cfg::B_STAGE + sub * Cfg::B_SUB, kb * 64,
                            n0 + sub * BN + static_cast<int>(rank) * Cfg::BH, p.b_batched ? y0 : 0,
                            p.b_batched ? b0 : 0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA only) =======================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(256, BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t ga_base = 0;  // A-stationary: first activation chunk of the current M pair
      int t;
      for (int it = 0; tile_index(it, t); ++it) {
        const int a = it % NBUF;
        mbar_wait(&tmem_empty[a], ((it / NBUF) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + a * Cfg::ACC_COLS;
        if (AS) {
          const int nidx = t % n_groups;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            const uint32_t g = ga_base + kc, slot = g % NA;
            if (nidx == 0) mbar_wait(&a_full[slot], (g / NA) & 1);
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t da = make_desc_k_sw128(smem_u32(smem_a + slot * A_STAGE));
            const uint64_t db = make_desc_k_sw128(smem_u32(smem_b + stage * Cfg::B_STAGE));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss_2cta(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
            umma_commit_2cta(&empty_bar[stage], 0b11);
            if (nidx == n_groups - 1) umma_commit_2cta(&a_empty[slot], 0b11);  // last N tile of this M pair
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (nidx == n_groups - 1) ga_base += p.kchunks;
          umma_commit_2cta(&tmem_full[a], 0b11);
          continue;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_desc_k_sw128(smem_u32(smem_a + stage * A_STAGE));
          const uint64_t db = make_desc_k_sw128(smem_u32(smem_b + stage * Cfg::B_STAGE));
          if (TR) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int sub = 0; sub < NSUB; ++sub)
                  umma_f16_ss_2cta(tmem_acc + sub * BN, da + ky * ((G2_TR_BW * 128) >> 4) + 2 * k,
                                   db + (ky * NSUB + sub) * (Cfg::B_SUB >> 4) + 2 * k, idesc, (kb | ky | k) != 0 ? 1u : 0u);
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
              for (int sub = 0; sub < NSUB; ++sub)
                umma_f16_ss_2cta(tmem_acc + sub * BN, da + 2 * k, db + sub * (Cfg::B_SUB >> 4) + 2 * k, idesc,
                                 (kb | k) != 0 ? 1u : 0u);
            }
          }
          // frees this pair's stage; pair 1 additionally releases pair 0's (whose producer also fills pair 1's A)
          umma_commit_2cta(&empty_bar[stage], CL == 2 ? 0b11 : (pair_id == 0 ? 0b0011 : 0b1111));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2cta(&tmem_full[a], static_cast<uint16_t>(0b11u << (2 * pair_id)));
      }
    }
  } else if (warp == 2) {
    // =========================== residual producer (TMA epilogue, both CTAs) =========
    // [128 rows x 32 columns] chunks of this CTA's residual tile, in the order the epilogue consumes them; runs ahead
    // of the epilogue by up to G2_RES_STAGES chunks, across tile boundaries
    if (NSUB == 1 && p.epi_tma && p.resid && lane == 0) {
      uint32_t gc = 0;
      int t;
      for (int i = 0; tile_index(i, t); ++i) {
        int x0, y0, b0, n0;
        tile_coords(t, x0, y0, b0, n0);
        const int nch = (max(0, min(BN, p.N - n0)) + 31) >> 5;
        for (int c = 0; c < nch; ++c, ++gc) {
          const uint32_t slot = gc % G2_RES_STAGES;
          mbar_wait(&res_empty[slot], ((gc / G2_RES_STAGES) & 1) ^ 1);
          mbar_expect_tx(&res_full[slot], G2_RES_STAGE);
          tma_load_4d(&p.mapRes, &res_full[slot], res_ring + slot * G2_RES_STAGE, n0 + c * 32, x0, y0, b0);
        }
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue (both CTAs, own 128 rows) ==================
    EpiTmaState est;
    int t;
    for (int it = 0; tile_index(it, t); ++it) {
      int x0, y0, b0, n0;
      tile_coords(t, x0, y0, b0, n0);
      const int a = it % NBUF;
      // the two warps of a lane quarter interleave 32-column chunks: twice the loads / stores in flight
      if (NSUB == 1 && p.epi_tma) {
        gemm_epilogue_tma<BN>(p, tmem_base + a * Cfg::ACC_COLS, warp, lane, x0, y0, b0, n0, &tmem_full[a], (it / NBUF) & 1,
                              (warp - 4) >> 2, epi_stage + (warp - 4) * 4096,
                              reinterpret_cast<float*>(epi_bias + (warp - 4) * 1024), res_ring, res_full, res_empty, est);
      } else {
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub)
          gemm_epilogue<BN>(p, tmem_base + a * Cfg::ACC_COLS + sub * BN, warp, lane, x0, y0, b0, n0 + sub * BN,
                            &tmem_full[a], (it / NBUF) & 1, (warp - 4) >> 2, 2, epi_stage + (warp - 4) * 2048);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&tmem_empty[a]), lead_rank));
    }
    if (p.epi_tma && lane == 0) bulk_wait_group<0>();  // every output slab has reached global memory
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int NSUB, int CL, int TR = 0>
static int set_attr2() {
  SDW_CUDA_OK(cudaFuncSetAttribute(gemm2_tc_kernel<BN, NSUB, CL, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Gemm2Cfg<BN, NSUB, TR>::SMEM_BYTES));
  return 0;
}

static int g_max_clusters4 = 0;
int gemm2_max_clusters4() { return g_max_clusters4; }

static bool g_init2 = false;
int gemm2_init() {
  if (g_init2) return 0;
  if (int e = set_attr2<128, 1, 2>()) return e;
  if (int e = set_attr2<160, 1, 2>()) return e;
  if (int e = set_attr2<192, 1, 2>()) return e;
  if (int e = set_attr2<256, 1, 2>()) return e;
  if (int e = set_attr2<160, 2, 2>()) return e;
  if (int e = set_attr2<160, 1, 4>()) return e;
  if (int e = set_attr2<256, 1, 4>()) return e;
  if (int e = set_attr2<128, 1, 2, 1>()) return e;
  if (int e = set_attr2<160, 1, 2, 1>()) return e;
  if (int e = set_attr2<192, 1, 2, 1>()) return e;
  if (int e = set_attr2<256, 1, 2, 1>()) return e;
  if (int e = set_attr2<160, 2, 2, 1>()) return e;
  {
    // how many 4-CTA clusters of this kernel the device can hold at once (GPC boundaries strand some SMs)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(G2_THREADS);
    cfg.dynamicSmemBytes = Gemm2Cfg<256, 1>::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm2_tc_kernel<256, 1, 4>, &cfg) != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      n = 32;
    }
    g_max_clusters4 = n;
  }
  g_init2 = true;
  return 0;
}

int launch_gemm2(const GemmLaunch& l, cudaStream_t stream) {
  if (int e = gemm2_init()) return e;
  if (l.tr) {
    if (l.cl != 2) {
      set_error("tap reuse is a CTA-pair variant");
      return 1;
    }
#define SDW_TR_CASE(BN_, NSUB_)                                                                                       \
  SDW_CUDA_OK(launch_pdl(gemm2_tc_kernel<BN_, NSUB_, 2, 1>, l.grid, dim3(G2_THREADS), Gemm2Cfg<BN_, NSUB_, 1>::SMEM_BYTES, \
                         stream, l.p))
    if (l.nsub == 2 && l.bn == 160) SDW_TR_CASE(160, 2);
    else if (l.nsub == 1 && l.bn == 128) SDW_TR_CASE(128, 1);
    else if (l.nsub == 1 && l.bn == 160) SDW_TR_CASE(160, 1);
    else if (l.nsub == 1 && l.bn == 192) SDW_TR_CASE(192, 1);
    else if (l.nsub == 1 && l.bn == 256) SDW_TR_CASE(256, 1);
    else {
      set_error("bad BLOCK_N / nsub for the tap-reuse kernel");
      return 1;
    }
#undef SDW_TR_CASE
    return 0;
  }
  if (l.cl == 4) {
    if (l.nsub != 1 || (l.bn != 160 && l.bn != 256)) {
      set_error("the 4-CTA cluster variant exists for BLOCK_N = 160 / 256, one accumulator");
      return 1;
    }
    if (l.bn == 160)
      SDW_CUDA_OK(launch_pdl(gemm2_tc_kernel<160, 1, 4>, l.grid, dim3(G2_THREADS), Gemm2Cfg<160, 1>::SMEM_BYTES, stream, l.p));
    else
      SDW_CUDA_OK(launch_pdl(gemm2_tc_kernel<256, 1, 4>, l.grid, dim3(G2_THREADS), Gemm2Cfg<256, 1>::SMEM_BYTES, stream, l.p));
    return 0;
  }
  if (l.nsub == 2) {
    if (l.bn != 160) {
      set_error("the two-accumulator variant exists for BLOCK_N = 160 only");
      return 1;
    }
    SDW_CUDA_OK(launch_pdl(gemm2_tc_kernel<160, 2, 2>, l.grid, dim3(G2_THREADS), Gemm2Cfg<160, 2>::SMEM_BYTES, stream, l.p));
    SDW_CUDA_OK(cudaGetLastError());
    return 0;
  }
  switch (l.bn) {
    case 128:
      SDW_CUDA_OK(launch_pdl(gemm2_tc_kernel<128, 1, 2>, l.grid, dim3(G2_THREADS), Gemm2Cfg<128, 1>::SMEM_BYTES, stream, l.p));
      break;
    case 160:
      SDW_CUDA_OK(launch_pdl(gemm2_tc_kernel<160, 1, 2>, l.grid, dim3(G2_THREADS), Gemm2Cfg<160, 1>::SMEM_BYTES, stream, l.p));
      break;
    case 192:
      SDW_CUDA_OK(launch_pdl(gemm2_tc_kernel<192, 1, 2>, l.grid, dim3(G2_THREADS), Gemm2Cfg<192, 1>::SMEM_BYTES, stream, l.p));
      break;
    case 256:
      SDW_CUDA_OK(launch_pdl(gemm2_tc_kernel<256, 1, 2>, l.grid, dim3(G2_THREADS), Gemm2Cfg<256, 1>::SMEM_BYTES, stream, l.p));
      break;
    default:
      set_error("bad BLOCK_N for the 2-CTA kernel");
      return 1;
  }
  SDW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sdw
