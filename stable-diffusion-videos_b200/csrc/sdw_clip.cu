// sdw_clip.cu — the CLIP text tower behind `embed_text` (stable_diffusion_pipeline.py:809-820) and the per-call ""
// encode (P:341-348), natively: token + position embedding, N pre-LN transformer layers with a causal mask, final
// LayerNorm; the caller takes last_hidden_state [B][77][hidden] (P:306, 819).
//
// SD-1.x: ViT-L/14 text tower (12 layers, 768 wide, 12 heads x 64, MLP 3072, quick-GELU); SD-2.x: OpenCLIP-H (23 used
// layers, 1024 wide, 16 heads x 64, MLP 4096, GELU) — both are configurations of this engine.  The linears run on the
// tcgen05 GEMM of sdw_gemm.cu (M = 77 B rows: one or two 128-row tiles), LayerNorm on sdw_norm.cu's kernel; the
// 77 x 77 causal attention per head and the embedding gather are small CUDA-core kernels here (13 GFLOP per prompt: the
// tower is a feed of the hot loop, not part of it).  State-dict names are transformers' `CLIPTextModel` keys.
#include "sdw_internal.h"
#include "sdw_ptx.cuh"

#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sdwalk.h"

namespace sdw {

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__global__ void clip_embed_kernel(const int32_t* __restrict__ ids, const __half* __restrict__ tok,
                                  const __half* __restrict__ pos, int P, int H, int vocab, __half* __restrict__ x) {
  const int row = blockIdx.x;  // b * P + p
  int id = ids[row];
  id = min(max(id, 0), vocab - 1);
  const __half2* t2 = reinterpret_cast<const __half2*>(tok + static_cast<int64_t>(id) * H);
  const __half2* p2 = reinterpret_cast<const __half2*>(pos + static_cast<int64_t>(row % P) * H);
  __half2* o2 = reinterpret_cast<__half2*>(x + static_cast<int64_t>(row) * H);
  for (int i = threadIdx.x; i < H / 2; i += blockDim.x) {
    const float2 a = __half22float2(t2[i]), b = __half22float2(p2[i]);
    o2[i] = __floats2half2_rn(a.x + b.x, a.y + b.y);
  }
}

// causal self-attention of one (head, sample): qkv [T][3H] (q | k | v column blocks, head h at columns h*64), out [T][H].
// 128 threads; K and V of the head staged in shared memory; warp w owns query rows w, w+4, ...
template <int MAXP>
__global__ void __launch_bounds__(128) clip_attn_kernel(const __half* __restrict__ qkv, int P, int H,
                                                        __half* __restrict__ out) {
  constexpr int D = 64;
  __shared__ __half ks[MAXP][D + 2];  // fp16 as stored; +2 keeps the per-lane rows on different banks
  __shared__ __half vs[MAXP][D];
  __shared__ float qs[4][D];
  const int head = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __half* base = qkv + static_cast<int64_t>(b) * P * 3 * H + head * D;
  for (int i = threadIdx.x; i < P * D; i += blockDim.x) {
    const int r = i / D, c = i % D;
    ks[r][c] = base[static_cast<int64_t>(r) * 3 * H + H + c];
    vs[r][c] = base[static_cast<int64_t>(r) * 3 * H + 2 * H + c];
  }
  __syncthreads();
  const float scale = 0.125f;  // 64^-1/2
  for (int i = warp; i < P; i += 4) {
    qs[warp][lane] = __half2float(base[static_cast<int64_t>(i) * 3 * H + lane]) * scale;
    qs[warp][lane + 32] = __half2float(base[static_cast<int64_t>(i) * 3 * H + lane + 32]) * scale;
    __syncwarp();
    float s[(MAXP + 31) / 32];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < (MAXP + 31) / 32; ++t) {
      const int j = t * 32 + lane;
      float acc = -INFINITY;
      if (j <= i) {  // causal mask: key j is visible to query i iff j <= i
        acc = 0.f;
#pragma unroll 16
        for (int c = 0; c < D; ++c) acc = fmaf(qs[warp][c], __half2float(ks[j][c]), acc);
      }
      s[t] = acc;
      m = fmaxf(m, acc);
    }
    m = warp_max(m);
    float l = 0.f;
#pragma unroll
    for (int t = 0; t < (MAXP + 31) / 32; ++t) {
      s[t] = (t * 32 + lane <= i) ? __expf(s[t] - m) : 0.f;
      l += s[t];
    }
    l = warp_sum(l);
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j <= i; ++j) {
      const float pj = __shfl_sync(0xffffffffu, s[j >> 5], j & 31);
      o0 = fmaf(pj, __half2float(vs[j][lane]), o0);
      o1 = fmaf(pj, __half2float(vs[j][lane + 32]), o1);
    }
    const float inv = 1.f / l;
    __half* orow = out + (static_cast<int64_t>(b) * P + i) * H + head * D;
    orow[lane] = __float2half_rn(o0 * inv);
    orow[lane + 32] = __float2half_rn(o1 * inv);
    __syncwarp();
  }
}

__global__ void clip_act_kernel(__half* __restrict__ x, int64_t n, int gelu_erf) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = __half2float(x[i]);
  const float y = gelu_erf ? 0.5f * v * (1.f + erff(v * 0.70710678118654752f)) : v / (1.f + __expf(-1.702f * v));
  x[i] = __float2half_rn(y);
}

// ---------------------------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------------------------
struct ClipParam {
  void* dst = nullptr;   // packed fp16 weight / fp32 vector / raw fp16 table
  int64_t numel = 0;
  int kind = 0;          // 0: fp32 vector, 1: linear weight [N][K] -> pack_weight at row offset, 2: raw fp16 copy
  int N = 0, K = 0;
  bool loaded = false;
};

struct ClipLayer {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *bqkv, *bo, *b1, *b2;
  __half *wqkv, *wo, *w1, *w2;
};

struct ClipEngine {
  sdw_clip_config cfg;
  bool dry = true;
  uint8_t* base = nullptr;
  size_t used = 0, cap = 0;
  std::map<std::string, ClipParam> params;
  std::vector<std::string> order;
  std::vector<ClipLayer> layers;
  __half *tok = nullptr, *pos = nullptr;
  float *lnf_g = nullptr, *lnf_b = nullptr;
  __half *x0 = nullptr, *x1 = nullptr, *h = nullptr, *qkv = nullptr, *ff = nullptr;
  int32_t* ids = nullptr;

  template <typename T>
  T* take(size_t n) {
    used = (used + 255) & ~size_t(255);
    T* p = dry ? nullptr : reinterpret_cast<T*>(base + used);
    used += n * sizeof(T);
    return p;
  }
  void reg(const std::string& name, void* dst, int64_t numel, int kind, int N = 0, int K = 0) {
    ClipParam p;
    p.dst = dst; p.numel = numel; p.kind = kind; p.N = N; p.K = K;
    if (!params.count(name)) order.push_back(name);
    params[name] = p;
  }
  void layout() {
    const sdw_clip_config& c = cfg;
    const int H = c.hidden, I = c.intermediate;
    used = 0;
    params.clear();
    order.clear();
    layers.assign(c.layers, ClipLayer{});
    tok = take<__half>(static_cast<size_t>(c.vocab) * H);
    pos = take<__half>(static_cast<size_t>(c.max_positions) * H);
    reg("text_model.embeddings.token_embedding.weight", tok, static_cast<int64_t>(c.vocab) * H, 2);
    reg("text_model.embeddings.position_embedding.weight", pos, static_cast<int64_t>(c.max_positions) * H, 2);
    for (int i = 0; i < c.layers; ++i) {
      ClipLayer& L = layers[i];
      const std::string p = "text_model.encoder.layers." + std::to_string(i) + ".";
      L.ln1_g = take<float>(H); L.ln1_b = take<float>(H); L.ln2_g = take<float>(H); L.ln2_b = take<float>(H);
      L.bqkv = take<float>(3 * H); L.bo = take<float>(H); L.b1 = take<float>(I); L.b2 = take<float>(H);
      L.wqkv = take<__half>(static_cast<size_t>(3) * H * H);
      L.wo = take<__half>(static_cast<size_t>(H) * H);
      L.w1 = take<__half>(static_cast<size_t>(I) * H);
      L.w2 = take<__half>(static_cast<size_t>(H) * I);
      reg(p + "layer_norm1.weight", L.ln1_g, H, 0); reg(p + "layer_norm1.bias", L.ln1_b, H, 0);
      reg(p + "layer_norm2.weight", L.ln2_g, H, 0); reg(p + "layer_norm2.bias", L.ln2_b, H, 0);
      const char* qkvn[3] = {"q_proj", "k_proj", "v_proj"};
      for (int k = 0; k < 3; ++k) {
        reg(p + "self_attn." + qkvn[k] + ".weight", dry ? nullptr : L.wqkv + static_cast<size_t>(k) * H * H,
            static_cast<int64_t>(H) * H, 1, H, H);
        reg(p + "self_attn." + qkvn[k] + ".bias", dry ? nullptr : L.bqkv + k * H, H, 0);
      }
      reg(p + "self_attn.out_proj.weight", L.wo, static_cast<int64_t>(H) * H, 1, H, H);
      reg(p + "self_attn.out_proj.bias", L.bo, H, 0);
      reg(p + "mlp.fc1.weight", L.w1, static_cast<int64_t>(I) * H, 1, I, H); reg(p + "mlp.fc1.bias", L.b1, I, 0);
      reg(p + "mlp.fc2.weight", L.w2, static_cast<int64_t>(H) * I, 1, H, I); reg(p + "mlp.fc2.bias", L.b2, H, 0);
    }
    lnf_g = take<float>(H); lnf_b = take<float>(H);
    reg("text_model.final_layer_norm.weight", lnf_g, H, 0);
    reg("text_model.final_layer_norm.bias", lnf_b, H, 0);
    const size_t T = static_cast<size_t>(c.max_batch) * c.max_positions;
    x0 = take<__half>(T * H); x1 = take<__half>(T * H); h = take<__half>(T * H);
    qkv = take<__half>(T * 3 * H); ff = take<__half>(T * I);
    ids = take<int32_t>(T);
  }
};

static int clip_linear(const __half* a, int64_t T, int K, const __half* w, int N, const float* bias, const __half* resid,
                       __half* out, cudaStream_t st) {
  GemmDesc d;
  d.A = a; d.C = K; d.W = static_cast<int>(T); d.H = 1; d.B = 1; d.sW = K;
  d.Wt = w; d.N = N; d.bias = bias; d.resid = resid; d.ldr = N; d.out = out; d.ldc = N;
  GemmLaunch L;
  if (int e = plan_gemm(d, &L)) return e;
  return launch_gemm(L, st);
}

}  // namespace sdw

using namespace sdw;

extern "C" {

int sdw_clip_create(const sdw_clip_config* cfg, sdw_clip** out) {
  SDW_REQUIRE(cfg && out, "null");
  SDW_REQUIRE(cfg->hidden % 64 == 0 && cfg->hidden == cfg->heads * 64, "CLIP text towers here have 64-wide heads");
  SDW_REQUIRE(cfg->intermediate % 64 == 0 && cfg->layers >= 1 && cfg->vocab >= 1, "bad CLIP configuration");
  SDW_REQUIRE(cfg->max_positions >= 1 && cfg->max_positions <= 96, "at most 96 positions (77 in every SD checkpoint)");
  SDW_REQUIRE(cfg->max_batch >= 1, "max_batch");
  ClipEngine* E = new ClipEngine();
  E->cfg = *cfg;
  E->dry = true;
  E->layout();
  E->cap = E->used;
  *out = reinterpret_cast<sdw_clip*>(E);
  return 0;
}

void sdw_clip_destroy(sdw_clip* e) { delete reinterpret_cast<ClipEngine*>(e); }

int sdw_clip_arena_bytes(const sdw_clip* e, uint64_t* bytes) {
  const ClipEngine* E = reinterpret_cast<const ClipEngine*>(e);
  SDW_REQUIRE(E && bytes, "null");
  *bytes = E->cap + 256;
  return 0;
}

int sdw_clip_bind(sdw_clip* e, void* arena, uint64_t bytes) {
  ClipEngine* E = reinterpret_cast<ClipEngine*>(e);
  SDW_REQUIRE(E && arena, "null");
  SDW_REQUIRE(bytes >= E->cap + 256, "arena too small");
  SDW_REQUIRE((reinterpret_cast<uintptr_t>(arena) & 255) == 0, "arena must be 256-byte aligned");
  E->base = static_cast<uint8_t*>(arena);
  E->dry = false;
  E->layout();
  return 0;
}

int sdw_clip_num_params(const sdw_clip* e) {
  const ClipEngine* E = reinterpret_cast<const ClipEngine*>(e);
  return E ? static_cast<int>(E->order.size()) : 0;
}

int sdw_clip_param_info(const sdw_clip* e, int index, const char** name, int64_t* numel) {
  const ClipEngine* E = reinterpret_cast<const ClipEngine*>(e);
  SDW_REQUIRE(E && index >= 0 && index < static_cast<int>(E->order.size()), "bad index");
  if (name) *name = E->order[index].c_str();
  if (numel) *numel = E->params.at(E->order[index]).numel;
  return 0;
}

int sdw_clip_load_param(sdw_clip* e, const char* name, const void* data_f16, int64_t numel, void* stream) {
  ClipEngine* E = reinterpret_cast<ClipEngine*>(e);
  SDW_REQUIRE(E && name && data_f16 && !E->dry, "null / engine not bound");
  auto it = E->params.find(name);
  SDW_REQUIRE(it != E->params.end(), "unknown CLIP parameter");
  ClipParam& p = it->second;
  SDW_REQUIRE(numel == p.numel, "CLIP parameter size mismatch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p.kind == 0) {
    if (int rc = half_to_float(static_cast<const __half*>(data_f16), static_cast<float*>(p.dst), numel, 0, st)) return rc;
  } else if (p.kind == 1) {
    if (int rc = pack_weight(data_f16, p.N, p.K, 1, 1, 0, p.dst, st)) return rc;
  } else {
    SDW_CUDA_OK(cudaMemcpyAsync(p.dst, data_f16, static_cast<size_t>(numel) * 2, cudaMemcpyDeviceToDevice, st));
  }
  p.loaded = true;
  return 0;
}

int sdw_clip_missing_params(const sdw_clip* e, const char** first_missing) {
  const ClipEngine* E = reinterpret_cast<const ClipEngine*>(e);
  if (!E) return -1;
  int n = 0;
  for (auto& name : E->order)
    if (!E->params.at(name).loaded) {
      if (n == 0 && first_missing) *first_missing = name.c_str();
      ++n;
    }
  return n;
}

int sdw_clip_forward(sdw_clip* e, const int32_t* ids, int B, void* out_f16, void* stream) {
  ClipEngine* E = reinterpret_cast<ClipEngine*>(e);
  SDW_REQUIRE(E && ids && out_f16 && !E->dry, "null / engine not bound");
  const sdw_clip_config& c = E->cfg;
  SDW_REQUIRE(B >= 1 && B <= c.max_batch, "batch exceeds max_batch");
  const char* missing = nullptr;
  SDW_REQUIRE(sdw_clip_missing_params(e, &missing) == 0, "CLIP parameters not loaded");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int P = c.max_positions, H = c.hidden, I = c.intermediate;
  const int64_t T = static_cast<int64_t>(B) * P;
  clip_embed_kernel<<<static_cast<unsigned>(T), 128, 0, st>>>(ids, E->tok, E->pos, P, H, c.vocab, E->x0);
  SDW_CUDA_OK(cudaGetLastError());
  __half *x = E->x0, *y = E->x1;
  for (int i = 0; i < c.layers; ++i) {
    const ClipLayer& L = E->layers[i];
    if (int rc = layernorm(x, H, T, H, L.ln1_g, L.ln1_b, c.eps, E->h, H, st)) return rc;
    if (int rc = clip_linear(E->h, T, H, L.wqkv, 3 * H, L.bqkv, nullptr, E->qkv, st)) return rc;
    clip_attn_kernel<96><<<dim3(c.heads, B), 128, 0, st>>>(E->qkv, P, H, E->h);
    SDW_CUDA_OK(cudaGetLastError());
    if (int rc = clip_linear(E->h, T, H, L.wo, H, L.bo, x, y, st)) return rc;
    std::swap(x, y);
    if (int rc = layernorm(x, H, T, H, L.ln2_g, L.ln2_b, c.eps, E->h, H, st)) return rc;
    if (int rc = clip_linear(E->h, T, H, L.w1, I, L.b1, nullptr, E->ff, st)) return rc;
    clip_act_kernel<<<static_cast<unsigned>((T * I + 255) / 256), 256, 0, st>>>(E->ff, T * I, c.act_gelu_erf);
    SDW_CUDA_OK(cudaGetLastError());
    if (int rc = clip_linear(E->ff, T, I, L.w2, H, L.b2, x, y, st)) return rc;
    std::swap(x, y);
  }
  return layernorm(x, H, T, H, E->lnf_g, E->lnf_b, c.eps, static_cast<__half*>(out_f16), H, st);
}

}  // extern "C"
