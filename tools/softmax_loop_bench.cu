// Instruction-mix microbenchmark for the attention softmax inner loop (no TMEM, no MMA): how many cycles does ONE
// softmax warp (or two sharing a scheduler) need per 32-column chunk of a score row, for several instruction mixes?
// MUFU.EX2 is 16 / clk / SM = 8 cycles per warp instruction per scheduler, so a 32-column chunk costs >= 256 cycles of
// the XU pipe per warp; the question is how close real mixes get with 1 and 2 warps per scheduler.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_loop_bench tools/softmax_loop_bench.cu
//
// Scores come from shared memory (LDS.128, stands in for tcgen05.ld), P goes back to shared memory (STS.128, stands in
// for tcgen05.st).  MODE: 0 pure MUFU | 1 shipped mix (FFMA2, 2 MUFU, FADD2 x2 acc, F2FP, FMNMX3) | 2 no running max,
// 4 sum accumulators | 3/4/5 = mode 2 with 1/4, 3/8, 1/2 of the exponentials as a packed degree-3 polynomial on the
// FMA pipe | 6 = mode 2 with 1/4 poly and the max kept.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t add2_rm(uint64_t a, uint64_t b) { uint64_t r; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

// 2^t for two values on the FMA pipe: floor via round-down magic add, degree-3 minimax on the fraction, exponent
// spliced in with one integer multiply-add per value
__device__ __forceinline__ void ex2_poly2(float t0, float t1, float& e0, float& e1) {
  t0 = fmaxf(t0, -126.f);
  t1 = fmaxf(t1, -126.f);
  const uint64_t magic = pk2(12582912.f, 12582912.f);
  const uint64_t t = pk2(t0, t1);
  const uint64_t xr = add2_rm(t, magic);
  const uint64_t f = sub2(t, sub2(xr, magic));
  uint64_t q = fma2(f, pk2(0.0780244991f, 0.0780244991f), pk2(0.2260671854f, 0.2260671854f));
  q = fma2(q, f, pk2(0.6958335042f, 0.6958335042f));
  q = fma2(q, f, pk2(0.9999251962f, 0.9999251962f));
  float q0, q1, r0, r1;
  upk2(q, q0, q1);
  upk2(xr, r0, r1);
  e0 = __int_as_float(__float_as_int(r0) * 8388608 + __float_as_int(q0));
  e1 = __int_as_float(__float_as_int(r1) * 8388608 + __float_as_int(q1));
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, long long* cyc, int iters) {
  extern __shared__ uint8_t smem[];
  float4* s_in = reinterpret_cast<float4*>(smem) + threadIdx.x;                      // [8][threads] float4: conflict-free
  uint4* s_out = reinterpret_cast<uint4*>(smem + blockDim.x * 128) + threadIdx.x;      // [4][threads] uint4
  for (int i = 0; i < 8; ++i) s_in[i * blockDim.x] = make_float4(-0.01f * (threadIdx.x & 31) - i, -1.f - i, -2.f - 0.1f * i, -3.f);
  __syncthreads();
  const float sl2 = 0.228f, mb = 0.5f;
  const uint64_t sl2_2 = pk2(sl2, sl2), nmb_2 = pk2(-mb, -mb);
  uint64_t sm[4] = {pk2(0.f, 0.f), pk2(0.f, 0.f), pk2(0.f, 0.f), pk2(0.f, 0.f)};
  float mx0 = -1e30f, mx1 = -1e30f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float x[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(s_in + i * blockDim.x))));
      x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
    }
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float e0, e1;
      if (MODE == 0) {
        e0 = ex2f(x[i]);
        e1 = ex2f(x[i + 1]);
        pk[i >> 1] = __float_as_uint(e0) ^ __float_as_uint(e1);
        continue;
      }
      if (MODE == 1 || MODE == 6) { mx0 = fmaxf(mx0, x[i]); mx1 = fmaxf(mx1, x[i + 1]); }
      float t0_, t1_;
      upk2(fma2(pk2(x[i], x[i + 1]), sl2_2, nmb_2), t0_, t1_);
      const int pair = i >> 1;
      const bool poly = (MODE == 3 || MODE == 6) ? (pair % 4 == 3) : (MODE == 4) ? (pair % 8 == 2 || pair % 8 == 5 || pair % 8 == 7)
                                                                   : (MODE == 5) ? (pair % 2 == 1) : false;
      if (poly) ex2_poly2(t0_, t1_, e0, e1);
      else { e0 = ex2f(t0_); e1 = ex2f(t1_); }
      const int a = (MODE == 1) ? (pair & 1) : (pair & 3);
      sm[a] = add2(sm[a], pk2(e0, e1));
      pk[pair] = pack_h2(e0, e1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(s_out + i * blockDim.x))),
                   "r"(pk[4 * i]), "r"(pk[4 * i + 1]), "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3]) : "memory");
  }
  const long long t1 = clock64();
  float s = mx0 + mx1;
  for (int i = 0; i < 4; ++i) { float a, b; upk2(sm[i], a, b); s += a + b; }
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_sched) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096, threads = 128 * warps_per_sched;
  const size_t smem = threads * (128 + 64);
  k<MODE><<<148, threads, smem>>>(out, cyc, 64);
  k<MODE><<<148, threads, smem>>>(out, cyc, iters);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  if (cudaDeviceSynchronize() != cudaSuccess) printf("CUDA error\n");
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += double(h[i]) / iters / 148;
  printf("%-52s warps/scheduler %d : %7.1f cycles per 32-column chunk per warp-set (XU floor %d)\n", name, warps_per_sched, avg,
         MODE == 0 ? 256 * warps_per_sched : 0);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w = 1; w <= 2; ++w) {
    run<0>("0 pure MUFU.EX2 x32", w);
    run<1>("1 shipped mix (max, 2 sum acc)", w);
    run<2>("2 no max, 4 sum acc", w);
    run<3>("3 no max, 1/4 poly", w);
    run<4>("4 no max, 3/8 poly", w);
    run<5>("5 no max, 1/2 poly", w);
    run<6>("6 max kept, 1/4 poly", w);
  }
  return 0;
}
