// tcgen05.ld throughput per SM: how fast can the softmax / epilogue warps pull fp32 accumulators out of TMEM?
// One CTA per SM allocates 512 columns; W warps (4 = one per lane quarter, 8 = two per quarter) issue back-to-back
// 32x32b.x32 loads (4 KB per warp instruction).  Prints bytes / clk / SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bench tools/tmem_bench.cu && ./tools/tmem_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(int iters, int unroll_wait, long long* cycles, float* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      const uint32_t a = base + ((it * 4 + c) & 15) * 32;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(a));
      if (unroll_wait == 1 || c == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(r[0] ^ r[31]);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 12345.678f) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tptr));
}

int main() {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 148 * 8);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int warps : {1, 4, 8}) {
    for (int uw : {1, 0}) {
      k<<<148, warps * 32>>>(16, uw, cyc, sink);
      k<<<148, warps * 32>>>(iters, uw, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += h[i];
      avg /= 148;
      const double bytes = double(warps) * iters * 4 * 4096;
      printf("warps %d wait-%s: %.0f cycles, %.1f B/clk/SM, %.1f cycles per x32 load per warp  (%s)\n", warps,
             uw ? "each" : "every-4", avg, bytes / avg, avg / (iters * 4), cudaGetErrorString(e));
    }
  }
  return 0;
}
