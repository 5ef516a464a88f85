// Throughput of the exponential forms the attention softmax could use (per SM, per clock):
//   ex2.approx.ftz.f32 (one value / MUFU op)  vs  ex2.approx.ftz.f16x2 (two values / instruction).
// nvcc -arch=sm_100a -O3 -o mufu_bench tools/mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  unsigned h[8];
  for (int i = 0; i < 8; ++i) {
    a[i] = -0.001f * (threadIdx.x + i);
    h[i] = 0xb000b000u + threadIdx.x + i;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16 %0, %0;" : "+h"(*reinterpret_cast<unsigned short*>(&h[i])));
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int per_instr) {
  float* out;
  cudaMalloc(&out, 4);
  const int iters = 4096, blocks = 148 * 2, threads = 1024;
  k<MODE><<<blocks, threads>>>(out, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double instr = double(blocks) * threads * iters * 8;
  printf("%-28s %8.3f ms  %7.2f Ginstr-lanes/s  %7.2f Gvalues/s\n", name, ms, instr / ms / 1e6, instr * per_instr / ms / 1e6);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.f16x2", 2);
  run<2>("ex2.approx.ftz.f16", 1);
  return 0;
}
