// Micro-benchmark: FP32 FMA throughput per SM with FFMA vs FFMA2 (fma.rn.f32x2), 1/2/4 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float2 acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 a = make_float2(s, s * 1.0001f), b = make_float2(1.0f - s, 0.999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (MODE == 0) {
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].x) : "f"(a.x), "f"(b.x));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].y) : "f"(a.y), "f"(b.y));
      } else {
        ffma2(acc[i], a, b);
      }
    }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) r += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int threads : {128, 256, 512, 1024}) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148, threads>>>(out, iters, 0.5f); else k<1><<<148, threads>>>(out, iters, 0.5f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * threads * iters * 64.0;
      printf("%s threads/SM=%4d: %.3f ms  %.1f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", mode ? "FFMA2" : "FFMA ", threads, ms,
             2 * fma / ms / 1e9, fma / 148 / (ms * 1e-3 * 1.965e9));
    }
  return 0;
}
