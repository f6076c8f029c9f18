#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ void ffma2(u64& d, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
#ifndef NPX
#define NPX 4
#endif
#ifndef KVV
#define KVV 8
#endif
extern "C" __global__ void __launch_bounds__(256, 1) k(const ulonglong2* __restrict__ gin, float* out, int n) {
  extern __shared__ ulonglong2 sm[];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = gin[i];
  __syncthreads();
  const ulonglong2* ap = sm + (threadIdx.x & 7) * 5;
  const ulonglong2* bp = sm + 512 + (threadIdx.x >> 3) * 3;
  u64 acc[NPX][9];
#pragma unroll
  for (int p = 0; p < NPX; ++p)
#pragma unroll
    for (int d = 0; d < 9; ++d) acc[p][d] = 0ull;
  for (int rep = 0; rep < n; ++rep) {
#pragma unroll 1
    for (int kv = 0; kv < KVV; ++kv) {
      ulonglong2 a[NPX];
#pragma unroll
      for (int p = 0; p < NPX; ++p) a[p] = ap[kv + p * KVV];
#pragma unroll
      for (int j = 0; j < NPX + 8; ++j) {
        const ulonglong2 b = bp[kv + j * KVV];
#pragma unroll
        for (int p = 0; p < NPX; ++p) { const int d = j - p; if (d >= 0 && d < 9) ffma2(acc[p][d], a[p].x, b.x); }
#pragma unroll
        for (int p = 0; p < NPX; ++p) { const int d = j - p; if (d >= 0 && d < 9) ffma2(acc[p][d], a[p].y, b.y); }
      }
    }
  }
  float r = 0;
#pragma unroll
  for (int p = 0; p < NPX; ++p)
#pragma unroll
    for (int d = 0; d < 9; ++d) { float2 v = *reinterpret_cast<float2*>(&acc[p][d]); r += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
#include <cstdio>
int main() {
  ulonglong2* gin; float* out;
  cudaMalloc(&gin, 2048 * 16); cudaMemset(gin, 0, 2048 * 16); cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int n = 2000;
  for (int threads : {128, 256, 384}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k<<<148, threads, 40960>>>(gin, out, n);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ffma2_per_warp = (double)n * KVV * NPX * 9 * 2;
    double cyc = ms * 1e-3 * 1.965e9;
    printf("NPX=%d threads=%d: %.3f ms, %.2f cycles per FFMA2 per SMSP (warps/SMSP=%d), FMA/clk/SM=%.1f  err=%s\n", NPX, threads, ms,
           cyc / (ffma2_per_warp * (threads / 128)), threads / 128, ffma2_per_warp * 2 * 32 * (threads / 32) / cyc, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
