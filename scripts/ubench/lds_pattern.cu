// Micro-benchmark: shared-memory wavefronts per LDS.128 / LDS.64 for lane->address patterns with duplicates.
// Question answered: which lane arrangements let the crossbar merge duplicate addresses (broadcast) in one wavefront?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <string>

template <int VEC>  // 4 = LDS.128, 2 = LDS.64, 1 = LDS.32
__global__ void __launch_bounds__(512) k(const int* __restrict__ pat, float* out, int iters, long long* cyc) {
  extern __shared__ float4 sm[];
  __shared__ unsigned long long tmin, tmax;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  if (threadIdx.x == 0) { tmin = ~0ull; tmax = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // every warp uses its own 512-byte-aligned window so that nothing can merge across warps
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + pat[lane] * 4 * VEC + (warp & 7) * 512;
  float a = 0, b = 0, c = 0, d = 0;
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      // results are overwritten, never consumed inside the loop: no ALU work competes with the loads
      if (VEC == 4) asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(base + (u & 3) * 8192));
      else if (VEC == 2) asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(base + (u & 3) * 8192));
      else asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(a) : "r"(base + (u & 3) * 8192));
    }
  }
  unsigned long long t1 = clock64();
  atomicMin(&tmin, t0); atomicMax(&tmax, t1);
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (long long)(tmax - tmin);
}

int main() {
  struct Pat { std::string name; int vec; int idx[32]; };
  std::vector<Pat> pats;
  auto add = [&](const char* n, int vec, auto f) { Pat p; p.name = n; p.vec = vec; for (int l = 0; l < 32; ++l) p.idx[l] = f(l); pats.push_back(p); };
  add("128: all same", 4, [](int l) { return 0; });
  add("128: l&7 (8 uniq, same in every quarter)", 4, [](int l) { return l & 7; });
  add("128: l>>1 (16 uniq, lane pairs)", 4, [](int l) { return l >> 1; });
  add("128: l>>2 (8 uniq, 2 per quarter)", 4, [](int l) { return l >> 2; });
  add("128: l>>3 (4 uniq, 1 per quarter)", 4, [](int l) { return l >> 3; });
  add("128: l (32 uniq)", 4, [](int l) { return l; });
  add("128: l&15 (16 uniq, halves same)", 4, [](int l) { return l & 15; });
  add("128: l&3 (4 uniq, in every quarter)", 4, [](int l) { return l & 3; });
  add("128: l&1 (2 uniq)", 4, [](int l) { return l & 1; });
  add("128: (l&1)+2*(l>>3) (8 uniq, 2/quarter, quarters differ)", 4, [](int l) { return (l & 1) + 2 * (l >> 3); });
  add("128: (l>>1&3)+4*(l>>3) (16 uniq, 4/quarter)", 4, [](int l) { return ((l >> 1) & 3) + 4 * (l >> 3); });
  add("128: (l&3)+4*(l>>3) (16 uniq, 4/quarter alt)", 4, [](int l) { return (l & 3) + 4 * (l >> 3); });
  add("128: (l&3)+4*(l>>4) (8 uniq, 4/quarter, halves differ)", 4, [](int l) { return (l & 3) + 4 * (l >> 4); });
  add("128: (l&7)+8*(l>>4) (16 uniq, 8/quarter, halves differ)", 4, [](int l) { return (l & 7) + 8 * (l >> 4); });
  add("128: 8 uniq stride 9 f4 same/quarter", 4, [](int l) { return (l & 7) * 9; });
  add("64: all same", 2, [](int l) { return 0; });
  add("64: l&15 (16 uniq, halves same)", 2, [](int l) { return l & 15; });
  add("64: l>>1 (16 uniq pairs)", 2, [](int l) { return l >> 1; });
  add("64: l (32 uniq)", 2, [](int l) { return l; });
  add("64: l&7", 2, [](int l) { return l & 7; });
  add("32: l", 1, [](int l) { return l; });
  add("32: l&7", 1, [](int l) { return l & 7; });
  int* dp; cudaMalloc(&dp, 32 * 4);
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  long long* dc; cudaMalloc(&dc, 8);
  const int iters = 2000, threads = 512;
  for (auto& p : pats) {
    cudaMemcpy(dp, p.idx, 128, cudaMemcpyHostToDevice);
    long long c = 0;
    for (int rep = 0; rep < 2; ++rep) {
      if (p.vec == 4) k<4><<<148, threads, 32768>>>(dp, out, iters, dc);
      else if (p.vec == 2) k<2><<<148, threads, 32768>>>(dp, out, iters, dc);
      else k<1><<<148, threads, 32768>>>(dp, out, iters, dc);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    const double n_inst = (double)iters * 16 * (threads / 32);  // warp-level LDS per SM
    printf("%-62s %.2f clk per warp-LDS\n", p.name.c_str(), c / n_inst);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
