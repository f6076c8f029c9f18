// Pointer-chase latency: L2-resident vs DRAM, plus a batched-gather latency probe (20 independent LDG.128 per lane,
// like the producer) on an otherwise idle GPU.
#include <cstdio>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
__global__ void chase(const unsigned* p, int n, unsigned* out, long long* cyc) {
  unsigned i = 0;
  for (int k = 0; k < 1024; ++k) i = p[i];  // warm
  long long t0 = clock64();
  for (int k = 0; k < n; ++k) i = p[i];
  long long t1 = clock64();
  *out = i; *cyc = (t1 - t0) / n;
}
__global__ void batch(const float4* src, size_t n4, int iters, float4* out, long long* cyc, int nload) {
  // every lane: nload independent 16-B loads at pseudo-random 128-B lines, then use all
  const int lane = threadIdx.x & 31;
  unsigned s = blockIdx.x * 7919u + (threadIdx.x >> 5) * 104729u + 12345u;
  float4 acc = make_float4(0, 0, 0, 0);
  long long tot = 0;
  for (int it = 0; it < iters; ++it) {
    float4 v[20];
    long long t0 = clock64();
#pragma unroll
    for (int j = 0; j < 20; ++j) {
      s = s * 1664525u + 1013904223u;
      size_t line = (size_t)(s >> 4) % (n4 / 8);
      if (j < nload) v[j] = __ldg(src + line * 8 + (lane & 7));
    }
#pragma unroll
    for (int j = 0; j < 20; ++j) if (j < nload) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    long long t1 = clock64();
    tot += t1 - t0;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = tot / iters;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  for (size_t mb : {16, 64, 2048}) {
    size_t n = mb * 1024 * 1024 / 4;
    size_t lines = n / 32;
    std::vector<unsigned> h(n, 0);
    std::vector<unsigned> perm(lines);
    for (size_t i = 0; i < lines; ++i) perm[i] = i;
    std::mt19937 g(1); std::shuffle(perm.begin() + 1, perm.end(), g);
    for (size_t i = 0; i < lines; ++i) h[(size_t)perm[i] * 32] = perm[(i + 1) % lines] * 32;
    unsigned* d; cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    unsigned* o; long long* c; cudaMalloc(&o, 4); cudaMalloc(&c, 8);
    chase<<<1, 1>>>(d, 20000, o, c); chase<<<1, 1>>>(d, 20000, o, c);
    long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
    printf("pointer chase over %zu MB: %lld cycles per load\n", mb, hc);
    for (int nload : {1, 4, 20})
      for (int blocks : {1, 148}) {
        float4* out; cudaMalloc(&out, 148 * 128 * 16);
        batch<<<blocks, 128>>>((const float4*)d, n / 4, 2000, out, c, nload);
        cudaDeviceSynchronize();
        cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
        printf("   batch of %2d LDG.128 per lane, %3d CTAs x 4 warps: %lld cycles per batch\n", nload, blocks, hc);
        cudaFree(out);
      }
    cudaFree(d);
  }
  return 0;
}
