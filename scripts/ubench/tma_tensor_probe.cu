// Probe of the TMA tensor copies the fused backward uses for its g' workspace (run on the GPU box):
//   1. tensor store of a [9 planes][32 px] shared-memory box into the 4-d tensor [rows][9][9][W], no swizzle
//   2. the same through CU_TENSOR_MAP_SWIZZLE_128B from a box that does NOT start on a 1024-byte boundary: is the
//      16-byte chunk XOR-ed with bits 7..9 of the shared-memory ADDRESS (key = absolute 128-byte row) or of the row
//      index inside the box?
//   3. tensor load of a [9 dyi][32 px] box starting at a negative / unaligned x (zero fill, shift)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_tensor_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void store_kernel(const __grid_constant__ CUtensorMap tm, int mode, int slot_off_rows, int x0, int dyi, int row) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* box = reinterpret_cast<float*>(smem + slot_off_rows * 128);
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * 32; i += blockDim.x) {
    const int pl = i / 32, px = i % 32;
    int key = 0;
    if (mode == 1) key = (slot_off_rows + pl) & 7;   // absolute address bits 7..9
    if (mode == 2) key = pl & 7;                     // row inside the box
    box[pl * 32 + ((((px >> 2) ^ key) << 2) | (px & 3))] = 1000.f * pl + px;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(&tm)),
                 "r"(smem_u32(box)), "r"(x0), "r"(0), "r"(dyi), "r"(row)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

__global__ void load_kernel(const __grid_constant__ CUtensorMap tm, float* out, int x, int dxi, int row) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* box = reinterpret_cast<float*>(smem + 1152);  // 128-byte aligned, not 1024
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(9 * 32 * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(box)),
                 "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(bar)), "r"(x), "r"(dxi), "r"(0), "r"(row)
                 : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
  }
  for (int i = tid; i < 9 * 32; i += blockDim.x) out[i] = box[i];
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  const bool only_load = argc > 1;
  const int load_x = argc > 1 ? atoi(argv[1]) : 0;
  const int box_mode = argc > 2 ? atoi(argv[2]) : 0;
  const int W = 40, WP = 40, ROWS = 6;
  float* ws;
  const size_t n = (size_t)ROWS * 81 * WP;
  CK(cudaMalloc(&ws, n * 4));
  cuuint64_t dims[4] = {(cuuint64_t)W, 9, 9, ROWS};
  cuuint64_t strides[3] = {WP * 4ull, WP * 4ull * 9, WP * 4ull * 81};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  cuuint32_t box_store[4] = {32, 9, 1, 1}, box_load[4] = {32, 1, 9, 1};
  if (box_mode == 1) { box_load[1] = 9; box_load[2] = 1; }
  CUtensorMap tm_none, tm_swz, tm_load;
  CUresult r;
  cuInit(0);
  r = cuTensorMapEncodeTiled(&tm_none, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ws, dims, strides, box_store, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode store/none: %d\n", (int)r);
  r = cuTensorMapEncodeTiled(&tm_swz, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ws, dims, strides, box_store, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode store/128B: %d\n", (int)r);
  r = cuTensorMapEncodeTiled(&tm_load, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ws, dims, strides, box_load, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode load/none: %d\n", (int)r);
  std::vector<float> h(n);
  auto check_store = [&](const char* name, int x0, int dyi, int row) {
    cudaMemcpy(h.data(), ws, n * 4, cudaMemcpyDeviceToHost);
    int bad = 0, wrote = 0;
    for (int pl = 0; pl < 9; ++pl)
      for (int px = 0; px < 32; ++px) {
        const int x = x0 + px;
        if (x >= W) continue;
        const float v = h[((size_t)(row * 9 + dyi) * 9 + pl) * WP + x];
        if (v != 1000.f * pl + px) ++bad;
        ++wrote;
      }
    int stray = 0;
    for (size_t i = 0; i < n; ++i) if (h[i] != -1.f) ++stray;
    printf("%-40s: %d of %d elements wrong, %d elements touched in total\n", name, bad, wrote, stray);
  };
  for (int mode = 0; mode < 3 && !only_load; ++mode) {
    for (int off = 0; off <= 9; off += 9) {
      if (mode == 0 && off) continue;
      CK(cudaMemset(ws, 0xbf, n * 4));  // -1.5-ish pattern; use explicit fill instead
      std::vector<float> f(n, -1.f);
      cudaMemcpy(ws, f.data(), n * 4, cudaMemcpyHostToDevice);
      store_kernel<<<1, 128, 8192>>>(mode == 0 ? tm_none : tm_swz, mode, off, 8, 3, 2);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("store mode %d off %d: %s\n", mode, off, cudaGetErrorString(e)); return 1; }
      char name[96];
      snprintf(name, sizeof name, "store mode %d (0 none,1 abs,2 rel) box at row %d", mode, off);
      check_store(name, 8, 3, 2);
    }
  }
  // partial box at the right edge (x0 = 32, W = 40): clipped?
  if (!only_load) {
    std::vector<float> f(n, -1.f);
    cudaMemcpy(ws, f.data(), n * 4, cudaMemcpyHostToDevice);
    store_kernel<<<1, 128, 8192>>>(tm_none, 0, 0, 32, 1, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("edge store: %s\n", cudaGetErrorString(e));
    check_store("edge store x0=32 (8 valid px)", 32, 1, 0);
  }
  // shifted load
  {
    std::vector<float> f(n);
    for (size_t i = 0; i < n; ++i) f[i] = (float)i;
    cudaMemcpy(ws, f.data(), n * 4, cudaMemcpyHostToDevice);
    float* out;
    CK(cudaMalloc(&out, 9 * 32 * 4));
    for (int x : {load_x}) {
      load_kernel<<<1, 128, 8192>>>(tm_load, out, x, box_mode == 1 ? 0 : 7, 4);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("load x=%d: %s\n", x, cudaGetErrorString(e)); return 1; }
      std::vector<float> o(9 * 32);
      cudaMemcpy(o.data(), out, 9 * 32 * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int dy = 0; dy < 9; ++dy)
        for (int px = 0; px < 32; ++px) {
          const int xs = x + px;
          const float want = (xs >= 0 && xs < W) ? (box_mode == 1 ? (float)(((size_t)(4 * 9 + 0) * 9 + dy) * WP + xs) : (float)(((size_t)(4 * 9 + dy) * 9 + 7) * WP + xs)) : 0.f;
          if (o[dy * 32 + px] != want) ++bad;
        }
      printf("load box at x=%d: %d of 288 wrong\n", x, bad);
    }
  }
  return 0;
}
