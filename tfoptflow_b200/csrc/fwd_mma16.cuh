// Fused warp + cost-volume forward for 16-bit features (the reference's fp16 regime, model_pwcnet.py:390-394; bf16 as
// an extension), search_range 4, C in {32, 64}: the banded correlation of one c1 row against one warped row is run as
// a band GEMM on the tensor cores.
//
// DELIBERATE DEVIATION from "no tensor cores" (SURVEY.md 7.3-1 allows it for the 16-bit regime): at 16-bit element size
// the path needs ~139 TFLOP/s to run at HBM speed, twice what the CUDA cores have, and the fp32 kernels' operand
// widening made fp16 SLOWER than fp32.  fp16 x fp16 products are exact in fp32 and are accumulated in fp32, which is
// more accurate than the reference's all-fp16 products, i.e. inside the 1e-2 tolerance.  fp32 inputs never take this
// path (TF32 would break 1e-5).
//
// Per (c1 row y, displacement dy, block of 16 pixels): S[x, x'] = sum_c c1[y,x,c] * warped[y+dy, x', c] for the 24
// neighbours x' = x0-4 .. x0+19, as 3 x (C/16) `mma.sync.m16n8k16` (fp32 accumulate); the 9 diagonals dx = x' - x + 4
// in [0, 9) are the cost volume (9/24 of the products are used).  Both operands are K-contiguous ([pixel][channel]),
// exactly what `ldmatrix` wants, so the NHWC layout needs no transposition.
//   * CTA = 8 x 32 output pixels, 8 warps.  Phase 1: sampling coordinates of the 16 x 40 halo pixels; phase 2: the
//     bilinear gather (corners as 16-byte loads, blended in fp32, rounded ONCE to the 16-bit type like the reference's
//     warp output) into a shared-memory tile - the warped features never reach HBM; phase 3: per warp 2 (row, block)
//     units x 9 dy x 3 n-tiles of MMAs, A fragments (c1) straight from global memory and kept for all 9 dy; leaky-relu
//     and 1/C on the accumulator fragments, which are scattered into a [pixel][81] staging tile; phase 4: the staged
//     tile leaves as whole 16-byte vectors (32 pixels x 81 channels are contiguous in NHWC).
//   * 92 KB of shared memory: two CTAs per SM overlap each other's gather and MMA phases.
// That is the first variant (`fwd_mma16_kernel`, option mma_tile 0).  The default is the resident-halo variant further
// down (`fwd_mma16_band_kernel`, namespace band): taller or narrower tiles per configuration, the four bilinear corners
// kept packed until they are blended (16 loads in flight per thread), no tile-wide staging buffer - every warp stages
// and writes its own (row, 16-pixel block) units - and search_range 8 as well as 4.
#pragma once
#include "common.cuh"

namespace pwc {
namespace mma16 {

constexpr int R = 4, ND = 9, D = 81;
constexpr int TH = 8, TW = 32, BH = TH + 2 * R, BW = TW + 2 * R;   // 16 x 40 halo tile
constexpr int NTHREADS = 256;

template <int C>
struct Cfg {
  static constexpr int RS = C + 8;                    // halves per halo pixel (16-byte pad: ldmatrix rows hit 8 bank groups)
  static constexpr int B_BYTES = BH * BW * RS * 2;
  static constexpr int STAGE_BYTES = TH * TW * D * 2; // [pixel][81] 16-bit
  static constexpr int COORD_BYTES = BH * BW * 16;    // aliases the staging tile (dead before the first scatter)
  static_assert(COORD_BYTES <= STAGE_BYTES, "coords alias the staging tile");
  static constexpr int SMEM_BYTES = B_BYTES + STAGE_BYTES;
  static_assert((RS * 2) % 16 == 0 && ((RS * 2) / 16) % 2 == 1, "ldmatrix row stride");
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldmatrix_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 8 consecutive 16-bit channels <-> two float4
template <typename T>
__device__ __forceinline__ void ld8(const T* p, float4& lo, float4& hi) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const T* h = reinterpret_cast<const T*>(&u);
  lo = make_float4(ElemTraits<T>::to_f(h[0]), ElemTraits<T>::to_f(h[1]), ElemTraits<T>::to_f(h[2]), ElemTraits<T>::to_f(h[3]));
  hi = make_float4(ElemTraits<T>::to_f(h[4]), ElemTraits<T>::to_f(h[5]), ElemTraits<T>::to_f(h[6]), ElemTraits<T>::to_f(h[7]));
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float4& lo, const float4& hi) {
  uint4 u;
  T* h = reinterpret_cast<T*>(&u);
  h[0] = ElemTraits<T>::from_f(lo.x); h[1] = ElemTraits<T>::from_f(lo.y); h[2] = ElemTraits<T>::from_f(lo.z); h[3] = ElemTraits<T>::from_f(lo.w);
  h[4] = ElemTraits<T>::from_f(hi.x); h[5] = ElemTraits<T>::from_f(hi.y); h[6] = ElemTraits<T>::from_f(hi.z); h[7] = ElemTraits<T>::from_f(hi.w);
  return u;
}

template <int C, bool HAS_FLOW, typename T, typename TF>
__global__ void __launch_bounds__(NTHREADS, 2)
fwd_mma16_kernel(const T* __restrict__ c1, const T* __restrict__ c2, const TF* __restrict__ flow, float flow_scale,
                 T* __restrict__ out, int B, int H, int W, int64_t ops, int tiles_x, int tiles_y) {
  using K = Cfg<C>;
  static_assert(sizeof(T) == 2, "16-bit features only");
  extern __shared__ __align__(128) unsigned char smem[];
  T* sB = reinterpret_cast<T*>(smem);                       // [BH * BW][RS] warped halo tile
  T* sOut = reinterpret_cast<T*>(smem + K::B_BYTES);        // [TH * TW][81]
  int4* sCoord = reinterpret_cast<int4*>(smem + K::B_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t = blockIdx.x;
  const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
  const int x0 = tx * TW, y0 = ty * TH;

  // ---- phase 1: sampling coordinates of every halo pixel (w = 0: outside the image -> zeros) ----
  for (int i = tid; i < BH * BW; i += NTHREADS) {
    const int rb = i / BW, u = i - rb * BW;
    const int y = y0 - R + rb, x = x0 - R + u;
    int4 cd = make_int4(0, 0, 0, 0);
    if (y >= 0 && y < H && x >= 0 && x < W) {
      if (HAS_FLOW) {
        const SampleCoord s = sample_coord<T, TF>(flow, flow_scale, b, y, x, H, W);
        cd = make_int4(s.pix, __float_as_int(s.ax), __float_as_int(s.ay), 1);
      } else {
        cd = make_int4((b * H + y) * W + x, 0, 0, 2);
      }
    }
    sCoord[i] = cd;
  }
  __syncthreads();

  // ---- phase 2: warped halo tile, 8 channels (16 bytes) per item ----
  {
    constexpr int CH8 = C / 8, NB = BH * BW * CH8, FB = 2;
    const size_t rowstride = (size_t)W * C;
#pragma unroll 1
    for (int i0 = tid; i0 < NB; i0 += FB * NTHREADS) {
      float4 tl[FB][2], tr[FB][2], bl[FB][2], br[FB][2];
      int4 cd[FB];
#pragma unroll
      for (int j = 0; j < FB; ++j) {
        const int i = i0 + j * NTHREADS;
        const int p = i / CH8, k8 = i - p * CH8;
        cd[j] = i < NB ? sCoord[p] : make_int4(0, 0, 0, 0);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        tl[j][0] = tl[j][1] = tr[j][0] = tr[j][1] = bl[j][0] = bl[j][1] = br[j][0] = br[j][1] = z;
        if (cd[j].w != 0) {
          const T* base = c2 + (size_t)cd[j].x * C + 8 * k8;
          ld8<T>(base, tl[j][0], tl[j][1]);
          if (cd[j].w == 1) {
            ld8<T>(base + C, tr[j][0], tr[j][1]);
            ld8<T>(base + rowstride, bl[j][0], bl[j][1]);
            ld8<T>(base + rowstride + C, br[j][0], br[j][1]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < FB; ++j) {
        const int i = i0 + j * NTHREADS;
        if (i >= NB) continue;
        const int p = i / CH8, k8 = i - p * CH8;
        float4 lo = tl[j][0], hi = tl[j][1];
        if (cd[j].w == 1) {
          const float ax = __int_as_float(cd[j].y), ay = __int_as_float(cd[j].z);
          lo = bilerp4(tl[j][0], tr[j][0], bl[j][0], br[j][0], ax, ay);
          hi = bilerp4(tl[j][1], tr[j][1], bl[j][1], br[j][1], ax, ay);
        }
        *reinterpret_cast<uint4*>(sB + (size_t)p * K::RS + 8 * k8) = pack8<T>(lo, hi);  // one rounding, like the reference's warp
      }
    }
  }
  __syncthreads();

  // ---- phase 3: band GEMMs ----
  const int g = lane >> 2, tq = lane & 3;
  const float inv_c = 1.0f / (float)C;
#pragma unroll 1
  for (int un = warp; un < TH * (TW / 16); un += NTHREADS / 32) {
    const int row = un >> 1, blk = un & 1;
    const int y = y0 + row;
    // A fragments (c1 row block, 16 pixels x C channels) from global memory, kept for all 9 dy.
    // a0: (pixel g, k 2t..2t+1)  a1: (pixel g+8, same k)  a2: (pixel g, k+8)  a3: (pixel g+8, k+8)
    uint32_t afr[C / 16][4];
    {
      const int yc = min(y, H - 1);
      const int xa = min(x0 + 16 * blk + g, W - 1), xb = min(x0 + 16 * blk + g + 8, W - 1);
      const uint32_t* pa = reinterpret_cast<const uint32_t*>(c1 + (((size_t)b * H + yc) * W + xa) * C);
      const uint32_t* pb = reinterpret_cast<const uint32_t*>(c1 + (((size_t)b * H + yc) * W + xb) * C);
#pragma unroll
      for (int ks = 0; ks < C / 16; ++ks) {
        afr[ks][0] = __ldg(pa + ks * 8 + tq);
        afr[ks][1] = __ldg(pb + ks * 8 + tq);
        afr[ks][2] = __ldg(pa + ks * 8 + 4 + tq);
        afr[ks][3] = __ldg(pb + ks * 8 + 4 + tq);
      }
    }
    T* so = sOut + (size_t)(row * TW + 16 * blk) * D;
#pragma unroll 1
    for (int dyi = 0; dyi < ND; ++dyi) {
      // warped row (row + dyi) of the halo tile, halo columns 16 blk .. 16 blk + 23  <->  x' = x0 + 16 blk - 4 + n
      const T* brow = sB + (size_t)((row + dyi) * BW + 16 * blk) * K::RS;
      float acc[3][4];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < C / 16; ++ks) {
          // ldmatrix.x2: lanes 0-7 give the rows (pixels n) of the k-low 8x8 matrix, lanes 8-15 of the k-high one
          uint32_t b0, b1;
          const T* src = brow + (size_t)(8 * nt + (lane & 7)) * K::RS + 16 * ks + ((lane >> 3) & 1) * 8;
          ldmatrix_x2(b0, b1, smem_u32(src));
          mma16816<T>(acc[nt], afr[ks], b0, b1);
        }
      }
      // accumulator (pixel m, neighbour n): m = g + 8 h, n = 8 nt + 2 t + j  ->  dx index = n - m in [0, 9)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int h = e >> 1, j = e & 1;
          const int m = g + 8 * h, dxi = 8 * nt + 2 * tq + j - m;
          if (dxi >= 0 && dxi < ND) {
            const float v = acc[nt][e] * inv_c;
            so[m * D + dyi * ND + dxi] = ElemTraits<T>::from_f(fmaxf(0.1f * v, v));
          }
        }
    }
  }
  __syncthreads();

  // ---- phase 4: staged tile -> global ----
  const int nrows = min(TH, H - y0), ncols = min(TW, W - x0);
  const bool vec = ops == D && ncols == TW && ((((size_t)b * H + y0) * W + x0) * D * 2) % 16 == 0 &&
                   ((size_t)W * D * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec) {
    constexpr int V = TW * D * 2 / 16;  // 16-byte vectors per tile row
    for (int i = tid; i < nrows * V; i += NTHREADS) {
      const int row = i / V, e = i - row * V;
      uint4* dst = reinterpret_cast<uint4*>(out + (((size_t)b * H + y0 + row) * W + x0) * D);
      dst[e] = reinterpret_cast<const uint4*>(sOut + (size_t)row * TW * D)[e];
    }
  } else {
    for (int i = tid; i < nrows * TW * D; i += NTHREADS) {
      const int row = i / (TW * D), r2 = i - row * (TW * D), px = r2 / D, ch = r2 - px * D;
      if (px < ncols) out[(((size_t)b * H + y0 + row) * W + x0 + px) * ops + ch] = sOut[i];
    }
  }
}

// ---- search_range 8 (BASELINE configs[4]: 289 channels), C in {32, 64} -------------------------------------------------
// Same band GEMM, 32 neighbours per 16-pixel block (17 of 32 diagonals used).  The +-8 halo makes small tiles
// expensive (an 8 x 32 tile would gather 4.5 halo pixels per output pixel), so the CTA keeps the warped halo of a
// 16 x 32 tile resident (32 x 48 pixels, 120 KB) and there is no tile-wide staging buffer: each warp stages one
// (row, 16-pixel block) unit - 16 pixels x 289 channels = 9 248 contiguous bytes of the output - in its own 9 KB of
// shared memory and writes it out itself, as 16-byte vectors, before it takes the next unit.
#ifndef PWC_BAND_FB
#define PWC_BAND_FB 4    /* gather items (4 corner loads each, kept packed) in flight per thread when two CTAs share an SM */
#endif
#ifndef PWC_BAND_FB1
#define PWC_BAND_FB1 6   /* the same for the one-CTA-per-SM configurations (r = 8, C = 96) */
#endif
#ifndef PWC_BAND_DYU
#define PWC_BAND_DYU 3   /* unroll of the displacement-row loop of the MMA phase */
#endif
namespace band {
// R = 4: 16-row (C = 32) or 8-row tiles of 32 pixels, two CTAs per SM up to C = 64.  R = 8, C = 32: 16-row tile, 8 warps
// (120 KB halo + 8 x 9 KB staging); C = 64: the halo pixels are twice as wide, so an 8-row tile and 6 warps (162 KB +
// 6 x 9 KB); C = 96 / 128: 16-pixel-wide tiles of 6 / 4 rows, one unit per warp (the +-8 halo makes these expensive -
// 5.5 / 10 halo pixels per output pixel - but still cheaper than 289 x C multiply-adds per pixel on the CUDA cores).
template <int R_, int C_>
struct Cfg {
  static constexpr int R = R_, C = C_, RS = C + 8;
  static constexpr int TW = (R == 8 && C >= 96) ? 16 : 32;
  static constexpr int ND = 2 * R + 1, D = ND * ND, BW = TW + 2 * R, NTILES = (16 + 2 * R) / 8;
  static constexpr int TH = C == 32 ? 16 : (R == 4 || C == 64) ? 8 : (C == 96 ? 6 : 4), BH = TH + 2 * R;
  static constexpr int NWARPS = R == 4 ? 8 : (C == 32 ? 8 : (C == 128 ? 4 : 6)), NT = 32 * NWARPS;
  static constexpr int CTAS = (R == 4 && C <= 64) ? 2 : 1;
  static constexpr int B_BYTES = BH * BW * RS * 2;
  static constexpr int UNIT_BYTES = (16 * D * 2 + 15) / 16 * 16;    // one warp's staging slot: [16 px][D] 16-bit
  static constexpr int COORD_BYTES = BH * BW * 16;                  // aliases the staging slots (dead before the MMAs)
  static constexpr int STAGE_BYTES = NWARPS * UNIT_BYTES > COORD_BYTES ? NWARPS * UNIT_BYTES : COORD_BYTES;
  static constexpr int SMEM_BYTES = B_BYTES + STAGE_BYTES;
  static_assert((16 * D * 2) % 16 == 0 && SMEM_BYTES * CTAS <= 232448 - 1024 * CTAS, "shared memory");
  static_assert((RS * 2) % 16 == 0 && ((RS * 2) / 16) % 2 == 1, "ldmatrix row stride");
};
}  // namespace band

// ---- the three phases of a tile ---------------------------------------------------------------------------------
// phase 1: sampling coordinates of every halo pixel (w = 0: outside the image -> zeros); `tid` / `nthr` = the thread's
// index inside the group of threads that works on the phase
template <typename K, bool HAS_FLOW, typename T, typename TF>
__device__ __forceinline__ void band_coords(const TF* __restrict__ flow, float flow_scale, int b, int y0, int x0, int H, int W,
                                            int4* __restrict__ sCoord, int tid, int nthr) {
  for (int i = tid; i < K::BH * K::BW; i += nthr) {
    const int rb = i / K::BW, u = i - rb * K::BW;
    const int y = y0 - K::R + rb, x = x0 - K::R + u;
    int4 cd = make_int4(0, 0, 0, 0);
    if (y >= 0 && y < H && x >= 0 && x < W) {
      if (HAS_FLOW) {
        const SampleCoord s = sample_coord<T, TF>(flow, flow_scale, b, y, x, H, W);
        cd = make_int4(s.pix, __float_as_int(s.ax), __float_as_int(s.ay), 1);
      } else {
        cd = make_int4((b * H + y) * W + x, 0, 0, 2);
      }
    }
    sCoord[i] = cd;
  }
}

// phase 2: warped halo tile, 8 channels (16 bytes) per item, FB items (4 corner loads each) in flight per thread.  The
// corners stay packed (4 registers each) until they are blended, so 4 items = 16 loads fit the registers of a
// two-CTAs-per-SM kernel
template <typename T>
__device__ __forceinline__ void unpack8(const uint4& u, float4& lo, float4& hi) {
  const T* h = reinterpret_cast<const T*>(&u);
  lo = make_float4(ElemTraits<T>::to_f(h[0]), ElemTraits<T>::to_f(h[1]), ElemTraits<T>::to_f(h[2]), ElemTraits<T>::to_f(h[3]));
  hi = make_float4(ElemTraits<T>::to_f(h[4]), ElemTraits<T>::to_f(h[5]), ElemTraits<T>::to_f(h[6]), ElemTraits<T>::to_f(h[7]));
}
template <typename K, int FB, typename T>
__device__ __forceinline__ void band_gather(const T* __restrict__ c2, int W, const int4* __restrict__ sCoord,
                                            T* __restrict__ sB, int tid, int nthr) {
  constexpr int C = K::C, CH8 = C / 8, NB = K::BH * K::BW * CH8;
  const size_t rowstride = (size_t)W * C;
#pragma unroll 1
  for (int i0 = tid; i0 < NB; i0 += FB * nthr) {
    uint4 tl[FB], tr[FB], bl[FB], br[FB];
    int4 cd[FB];
#pragma unroll
    for (int j = 0; j < FB; ++j) {
      const int i = i0 + j * nthr;
      const int p = i / CH8, k8 = i - p * CH8;
      cd[j] = i < NB ? sCoord[p] : make_int4(0, 0, 0, 0);
      tl[j] = tr[j] = bl[j] = br[j] = make_uint4(0u, 0u, 0u, 0u);   // +0.0 in either 16-bit format
      if (cd[j].w != 0) {
        const T* base = c2 + (size_t)cd[j].x * C + 8 * k8;
        tl[j] = __ldg(reinterpret_cast<const uint4*>(base));
        if (cd[j].w == 1) {
          tr[j] = __ldg(reinterpret_cast<const uint4*>(base + C));
          bl[j] = __ldg(reinterpret_cast<const uint4*>(base + rowstride));
          br[j] = __ldg(reinterpret_cast<const uint4*>(base + rowstride + C));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < FB; ++j) {
      const int i = i0 + j * nthr;
      if (i >= NB) continue;
      const int p = i / CH8, k8 = i - p * CH8;
      uint4 v = tl[j];   // no warp: the features as they are
      if (cd[j].w == 1) {
        const float ax = __int_as_float(cd[j].y), ay = __int_as_float(cd[j].z);
        float4 a0, a1, b0, b1, c0, c1, d0, d1;
        unpack8<T>(tl[j], a0, a1); unpack8<T>(tr[j], b0, b1); unpack8<T>(bl[j], c0, c1); unpack8<T>(br[j], d0, d1);
        v = pack8<T>(bilerp4(a0, b0, c0, d0, ax, ay), bilerp4(a1, b1, c1, d1, ax, ay));   // one rounding, like the reference's warp
      }
      *reinterpret_cast<uint4*>(sB + (size_t)p * K::RS + 8 * k8) = v;
    }
  }
}

// phase 3: band GEMMs, one (row, 16-pixel block) unit at a time per warp, written out by the same warp from its own
// staging slot `so`; `wg` / `nwg` = the warp's index inside the group of warps that shares the tile
template <typename K, typename T>
__device__ __forceinline__ void band_units(const T* __restrict__ c1, T* __restrict__ out, const T* __restrict__ sB,
                                           T* __restrict__ so, int b, int y0, int x0, int H, int W, int64_t ops, int wg,
                                           int nwg, int lane) {
  constexpr int ND = K::ND, D = K::D, TH = K::TH, TW = K::TW, BW = K::BW, C = K::C, RS = K::RS, NTILES = K::NTILES;
  const int g = lane >> 2, tq = lane & 3;
  const float inv_c = 1.0f / (float)C;
#pragma unroll 1
  for (int un = wg; un < TH * (TW / 16); un += nwg) {
    const int row = un / (TW / 16), blk = un % (TW / 16);
    const int y = y0 + row, xu = x0 + 16 * blk;
    if (y >= H || xu >= W) continue;   // warp-uniform
    uint32_t afr[C / 16][4];
    {
      const int xa = min(xu + g, W - 1), xb = min(xu + g + 8, W - 1);
      const uint32_t* pa = reinterpret_cast<const uint32_t*>(c1 + (((size_t)b * H + y) * W + xa) * C);
      const uint32_t* pb = reinterpret_cast<const uint32_t*>(c1 + (((size_t)b * H + y) * W + xb) * C);
#pragma unroll
      for (int ks = 0; ks < C / 16; ++ks) {
        afr[ks][0] = __ldg(pa + ks * 8 + tq);
        afr[ks][1] = __ldg(pb + ks * 8 + tq);
        afr[ks][2] = __ldg(pa + ks * 8 + 4 + tq);
        afr[ks][3] = __ldg(pb + ks * 8 + 4 + tq);
      }
    }
    constexpr int DYU = PWC_BAND_DYU;
#pragma unroll DYU
    for (int dyi = 0; dyi < ND; ++dyi) {
      // warped row (row + dyi) of the halo tile, halo columns 16 blk .. 16 blk + 15 + 2 R  <->  x' = x0 + 16 blk - R + n
      const T* brow = sB + (size_t)((row + dyi) * BW + 16 * blk) * RS;
      float acc[NTILES][4];
#pragma unroll
      for (int nt = 0; nt < NTILES; ++nt) {
        acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < C / 16; ++ks) {
          uint32_t b0, b1;
          const T* src = brow + (size_t)(8 * nt + (lane & 7)) * RS + 16 * ks + ((lane >> 3) & 1) * 8;
          ldmatrix_x2(b0, b1, smem_u32(src));
          mma16816<T>(acc[nt], afr[ks], b0, b1);
        }
      }
      // accumulator (pixel m, neighbour n): m = g + 8 h, n = 8 nt + 2 t + j  ->  dx index = n - m in [0, ND)
#pragma unroll
      for (int nt = 0; nt < NTILES; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int h = e >> 1, j = e & 1;
          const int m = g + 8 * h, dxi = 8 * nt + 2 * tq + j - m;
          if (dxi >= 0 && dxi < ND) {
            const float v = acc[nt][e] * inv_c;
            so[m * D + dyi * ND + dxi] = ElemTraits<T>::from_f(fmaxf(0.1f * v, v));
          }
        }
    }
    __syncwarp();
    const int npx = min(16, W - xu);
    T* dst = out + (((size_t)b * H + y) * W + xu) * ops;
    if (ops == D && npx == 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      for (int i = lane; i < 16 * D * 2 / 16; i += 32) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(so)[i];
    } else {
      for (int i = lane; i < npx * D; i += 32) {
        const int px = i / D, ch = i - px * D;
        dst[(size_t)px * ops + ch] = so[i];
      }
    }
    __syncwarp();   // the slot is rewritten by the next unit
  }
}

template <int RR, int CC, bool HAS_FLOW, typename T, typename TF>
__global__ void __launch_bounds__(band::Cfg<RR, CC>::NT, band::Cfg<RR, CC>::CTAS)
fwd_mma16_band_kernel(const T* __restrict__ c1, const T* __restrict__ c2, const TF* __restrict__ flow, float flow_scale,
                      T* __restrict__ out, int B, int H, int W, int64_t ops, int tiles_x, int tiles_y) {
  using K = band::Cfg<RR, CC>;
  static_assert(sizeof(T) == 2, "16-bit features only");
  extern __shared__ __align__(128) unsigned char smem[];
  T* sB = reinterpret_cast<T*>(smem);                                   // [BH * BW][RS] warped halo tile
  int4* sCoord = reinterpret_cast<int4*>(smem + K::B_BYTES);           // aliases the staging slots
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  T* so = reinterpret_cast<T*>(smem + K::B_BYTES + warp * K::UNIT_BYTES);   // this warp's [16 px][D] staging slot
  const int t = blockIdx.x;
  const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
  const int x0 = tx * K::TW, y0 = ty * K::TH;
  band_coords<K, HAS_FLOW, T, TF>(flow, flow_scale, b, y0, x0, H, W, sCoord, tid, K::NT);
  __syncthreads();
  band_gather<K, (K::CTAS == 2 ? PWC_BAND_FB : PWC_BAND_FB1), T>(c2, W, sCoord, sB, tid, K::NT);   // 128 registers with two CTAs per SM
  __syncthreads();   // the coordinates are dead from here on: their memory becomes the staging slots
  band_units<K, T>(c1, out, sB, so, b, y0, x0, H, W, ops, warp, K::NWARPS, lane);
}

// Returns 0 when launched, 1 when the shape is not covered, -1 on a CUDA error.
template <typename T, typename TF>
int launch(const T* c1, const T* c2, const TF* flow, float flow_scale, T* out, int B, int H, int W, int C, int64_t ops,
           cudaStream_t st) {
  if (sizeof(T) != 2 || (C != 32 && C != 64) || ops < D) return 1;
  if ((reinterpret_cast<uintptr_t>(c1) | reinterpret_cast<uintptr_t>(c2)) & 15) return 1;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const long long tiles = (long long)tiles_x * tiles_y * B;
  if (tiles <= 0 || tiles >= (1LL << 31)) return 1;
  auto go = [&](auto kern, int smem_bytes) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100) != cudaSuccess) return -1;
    kern<<<(unsigned)tiles, NTHREADS, smem_bytes, st>>>(c1, c2, flow, flow_scale, out, B, H, W, ops, tiles_x, tiles_y);
    return 0;
  };
  if (C == 32)
    return flow ? go(fwd_mma16_kernel<32, true, T, TF>, Cfg<32>::SMEM_BYTES) : go(fwd_mma16_kernel<32, false, T, TF>, Cfg<32>::SMEM_BYTES);
  return flow ? go(fwd_mma16_kernel<64, true, T, TF>, Cfg<64>::SMEM_BYTES) : go(fwd_mma16_kernel<64, false, T, TF>, Cfg<64>::SMEM_BYTES);
}

// the resident-halo variant: search_range 4 (C in {32, 64, 96, 128}) and search_range 8 (the same)
template <int RR, int CC, typename T, typename TF>
int launch_band_cfg(const T* c1, const T* c2, const TF* flow, float flow_scale, T* out, int B, int H, int W, int64_t ops,
                    cudaStream_t st) {
  using K = band::Cfg<RR, CC>;
  const int tiles_x = (W + K::TW - 1) / K::TW, tiles_y = (H + K::TH - 1) / K::TH;
  const long long tiles = (long long)tiles_x * tiles_y * B;
  if (tiles <= 0 || tiles >= (1LL << 31)) return 1;
  auto go = [&](auto kern) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100) != cudaSuccess) return -1;
    kern<<<(unsigned)tiles, K::NT, K::SMEM_BYTES, st>>>(c1, c2, flow, flow_scale, out, B, H, W, ops, tiles_x, tiles_y);
    return 0;
  };
  return flow ? go(fwd_mma16_band_kernel<RR, CC, true, T, TF>) : go(fwd_mma16_band_kernel<RR, CC, false, T, TF>);
}
template <int RR, typename T, typename TF>
int launch_band(const T* c1, const T* c2, const TF* flow, float flow_scale, T* out, int B, int H, int W, int C, int64_t ops,
                cudaStream_t st) {
  if (sizeof(T) != 2 || ops < band::Cfg<RR, 32>::D) return 1;
  if ((reinterpret_cast<uintptr_t>(c1) | reinterpret_cast<uintptr_t>(c2)) & 15) return 1;
  switch (C) {
    case 32: return launch_band_cfg<RR, 32, T, TF>(c1, c2, flow, flow_scale, out, B, H, W, ops, st);
    case 64: return launch_band_cfg<RR, 64, T, TF>(c1, c2, flow, flow_scale, out, B, H, W, ops, st);
    case 96: return launch_band_cfg<RR, 96, T, TF>(c1, c2, flow, flow_scale, out, B, H, W, ops, st);
    case 128: return launch_band_cfg<RR, 128, T, TF>(c1, c2, flow, flow_scale, out, B, H, W, ops, st);
    default: return 1;
  }
}
template <typename T, typename TF>
int launch_r8(const T* c1, const T* c2, const TF* flow, float flow_scale, T* out, int B, int H, int W, int C, int64_t ops,
              cudaStream_t st) {
  return launch_band<8, T, TF>(c1, c2, flow, flow_scale, out, B, H, W, C, ops, st);
}

}  // namespace mma16
}  // namespace pwc
