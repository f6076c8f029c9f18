// Tiled fused warp + cost-volume forward (v1).
//
// One CTA owns a TH x TW tile of output pixels.  Channels are consumed in chunks of CK: for each
// chunk the CTA (phase 1) stages the c1 tile and the *warped* c2 tile + halo into shared memory -
// the bilinear gather is evaluated here, so the warped features never exist in HBM - and
// (phase 2) every thread accumulates a register tile of 2 rows x PX pixels x (2r+1) dx for one dy.
// Zero padding is applied after warping (core_costvol.py:27), edge clamping inside the warp
// (core_warp.py:101-104).  Shared-memory rows are padded by one 16-B slot every PX pixels so the
// 8 x-groups of a quarter-warp hit 8 different bank groups with compile-time offsets.
#pragma once
#include "common.cuh"

namespace pwc {
namespace tile {

// TW x PX: 32 x 4 (256-pixel tiles, 2 CTAs per SM) for levels with enough tiles to fill the GPU; 16 x 2 (128-pixel tiles,
// 48 KB, 3 CTAs per SM) for the short pyramid levels, which are latency-bound and want more, lighter CTAs.
// search_range 8 (BASELINE configs[4], 289 channels): 16 x 2 with 17 warps (one dy each, 68 accumulators per thread,
// 112 registers), 148 KB of output staging, one CTA per SM; that workload is FMA-bound.
template <int R_, int TW_ = 32, int PX_ = 4>
struct Cfg {
  static constexpr int R = R_;
  static constexpr int TH = 8, TW = TW_, PX = PX_;
  static constexpr int ND = 2 * R + 1, D = ND * ND;
  static constexpr int NXG = TW / PX;
  static constexpr int RPAIRS = TH / 2;
  static constexpr int NTHREADS = NXG * RPAIRS * ND;
  static constexpr int BW = TW + 2 * R, BH = TH + 2 * R;
  static constexpr int A_RS = TW + TW / PX;
  static constexpr int B_RS = BW + (BW + PX - 1) / PX;
  static constexpr int CK = 16, KV = CK / 4;
  // plane strides are padded to 2 (mod 8) float4 so that the 4 kv lanes x 2 pixels of a quarter-warp store to
  // 8 different bank groups in the fill phase (ncu: 16.7M store bank conflicts per launch without the pad)
  static constexpr int pad2(int n) { return n + ((2 - n % 8) + 8) % 8; }
  static constexpr int A_PLANE = pad2(TH * A_RS), B_PLANE = pad2(BH * B_RS);
  static constexpr int STAGE_BYTES = KV * (A_PLANE + B_PLANE) * 16;
  static constexpr int OUT_BYTES = TH * TW * D * 4;
  static constexpr int MAIN_BYTES = STAGE_BYTES > OUT_BYTES ? STAGE_BYTES : OUT_BYTES;
  static constexpr int COORD_N = BH * BW;
  static constexpr int SMEM_BYTES = MAIN_BYTES + COORD_N * 16;
  static_assert(NTHREADS % 32 == 0, "whole warps");
  static_assert(NXG * RPAIRS == 32, "one dy per warp");
};

template <int R, int TWC, int PXC, typename T, typename TF>
#ifndef PWC_TILE_MAXNREG
#define PWC_TILE_MAXNREG 96   /* 2 CTAs/SM (9 warps x 96 regs x 2); measured faster than 112 regs x 1 CTA */
#endif
__global__ void __maxnreg__(R == 8 ? 96 : (TWC == 32 ? PWC_TILE_MAXNREG : 72))  // r = 8: 17 warps are accounted as 20 (allocation granularity 4): 20 x 32 x 96 <= 64 K
fwd_tile_kernel(const T* __restrict__ c1, const T* __restrict__ c2, const TF* __restrict__ flow,
                float flow_scale, T* __restrict__ out, int B, int H, int W, int C,
                int64_t out_pixel_stride, int tiles_x, int tiles_y) {
  using K = Cfg<R, TWC, PXC>;
  constexpr int TH = K::TH, TW = K::TW, PX = K::PX, ND = K::ND, D = K::D, KV = K::KV;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* sA = reinterpret_cast<float4*>(smem_raw);
  float4* sB = sA + KV * K::A_PLANE;
  float* sOut = reinterpret_cast<float*>(smem_raw);
  int4* sCoord = reinterpret_cast<int4*>(smem_raw + K::MAIN_BYTES);

  const int tid = threadIdx.x;
  const int t = blockIdx.x;
  const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
  const int x0 = tx * TW, y0 = ty * TH;

  // ---- phase 0: sampling coordinates of every halo pixel (zero padding => invalid) ----
  for (int i = tid; i < K::COORD_N; i += K::NTHREADS) {
    const int rb = i / K::BW, u = i - rb * K::BW;
    const int y = y0 - R + rb, x = x0 - R + u;
    int4 cd = make_int4(0, 0, 0, 0);
    if (y >= 0 && y < H && x >= 0 && x < W) {
      if (flow != nullptr) {
        const SampleCoord s = sample_coord<T, TF>(flow, flow_scale, b, y, x, H, W);
        cd = make_int4(s.pix, __float_as_int(s.ax), __float_as_int(s.ay), 1);
      } else {
        cd = make_int4((b * H + y) * W + x, 0, 0, 2);
      }
    }
    sCoord[i] = cd;
  }

  // ---- thread -> (x-group, row pair, dy) ----
  const int xg = tid % K::NXG;
  const int rp = (tid / K::NXG) % K::RPAIRS;
  const int dyi = tid / (K::NXG * K::RPAIRS);
  const int rows[2] = {rp, rp + K::RPAIRS};

  float acc[2][PX][ND];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int d = 0; d < ND; ++d) acc[h][p][d] = 0.0f;

  for (int c0 = 0; c0 < C; c0 += K::CK) {
    __syncthreads();  // coords visible / previous chunk consumed
    // ---- phase 1a: c1 tile (FA loads in flight per thread) ----
    {
      constexpr int NA = TH * TW * KV, FA = R == 8 ? 2 : 4;
#pragma unroll 1
      for (int i0 = tid; i0 < NA; i0 += FA * K::NTHREADS) {
        float4 v[FA];
#pragma unroll
        for (int j = 0; j < FA; ++j) {
          const int i = i0 + j * K::NTHREADS;
          const int kv = i % KV, p = i / KV;
          const int row = p / TW, col = p - row * TW;
          const int y = y0 + row, x = x0 + col, c = c0 + 4 * kv;
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < NA && y < H && x < W && c < C) v[j] = ld4(c1 + (((size_t)b * H + y) * W + x) * C + c);
        }
#pragma unroll
        for (int j = 0; j < FA; ++j) {
          const int i = i0 + j * K::NTHREADS;
          const int kv = i % KV, p = i / KV;
          const int row = p / TW, col = p - row * TW;
          if (i < NA) sA[kv * K::A_PLANE + row * K::A_RS + col + col / PX] = v[j];
        }
      }
    }
    // ---- phase 1b: warped c2 tile + halo (FB items = 4*FB corner loads in flight per thread) ----
    {
      constexpr int NB = K::COORD_N * KV, FB = R == 8 ? 1 : 3;  // r = 8: 68 accumulators leave room for one item
      const size_t rowstride = (size_t)W * C;
#pragma unroll 1
      for (int i0 = tid; i0 < NB; i0 += FB * K::NTHREADS) {
        float4 tl[FB], tr[FB], bl[FB], br[FB];
        int4 cd[FB];
#pragma unroll
        for (int j = 0; j < FB; ++j) {
          const int i = i0 + j * K::NTHREADS;
          const int kv = i % KV, p = i / KV;
          cd[j] = make_int4(0, 0, 0, 0);
          if (i < NB) cd[j] = sCoord[p];
          const int c = c0 + 4 * kv;
          if (c >= C) cd[j].w = 0;
          tl[j] = tr[j] = bl[j] = br[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cd[j].w != 0) {
            const T* base = c2 + (size_t)cd[j].x * C + c;
            tl[j] = ld4(base);
            if (cd[j].w == 1) {
              tr[j] = ld4(base + C);
              bl[j] = ld4(base + rowstride);
              br[j] = ld4(base + rowstride + C);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < FB; ++j) {
          const int i = i0 + j * K::NTHREADS;
          if (i >= NB) continue;
          const int kv = i % KV, p = i / KV;
          float4 v = tl[j];
          if (cd[j].w == 1)
            v = round4<T>(bilerp4(tl[j], tr[j], bl[j], br[j], __int_as_float(cd[j].y), __int_as_float(cd[j].z)));
          const int rb = p / K::BW, u = p - rb * K::BW;
          sB[kv * K::B_PLANE + rb * K::B_RS + u + u / PX] = v;
        }
      }
    }
    __syncthreads();
    // ---- phase 2: register-tiled correlation ----
#pragma unroll
    for (int kv = 0; kv < KV; ++kv) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4* ap = sA + kv * K::A_PLANE + rows[h] * K::A_RS + xg * (PX + 1);
        const float4* bp = sB + kv * K::B_PLANE + (rows[h] + dyi) * K::B_RS + xg * (PX + 1);
        float4 a[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) a[p] = ap[p];
#pragma unroll
        for (int k = 0; k < PX + 2 * R; ++k) {
          const float4 bv = bp[k + k / PX];
#pragma unroll
          for (int p = 0; p < PX; ++p) {
            const int dxi = k - p;
            if (dxi >= 0 && dxi < ND) fma4(acc[h][p][dxi], a[p], bv);
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- epilogue: 1/C, leaky-relu, transpose through shared memory, coalesced store ----
  const float inv_c = 1.0f / (float)C;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int d = 0; d < ND; ++d)
        sOut[(rows[h] * TW + xg * PX + p) * D + dyi * ND + d] = lrelu01(acc[h][p][d] * inv_c);
  __syncthreads();

  const int nrows = min(TH, H - y0), ncols = min(TW, W - x0);
  for (int row = 0; row < nrows; ++row) {
    const size_t gp = ((size_t)b * H + y0 + row) * W + x0;
    const float* src = sOut + row * TW * D;
    const int n = ncols * D;
    bool vec = false;
    if (sizeof(T) == 4) vec = (out_pixel_stride == D) && (((gp * D) & 3) == 0) && ((n & 3) == 0) &&
                              ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec) {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + gp * D);
      const float4* s4 = reinterpret_cast<const float4*>(src);
      for (int e = tid; e < (n >> 2); e += K::NTHREADS) dst[e] = s4[e];
    } else {
      for (int e = tid; e < n; e += K::NTHREADS) {
        const int pix = e / D, ch = e - pix * D;
        out[(gp + pix) * out_pixel_stride + ch] = ElemTraits<T>::from_f(src[e]);
      }
    }
  }
}

}  // namespace tile
}  // namespace pwc
