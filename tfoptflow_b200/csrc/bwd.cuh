// Backward kernels (v1): what TF 1.10 autodiff emits for cost_volume / dense_image_warp
// (SURVEY.md section 8a row G), written as gather-form kernels plus one scatter for the warp.
//   g'[p,d]    = g[p,d] * (out[p,d] > 0 ? 1 : 0.1)        (ties go to the 0.1 branch)
//   dc1[p,c]   = 1/C * sum_d g'[p,d]   * warpZ[p+d,c]
//   dwarp[q,c] = 1/C * sum_d g'[q-d,d] * c1[q-d,c]          (q-d inside the image)
//   dc2        = scatter-add of dwarp with the 4 bilinear weights
//   dflow      = -(d/dalpha . dwarp) * clamp-mask * flow_scale
#pragma once
#include "common.cuh"
#ifndef PWC_BWD_MINBLOCKS
#define PWC_BWD_MINBLOCKS 3   /* 3 blocks of 256 per SM (<= 85 registers): level-2 fwd+bwd 2.16 -> 1.72 ms; 4 blocks: 1.93 ms */
#endif

namespace pwc {

template <int VEC> struct Vec { float v[VEC]; };

template <int VEC, typename T>
__device__ __forceinline__ Vec<VEC> ldv(const T* p) {
  Vec<VEC> r;
  if (VEC == 4) {
    const float4 f = ld4(p);
    r.v[0] = f.x; r.v[1 % VEC] = f.y; r.v[2 % VEC] = f.z; r.v[3 % VEC] = f.w;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = ElemTraits<T>::to_f(p[i]);
  }
  return r;
}
template <int VEC, typename T>
__device__ __forceinline__ void stv(T* p, const Vec<VEC>& r) {
  if (VEC == 4) {
    st4(p, make_float4(r.v[0], r.v[1 % VEC], r.v[2 % VEC], r.v[3 % VEC]));
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = ElemTraits<T>::from_f(r.v[i]);
  }
}

template <typename T>
__device__ __forceinline__ float gprime(const T* __restrict__ g, const T* __restrict__ out, size_t p, int d,
                                        int64_t gps, int64_t ops) {
  const float gv = ElemTraits<T>::to_f(g[p * gps + d]);
  const float ov = ElemTraits<T>::to_f(out[p * ops + d]);
  return gv * (ov > 0.0f ? 1.0f : 0.1f);
}

// MODE 0: dc1[p,c]   (other = warp, reads other[p+d])
// MODE 1: dwarp[q,c] (other = c1,   reads g'[q-d,d] and other[q-d])
template <int MODE, int VEC, typename T, typename TO>
__global__ void __launch_bounds__(256)
bwd_corr_kernel(const T* __restrict__ g, const T* __restrict__ out, const T* __restrict__ other,
                TO* __restrict__ dst, int B, int H, int W, int C, int r, int64_t gps, int64_t ops) {
  const int n = 2 * r + 1;
  const int CV = C / VEC;
  const int64_t total = (int64_t)B * H * W * CV;
  const float inv_c = 1.0f / (float)C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % CV) * VEC;
    const int64_t p = idx / CV;
    const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((int64_t)W * H));
    Vec<VEC> acc;
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc.v[i] = 0.0f;
    for (int dyi = 0; dyi < n; ++dyi) {
      const int yy = MODE == 0 ? y + dyi - r : y - (dyi - r);
      if (yy < 0 || yy >= H) continue;
      for (int dxi = 0; dxi < n; ++dxi) {
        const int xx = MODE == 0 ? x + dxi - r : x - (dxi - r);
        if (xx < 0 || xx >= W) continue;
        const size_t pp = ((size_t)b * H + yy) * W + xx;
        const float gp = gprime<T>(g, out, MODE == 0 ? (size_t)p : pp, dyi * n + dxi, gps, ops);
        const Vec<VEC> o = ldv<VEC, T>(other + pp * C + c);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc.v[i] = fmaf(gp, o.v[i], acc.v[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc.v[i] *= inv_c;
    stv<VEC, TO>(dst + (size_t)p * C + c, acc);
  }
}


// Register-tiled variant for r = 4 (the model's search range): one thread owns PXT = 4 pixels along x and CVT = 2
// channel quads.  Every neighbour vector is loaded once per thread and used for up to 4 (pixel, dx) combinations, and
// every g' value serves 8 channels: ~2.6x less L1 traffic per FMA than the one-pixel kernel, which is bound by L1
// bandwidth (81 LDG.128 per output float4).  Needs W % 4 == 0 and C % 8 == 0.
//   MODE 0: dc1[p,c]   += g'[p, (dy,dx)] * warp[p + (dy,dx), c]
//   MODE 1: dwarp[q,c] += g'[s, (dy,dx)] * c1[s, c],  s = q - (dy,dx)
template <int MODE, typename T, typename TO>
__global__ void __launch_bounds__(256, PWC_BWD_MINBLOCKS)
bwd_corr_r4_tiled_kernel(const T* __restrict__ g, const T* __restrict__ out, const T* __restrict__ other,
                         TO* __restrict__ dst, int B, int H, int W, int C, int64_t gps, int64_t ops) {
  constexpr int R = 4, N = 9, PXT = 4, CVT = 2;
  const int CG = C / (4 * CVT), WG = W / PXT;
  const int64_t total = (int64_t)B * H * WG * CG;
  const float inv_c = 1.0f / (float)C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % CG) * 4 * CVT;
    const int64_t pg = idx / CG;
    const int x0 = (int)(pg % WG) * PXT, y = (int)((pg / WG) % H), b = (int)(pg / ((int64_t)WG * H));
    float acc[PXT][4 * CVT];
#pragma unroll
    for (int i = 0; i < PXT; ++i)
#pragma unroll
      for (int k = 0; k < 4 * CVT; ++k) acc[i][k] = 0.0f;
    const size_t row0 = ((size_t)b * H + y) * W;
#pragma unroll 1
    for (int dyi = 0; dyi < N; ++dyi) {
      const int yy = MODE == 0 ? y + dyi - R : y - (dyi - R);  // row of the neighbour vector
      if (yy < 0 || yy >= H) continue;
      const size_t rown = ((size_t)b * H + yy) * W;
#pragma unroll
      for (int k = 0; k < PXT + 2 * R; ++k) {  // neighbour column xx = x0 - R + k
        const int xx = x0 - R + k;
        if (xx < 0 || xx >= W) continue;
        float o[4 * CVT];
#pragma unroll
        for (int v = 0; v < CVT; ++v) {
          const float4 f = ld4(other + (rown + xx) * C + c + 4 * v);
          o[4 * v] = f.x; o[4 * v + 1] = f.y; o[4 * v + 2] = f.z; o[4 * v + 3] = f.w;
        }
#pragma unroll
        for (int i = 0; i < PXT; ++i) {
          // MODE 0: xx = x0 + i + (dxi - R)  ->  dxi = k - i ;  MODE 1: xx = x0 + i - (dxi - R)  ->  dxi = 2R + i - k
          const int dxi = MODE == 0 ? k - i : 2 * R + i - k;
          if (dxi < 0 || dxi >= N) continue;
          const float gp = gprime<T>(g, out, MODE == 0 ? row0 + x0 + i : rown + xx, dyi * N + dxi, gps, ops);
#pragma unroll
          for (int q = 0; q < 4 * CVT; ++q) acc[i][q] = fmaf(gp, o[q], acc[i][q]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < PXT; ++i)
#pragma unroll
      for (int v = 0; v < CVT; ++v) {
        Vec<4> r4;
#pragma unroll
        for (int q = 0; q < 4; ++q) r4.v[q] = acc[i][4 * v + q] * inv_c;
        stv<4, TO>(dst + (row0 + x0 + i) * C + c + 4 * v, r4);
      }
  }
}

// dense_image_warp backward: one warp per output pixel, lanes stride over channels.
// dimg_acc is an fp32 accumulator (zero-initialised by the caller); dflow gets -(dalpha)*mask*scale.
// LPP = lanes per pixel (8 / 16 / 32): with 32 or 64 channels a whole warp per pixel would leave 3/4 or 1/2 of the
// lanes idle, so a warp takes 32 / LPP pixels.
template <int VEC, int LPP, typename T, typename TF, typename TG>
__global__ void __launch_bounds__(256)
warp_bwd_kernel(const TG* __restrict__ gout, const T* __restrict__ img, const TF* __restrict__ flow,
                float flow_scale, float* __restrict__ dimg_acc, TF* __restrict__ dflow,
                int B, int H, int W, int C) {
  constexpr int PPW = 32 / LPP;  // pixels per warp
  const int lane = threadIdx.x & (LPP - 1), sub = (threadIdx.x & 31) / LPP;
  const int64_t npix = (int64_t)B * H * W;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // the trip count is warp-uniform (shuffles below): pixels beyond the end are clamped and not written
  for (int64_t p0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * PPW; p0 < npix; p0 += warps_total * PPW) {
    const bool valid = p0 + sub < npix;
    const int64_t p = valid ? p0 + sub : npix - 1;
    const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((int64_t)W * H));
    const SampleCoord s = sample_coord<T, TF>(flow, flow_scale, b, y, x, H, W);
    const float ax = s.ax, ay = s.ay;
    const float w_tl = (1.f - ay) * (1.f - ax), w_tr = (1.f - ay) * ax, w_bl = ay * (1.f - ax), w_br = ay * ax;
    float dax = 0.f, day = 0.f;
    const size_t i_tl = (size_t)s.pix * C, i_tr = i_tl + C, i_bl = i_tl + (size_t)W * C, i_br = i_bl + C;
    for (int c = lane * VEC; valid && c < C; c += LPP * VEC) {
      const Vec<VEC> gw = ldv<VEC, TG>(gout + (size_t)p * C + c);
      const Vec<VEC> tl = ldv<VEC, T>(img + i_tl + c), tr = ldv<VEC, T>(img + i_tr + c);
      const Vec<VEC> bl = ldv<VEC, T>(img + i_bl + c), br = ldv<VEC, T>(img + i_br + c);
      if (VEC == 4) {
        atomicAdd(reinterpret_cast<float4*>(dimg_acc + i_tl + c),
                  make_float4(gw.v[0] * w_tl, gw.v[1 % VEC] * w_tl, gw.v[2 % VEC] * w_tl, gw.v[3 % VEC] * w_tl));
        atomicAdd(reinterpret_cast<float4*>(dimg_acc + i_tr + c),
                  make_float4(gw.v[0] * w_tr, gw.v[1 % VEC] * w_tr, gw.v[2 % VEC] * w_tr, gw.v[3 % VEC] * w_tr));
        atomicAdd(reinterpret_cast<float4*>(dimg_acc + i_bl + c),
                  make_float4(gw.v[0] * w_bl, gw.v[1 % VEC] * w_bl, gw.v[2 % VEC] * w_bl, gw.v[3 % VEC] * w_bl));
        atomicAdd(reinterpret_cast<float4*>(dimg_acc + i_br + c),
                  make_float4(gw.v[0] * w_br, gw.v[1 % VEC] * w_br, gw.v[2 % VEC] * w_br, gw.v[3 % VEC] * w_br));
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          atomicAdd(dimg_acc + i_tl + c + i, gw.v[i] * w_tl);
          atomicAdd(dimg_acc + i_tr + c + i, gw.v[i] * w_tr);
          atomicAdd(dimg_acc + i_bl + c + i, gw.v[i] * w_bl);
          atomicAdd(dimg_acc + i_br + c + i, gw.v[i] * w_br);
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float top = ax * (tr.v[i] - tl.v[i]) + tl.v[i];
        const float bot = ax * (br.v[i] - bl.v[i]) + bl.v[i];
        dax = fmaf(gw.v[i], (1.f - ay) * (tr.v[i] - tl.v[i]) + ay * (br.v[i] - bl.v[i]), dax);
        day = fmaf(gw.v[i], bot - top, day);
      }
    }
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) {
      dax += __shfl_xor_sync(0xffffffffu, dax, o);
      day += __shfl_xor_sync(0xffffffffu, day, o);
    }
    if (lane == 0 && valid) {
      dflow[2 * p] = ElemTraits<TF>::from_f(-day * s.my * flow_scale);
      dflow[2 * p + 1] = ElemTraits<TF>::from_f(-dax * s.mx * flow_scale);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) cast_from_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = ElemTraits<T>::from_f(src[i]);
}

}  // namespace pwc
