// Generic (any C, any r, any alignment) forward kernels: the correctness baseline on the GPU and
// the path for shapes the tuned kernels do not cover.  One thread per output element.
#pragma once
#include "common.cuh"

namespace pwc {

// out[p, d] = lrelu( 1/C * sum_c c1[p,c] * warpZ[p+d, c] ), warp evaluated on the fly.
template <typename T, typename TF>
__global__ void __launch_bounds__(256)
fwd_generic_kernel(const T* __restrict__ c1, const T* __restrict__ c2, const TF* __restrict__ flow,
                   float flow_scale, T* __restrict__ out, int B, int H, int W, int C, int r,
                   int64_t out_pixel_stride) {
  const int n = 2 * r + 1, D = n * n;
  const int64_t total = (int64_t)B * H * W * D;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(idx % D);
    const int64_t p = idx / D;
    const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((int64_t)W * H));
    const int dy = d / n - r, dx = d % n - r;
    const int yy = y + dy, xx = x + dx;
    float acc = 0.0f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const T* a = c1 + (size_t)p * C;
      if (flow != nullptr) {
        const SampleCoord s = sample_coord<T, TF>(flow, flow_scale, b, yy, xx, H, W);
        for (int c = 0; c < C; ++c) acc = fmaf(ElemTraits<T>::to_f(a[c]), sample1<T>(c2, s, W, C, c), acc);
      } else {
        const T* w = c2 + (((size_t)b * H + yy) * W + xx) * C;
        for (int c = 0; c < C; ++c) acc = fmaf(ElemTraits<T>::to_f(a[c]), ElemTraits<T>::to_f(w[c]), acc);
      }
    }
    out[(size_t)p * out_pixel_stride + d] = ElemTraits<T>::from_f(lrelu01(acc * (1.0f / (float)C)));
  }
}

// Un-warped cost volume (the top pyramid level, model_pwcnet.py:1554) for small images: one thread per output
// element, 16-byte loads.  A level-6 problem (7 x 16 pixels per image) has too few tiles to fill the GPU with the
// tiled kernel; here every (pixel, displacement) is its own thread and the operands come out of L1/L2.
template <typename T>
__global__ void __launch_bounds__(256)
fwd_noflow_vec4_kernel(const T* __restrict__ c1, const T* __restrict__ c2, T* __restrict__ out, int B, int H, int W,
                       int C, int r, int64_t out_pixel_stride) {
  const int n = 2 * r + 1, D = n * n;
  const int64_t total = (int64_t)B * H * W * D;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(idx % D);
    const int64_t p = idx / D;
    const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((int64_t)W * H));
    const int yy = y + d / n - r, xx = x + d % n - r;
    float a0 = 0.0f, a1 = 0.0f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const T* a = c1 + (size_t)p * C;
      const T* w = c2 + (((size_t)b * H + yy) * W + xx) * C;
      int c = 0;
      for (; c + 8 <= C; c += 8) {
        const float4 u0 = ld4(a + c), v0 = ld4(w + c), u1 = ld4(a + c + 4), v1 = ld4(w + c + 4);
        fma4(a0, u0, v0);
        fma4(a1, u1, v1);
      }
      if (c < C) fma4(a0, ld4(a + c), ld4(w + c));
    }
    out[(size_t)p * out_pixel_stride + d] = ElemTraits<T>::from_f(lrelu01((a0 + a1) * (1.0f / (float)C)));
  }
}

// Same, register-tiled: one thread owns 2 pixels along x and 3 consecutive dx of one dy (6 outputs) and loads the 2 c1
// vectors and the 4 c2 columns they touch once: 6 instead of 12 vector loads per 6 outputs.  r = 4, W even.
template <typename T>
__global__ void __launch_bounds__(256)
fwd_noflow_r4_tiled_kernel(const T* __restrict__ c1, const T* __restrict__ c2, T* __restrict__ out, int B, int H, int W,
                           int C, int64_t out_pixel_stride) {
  constexpr int R = 4, N = 9;
  const int WP = W / 2;
  const int64_t total = (int64_t)B * H * WP * N * 3;
  const float inv_c = 1.0f / (float)C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int t3 = (int)(idx % 3), dyi = (int)((idx / 3) % N);
    const int64_t pp = idx / (3 * N);
    const int x0 = (int)(pp % WP) * 2, y = (int)((pp / WP) % H), b = (int)(pp / ((int64_t)WP * H));
    const int yy = y + dyi - R;
    float acc[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    const size_t row0 = ((size_t)b * H + y) * W;
    if (yy >= 0 && yy < H) {
      const size_t rown = ((size_t)b * H + yy) * W;
      const int xb = x0 + 3 * t3 - R;  // first c2 column; columns xb .. xb+3
      const T* a0 = c1 + (row0 + x0) * C;
      const T* a1 = a0 + C;
      for (int c = 0; c < C; c += 4) {
        const float4 u0 = ld4(a0 + c), u1 = ld4(a1 + c);
        float4 w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int xx = xb + k;
          w[k] = (xx >= 0 && xx < W) ? ld4(c2 + (rown + xx) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // pixel i, dx index 3*t3 + j  ->  column k = i + j
          fma4(acc[0][j], u0, w[j]);
          fma4(acc[1][j], u1, w[j + 1]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        out[(row0 + x0 + i) * out_pixel_stride + dyi * N + 3 * t3 + j] = ElemTraits<T>::from_f(lrelu01(acc[i][j] * inv_c));
  }
}

// dense_image_warp forward, one thread per (pixel, channel)
template <typename T, typename TF>
__global__ void __launch_bounds__(256)
warp_fwd_scalar_kernel(const T* __restrict__ img, const TF* __restrict__ flow, float flow_scale,
                       T* __restrict__ out, int B, int H, int W, int C) {
  const int64_t total = (int64_t)B * H * W * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int64_t p = idx / C;
    const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((int64_t)W * H));
    const SampleCoord s = sample_coord<T, TF>(flow, flow_scale, b, y, x, H, W);
    out[idx] = ElemTraits<T>::from_f(sample1<T>(img, s, W, C, c));
  }
}

// dense_image_warp forward, one thread per (pixel, 4 channels); needs C % 4 == 0 and aligned pointers
template <typename T, typename TF>
__global__ void __launch_bounds__(256)
warp_fwd_vec4_kernel(const T* __restrict__ img, const TF* __restrict__ flow, float flow_scale,
                     T* __restrict__ out, int B, int H, int W, int C) {
  const int C4 = C >> 2;
  const int64_t total = (int64_t)B * H * W * C4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C4) * 4;
    const int64_t p = idx / C4;
    const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((int64_t)W * H));
    const SampleCoord s = sample_coord<T, TF>(flow, flow_scale, b, y, x, H, W);
    st4(out + (size_t)p * C + c, sample4<T>(img, s, W, C, c));
  }
}

}  // namespace pwc
