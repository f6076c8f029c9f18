// Shared device helpers for the warp + cost-volume kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pwc {

// ---- element types -------------------------------------------------------------------
template <typename T> struct ElemTraits;
template <> struct ElemTraits<float> {
  static constexpr bool kIsHalf = false;
  __device__ static __forceinline__ float to_f(float v) { return v; }
  __device__ static __forceinline__ float from_f(float v) { return v; }
  // round-trip through the storage type (identity for fp32)
  __device__ static __forceinline__ float round_f(float v) { return v; }
};
template <> struct ElemTraits<__half> {
  static constexpr bool kIsHalf = true;
  __device__ static __forceinline__ float to_f(__half v) { return __half2float(v); }
  __device__ static __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  __device__ static __forceinline__ float round_f(float v) { return __half2float(__float2half_rn(v)); }
};
template <> struct ElemTraits<__nv_bfloat16> {
  static constexpr bool kIsHalf = true;
  __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  __device__ static __forceinline__ float round_f(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

// 4 consecutive elements -> float4.  Pointer must be aligned to 4 elements.
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const __half* p) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  __half2 a = *reinterpret_cast<__half2*>(&u.x), b = *reinterpret_cast<__half2*>(&u.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
// streaming variants: do not allocate in L1 (with ~220 KB of shared memory per CTA the L1 is tiny and its
// line allocation, not DRAM, bounds the number of gathers in flight)
__device__ __forceinline__ float4 ld4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld4_stream(const __half* p) {
  uint2 u;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(u.x), "=r"(u.y) : "l"(p));
  __half2 a = *reinterpret_cast<__half2*>(&u.x), b = *reinterpret_cast<__half2*>(&u.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ float4 ld4_stream(const __nv_bfloat16* p) {
  uint2 u;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(u.x), "=r"(u.y) : "l"(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ float lrelu01(float x) { return fmaxf(0.1f * x, x); }  // TF1.10 leaky_relu

// ---- bilinear sampling coordinates (tfoptflow/core_warp.py:93-119,189-194) ------------
// pix   : linear index (b*H + y0)*W + x0 of the top-left corner (clamped floor)
// ax/ay : clamped interpolation weights in x / y
// mx/my : 1 where the clamp passes gradient (0 < q - floor <= 1), else 0  (TF Maximum/Minimum grad)
struct SampleCoord {
  int pix;
  float ax, ay;
  float mx, my;
};

template <typename TF>
__device__ __forceinline__ float query_coord(int grid, TF flow_v, float scale);
// fp32: fl(grid - fl(flow * scale)), two roundings exactly like the two TF ops (no FMA contraction)
template <>
__device__ __forceinline__ float query_coord<float>(int grid, float flow_v, float scale) {
  return __fsub_rn((float)grid, __fmul_rn(flow_v, scale));
}
// fp16: every op result is an fp16 tensor in the reference (model_pwcnet.py:390-394, core_warp.py:191-194)
template <>
__device__ __forceinline__ float query_coord<__half>(int grid, __half flow_v, float scale) {
  float s = __half2float(__float2half_rn(__half2float(flow_v) * scale));  // exact product, one rounding
  return __half2float(__double2half((double)grid - (double)s));           // exact difference, one rounding
}
template <>
__device__ __forceinline__ float query_coord<__nv_bfloat16>(int grid, __nv_bfloat16 flow_v, float scale) {
  float s = __bfloat162float(__float2bfloat16_rn(__bfloat162float(flow_v) * scale));
  return __bfloat162float(__double2bfloat16((double)grid - (double)s));
}

// TImg decides the dtype alpha is cast to (core_warp.py:112).  f0 / f1: flow of the pixel (rows / columns).
template <typename TImg, typename TF>
__device__ __forceinline__ SampleCoord sample_coord_v(TF f0, TF f1, float flow_scale, int b, int y, int x, int H, int W) {
  const float qy = query_coord<TF>(y, f0, flow_scale);
  const float qx = query_coord<TF>(x, f1, flow_scale);
  const float fy = fminf(fmaxf(0.0f, floorf(qy)), (float)(H - 2));
  const float fx = fminf(fmaxf(0.0f, floorf(qx)), (float)(W - 2));
  // queries - floor in the query dtype (exact in fp32 here because both operands are
  // representable in TF and |q - floor| is either small or gets clamped), then cast to grid dtype
  float ay = ElemTraits<TF>::round_f(__fsub_rn(qy, fy));
  float ax = ElemTraits<TF>::round_f(__fsub_rn(qx, fx));
  ay = ElemTraits<TImg>::round_f(ay);
  ax = ElemTraits<TImg>::round_f(ax);
  SampleCoord s;
  s.my = (ay > 0.0f && ay <= 1.0f) ? 1.0f : 0.0f;
  s.mx = (ax > 0.0f && ax <= 1.0f) ? 1.0f : 0.0f;
  s.ay = fminf(fmaxf(0.0f, ay), 1.0f);
  s.ax = fminf(fmaxf(0.0f, ax), 1.0f);
  s.pix = (b * H + (int)fy) * W + (int)fx;
  return s;
}
template <typename TImg, typename TF>
__device__ __forceinline__ SampleCoord sample_coord(const TF* __restrict__ flow, float flow_scale,
                                                    int b, int y, int x, int H, int W) {
  const size_t p = ((size_t)b * H + y) * W + x;
  return sample_coord_v<TImg, TF>(flow[2 * p], flow[2 * p + 1], flow_scale, b, y, x, H, W);
}

// lerp exactly as the reference evaluates it: alpha * (hi - lo) + lo with three roundings
__device__ __forceinline__ float lerp3(float a, float lo, float hi) {
  return __fadd_rn(__fmul_rn(a, __fsub_rn(hi, lo)), lo);
}
__device__ __forceinline__ float4 bilerp4(float4 tl, float4 tr, float4 bl, float4 br, float ax, float ay) {
  float4 o;
  o.x = lerp3(ay, lerp3(ax, tl.x, tr.x), lerp3(ax, bl.x, br.x));
  o.y = lerp3(ay, lerp3(ax, tl.y, tr.y), lerp3(ax, bl.y, br.y));
  o.z = lerp3(ay, lerp3(ax, tl.z, tr.z), lerp3(ax, bl.z, br.z));
  o.w = lerp3(ay, lerp3(ax, tl.w, tr.w), lerp3(ax, bl.w, br.w));
  return o;
}
template <typename T>
__device__ __forceinline__ float4 round4(float4 v) {
  if (ElemTraits<T>::kIsHalf) {
    v.x = ElemTraits<T>::round_f(v.x);
    v.y = ElemTraits<T>::round_f(v.y);
    v.z = ElemTraits<T>::round_f(v.z);
    v.w = ElemTraits<T>::round_f(v.w);
  }
  return v;
}

// bilinear sample of 4 consecutive channels of `img` (NHWC, C channels) at `s`
template <typename T>
__device__ __forceinline__ float4 sample4(const T* __restrict__ img, const SampleCoord& s, int W, int C, int c) {
  const T* base = img + (size_t)s.pix * C + c;
  const float4 tl = ld4(base), tr = ld4(base + C);
  const float4 bl = ld4(base + (size_t)W * C), br = ld4(base + (size_t)W * C + C);
  return round4<T>(bilerp4(tl, tr, bl, br, s.ax, s.ay));
}
template <typename T>
__device__ __forceinline__ float sample1(const T* __restrict__ img, const SampleCoord& s, int W, int C, int c) {
  const T* base = img + (size_t)s.pix * C + c;
  const float tl = ElemTraits<T>::to_f(base[0]), tr = ElemTraits<T>::to_f(base[C]);
  const float bl = ElemTraits<T>::to_f(base[(size_t)W * C]), br = ElemTraits<T>::to_f(base[(size_t)W * C + C]);
  return ElemTraits<T>::round_f(lerp3(s.ay, lerp3(s.ax, tl, tr), lerp3(s.ax, bl, br)));
}

__device__ __forceinline__ void fma4(float& acc, const float4& a, const float4& b) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
}

}  // namespace pwc
