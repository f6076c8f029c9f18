// C ABI of libpwc_b200.so (see include/pwc_b200.h).  Host-side dispatch only; kernels live in *.cuh.
#include "../../include/pwc_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "bwd.cuh"
#include "bwd_fused.cuh"
#ifndef PWC_BWD_NA1
#define PWC_BWD_NA1 4    /* K1 with a flow: 4 A builders + 8 gather warps (level 2: 6 + 6 592 us, 5 + 7 583, 4 + 8 569, 7 + 5 600) */
#endif
#ifndef PWC_BWD_NA1N
#define PWC_BWD_NA1N 8   /* K1 without a flow: 8 A producers + 4 copy warps */
#endif
#ifndef PWC_BWD_NA2
#define PWC_BWD_NA2 2    /* K2 with a flow: 2 A-row warps + 2 copy warps + 8 scatter warps (one per staging slot) */
#endif
#ifndef PWC_BWD_NA2N
#define PWC_BWD_NA2N 6   /* K2 without a flow: 6 A-row warps + 6 copy warps */
#endif
#include "common.cuh"
#include "fwd_generic.cuh"
#include "fwd_tile.cuh"
#include "fwd_mma16.cuh"
#ifdef PWC_HAVE_STREAM_KERNEL
#include "fwd_stream.cuh"
#endif

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_fwd_impl{0};
std::atomic<int> g_stream_ctas{0};
std::atomic<int> g_stream_pf{4};
std::atomic<int> g_stream_dbg{0};
std::atomic<int> g_mma_tile{1};  // 16-bit band GEMM, r = 4: 0 = 8x32 tile with a tile-wide staging buffer, 1 = resident-halo variant
std::atomic<int> g_bwd_pf{1};    // fused backward: L2 prefetch distance of the producers (turns of their row loop)
std::atomic<int> g_bwd_dbg{0};   // fused backward: timing experiments only (see bwd_fused.cuh Args::dbg)
// levels at least this tall take the streaming kernel (it pays ~5 warm-up slabs per strip segment): measured
// 28 rows 84 us streaming vs 91 us tiled, 14 rows 56 vs 46 us
constexpr int kTallMinH = 28;
std::atomic<int> g_overlap{1};
std::atomic<int> g_bwd_tiled{1};
std::atomic<int> g_bwd_impl{0};   // 0 = auto, 1 = unfused v1 kernels, 2 = fused kernels (error when not covered)
std::atomic<int> g_tile_small{1};
std::atomic<int> g_side_reverse{1};
std::atomic<int> g_side_sms{16};          // SMs the tall levels' persistent kernels leave to the short levels of a pyramid pass
thread_local int t_reserve_sms = 0;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define PWC_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) return fail(PWC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

#define PWC_LAUNCH_CHECK()                                                                      \
  do {                                                                                          \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                         \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) return fail(PWC_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e__)); \
  } while (0)

struct DevInfo {
  int major = -1, minor = 0, sms = 0;
};
DevInfo g_dev[64];
std::mutex g_dev_mu;

int device_info(DevInfo* out) {
  int dev = 0;
  PWC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(PWC_ERR_ARCH, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (g_dev[dev].major < 0) {
    DevInfo d;
    PWC_CUDA(cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev));
    PWC_CUDA(cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, dev));
    PWC_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    g_dev[dev] = d;
  }
  *out = g_dev[dev];
  if (out->major != 10)
    return fail(PWC_ERR_ARCH, "libpwc_b200 is built for sm_100a only; device %d is sm_%d%d (no fallback path)", dev,
                out->major, out->minor);
  return PWC_OK;
}

size_t elt_size(int dtype) { return dtype == PWC_F32 ? 4 : 2; }
bool dtype_ok(int dtype) { return dtype == PWC_F32 || dtype == PWC_F16 || dtype == PWC_BF16; }
bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
size_t up256(size_t n) { return (n + 255) & ~size_t(255); }

int check_shape(int B, int H, int W, int C, int r, bool has_flow) {
  if (B < 0 || H < 0 || W < 0 || C <= 0 || r < 0) return fail(PWC_ERR_BAD_SHAPE, "negative dimension");
  if ((int64_t)B * H * W >= (int64_t)1 << 31) return fail(PWC_ERR_BAD_SHAPE, "B*H*W must be < 2^31");
  if (has_flow && (H < 2 || W < 2))
    return fail(PWC_ERR_BAD_SHAPE, "dense_image_warp needs height >= 2 and width >= 2 (got %dx%d)", H, W);
  if (r > 64) return fail(PWC_ERR_BAD_SHAPE, "search_range %d too large", r);
  return PWC_OK;
}

int grid_for(int64_t total, int block, int sms) {
  int64_t g = (total + block - 1) / block;
  const int64_t cap = (int64_t)sms * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- forward dispatch ----------------------------------------------------------------
template <typename T, typename TF>
int fwd_typed(const void* c1, const void* c2, const void* flow, float flow_scale, void* out, int B, int H, int W,
              int C, int r, int64_t ops, cudaStream_t st, const DevInfo& dev) {
  const int n = 2 * r + 1;
  const int64_t D = (int64_t)n * n;
  if (ops == 0) ops = D;
  if (ops < D) return fail(PWC_ERR_BAD_ARG, "out_pixel_stride %lld < (2r+1)^2", (long long)ops);
  if ((int64_t)B * H * W == 0) return PWC_OK;
  const size_t va = sizeof(T) * 4;  // vector alignment for 4-element loads
  const bool vec_ok = (C % 4 == 0) && aligned(c1, va) && aligned(c2, va);
  int impl = g_fwd_impl.load();
  const bool tuned_ok = (r == 4 || r == 8) && vec_ok;
  // auto: the streaming kernel wins on tall levels (it pays ~5 warm-up slabs per strip segment), the tiled kernel
  // on short ones; measured crossover between H = 28 and H = 56 (profiles/r01_notes.md)
  if (impl == 0) impl = !tuned_ok ? 1 : (r == 4 && H >= kTallMinH ? 3 : 2);
  // 16-bit features, C = 32 / 64 / 96 / 128: the tensor-core band GEMM beats the widening CUDA-core loops (level 2
  // fp16: 216 vs 277 us, level 3: 147 vs 160 us; same error against the fp16-regime oracle); shapes it does not cover
  // fall back to the choice above
  const int impl_fallback = impl;
  if (g_fwd_impl.load() == 0 && sizeof(T) == 2 && tuned_ok && r == 4 &&
      (g_mma_tile.load() >= 1 ? (C == 32 || C == 64 || C == 96 || C == 128) : ((C == 32 || C == 64) && H >= kTallMinH)))
    impl = 4;   // any height: level 5 (14 rows, C = 128) takes 24 us against 56 us on the tiled kernel
  // search_range 8 (289 channels, configs[4]): 17 of 32 diagonals of the band GEMM are used, the CUDA-core tile kernel
  // is FMA-bound at 11 % of its peak
  // (small levels stay on the tile kernel: 2x34x60x32 takes 37 us there, 52 us here)
  if (g_fwd_impl.load() == 0 && sizeof(T) == 2 && tuned_ok && r == 8 && (C == 32 || C == 64 || C == 96 || C == 128) &&
      (int64_t)B * H * W >= 16384)
    impl = 4;
  if (impl == 3 && r != 4) impl = 2;  // the streaming kernel is written for search_range 4
  // un-warped small level (the pyramid top): too few tiles for the tiled kernel, one thread per output element wins
  if (g_fwd_impl.load() == 0 && flow == nullptr && vec_ok && (int64_t)B * H * W <= 16384) {
    if (r == 4 && W % 2 == 0) {
      const int64_t total = (int64_t)B * H * (W / 2) * 27;
      pwc::fwd_noflow_r4_tiled_kernel<T><<<grid_for(total, 256, dev.sms), 256, 0, st>>>((const T*)c1, (const T*)c2, (T*)out,
                                                                                      B, H, W, C, ops);
      PWC_LAUNCH_CHECK();
      return PWC_OK;
    }
    const int64_t total = (int64_t)B * H * W * D;
    pwc::fwd_noflow_vec4_kernel<T><<<grid_for(total, 256, dev.sms), 256, 0, st>>>((const T*)c1, (const T*)c2, (T*)out, B, H,
                                                                                W, C, r, ops);
    PWC_LAUNCH_CHECK();
    return PWC_OK;
  }
  if (impl >= 2 && !tuned_ok) impl = 1;
  if (impl == 4) {  // 16-bit band GEMM on the tensor cores (fwd_mma16.cuh): r = 4 or 8, C in {32, 64}
    int rc = 1;
    if constexpr (sizeof(T) == 2)
    {
      if (r == 4) {
        const int mt = g_mma_tile.load();
        rc = 1;
        if (rc == 1 && mt >= 1) rc = pwc::mma16::launch_band<4, T, TF>((const T*)c1, (const T*)c2, (const TF*)flow, flow_scale, (T*)out, B, H, W, C, ops, st);
        if (rc == 1 && mt == 0) rc = pwc::mma16::launch<T, TF>((const T*)c1, (const T*)c2, (const TF*)flow, flow_scale, (T*)out, B, H, W, C, ops, st);
      }
      if (r == 8) rc = pwc::mma16::launch_r8<T, TF>((const T*)c1, (const T*)c2, (const TF*)flow, flow_scale, (T*)out, B, H, W, C, ops, st);
    }
    if (rc == 0) { PWC_LAUNCH_CHECK(); return PWC_OK; }
    if (rc < 0) return fail(PWC_ERR_CUDA, "mma16 kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    impl = g_fwd_impl.load() == 0 ? impl_fallback : (r == 4 && H >= kTallMinH ? 3 : 2);  // shape not covered
  }
#ifdef PWC_HAVE_STREAM_KERNEL
  if (impl == 3) {
    int rc = 1;
    // instantiated for the dtype pairs the reference uses (fp32/fp32, fp16/fp16); others take the tiled kernel
    if constexpr ((std::is_same<T, float>::value && std::is_same<TF, float>::value) ||
                  (std::is_same<T, __half>::value && std::is_same<TF, __half>::value)) {
      int ctas = g_stream_ctas.load();
      if (ctas <= 0 && t_reserve_sms > 0 && t_reserve_sms < dev.sms) ctas = dev.sms - t_reserve_sms;
      rc = pwc::stream::launch<T, TF>((const T*)c1, (const T*)c2, (const TF*)flow, flow_scale, (T*)out, B, H, W, C,
                                      ops, st, dev.sms, ctas, g_stream_pf.load(), g_stream_dbg.load());
    }
    if (rc == 0) { PWC_LAUNCH_CHECK(); return PWC_OK; }
    if (rc < 0) return fail(PWC_ERR_CUDA, "stream kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    impl = 2;  // rc > 0: shape not covered
  }
#else
  if (impl == 3) impl = 2;
#endif
  if (impl == 2) {
    auto launch_tile = [&](auto cfg, auto kern, int slot) -> int {
      using K = decltype(cfg);
      static std::atomic<uint64_t> attr_set[3];  // per (T,TF) instantiation and tile shape; one bit per device
      int devid = 0;
      PWC_CUDA(cudaGetDevice(&devid));
      if (!((attr_set[slot].load() >> devid) & 1)) {
        PWC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES));
        // co-resident CTAs per SM hide each other's fill/compute phases
        PWC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        if (getenv("PWC_DEBUG_OCC")) {
          int nb = 0;
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, K::NTHREADS, K::SMEM_BYTES);
          fprintf(stderr, "[pwc] tile kernel %dx%d occupancy: %d CTAs/SM\n", K::TH, K::TW, nb);
        }
        attr_set[slot].fetch_or(uint64_t(1) << devid);
      }
      const int tiles_x = (W + K::TW - 1) / K::TW, tiles_y = (H + K::TH - 1) / K::TH;
      const int64_t tiles = (int64_t)tiles_x * tiles_y * B;
      if (tiles >= (int64_t)1 << 31) return fail(PWC_ERR_BAD_SHAPE, "too many tiles");
      kern<<<(unsigned)tiles, K::NTHREADS, K::SMEM_BYTES, st>>>((const T*)c1, (const T*)c2, (const TF*)flow, flow_scale,
                                                               (T*)out, B, H, W, C, ops, tiles_x, tiles_y);
      PWC_LAUNCH_CHECK();
      return PWC_OK;
    };
    // short levels: fewer full-size tiles than SMs -> the small tile shape (more, lighter CTAs; measured: level 5
    // 76 -> 46 us, while level 4 with 256 full-size tiles is faster as it is: 91 vs 119 us)
    if (r == 8) return launch_tile(pwc::tile::Cfg<8, 16, 2>(), pwc::tile::fwd_tile_kernel<8, 16, 2, T, TF>, 2);
    const int64_t big_tiles = (int64_t)((W + 31) / 32) * ((H + 7) / 8) * B;
    if (big_tiles < (int64_t)dev.sms && g_tile_small.load())
      return launch_tile(pwc::tile::Cfg<4, 16, 2>(), pwc::tile::fwd_tile_kernel<4, 16, 2, T, TF>, 1);
    return launch_tile(pwc::tile::Cfg<4, 32, 4>(), pwc::tile::fwd_tile_kernel<4, 32, 4, T, TF>, 0);
  }
  const int64_t total = (int64_t)B * H * W * D;
  pwc::fwd_generic_kernel<T, TF><<<grid_for(total, 256, dev.sms), 256, 0, st>>>(
      (const T*)c1, (const T*)c2, (const TF*)flow, flow_scale, (T*)out, B, H, W, C, r, ops);
  PWC_LAUNCH_CHECK();
  return PWC_OK;
}

template <typename... A>
int fwd_dispatch(int dtype, int flow_dtype, const void* flow, A... a) {
  // flow dtype must be the feature dtype or fp32 (the reference allows any mix; these are the ones it uses)
  if (flow == nullptr) flow_dtype = dtype;
  if (dtype == PWC_F32 && flow_dtype == PWC_F32) return fwd_typed<float, float>(a...);
  if (dtype == PWC_F16 && flow_dtype == PWC_F16) return fwd_typed<__half, __half>(a...);
  if (dtype == PWC_F16 && flow_dtype == PWC_F32) return fwd_typed<__half, float>(a...);
  if (dtype == PWC_BF16 && flow_dtype == PWC_BF16) return fwd_typed<__nv_bfloat16, __nv_bfloat16>(a...);
  if (dtype == PWC_BF16 && flow_dtype == PWC_F32) return fwd_typed<__nv_bfloat16, float>(a...);
  return fail(PWC_ERR_BAD_DTYPE, "unsupported (feature dtype %d, flow dtype %d) combination", dtype, flow_dtype);
}

int fwd_entry(const void* c1, const void* c2, const void* flow, float flow_scale, void* out, int B, int H, int W,
              int C, int r, int dtype, int flow_dtype, int64_t ops, void* stream) {
  if (!dtype_ok(dtype) || (flow && !dtype_ok(flow_dtype))) return fail(PWC_ERR_BAD_DTYPE, "unknown dtype");
  int rc = check_shape(B, H, W, C, r, flow != nullptr);
  if (rc) return rc;
  if ((int64_t)B * H * W > 0 && (!c1 || !c2 || !out)) return fail(PWC_ERR_BAD_ARG, "null tensor pointer");
  if (!aligned(c1, elt_size(dtype)) || !aligned(c2, elt_size(dtype)) || !aligned(out, elt_size(dtype)) ||
      (flow && !aligned(flow, elt_size(flow_dtype))))
    return fail(PWC_ERR_ALIGN, "tensor pointer not aligned to its element size");
  DevInfo dev;
  rc = device_info(&dev);
  if (rc) return rc;
  return fwd_dispatch(dtype, flow_dtype, flow, c1, c2, flow, flow_scale, out, B, H, W, C, r, ops,
                      (cudaStream_t)stream, dev);
}

// ---- standalone warp -------------------------------------------------------------------
template <typename T, typename TF>
int warp_fwd_typed(const void* img, const void* flow, float scale, void* out, int B, int H, int W, int C,
                   cudaStream_t st, const DevInfo& dev) {
  if ((int64_t)B * H * W == 0) return PWC_OK;
  const size_t va = sizeof(T) * 4;
  if (C % 4 == 0 && aligned(img, va) && aligned(out, va)) {
    const int64_t total = (int64_t)B * H * W * (C / 4);
    pwc::warp_fwd_vec4_kernel<T, TF><<<grid_for(total, 256, dev.sms), 256, 0, st>>>((const T*)img, (const TF*)flow,
                                                                                   scale, (T*)out, B, H, W, C);
  } else {
    const int64_t total = (int64_t)B * H * W * C;
    pwc::warp_fwd_scalar_kernel<T, TF><<<grid_for(total, 256, dev.sms), 256, 0, st>>>((const T*)img, (const TF*)flow,
                                                                                     scale, (T*)out, B, H, W, C);
  }
  PWC_LAUNCH_CHECK();
  return PWC_OK;
}

int warp_fwd_entry(const void* img, const void* flow, float scale, void* out, int B, int H, int W, int C, int dtype,
                   int flow_dtype, cudaStream_t st) {
  if (!dtype_ok(dtype) || !dtype_ok(flow_dtype)) return fail(PWC_ERR_BAD_DTYPE, "unknown dtype");
  int rc = check_shape(B, H, W, C, 0, true);
  if (rc) return rc;
  if ((int64_t)B * H * W > 0 && (!img || !flow || !out)) return fail(PWC_ERR_BAD_ARG, "null tensor pointer");
  if (!aligned(img, elt_size(dtype)) || !aligned(out, elt_size(dtype)) || !aligned(flow, elt_size(flow_dtype)))
    return fail(PWC_ERR_ALIGN, "tensor pointer not aligned to its element size");
  DevInfo dev;
  rc = device_info(&dev);
  if (rc) return rc;
  if (dtype == PWC_F32 && flow_dtype == PWC_F32) return warp_fwd_typed<float, float>(img, flow, scale, out, B, H, W, C, st, dev);
  if (dtype == PWC_F16 && flow_dtype == PWC_F16) return warp_fwd_typed<__half, __half>(img, flow, scale, out, B, H, W, C, st, dev);
  if (dtype == PWC_F16 && flow_dtype == PWC_F32) return warp_fwd_typed<__half, float>(img, flow, scale, out, B, H, W, C, st, dev);
  if (dtype == PWC_BF16 && flow_dtype == PWC_BF16) return warp_fwd_typed<__nv_bfloat16, __nv_bfloat16>(img, flow, scale, out, B, H, W, C, st, dev);
  if (dtype == PWC_BF16 && flow_dtype == PWC_F32) return warp_fwd_typed<__nv_bfloat16, float>(img, flow, scale, out, B, H, W, C, st, dev);
  return fail(PWC_ERR_BAD_DTYPE, "unsupported (image dtype %d, flow dtype %d) combination", dtype, flow_dtype);
}

// ---- backward --------------------------------------------------------------------------
template <int MODE, typename T, typename TO>
int bwd_corr_launch(const void* g, const void* out, const void* other, void* dst, int B, int H, int W, int C, int r,
                    int64_t gps, int64_t ops, cudaStream_t st, const DevInfo& dev) {
  const size_t va = sizeof(T) * 4;
  if (r == 4 && C % 8 == 0 && W % 4 == 0 && aligned(other, va) && aligned(dst, sizeof(TO) * 4) && g_bwd_tiled.load()) {
    const int64_t total = (int64_t)B * H * (W / 4) * (C / 8);
    pwc::bwd_corr_r4_tiled_kernel<MODE, T, TO><<<grid_for(total, 256, dev.sms), 256, 0, st>>>(
        (const T*)g, (const T*)out, (const T*)other, (TO*)dst, B, H, W, C, gps, ops);
  } else if (C % 4 == 0 && aligned(other, va) && aligned(dst, sizeof(TO) * 4)) {
    const int64_t total = (int64_t)B * H * W * (C / 4);
    pwc::bwd_corr_kernel<MODE, 4, T, TO><<<grid_for(total, 256, dev.sms), 256, 0, st>>>(
        (const T*)g, (const T*)out, (const T*)other, (TO*)dst, B, H, W, C, r, gps, ops);
  } else {
    const int64_t total = (int64_t)B * H * W * C;
    pwc::bwd_corr_kernel<MODE, 1, T, TO><<<grid_for(total, 256, dev.sms), 256, 0, st>>>(
        (const T*)g, (const T*)out, (const T*)other, (TO*)dst, B, H, W, C, r, gps, ops);
  }
  PWC_LAUNCH_CHECK();
  return PWC_OK;
}

template <typename T, typename TF, typename TG>
int warp_bwd_launch(const void* gout, const void* img, const void* flow, float scale, float* dimg_acc, void* dflow,
                    int B, int H, int W, int C, cudaStream_t st, const DevInfo& dev) {
  const int64_t npix = (int64_t)B * H * W;
  if (C % 4 == 0 && aligned(gout, sizeof(TG) * 4) && aligned(img, sizeof(T) * 4) && aligned(dimg_acc, 16)) {
    const int lpp = C <= 32 ? 8 : (C <= 64 ? 16 : 32);  // lanes per pixel, 4 channels per lane and step
    const int grid = grid_for(npix * lpp, 256, dev.sms);
    auto go = [&](auto kern) {
      kern<<<grid, 256, 0, st>>>((const TG*)gout, (const T*)img, (const TF*)flow, scale, dimg_acc, (TF*)dflow, B, H, W, C);
    };
    if (lpp == 8) go(pwc::warp_bwd_kernel<4, 8, T, TF, TG>);
    else if (lpp == 16) go(pwc::warp_bwd_kernel<4, 16, T, TF, TG>);
    else go(pwc::warp_bwd_kernel<4, 32, T, TF, TG>);
  } else {
    const int grid = grid_for(npix * 32, 256, dev.sms);
    pwc::warp_bwd_kernel<1, 32, T, TF, TG><<<grid, 256, 0, st>>>((const TG*)gout, (const T*)img, (const TF*)flow, scale,
                                                                dimg_acc, (TF*)dflow, B, H, W, C);
  }
  PWC_LAUNCH_CHECK();
  return PWC_OK;
}

template <typename T>
int cast_launch(const float* src, void* dst, int64_t n, cudaStream_t st, const DevInfo& dev) {
  pwc::cast_from_f32_kernel<T><<<grid_for(n, 256, dev.sms), 256, 0, st>>>(src, (T*)dst, n);
  PWC_LAUNCH_CHECK();
  return PWC_OK;
}

struct BwdWs {
  size_t warp_off, dwarp_off, acc_off, total;
};
BwdWs bwd_ws_layout(int B, int H, int W, int C, int dtype, bool has_flow) {
  const size_t n = (size_t)B * H * W * C;
  BwdWs w{};
  size_t off = 0;
  w.warp_off = off;
  if (has_flow) off += up256(n * elt_size(dtype));
  w.dwarp_off = off;
  if (has_flow) off += up256(n * 4);
  w.acc_off = off;
  if (has_flow && dtype != PWC_F32) off += up256(n * 4);
  w.total = off;
  return w;
}

template <typename T>
int cv_bwd_typed(const void* g, const void* out, const void* c1, const void* warp, void* dc1, void* dwarp, int B, int H,
                 int W, int C, int r, int64_t gps, int64_t ops, cudaStream_t st, const DevInfo& dev) {
  int rc = bwd_corr_launch<0, T, T>(g, out, warp, dc1, B, H, W, C, r, gps, ops, st, dev);
  if (rc) return rc;
  return bwd_corr_launch<1, T, T>(g, out, c1, dwarp, B, H, W, C, r, gps, ops, st, dev);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (the library links cudart statically and not
// libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int encode_tensor_map(CUtensorMap* tm, void* base, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                      const cuuint32_t* estr, CUtensorMapSwizzle swz) {
  static std::atomic<EncodeTiledFn> fn{nullptr};
  EncodeTiledFn f = fn.load();
  if (f == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr ||
        q != cudaDriverEntryPointSuccess)
      return fail(PWC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    f = reinterpret_cast<EncodeTiledFn>(p);
    fn.store(f);
  }
  const CUresult r = f(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PWC_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return PWC_OK;
}

// Fused backward (bwd_fused.cuh): K1 = dc1 + transposed g' workspace, K2 = dwarp -> dc2 / dflow.  Two kernel launches
// per level (plus a memset of the scatter target, and a cast for 16-bit tensors).
template <typename T, typename TF>
int bwd_fused_launch(const void* g, const void* out, const void* c1, const void* c2, const void* flow, float scale,
                     void* dc1, void* dc2, void* dflow, int B, int H, int W, int C, void* ws, size_t ws_bytes,
                     cudaStream_t st, const DevInfo& dev) {
  namespace bf = pwc::bwdf;
  const bool hf = flow != nullptr;
  const bf::Plan pl = bf::make_plan(B, H, W, C, sizeof(T), sizeof(TF), hf, dev.sms);
  if (ws_bytes < pl.total || ws == nullptr)
    return fail(PWC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", pl.total, ws_bytes);
  if (!aligned(ws, 256)) return fail(PWC_ERR_ALIGN, "workspace must be 256-byte aligned");
  char* base = (char*)ws;
  const size_t n = (size_t)B * H * W * C, npix = (size_t)B * H * W;
  bf::Args a{};
  a.g = g; a.out = out; a.flow = flow; a.c2 = c2;
  a.ws = (float*)(base + pl.ws_off);
  a.B = B; a.H = H; a.W = W; a.C = C;
  a.nsx = pl.nsx; a.nslices = pl.nslices; a.NPAIR = pl.npair; a.Q = pl.Q; a.wpitch = pl.wpitch;
  a.flow_scale = scale; a.inv_c = 1.0f / (float)C;
  a.dbg = g_bwd_dbg.load();
  a.pf = g_bwd_pf.load();
  if ((long long)B * pl.nsx * pl.nslices >= (1LL << 31)) return fail(PWC_ERR_BAD_SHAPE, "too many strips");
  const int grid = pl.grid;  // persistent: one CTA per SM walks its share of the (strip, row pair) units
  auto prep = [&](auto kern) -> int {
    PWC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf::SMEM_BYTES));
    return PWC_OK;
  };
  // warp split (A producers of 12 non-consumer warps): K1 with a flow also gathers the warped rows, K2 mostly shifts
  // the transposed g' rows into place
  constexpr int NA1 = PWC_BWD_NA1, NA1N = PWC_BWD_NA1N, NA2 = PWC_BWD_NA2, NA2N = PWC_BWD_NA2N;
  auto k1 = hf ? bf::bwd_fused_kernel<0, true, NA1, T, TF> : bf::bwd_fused_kernel<0, false, NA1N, T, TF>;
  auto k2 = hf ? bf::bwd_fused_kernel<1, true, NA2, T, TF> : bf::bwd_fused_kernel<1, false, NA2N, T, TF>;
  int rc = prep(k1);
  if (rc) return rc;
  if ((rc = prep(k2))) return rc;
  // the scatter targets of K2 (red.global.add) are zeroed between the two kernels, not before K1: the zeros are then
  // still in L2 when the atomics arrive (before K1 they were evicted by its 1.2 GB of traffic and every atomic paid a
  // DRAM read)
  float* acc = nullptr;
  float* dfacc = nullptr;
  if (hf) {
    acc = pl.acc_in_ws ? (float*)(base + pl.acc_off) : (float*)dc2;
    if (pl.nslices > 1) dfacc = pl.dflow_in_ws ? (float*)(base + pl.dflow_off) : (float*)dflow;
  }
  // the g' workspace as a 4-d tensor [B * H][dyi][dxi][x] (see bwd_fused.cuh): K1 stores 32 px x 9-plane boxes through the
  // 128-byte swizzle, K2 loads 32 px x 9 dyi boxes (and two 4 px x 3 plane x 9 dyi remainder boxes) at multiples of 4 px
  CUtensorMap tm_store, tm_main, tm_rem;
  {
    const cuuint64_t wp = (cuuint64_t)pl.wpitch;
    const cuuint64_t dims[4] = {(cuuint64_t)W, 9, 9, (cuuint64_t)B * H};
    const cuuint64_t strides[3] = {wp * 4, wp * 4 * 9, wp * 4 * 81};
    const cuuint32_t box_store[4] = {32, 9, 1, 1}, box_main[4] = {32, 1, 9, 1}, box_rem[4] = {4, 3, 9, 1}, estr[4] = {1, 1, 1, 1};
    if ((rc = encode_tensor_map(&tm_store, a.ws, dims, strides, box_store, estr, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = encode_tensor_map(&tm_main, a.ws, dims, strides, box_main, estr, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
    if ((rc = encode_tensor_map(&tm_rem, a.ws, dims, strides, box_rem, estr, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  }
  a.feat = c2; a.dst = dc1;
  k1<<<(unsigned)grid, bf::NTHREADS, bf::Lay<0>::SMEM_BYTES, st>>>(a, tm_store, tm_store);
  PWC_LAUNCH_CHECK();
  if (acc) PWC_CUDA(cudaMemsetAsync(acc, 0, n * 4, st));
  if (dfacc) PWC_CUDA(cudaMemsetAsync(dfacc, 0, npix * 2 * 4, st));
  a.feat = c1; a.dst = dc2; a.acc = acc; a.dflow = dflow; a.dflow_acc = dfacc;
  k2<<<(unsigned)grid, bf::NTHREADS, bf::Lay<1>::SMEM_BYTES, st>>>(a, tm_main, tm_rem);
  PWC_LAUNCH_CHECK();
  if (hf && pl.acc_in_ws && (rc = cast_launch<T>(acc, dc2, (int64_t)n, st, dev))) return rc;
  if (hf && dfacc && pl.dflow_in_ws && (rc = cast_launch<TF>(dfacc, dflow, (int64_t)npix * 2, st, dev))) return rc;
  return PWC_OK;
}

template <typename T, typename TF>
int fused_bwd_typed(const void* g, const void* out, const void* c1, const void* c2, const void* flow, float scale,
                    void* dc1, void* dc2, void* dflow, int B, int H, int W, int C, int r, int dtype, int64_t gps,
                    int64_t ops, void* ws, size_t ws_bytes, cudaStream_t st, const DevInfo& dev) {
  {
    const int impl = g_bwd_impl.load();
    const bool ok = pwc::bwdf::covered(C, r, gps, ops, c1, c2, dc1, dc2, sizeof(T));
    if (impl == 2 && !ok)
      return fail(PWC_ERR_BAD_ARG, "bwd_impl=2: the fused backward needs r == 4, C %% 4 == 0 (C %% 8 for 16-bit tensors), dense g / out and 16-byte aligned tensors");
    // auto: below ~1 K pixels the two persistent kernels' set-up costs more than the unfused launches (8x6x8x196: 38 vs 34 us)
    if (impl != 1 && ok && (impl == 2 || (int64_t)B * H * W >= 1024))
      return bwd_fused_launch<T, TF>(g, out, c1, c2, flow, scale, dc1, dc2, dflow, B, H, W, C, ws, ws_bytes, st, dev);
  }
  if (flow == nullptr) return cv_bwd_typed<T>(g, out, c1, c2, dc1, dc2, B, H, W, C, r, gps, ops, st, dev);
  const BwdWs L = bwd_ws_layout(B, H, W, C, dtype, true);
  if (ws_bytes < L.total || ws == nullptr)
    return fail(PWC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", L.total, ws_bytes);
  if (!aligned(ws, 256)) return fail(PWC_ERR_ALIGN, "workspace must be 256-byte aligned");
  char* base = (char*)ws;
  void* wbuf = base + L.warp_off;
  float* dwbuf = (float*)(base + L.dwarp_off);
  const size_t n = (size_t)B * H * W * C;
  int rc = warp_fwd_typed<T, TF>(c2, flow, scale, wbuf, B, H, W, C, st, dev);
  if (rc) return rc;
  rc = bwd_corr_launch<0, T, T>(g, out, wbuf, dc1, B, H, W, C, r, gps, ops, st, dev);
  if (rc) return rc;
  rc = bwd_corr_launch<1, T, float>(g, out, c1, dwbuf, B, H, W, C, r, gps, ops, st, dev);
  if (rc) return rc;
  float* acc = sizeof(T) == 4 ? (float*)dc2 : (float*)(base + L.acc_off);
  PWC_CUDA(cudaMemsetAsync(acc, 0, n * 4, st));
  rc = warp_bwd_launch<T, TF, float>(dwbuf, c2, flow, scale, acc, dflow, B, H, W, C, st, dev);
  if (rc) return rc;
  if (sizeof(T) != 4) rc = cast_launch<T>(acc, dc2, (int64_t)n, st, dev);
  return rc;
}

template <typename T, typename TF>
int warp_bwd_typed(const void* gout, const void* img, const void* flow, void* dimg, void* dflow, int B, int H, int W,
                   int C, void* ws, size_t ws_bytes, cudaStream_t st, const DevInfo& dev) {
  const size_t n = (size_t)B * H * W * C;
  float* acc = (float*)dimg;
  if (sizeof(T) != 4) {
    if (ws == nullptr || ws_bytes < up256(n * 4))
      return fail(PWC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", up256(n * 4), ws_bytes);
    if (!aligned(ws, 256)) return fail(PWC_ERR_ALIGN, "workspace must be 256-byte aligned");
    acc = (float*)ws;
  }
  PWC_CUDA(cudaMemsetAsync(acc, 0, n * 4, st));
  int rc = warp_bwd_launch<T, TF, T>(gout, img, flow, 1.0f, acc, dflow, B, H, W, C, st, dev);
  if (rc) return rc;
  if (sizeof(T) != 4) rc = cast_launch<T>(acc, dimg, (int64_t)n, st, dev);
  return rc;
}

}  // namespace

// =========================================================================================
extern "C" {

int pwc_abi_version(void) { return PWC_ABI_VERSION; }
const char* pwc_last_error(void) { return g_err.c_str(); }
uint64_t pwc_launch_count(void) { return g_launches.load(); }

int pwc_set_option(const char* key, int value) {
  if (!key) return fail(PWC_ERR_BAD_ARG, "null option key");
  if (!strcmp(key, "fwd_impl")) {
    if (value < 0 || value > 4) return fail(PWC_ERR_BAD_ARG, "fwd_impl must be 0..4");
    g_fwd_impl.store(value);
    return PWC_OK;
  }
  if (!strcmp(key, "stream_ctas")) {
    g_stream_ctas.store(value);
    return PWC_OK;
  }
  if (!strcmp(key, "stream_prefetch")) {
    g_stream_pf.store(value);
    return PWC_OK;
  }
  if (!strcmp(key, "overlap_levels")) {
    g_overlap.store(value);
    return PWC_OK;
  }
  if (!strcmp(key, "bwd_impl")) {
    if (value < 0 || value > 2) return fail(PWC_ERR_BAD_ARG, "bwd_impl must be 0..2");
    g_bwd_impl.store(value);
    return PWC_OK;
  }
  if (!strcmp(key, "bwd_tiled")) {  // 0: one-pixel-per-thread backward correlation kernels
    g_bwd_tiled.store(value);
    return PWC_OK;
  }
  if (!strcmp(key, "side_reverse")) {  // order in which the short levels are launched on the side streams
    g_side_reverse.store(value);
    return PWC_OK;
  }
  if (!strcmp(key, "side_sms")) {  // SMs left to the short levels while the tall levels' persistent kernels run
    if (value < 0 || value > 64) return fail(PWC_ERR_BAD_ARG, "side_sms must be 0..64");
    g_side_sms.store(value);
    return PWC_OK;
  }
  if (!strcmp(key, "tile_small")) {  // 0: always the 8x32 tile shape
    g_tile_small.store(value);
    return PWC_OK;
  }
  if (!strcmp(key, "stream_debug")) {  // timing experiments only; results are wrong when non-zero
    g_stream_dbg.store(value);
    return PWC_OK;
  }
  if (!strcmp(key, "mma_tile")) {
    if (value < 0 || value > 1) return fail(PWC_ERR_BAD_ARG, "mma_tile must be 0 or 1");
    g_mma_tile.store(value);
    return PWC_OK;
  }
  if (!strcmp(key, "bwd_prefetch")) {
    if (value < 0 || value > 8) return fail(PWC_ERR_BAD_ARG, "bwd_prefetch must be 0..8");
    g_bwd_pf.store(value);
    return PWC_OK;
  }
  if (!strcmp(key, "bwd_debug")) {  // timing experiments only; results are wrong when non-zero
    g_bwd_dbg.store(value);
    return PWC_OK;
  }
  return fail(PWC_ERR_BAD_ARG, "unknown option '%s'", key);
}
int pwc_debug_read_trace(void* dst_host, int max_bytes) {
#ifdef PWC_HAVE_STREAM_KERNEL
  if (!dst_host || max_bytes <= 0) return fail(PWC_ERR_BAD_ARG, "null argument");
  size_t n = sizeof(unsigned long long) * (pwc::stream::NCW + pwc::stream::NPW) * pwc::stream::TRACE_LEN;
  if ((size_t)max_bytes < n) n = (size_t)max_bytes;
  if (cudaDeviceSynchronize() != cudaSuccess ||
      cudaMemcpyFromSymbol(dst_host, pwc::stream::g_trace, n, 0, cudaMemcpyDeviceToHost) != cudaSuccess)
    return fail(PWC_ERR_CUDA, "trace copy failed: %s", cudaGetErrorString(cudaGetLastError()));
  return (int)n;
#else
  return fail(PWC_ERR_BAD_ARG, "built without the streaming kernel");
#endif
}
int pwc_get_option(const char* key, int* value) {
  if (!key || !value) return fail(PWC_ERR_BAD_ARG, "null argument");
  if (!strcmp(key, "fwd_impl")) { *value = g_fwd_impl.load(); return PWC_OK; }
  if (!strcmp(key, "stream_ctas")) { *value = g_stream_ctas.load(); return PWC_OK; }
  if (!strcmp(key, "stream_prefetch")) { *value = g_stream_pf.load(); return PWC_OK; }
  if (!strcmp(key, "stream_debug")) { *value = g_stream_dbg.load(); return PWC_OK; }
  if (!strcmp(key, "bwd_debug")) { *value = g_bwd_dbg.load(); return PWC_OK; }
  if (!strcmp(key, "bwd_prefetch")) { *value = g_bwd_pf.load(); return PWC_OK; }
  if (!strcmp(key, "mma_tile")) { *value = g_mma_tile.load(); return PWC_OK; }
  if (!strcmp(key, "overlap_levels")) { *value = g_overlap.load(); return PWC_OK; }
  if (!strcmp(key, "bwd_impl")) { *value = g_bwd_impl.load(); return PWC_OK; }
  if (!strcmp(key, "side_sms")) { *value = g_side_sms.load(); return PWC_OK; }
  if (!strcmp(key, "bwd_tiled")) { *value = g_bwd_tiled.load(); return PWC_OK; }
  if (!strcmp(key, "side_reverse")) { *value = g_side_reverse.load(); return PWC_OK; }
  if (!strcmp(key, "tile_small")) { *value = g_tile_small.load(); return PWC_OK; }
  return fail(PWC_ERR_BAD_ARG, "unknown option '%s'", key);
}

int pwc_cost_volume_fwd(const void* c1, const void* warp, void* out, int B, int H, int W, int C, int r, int dtype,
                        int64_t out_pixel_stride, void* cuda_stream) {
  return fwd_entry(c1, warp, nullptr, 1.0f, out, B, H, W, C, r, dtype, dtype, out_pixel_stride, cuda_stream);
}

int pwc_dense_image_warp_fwd(const void* image, const void* flow, void* out, int B, int H, int W, int C, int img_dtype,
                             int flow_dtype, void* cuda_stream) {
  return warp_fwd_entry(image, flow, 1.0f, out, B, H, W, C, img_dtype, flow_dtype, (cudaStream_t)cuda_stream);
}

int pwc_warp_cost_volume_fwd(const void* c1, const void* c2, const void* flow, float flow_scale, void* out, int B,
                             int H, int W, int C, int r, int dtype, int flow_dtype, int64_t out_pixel_stride,
                             void* cuda_stream) {
  return fwd_entry(c1, c2, flow, flow_scale, out, B, H, W, C, r, dtype, flow_dtype, out_pixel_stride, cuda_stream);
}

// The levels of one pass are independent operator calls (their inputs are given), and the short levels are
// latency-bound with far fewer CTAs than SMs, so they are forked onto side streams and joined at the end:
// cuda_stream still sees "all levels done" as one unit of stream-ordered work.
namespace {
struct SideStreams {
  cudaStream_t s[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // 0-2: short levels, 3-5: tall levels after the first
  cudaEvent_t fork = nullptr, join[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ok = false;
};
SideStreams* side_streams_for_device() {
  static thread_local SideStreams per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStreams& S = per_dev[dev];
  if (!S.ok) {
    for (int i = 0; i < 6; ++i) {
      if (cudaStreamCreateWithFlags(&S.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&S.join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&S.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    S.ok = true;
  }
  return &S;
}
}  // namespace

int pwc_pyramid_fwd(const pwc_level_t* levels, int n_levels, int B, int r, int dtype, int flow_dtype,
                    void* cuda_stream) {
  if (n_levels < 0 || (n_levels > 0 && !levels)) return fail(PWC_ERR_BAD_ARG, "bad level list");
  // every level is validated before anything is launched, so a bad argument cannot leave forked work behind
  if (!dtype_ok(dtype) || !dtype_ok(flow_dtype)) return fail(PWC_ERR_BAD_DTYPE, "unknown dtype");
  for (int i = 0; i < n_levels; ++i) {
    const pwc_level_t& L = levels[i];
    int rc = check_shape(B, L.H, L.W, L.C, r, L.flow != nullptr);
    if (rc) return rc;
    if ((int64_t)B * L.H * L.W > 0 && (!L.c1 || !L.c2 || !L.out)) return fail(PWC_ERR_BAD_ARG, "level %d: null tensor pointer", i);
    if (L.out_pixel_stride != 0 && L.out_pixel_stride < (int64_t)(2 * r + 1) * (2 * r + 1))
      return fail(PWC_ERR_BAD_ARG, "level %d: out_pixel_stride %lld < (2r+1)^2", i, (long long)L.out_pixel_stride);
  }
  cudaStream_t main_st = (cudaStream_t)cuda_stream;
  SideStreams* S = (n_levels > 1 && g_overlap.load()) ? side_streams_for_device() : nullptr;
  int n_side = 0, rc = PWC_OK;
  bool used[6] = {false, false, false, false, false, false};
  // The short levels are latency-bound (few CTAs each) and the tall levels' persistent kernels would otherwise own
  // every SM: the tall kernels leave `side_sms` SMs free, the first of them is launched (its CTAs are placed at once),
  // then the short levels are forked onto side streams - they can only land on the free SMs and run there NEXT TO
  // the tall levels instead of after them - and the remaining tall levels follow, in order, on the caller's stream.
  int n_short = 0;
  for (int i = 0; i < n_levels; ++i) n_short += (S && levels[i].H < kTallMinH) ? 1 : 0;
  const int reserve = n_short > 0 && n_short < n_levels ? g_side_sms.load() : 0;
  if (S && cudaEventRecord(S->fork, main_st) != cudaSuccess) return fail(PWC_ERR_CUDA, "cudaEventRecord failed");
  bool forked = false;
  auto fork_short = [&]() {
    forked = true;
    // largest first: its CTAs take whole SMs; the lighter kernels then fill what is left
    for (int ii = 0; ii < n_levels && rc == PWC_OK; ++ii) {
      const int i = g_side_reverse.load() ? n_levels - 1 - ii : ii;
      const pwc_level_t& L = levels[i];
      if (L.H >= kTallMinH) continue;
      const int k = n_side++ % 3;
      if (!used[k]) {
        if (cudaStreamWaitEvent(S->s[k], S->fork, 0) != cudaSuccess) { rc = fail(PWC_ERR_CUDA, "cudaStreamWaitEvent failed"); return; }
        used[k] = true;
      }
      rc = fwd_entry(L.c1, L.c2, L.flow, L.flow_scale, L.out, B, L.H, L.W, L.C, r, dtype, flow_dtype,
                     L.out_pixel_stride, S->s[k]);
    }
  };
  // overlap_levels 2: the tall levels after the first go to streams of their own as well - they are independent too, and
  // the CTAs of the next persistent kernel then start on the SMs the previous one's early finishers leave, instead of
  // after its last CTA
  int n_tall = 0;
  for (int i = 0; i < n_levels && rc == PWC_OK; ++i) {
    const pwc_level_t& L = levels[i];
    if (S && L.H < kTallMinH) continue;
    cudaStream_t st = main_st;
    if (S && g_overlap.load() >= 2 && n_tall > 0) {
      const int k = 3 + (n_tall - 1) % 3;
      if (!used[k]) {
        if (cudaStreamWaitEvent(S->s[k], S->fork, 0) != cudaSuccess) { rc = fail(PWC_ERR_CUDA, "cudaStreamWaitEvent failed"); break; }
        used[k] = true;
      }
      st = S->s[k];
    }
    ++n_tall;
    t_reserve_sms = reserve;
    rc = fwd_entry(L.c1, L.c2, L.flow, L.flow_scale, L.out, B, L.H, L.W, L.C, r, dtype, flow_dtype,
                   L.out_pixel_stride, st);
    t_reserve_sms = 0;
    if (S && !forked && rc == PWC_OK) fork_short();
  }
  if (S && !forked && rc == PWC_OK) fork_short();  // no tall level at all
  // join whatever was forked, also on the error path: the caller's stream must not be left racing side-stream work
  if (S)
    for (int k = 0; k < 6; ++k)
      if (used[k]) {
        if (cudaEventRecord(S->join[k], S->s[k]) != cudaSuccess || cudaStreamWaitEvent(main_st, S->join[k], 0) != cudaSuccess)
          if (rc == PWC_OK) rc = fail(PWC_ERR_CUDA, "joining a side stream failed");
      }
  return rc;
}

// ---- host-buffer path --------------------------------------------------------------------
struct pwc_host_ctx {
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_run;
  struct Buf { void* p = nullptr; size_t n = 0; };
  std::vector<Buf> bufs;  // 4 per level: c1, c2, flow, out
};

int pwc_host_ctx_destroy(pwc_host_ctx* c);

int pwc_host_ctx_create(pwc_host_ctx** ctx) {
  if (!ctx) return fail(PWC_ERR_BAD_ARG, "null ctx pointer");
  DevInfo dev;
  int rc = device_info(&dev);
  if (rc) return rc;
  pwc_host_ctx* c = new (std::nothrow) pwc_host_ctx();
  if (!c) return fail(PWC_ERR_BAD_ARG, "out of host memory");
  if (cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->s_run, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking) != cudaSuccess) {
    const cudaError_t e = cudaGetLastError();
    pwc_host_ctx_destroy(c);  // frees whatever was created
    return fail(PWC_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  *ctx = c;
  return PWC_OK;
}

int pwc_host_ctx_destroy(pwc_host_ctx* c) {
  if (!c) return PWC_OK;
  for (auto& b : c->bufs)
    if (b.p) cudaFree(b.p);
  for (auto e : c->ev_in) cudaEventDestroy(e);
  for (auto e : c->ev_run) cudaEventDestroy(e);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_run) cudaStreamDestroy(c->s_run);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  delete c;
  return PWC_OK;
}

static int ctx_reserve(pwc_host_ctx* c, size_t slot, size_t bytes) {
  if (c->bufs.size() <= slot) c->bufs.resize(slot + 1);
  auto& b = c->bufs[slot];
  if (b.n < bytes) {
    if (b.p) PWC_CUDA(cudaFree(b.p));
    b.p = nullptr;
    b.n = 0;
    PWC_CUDA(cudaMalloc(&b.p, bytes));
    b.n = bytes;
  }
  return PWC_OK;
}

int pwc_pyramid_fwd_host(pwc_host_ctx* c, const pwc_level_t* lv, int n_levels, int B, int r, int dtype,
                         int flow_dtype) {
  if (!c || n_levels < 0 || (n_levels > 0 && !lv)) return fail(PWC_ERR_BAD_ARG, "bad arguments");
  if (!dtype_ok(dtype) || !dtype_ok(flow_dtype)) return fail(PWC_ERR_BAD_DTYPE, "unknown dtype");
  const size_t es = elt_size(dtype), fs = elt_size(flow_dtype);
  const int64_t D = (int64_t)(2 * r + 1) * (2 * r + 1);
  // The batch is cut into chunks (batch elements are independent) so that the upload of chunk k+1, the
  // kernels of chunk k and the download of chunk k-1 overlap: PCIe is full duplex and the copies, not the kernels,
  // bound this path (822 MB per step at the headline workload against < 1 ms of kernels).
  // Chunk sizes ramp up and down (quarter, half, full ... full, half, quarter of B/8): the first upload and the last
  // download are the only copies that nothing overlaps, so they are kept short.
  int cstart[40];
  int nch = 0;
  {
    const char* env = getenv("PWC_HOST_CHUNK");  // experiments: images per full chunk
    const int full = env && atoi(env) > 0 ? atoi(env) : (B >= 8 ? B / 8 : 1);
    const int ramp[2] = {full >= 4 ? full / 4 : 1, full >= 2 ? full / 2 : 1};
    int b0 = 0;
    auto push = [&](int n) {
      if (n > B - b0) n = B - b0;
      if (n > 0 && nch < 39) { cstart[nch++] = b0; b0 += n; }
    };
    if (B >= 16) { push(ramp[0]); push(ramp[1]); }
    const int tail = B >= 16 ? ramp[0] + ramp[1] : 0;
    while (B - b0 > tail && nch < 36) push(full <= B - b0 - tail ? full : B - b0 - tail);
    if (B >= 16) { push(ramp[1]); push(ramp[0]); }
    if (b0 < B) push(B - b0);
    cstart[nch] = B;
  }
  while ((int)c->ev_in.size() < n_levels * nch) {
    cudaEvent_t e1, e2;
    PWC_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
    PWC_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    c->ev_in.push_back(e1);
    c->ev_run.push_back(e2);
  }
  for (int i = 0; i < n_levels; ++i) {
    const pwc_level_t& L = lv[i];
    int rc = check_shape(B, L.H, L.W, L.C, r, L.flow != nullptr);
    if (rc) return rc;
    // the download copies dense rows: a strided host `out` would have its other channels overwritten
    if (L.out_pixel_stride != 0 && L.out_pixel_stride != D)
      return fail(PWC_ERR_BAD_ARG, "pwc_pyramid_fwd_host: out_pixel_stride must be 0 (dense host output), got %lld",
                  (long long)L.out_pixel_stride);
    const size_t npix = (size_t)B * L.H * L.W;
    const int64_t ops = D;
    if ((rc = ctx_reserve(c, 4 * i + 0, npix * L.C * es))) return rc;
    if ((rc = ctx_reserve(c, 4 * i + 1, npix * L.C * es))) return rc;
    if (L.flow && (rc = ctx_reserve(c, 4 * i + 2, npix * 2 * fs))) return rc;
    if ((rc = ctx_reserve(c, 4 * i + 3, npix * ops * es))) return rc;
  }
  const bool timing = getenv("PWC_HOST_TIMING") != nullptr;
  const int tmode = timing ? atoi(getenv("PWC_HOST_TIMING")) : 0;  // 2: no downloads, 3: uploads only (results invalid)  // diagnostics: per-stream elapsed times on stderr
  cudaEvent_t tev[6] = {};
  if (timing) {
    for (auto& e : tev) cudaEventCreate(&e);
    cudaEventRecord(tev[0], c->s_in); cudaEventRecord(tev[1], c->s_run); cudaEventRecord(tev[2], c->s_out);
  }
  for (int ch = 0; ch < nch; ++ch) {
    const int b0 = cstart[ch], bn = cstart[ch + 1] - b0;
    if (bn <= 0) continue;
    for (int i = 0; i < n_levels; ++i) {
      const pwc_level_t& L = lv[i];
      // small levels are not worth cutting up (a copy costs ~10 us whatever its size): whole batch with chunk 0
      const bool whole = (size_t)B * L.H * L.W * L.C * es <= ((size_t)32 << 20);
      if (whole && ch != 0) continue;
      const int lb0 = whole ? 0 : b0, lbn = whole ? B : bn;
      const size_t pix0 = (size_t)lb0 * L.H * L.W, npix = (size_t)lbn * L.H * L.W;
      if (npix == 0) continue;
      const int64_t ops = D;
      const size_t o_feat = pix0 * L.C * es, nb_feat = npix * L.C * es;
      const size_t o_flow = pix0 * 2 * fs, nb_flow = npix * 2 * fs;
      const size_t o_out = pix0 * ops * es, nb_out = npix * ops * es;
      char* d_c1 = (char*)c->bufs[4 * i + 0].p + o_feat;
      char* d_c2 = (char*)c->bufs[4 * i + 1].p + o_feat;
      char* d_fl = L.flow ? (char*)c->bufs[4 * i + 2].p + o_flow : nullptr;
      char* d_out = (char*)c->bufs[4 * i + 3].p + o_out;
      const int e = ch * n_levels + i;
      PWC_CUDA(cudaMemcpyAsync(d_c1, (const char*)L.c1 + o_feat, nb_feat, cudaMemcpyHostToDevice, c->s_in));
      PWC_CUDA(cudaMemcpyAsync(d_c2, (const char*)L.c2 + o_feat, nb_feat, cudaMemcpyHostToDevice, c->s_in));
      if (L.flow)
        PWC_CUDA(cudaMemcpyAsync(d_fl, (const char*)L.flow + o_flow, nb_flow, cudaMemcpyHostToDevice, c->s_in));
      PWC_CUDA(cudaEventRecord(c->ev_in[e], c->s_in));
      PWC_CUDA(cudaStreamWaitEvent(c->s_run, c->ev_in[e], 0));
      int rc = tmode >= 3 ? 0 : fwd_entry(d_c1, d_c2, d_fl, L.flow_scale, d_out, lbn, L.H, L.W, L.C, r, dtype, flow_dtype, ops, c->s_run);
      if (rc) return rc;
      PWC_CUDA(cudaEventRecord(c->ev_run[e], c->s_run));
      PWC_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_run[e], 0));
      if (tmode < 2) PWC_CUDA(cudaMemcpyAsync((char*)L.out + o_out, d_out, nb_out, cudaMemcpyDeviceToHost, c->s_out));
    }
  }
  if (timing) { cudaEventRecord(tev[3], c->s_in); cudaEventRecord(tev[4], c->s_run); cudaEventRecord(tev[5], c->s_out); }
  PWC_CUDA(cudaStreamSynchronize(c->s_out));
  PWC_CUDA(cudaStreamSynchronize(c->s_in));
  if (timing) {
    float a = 0, b = 0, d = 0, e = 0;
    cudaEventElapsedTime(&a, tev[0], tev[3]); cudaEventElapsedTime(&b, tev[0], tev[4]);
    cudaEventElapsedTime(&d, tev[0], tev[5]); cudaEventElapsedTime(&e, tev[2], tev[5]);
    fprintf(stderr, "[pwc host] upload stream done at %.2f ms, kernels at %.2f ms, download at %.2f ms (download stream busy span %.2f ms), %d chunks\n", a, b, d, e, nch);
    for (auto& ev : tev) cudaEventDestroy(ev);
  }
  return PWC_OK;
}

// ---- backward ------------------------------------------------------------------------------
size_t pwc_bwd_workspace_bytes(int B, int H, int W, int C, int r, int dtype, int has_flow) {
  if (B < 0 || H < 0 || W < 0 || C < 0) return 0;
  // covers both the fused backward and the standalone warp backward (which needs only the fp32 accumulator)
  size_t need = bwd_ws_layout(B, H, W, C, dtype, has_flow != 0).total;
  if (r == pwc::bwdf::R && C % (elt_size(dtype) == 4 ? 4 : 8) == 0 && C > 0 && B > 0 && H > 0 && W > 0) {
    // flow dtype unknown here: assume a 16-bit flow (the larger layout)
    const size_t f = pwc::bwdf::make_plan(B, H, W, C, elt_size(dtype), 2, has_flow != 0, 148).total;
    if (f > need) need = f;
  }
  return need;
}

int pwc_cost_volume_bwd(const void* g, const void* out, const void* c1, const void* warp, void* dc1, void* dwarp, int B,
                        int H, int W, int C, int r, int dtype, int64_t gps, int64_t ops, void* cuda_stream) {
  if (!dtype_ok(dtype)) return fail(PWC_ERR_BAD_DTYPE, "unknown dtype");
  int rc = check_shape(B, H, W, C, r, false);
  if (rc) return rc;
  if ((int64_t)B * H * W == 0) return PWC_OK;
  if (!g || !out || !c1 || !warp || !dc1 || !dwarp) return fail(PWC_ERR_BAD_ARG, "null tensor pointer");
  const int64_t D = (int64_t)(2 * r + 1) * (2 * r + 1);
  if (gps == 0) gps = D;
  if (ops == 0) ops = D;
  DevInfo dev;
  rc = device_info(&dev);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (dtype == PWC_F32) return cv_bwd_typed<float>(g, out, c1, warp, dc1, dwarp, B, H, W, C, r, gps, ops, st, dev);
  if (dtype == PWC_F16) return cv_bwd_typed<__half>(g, out, c1, warp, dc1, dwarp, B, H, W, C, r, gps, ops, st, dev);
  return cv_bwd_typed<__nv_bfloat16>(g, out, c1, warp, dc1, dwarp, B, H, W, C, r, gps, ops, st, dev);
}

int pwc_dense_image_warp_bwd(const void* gout, const void* image, const void* flow, void* dimage, void* dflow, int B,
                             int H, int W, int C, int img_dtype, int flow_dtype, void* workspace,
                             size_t workspace_bytes, void* cuda_stream) {
  if (!dtype_ok(img_dtype) || !dtype_ok(flow_dtype)) return fail(PWC_ERR_BAD_DTYPE, "unknown dtype");
  int rc = check_shape(B, H, W, C, 0, true);
  if (rc) return rc;
  if ((int64_t)B * H * W == 0) return PWC_OK;
  if (!gout || !image || !flow || !dimage || !dflow) return fail(PWC_ERR_BAD_ARG, "null tensor pointer");
  DevInfo dev;
  rc = device_info(&dev);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
#define PWC_WB(T, TF) return warp_bwd_typed<T, TF>(gout, image, flow, dimage, dflow, B, H, W, C, workspace, workspace_bytes, st, dev)
  if (img_dtype == PWC_F32 && flow_dtype == PWC_F32) PWC_WB(float, float);
  if (img_dtype == PWC_F16 && flow_dtype == PWC_F16) PWC_WB(__half, __half);
  if (img_dtype == PWC_F16 && flow_dtype == PWC_F32) PWC_WB(__half, float);
  if (img_dtype == PWC_BF16 && flow_dtype == PWC_BF16) PWC_WB(__nv_bfloat16, __nv_bfloat16);
  if (img_dtype == PWC_BF16 && flow_dtype == PWC_F32) PWC_WB(__nv_bfloat16, float);
#undef PWC_WB
  return fail(PWC_ERR_BAD_DTYPE, "unsupported (image dtype %d, flow dtype %d) combination", img_dtype, flow_dtype);
}

int pwc_warp_cost_volume_bwd(const void* g, const void* out, const void* c1, const void* c2, const void* flow,
                             float flow_scale, void* dc1, void* dc2, void* dflow, int B, int H, int W, int C, int r,
                             int dtype, int flow_dtype, int64_t gps, int64_t ops, void* workspace,
                             size_t workspace_bytes, void* cuda_stream) {
  if (!dtype_ok(dtype) || (flow && !dtype_ok(flow_dtype))) return fail(PWC_ERR_BAD_DTYPE, "unknown dtype");
  int rc = check_shape(B, H, W, C, r, flow != nullptr);
  if (rc) return rc;
  if ((int64_t)B * H * W == 0) return PWC_OK;
  if (!g || !out || !c1 || !c2 || !dc1 || !dc2 || (flow && !dflow)) return fail(PWC_ERR_BAD_ARG, "null tensor pointer");
  const int64_t D = (int64_t)(2 * r + 1) * (2 * r + 1);
  if (gps == 0) gps = D;
  if (ops == 0) ops = D;
  DevInfo dev;
  rc = device_info(&dev);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (!flow) flow_dtype = dtype;
#define PWC_FB(T, TF)                                                                                              \
  return fused_bwd_typed<T, TF>(g, out, c1, c2, flow, flow_scale, dc1, dc2, dflow, B, H, W, C, r, dtype, gps, ops, \
                                workspace, workspace_bytes, st, dev)
  if (dtype == PWC_F32 && flow_dtype == PWC_F32) PWC_FB(float, float);
  if (dtype == PWC_F16 && flow_dtype == PWC_F16) PWC_FB(__half, __half);
  if (dtype == PWC_F16 && flow_dtype == PWC_F32) PWC_FB(__half, float);
  if (dtype == PWC_BF16 && flow_dtype == PWC_BF16) PWC_FB(__nv_bfloat16, __nv_bfloat16);
  if (dtype == PWC_BF16 && flow_dtype == PWC_F32) PWC_FB(__nv_bfloat16, float);
#undef PWC_FB
  return fail(PWC_ERR_BAD_DTYPE, "unsupported (feature dtype %d, flow dtype %d) combination", dtype, flow_dtype);
}

}  // extern "C"
