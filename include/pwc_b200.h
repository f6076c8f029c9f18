/* pwc_b200.h - C ABI of the B200-native warp + cost-volume library (libpwc_b200.so).
 *
 * Drop-in boundary for PWC-Net's per-pyramid-level hot path in philferriere/tfoptflow.
 * The reference has NO FFI for this path: it is two plain Python functions built from
 * stock TensorFlow ops.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference checkout):
 *
 *   pwc_cost_volume_*        <- cost_volume(c1, warp, search_range, name)   tfoptflow/core_costvol.py:20-40
 *   pwc_dense_image_warp_*   <- dense_image_warp(image, flow, name)         tfoptflow/core_warp.py:153-202
 *                               (_interpolate_bilinear                      tfoptflow/core_warp.py:42-150)
 *   pwc_warp_cost_volume_*   <- ModelPWCNet.warp + ModelPWCNet.corr fused   tfoptflow/model_pwcnet.py:1178,1277
 *                               incl. the `up_flow * scaler` multiply       tfoptflow/model_pwcnet.py:1560-1561
 *   pwc_pyramid_fwd*         <- the coarse-to-fine loop's hot-path calls    tfoptflow/model_pwcnet.py:1550-1567
 *   *_bwd                    <- what TF autodiff emits for the above        (selected at model_pwcnet.py:547-550)
 *
 * Conventions
 *   - Tensors are NHWC, dense, channels contiguous; `*_pixel_stride` (elements) lets the cost
 *     volume be written into / read from a channel slice of a wider buffer (the concat of
 *     model_pwcnet.py:1424).  Pass 0 for "dense" (= (2r+1)^2).
 *   - The caller owns every buffer (outputs and workspace included); the library keeps no
 *     reference after return and never frees caller memory.
 *   - All device entry points are asynchronous on `cuda_stream` (a cudaStream_t, NULL = legacy
 *     default stream) and run on the calling thread's current device.  No hidden syncs.
 *   - Return 0 on success, a negative pwc_status otherwise; pwc_last_error() gives a
 *     thread-local message.  No C++ exception crosses this boundary.
 *   - sm_100 (B200) only: other devices are refused with PWC_ERR_ARCH.  There is no CPU path.
 */
#ifndef PWC_B200_H_
#define PWC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PWC_ABI_VERSION 1

typedef enum { PWC_F32 = 0, PWC_F16 = 1, PWC_BF16 = 2 } pwc_dtype;

typedef enum {
  PWC_OK = 0,
  PWC_ERR_BAD_ARG = -1,   /* null pointer, negative size, unknown option */
  PWC_ERR_BAD_SHAPE = -2, /* H or W < 2 with a flow (core_warp.py:101), B*H*W >= 2^31, r < 0 ... */
  PWC_ERR_BAD_DTYPE = -3,
  PWC_ERR_ALIGN = -4,     /* pointer not aligned to its element size */
  PWC_ERR_ARCH = -5,      /* not an sm_100 device */
  PWC_ERR_CUDA = -6,      /* a CUDA runtime call failed; message has the cudaError string */
  PWC_ERR_WORKSPACE = -7  /* workspace too small */
} pwc_status;

int pwc_abi_version(void);
const char* pwc_last_error(void);

/* Number of kernels this library has launched on this process so far (all threads). */
uint64_t pwc_launch_count(void);

/* Tuning / debug options (process-wide).  pwc_set_option returns PWC_ERR_BAD_ARG for an unknown key or an out-of-range
 * value; pwc_get_option reads back every key pwc_set_option accepts.
 *   "fwd_impl"        0 = auto (default: streaming kernel for r == 4 and H >= 28 - the band GEMM instead for 16-bit
 *                     features with C in {32, 64, 96, 128} at any height - tiled kernel for short levels and r == 8 (band GEMM for 16-bit features,
 *                     C in {32, 64, 96, 128}, >= 16 K pixels), generic otherwise),
 *                     1 = generic scalar kernel, 2 = tiled kernel, 3 = streaming kernel,
 *                     4 = tensor-core band GEMM for 16-bit features (r == 4 or 8, C in {32, 64, 96, 128}; other
 *                     shapes fall back)
 *   "bwd_impl"        0 = auto (default: fused K1/K2 kernels for r == 4, C % 4 == 0 (16-bit: C % 8), dense g / out; else v1),
 *                     1 = unfused v1 kernels, 2 = fused kernels or PWC_ERR_BAD_ARG when the shape is not covered
 *   "overlap_levels"  1 (default) = pwc_pyramid_fwd forks the short levels onto side streams, 2 = the tall levels after
 *                     the first get streams of their own as well, 0 = one stream
 *   "side_sms"        SMs the tall levels' persistent kernels leave free for those short levels (default 16, 0..64)
 *   "side_reverse"    1 (default) = fork the short levels largest first
 *   "tile_small"      1 (default) = 8x16 tiles for levels with fewer 8x32 tiles than SMs
 *   "bwd_tiled"       1 (default) = register-tiled v1 backward correlation kernels, 0 = one pixel per thread
 *   "stream_ctas"     CTA count of the streaming kernel (0 = one per SM, default)
 *   "stream_prefetch" L2 prefetch distance of the streaming kernel in slabs (default 4, 0 = off)
 *   "stream_debug"    DIAGNOSTICS ONLY, default 0: bits 1 / 2 / 4 skip the correlation / the gathers / the stores
 *                     of the streaming kernel (RESULTS ARE WRONG; bench.py refuses to report with them set), bit 8
 *                     forces the LDG path for c1, bit 16 records the event trace read by pwc_debug_read_trace
 *   "mma_tile"        16-bit band GEMM for r == 4: 1 (default) = resident-halo kernel (16- or 8-row tiles, every warp stages and
 *                     writes its own units; also covers C == 96 and 128), 0 = 8x32 tiles with a tile-wide staging buffer
 *   "bwd_prefetch"    L2 prefetch distance of the fused backward's producer warps, in turns of their row loops
 *                     (default 1: two turns ahead re-reads 40 % of the
 *                     prefetched lines from DRAM, profiles/r02_notes.md; 0 = off, 0..8)
 *   "bwd_debug"       DIAGNOSTICS ONLY, default 0: bits 1 / 2 / 4 / 8 / 16 skip the FMA loop / the g, out (K1) or
 *                     workspace (K2) loads / the window-row loads / the scatter / the workspace stores of the fused
 *                     backward (RESULTS ARE WRONG; bench.py refuses to report with it set) */
int pwc_set_option(const char* key, int value);
int pwc_get_option(const char* key, int* value);

/* ---- forward ------------------------------------------------------------------------ */

/* cost_volume(c1, warp, search_range): out[b,i,j,(dy+r)(2r+1)+(dx+r)] =
 * lrelu_0.1( mean_c c1[b,i,j,c] * warpZ[b,i+dy,j+dx,c] ), warpZ zero outside the image. */
int pwc_cost_volume_fwd(const void* c1, const void* warp, void* out,
                        int B, int H, int W, int C, int r, int dtype,
                        int64_t out_pixel_stride, void* cuda_stream);

/* dense_image_warp(image, flow): bilinear sample of image at (i - flow[...,0], j - flow[...,1])
 * with the reference's clamping rules.  Output has image's dtype. */
int pwc_dense_image_warp_fwd(const void* image, const void* flow, void* out,
                             int B, int H, int W, int C, int img_dtype, int flow_dtype,
                             void* cuda_stream);

/* cost_volume(c1, dense_image_warp(c2, flow * flow_scale), r) without materialising the warp.
 * flow == NULL means "no warp" (top pyramid level). */
int pwc_warp_cost_volume_fwd(const void* c1, const void* c2, const void* flow, float flow_scale,
                             void* out, int B, int H, int W, int C, int r,
                             int dtype, int flow_dtype, int64_t out_pixel_stride, void* cuda_stream);

typedef struct {
  const void* c1;
  const void* c2;
  const void* flow; /* NULL = no warp */
  void* out;
  int H, W, C;
  float flow_scale;
  int64_t out_pixel_stride; /* 0 = dense */
} pwc_level_t;

/* All hot-path calls of one pass of the pyramid loop, device buffers. */
int pwc_pyramid_fwd(const pwc_level_t* levels, int n_levels, int B, int r,
                    int dtype, int flow_dtype, void* cuda_stream);

/* Same with HOST buffers (pinned or pageable): copies inputs to device scratch owned by the
 * context, runs the kernels, copies the cost volumes back, and returns after the last copy
 * has completed.  The context caches its scratch buffers and streams between calls. */
typedef struct pwc_host_ctx pwc_host_ctx;
int pwc_host_ctx_create(pwc_host_ctx** ctx);
int pwc_host_ctx_destroy(pwc_host_ctx* ctx);
int pwc_pyramid_fwd_host(pwc_host_ctx* ctx, const pwc_level_t* host_levels, int n_levels,
                         int B, int r, int dtype, int flow_dtype);

/* Diagnostics: copies the streaming kernel's event trace (CTA 0; recorded when option stream_debug has bit 16
 * set) into dst_host; returns the number of bytes copied or a negative error.  Synchronises the device. */
int pwc_debug_read_trace(void* dst_host, int max_bytes);

/* ---- backward ----------------------------------------------------------------------- */

/* Scratch needed by the *_bwd entry points for this problem (bytes, 256-aligned). */
size_t pwc_bwd_workspace_bytes(int B, int H, int W, int C, int r, int dtype, int has_flow);

/* g, out: [B,H,W,*] with g_pixel_stride / out_pixel_stride (0 = dense); out is the forward
 * result (gives the leaky-relu mask).  dc1, dwarp: [B,H,W,C] in `dtype`. */
int pwc_cost_volume_bwd(const void* g, const void* out, const void* c1, const void* warp,
                        void* dc1, void* dwarp, int B, int H, int W, int C, int r, int dtype,
                        int64_t g_pixel_stride, int64_t out_pixel_stride, void* cuda_stream);

/* gout: [B,H,W,C] in img_dtype.  dimage [B,H,W,C] img_dtype, dflow [B,H,W,2] flow_dtype. */
int pwc_dense_image_warp_bwd(const void* gout, const void* image, const void* flow,
                             void* dimage, void* dflow, int B, int H, int W, int C,
                             int img_dtype, int flow_dtype,
                             void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Backward of pwc_warp_cost_volume_fwd.  flow == NULL: dflow is ignored, dc2 = dwarp. */
int pwc_warp_cost_volume_bwd(const void* g, const void* out, const void* c1, const void* c2,
                             const void* flow, float flow_scale,
                             void* dc1, void* dc2, void* dflow,
                             int B, int H, int W, int C, int r, int dtype, int flow_dtype,
                             int64_t g_pixel_stride, int64_t out_pixel_stride,
                             void* workspace, size_t workspace_bytes, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* PWC_B200_H_ */
