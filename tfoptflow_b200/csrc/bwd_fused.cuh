// Fused backward of warp + cost volume for search_range 4, two launches per level.
//
//   g'[p,d]    = g[p,d] * (out[p,d] > 0 ? 1 : 0.1) / C                       (leaky-relu mask, channel mean)
//   dc1[p,c]   = sum_d g'[p,d]   * warpZ[p + d, c]                           K1
//   dwarp[q,c] = sum_e g'[q+e,-e] * c1[q + e, c]        (e = -d)             K2
//   dc2        = bilinear scatter of dwarp, dflow = -(dwarp . d warp/d alpha) * clamp mask * flow_scale   K2
//
// Both sums have the same shape: out_row[px][c] = sum_{ey,ex} A[ey][ex][px] * Bwin[row + ey - 4][px + ex - 4][c], a
// banded correlation of 81 scalar planes A against a 9-row window of feature rows.  One consumer routine serves both:
//   * a CTA owns (image, 32-pixel strip, band of rows, 32-channel slice) and marches down the band;
//   * B rows live in a shared-memory ring (K1: rows of c2 warped on the fly by gather warps, never in HBM; K2: rows of c1);
//   * A rows (81 x 32 fp32 per strip row) live in a second ring.  K1 builds them from g and out and also sends them, as
//     they are, to a workspace tensor [B * H (TARGET row)][dyi][dxi][x]: 9 TMA tensor stores per row (one plane block
//     per target row, whole 128-byte lines).  K2 fetches the A row of a dwarp row from there with TMA tensor loads that
//     start ex - 4 pixels to the side, rounded down to a multiple of 4 (tensor copies want 16-byte aligned starts), and
//     shifts the remaining 0..3 pixels in shared memory;
//   * 4 consumer warps (one per scheduler) each own a PAIR of output rows: lanes 0-15 row 2p, lanes 16-31 row 2p+1,
//     8 pixels x 8 channels per lane (32 packed f32x2 accumulators).  The two halves walk the window skewed by one
//     row, so both read the SAME window row in every step: the B operand (LDS.128, the expensive one) is shared by
//     the two halves, the A operand is a scalar-broadcast FFMA2 source shared by the 4 channel groups of a pixel group.
//   * K2 hands finished dwarp rows to 8 scatter warps, one per staging slot (red.global.add.v4.f32 into an fp32 dc2,
//     warp-shuffle dflow).
// Synchronisation: mbarriers only (full per ring slot, done per task); no __syncthreads after the role split.
// A parity wait also passes when the barrier is one whole phase BEHIND the one waited for (the parity then names the
// preceding phase), so every barrier here has either a single producer and a single consumer that alternate strictly
// (staging slots), or two alternating barrier sets per data slot (rings: index i % 2N, parity (i / 2N) & 1), or a
// dependency chain that keeps waiters less than one lap ahead (task-done barriers, see wait_window_free).
#pragma once
#include <cuda.h>   // CUtensorMap (types only: the encode entry point is fetched through the runtime)
#include "common.cuh"

#ifndef PWC_BWD_GATHER_DEPTH
#define PWC_BWD_GATHER_DEPTH 2   /* K1 gather: steps (4 corner loads each) in flight ahead of the one being blended */
#endif
#ifndef PWC_BWD_SCATTER_NB
#define PWC_BWD_SCATTER_NB 2   /* K2 scatter: 2 = corners a quarter row at a time (8 LDG.128 in flight), 4 = half a row */
#endif
#ifndef PWC_BWD_LANEMAP
#define PWC_BWD_LANEMAP 1   /* 0: (cg, pg, h) lane order of the first version */
#endif

namespace pwc {
namespace bwdf {

typedef unsigned long long u64;

constexpr int R = 4, ND = 9, D = 81;
constexpr int TW = 32;                      // strip width in pixels
constexpr int KS = 32;                      // channels per CTA
constexpr int WPX = TW + 2 * R;             // window pixels per row (halo of 4 each side)
constexpr int WROW = WPX * 8 + WPX / 8;     // float4 per window row: 8 per pixel + 1 pad per 8 pixels
constexpr int AROW = D * TW / 4;            // float4 per A row
constexpr int ASLICE = ND * TW / 4;         // float4 per (row, ey) slice
constexpr int SROW = TW * 8 + TW / 8;       // float4 per staged dwarp row (same padding as the window)
constexpr int NG = 10;                      // A ring slots: 4 tasks x 2 rows in flight, + 2 ahead
constexpr int NT = 16;                      // task-done barriers
constexpr int NDS = 8;                      // staged dwarp rows: two (one per half) per consumer warp, drained by its own scatter warp
constexpr int NCW = 4;                      // consumer warps; the other 12 are split per kernel (template NAW_)
constexpr int NTHREADS = 512;
constexpr int REGS_CONSUMER = 200, REGS_OTHER = 104;   // 128 * 200 + 384 * 104 = 512 * 128
constexpr int AREM = 2 * 512;               // K2: the two remainder boxes of an A row (2 x 432 bytes, 128-byte aligned)

// Shared-memory layout per kernel.  K1: 18 window rows (4 tasks in flight span 16, + 2 ahead).  K2: its A slots also
// hold the remainder boxes, paid for with a 16-row window ring (the two rows a task frees are needed by the same warp's
// next task only in its last two steps)
template <int MODE>
struct Lay {
  static constexpr int NW = MODE == 0 ? 18 : 16;
  static constexpr size_t ASTRIDE = (size_t)AROW * 16 + (MODE == 0 ? 0 : AREM);   // bytes per A slot
  static constexpr size_t OFF_A = 0;                 // TMA tensor source / destination: 128-byte aligned slots
  static constexpr size_t OFF_Z = OFF_A + (size_t)NG * ASTRIDE;
  static constexpr size_t OFF_W = OFF_Z + (size_t)ASLICE * 16;
  static constexpr size_t OFF_S = OFF_W + (size_t)NW * WROW * 16;
  static constexpr size_t OFF_BAR = OFF_S + (size_t)NDS * SROW * 16;
  static constexpr int NBAR = 2 * NW + 2 * NG + NT + 2 * NDS + (MODE == 0 ? 0 : 2 * NG);
  static constexpr size_t SMEM_BYTES = OFF_BAR + (size_t)NBAR * 8;
  static_assert(ASTRIDE % 128 == 0 && OFF_Z % 128 == 0 && (ND * TW * 4) % 128 == 0, "TMA tensor copies want 128-byte aligned shared memory");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};
constexpr size_t SMEM_BYTES = Lay<0>::SMEM_BYTES > Lay<1>::SMEM_BYTES ? Lay<0>::SMEM_BYTES : Lay<1>::SMEM_BYTES;

struct Args {
  const void* g;        // K1: upstream gradient [B,H,W,81]
  const void* out;      // K1: forward result    [B,H,W,81]
  const void* feat;     // K1: c2 (warped on the fly)   K2: c1
  const void* c2;       // K2: corners for dflow
  const void* flow;     // NULL = no warp
  void* dst;            // K1: dc1   K2 without flow: dc2
  float* acc;           // K2 with flow: fp32 dc2 accumulator (zeroed)
  void* dflow;          // K2 with flow, single slice: flow gradient
  float* dflow_acc;     // K2 with flow, several slices: fp32 accumulator (zeroed)
  float* ws;            // g' workspace, the tensor [B * H (target row)][dyi][dxi][wpitch] behind the kernels' tensor map
  int B, H, W, C;
  int nsx, nslices, NPAIR;   // strips per row, 32-channel slices, row pairs per image
  int wpitch;                // row pitch of a workspace plane in floats (W rounded up to 4)
  long long Q;               // B * nsx * nslices * NPAIR (strip, row pair) units, split evenly over the CTAs
  float flow_scale, inv_c;
  int pf;    // L2 prefetch distance of the producers, in turns of their row loop (option bwd_prefetch; 0 = off)
  int dbg;   // timing experiments only (option bwd_debug): 1 = no FMA loop, 2 = A rows without loads, 4 = window rows
             // without loads, 8 = no scatter (K2), 16 = no workspace stores (K1); results are wrong when non-zero
};

// ---- mbarrier / bulk-copy helpers ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA tensor copies (SASS: UTMALDG / UTMASTG), 4-d, coordinates innermost first
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// L2 prefetch of a contiguous span / of one 128-byte line per pixel (the producers are latency-bound: one warp per
// row with a few dozen loads in flight; pulling the rows of the next turn (option bwd_prefetch) into L2 cuts each exposure from DRAM
// latency to L2 latency)
__device__ __forceinline__ void prefetch_span(const void* p, int nbytes, int lane) {
  const char* c = reinterpret_cast<const char*>(p);
  for (int o = lane * 128; o < nbytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + o));
}
__device__ __forceinline__ void prefetch_pixels(const void* p, int npx, size_t pixel_bytes, int lane) {
  const char* c = reinterpret_cast<const char*>(p);
  for (int i = lane; i < npx; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + (size_t)i * pixel_bytes));
}

// scalar x packed pair (ptxas folds the {a, a} pair into FFMA2's scalar-broadcast operand form)
__device__ __forceinline__ void ffma2s(u64& d, float a, u64 b) {
  u64 aa;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(aa), "l"(b));
}
__device__ __forceinline__ float2 unpack2(u64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// float index of (pixel px, plane d) inside a K1 A row: [d][px], the 8 float4 groups of a plane XOR-ed with key & 7, where
// key = slot + d = bits 7..9 of the element's shared-memory address (the A ring starts at a 1024-byte boundary, a slot
// is 81 x 128 bytes, a plane 128 bytes): the TMA unit's 128-byte swizzle, which the tensor stores undo on the way out.
// The builder's stores (lanes = 4 pixels x 8 consecutive planes) hit 32 distinct banks with it.
__device__ __forceinline__ int a_index(int px, int d, int key) { return d * TW + ((((px >> 2) ^ (key & 7)) << 2) | (px & 3)); }

template <typename T>
__device__ __forceinline__ float ld1(const T* p) { return ElemTraits<T>::to_f(__ldg(p)); }
template <>
__device__ __forceinline__ float ld1<__half>(const __half* p) {
  return __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(p))));
}
template <>
__device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float((unsigned)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
}

// 8 consecutive channels -> global, in the tensor dtype
// (n = channels of the 8 that exist: a partial last slice has a multiple of 4 (fp32) / 8 (16-bit) of them)
__device__ __forceinline__ void st8(float* p, const float (&v)[8], int n) {
  if (n >= 4) st4(p, make_float4(v[0], v[1], v[2], v[3]));
  if (n >= 8) st4(p + 4, make_float4(v[4], v[5], v[6], v[7]));
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&v)[8], int n) {  // 16-bit: one 16-byte store
  if (n < 8) return;
  uint4 u;
  uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const T lo = ElemTraits<T>::from_f(v[2 * k]), hi = ElemTraits<T>::from_f(v[2 * k + 1]);
    w[k] = (uint32_t)(*reinterpret_cast<const unsigned short*>(&lo)) |
           ((uint32_t)(*reinterpret_cast<const unsigned short*>(&hi)) << 16);
  }
  *reinterpret_cast<uint4*>(p) = u;
}

struct Unit {
  int b, x0, y0, rows, P, U, cs, sx;   // rows = valid rows of the band, P = row pairs, U = window rows
  int cv;                              // valid channels of slice cs: 32, less in the last slice when C % 32 != 0
};

// The kernel is persistent: CTA i owns the units [Q i / G, Q (i + 1) / G) of the linearised (image, strip, slice, row pair)
// space, cut into per-strip bands, and every warp role walks the same list of bands.  Ring positions and barrier
// phases are counted GLOBALLY over the CTA's life (tasks: Pb + p, window rows: Ub + u, A rows: 2 (Pb + p) + h), so the
// producers run ahead into the next band while the consumers finish the current one - no drain / refill between bands.
struct Walk {
  long long q, q1;
  int Pb, Ub;          // tasks / window rows of the bands before the current one
  int pP, pPb, pUb;    // the previous band: its P and bases (pP == 0: none)
  __device__ __forceinline__ void init(const Args& a) {
    q = a.Q * blockIdx.x / gridDim.x;
    q1 = a.Q * (blockIdx.x + 1) / gridDim.x;
    Pb = Ub = pP = pPb = pUb = 0;
  }
  __device__ __forceinline__ bool next(const Args& a, Unit& un) const {
    if (q >= q1) return false;
    const long long strip = q / a.NPAIR;
    const int p0 = (int)(q - strip * a.NPAIR);
    const long long rem = q1 - q;
    const int n = (long long)(a.NPAIR - p0) < rem ? a.NPAIR - p0 : (int)rem;
    int id = (int)strip;
    un.cs = id % a.nslices; id /= a.nslices;
    un.cv = min(KS, a.C - un.cs * KS);
    un.sx = id % a.nsx;
    un.b = id / a.nsx;
    un.x0 = un.sx * TW;
    un.y0 = 2 * p0;
    un.rows = min(2 * n, a.H - un.y0);
    un.P = n;
    un.U = 2 * n + 2 * R;
    return true;
  }
  __device__ __forceinline__ void advance(const Unit& un) {
    pP = un.P; pPb = Pb; pUb = Ub;
    Pb += un.P; Ub += un.U; q += un.P;
  }
};

// wait until every task that reads the window row with GLOBAL index `ug` has finished.  Row u of a band is read by its
// tasks p with 2p <= u <= 2p + 9, the last one being min(u / 2, P - 1); tasks are spread round-robin (global index) over
// the NCW consumer warps and complete in order per warp, so the last NCW task indices cover "all tasks <= plast".
__device__ __forceinline__ void wait_window_free(uint64_t* done, int ug, const Walk& w, int P) {
  int plast;
  if (ug >= w.Ub) plast = w.Pb + min((ug - w.Ub) >> 1, P - 1);
  else if (w.pP > 0 && ug >= w.pUb) plast = w.pPb + min((ug - w.pUb) >> 1, w.pP - 1);
  else plast = w.pPb - 1;  // two or more bands back: everything before the previous band
#pragma unroll
  for (int k = 0; k < NCW; ++k) {
    const int p = plast - k;
    if (p >= 0) mbar_wait(&done[p % NT], (p / NT) & 1);
  }
}
// first local index >= 0 whose global index (base + i) is congruent to `me` modulo `n`
__device__ __forceinline__ int first_mine(int base, int me, int n) { return ((me - base) % n + n) % n; }

// ---- B producers -------------------------------------------------------------------------------
// window row u <-> image row y0 - 4 + u, pixels x0 - 4 .. x0 + 35, channels [32 cs, 32 cs + 32); zero outside the image.
// 8 lanes per pixel (one float4 each), 4 pixels per step, 10 steps per row.  The flow of all 10 steps is loaded and
// turned into corner indices / weights first (one latency exposure), then the corners are gathered in a software pipeline.
template <bool WARP, int NPROD, int NW, typename T, typename TF>
__device__ __forceinline__ void produce_window(const Args& a, const Unit& un, const Walk& w, float4* sW,
                                               uint64_t* full_w, uint64_t* done, int pi, int lane) {
  const T* __restrict__ feat = reinterpret_cast<const T*>(a.feat);
  const TF* __restrict__ flow = reinterpret_cast<const TF*>(a.flow);
  const int psub = lane >> 3, f4 = lane & 7;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int NIT = WPX / 4;           // 10
  const T* base = feat + (size_t)un.cs * KS + 4 * f4;
  const bool cok = 4 * f4 < un.cv;
  auto prefetch_row = [&](int u) {
    const int y = un.y0 - R + u;
    if (u >= un.U || y < 0 || y >= a.H) return;
    const int xlo = max(un.x0 - 2 * R, 0), xhi = min(un.x0 + TW + 2 * R, a.W);
    const size_t rp = ((size_t)un.b * a.H + y) * a.W + xlo;
    prefetch_pixels(feat + rp * a.C + (size_t)un.cs * KS, xhi - xlo, (size_t)a.C * sizeof(T), lane);
    if (WARP) prefetch_span(flow + rp * 2, (xhi - xlo) * 2 * (int)sizeof(TF), lane);
  };
  const int u0 = first_mine(w.Ub, pi, NPROD);  // rows are dealt round-robin by GLOBAL index
  for (int k = 0; k < a.pf; ++k) prefetch_row(u0 + k * NPROD);
  for (int u = u0; u < un.U; u += NPROD) {
    const int ug = w.Ub + u;
    if (a.pf > 0) prefetch_row(u + a.pf * NPROD);
    if (ug >= NW) wait_window_free(done, ug - NW, w, un.P);
    float4* Wd = sW + (size_t)(ug % NW) * WROW;
    const int y = un.y0 - R + u;
    if (y < 0 || y >= a.H || (a.dbg & 4)) {
      for (int i = lane; i < WROW; i += 32) Wd[i] = zero;
    } else {
      int pix[NIT];
      float wx[NIT], wy[NIT];
      unsigned inmask = 0;
      {
        TF f0[NIT], f1[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int x = un.x0 - R + it * 4 + psub;
          const bool in = x >= 0 && x < a.W;
          inmask |= in ? (1u << it) : 0u;
          const int xc = min(max(x, 0), a.W - 1);
          pix[it] = (un.b * a.H + y) * a.W + xc;
          if (WARP) {
            f0[it] = flow[2 * (size_t)pix[it]];
            f1[it] = flow[2 * (size_t)pix[it] + 1];
          }
        }
        if (WARP) {
#pragma unroll
          for (int it = 0; it < NIT; ++it) {
            const int x = un.x0 - R + it * 4 + psub;
            const int xc = min(max(x, 0), a.W - 1);
            const SampleCoord s = sample_coord_v<T, TF>(f0[it], f1[it], a.flow_scale, un.b, y, xc, a.H, a.W);
            pix[it] = s.pix;
            wx[it] = s.ax;
            wy[it] = s.ay;
          }
        }
      }
      // software pipeline over the 10 steps of the row: the corner loads of step i + GD are issued before step i is
      // blended, so a lane keeps 4 GD .. 4 (GD + 1) LDG.128 in flight for the whole row instead of one latency exposure
      // per batch (fully unrolled; the ring of corner registers is indexed at compile time).  Plain copies (K2) issue
      // the whole row at once.
      constexpr int GD = WARP ? PWC_BWD_GATHER_DEPTH : NIT - 1;
      float4 tl[GD + 1], tr[GD + 1], bl[GD + 1], br[GD + 1];
#pragma unroll
      for (int i = 0; i < NIT + GD; ++i) {
        if (i < NIT) {
          const int k = i % (GD + 1);
          const T* q = base + (size_t)pix[i] * a.C;
          tl[k] = tr[k] = bl[k] = br[k] = zero;
          if (cok) {   // channels past C in a partial last slice read as zeros
            tl[k] = ld4(q);
            if (WARP) {
              tr[k] = ld4(q + a.C);
              bl[k] = ld4(q + (size_t)a.W * a.C);
              br[k] = ld4(q + (size_t)a.W * a.C + a.C);
            }
          }
        }
        if (i >= GD) {
          const int it = i - GD, k = it % (GD + 1), px = it * 4 + psub;
          float4 v = tl[k];
          if (WARP) {
            if (ElemTraits<T>::kIsHalf) {
              v = round4<T>(bilerp4(tl[k], tr[k], bl[k], br[k], wx[it], wy[it]));  // the reference's warp is a 16-bit tensor
            } else {
              const float w11 = wx[it] * wy[it], w01 = wx[it] - w11, w10 = wy[it] - w11, w00 = (1.0f - wx[it]) - w10;
              v.x = fmaf(w11, br[k].x, fmaf(w10, bl[k].x, fmaf(w01, tr[k].x, w00 * tl[k].x)));
              v.y = fmaf(w11, br[k].y, fmaf(w10, bl[k].y, fmaf(w01, tr[k].y, w00 * tl[k].y)));
              v.z = fmaf(w11, br[k].z, fmaf(w10, bl[k].z, fmaf(w01, tr[k].z, w00 * tl[k].z)));
              v.w = fmaf(w11, br[k].w, fmaf(w10, bl[k].w, fmaf(w01, tr[k].w, w00 * tl[k].w)));
            }
          }
          Wd[px * 8 + (px >> 3) + f4] = ((inmask >> it) & 1u) ? v : zero;
        }
      }
    }
    mbar_arrive(&full_w[ug % (2 * NW)]);
  }
}

// ---- A producers -------------------------------------------------------------------------------
// K1: A[d][px] = g * mask / C for strip row j, then the 9 plane blocks go to the workspace by TMA tensor stores.
//
// Workspace: a 4-d fp32 tensor [B * H (TARGET row)][dyi 9][dxi 9][x] over the whole image width (row pitch W rounded up
// to 4).  Plane block dyi of SOURCE row y lands in target row y + dyi - 4, at the source's own x: one tensor store per
// block (box 32 x 9 planes, whole aligned 128-byte lines - no partial sectors), clipped at the image edge by the
// tensor map.  The K2 consumer wants, for dwarp row y and strip x0, A[ey][ex][px] = g' of the source pixel
// (y + ey - 4, x0 + px + ex - 4) for displacement (8 - ey, 8 - ex): that is the box (x0 + ex - 4, dxi = 8 - ex,
// dyi = 0 .. 8, row y) of this tensor, so K2 fetches an A row with tensor loads that start (a multiple of 4 pixels) to
// the side - the TMA unit does the coarse shift and zero-fills what lies outside the image, see produce_a_k2.  Blocks
// whose source row is outside the image are written as zeros (from sZ) by the builder of the row that shares their index.
// The A slot is swizzled the way the tensor map's 128-byte mode expects (16-byte chunk ^= bits 7..9 of the address).
template <int NAW, typename T>
__device__ __forceinline__ void produce_a_k1(const Args& a, const Unit& un, const Walk& w, float* sA, uint64_t* full_a,
                                             const float4* sZ, uint64_t* done, uint64_t* stored, int ai, int lane,
                                             int& pending_jg, const CUtensorMap* tm) {
  const T* __restrict__ g = reinterpret_cast<const T*>(a.g);
  const T* __restrict__ out = reinterpret_cast<const T*>(a.out);
  const int psub = lane >> 3, dq = lane & 7;
  const int npx = min(TW, a.W - un.x0);
  auto prefetch_row = [&](int j) {
    if (j >= un.rows) return;
    const size_t rp = (((size_t)un.b * a.H + un.y0 + j) * a.W + un.x0) * D;
    prefetch_span(g + rp, npx * D * (int)sizeof(T), lane);
    prefetch_span(out + rp, npx * D * (int)sizeof(T), lane);
  };
  // `pending_jg` (caller's, lives across bands): the row whose tensor stores are not yet known to have read their A slot.
  // stored[slot] tells the builder of the slot's next occupant that they have (one barrier per slot is enough: the
  // task-done wait that precedes it orders the phases)
  const int j0 = first_mine(2 * w.Pb, ai, NAW);
  for (int k = 0; k < a.pf; ++k) prefetch_row(j0 + k * NAW);
  for (int j = j0; j < 2 * un.P; j += NAW) {
    const int jg = 2 * w.Pb + j;
    if (a.pf > 0) prefetch_row(j + a.pf * NAW);
    if (jg >= NG) {
      const int p = (jg - NG) >> 1;
      mbar_wait(&done[p % NT], (p / NT) & 1);                       // the consumers are done with the slot
      mbar_wait(&stored[(jg - NG) % NG], ((jg - NG) / NG) & 1);     // and so are the tensor stores of its last occupant
    }
    const int slot = jg % NG;
    float* A = sA + (size_t)slot * (AROW * 4);
    const int y = un.y0 + j;
    if (j >= un.rows || (a.dbg & 2)) {
      for (int i = lane; i < AROW; i += 32) reinterpret_cast<float4*>(A)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const size_t rowpix = ((size_t)un.b * a.H + y) * a.W + un.x0;
#pragma unroll 1
      for (int grp = 0; grp < TW / 4; grp += 2) {   // 2 groups x 11 plane octets x (g, out): 44 loads in flight per lane
        float gv[2][11], ov[2][11];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int px = (grp + q) * 4 + psub;
          const bool pin = un.x0 + px < a.W;
          const T* gp = g + (rowpix + px) * D;
          const T* op = out + (rowpix + px) * D;
#pragma unroll
          for (int o = 0; o < 11; ++o) {
            const int d = o * 8 + dq;
            const bool ok = pin && d < D;
            gv[q][o] = ok ? ld1<T>(gp + d) : 0.f;
            ov[q][o] = ok ? ld1<T>(op + d) : 1.f;
          }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int px = (grp + q) * 4 + psub;
#pragma unroll
          for (int o = 0; o < 11; ++o) {
            const int d = o * 8 + dq;
            if (d < D) A[a_index(px, d, slot + d)] = (gv[q][o] * (ov[q][o] > 0.0f ? 1.0f : 0.1f)) * a.inv_c;
          }
        }
      }
    }
    if (j < un.rows) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my stores, before TMA reads them
    __syncwarp();
    if (j < un.rows) {
      // my previous row's stores have had a whole row's time to read their slot: confirm it and hand the slot on
      if (pending_jg >= 0) {
        if (lane < ND) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&stored[pending_jg % NG]);
      }
      if (lane < ND) {
        const int yq = y + lane - R;  // plane block dyi = lane lands in target row y + dyi - 4
        if (yq >= 0 && yq < a.H && !(a.dbg & 16)) tma_store_4d(tm, A + lane * ND * TW, un.x0, 0, lane, un.b * a.H + yq);
        // target row y itself: block dyi = lane would come from source row y + 4 - dyi
        const int ys = y + R - lane;
        if ((ys < 0 || ys >= a.H) && !(a.dbg & 16)) tma_store_4d(tm, sZ, un.x0, 0, lane, un.b * a.H + y);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      pending_jg = jg;
    } else if (lane == 0) {
      mbar_arrive(&stored[slot]);   // nothing is stored from a row past the end of the band
    }
    mbar_arrive(&full_a[jg % (2 * NG)]);
  }
}

// K2: A[ex][dyi][px] = g' of the source pixel (y + 4 - dyi, x0 + px + ex - 4) for displacement (dyi, dxi = 8 - ex) = plane
// (dyi, 8 - ex) of workspace row y at x = x0 + px + s, s = ex - 4.  A tensor load must start at a multiple of 16 bytes
// (a box at any other x faults: scripts/ubench/tma_tensor_probe.cu), so the TMA unit does the coarse part of the shift,
// c = 4 floor(s / 4) in {-4, 0, 4}: one box (32 px x 9 dyi) per ex at x0 + c, plus two small boxes (4 px x 3 planes x
// 9 dyi) with the 4 pixels that follow the main boxes of ex = 1..3 and ex = 5..7.  The remaining r = s - c = ex & 3 pixels
// are shifted in place, in shared memory, by the warp that issued the copies (54 rows of 32 floats: lane = pixel, all
// rows read, then all rows written).  No global latency is exposed: the copies of a warp's next row are in flight
// while it shifts the current one.  landed[] counts the copies' bytes, full_a is the hand-over to the consumers.
constexpr uint32_t A_TX_BYTES = D * TW * 4 + 2 * (4 * 3 * ND * 4);
__device__ __forceinline__ void k2_issue_row(const Args& a, const Unit& un, int j, int jg, unsigned char* sA,
                                             uint64_t* landed, uint64_t* done, int lane, const CUtensorMap* tm_main,
                                             const CUtensorMap* tm_rem, int pf) {
  if (jg >= NG) {
    const int p = (jg - NG) >> 1;
    mbar_wait(&done[p % NT], (p / NT) & 1);
  }
  if (j >= un.rows || (a.dbg & 2)) return;   // a zero row: filled at hand-over time
  if (lane == 0) {
    unsigned char* A = sA + (size_t)(jg % NG) * Lay<1>::ASTRIDE;
    uint64_t* bar = &landed[jg % (2 * NG)];
    const int row = un.b * a.H + un.y0 + j;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the consumers' reads of the slot come first
    mbar_arrive_expect_tx(bar, A_TX_BYTES);
#pragma unroll 1
    for (int ex = 0; ex < ND; ++ex)
      tma_load_4d(A + ex * (ND * TW * 4), tm_main, bar, un.x0 + ((ex - R) & ~3), 2 * R - ex, 0, row);
    // ex = 1..3 (c = -4): pixels x0 + 28 .. 31 of planes dxi = 7, 6, 5;  ex = 5..7 (c = 0): x0 + 32 .. 35 of dxi = 3, 2, 1
    tma_load_4d(A + AROW * 16, tm_rem, bar, un.x0 + TW - R, 5, 0, row);
    tma_load_4d(A + AROW * 16 + AREM / 2, tm_rem, bar, un.x0 + TW, 1, 0, row);
  } else if (pf > 0 && lane <= ND && j + pf < un.rows) {
    // pull a row further down the band into L2: lanes 1..9 take one main box each (the remainder boxes share its lines)
    const int ex = lane - 1;
    tma_prefetch_4d(tm_main, un.x0 + ((ex - R) & ~3), 2 * R - ex, 0, un.b * a.H + un.y0 + j + pf);
  }
}
__device__ __forceinline__ void k2_finish_row(const Args& a, const Unit& un, int j, int jg, unsigned char* sA,
                                              uint64_t* landed, uint64_t* full_a, int lane) {
  float* A = reinterpret_cast<float*>(sA + (size_t)(jg % NG) * Lay<1>::ASTRIDE);
  if (j >= un.rows || (a.dbg & 2)) {
    for (int i = lane; i < AROW; i += 32) reinterpret_cast<float4*>(A)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    mbar_wait(&landed[jg % (2 * NG)], (jg / (2 * NG)) & 1);
    const float* rem = A + AROW * 4;
    float v[6 * ND];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int ex = k < 3 ? k + 1 : k + 2, r = ex & 3;          // 1, 2, 3, 5, 6, 7
      const float* rbox = rem + (k < 3 ? 0 : AREM / 8);           // [dyi][3 planes][4 px]; plane dxi = 8 - ex is box plane 3 - r
#pragma unroll
      for (int dyi = 0; dyi < ND; ++dyi)
        v[k * ND + dyi] = lane + r < TW ? A[(ex * ND + dyi) * TW + lane + r] : rbox[(dyi * 3 + (3 - r)) * 4 + lane + r - TW];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int ex = k < 3 ? k + 1 : k + 2;
#pragma unroll
      for (int dyi = 0; dyi < ND; ++dyi) A[(ex * ND + dyi) * TW + lane] = v[k * ND + dyi];
    }
  }
  mbar_arrive(&full_a[jg % (2 * NG)]);
}
template <int NAW>
__device__ __forceinline__ void produce_a_k2(const Args& a, const Unit& un, const Walk& w, unsigned char* sA,
                                             uint64_t* full_a, uint64_t* landed, uint64_t* done, int ai, int lane,
                                             const CUtensorMap* tm_main, const CUtensorMap* tm_rem) {
  // my rows of this band: every NAW-th, by global index; the copies of row j + NAW are issued before row j is finished
  const int j0 = first_mine(2 * w.Pb, ai, NAW);
  if (j0 < 2 * un.P) k2_issue_row(a, un, j0, 2 * w.Pb + j0, sA, landed, done, lane, tm_main, tm_rem, a.pf * NAW);
  for (int j = j0; j < 2 * un.P; j += NAW) {
    if (j + NAW < 2 * un.P) k2_issue_row(a, un, j + NAW, 2 * w.Pb + j + NAW, sA, landed, done, lane, tm_main, tm_rem, a.pf * NAW);
    k2_finish_row(a, un, j, 2 * w.Pb + j, sA, landed, full_a, lane);
  }
}

// ---- consumer ----------------------------------------------------------------------------------
// MODE 0: K1, store dc1.  MODE 1: K2; without a flow the rows are dc2 and are stored, with a flow they are staged for
// the scatter warps.
template <int MODE, bool HAS_FLOW, typename T>
__device__ __forceinline__ void consume(const Args& a, const Unit& un, const Walk& w, const float4* sW, const float4* sA,
                                        const float4* sZ, float4* sS, uint64_t* full_w, uint64_t* full_a,
                                        uint64_t* done, uint64_t* full_d, uint64_t* empty_d, int warp, int lane,
                                        int& nstaged) {
  // Lane order (h, cg, pg) from the lowest bit: every aligned group of 4 lanes reads 2 distinct B words (cg) and 2
  // distinct A words (h), the pattern the crossbar serves in 2 cycles per LDS.128 (scripts/ubench/lds_pattern.cu);
  // with (cg, pg, h) the four channel groups of a lane quad made every B load a 4-cycle one.
#if PWC_BWD_LANEMAP == 0
  const int h = lane >> 4, s = lane & 15, pg = s >> 2, cg = s & 3;
#else
  const int h = lane & 1, cg = (lane >> 1) & 3, pg = lane >> 3;
#endif
  const int g0 = 2 * pg, g1 = 2 * pg + 1;
  for (int p = first_mine(w.Pb, warp, NCW); p < un.P; p += NCW) {
    const int pg_ = w.Pb + p, jg = 2 * pg_;   // global task / A row index
    mbar_wait(&full_a[jg % (2 * NG)], (jg / (2 * NG)) & 1);
    mbar_wait(&full_a[(jg + 1) % (2 * NG)], ((jg + 1) / (2 * NG)) & 1);
    constexpr int NW = Lay<MODE>::NW;
    const int aslot = (jg + h) % NG;
    const float4* Abase = sA + (size_t)aslot * (Lay<MODE>::ASTRIDE / 16);
    u64 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][q] = 0ull;

#pragma unroll 1
    for (int t = 0; t < ND + 1; ++t) {
      const int u = w.Ub + 2 * p + t;   // global window row
      mbar_wait(&full_w[u % (2 * NW)], (u / (2 * NW)) & 1);
      if (a.dbg & 1) continue;
      const ulonglong2* Wr = reinterpret_cast<const ulonglong2*>(sW + (size_t)(u % NW) * WROW + pg * 65 + 2 * cg);
      const int ey = t - h;  // the two halves use the same window row for different displacements
      const bool eyv = ey >= 0 && ey < ND;
      // K1 rows: [ey][ex][px] with the TMA swizzle (key = slot + plane, see a_index).  K2 rows: [ex][dyi = 8 - ey][px] as the
      // tensor loads deliver them.  A half without a displacement in this step reads the zero slice.
      const float4* Ar = !eyv ? sZ : (MODE == 0 ? Abase + ey * ASLICE : Abase + (2 * R - ey) * (TW / 4));
      const int akey = eyv ? aslot + ey * ND : 0;   // zero slice: any permutation will do
      const int astep = eyv ? ND * TW / 4 : 0;
      auto aidx = [&](int ex, int grp) { return MODE == 0 ? ex * (TW / 4) + (grp ^ ((akey + ex) & 7)) : ex * astep + grp; };
      float4 a0[ND], a1[ND];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < ND) a0[j] = Ar[aidx(j, g0)];
        if (j >= 4 && j - 4 < ND) a1[j - 4] = Ar[aidx(j - 4, g1)];
        const ulonglong2 b0 = Wr[j * 8 + (j >> 3)], b1 = Wr[j * 8 + (j >> 3) + 1];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int ex = j - i;
          if (ex >= 0 && ex < ND) {
            const float av = i < 4 ? comp(a0[ex], i) : comp(a1[ex], i - 4);
            ffma2s(acc[i][0], av, b0.x);
            ffma2s(acc[i][1], av, b0.y);
            ffma2s(acc[i][2], av, b1.x);
            ffma2s(acc[i][3], av, b1.y);
          }
        }
      }
    }

    // epilogue: row j of the band, pixels x0 + 8 pg .. + 7, channels 32 cs + 8 cg .. + 7
    const int j = 2 * p + h;
    const bool rowv = j < un.rows;
    if (MODE == 1 && HAS_FLOW) {
      // this warp's own two staging slots (one per half), drained by scatter warp `warp`: one producer, one consumer
      // (`nstaged` = rows this half has staged so far, over all bands: rows past the end of a band are not staged)
      const int slot = 2 * warp + h;
      if (rowv) {
        if (nstaged > 0) mbar_wait(&empty_d[slot], (nstaged - 1) & 1);
        ++nstaged;
        float4* S = sS + (size_t)slot * SROW + pg * 65 + 2 * cg;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 v0 = unpack2(acc[i][0]), v1 = unpack2(acc[i][1]), v2 = unpack2(acc[i][2]), v3 = unpack2(acc[i][3]);
          S[i * 8] = make_float4(v0.x, v0.y, v1.x, v1.y);
          S[i * 8 + 1] = make_float4(v2.x, v2.y, v3.x, v3.y);
        }
        mbar_arrive(&full_d[slot]);
      }
    } else if (rowv) {
      T* dst = reinterpret_cast<T*>(a.dst) +
               (((size_t)un.b * a.H + un.y0 + j) * a.W + un.x0 + 8 * pg) * a.C + un.cs * KS + 8 * cg;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (un.x0 + 8 * pg + i < a.W) {
          const float2 v0 = unpack2(acc[i][0]), v1 = unpack2(acc[i][1]), v2 = unpack2(acc[i][2]), v3 = unpack2(acc[i][3]);
          const float v[8] = {v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, v3.x, v3.y};
          st8(dst + (size_t)i * a.C, v, un.cv - 8 * cg);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&done[pg_ % NT]);
  }
}

// ---- K2 scatter warps ----------------------------------------------------------------------------
// dense_image_warp backward for one staged dwarp row: 8 lanes per pixel (4 channels each), 4 pixels per step.
// Scatter warp (cw, h) drains staging slot 2 cw + h: row 2p + h of every task p of consumer warp cw (one producer, one
// consumer per slot; `nuse` = rows taken from the slot so far, over all bands).
template <typename T, typename TF>
__device__ __forceinline__ void scatter_rows(const Args& a, const Unit& un, const Walk& w, const float4* sS,
                                             uint64_t* full_d, uint64_t* empty_d, int cw, int h, int lane, int& nuse) {
  const T* __restrict__ c2 = reinterpret_cast<const T*>(a.c2);
  const TF* __restrict__ flow = reinterpret_cast<const TF*>(a.flow);
  const int psub = lane >> 3, f4 = lane & 7;
  const int slot = 2 * cw + h;
  auto load_flow = [&](int p, TF& f0, TF& f1) {  // flow of this lane's pixel of row 2p + h
    const int y = un.y0 + 2 * p + h;
    const int x = min(un.x0 + lane, a.W - 1);
    const size_t pp = ((size_t)un.b * a.H + y) * a.W + x;
    f0 = flow[2 * pp];
    f1 = flow[2 * pp + 1];
  };
  auto row_ok = [&](int p) { return p < un.P && 2 * p + h < un.rows; };
  int p = first_mine(w.Pb, cw, NCW);
  TF nf0 = TF(0.f), nf1 = TF(0.f);
  if (row_ok(p)) load_flow(p, nf0, nf1);
  for (; row_ok(p); p += NCW) {
    const int j = 2 * p + h;
    const bool have_next = row_ok(p + NCW);
    {  // corners of the row this warp scatters after the next one, assuming moderate flows
      const int yn = un.y0 + j + 2 * a.pf * NCW;
      if (a.pf > 0 && yn < a.H) {
        const int xlo = max(un.x0 - 2 * R, 0), xhi = min(un.x0 + TW + 2 * R, a.W);
        prefetch_pixels(c2 + (((size_t)un.b * a.H + yn) * a.W + xlo) * a.C + (size_t)un.cs * KS, xhi - xlo,
                        (size_t)a.C * sizeof(T), lane);
      }
    }
    // Latency plan: the flow of the whole row is read once (lane = pixel) and turned into coordinates, then the c2
    // corners are loaded a quarter row at a time (8 LDG.128 in flight per lane; more does not fit the 104 registers of
    // this role) - the first quarter BEFORE waiting for the staged dwarp row, so that it overlaps the consumer's work.
    const int y = un.y0 + j;
    const SampleCoord sc = sample_coord_v<T, TF>(nf0, nf1, a.flow_scale, un.b, y, min(un.x0 + lane, a.W - 1), a.H, a.W);
    const float4* S = sS + (size_t)slot * SROW;
    constexpr int NB = PWC_BWD_SCATTER_NB;  // steps (4 pixels each) per batch of corner loads
#pragma unroll 1
    for (int half = 0; half < TW / 4 / NB; ++half) {
      float4 tl[NB], tr[NB], bl[NB], br[NB];
      int pixs[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const int px = (half * NB + k) * 4 + psub;
        pixs[k] = __shfl_sync(0xffffffffu, sc.pix, px);
        const size_t i_tl = (size_t)pixs[k] * a.C + un.cs * KS + 4 * f4;
        // always a valid address: the coordinates of columns past the image edge were taken at the clamped column
        if ((a.dbg & 8) || 4 * f4 >= un.cv) { tl[k] = tr[k] = bl[k] = br[k] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
        tl[k] = ld4(c2 + i_tl);
        tr[k] = ld4(c2 + i_tl + a.C);
        bl[k] = ld4(c2 + i_tl + (size_t)a.W * a.C);
        br[k] = ld4(c2 + i_tl + (size_t)a.W * a.C + a.C);
      }
      if (half == 0) {
        if (have_next) load_flow(p + NCW, nf0, nf1);  // consumed at the top of the next row
        mbar_wait(&full_d[slot], nuse & 1);
        ++nuse;
      }
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const int px = (half * NB + k) * 4 + psub, x = un.x0 + px;
        const bool valid = x < a.W && !(a.dbg & 8);
        const bool cok = 4 * f4 < un.cv;   // lanes of channels past C add nothing
        const float ax = __shfl_sync(0xffffffffu, sc.ax, px), ay = __shfl_sync(0xffffffffu, sc.ay, px);
        const float mx = __shfl_sync(0xffffffffu, sc.mx, px), my = __shfl_sync(0xffffffffu, sc.my, px);
        float dax = 0.f, day = 0.f;
        if (valid && cok) {
          const size_t i_tl = (size_t)pixs[k] * a.C + un.cs * KS + 4 * f4, i_tr = i_tl + a.C;
          const size_t i_bl = i_tl + (size_t)a.W * a.C, i_br = i_bl + a.C;
          const float4 gw = S[px * 8 + (px >> 3) + f4];
          const float w_tl = (1.f - ay) * (1.f - ax), w_tr = (1.f - ay) * ax, w_bl = ay * (1.f - ax), w_br = ay * ax;
          atomicAdd(reinterpret_cast<float4*>(a.acc + i_tl), make_float4(gw.x * w_tl, gw.y * w_tl, gw.z * w_tl, gw.w * w_tl));
          atomicAdd(reinterpret_cast<float4*>(a.acc + i_tr), make_float4(gw.x * w_tr, gw.y * w_tr, gw.z * w_tr, gw.w * w_tr));
          atomicAdd(reinterpret_cast<float4*>(a.acc + i_bl), make_float4(gw.x * w_bl, gw.y * w_bl, gw.z * w_bl, gw.w * w_bl));
          atomicAdd(reinterpret_cast<float4*>(a.acc + i_br), make_float4(gw.x * w_br, gw.y * w_br, gw.z * w_br, gw.w * w_br));
          const float gv[4] = {gw.x, gw.y, gw.z, gw.w};
          const float vtl[4] = {tl[k].x, tl[k].y, tl[k].z, tl[k].w}, vtr[4] = {tr[k].x, tr[k].y, tr[k].z, tr[k].w};
          const float vbl[4] = {bl[k].x, bl[k].y, bl[k].z, bl[k].w}, vbr[4] = {br[k].x, br[k].y, br[k].z, br[k].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float top = ax * (vtr[i] - vtl[i]) + vtl[i];
            const float bot = ax * (vbr[i] - vbl[i]) + vbl[i];
            dax = fmaf(gv[i], (1.f - ay) * (vtr[i] - vtl[i]) + ay * (vbr[i] - vbl[i]), dax);
            day = fmaf(gv[i], bot - top, day);
          }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          dax += __shfl_xor_sync(0xffffffffu, dax, o);
          day += __shfl_xor_sync(0xffffffffu, day, o);
        }
        if (valid && f4 == 0) {
          const size_t pidx = ((size_t)un.b * a.H + y) * a.W + x;
          const float d0 = -day * my * a.flow_scale, d1 = -dax * mx * a.flow_scale;
          if (a.dflow_acc != nullptr) {
            atomicAdd(a.dflow_acc + 2 * pidx, d0);
            atomicAdd(a.dflow_acc + 2 * pidx + 1, d1);
          } else {
            TF* df = reinterpret_cast<TF*>(a.dflow);
            df[2 * pidx] = ElemTraits<TF>::from_f(d0);
            df[2 * pidx + 1] = ElemTraits<TF>::from_f(d1);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_d[slot]);
  }
}

// ---- the kernel -----------------------------------------------------------------------------------
// Warp roles after the 4 consumers.  K1: NAW_ A builders, the other 12 - NAW_ gather the warped rows.  K2: NAW_ warps
// fetch the A rows, 8 scatter warps (with a flow: one per staging slot), the rest copy c1 rows into the window ring.
template <int MODE, bool HAS_FLOW, int NAW_, typename T, typename TF>
__global__ void __launch_bounds__(NTHREADS, 1)
bwd_fused_kernel(const Args a, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap2) {
  using L = Lay<MODE>;
  constexpr int NW = L::NW;
  extern __shared__ __align__(1024) unsigned char smem[];
  float4* sW = reinterpret_cast<float4*>(smem + L::OFF_W);
  float4* sA = reinterpret_cast<float4*>(smem + L::OFF_A);
  float4* sZ = reinterpret_cast<float4*>(smem + L::OFF_Z);
  float4* sS = reinterpret_cast<float4*>(smem + L::OFF_S);
  uint64_t* full_w = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* full_a = full_w + 2 * NW;
  uint64_t* done = full_a + 2 * NG;
  uint64_t* full_d = done + NT;
  uint64_t* empty_d = full_d + NDS;
  uint64_t* landed = empty_d + NDS;   // K2 only

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2 * NW; ++i) mbar_init(&full_w[i], 32);
    for (int i = 0; i < 2 * NG; ++i) mbar_init(&full_a[i], 32);
    for (int i = 0; i < NT; ++i) mbar_init(&done[i], 1);
    for (int i = 0; i < NDS; ++i) { mbar_init(&full_d[i], MODE == 0 ? 1 : 16); mbar_init(&empty_d[i], 1); }   // K1: `stored`
    if (MODE == 1)
      for (int i = 0; i < 2 * NG; ++i) mbar_init(&landed[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async (TMA) proxy
  }
  for (int i = tid; i < ASLICE; i += NTHREADS) sZ[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // sZ is a TMA source in K1
  __syncthreads();

  // every role walks the CTA's bands in the same order with its own copy of the walk state
  Walk w;
  w.init(a);
  Unit un;
  if (warp < NCW) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_CONSUMER));
    int nstaged = 0;
    while (w.next(a, un)) {
      consume<MODE, HAS_FLOW, T>(a, un, w, sW, sA, sZ, sS, full_w, full_a, done, full_d, empty_d, warp, lane, nstaged);
      w.advance(un);
    }
    return;
  }
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_OTHER));
  constexpr int NHELP = NTHREADS / 32 - NCW;   // 12
  const int hi = warp - NCW;
  if constexpr (MODE == 0) {
    constexpr int NGW = NHELP - NAW_;
    static_assert(NGW >= 1, "K1 warp split");
    static_assert(NG <= 2 * NDS, "K1's stored barriers live where K2 has its staging barriers");
    uint64_t* stored = full_d;
    int pending_jg = -1;
    while (w.next(a, un)) {
      if (hi < NAW_)
        produce_a_k1<NAW_, T>(a, un, w, reinterpret_cast<float*>(sA), full_a, sZ, done, stored, hi, lane, pending_jg, &tmap);
      else produce_window<HAS_FLOW, NGW, NW, T, TF>(a, un, w, sW, full_w, done, hi - NAW_, lane);
      w.advance(un);
    }
    // my last tensor stores must have read their shared-memory source before the CTA's memory goes away
    if (hi < NAW_ && pending_jg >= 0) {
      if (lane < ND) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&stored[pending_jg % NG]);
    }
  } else {
    constexpr int NSW = HAS_FLOW ? 2 * NCW : 0;   // one scatter warp per staging slot
    constexpr int NWW = NHELP - NAW_ - NSW;
    static_assert(NWW >= 1, "window producers");
    int nuse = 0;
    while (w.next(a, un)) {
      if (hi < NAW_)
        produce_a_k2<NAW_>(a, un, w, smem + L::OFF_A, full_a, landed, done, hi, lane, &tmap, &tmap2);
      else if (hi < NAW_ + NWW) produce_window<false, NWW, NW, T, TF>(a, un, w, sW, full_w, done, hi - NAW_, lane);
      else scatter_rows<T, TF>(a, un, w, sS, full_d, empty_d, (hi - NAW_ - NWW) >> 1, (hi - NAW_ - NWW) & 1, lane, nuse);
      w.advance(un);
    }
  }
}

// Workspace: transposed g' (fp32), then - with a flow - an fp32 dc2 accumulator for 16-bit tensors and an fp32 dflow
// accumulator when the channels are processed in several slices.
struct Plan {
  int nsx, nslices, npair, wpitch;   // wpitch = row pitch of the workspace planes in floats (W rounded up to 4)
  long long Q;   // (image, strip, slice, row pair) units
  int grid;      // persistent CTAs (one per SM: the kernel uses all of an SM's shared memory)
  size_t ws_off, acc_off, dflow_off, total;
  bool acc_in_ws, dflow_in_ws;
};
inline size_t up256(size_t n) { return (n + 255) & ~size_t(255); }

inline Plan make_plan(int B, int H, int W, int C, size_t elt, size_t flow_elt, bool has_flow, int sms) {
  Plan p{};
  p.nsx = (W + TW - 1) / TW;
  p.nslices = (C + KS - 1) / KS;
  p.npair = (H + 1) / 2;
  p.Q = (long long)B * p.nsx * p.nslices * p.npair;
  // a band shorter than ~4 row pairs pays more for its 8 halo rows than for its own: no more CTAs than that allows
  long long g = p.Q / 4 > 0 ? p.Q / 4 : 1;
  if (g > sms) g = sms;
  p.grid = (int)g;
  size_t off = 0;
  p.ws_off = off;
  p.wpitch = (W + 3) & ~3;
  off += up256((size_t)B * H * D * p.wpitch * 4);
  p.acc_in_ws = has_flow && elt != 4;
  p.acc_off = off;
  if (p.acc_in_ws) off += up256((size_t)B * H * W * C * 4);
  p.dflow_in_ws = has_flow && p.nslices > 1 && flow_elt != 4;
  p.dflow_off = off;
  if (p.dflow_in_ws) off += up256((size_t)B * H * W * 2 * 4);
  p.total = off;
  return p;
}

inline bool covered(int C, int r, int64_t gps, int64_t ops, const void* c1, const void* c2, const void* dc1,
                    const void* dc2, size_t elt) {
  auto al = [&](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  // channels in slices of 32; the last slice may be partial as long as 16-byte vectors never straddle the end of a pixel
  return r == R && C % (elt == 4 ? 4 : 8) == 0 && gps == D && ops == D && al(c1) && al(c2) && al(dc1) && al(dc2);
}

}  // namespace bwdf
}  // namespace pwc
