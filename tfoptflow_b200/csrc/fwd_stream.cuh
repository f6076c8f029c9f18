// Streaming fused warp + cost-volume forward (search_range 4, C in {16, 32, 64, 96, 128}).
//
// Persistent, warp-specialised kernel, one CTA per SM.  The image is cut into vertical strips of
// TW = 8 * NXG pixels (32 / 16 / 8 for C <= 32 / 64 / >= 96, so that all channels of the rows in flight fit in
// shared memory); every CTA owns a contiguous range of (strip, row-pair) work and marches down it.
//   * NPW producer warps evaluate the bilinear warp of c2 (8 lanes per pixel read whole 128-byte lines of every
//     corner, 4-tap weighted sum in packed FFMA2) and stream the *warped* rows ("slabs" = warped rows 2m-1, 2m) into
//     a shared-memory ring guarded by mbarriers.  The warped features never reach HBM; the 1/C of the channel mean is
//     folded into the bilinear weights.  The matching c1 row pairs arrive through TMA bulk copies
//     (cp.async.bulk -> mbarrier complete_tx; no registers, no issue slots) for fp32, by LDG/STS for 16-bit inputs.
//     Each producer warp owns every NPW-th slab, so NPW gathers are in flight.
//   * NCW compute warps pull tasks in order.  Every task is 32 independent lane-tasks: one c1 row against one
//     warped row, 8 pixels x 9 dx.  Accumulators are channel PAIRS (72 x f32x2) so that the inner loop is packed
//     FFMA2: one issue slot per two FMAs, which leaves issue bandwidth for the LDS.128 operand stream, the producers
//     and the epilogue.  Lanes are laid out (warped row, c1 row, x-group): every aligned group of 4 lanes sees at most
//     2 distinct c1 operands and 2 distinct warped operands, the pattern the LDS crossbar serves in 2 cycles per
//     LDS.128 instead of 4 (scripts/ubench/lds_pattern.cu).  A slab sees c1 rows 2m-4 .. 2m+3 (full tasks F); the two
//     remaining displacements (row 2m+4 with dy=-4, row 2m-5 with dy=+4) are done by one edge task E per 4 slabs,
//     with the channel range split over 4/NXG lanes and reduced by shuffles.
//   * Results go through a per-warp transposition buffer: each store instruction writes runs of 9 / 18 channels.
// Shared-memory layout is pixel-major ([pixel][channel], what TMA delivers) with 16-byte pads chosen so that the
// operands of one LDS land in distinct bank groups: row stride = 1, pair slot = 2, x-group = 4 (c1) / 2 (warped)
// mod 8 float4.
// Zero padding is applied after warping (core_costvol.py:27); the warp clamps (core_warp.py:101-104).
// The ring protocol (two alternating mbarrier sets per slot) is modelled in tests/stream_protocol_sim.py.
#pragma once
#include <type_traits>
#include "common.cuh"

namespace pwc {
namespace stream {

typedef unsigned long long u64;

constexpr int R = 4, ND = 9, D = 81;
constexpr int NCW = 8;   // compute warps (two warpgroups)
constexpr int NPW = 4;   // producer warps (one warpgroup)
constexpr int MAXSEG = 8;
constexpr int GS = 4;    // slabs per task group (one edge task per group)
constexpr size_t SMEM_LIMIT = 232448;
constexpr int GATHER_DEPTH = 3;   // gather items in flight ahead of the one being blended (3 / 4 / 5 measured equal)

constexpr int pad_to(int v, int r) { return v + ((r - v % 8) + 8) % 8; }  // smallest v' >= v with v' % 8 == r
constexpr int pad_to16(int v, int r) { return v + ((r - v % 16) + 16) % 16; }

// PX = pixels per lane-task (8, or 6 where the 72 packed accumulators of 8 pixels leave ptxas no room for a clean
// rolled channel loop: 54 pairs + operands fit 184 registers and give the producers 136)
// AB = bytes per (pixel, channel quad) in the c1 ring: 16 (fp32) or 8 (raw 16-bit elements delivered by TMA and widened
// in the correlation loop)
template <int KV_, int NXG_, int PX_ = 8, int AB_ = 16>
struct Geo {
  static constexpr int KV = KV_, NXG = NXG_, PX = PX_, AB = AB_;
  static constexpr int TW = PX * NXG;                 // strip width
  static constexpr int BU = TW + 2 * R;               // warped pixels per row (with halo)
  static constexpr int NPIX = 2 * BU;                 // pixels per slab
  // float4 units.  c1 pair slot: [sub(2)][xg][p(8)][kv] ; warped slab slot: [bsel(2)][u(BU)][kv], 2 pads per 8 pixels
  // c1 strides in units of AB bytes.  16-byte units: row 1, pair slot 2, x-group 4 (mod 8 = one 128-byte bank window).
  // 8-byte units (16-bit c1): everything even (TMA wants 16-byte destinations): row 2, slot 4, x-group 8 (mod 16).
  static constexpr int AXG = AB == 16 ? pad_to(PX * KV, 4) : pad_to16(PX * KV, 8);
  static constexpr int AROW = AB == 16 ? pad_to(NXG * AXG, 1) : pad_to16(NXG * AXG, 2);
  static constexpr int ASLOT = 2 * AROW;
  static constexpr int BXPAD = 2;
  static constexpr int BXG = PX * KV + BXPAD;
  static constexpr int BROW = pad_to(BU * KV + ((BU + PX - 1) / PX) * BXPAD, 1);
  static constexpr int BSLOT = 2 * BROW;
  static_assert(AB == 8 || (AXG % 8 == 4 && AROW % 8 == 1 && ASLOT % 8 == 2), "c1 bank-group strides");
  static_assert(AB == 16 || (AXG % 16 == 8 && AROW % 16 == 2 && ASLOT % 16 == 4), "16-bit c1 strides");
  static_assert(BROW % 8 == 1, "warped-row stride");
  static constexpr int NH = 16 / NXG;                 // c1 rows (hi values) per full task
  static constexpr int NF = 8 * GS / NH;              // full tasks per group: 8 / 4 / 2
  static constexpr int TPG = NF + 1;                  // + one edge task
  static constexpr int KP = 4 / NXG;                  // channel split of the edge task
  // channel-loop unroll: rolled (1) where the accumulators leave room, else long stretches so that the accumulator
  // shuffling ptxas does at every back-edge of a register-starved loop is amortised
  static constexpr int KU = PX <= 6 ? 1 : ((KV / KP) % 8 == 0 ? 8 : ((KV / KP) % 6 == 0 ? 6 : 4));
  // setmaxnreg moves registers inside the CTA allocation (384 x 168): 2 * REGS_COMPUTE + REGS_PRODUCER = 504
  static constexpr int NCWG = NCW;                    // compute warps (two per scheduler)
  static constexpr int NTHREADS = 32 * (NCWG + NPW);
  static constexpr int REGS_COMPUTE = PX <= 6 ? 192 : 200;
  static constexpr int REGS_PRODUCER = 504 - 2 * REGS_COMPUTE;
  static constexpr int EMPTYB_COUNT = (NXG == 4 ? 2 : 1) + 1;
  static constexpr int LPP = KV >= 8 ? 8 : KV;        // producer lanes per pixel
  static constexpr int KQ = KV / LPP;                 // kv chunks per lane
  // ring lengths (slab slots / c1 pair slots) per geometry: as long as shared memory allows (6 / 8 / 12 slab slots
  // measured 328 / 243 / 225 us on level 2).  NAP is even: the wrap-around keeps consecutive pairs 2 bank groups apart
  static constexpr int NSLAB = (KV == 32 && NXG == 2) ? 4 : (KV == 24 && NXG == 2) ? 6 : KV == 32 ? 6 : (KV == 16 ? (PX <= 6 ? 10 : 6) : (KV == 8 ? 12 : 7));
  static constexpr int NAP = (KV == 32 && NXG == 2) ? 8 : (KV == 24 && NXG == 2) ? 10 : KV == 32 ? 12 : (KV == 8 ? 14 : (KV == 16 && PX <= 6 ? 14 : 12));
  // transposition buffer per compute warp: two pixels, double buffered (staging the whole task costs 41 KB that buy
  // more as four extra slab slots)
  static constexpr int STAGE_WORDS = 640;
  static constexpr size_t OFF_B = 0;
  static constexpr size_t OFF_A = OFF_B + (size_t)NSLAB * BSLOT * 16;
  static constexpr size_t OFF_STAGE = (OFF_A + (size_t)NAP * ASLOT * AB + 15) / 16 * 16;
  static constexpr size_t OFF_SCRW = OFF_STAGE + (size_t)NCWG * STAGE_WORDS * 4;   // float4 weights per pixel
  static constexpr size_t OFF_SCRP = OFF_SCRW + (size_t)NPW * NPIX * 16;          // int2 {corner pixel, smem offset}
  static constexpr size_t OFF_BAR = OFF_SCRP + (size_t)NPW * NPIX * 8;
  static constexpr size_t OFF_SEG = OFF_BAR + (size_t)(4 * NSLAB + 4 * NAP) * 8;
  static constexpr size_t OFF_ACNT = OFF_SEG + MAXSEG * 32 + 64;                  // consumer arrivals per c1 pair slot
  static constexpr size_t SMEM_BYTES = (OFF_ACNT + NAP * 4 + 15) / 16 * 16;
  static_assert(SMEM_BYTES <= SMEM_LIMIT, "shared memory budget");
  static_assert(KV % LPP == 0 && KV % KP == 0 && NAP % 2 == 0, "channel split / ring");
  // An edge task waits for the 4 slabs of its group and for the c1 pairs l-5 .. l+3 of those slabs at once: 9 pairs
  // and 4 slabs must be resident together or the task can never start (found by scripts/stress_stream.py: 8 pair
  // slots deadlock).  One spare of each keeps the producers from stalling on the group boundary.
  static_assert(NAP >= 10 && NSLAB >= 5, "ring too short for an edge task");
};

struct Seg {
  int b, x0, p0, p1;   // batch, strip origin, pair range [p0, p1)
  int nslab;           // p1 - p0 + 5 slabs: m = p0-2 .. p1+2
  int task0, j0, a0;   // first global task / slab / pair index of this segment
};

struct Params {
  int B, H, W, C;
  int64_t ops;      // out pixel stride (elements)
  float flow_scale;
  int NPAIR;        // ceil(H / 2)
  int nstrip_x;     // ceil(W / TW)
  long long Q;      // total (strip, pair) units
  int pf_dist;      // L2 prefetch distance in slabs (0 = off)
  int dbg;          // timing experiments only: 1 = skip correlate, 2 = skip gathers, 4 = skip stores; 8 = c1 by LDG
};

// ---- event trace (diagnostics, stream_debug bit 16): CTA 0 records (clock << 16 | code << 12 | value) per warp ----
constexpr int TRACE_LEN = 512;
__device__ unsigned long long g_trace[16][TRACE_LEN];
#define PWC_TR(code, val)                                                                              \
  do {                                                                                                 \
    if ((P.dbg & 16) && blockIdx.x == 0 && lane == 0 && tr_n < TRACE_LEN)                              \
      g_trace[warp][tr_n++] = ((unsigned long long)clock64() << 16) | ((code) << 12) | ((val) & 0xfff); \
  } while (0)

// ---- mbarrier / TMA helpers ----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ int atom_add_acq_rel_cta(int* p, int v) {
  int old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // non-blocking
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- packed fp32 (sm_100 FFMA2 / FMUL2) ----------------------------------------------------
__device__ __forceinline__ void ffma2(u64& d, u64 a, u64 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
// scalar weight x packed pair: ptxas folds the {w, w} pair into the scalar-broadcast operand form
__device__ __forceinline__ void ffma2s(u64& d, float w, u64 b) {
  u64 ww;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(ww), "l"(b));
}
__device__ __forceinline__ u64 fmul2s(float w, u64 b) {
  u64 ww, r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(ww), "l"(b));
  return r;
}
__device__ __forceinline__ u64 pack2(float x, float y) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float2 unpack2(u64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

// ---- producer ------------------------------------------------------------------------------
// pull the leading edge of a future slab (2 rows of c2 and flow) towards L2, one 128-byte line per lane and step
// (c1 needs no help: its TMA copies are issued a ring length ahead)
template <typename G, typename T, typename TF>
__device__ __forceinline__ void prefetch_rows(const T* c2, const TF* flow, const Params& P, const Seg& sg, int y0,
                                              int lane) {
  const int xlo = max(sg.x0 - 8, 0), xhi = min(sg.x0 + G::TW + 8, P.W);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int y = y0 + r;
    if (y < 0 || y >= P.H) continue;
    const size_t rowpix = ((size_t)sg.b * P.H + y) * P.W;
    const char* lo = reinterpret_cast<const char*>(c2 + (rowpix + xlo) * P.C);
    const int nb = (xhi - xlo) * P.C * (int)sizeof(T);
    for (int o = lane * 128; o < nb; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(lo + o));
    if (flow != nullptr) {
      const char* fl = reinterpret_cast<const char*>(flow + (rowpix + xlo) * 2);
      const int fb = (xhi - xlo) * 2 * (int)sizeof(TF);
      if (lane * 128 < fb) asm volatile("prefetch.global.L2 [%0];" ::"l"(fl + lane * 128));
    }
  }
}

// c1 pair `ag` (global pair index; segment looked up in the table) by TMA, issued by a single lane: one bulk copy per
// (row, PX-pixel group), the elements exactly as stored (fp32, or 16-bit widened later by the compute warps).
// Used for the initial fill and by the consumer that frees a slot (credit return: no producer ever waits for c1 space).
template <typename G, typename T>
__device__ __forceinline__ void refill_pair_tma(const T* __restrict__ c1, const Params& P, const Seg* segs, int nseg,
                                                int ag, unsigned char* __restrict__ sA, uint64_t* fullA) {
  static_assert(sizeof(T) * 4 == G::AB, "c1 ring element size");
  int si = 0;
  while (si + 1 < nseg && segs[si + 1].a0 <= ag) ++si;
  const int b = segs[si].b, x0 = segs[si].x0, k = segs[si].p0 + (ag - segs[si].a0);
  uint64_t* bar = &fullA[ag % (2 * G::NAP)];
  unsigned char* slot = sA + (size_t)(ag % G::NAP) * G::ASLOT * G::AB;
  const int rows = min(2, P.H - 2 * k), px = min(G::TW, P.W - x0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // consumers' generic reads of the slot come first
  mbar_arrive_expect_tx(bar, (uint32_t)(rows * px * P.C * (int)sizeof(T)));
  for (int row = 0; row < rows; ++row)
    for (int xg = 0; xg * G::PX < px; ++xg) {
      const int x = x0 + G::PX * xg, npx = min(G::PX, P.W - x);
      bulk_g2s(slot + (size_t)(row * G::AROW + xg * G::AXG) * G::AB,
               c1 + (((size_t)b * P.H + 2 * k + row) * P.W + x) * P.C, (uint32_t)(npx * P.C * (int)sizeof(T)), bar);
    }
}

// c1 rows 2k, 2k+1 of the strip by LDG / STS (16-bit inputs are widened to fp32 on the way).
// All loads are issued before any use (one latency exposure).
template <typename G, typename T>
__device__ __forceinline__ void produce_pair(const T* __restrict__ c1, const Params& P, const Seg& sg, int k,
                                             float4* __restrict__ sA_slot, int lane) {
  constexpr int ITEMS = 2 * G::TW * G::KV / 32;  // per lane
  float4 v[ITEMS];
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int id = it * 32 + lane;
    const int kv = id % G::KV, pxl = id / G::KV;
    const int row = pxl / G::TW, col = pxl - row * G::TW;
    const int y = 2 * k + row, x = sg.x0 + col;
    v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < P.H && x < P.W) v[it] = ld4(c1 + (((size_t)sg.b * P.H + y) * P.W + x) * P.C + 4 * kv);
  }
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int id = it * 32 + lane;
    const int kv = id % G::KV, pxl = id / G::KV;
    const int row = pxl / G::TW, col = pxl - row * G::TW;
    sA_slot[row * G::AROW + (col / G::PX) * G::AXG + (col % G::PX) * G::KV + kv] = v[it];
  }
}

// warped rows 2m-1 (bsel 0) and 2m (bsel 1), u = 0..BU-1 <-> x = x0 - 4 + u.
// wsc[e] = the 4 bilinear weights of slab pixel e (already times 1/C for fp32; zero outside the image),
// psc[e] = {linear index of the top-left corner, float4 offset of the pixel inside the slab slot}.
// Software pipeline: the corner loads of item i + GD are issued before item i is blended, so every lane keeps
// 4 * GD .. 4 * (GD + 1) LDG.128 in flight for the whole slab instead of one latency exposure per batch (the loop is
// fully unrolled; the ring of corner registers is indexed at compile time).
template <typename G, bool HAS_FLOW, typename T>
__device__ __forceinline__ void gather_slab(const T* __restrict__ c2, const Params& P, float4* __restrict__ sB_slot,
                                            const float4* __restrict__ wsc, const int2* __restrict__ psc, int lane,
                                            float inv_c) {
  // LPP lanes per pixel, KQ channel chunks per lane
  constexpr int PPI = 32 / G::LPP;               // pixels per warp iteration
  constexpr int NITEMS = G::NPIX * G::KQ / PPI;  // per-lane items of the slab
  constexpr int GD = HAS_FLOW ? GATHER_DEPTH : 6;  // items in flight ahead of the one being blended
  static_assert((G::NPIX * G::KQ) % PPI == 0 && NITEMS > GD, "gather items");
  constexpr int CELT = G::KV * 4;                // channels (compile-time: corner offsets become immediates)
  constexpr bool HALF = ElemTraits<T>::kIsHalf;
  const int sub = lane / G::LPP, kvl = lane % G::LPP;
  const T* c2k = c2 + 4 * kvl;
  const T* c2k1 = c2k + (size_t)P.W * CELT;
  float4 tl[GD + 1], tr[GD + 1], bl[GD + 1], br[GD + 1];
  int off[GD + 1];
#pragma unroll
  for (int i = 0; i < NITEMS + GD; ++i) {
    if (i < NITEMS) {  // item -> (pixel group, kv chunk): the chunks of a pixel are adjacent items
      const int r = i % (GD + 1);
      const int e = (i / G::KQ) * PPI + sub, q = i % G::KQ;
      const int2 ps = psc[e];
      off[r] = ps.y + kvl + q * G::LPP;   // negative: pixel outside the image (see slab_coords)
      const size_t o = (size_t)ps.x * CELT + 4 * q * G::LPP;
      tl[r] = ld4(c2k + o);
      if (HAS_FLOW) {
        tr[r] = ld4(c2k + o + CELT);
        bl[r] = ld4(c2k1 + o);
        br[r] = ld4(c2k1 + o + CELT);
      }
    }
    if (i >= GD) {
      const int c = i - GD, r = c % (GD + 1);
      const int e = (c / G::KQ) * PPI + sub;
      const float4 w = wsc[e];
      u64 lo = fmul2s(w.x, pack2(tl[r].x, tl[r].y)), hi = fmul2s(w.x, pack2(tl[r].z, tl[r].w));
      if (HAS_FLOW) {
        ffma2s(lo, w.y, pack2(tr[r].x, tr[r].y)); ffma2s(hi, w.y, pack2(tr[r].z, tr[r].w));
        ffma2s(lo, w.z, pack2(bl[r].x, bl[r].y)); ffma2s(hi, w.z, pack2(bl[r].z, bl[r].w));
        ffma2s(lo, w.w, pack2(br[r].x, br[r].y)); ffma2s(hi, w.w, pack2(br[r].z, br[r].w));
      }
      const float2 a = unpack2(lo), b = unpack2(hi);
      float4 v = make_float4(a.x, a.y, b.x, b.y);
      if (HALF) {  // the reference's warp output is a 16-bit tensor: round, then apply the channel-mean factor
        v = round4<T>(v);
        v.x *= inv_c; v.y *= inv_c; v.z *= inv_c; v.w *= inv_c;
      }
      // outside the image the cost volume sees zeros whatever the features hold: select, do not multiply (0 * NaN)
      const bool outside = off[r] < 0;
      sB_slot[off[r] & 0x7fffffff] = outside ? make_float4(0.f, 0.f, 0.f, 0.f) : v;
    }
  }
}

// flow of the slab's pixels (lane handles pixels e = i * 32 + lane); loads only, consumed by slab_coords
template <typename G, bool HAS_FLOW, typename TF>
struct SlabFlow {
  static constexpr int NP1 = (G::NPIX + 31) / 32;
  TF f0[NP1], f1[NP1];
  __device__ __forceinline__ void load(const TF* __restrict__ flow, const Params& P, const Seg& sg, int m, int lane) {
#pragma unroll
    for (int i = 0; i < NP1; ++i) {
      const int e = i * 32 + lane;
      const int bsel = e / G::BU, u = e - bsel * G::BU;
      const int y = 2 * m - 1 + bsel, x = sg.x0 - R + u;
      const bool inside = e < G::NPIX && y >= 0 && y < P.H && x >= 0 && x < P.W;
      f0[i] = f1[i] = TF(0.0f);
      if (HAS_FLOW && inside) {
        const size_t p = ((size_t)sg.b * P.H + y) * P.W + x;
        f0[i] = flow[2 * p];
        f1[i] = flow[2 * p + 1];
      }
    }
  }
};

// per pixel weights / corner index -> per-warp scratch
template <typename G, bool HAS_FLOW, typename T, typename TF>
__device__ __forceinline__ void slab_coords(const SlabFlow<G, HAS_FLOW, TF>& F, const Params& P, const Seg& sg, int m,
                                            float4* __restrict__ wsc, int2* __restrict__ psc, int lane, float inv_c) {
  const float wscale = ElemTraits<T>::kIsHalf ? 1.0f : inv_c;
#pragma unroll
  for (int i = 0; i < SlabFlow<G, HAS_FLOW, TF>::NP1; ++i) {
    const int e = i * 32 + lane;
    if (e < G::NPIX) {
      const int bsel = e / G::BU, u = e - bsel * G::BU;
      const int y = 2 * m - 1 + bsel, x = sg.x0 - R + u;
      const bool inside = y >= 0 && y < P.H && x >= 0 && x < P.W;
      // outside the image the cost volume sees zeros (padding after warping): zero weights on a valid address
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      int pix = sg.b * P.H * P.W;
      if (inside) {
        if (HAS_FLOW) {
          const SampleCoord s = sample_coord_v<T, TF>(F.f0[i], F.f1[i], P.flow_scale, sg.b, y, x, P.H, P.W);
          pix = s.pix;
          const float w11 = s.ax * s.ay;
          const float w01 = s.ax - w11, w10 = s.ay - w11;
          const float w00 = (1.0f - s.ax) - w10;
          w = make_float4(w00 * wscale, w01 * wscale, w10 * wscale, w11 * wscale);
        } else {
          pix += y * P.W + x;
          w.x = wscale;
        }
      }
      wsc[e] = w;
      psc[e] = make_int2(pix, (bsel * G::BROW + u * G::KV + (u / G::PX) * G::BXPAD) | (inside ? 0 : (int)0x80000000));
    }
  }
}

// ---- per-lane correlation: 1 row x 8 pixels x 9 dx against ONE warped row, kv in [kv0, kv1) ------------
// acc[p][d] = (sum over even channel pairs, sum over odd channel pairs) of c1[p] * warped[p + d]
// 4 channels of c1 at ring position i, widened to two f32x2 operands
template <typename G, typename T>
__device__ __forceinline__ ulonglong2 load_a(const unsigned char* __restrict__ ap, int i) {
  if constexpr (G::AB == 16) {
    return reinterpret_cast<const ulonglong2*>(ap)[i];
  } else {
    const uint2 r = reinterpret_cast<const uint2*>(ap)[i];
    float2 lo, hi;
    if constexpr (std::is_same<T, __half>::value) {
      lo = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
      hi = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    } else {  // bf16: the upper 16 bits of an fp32
      lo = make_float2(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u));
      hi = make_float2(__uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
    }
    ulonglong2 o;
    o.x = pack2(lo.x, lo.y);
    o.y = pack2(hi.x, hi.y);
    return o;
  }
}

template <typename G, typename T>
__device__ __forceinline__ void correlate(const unsigned char* __restrict__ ap, const ulonglong2* __restrict__ bp,
                                          int kv0, int kv1, u64 (&acc)[G::PX][ND]) {
  // Unrolled by KU channel quads: with a rolled loop ptxas rotates the 72 accumulator pairs through ~100 MOVs per
  // iteration.  kv1 - kv0 is KV or KV / KP, both multiples of KU.
  constexpr int KU = G::KU;
  if (G::KP == 1) { kv0 = 0; kv1 = G::KV; }  // compile-time trip count
#pragma unroll 1
  for (int kvb = kv0; kvb < kv1; kvb += KU) {
#pragma unroll
    for (int ku = 0; ku < KU; ++ku) {
      const int kv = kvb + ku;
      ulonglong2 a[G::PX];
#pragma unroll
      for (int p = 0; p < G::PX; ++p) a[p] = load_a<G, T>(ap, kv + p * G::KV);
#pragma unroll
      for (int j = 0; j < G::PX + 2 * R; ++j) {
        const ulonglong2 b = bp[kv + j * G::KV + (j / G::PX) * G::BXPAD];
        // all first halves, then all second halves: the two FFMA2 of one accumulator are >= 8 issue slots apart
#pragma unroll
        for (int p = 0; p < G::PX; ++p) {
          const int d = j - p;
          if (d >= 0 && d < ND) ffma2(acc[p][d], a[p].x, b.x);
        }
#pragma unroll
        for (int p = 0; p < G::PX; ++p) {
          const int d = j - p;
          if (d >= 0 && d < ND) ffma2(acc[p][d], a[p].y, b.y);
        }
      }
    }
  }
}

template <int KV, int NXG, int PXL, bool HAS_FLOW, bool ATMA, typename T, typename TF>
__global__ void __launch_bounds__(32 * (NCW + NPW), 1)
fwd_stream_kernel(const T* __restrict__ c1, const T* __restrict__ c2, const TF* __restrict__ flow,
                  T* __restrict__ out, const Params P) {
  using G = Geo<KV, NXG, PXL, ATMA ? (int)sizeof(T) * 4 : 16>;
  extern __shared__ __align__(128) unsigned char smem[];
  float4* sB = reinterpret_cast<float4*>(smem + G::OFF_B);
  unsigned char* sA = smem + G::OFF_A;
  float* sStage = reinterpret_cast<float*>(smem + G::OFF_STAGE);
  // Two alternating barrier sets per data slot (index j % 2N, parity (j / 2N) & 1): with one set a waiter that
  // is two uses ahead of a slow producer would pass a parity wait prematurely (tests/stream_protocol_sim.py).
  constexpr int NBB = 2 * G::NSLAB, NBA = 2 * G::NAP;
  uint64_t* fullB = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);
  uint64_t* emptyB = fullB + NBB;
  uint64_t* fullA = emptyB + NBB;
  uint64_t* emptyA = fullA + NBA;
  Seg* segs = reinterpret_cast<Seg*>(smem + G::OFF_SEG);
  int* s_misc = reinterpret_cast<int*>(smem + G::OFF_SEG + MAXSEG * 32);  // [0] nseg, [1] total tasks, [2] task counter, [3] pairs
  int* a_cnt = reinterpret_cast<int*>(smem + G::OFF_ACNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int tr_n = 0;

  if (tid == 0) {
    for (int i = 0; i < NBB; ++i) { mbar_init(&fullB[i], 32); mbar_init(&emptyB[i], G::EMPTYB_COUNT); }
    for (int i = 0; i < NBA; ++i) { mbar_init(&fullA[i], ATMA ? 1 : 32); mbar_init(&emptyA[i], 6); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async (TMA) proxy
    // this CTA's contiguous range of (strip, pair) units, cut into per-strip segments
    const long long qlo = P.Q * blockIdx.x / gridDim.x, qhi = P.Q * (blockIdx.x + 1) / gridDim.x;
    int n = 0, task = 0, j = 0, a = 0;
    long long q = qlo;
    while (q < qhi && n < MAXSEG) {
      const long long strip = q / P.NPAIR;
      const int p0 = (int)(q - strip * P.NPAIR);
      const long long rem = qhi - q;
      const int p1 = (int)((long long)(P.NPAIR - p0) < rem ? P.NPAIR : p0 + rem);
      Seg s;
      s.b = (int)(strip / P.nstrip_x);
      s.x0 = (int)(strip % P.nstrip_x) * G::TW;
      s.p0 = p0; s.p1 = p1; s.nslab = p1 - p0 + 5;
      s.task0 = task; s.j0 = j; s.a0 = a;
      task += G::TPG * ((s.nslab + GS - 1) / GS);
      j += s.nslab;
      a += p1 - p0;
      segs[n++] = s;
      q += p1 - p0;
    }
    s_misc[0] = n; s_misc[1] = task; s_misc[2] = 0; s_misc[3] = a;
    for (int i = 0; i < G::NAP; ++i) a_cnt[i] = 0;
  }
  __syncthreads();
  const int nseg = s_misc[0];

  if (warp >= G::NCWG) {
    // =========================== producers ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G::REGS_PRODUCER));
    const int pwarp = warp - G::NCWG;
    float4* wsc = reinterpret_cast<float4*>(smem + G::OFF_SCRW) + pwarp * G::NPIX;
    int2* psc = reinterpret_cast<int2*>(smem + G::OFF_SCRP) + pwarp * G::NPIX;
    const float inv_c = 1.0f / (float)P.C;
    // slabs owned by this warp, in order; the flow of the next owned slab is loaded while the current one is gathered
    int si = 0, l = -1;
    auto advance = [&]() {  // -> next (si, l) with (j0 + l) % NPW == pwarp, or si == nseg
      for (;;) {
        if (si >= nseg) return;
        ++l;
        if (l >= segs[si].nslab) { ++si; l = -1; continue; }
        if ((segs[si].j0 + l) % NPW == pwarp) return;
      }
    };
    auto slab_live = [&](int m) { return 2 * m >= 0 && 2 * m - 1 < P.H; };  // else entirely outside the image
    if constexpr (ATMA) {  // initial fill of the c1 ring; afterwards the consumer that frees a slot refills it
      const int npairs = s_misc[3];
      if (pwarp == 0 && lane < G::NAP && lane < npairs) refill_pair_tma<G, T>(c1, P, segs, nseg, lane, sA, fullA);
    }
    SlabFlow<G, HAS_FLOW, TF> F;
    advance();
    if (si < nseg && slab_live(segs[si].p0 - 2 + l)) F.load(flow, P, segs[si], segs[si].p0 - 2 + l, lane);
    while (si < nseg) {
      const Seg sg = segs[si];
      const int jg = sg.j0 + l, m = sg.p0 - 2 + l;
      PWC_TR(1, jg);
      if (P.pf_dist > 0) prefetch_rows<G, T, TF>(c2, flow, P, sg, 2 * (m + P.pf_dist) + 2, lane);
      if constexpr (!ATMA) {
        if (sg.p0 + l < sg.p1) {
          const int ag = sg.a0 + l, slot = ag % G::NAP, prev = ag - G::NAP;
          if (prev >= 0) mbar_wait(&emptyA[prev % NBA], (prev / NBA) & 1);  // previous occupant fully consumed
          if (!(P.dbg & 2))
            produce_pair<G, T>(c1, P, sg, sg.p0 + l, reinterpret_cast<float4*>(sA) + (size_t)slot * G::ASLOT, lane);
          mbar_arrive(&fullA[ag % NBA]);
        }
      }
      PWC_TR(2, jg);
      const int slot = jg % G::NSLAB, prev = jg - G::NSLAB;
      const bool live = slab_live(m) && !(P.dbg & 2);
      __syncwarp();  // every lane is done reading the previous slab's scratch
      if (live) slab_coords<G, HAS_FLOW, T, TF>(F, P, sg, m, wsc, psc, lane, inv_c);
      __syncwarp();
      advance();
      if (si < nseg && slab_live(segs[si].p0 - 2 + l)) F.load(flow, P, segs[si], segs[si].p0 - 2 + l, lane);
      PWC_TR(3, jg);
      if (prev >= 0) mbar_wait(&emptyB[prev % NBB], (prev / NBB) & 1);
      PWC_TR(4, jg);
      if (live) gather_slab<G, HAS_FLOW, T>(c2, P, sB + (size_t)slot * G::BSLOT, wsc, psc, lane, inv_c);
      mbar_arrive(&fullB[jg % NBB]);
      PWC_TR(5, jg);
    }
    return;
  }

  // =========================== compute warps ===========================
  // Full task f of group g: lane = lo + 2*hi + 2*NH*xg ; combo c = f*NH + hi -> slab 4g + c/8, c1 row 2m-4 + c%8,
  // lo = warped row of the slab.  Edge task: lane = lo + 2*xg + 2*NXG*hi, hi = (slab ms, channel part); lo = half:
  //   half 0: c1 row 2m+4 vs warped row 2m (dy=-4) ; half 1: c1 row 2m-5 vs warped row 2m-1 (dy=+4).
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(G::REGS_COMPUTE));
  float* stage = sStage + warp * G::STAGE_WORDS;
  const int total_tasks = s_misc[1];
  const int lo = lane & 1;
  while (true) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&s_misc[2], 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= total_tasks) break;
    PWC_TR(8, t);
    int si = 0;
    while (si + 1 < nseg && segs[si + 1].task0 <= t) ++si;
    const Seg sg = segs[si];
    const int tl = t - sg.task0, g = tl / G::TPG, rr = tl - G::TPG * g;
    const bool is_edge = rr == G::NF;

    int l, k, sub, bsel, xg, kv0 = 0, kv1 = KV;
    bool arrive_b, arrive_a;
    if (!is_edge) {
      const int hi = (lane >> 1) % G::NH;
      xg = lane / (2 * G::NH);
      const int c = rr * G::NH + hi;
      l = GS * g + (c >> 3);
      const int r8 = c & 7;
      k = sg.p0 - 2 + l - 2 + (r8 >> 1);
      sub = r8 & 1;
      bsel = lo;
      arrive_b = lo == 0 && xg == 0 && (hi == 0 || r8 == 0);
      arrive_a = lo == 0 && xg == 0 && sub == 0;
    } else {
      const int hi = lane / (2 * NXG);
      xg = (lane >> 1) % NXG;
      const int ms = hi / G::KP, part = hi % G::KP;
      l = GS * g + ms;
      k = sg.p0 - 2 + l + (lo == 0 ? 2 : -3);
      sub = lo;
      bsel = 1 - lo;
      kv0 = part * (KV / G::KP);
      kv1 = kv0 + KV / G::KP;
      arrive_b = lo == 0 && xg == 0 && part == 0;
      arrive_a = xg == 0 && part == 0;
    }
    if (!is_edge && GS * g + ((rr * G::NH) >> 3) >= sg.nslab) continue;  // null task (warp-uniform)
    const bool slab_ok = l < sg.nslab;
    const int m = sg.p0 - 2 + l;
    const bool pair_ok = slab_ok && k >= sg.p0 && k < sg.p1;
    const int jg = sg.j0 + (slab_ok ? l : 0), sb = jg % G::NSLAB;
    const int ag = sg.a0 + (pair_ok ? k - sg.p0 : 0), sa = ag % G::NAP;
    if (slab_ok) mbar_wait(&fullB[jg % NBB], (jg / NBB) & 1);
    if (pair_ok) mbar_wait(&fullA[ag % NBA], (ag / NBA) & 1);
    PWC_TR(9, t);
    const int yb = 2 * m - 1 + bsel, ya = 2 * k + sub;
    const bool store = pair_ok && ya < P.H;
    const bool active = store && yb >= 0 && yb < P.H;

    u64 acc[G::PX][ND];
#pragma unroll
    for (int p = 0; p < G::PX; ++p)
#pragma unroll
      for (int d = 0; d < ND; ++d) acc[p][d] = 0ull;
    if (active && !(P.dbg & 1)) {
      const unsigned char* ap = sA + ((size_t)sa * G::ASLOT + sub * G::AROW + xg * G::AXG) * G::AB;
      const ulonglong2* bp = reinterpret_cast<const ulonglong2*>(sB + (size_t)sb * G::BSLOT + bsel * G::BROW + xg * G::BXG);
      correlate<G, T>(ap, bp, kv0, kv1, acc);
    }
    __syncwarp();
    PWC_TR(10, t);
    // release the ring slots: one arrival per (task, slab) and per (task, slab, pair)
    if (slab_ok && arrive_b) mbar_arrive(&emptyB[jg % NBB]);
    if (pair_ok && arrive_a) {
      if constexpr (ATMA) {
        // The arrival is an acq_rel atomic at CTA scope: the other warps' reads of the slot (their arrivals release them)
        // happen-before the refilling lane's fence.proxy.async + TMA overwrite.  CTA scope orders shared-memory accesses
        // only; it does not wait for the previous task's global stores.
        if (atom_add_acq_rel_cta(&a_cnt[sa], 1) == 5) {  // last of the 6 arrivals: the slot is free, refill it with pair ag + NAP
          a_cnt[sa] = 0;
          if (ag + G::NAP < s_misc[3]) refill_pair_tma<G, T>(c1, P, segs, nseg, ag + G::NAP, sA, fullA);
          if ((P.dbg & 16) && blockIdx.x == 0 && lane == 0 && tr_n < TRACE_LEN) g_trace[warp][tr_n++] = ((unsigned long long)clock64() << 16) | (13 << 12) | (lane & 0xfff);
        }
      } else {
        mbar_arrive(&emptyA[ag % NBA]);
      }
    }

    // tasks that only exist to keep the ring arithmetic regular (warm-up slabs above / below the image) write nothing
    if (!__any_sync(0xffffffffu, store)) {
      PWC_TR(11, t);
      continue;
    }
    // epilogue: lane L owns, for each pixel p, the 9 channels [dyi*9, dyi*9+9) of pixel (ya, xb+p).
    // Runs go through a per-warp transposition buffer so that each store instruction writes whole runs.
    const int xb = sg.x0 + xg * G::PX;
    const int dyi = yb - ya + R;
    const bool writer = store && kv0 == 0;
    const long long my_off = writer ? (long long)((((size_t)sg.b * P.H + ya) * P.W + xb) * P.ops + dyi * ND) : 0;
    const int my_nv = writer ? P.W - xb : 0;
    auto stage_pixel = [&](int p, float* stg) {
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const float2 h = unpack2(acc[p][d]);
        float v = h.x + h.y;
        if (G::KP > 1 && is_edge) {  // channel parts sit in adjacent hi values: butterfly over those lane bits
          if (G::KP >= 2) v += __shfl_xor_sync(0xffffffffu, v, 2 * NXG);
          if (G::KP >= 4) v += __shfl_xor_sync(0xffffffffu, v, 4 * NXG);
        }
        stg[lane * 9 + d] = fmaxf(0.1f * v, v);
      }
    };
    {
      T* optr[9];
      int nv[9];
#pragma unroll
      for (int it = 0; it < 9; ++it) {
        const int idx = it * 32 + lane;
        const int l2 = idx / 9, w = idx - l2 * 9;
        const long long o = __shfl_sync(0xffffffffu, my_off, l2);
        nv[it] = __shfl_sync(0xffffffffu, my_nv, l2);
        optr[it] = out + o + w;
      }
      // software pipeline over the pixels: pixel p + 1 is staged while pixel p is read back and stored
      stage_pixel(0, stage);
      __syncwarp();
#pragma unroll
      for (int p = 0; p < G::PX; ++p) {
        if (p + 1 < G::PX) stage_pixel(p + 1, stage + ((p + 1) & 1) * 320);
        const float* stg = stage + (p & 1) * 320;
#pragma unroll
        for (int it = 0; it < 9; ++it)
          if (p < nv[it] && !(P.dbg & 4)) optr[it][(size_t)p * P.ops] = ElemTraits<T>::from_f(stg[it * 32 + lane]);
        __syncwarp();
      }
    }
    PWC_TR(11, t);
  }
}

// Host-side launcher.  Returns 0 when launched, 1 when the shape is not covered, -1 on a CUDA error.
template <int KV, int NXG, int PXL, typename T, typename TF>
int launch_cfg(const T* c1, const T* c2, const TF* flow, float flow_scale, T* out, int B, int H, int W, int C,
               int64_t ops, cudaStream_t st, int sms, int ctas_override, int pf_dist, int dbg) {
  using G = Geo<KV, NXG, PXL, 16>;  // strip geometry (the c1 ring element size only changes the shared-memory budget)
  Params P;
  P.B = B; P.H = H; P.W = W; P.C = C; P.ops = ops; P.flow_scale = flow_scale;
  P.NPAIR = (H + 1) / 2;
  P.nstrip_x = (W + G::TW - 1) / G::TW;
  P.Q = (long long)B * P.nstrip_x * P.NPAIR;
  P.pf_dist = pf_dist;
  P.dbg = dbg;
  if (P.Q <= 0) return 1;
  long long grid = ctas_override > 0 ? ctas_override : sms;
  if (grid > P.Q) grid = P.Q;
  // every CTA must fit its range into MAXSEG per-strip segments
  const long long per = (P.Q + grid - 1) / grid;
  if (per / P.NPAIR + 2 > MAXSEG) return 1;
  auto go = [&](auto kern, size_t smem_bytes) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess)
      return -1;
    kern<<<(unsigned)grid, G::NTHREADS, smem_bytes, st>>>(c1, c2, flow, out, P);
    return 0;
  };
  // TMA needs 16-byte aligned global rows: C % 8 == 0 for 16-bit elements, C % 4 == 0 for fp32, and a 16-byte aligned
  // base (a 16-bit view may start on an 8-byte boundary: LDG path then); stream_debug bit 8 forces the LDG/STS c1 path
  if (!(dbg & 8) && (C * (int)sizeof(T)) % 16 == 0 && reinterpret_cast<uintptr_t>(c1) % 16 == 0) {
    using GT = Geo<KV, NXG, PXL, (int)sizeof(T) * 4>;
    return flow != nullptr ? go(fwd_stream_kernel<KV, NXG, PXL, true, true, T, TF>, GT::SMEM_BYTES)
                           : go(fwd_stream_kernel<KV, NXG, PXL, false, true, T, TF>, GT::SMEM_BYTES);
  }
  return flow != nullptr ? go(fwd_stream_kernel<KV, NXG, PXL, true, false, T, TF>, G::SMEM_BYTES)
                         : go(fwd_stream_kernel<KV, NXG, PXL, false, false, T, TF>, G::SMEM_BYTES);
}

// Covered: any output pixel stride >= 81 (a channel slice of the concat buffer, model_pwcnet.py:1424),
// C in {16, 32, 64, 96, 128}.  Everything else returns 1 (caller falls back).
template <typename T, typename TF>
int launch(const T* c1, const T* c2, const TF* flow, float flow_scale, T* out, int B, int H, int W, int C,
           int64_t ops, cudaStream_t st, int sms, int ctas_override, int pf_dist, int dbg = 0) {
  if (ops < D) return 1;
#define PWC_STREAM_GO(KV, NXG, PXL) \
  return launch_cfg<KV, NXG, PXL, T, TF>(c1, c2, flow, flow_scale, out, B, H, W, C, ops, st, sms, ctas_override, pf_dist, dbg)
  switch (C) {
    case 16: PWC_STREAM_GO(4, 4, 6);
    case 32: PWC_STREAM_GO(8, 4, 6);
    case 64: PWC_STREAM_GO(16, 2, 6);
    case 96: PWC_STREAM_GO(24, 2, 6);
    case 128: PWC_STREAM_GO(32, 1, 6);
    default: return 1;
  }
#undef PWC_STREAM_GO
}

}  // namespace stream
}  // namespace pwc
