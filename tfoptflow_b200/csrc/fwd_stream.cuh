// Streaming fused warp + cost-volume forward (search_range 4, C in {16, 32, 64, 96, 128}).
//
// Persistent, warp-specialised kernel, one CTA per SM.  The image is cut into vertical strips of
// TW = 8 * NXG pixels (32 / 16 / 8 for C <= 32 / 64 / >= 96, so that all channels of the rows in flight fit in
// shared memory); every CTA owns a contiguous range of (strip, row-pair) work and marches down it.
//   * NPW producer warps evaluate the bilinear warp of c2 (8 lanes per pixel read whole 128-byte lines of every
//     corner) and stream the *warped* rows ("slabs" = warped rows 2m-1, 2m), plus the matching c1 row pairs, into
//     two shared-memory rings guarded by mbarriers.  Warped features never reach HBM.  Each producer warp owns every
//     NPW-th slab, so NPW gathers are in flight.
//   * NCW compute warps pull tasks in order.  Every task is 32 independent lane-tasks: one c1 row against one
//     warped row, 8 pixels x 9 dx, 72 accumulators, every LDS.128 operand feeding up to 36 FMAs.  Lanes are laid out
//     (warped row, x-group, c1 row) so that c1 operands are shared by the two warped rows of a slab and warped
//     operands by all c1 rows (LDS broadcast).  A slab sees c1 rows 2m-4 .. 2m+3 (full tasks F); the two remaining
//     displacements (row 2m+4 with dy=-4, row 2m-5 with dy=+4) are done by one edge task E per 4 slabs, with the
//     channel range split over 4/NXG lanes and reduced by shuffles.
//   * Results go through a per-warp transposition buffer: each store instruction writes runs of 9 / 18 channels.
// Zero padding is applied after warping (core_costvol.py:27); the warp clamps (core_warp.py:101-104).
// The ring protocol (two alternating mbarrier sets per slot) is modelled in tests/stream_protocol_sim.py.
#pragma once
#include "common.cuh"

namespace pwc {
namespace stream {

constexpr int R = 4, ND = 9, D = 81, PX = 8;
constexpr int NCW = 8;   // compute warps
constexpr int NPW = 4;   // producer warps
constexpr int NTHREADS = 32 * (NCW + NPW);
constexpr int MAXSEG = 8;
constexpr int GS = 4;    // slabs per task group (one edge task per group)
constexpr size_t SMEM_LIMIT = 232448;

template <int KV_, int NXG_>
struct Geo {
  static constexpr int KV = KV_, NXG = NXG_;
  static constexpr int TW = PX * NXG;                 // strip width
  static constexpr int BU = TW + 2 * R;               // warped pixels per row (with halo)
  static constexpr int NPIX = 2 * BU;                 // pixels per slab
  static constexpr int B_USLOTS = 9 * NXG + 8;        // slot(u) = u + u/8
  static constexpr int A_ROWSLOTS = 9 * NXG;
  static constexpr int B_PLANE = B_USLOTS * 2 + 1;    // float4 per kv plane, odd => kv-strided stores conflict-free
  static constexpr int A_PLANE = A_ROWSLOTS * 2 + 1;
  static constexpr int BSLOT = KV * B_PLANE;
  static constexpr int ASLOT = KV * A_PLANE + 1;      // +1: consecutive pair slots start in different banks
  static constexpr int NH = 16 / NXG;                 // c1 rows (hi values) per full task
  static constexpr int NF = 8 * GS / NH;              // full tasks per group: 8 / 4 / 2
  static constexpr int TPG = NF + 1;                  // + one edge task
  static constexpr int KP = 4 / NXG;                  // channel split of the edge task
  static constexpr int EMPTYB_COUNT = (NXG == 4 ? 2 : 1) + 1;
  static constexpr int LPP = KV >= 8 ? 8 : KV;        // producer lanes per pixel
  static constexpr int KQ = KV / LPP;                 // kv chunks per lane
  static constexpr int NSLAB = KV <= 8 ? 7 : (KV == 24 ? 7 : 6);
  static constexpr int NAP = KV <= 8 ? 12 : (KV == 32 ? 10 : 12);
  static constexpr int STAGE_WORDS = 640;             // per compute warp (2 x 320, double buffered)
  static constexpr size_t OFF_B = 0;
  static constexpr size_t OFF_A = OFF_B + (size_t)NSLAB * BSLOT * 16;
  static constexpr size_t OFF_STAGE = OFF_A + (size_t)NAP * ASLOT * 16;
  static constexpr size_t OFF_SCRATCH = OFF_STAGE + (size_t)NCW * STAGE_WORDS * 4;
  static constexpr size_t OFF_BAR = OFF_SCRATCH + (size_t)NPW * NPIX * 16;
  static constexpr size_t OFF_SEG = OFF_BAR + (size_t)(4 * NSLAB + 4 * NAP) * 8;
  static constexpr size_t SMEM_BYTES = OFF_SEG + MAXSEG * 32 + 64;
  static_assert(SMEM_BYTES <= SMEM_LIMIT, "shared memory budget");
  static_assert(KV % LPP == 0 && KV % KP == 0, "channel split");
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
  int dbg;          // timing experiments only: 1 = skip correlate, 2 = skip gathers, 4 = skip stores
};

// ---- mbarrier helpers (generic-proxy producers/consumers only) ---------------------------
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

// ---- producer ------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// pull the leading edge of a future slab (2 rows of c2, c1 and flow) towards L2
template <typename G, typename T, typename TF>
__device__ __forceinline__ void prefetch_rows(const T* c1, const T* c2, const TF* flow, const Params& P, const Seg& sg,
                                              int y0, int lane) {
  // lanes 0,1: c2 rows y0, y0+1 ; lanes 2,3: c1 ; lanes 4,5: flow.  Ranges are rounded down to 16 bytes.
  const int r = lane & 1, what = lane >> 1;
  const int y = y0 + r;
  if (what > 2 || y < 0 || y >= P.H) return;
  if (what == 2 && flow == nullptr) return;
  const int xlo = max(sg.x0 - 8, 0), xhi = min(sg.x0 + G::TW + 8, P.W);
  const size_t rowpix = ((size_t)sg.b * P.H + y) * P.W;
  uintptr_t lo, hi;
  if (what == 2) {
    lo = reinterpret_cast<uintptr_t>(flow + (rowpix + xlo) * 2);
    hi = reinterpret_cast<uintptr_t>(flow + (rowpix + xhi) * 2);
  } else {
    const T* base = what == 0 ? c2 : c1;
    lo = reinterpret_cast<uintptr_t>(base + (rowpix + xlo) * P.C);
    hi = reinterpret_cast<uintptr_t>(base + (rowpix + xhi) * P.C);
  }
  lo &= ~uintptr_t(15);
  hi &= ~uintptr_t(15);
  if (hi > lo) bulk_prefetch_l2(reinterpret_cast<const void*>(lo), (uint32_t)(hi - lo));
}

// c1 rows 2k, 2k+1 of the strip, scaled by 1/C (the channel mean is folded into c1)
template <typename G, typename T>
__device__ __forceinline__ void produce_pair(const T* __restrict__ c1, const Params& P, const Seg& sg, int k,
                                             float4* __restrict__ sA_slot, int lane, float inv_c) {
  constexpr int ITEMS = 2 * G::TW * G::KV / 32;  // per lane
  constexpr int BATCH = ITEMS;                   // one batch: all loads in flight together (one latency exposure)
#pragma unroll 1
  for (int b0 = 0; b0 < ITEMS; b0 += BATCH) {
    float4 v[BATCH];
#pragma unroll
    for (int it = 0; it < BATCH; ++it) {
      const int id = (b0 + it) * 32 + lane;
      const int kv = id % G::KV, pxl = id / G::KV;
      const int row = pxl / G::TW, col = pxl - row * G::TW;
      const int y = 2 * k + row, x = sg.x0 + col;
      v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y < P.H && x < P.W) v[it] = ld4(c1 + (((size_t)sg.b * P.H + y) * P.W + x) * P.C + 4 * kv);
    }
#pragma unroll
    for (int it = 0; it < BATCH; ++it) {
      const int id = (b0 + it) * 32 + lane;
      const int kv = id % G::KV, pxl = id / G::KV;
      const int row = pxl / G::TW, col = pxl - row * G::TW;
      float4 w = v[it];
      w.x *= inv_c; w.y *= inv_c; w.z *= inv_c; w.w *= inv_c;
      sA_slot[kv * G::A_PLANE + row * G::A_ROWSLOTS + col + col / PX] = w;
    }
  }
}

// bilinear blend with FMA lerps (the fused path is checked against the oracle with a tolerance; the standalone
// dense_image_warp kernel keeps the reference's exact op order)
__device__ __forceinline__ float4 bilerp4_fast(const float4& tl, const float4& tr, const float4& bl, const float4& br,
                                               float ax, float ay) {
  float4 o;
  float t, b;
  t = fmaf(ax, tr.x - tl.x, tl.x); b = fmaf(ax, br.x - bl.x, bl.x); o.x = fmaf(ay, b - t, t);
  t = fmaf(ax, tr.y - tl.y, tl.y); b = fmaf(ax, br.y - bl.y, bl.y); o.y = fmaf(ay, b - t, t);
  t = fmaf(ax, tr.z - tl.z, tl.z); b = fmaf(ax, br.z - bl.z, bl.z); o.z = fmaf(ay, b - t, t);
  t = fmaf(ax, tr.w - tl.w, tl.w); b = fmaf(ax, br.w - bl.w, bl.w); o.w = fmaf(ay, b - t, t);
  return o;
}

// warped rows 2m-1 (bsel 0) and 2m (bsel 1), u = 0..BU-1 <-> x = x0 - 4 + u
template <typename G, bool HAS_FLOW, bool ALL_IN, typename T>
__device__ __forceinline__ void gather_slab(const T* __restrict__ c2, const Params& P, float4* __restrict__ sB_slot,
                                            const int4* __restrict__ scratch, int lane) {
  // LPP lanes per pixel, KQ channel chunks per lane; NIT items are loaded before any is used.
  // ALL_IN: every pixel of the slab lies inside the image (warp-uniform), so there is no per-item predicate.
  constexpr int PPI = 32 / G::LPP;               // pixels per warp iteration
  constexpr int NITEMS = G::NPIX * G::KQ / PPI;  // per-lane items of the slab
  constexpr int NIT = NITEMS % 4 == 0 ? 4 : 2;
  static_assert(NITEMS % NIT == 0, "gather batching");
  constexpr int CELT = G::KV * 4;                // channels (compile-time: corner offsets become immediates)
  const int sub = lane / G::LPP, kvl = lane % G::LPP;
  const size_t rowstride = (size_t)P.W * CELT;
#pragma unroll 1
  for (int i0 = 0; i0 < NITEMS; i0 += NIT) {
    float4 tl[NIT], tr[NIT], bl[NIT], br[NIT];
    int4 cd[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int item = i0 + it;  // item -> (pixel group, kv chunk): the chunks of a pixel are adjacent items
      const int e = (item / G::KQ) * PPI + sub, q = item % G::KQ;
      cd[it] = scratch[e];
      const int kv = kvl + q * G::LPP;
      cd[it].w += kv * G::B_PLANE;
      const bool ok = ALL_IN || cd[it].x >= 0;
      const T* base = c2 + (size_t)(ok ? cd[it].x : 0) * CELT + 4 * kv;
      tl[it] = ld4(base);
      if (HAS_FLOW) {
        tr[it] = ld4(base + CELT);
        bl[it] = ld4(base + rowstride);
        br[it] = ld4(base + rowstride + CELT);
      }
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      float4 v = tl[it];
      if (HAS_FLOW)
        v = round4<T>(bilerp4_fast(tl[it], tr[it], bl[it], br[it], __int_as_float(cd[it].y), __int_as_float(cd[it].z)));
      if (!ALL_IN && cd[it].x < 0) v = make_float4(0.f, 0.f, 0.f, 0.f);
      sB_slot[cd[it].w] = v;
    }
  }
}

template <typename G, bool HAS_FLOW, typename T, typename TF>
__device__ __forceinline__ void produce_slab(const T* __restrict__ c2, const TF* __restrict__ flow, const Params& P,
                                             const Seg& sg, int m, float4* __restrict__ sB_slot,
                                             int4* __restrict__ scratch, int lane) {
  __syncwarp();  // every lane is done reading the previous slab's scratch
  // phase 1: per pixel {top-left corner index or -1, ax, ay, smem offset} -> per-warp scratch.
  // All flow loads are issued before any is used (one latency exposure for the whole slab).
  constexpr int NP1 = (G::NPIX + 31) / 32;
  TF f0[NP1], f1[NP1];
  bool inside[NP1];
  bool all_in = true;
#pragma unroll
  for (int i = 0; i < NP1; ++i) {
    const int e = i * 32 + lane;
    const int bsel = e / G::BU, u = e - bsel * G::BU;
    const int y = 2 * m - 1 + bsel, x = sg.x0 - R + u;
    inside[i] = e < G::NPIX && y >= 0 && y < P.H && x >= 0 && x < P.W;
    if (e < G::NPIX && !inside[i]) all_in = false;
    f0[i] = f1[i] = TF(0.0f);
    if (HAS_FLOW && inside[i]) {
      const size_t p = ((size_t)sg.b * P.H + y) * P.W + x;
      f0[i] = flow[2 * p];
      f1[i] = flow[2 * p + 1];
    }
  }
#pragma unroll
  for (int i = 0; i < NP1; ++i) {
    const int e = i * 32 + lane;
    if (e < G::NPIX) {
      const int bsel = e / G::BU, u = e - bsel * G::BU;
      const int y = 2 * m - 1 + bsel, x = sg.x0 - R + u;
      int4 cd = make_int4(-1, 0, 0, (u + u / PX) * 2 + bsel);
      if (inside[i]) {
        if (HAS_FLOW) {
          const SampleCoord s = sample_coord_v<T, TF>(f0[i], f1[i], P.flow_scale, sg.b, y, x, P.H, P.W);
          cd.x = s.pix; cd.y = __float_as_int(s.ax); cd.z = __float_as_int(s.ay);
        } else {
          cd.x = (sg.b * P.H + y) * P.W + x;
        }
      }
      scratch[e] = cd;
    }
  }
  all_in = __all_sync(0xffffffffu, all_in);  // also orders the scratch writes before the reads below
  if (all_in) gather_slab<G, HAS_FLOW, true, T>(c2, P, sB_slot, scratch, lane);
  else        gather_slab<G, HAS_FLOW, false, T>(c2, P, sB_slot, scratch, lane);
}

// ---- per-lane correlation: 1 row x 8 pixels x 9 dx against ONE warped row, kv in [kv0, kv1) ------------
template <typename G>
__device__ __forceinline__ void correlate(const float4* __restrict__ ap, const float4* __restrict__ bp, int kv0,
                                          int kv1, float (&acc)[PX][ND]) {
#pragma unroll 1
  for (int kv = kv0; kv < kv1; ++kv) {
    // pixel-major order: the warped window slides (9 live float4), the c1 operand lives for one pixel only
    float4 bw[PX + 2 * R];
#pragma unroll
    for (int k = 0; k < ND; ++k) bw[k] = bp[kv * G::B_PLANE + (k + k / PX) * 2];
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      if (p > 0) bw[p + ND - 1] = bp[kv * G::B_PLANE + ((p + ND - 1) + (p + ND - 1) / PX) * 2];
      const float4 a = ap[kv * G::A_PLANE + p];
#pragma unroll
      for (int d = 0; d < ND; ++d) fma4(acc[p][d], a, bw[p + d]);
    }
  }
}

template <int KV, int NXG, bool HAS_FLOW, typename T, typename TF>
__global__ void __launch_bounds__(NTHREADS, 1)
fwd_stream_kernel(const T* __restrict__ c1, const T* __restrict__ c2, const TF* __restrict__ flow,
                  T* __restrict__ out, const Params P) {
  using G = Geo<KV, NXG>;
  extern __shared__ __align__(16) unsigned char smem[];
  float4* sB = reinterpret_cast<float4*>(smem + G::OFF_B);
  float4* sA = reinterpret_cast<float4*>(smem + G::OFF_A);
  float* sStage = reinterpret_cast<float*>(smem + G::OFF_STAGE);
  // Two alternating barrier sets per data slot (index j % 2N, parity (j / 2N) & 1): with one set a waiter that
  // is two uses ahead of a slow producer would pass a parity wait prematurely (tests/stream_protocol_sim.py).
  constexpr int NBB = 2 * G::NSLAB, NBA = 2 * G::NAP;
  uint64_t* fullB = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);
  uint64_t* emptyB = fullB + NBB;
  uint64_t* fullA = emptyB + NBB;
  uint64_t* emptyA = fullA + NBA;
  Seg* segs = reinterpret_cast<Seg*>(smem + G::OFF_SEG);
  int* s_misc = reinterpret_cast<int*>(smem + G::OFF_SEG + MAXSEG * 32);  // [0] nseg, [1] total tasks, [2] task counter

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < NBB; ++i) { mbar_init(&fullB[i], 32); mbar_init(&emptyB[i], G::EMPTYB_COUNT); }
    for (int i = 0; i < NBA; ++i) { mbar_init(&fullA[i], 32); mbar_init(&emptyA[i], 6); }
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
    s_misc[0] = n; s_misc[1] = task; s_misc[2] = 0;
  }
  __syncthreads();
  const int nseg = s_misc[0];

  if (warp >= NCW) {
    // =========================== producers ===========================
    const int pwarp = warp - NCW;
    int4* scratch = reinterpret_cast<int4*>(smem + G::OFF_SCRATCH) + pwarp * G::NPIX;
    const float inv_c = 1.0f / (float)P.C;
    for (int si = 0; si < nseg; ++si) {
      const Seg sg = segs[si];
      for (int l = 0; l < sg.nslab; ++l) {
        const int jg = sg.j0 + l;
        if (jg % NPW != pwarp) continue;
        const int m = sg.p0 - 2 + l;
        if (P.pf_dist > 0) prefetch_rows<G, T, TF>(c1, c2, flow, P, sg, 2 * (m + P.pf_dist) + 2, lane);
        if (sg.p0 + l < sg.p1) {
          const int ag = sg.a0 + l, slot = ag % G::NAP, prev = ag - G::NAP;
          if (prev >= 0) mbar_wait(&emptyA[prev % NBA], (prev / NBA) & 1);  // previous occupant fully consumed
          if (!(P.dbg & 2)) produce_pair<G, T>(c1, P, sg, sg.p0 + l, sA + (size_t)slot * G::ASLOT, lane, inv_c);
          mbar_arrive(&fullA[ag % NBA]);
        }
        const int slot = jg % G::NSLAB, prev = jg - G::NSLAB;
        if (prev >= 0) mbar_wait(&emptyB[prev % NBB], (prev / NBB) & 1);
        if (2 * m >= 0 && 2 * m - 1 < P.H && !(P.dbg & 2))  // slabs entirely outside the image are never read
          produce_slab<G, HAS_FLOW, T, TF>(c2, flow, P, sg, m, sB + (size_t)slot * G::BSLOT, scratch, lane);
        mbar_arrive(&fullB[jg % NBB]);
      }
    }
    return;
  }

  // =========================== compute warps ===========================
  // lane = lo + 2*xg + 2*NXG*hi.  Full task f of group g: combo c = f*NH + hi -> slab 4g + c/8, c1 row 2m-4 + c%8,
  // lo = warped row of the slab.  Edge task: hi = (slab ms, channel part); lo = half:
  //   half 0: c1 row 2m+4 vs warped row 2m (dy=-4) ; half 1: c1 row 2m-5 vs warped row 2m-1 (dy=+4).
  float* stage = sStage + warp * G::STAGE_WORDS;
  const int total_tasks = s_misc[1];
  const int lo = lane & 1, xg = (lane >> 1) % NXG, hi = lane / (2 * NXG);
  while (true) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&s_misc[2], 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= total_tasks) break;
    int si = 0;
    while (si + 1 < nseg && segs[si + 1].task0 <= t) ++si;
    const Seg sg = segs[si];
    const int tl = t - sg.task0, g = tl / G::TPG, rr = tl - G::TPG * g;
    const bool is_edge = rr == G::NF;

    int l, k, sub, bsel, kv0 = 0, kv1 = KV;
    bool arrive_b, arrive_a;
    if (!is_edge) {
      const int c = rr * G::NH + hi;
      l = GS * g + (c >> 3);
      const int r8 = c & 7;
      k = sg.p0 - 2 + l - 2 + (r8 >> 1);
      sub = r8 & 1;
      bsel = lo;
      arrive_b = lo == 0 && xg == 0 && (hi == 0 || r8 == 0);
      arrive_a = lo == 0 && xg == 0 && sub == 0;
    } else {
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
    const int yb = 2 * m - 1 + bsel, ya = 2 * k + sub;
    const bool store = pair_ok && ya < P.H;
    const bool active = store && yb >= 0 && yb < P.H;

    float acc[PX][ND];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int d = 0; d < ND; ++d) acc[p][d] = 0.f;
    if (active && !(P.dbg & 1)) {
      const float4* ap = sA + (size_t)sa * G::ASLOT + sub * G::A_ROWSLOTS + xg * (PX + 1);
      const float4* bp = sB + (size_t)sb * G::BSLOT + xg * (PX + 1) * 2 + bsel;
      correlate<G>(ap, bp, kv0, kv1, acc);
    }
    __syncwarp();
    // release the ring slots: one arrival per (task, slab) and per (task, slab, pair)
    if (slab_ok && arrive_b) mbar_arrive(&emptyB[jg % NBB]);
    if (pair_ok && arrive_a) mbar_arrive(&emptyA[ag % NBA]);
    if (G::KP > 1 && is_edge) {  // channel parts sit in adjacent hi values: butterfly over those lane bits
#pragma unroll
      for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          float v = acc[p][d];
          if (G::KP >= 2) v += __shfl_xor_sync(0xffffffffu, v, 2 * NXG);
          if (G::KP >= 4) v += __shfl_xor_sync(0xffffffffu, v, 4 * NXG);
          acc[p][d] = v;
        }
    }

    // epilogue: lane L owns, for each pixel p, the 9 channels [dyi*9, dyi*9+9) of pixel (ya, xb+p).
    // Runs go through a per-warp transposition buffer so that each store instruction writes whole runs.
    const int xb = sg.x0 + xg * PX;
    const int dyi = yb - ya + R;
    const bool writer = store && kv0 == 0;
    const long long my_off = writer ? (long long)((((size_t)sg.b * P.H + ya) * P.W + xb) * D + dyi * ND) : 0;
    const int my_nv = writer ? P.W - xb : 0;
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
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      float* stg = stage + (p & 1) * (G::STAGE_WORDS / 2);
#pragma unroll
      for (int d = 0; d < ND; ++d) stg[lane * 9 + d] = fmaxf(0.1f * acc[p][d], acc[p][d]);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 9; ++it)
        if (p < nv[it] && !(P.dbg & 4)) optr[it][p * D] = ElemTraits<T>::from_f(stg[it * 32 + lane]);
    }
    __syncwarp();
  }
}

// Host-side launcher.  Returns 0 when launched, 1 when the shape is not covered, -1 on a CUDA error.
template <int KV, int NXG, typename T, typename TF>
int launch_cfg(const T* c1, const T* c2, const TF* flow, float flow_scale, T* out, int B, int H, int W, int C,
               cudaStream_t st, int sms, int ctas_override, int pf_dist, int dbg) {
  using G = Geo<KV, NXG>;
  Params P;
  P.B = B; P.H = H; P.W = W; P.C = C; P.ops = D; P.flow_scale = flow_scale;
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
  auto go = [&](auto kern) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES) != cudaSuccess)
      return -1;
    kern<<<(unsigned)grid, NTHREADS, G::SMEM_BYTES, st>>>(c1, c2, flow, out, P);
    return 0;
  };
  return flow != nullptr ? go(fwd_stream_kernel<KV, NXG, true, T, TF>) : go(fwd_stream_kernel<KV, NXG, false, T, TF>);
}

// Covered: dense output (pixel stride 81), C in {16, 32, 64, 96, 128}.  Everything else returns 1 (caller falls back).
template <typename T, typename TF>
int launch(const T* c1, const T* c2, const TF* flow, float flow_scale, T* out, int B, int H, int W, int C,
           int64_t ops, cudaStream_t st, int sms, int ctas_override, int pf_dist, int dbg = 0) {
  if (ops != D) return 1;
#define PWC_STREAM_GO(KV, NXG) \
  return launch_cfg<KV, NXG, T, TF>(c1, c2, flow, flow_scale, out, B, H, W, C, st, sms, ctas_override, pf_dist, dbg)
  switch (C) {
    case 16: PWC_STREAM_GO(4, 4);
    case 32: PWC_STREAM_GO(8, 4);
    case 64: PWC_STREAM_GO(16, 2);
    case 96: PWC_STREAM_GO(24, 1);
    case 128: PWC_STREAM_GO(32, 1);
    default: return 1;
  }
#undef PWC_STREAM_GO
}

}  // namespace stream
}  // namespace pwc
