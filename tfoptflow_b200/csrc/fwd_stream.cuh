// Streaming fused warp + cost-volume forward (v2) for search_range 4 and C in {16, 32}.
//
// Persistent, warp-specialised kernel, one CTA per SM.  The image is cut into vertical strips of
// TW = 32 pixels; every CTA owns a contiguous range of (strip, row-pair) work and marches down it.
//   * 4 producer warps evaluate the bilinear warp of c2 (whole 128-byte pixel vectors per corner, so
//     every gather touches full cache lines) and stream the *warped* rows, plus the matching c1 rows,
//     into two shared-memory rings guarded by mbarriers.  Warped features never reach HBM.
//   * 4 compute warps pull tasks in order.  A full task F(m) takes one "slab" (warped rows 2m-1, 2m)
//     against the 4 row pairs that can see it; each lane owns a register tile of 2 rows x 8 pixels x
//     9 dx (144 accumulators) for one (pair, B row): row 2k uses dy = D, row 2k+1 uses dy = D-1 with
//     the SAME warped row, so every shared-memory operand feeds 18 FMAs per LDS.128.  Lanes of a warp
//     are arranged (pair, B row, x-group) so that c1 operands are shared by the two B rows and warped
//     operands by the four pairs (broadcast within the LDS wavefront).
//     The two displacements the 8-row window cannot reach (row 2m+4 with dy=-4, row 2m-5 with dy=+4)
//     are done by an edge task E(g) per group of 4 slabs (72 accumulators per lane).
//   * Results go through a per-warp transposition buffer so global stores are runs of 18 (9)
//     contiguous floats.
// Zero padding is applied after warping (core_costvol.py:27); the warp clamps (core_warp.py:101-104).
#pragma once
#include "common.cuh"

namespace pwc {
namespace stream {

#ifndef PWC_STREAM_LD4
#define PWC_STREAM_LD4 ld4
#endif
#define LD4 PWC_STREAM_LD4

constexpr int R = 4, ND = 9, D = 81;
constexpr int PX = 8, NXG = 4, TW = PX * NXG;  // 32-pixel strips
constexpr int NCW = 8;                          // compute warps
#ifndef PWC_STREAM_NPW
#define PWC_STREAM_NPW 4
#endif
constexpr int NPW = PWC_STREAM_NPW;             // producer warps
constexpr int REGS_COMPUTE = 168, REGS_PRODUCER = 88;  // setmaxnreg budgets (NPW == 8: 256*168 + 256*88 = 64K)
constexpr int NTHREADS = 32 * (NCW + NPW);
constexpr int NPROD = 32 * NPW;
constexpr int BU = TW + 2 * R;                  // 40 warped pixels per row
constexpr int B_USLOTS = BU + BU / PX - 1;      // 44 (slot(u) = u + u/8)
constexpr int A_ROWSLOTS = TW + TW / PX;        // 36
constexpr int MAXSEG = 8;
constexpr int TASKS_PER_GROUP = 9;               // Fa,Fb for 4 slabs + 1 edge task

template <int KV>
struct Geo {
  static constexpr int B_PLANE = B_USLOTS * 2 + 1;      // 89 float4, odd => kv-strided stores conflict-free
  static constexpr int A_PLANE = A_ROWSLOTS * 2 + 1;    // 73
  static constexpr int BSLOT = KV * B_PLANE;            // float4 per slab slot
  static constexpr int ASLOT = KV * A_PLANE;            // float4 per pair slot
#ifndef PWC_STREAM_NSLAB
#define PWC_STREAM_NSLAB 7
#define PWC_STREAM_NAP 12
#endif
  static constexpr int NSLAB = KV == 8 ? PWC_STREAM_NSLAB : 12;        // ring depths
  static constexpr int NAP = KV == 8 ? PWC_STREAM_NAP : 18;
  static constexpr int STAGE_WORDS = 640;               // per compute warp (2 x 320, double buffered)
  static constexpr size_t OFF_B = 0;
  static constexpr size_t OFF_A = OFF_B + (size_t)NSLAB * BSLOT * 16;
  static constexpr size_t OFF_STAGE = OFF_A + (size_t)NAP * ASLOT * 16;
  static constexpr size_t OFF_SCRATCH = OFF_STAGE + (size_t)NCW * STAGE_WORDS * 4;
  static constexpr size_t OFF_BAR = OFF_SCRATCH + (size_t)NPW * 2 * BU * 16;
  static constexpr size_t OFF_SEG = OFF_BAR + (size_t)(4 * NSLAB + 4 * NAP) * 8;
  static constexpr size_t SMEM_BYTES = OFF_SEG + MAXSEG * 32 + 64;
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
// Each producer warp owns every NPW-th slab of the CTA's sequence (and the c1 pair that travels with it),
// so NPW slabs are being gathered at any time and their global-memory latencies overlap.
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <int KV, typename T>
__device__ __forceinline__ void produce_pair(const T* __restrict__ c1, const Params& P, const Seg& sg, int k,
                                             float4* __restrict__ sA_slot, int lane, float inv_c) {
  // rows 2k, 2k+1 of the strip: 2 x 32 pixels x KV vectors, KV lanes per pixel
  constexpr int ITEMS = 2 * TW * KV / 32;  // per lane
  constexpr int BATCH = 8;
#pragma unroll
  for (int b0 = 0; b0 < ITEMS; b0 += BATCH) {
    float4 v[BATCH];
#pragma unroll
    for (int it = 0; it < BATCH; ++it) {
      const int id = (b0 + it) * 32 + lane;
      const int kv = id % KV, pxl = id / KV;
      const int row = pxl / TW, col = pxl - row * TW;
      const int y = 2 * k + row, x = sg.x0 + col;
      v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y < P.H && x < P.W) v[it] = LD4(c1 + (((size_t)sg.b * P.H + y) * P.W + x) * P.C + 4 * kv);
    }
#pragma unroll
    for (int it = 0; it < BATCH; ++it) {
      const int id = (b0 + it) * 32 + lane;
      const int kv = id % KV, pxl = id / KV;
      const int row = pxl / TW, col = pxl - row * TW;
      float4 w = v[it];
      w.x *= inv_c; w.y *= inv_c; w.z *= inv_c; w.w *= inv_c;  // mean over channels folded in (exact: C = 2^n)
      sA_slot[kv * Geo<KV>::A_PLANE + row * A_ROWSLOTS + col + col / PX] = w;
    }
  }
}

template <int KV, typename T, typename TF>
__device__ __forceinline__ void produce_slab(const T* __restrict__ c2, const TF* __restrict__ flow, const Params& P,
                                             const Seg& sg, int m, float4* __restrict__ sB_slot,
                                             int4* __restrict__ scratch, int lane) {
  // warped rows 2m-1 (bsel 0) and 2m (bsel 1), u = 0..39 <-> x = x0 - 4 + u : 80 pixels
  constexpr int NPIX = 2 * BU;
  __syncwarp();  // every lane is done reading the previous slab's scratch
  // phase 1: sampling coordinates of the 80 pixels -> per-warp scratch
#pragma unroll
  for (int e0 = 0; e0 < NPIX; e0 += 32) {
    const int e = e0 + lane;
    if (e < NPIX) {
      const int bsel = e / BU, u = e - bsel * BU;
      const int y = 2 * m - 1 + bsel, x = sg.x0 - R + u;
      int4 cd = make_int4(0, 0, 0, 0);
      if (y >= 0 && y < P.H && x >= 0 && x < P.W) {
        if (flow != nullptr) {
          const SampleCoord s = sample_coord<T, TF>(flow, P.flow_scale, sg.b, y, x, P.H, P.W);
          cd = make_int4(s.pix, __float_as_int(s.ax), __float_as_int(s.ay), 1);
        } else {
          cd = make_int4((sg.b * P.H + y) * P.W + x, 0, 0, 2);
        }
      }
      scratch[e] = cd;
    }
  }
  __syncwarp();
  // phase 2: KV lanes per pixel gather the 4 corners (full pixel vectors) and lerp
  constexpr int PPI = 32 / KV;        // pixels per iteration
  constexpr int NIT = NPW == 8 ? 2 : 5;  // iterations per batch (all loads of a batch are issued before any use)
  constexpr int PXB = NIT * PPI;      // pixels per batch
  static_assert(NPIX % PXB == 0, "batching");
  const int kv = lane % KV, sub = lane / KV;
#pragma unroll 1
  for (int p0 = 0; p0 < NPIX; p0 += PXB) {
    float4 tl[NIT], tr[NIT], bl[NIT], br[NIT];
    int4 cd[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      cd[it] = scratch[p0 + it * PPI + sub];
      tl[it] = tr[it] = bl[it] = br[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cd[it].w > 0) {
        const T* base = c2 + (size_t)cd[it].x * P.C + 4 * kv;
        tl[it] = LD4(base);
        if (cd[it].w == 1) {
          tr[it] = LD4(base + P.C);
          bl[it] = LD4(base + (size_t)P.W * P.C);
          br[it] = LD4(base + (size_t)P.W * P.C + P.C);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      float4 v = tl[it];
      if (cd[it].w == 1)
        v = round4<T>(bilerp4(tl[it], tr[it], bl[it], br[it], __int_as_float(cd[it].y), __int_as_float(cd[it].z)));
      const int e = p0 + it * PPI + sub;
      const int bsel = e / BU, u = e - bsel * BU;
      sB_slot[kv * Geo<KV>::B_PLANE + (u + u / PX) * 2 + bsel] = v;
    }
  }
}

// pull the leading edge of a future slab (2 rows of c2, flow and c1) towards L2 with bulk prefetches
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
template <typename T, typename TF>
__device__ __forceinline__ void prefetch_rows(const T* c1, const T* c2, const TF* flow, const Params& P, const Seg& sg,
                                              int y0, int lane) {
  // lanes 0,1: c2 rows y0, y0+1 ; lanes 2,3: c1 ; lanes 4,5: flow.  Ranges are rounded to 16 bytes.
  const int r = lane & 1, what = lane >> 1;
  const int y = y0 + r;
  if (what > 2 || y < 0 || y >= P.H) return;
  if (what == 2 && flow == nullptr) return;
  const int xlo = max(sg.x0 - 8, 0), xhi = min(sg.x0 + TW + 8, P.W);
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
  hi &= ~uintptr_t(15);  // round down: never touch bytes past the tensor
  if (hi > lo) bulk_prefetch_l2(reinterpret_cast<const void*>(lo), (uint32_t)(hi - lo));
}

// ---- per-lane correlation: 1 row x 8 pixels x 9 dx against ONE warped row --------------------
template <int KV>
__device__ __forceinline__ void correlate(const float4* __restrict__ ap, const float4* __restrict__ bp,
                                          float (&acc)[PX][ND]) {
#pragma unroll 1
  for (int kv = 0; kv < KV; ++kv) {
    // pixel-major order: the warped window slides (9 live float4), the c1 operand lives for one pixel only
    float4 bw[PX + 2 * R];
#pragma unroll
    for (int k = 0; k < ND; ++k) bw[k] = bp[kv * Geo<KV>::B_PLANE + (k + k / PX) * 2];
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      if (p > 0) bw[p + ND - 1] = bp[kv * Geo<KV>::B_PLANE + ((p + ND - 1) + (p + ND - 1) / PX) * 2];
      const float4 a = ap[kv * Geo<KV>::A_PLANE + p];
#pragma unroll
      for (int d = 0; d < ND; ++d) fma4(acc[p][d], a, bw[p + d]);
    }
  }
}

template <int KV, bool DENSE, typename T, typename TF>
__global__ void __launch_bounds__(NTHREADS, 1)
fwd_stream_kernel(const T* __restrict__ c1, const T* __restrict__ c2, const TF* __restrict__ flow,
                  T* __restrict__ out, const Params P) {
  using G = Geo<KV>;
  extern __shared__ __align__(16) unsigned char smem[];
  float4* sB = reinterpret_cast<float4*>(smem + G::OFF_B);
  float4* sA = reinterpret_cast<float4*>(smem + G::OFF_A);
  float* sStage = reinterpret_cast<float*>(smem + G::OFF_STAGE);
  uint64_t* fullB = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);
  // Two alternating barrier sets per data slot (index j % 2N, parity (j / 2N) & 1): with one set a waiter that
  // is two uses ahead of a slow producer would pass a parity wait prematurely (tests/stream_protocol_sim.py).
  constexpr int NBB = 2 * G::NSLAB, NBA = 2 * G::NAP;
  uint64_t* emptyB = fullB + NBB;
  uint64_t* fullA = emptyB + NBB;
  uint64_t* emptyA = fullA + NBA;
  Seg* segs = reinterpret_cast<Seg*>(smem + G::OFF_SEG);
  int* s_misc = reinterpret_cast<int*>(smem + G::OFF_SEG + MAXSEG * 32);  // [0] nseg, [1] total tasks, [2] task counter

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < NBB; ++i) { mbar_init(&fullB[i], 32); mbar_init(&emptyB[i], 3); }
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
      s.x0 = (int)(strip % P.nstrip_x) * TW;
      s.p0 = p0; s.p1 = p1; s.nslab = p1 - p0 + 5;
      s.task0 = task; s.j0 = j; s.a0 = a;
      task += TASKS_PER_GROUP * ((s.nslab + 3) / 4);
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
    if (NPW == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
    const int pwarp = warp - NCW;
    int4* scratch = reinterpret_cast<int4*>(smem + G::OFF_SCRATCH) + pwarp * (2 * BU);
    const float inv_c = 1.0f / (float)P.C;
    for (int si = 0; si < nseg; ++si) {
      const Seg sg = segs[si];
      for (int l = 0; l < sg.nslab; ++l) {
        const int jg = sg.j0 + l;
        if (jg % NPW != pwarp) continue;
        const int m = sg.p0 - 2 + l;
        // rows that the slab this warp will gather next (NPW slabs ahead) adds to the working set
        if (P.pf_dist > 0) prefetch_rows<T, TF>(c1, c2, flow, P, sg, 2 * (m + P.pf_dist) + 2, lane);
        if (sg.p0 + l < sg.p1) {
          const int ag = sg.a0 + l, slot = ag % G::NAP, prev = ag - G::NAP;
          if (prev >= 0) mbar_wait(&emptyA[prev % NBA], (prev / NBA) & 1);  // previous occupant fully consumed
          if (!(P.dbg & 2)) produce_pair<KV, T>(c1, P, sg, sg.p0 + l, sA + (size_t)slot * G::ASLOT, lane, inv_c);
          mbar_arrive(&fullA[ag % NBA]);
        }
        const int slot = jg % G::NSLAB, prev = jg - G::NSLAB;
        if (prev >= 0) mbar_wait(&emptyB[prev % NBB], (prev / NBB) & 1);
        if (2 * m >= 0 && 2 * m - 1 < P.H && !(P.dbg & 2))  // slabs entirely outside the image are never read
          produce_slab<KV, T, TF>(c2, flow, P, sg, m, sB + (size_t)slot * G::BSLOT, scratch, lane);
        mbar_arrive(&fullB[jg % NBB]);
      }
    }
    return;
  }

  // =========================== compute warps ===========================
  // Every task is 32 independent lane-tasks: (one c1 row, one warped row, 8 pixels, 9 dx) -> a run of 9
  // output channels for each of 8 pixels.  The three task kinds differ only in the lane -> work mapping:
  //   Fa(l): slab l vs c1 rows 2m-4 .. 2m-1      Fb(l): slab l vs c1 rows 2m .. 2m+3
  //          lane = bsel + 2*xg + 8*rq   (bsel: warped row 2m-1 / 2m;  rq: c1 row within the 4)
  //   E(g):  the two displacements out of reach of the 8-row window, for the 4 slabs of group g
  //          lane = half + 2*xg + 8*ms   (half 0: c1 row 2m+4 vs warped row 2m, dy=-4;
  //                                       half 1: c1 row 2m-5 vs warped row 2m-1, dy=+4)
  if (NPW == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_COMPUTE));
  float* stage = sStage + warp * G::STAGE_WORDS;
  const int total_tasks = s_misc[1];
  const int xg = (lane >> 1) & 3, lo = lane & 1, hi = lane >> 3;
  while (true) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&s_misc[2], 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= total_tasks) break;
    int si = 0;
    while (si + 1 < nseg && segs[si + 1].task0 <= t) ++si;
    const Seg sg = segs[si];
    const int tl = t - sg.task0, g = tl / TASKS_PER_GROUP, rr = tl - TASKS_PER_GROUP * g;

    int l, k, sub, bsel;  // slab (segment-local), c1 pair, row within pair, warped row within slab
    if (rr < 8) {
      l = 4 * g + (rr >> 1);
      k = sg.p0 - 2 + l - 2 + 2 * (rr & 1) + (hi >> 1);
      sub = hi & 1;
      bsel = lo;
    } else {
      l = 4 * g + hi;
      k = sg.p0 - 2 + l + (lo == 0 ? 2 : -3);
      sub = lo;
      bsel = 1 - lo;
    }
    if (rr < 8 && l >= sg.nslab) continue;  // null task of a ragged last group (warp-uniform)
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
      const float4* ap = sA + (size_t)sa * G::ASLOT + sub * A_ROWSLOTS + xg * (PX + 1);
      const float4* bp = sB + (size_t)sb * G::BSLOT + xg * (PX + 1) * 2 + bsel;
      correlate<KV>(ap, bp, acc);
    }
    __syncwarp();
    // release the ring slots: one arrival per (task, slab) and per (task, pair)
    if (rr < 8) {
      if (lane == 0) mbar_arrive(&emptyB[jg % NBB]);
      if (pair_ok && (lane & 15) == 0) mbar_arrive(&emptyA[ag % NBA]);
    } else {
      if (slab_ok && (lane & 7) == 0) mbar_arrive(&emptyB[jg % NBB]);
      if (pair_ok && (lane & 6) == 0) mbar_arrive(&emptyA[ag % NBA]);
    }

    // epilogue: lane L owns, for each pixel p, the 9 channels [dyi*9, dyi*9+9) of pixel (ya, xb+p).
    // Runs go through a per-warp transposition buffer so that each store instruction writes whole runs.
    const int xb = sg.x0 + xg * PX;
    const int dyi = yb - ya + R;
    const long long my_off = store ? (long long)((((size_t)sg.b * P.H + ya) * P.W + xb) * P.ops + dyi * ND) : 0;
    const int my_nv = store ? P.W - xb : 0;
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
    const size_t pstride = DENSE ? (size_t)D : (size_t)P.ops;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      float* stg = stage + (p & 1) * (G::STAGE_WORDS / 2);
#pragma unroll
      for (int d = 0; d < ND; ++d) stg[lane * 9 + d] = fmaxf(0.1f * acc[p][d], acc[p][d]);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 9; ++it)
        if (p < nv[it] && !(P.dbg & 4)) optr[it][p * pstride] = ElemTraits<T>::from_f(stg[it * 32 + lane]);
    }
    __syncwarp();
  }
}

// Host-side launcher.  Returns 0 when launched, 1 when the shape is not covered, -1 on a CUDA error.
template <typename T, typename TF>
int launch(const T* c1, const T* c2, const TF* flow, float flow_scale, T* out, int B, int H, int W, int C,
           int64_t ops, cudaStream_t st, int sms, int ctas_override, int pf_dist, int dbg = 0) {
  if (C != 32 && C != 16) return 1;
  Params P;
  P.B = B; P.H = H; P.W = W; P.C = C; P.ops = ops; P.flow_scale = flow_scale;
  P.NPAIR = (H + 1) / 2;
  P.nstrip_x = (W + TW - 1) / TW;
  P.Q = (long long)B * P.nstrip_x * P.NPAIR;
  P.pf_dist = pf_dist;
  P.dbg = dbg;
  if (P.Q <= 0) return 1;
  long long grid = ctas_override > 0 ? ctas_override : sms;
  if (grid > P.Q) grid = P.Q;
  // every CTA must fit its range into MAXSEG per-strip segments
  const long long per = (P.Q + grid - 1) / grid;
  if (per / P.NPAIR + 2 > MAXSEG) return 1;
  const bool dense = ops == D;
  auto go = [&](auto kern, size_t smem_bytes) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return -1;
    kern<<<(unsigned)grid, NTHREADS, smem_bytes, st>>>(c1, c2, flow, out, P);
    return 0;
  };
  int rc;
  if (C == 32) rc = dense ? go(fwd_stream_kernel<8, true, T, TF>, Geo<8>::SMEM_BYTES)
                          : go(fwd_stream_kernel<8, false, T, TF>, Geo<8>::SMEM_BYTES);
  else         rc = dense ? go(fwd_stream_kernel<4, true, T, TF>, Geo<4>::SMEM_BYTES)
                          : go(fwd_stream_kernel<4, false, T, TF>, Geo<4>::SMEM_BYTES);
  if (rc) return rc;
  return 0;
}

}  // namespace stream
}  // namespace pwc
