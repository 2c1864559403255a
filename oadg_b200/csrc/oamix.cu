// OA-Mix plan executor: the device side of OAMix.oamix (reference oa_mix.py:207-309).
//
// Data layout in HBM (all caller- or workspace-owned):
//   frames      u8 HWC, pitch 3*W, one per source image / branch ping-pong / bbo chain (S, T)
//   profiles    per gt box two float32 vectors ux[W], uy[H]; blurred mask(y,x) = uy[y]*ux[x]
//               (the reference materialises a 25 MB float HxWx3 mask per box, oa_mix.py:75-93)
//   hist / lut  per lane input 3x256 u32 histogram (+ luma sum), per LUT op 3x256 u8 table
//   plan        the host-sampled plan blob (oadg.h records) + work tables + tensor maps + ready ring, one H2D copy
//
// Two launches per call (a call may carry the views of SEVERAL loader batches: they are independent, and the more
// lanes the queue holds the fewer CTAs ever wait for a dependency):
//   oamix_chain_kernel   ONE persistent launch (four independent 256-thread CTAs per SM) that drains the host-built
//                        work queue (oamix_exec.h): mask profiles, union masks, histograms, LUTs, the bboxes-only
//                        chains level by level and every depth step of every (view, branch) lane.  Items are cut
//                        into tiles; an item is appended to the READY RING when the last tile of its last
//                        dependency is published, CTAs take tiles from the ring in order (no grid-wide barrier).
//                        The affine gathers (bboxes-only blends, bg-only steps) stage the source rectangle of every
//                        64 x 16 sub-tile with TMA (cp.async.bulk.tensor.2d, zero fill = BORDER_CONSTANT) into a
//                        two-stage mbarrier ring, so the staging of sub-tile k+1 overlaps the taps of sub-tile k.
//   mix_kernel           branch mixing + object-aware mixing of all views (oa_mix.py:236,281-309)
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "oadg_common.cuh"
#include "oadg_tma.cuh"
#include "oamix_exec.h"
#include "oamix_tile.h"

namespace oadg {
namespace {

// The phase handlers are separate device functions: compiled into one monolithic kernel body, ptxas (12.9) produced
// wrong code for the mixed-tile path (caught by tests/test_gpu_oamix.py::test_single_op_plans_match_host_arithmetic).
#ifdef OADG_INLINE_HANDLERS
#define OADG_HANDLER __forceinline__
#else
#define OADG_HANDLER __noinline__
#endif

constexpr int kCT = 256;     // WORKER threads per CTA of the chain kernel (warps 0-7): they run the tile handlers
constexpr int kSchedThreads = 32;            // warp 8: the tile scheduler (claims the next tile, publishes the last one)
constexpr int kCtaThreads = kCT + kSchedThreads;
#ifndef OADG_CTAS_PER_SM_MAX
#define OADG_CTAS_PER_SM_MAX 3
#endif
constexpr int kCtaPerSm = OADG_CTAS_PER_SM_MAX;   // independent CTAs per SM: tiles of different kinds overlap on an SM
// named barriers: 1 = the workers' block barrier; 2,3 = "slot k holds the next tile" (scheduler arrives, workers sync)
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_all(int id) { asm volatile("bar.sync %0, 288;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive_all(int id) { asm volatile("bar.arrive %0, 288;" ::"r"(id) : "memory"); }

// ---- dynamic shared memory of a CTA ------------------------------------------------------------------------
// [0, 2 * kStageBytes)   two gather stages: 3 TMA boxes of 16 rows x 256 B (frame) + 3 boxes of 16 rows x 96 B (mask)
//                        (aliased by the profile tiles' prefix sums, the histogram tiles' privatised bins and the
//                        hand-staged rows of frames without a tensor map)
// [kColOff, ...)         per tile column: cv2's adelta / bdelta terms of the affine map (int2)
// [kUxOff, ...)          per tile column: the x-profile of the blended box (float)
constexpr int kSubW = 64, kSubH = 16;
constexpr int kMaxSrcRows = 3 * kGatherBoxRows;                    // 48 source rows per sub-tile
constexpr int kMaxSrcPx = (kGatherImgBoxBytes - 15) / 3;           // 80 source pixels per row (TMA boxes start at 16-byte columns)
constexpr int kStageImgBytes = kMaxSrcRows * kGatherImgBoxBytes;   // 12288
constexpr int kStageMaskBytes = kMaxSrcRows * kGatherMaskBoxBytes; // 4608
constexpr int kStageBytes = kStageImgBytes + kStageMaskBytes;      // 16896 (a multiple of 128)
constexpr int kMaxTileW = 512;
constexpr int kColOff = 2 * kStageBytes;
constexpr int kUxOff = kColOff + kMaxTileW * 8;
constexpr int kDynSmem = kUxOff + kMaxTileW * 4;                   // 39936 B
constexpr int kHandStage = 2 * kStageBytes;                        // bytes the hand-staged fallback may use

struct RegOp {      // op parameters of one region, staged in shared memory for per-pixel tiles
  int32_t kind, p0, p1;
  float factor;
  double minv[6];
};

struct BboStage {    // one bbo job staged for the CTA (bbo_r_tile / bbo_c_segment)
  double minv[6];
  int32_t rect[4];
  int32_t W, H, gt, n_excl;
  const uint8_t* X;
  uint8_t* Y;
  int32_t x_map, pad;
  int32_t excl[16][4];   // supports of the next level's boxes (catch-up exclusion)
};

struct TileSlot {      // the scheduler warp hands the workers one tile per slot
  int item;            // -1: the queue is drained
  int tile;            // tile index within the item (after the item's permutation)
  Item I;
};
struct ChainSmem {
  TileSlot slot[2];
  unsigned done_seq;   // tiles the workers have finished (written by worker thread 0, read by the scheduler)
  int rowoff[2][96];   // hand-staged gathers: byte offset of every staged source row (frame rows, mask rows)
  int cand[16];        // mask tiles: the gt boxes whose support meets the tile
  int step_class;      // measurement aid: class of the last step tile (7 stream, 8 staged bg, 9 mixed / per pixel)
  int bs_key;          // which bbo job `bs` currently holds; -1 = none
  int sub_cls[kMaxTileW / kSubW];   // step tiles: class of every sub-tile (see step_tile)
  int gather_rect[2][4];            // per gather stage: source x0, y0, rows (-1: not staged, 0: outside the frame), -
  __align__(8) uint64_t full[2];    // mbarriers of the two gather stages
  ChainArgs args;      // the kernel arguments, copied once: the (non-inlined) handlers read them from shared memory
  BboStage bs;
  __align__(16) uint8_t lut[OADG_MAX_REGIONS * 768];
  RegOp rop[OADG_MAX_REGIONS];
  Lane lane;
  double red[32];
  double ksum;
  double div255[256];  // i / 255 in float64 (the reference divides a uint8 array by the python int 255, bbox_augmentation.py:267)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}

// ---- the work queue ---------------------------------------------------------------------------------------
// Every item has a counter of outstanding dependency tiles (`pending`, initialised by the host).  The CTA that
// publishes the last outstanding dependency tile of an item makes the item READY: it reserves a range of TICKETS for
// the item's tiles (one atomic on the publish cursor) and writes them (warp 0, one ticket per lane and round).  A CTA
// that wants work takes the next ticket number (one atomic on the claim cursor) and reads tickets[T] with an acquire
// load, waiting if that ticket is not published yet: tiles are handed out in the order in which their items became
// ready, with two memory round trips per claim and no per-item hot spot.  Publishing a tile is bar.sync + fence by
// thread 0, then one atomic per successor; the CTA bar.syncs before it touches a claimed tile's data.  All CTAs are
// co-resident (cooperative launch) and a waiting CTA holds no unfinished tile, so the scheme cannot deadlock; a CTA
// that waits longer than 2 s raises the sticky fault flag (the host turns it into OADG_E_PLAN) instead of hanging
// the GPU.
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ unsigned ld_acquire_cta_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// scheduler warp (all 32 lanes), after the workers reported the tile finished (done_seq): the tile's results become
// visible device-wide, and every successor whose last outstanding dependency tile this was gets its tickets.  One
// lane per successor: the successor list, the dependency counters and the ticket reservations each cost ONE round
// trip for the whole list.
__device__ __forceinline__ void publish_tile(const ChainArgs& A, const Item& I) {
  const int lane = threadIdx.x & 31;
  fence_acq_rel_gpu();        // release: the tile's stores (ordered before this warp by the barrier) before the counters
  tma::fence_proxy_async();   // ... and before TMA reads (async proxy) of later items
  if (A.debug & 64) return;   // fault injection (tests): no successor is ever released, the queue starves
  for (int base = 0; base < I.succ_count; base += 32) {
    int s = -1, nt = 0;
    unsigned long long first = 0;
    if (base + lane < I.succ_count) {
      s = A.succ[I.succ_first + base + lane];
      if (atomicSub(A.pending + s, 1) == 1) {   // the last outstanding dependency tile: the successor is ready
        nt = A.items[s].ntiles;
        fence_acq_rel_gpu();                    // acquire the other dependencies' results before handing out the item
        first = atomicAdd(A.ring, (unsigned long long)nt);
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, nt > 0);
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      const int sj = __shfl_sync(0xffffffffu, s, j), ntj = __shfl_sync(0xffffffffu, nt, j);
      const unsigned long long fj = __shfl_sync(0xffffffffu, first, j);
      for (int t = lane; t < ntj; t += 32)
        st_release_u64(A.tickets + fj + t, ((unsigned long long)(unsigned)(sj + 1) << 32) | (unsigned)t);
    }
  }
}

// ------------------------------------------------------------------------------------
// blurred-mask profile of one (gt box, axis) (oa_mix.py:78-91): indicator on the 1/sr canvas -> GaussianBlur
// (separable, BORDER_REFLECT_101, float32 kernel from getGaussianKernel) -> bilinear cv2.resize to full resolution.
// ------------------------------------------------------------------------------------
__device__ OADG_HANDLER void profile_tile(const ChainArgs& A, ChainSmem& S, uint8_t* dyn, int obj) {
  const DevPlan& P = A.P;
  const int sr = 4;
  const int g = obj >> 1, axis = obj & 1;
  const oadg_gt_t G = P.gts[g];
  const oadg_view_t& V = P.views[G.view];
  const int n_hi = axis == 0 ? V.W : V.H;
  const int n_lo = n_hi / sr;
  const int lo = G.lo[axis], hi = G.lo[axis + 2];
  const int ks = axis == 0 ? G.kx : G.ky;
  const double sigma = axis == 0 ? G.sigma_x : G.sigma_y;
  double* K = reinterpret_cast<double*>(dyn);                  // [1537] exclusive prefix sums of the gaussian kernel
  float* p = reinterpret_cast<float*>(dyn + 1538 * 8);         // [1024] low-res blurred profile
  float* out = (axis == 0 ? A.prof_x + (size_t)g * P.max_w : A.prof_y + (size_t)g * P.max_h);
  const int tid = threadIdx.x;
  worker_sync();  // the shared buffers may still be in use by the previous tile
  if (n_lo <= 0) {
    for (int d = tid; d < n_hi; d += kCT) out[d] = 0.f;
    return;
  }
  if (G.blur) {
    // cv::getGaussianKernel(ks, sigma, CV_32F): exp(-x^2/(2 sigma^2)) in double, normalised, cast to float32
    const double s2 = -0.5 / (sigma * sigma);
    double part = 0.0;
    for (int i = tid; i < ks; i += kCT) {
      double x = i - (ks - 1) * 0.5;
      part += exp(s2 * x * x);
    }
    part = warp_sum(part);
    if ((tid & 31) == 0) S.red[tid >> 5] = part;
    worker_sync();
    if (tid == 0) {
      double t = 0;
      for (int w = 0; w < kCT / 32; ++w) t += S.red[w];
      S.ksum = 1.0 / t;
    }
    worker_sync();
    const double ksum = S.ksum;
    // K[j] = sum of the float32 taps 0..j-1 in float64 (the blur of an indicator is a difference of prefix sums)
    for (int i = tid; i <= ks; i += kCT) {
      double x = (i - 1) - (ks - 1) * 0.5;
      K[i] = i == 0 ? 0.0 : (double)(float)(exp(s2 * x * x) * ksum);
    }
    worker_sync();
    for (int off = 1; off <= ks; off <<= 1) {  // Hillis-Steele inclusive scan over K[0..ks]
      double v[8];
      int n = 0;
      for (int i = tid; i <= ks; i += kCT, ++n) v[n] = i >= off ? K[i] + K[i - off] : K[i];
      worker_sync();
      n = 0;
      for (int i = tid; i <= ks; i += kCT, ++n) K[i] = v[n];
      worker_sync();
    }
    const int r = ks / 2;
    auto range = [&](int a, int b) {  // sum of taps j in [a, b) clipped to [0, ks)
      a = max(a, 0);
      b = min(b, ks);
      return b > a ? K[b] - K[a] : 0.0;
    };
    auto span = [&](int a, int b) {   // how many taps that is
      a = max(a, 0);
      b = min(b, ks);
      return b > a ? b - a : 0;
    };
    // A window that lies entirely inside the box (only boxes that span the frame get there, through the reflected
    // border) carries cv2's saturated value, whose last bit decides every truncation of the blend: the x profile
    // carries 1.0 and the y profile the constant of cv_saturated_col / _row (oamix_math.h); kSatMark is replaced below.
    constexpr float kSatMark = 2.0f;
    if (tid == 0) S.red[0] = 0.0;
    worker_sync();
    bool any_sat = false;
    if (r <= n_lo - 1) {
      // BORDER_REFLECT_101 with at most one reflection per side: tap j reads q = x + j - r, -q or 2(n_lo-1) - q
      for (int x = tid; x < n_lo; x += kCT) {
        const int s = r - x;  // j = q + s
        const int m2 = 2 * (n_lo - 1);
        double acc = range(lo + s, hi + s);                                    // q in [lo, hi)
        acc += range(-hi + 1 + s, min(-lo, -1) + 1 + s);                       // q in [-hi+1, min(-lo,-1)]
        acc += range(max(m2 - hi + 1, n_lo) + s, m2 - lo + 1 + s);             // q in [max(m2-hi+1,n_lo), m2-lo]
        const int cov = span(lo + s, hi + s) + span(-hi + 1 + s, min(-lo, -1) + 1 + s) +
                        span(max(m2 - hi + 1, n_lo) + s, m2 - lo + 1 + s);
        p[x] = (float)acc;
        if (cov == ks) {
          p[x] = axis == 0 ? 1.f : kSatMark;
          any_sat = true;
        }
      }
    } else {
      const int period = 2 * (n_lo - 1);
      for (int x = tid; x < n_lo; x += kCT) {
        double acc = 0.0;
        int cov = 0;
        for (int j = 0; j < ks; ++j) {
          int q = x + j - r;
          if (n_lo == 1) q = 0;
          else {
            if (q < 0) q = -q;
            q %= period;
            if (q >= n_lo) q = period - q;
          }
          if (q >= lo && q < hi) {
            acc += K[j + 1] - K[j];
            ++cov;
          }
        }
        p[x] = (float)acc;
        if (cov == ks) {
          p[x] = axis == 0 ? 1.f : kSatMark;
          any_sat = true;
        }
      }
    }
    if (axis == 1) {
      if (any_sat) S.red[0] = 1.0;   // same value from every writer
      worker_sync();
      if (S.red[0] != 0.0) {         // rare: both kernels as float32 taps, then the two float sums in cv2's order
        float* kyf = reinterpret_cast<float*>(dyn + 1538 * 8 + 4096);
        float* kxf = kyf + 1540;
        const int ksx = G.kx;
        const double s2x = -0.5 / (G.sigma_x * G.sigma_x);
        worker_sync();
        double px = 0.0;
        for (int i = tid; i < ksx; i += kCT) {
          double x = i - (ksx - 1) * 0.5;
          px += exp(s2x * x * x);
        }
        px = warp_sum(px);
        if ((tid & 31) == 0) S.red[tid >> 5] = px;
        worker_sync();
        double tx = 0;
        for (int w = 0; w < kCT / 32; ++w) tx += S.red[w];
        const double ksumx = 1.0 / tx;
        for (int i = tid; i < ksx; i += kCT) {
          double x = i - (ksx - 1) * 0.5;
          kxf[i] = (float)(exp(s2x * x * x) * ksumx);
        }
        for (int i = tid; i < ks; i += kCT) {
          double x = i - (ks - 1) * 0.5;
          kyf[i] = (float)(exp(s2 * x * x) * ksum);
        }
        worker_sync();
        if (tid == 0) S.ksum = (double)cv_saturated_col(kyf, ks, cv_saturated_row(kxf, ksx));
        worker_sync();
        const float sat = (float)S.ksum;
        for (int x = tid; x < n_lo; x += kCT)
          if (p[x] == kSatMark) p[x] = sat;
      }
    }
  } else {
    for (int x = tid; x < n_lo; x += kCT) p[x] = (x >= lo && x < hi) ? 1.f : 0.f;
  }
  worker_sync();
  // cv2.resize(f32, INTER_LINEAR): fx = (float)((dx+0.5)*scale - 0.5)
  const double scale = (double)n_lo / (double)n_hi;
  for (int d = tid; d < n_hi; d += kCT) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    float t = fsub(f, (float)s);
    if (s < 0) { s = 0; t = 0.f; }
    if (s >= n_lo - 1) { s = n_lo - 1; t = 0.f; }
    int s1 = min(s + 1, n_lo - 1);
    // OpenCV's float32 bilinear resize evaluates a + t * (b - a) with one fused multiply-add (checked bit for bit
    // against cv2.resize on 1.1 M values, horizontal and vertical pass alike); a constant run stays constant
    out[d] = ffma(t, fsub(p[s1], p[s]), p[s]);
  }
}
// union of the blurred gt masks of a view (np.max(mask_bboxes, axis=0), bbox_augmentation.py:260) as float32 and as
// uint8(mask*255): written once per batch, read by every bg-only op.  Tile = 256 x 8 px, one column x 8 rows per
// thread; the boxes whose support meets the tile are listed once per tile, and a thread keeps the x-profile value of
// each listed box for its column in registers.
__device__ OADG_HANDLER void mask_tile(const ChainArgs& A, ChainSmem& S, uint8_t* dyn, int view, int local, int tx) {
  const DevPlan& P = A.P;
  const oadg_view_t& V = P.views[view];
  const int x0 = (local % tx) * kMaskTileW, y0 = (local / tx) * kMaskTileH;
  const int x1 = min(x0 + kMaskTileW, V.W), y1 = min(y0 + kMaskTileH, V.H);
  worker_sync();
  if (threadIdx.x < 32) {  // lane k tests box k (+32, ...): one round trip instead of a serial walk
    if (threadIdx.x == 0) S.bs_key = -1;   // bs.excl is reused below
    int n = 0;
    for (int k0 = 0; k0 < V.n_gt; k0 += 32) {
      const int k = k0 + threadIdx.x;
      int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      bool hit = false;
      if (k < V.n_gt) {
        const int32_t* s = P.gts[V.gt_first + k].supp;
        s0 = s[0]; s1 = s[1]; s2 = s[2]; s3 = s[3];
        hit = s0 < x1 && s2 > x0 && s1 < y1 && s3 > y0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      const int slot = n + __popc(bal & ((1u << threadIdx.x) - 1u));
      if (hit && slot < 16) {
        S.bs.excl[slot][0] = s0; S.bs.excl[slot][1] = s1; S.bs.excl[slot][2] = s2; S.bs.excl[slot][3] = s3;
        S.cand[slot] = V.gt_first + k;
      }
      n += __popc(bal);
    }
    if (threadIdx.x == 0) S.bs.n_excl = n;
  }
  worker_sync();
  const int n = S.bs.n_excl;
  const size_t base = (size_t)view * P.mask_stride;
  if (n <= 16 && (V.W & 3) == 0 && (P.mask_stride & 3) == 0 && x1 - x0 == kMaskTileW) {
    // full-width tile of a frame whose rows are 16-byte periodic: the listed boxes' profile slices go to shared
    // memory (zeros outside a box's support: the product is then +0 and never wins the max), every thread owns 4
    // consecutive columns of 8 rows and stores one float4 + one uint32 per row
    const int t = threadIdx.x, g = t & 63, r0 = t >> 6;
    float* spx = reinterpret_cast<float*>(dyn);          // [n][256]
    float* spy = spx + n * kMaskTileW;                    // [n][32]
    for (int i = t; i < n * kMaskTileW; i += kCT) {
      const int c = i >> 8, x = x0 + (i & 255);
      spx[i] = (x >= S.bs.excl[c][0] && x < S.bs.excl[c][2]) ? A.prof_x[(size_t)S.cand[c] * P.max_w + x] : 0.f;
    }
    for (int i = t; i < n * kMaskTileH; i += kCT) {
      const int c = i / kMaskTileH, y = y0 + i % kMaskTileH;
      spy[i] = (y < y1 && y >= S.bs.excl[c][1] && y < S.bs.excl[c][3]) ? A.prof_y[(size_t)S.cand[c] * P.max_h + y] : 0.f;
    }
    worker_sync();
    for (int r = r0; r < y1 - y0; r += 4) {
      float m[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < n; ++c) {
        const float uy = spy[c * kMaskTileH + r];
        const float4 ux = *reinterpret_cast<const float4*>(spx + c * kMaskTileW + 4 * g);
        const float v0 = fmul(uy, ux.x), v1 = fmul(uy, ux.y), v2 = fmul(uy, ux.z), v3 = fmul(uy, ux.w);
        m[0] = v0 > m[0] ? v0 : m[0]; m[1] = v1 > m[1] ? v1 : m[1];
        m[2] = v2 > m[2] ? v2 : m[2]; m[3] = v3 > m[3] ? v3 : m[3];
      }
      const size_t o = base + (size_t)(y0 + r) * V.W + x0 + 4 * g;
      *reinterpret_cast<float4*>(A.maskf + o) = make_float4(m[0], m[1], m[2], m[3]);
      // uint8(mask * 255): truncation of a non-negative float32 (mask_to_u8, oamix_math.h) without the F2I instruction
      *reinterpret_cast<uint32_t*>(A.masku + o) =
          (uint32_t)f32_trunc_u8(fmul(m[0], 255.0f)) | (uint32_t)f32_trunc_u8(fmul(m[1], 255.0f)) << 8 |
          (uint32_t)f32_trunc_u8(fmul(m[2], 255.0f)) << 16 | (uint32_t)f32_trunc_u8(fmul(m[3], 255.0f)) << 24;
    }
    return;
  }
  const int x = x0 + (threadIdx.x & 255);
  if (x >= x1) return;
  if (n > 8) {  // many overlapping boxes: the plain per-pixel walk
    for (int y = y0; y < y1; ++y) mask_pixel(P, view, x, y, A.maskf, A.masku);
    return;
  }
  float ux[8];
#pragma unroll
  for (int c = 0; c < 8; ++c)
    ux[c] = (c < n && x >= S.bs.excl[c][0] && x < S.bs.excl[c][2]) ? A.prof_x[(size_t)S.cand[c] * P.max_w + x] : -1.f;
  for (int y = y0; y < y1; ++y) {
    float m = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < n && ux[c] >= 0.f && y >= S.bs.excl[c][1] && y < S.bs.excl[c][3]) {
        const float v = fmul(A.prof_y[(size_t)S.cand[c] * P.max_h + y], ux[c]);
        m = v > m ? v : m;
      }
    }
    const size_t o = base + (size_t)y * V.W + x;
    A.maskf[o] = m;
    A.masku[o] = (uint8_t)mask_to_u8(m);
  }
}
// per-channel histogram + luma sum of a lane's input frame (PIL Image.histogram()); tile = 32768 px (linear)
__device__ OADG_HANDLER void hist_tile(const Lane& L, unsigned* hist, int local, unsigned long long& lsum) {
  const int tid = threadIdx.x;
  const size_t npx = (size_t)L.H * L.W;
  const size_t p0 = (size_t)local * kHistTilePx;
  const size_t p1 = p0 + kHistTilePx < npx ? p0 + kHistTilePx : npx;
  unsigned* my = hist + ((tid >> 5) & 1) * 768;
  if ((((uintptr_t)L.in) & 15) == 0) {
    // 16 px = 48 B = 3 x uint4 per iteration: the channel of byte k is k % 3 at a compile-time phase
    const size_t c1 = p1 / kChunkPx;
    for (size_t i = p0 / kChunkPx + tid; i < c1; i += kCT) {
      Chunk c;
      chunk_load(L.in + i * 48, kChunkPx, true, c);
#pragma unroll
      for (int px = 0; px < kChunkPx; ++px) {
        const int c0 = chunk_get(c, px * 3), c1v = chunk_get(c, px * 3 + 1), c2 = chunk_get(c, px * 3 + 2);
        atomicAdd(&my[c0], 1u);
        atomicAdd(&my[256 + c1v], 1u);
        atomicAdd(&my[512 + c2], 1u);
        lsum += (unsigned)pil_luma(c0, c1v, c2);
      }
    }
    for (size_t i = c1 * kChunkPx + tid; i < p1; i += kCT) {  // ragged tail of the frame (last tile only)
      const uint8_t* p = L.in + i * 3;
      int c0 = ldb(p), c1v = ldb(p + 1), c2 = ldb(p + 2);
      atomicAdd(&my[c0], 1u);
      atomicAdd(&my[256 + c1v], 1u);
      atomicAdd(&my[512 + c2], 1u);
      lsum += (unsigned)pil_luma(c0, c1v, c2);
    }
  } else {
    for (size_t i = p0 + tid; i < p1; i += kCT) {
      const uint8_t* p = L.in + i * 3;
      int c0 = ldb(p), c1v = ldb(p + 1), c2 = ldb(p + 2);
      atomicAdd(&my[c0], 1u);
      atomicAdd(&my[256 + c1v], 1u);
      atomicAdd(&my[512 + c2], 1u);
      lsum += (unsigned)pil_luma(c0, c1v, c2);
    }
  }
}
__device__ OADG_HANDLER void hist_begin(unsigned* hist) {
  worker_sync();
  for (int i = threadIdx.x; i < 2 * 768; i += kCT) hist[i] = 0;
  worker_sync();
}
__device__ OADG_HANDLER void hist_flush(const ChainArgs& A, const unsigned* hist, int slot, unsigned long long& lsum) {
  worker_sync();
  unsigned* dst = A.hist + (size_t)slot * 768;
  for (int i = threadIdx.x; i < 768; i += kCT) {
    unsigned s = 0;
#pragma unroll
    for (int w = 0; w < 2; ++w) s += hist[w * 768 + i];
    if (s) atomicAdd(dst + i, s);
  }
  lsum = warp_sum(lsum);
  if ((threadIdx.x & 31) == 0 && lsum) atomicAdd(A.luma + slot, lsum);
  lsum = 0;
  worker_sync();
}
// one LUT op: PIL.ImageOps autocontrast / equalize from the finished histogram, or a closed-form table
__device__ OADG_HANDLER void lut_tile(const ChainArgs& A, ChainSmem& S, int job) {
  const LutJob J = A.lutjobs[job];
  const oadg_op_t& op = A.P.ops[J.op];
  uint8_t* out = A.luts + (size_t)op.lut * 768;
  const int tid = threadIdx.x;
  worker_sync();
  if (op.kind == OADG_OP_AUTOCONTRAST || op.kind == OADG_OP_EQUALIZE) {
    // one thread per bin, channel after channel (same integer / float64 arithmetic as lut_autocontrast_ch /
    // lut_equalize_ch in oamix_math.h, with the scans done by ballots and a block prefix sum)
    unsigned* msk = reinterpret_cast<unsigned*>(S.red);        // [8] non-empty-bin masks of the 8 warps
    unsigned* wsum = reinterpret_cast<unsigned*>(S.red) + 8;   // [8] per-warp histogram sums
    const int lane = tid & 31, warp = tid >> 5;
    for (int c = 0; c < 3; ++c) {
      const unsigned hv = A.hist[(size_t)J.hist_slot * 768 + c * 256 + tid];
      const unsigned bal = __ballot_sync(0xffffffffu, hv != 0u);
      unsigned incl = hv;   // inclusive prefix sum inside the warp
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      worker_sync();
      if (lane == 31) wsum[warp] = incl;
      if (lane == 0) msk[warp] = bal;
      worker_sync();
      int lo = 256, hi = -1, nnz = 0;
      unsigned total = 0, before = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const unsigned m = msk[w];
        if (m) {
          if (lo == 256) lo = w * 32 + __ffs(m) - 1;
          hi = w * 32 + 31 - __clz(m);
        }
        nnz += __popc(m);
        total += wsum[w];
        if (w < warp) before += wsum[w];
      }
      const unsigned excl = before + incl - hv;   // sum of the bins below this one
      uint8_t r = (uint8_t)tid;
      if (op.kind == OADG_OP_AUTOCONTRAST) {
        if (hi > lo) {
          const double scale = 255.0 / (double)(hi - lo);
          const double offset = dmul((double)(-lo), scale);
          const int v = (int)dadd(dmul((double)tid, scale), offset);
          r = (uint8_t)imin(imax(v, 0), 255);
        }
      } else {
        const unsigned last = hi >= 0 ? A.hist[(size_t)J.hist_slot * 768 + c * 256 + hi] : 0u;
        const unsigned step = nnz <= 1 ? 0u : (total - last) / 255u;
        if (step) {
          const unsigned v = (step / 2 + excl) / step;
          r = (uint8_t)(v > 255u ? 255u : v);   // Image.point clips list entries to 8 bits
        }
      }
      out[c * 256 + tid] = r;
    }
    return;
  }
  if (tid < 256) {
    const oadg_view_t& V = A.P.views[J.view];
    const double lsum = J.hist_slot >= 0 ? (double)A.luma[J.hist_slot] : 0.0;
    const uint8_t v = lut_simple_at(op, tid, lsum, (double)((long long)V.H * V.W));
    out[tid] = v;
    out[256 + tid] = v;
    out[512 + tid] = v;
  }
}

// T (and S when the chain has a second level) = copy of the lane input; tiles [l0, l1) of 64 KB
__device__ OADG_HANDLER void copy_segment(const Chain& C, size_t nbytes, bool both, int l0, int l1) {
  const size_t b0 = (size_t)l0 * kCopyTileBytes;
  const size_t b1 = (size_t)l1 * kCopyTileBytes < nbytes ? (size_t)l1 * kCopyTileBytes : nbytes;
  const int tid = threadIdx.x;
  if (((((uintptr_t)C.in) | ((uintptr_t)C.S) | ((uintptr_t)C.T)) & 15) == 0) {
    const uint4* s = reinterpret_cast<const uint4*>(C.in);
    uint4* d = reinterpret_cast<uint4*>(C.T);
    uint4* d2 = reinterpret_cast<uint4*>(C.S);
    const size_t v1 = b1 / 16;
#pragma unroll 4
    for (size_t i = b0 / 16 + tid; i < v1; i += kCT) {
      const uint4 v = s[i];
      d[i] = v;
      if (both) d2[i] = v;
    }
    for (size_t i = v1 * 16 + tid; i < b1; i += kCT) {
      C.T[i] = C.in[i];
      if (both) C.S[i] = C.in[i];
    }
  } else {
    for (size_t i = b0 + tid; i < b1; i += kCT) {
      C.T[i] = C.in[i];
      if (both) C.S[i] = C.in[i];
    }
  }
}
// ---- staged affine gathers ---------------------------------------------------------------------------------
// Both geometric op families (bboxes-only blends and bg-only ops) resample a frame through cv::warpAffine's
// fixed-point bilinear map.  A CTA works on sub-tiles of 64 x 16 output pixels: the source rectangle the sub-tile
// reads (exact: the map is monotone in x and in y, so its extremes sit at the sub-tile corners) is copied into
// shared memory with 16-byte vector loads of whole row spans, then every thread resamples 4 consecutive pixels
// from shared memory and writes 12 bytes.  Out-of-frame taps read 0 (BORDER_CONSTANT).

struct StageView {
  const uint8_t* sm;   // staged rows: row r holds the 16-byte aligned global span that covers source row by0 + r
  int pitch;           // bytes per staged row (multiple of 16)
  int bx0, by0, bx1, by1;
  uint32_t lo;         // low bits of the frame pointer: byte phase of a row = (lo + (y*W + bx0)*C) & 15
  int W, H, C;
};

__device__ __forceinline__ void warp_coord(const double* m, int x, int y, int& sx, int& sy, int& fx, int& fy) {
  const int X = (cv_round(dmul(dadd(dmul(m[1], (double)y), m[2]), 1024.0)) + 16 + cv_round(dmul(dmul(m[0], (double)x), 1024.0))) >> 5;
  const int Y = (cv_round(dmul(dadd(dmul(m[4], (double)y), m[5]), 1024.0)) + 16 + cv_round(dmul(dmul(m[3], (double)x), 1024.0))) >> 5;
  sx = imin(imax(X >> 5, -32768), 32767);
  sy = imin(imax(Y >> 5, -32768), 32767);
  fx = X & 31;
  fy = Y & 31;
}
// the same map with the row terms (functions of y alone) computed once per thread
struct WarpRowTerm {
  int X0, Y0;
};
__device__ __forceinline__ WarpRowTerm warp_row_term(const double* m, int y) {
  WarpRowTerm r;
  r.X0 = cv_round(dmul(dadd(dmul(m[1], (double)y), m[2]), 1024.0)) + 16;
  r.Y0 = cv_round(dmul(dadd(dmul(m[4], (double)y), m[5]), 1024.0)) + 16;
  return r;
}
__device__ __forceinline__ void warp_coord_row(const double* m, WarpRowTerm r, int x, int& sx, int& sy, int& fx, int& fy) {
  const int X = (r.X0 + cv_round(dmul(dmul(m[0], (double)x), 1024.0))) >> 5;
  const int Y = (r.Y0 + cv_round(dmul(dmul(m[3], (double)x), 1024.0))) >> 5;
  sx = imin(imax(X >> 5, -32768), 32767);
  sy = imin(imax(Y >> 5, -32768), 32767);
  fx = X & 31;
  fy = Y & 31;
}
// source rectangle of the output rectangle [x0,x1) x [y0,y1), clamped to the frame; false when it is empty
__device__ __forceinline__ bool warp_src_rect(const double* m, int x0, int y0, int x1, int y1, int W, int H, int r[4]) {
  int sxa = 1 << 30, sxb = -(1 << 30), sya = 1 << 30, syb = -(1 << 30);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int sx, sy, fx, fy;
    warp_coord(m, (k & 1) ? x1 - 1 : x0, (k & 2) ? y1 - 1 : y0, sx, sy, fx, fy);
    sxa = imin(sxa, sx); sxb = imax(sxb, sx);
    sya = imin(sya, sy); syb = imax(syb, sy);
  }
  r[0] = imax(sxa, 0);
  r[1] = imax(sya, 0);
  r[2] = imin(sxb + 2, W);
  r[3] = imin(syb + 2, H);
  return r[2] > r[0] && r[3] > r[1];
}
__device__ __forceinline__ int stage_pitch(int bx0, int bx1, int C) { return (15 + (bx1 - bx0) * C + 15) & ~15; }
// copy source rows [by0,by1) x [bx0,bx1) of a C-byte-per-pixel frame into shared memory (all threads)
__device__ OADG_HANDLER void stage_rows(uint8_t* sm, int pitch, const uint8_t* base, int W, int H, int C, int bx0, int by0,
                                        int rows) {
  const int vpr = pitch >> 4;
  const uint8_t* end = base + (size_t)H * W * C;
  for (int i = threadIdx.x; i < rows * vpr; i += kCT) {
    const int rr = i / vpr, vv = i - rr * vpr;
    const uint8_t* g = base + ((size_t)(by0 + rr) * W + bx0) * C;
    const uint8_t* ga = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(g) & ~(uintptr_t)15) + vv * 16;
    uint4 val;
    if (ga >= base && ga + 16 <= end) {
      val = *reinterpret_cast<const uint4*>(ga);
    } else {  // the vector straddles the frame's first / last bytes
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      for (int k = 0; k < 16; ++k)
        if (ga + k >= base && ga + k < end) w[k >> 2] |= (uint32_t)ga[k] << ((k & 3) * 8);
      val = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(sm + (size_t)rr * pitch + vv * 16) = val;
  }
}
__device__ __forceinline__ StageView make_view(uint8_t* sm, const uint8_t* base, int W, int H, int C, const int r[4]) {
  StageView v;
  v.sm = sm;
  v.bx0 = r[0]; v.by0 = r[1]; v.bx1 = r[2]; v.by1 = r[3];
  v.W = W; v.H = H; v.C = C;
  v.lo = (uint32_t)(uintptr_t)base;
  v.pitch = stage_pitch(r[0], r[2], C);
  return v;
}
// first byte of source pixel (sx, sy) in the staged rows (the pixel must lie inside the staged rectangle)
__device__ __forceinline__ const uint8_t* staged_px(const StageView& v, int sx, int sy) {
  const uint32_t phase = (v.lo + (uint32_t)((sy * v.W + v.bx0) * v.C)) & 15u;
  return v.sm + (sy - v.by0) * v.pitch + phase + (sx - v.bx0) * v.C;
}
// one warped u8x3 pixel from staged rows (same taps and weights as warp_fetch3)
__device__ __forceinline__ void staged_fetch3(const StageView& v, int sx, int sy, int fx, int fy, int out[3]) {
  const bool x0 = (unsigned)sx < (unsigned)v.W, x1 = fx != 0 && (unsigned)(sx + 1) < (unsigned)v.W;
  const bool y0 = (unsigned)sy < (unsigned)v.H, y1 = fy != 0 && (unsigned)(sy + 1) < (unsigned)v.H;
  const uint8_t* r0 = staged_px(v, sx, sy);
  const uint8_t* r1 = staged_px(v, sx, sy + 1);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int v00 = (y0 && x0) ? r0[c] : 0, v01 = (y0 && x1) ? r0[3 + c] : 0;
    const int v10 = (y1 && x0) ? r1[c] : 0, v11 = (y1 && x1) ? r1[3 + c] : 0;
    out[c] = bilerp_fix(v00, v01, v10, v11, fx, fy);
  }
}
__device__ __forceinline__ int staged_fetch1(const StageView& v, int sx, int sy, int fx, int fy) {
  const bool x0 = (unsigned)sx < (unsigned)v.W, x1 = fx != 0 && (unsigned)(sx + 1) < (unsigned)v.W;
  const bool y0 = (unsigned)sy < (unsigned)v.H, y1 = fy != 0 && (unsigned)(sy + 1) < (unsigned)v.H;
  const uint8_t* r0 = staged_px(v, sx, sy);
  const uint8_t* r1 = staged_px(v, sx, sy + 1);
  return bilerp_fix((y0 && x0) ? r0[0] : 0, (y0 && x1) ? r0[1] : 0, (y1 && x0) ? r1[0] : 0, (y1 && x1) ? r1[1] : 0, fx, fy);
}
// row offset table of a staged view: rowoff[r] = r * pitch + byte phase of source row by0 + r  (threads 0..rows-1)
__device__ __forceinline__ void fill_rowoff(int* rowoff, const StageView& v, int rows) {
  const int r = threadIdx.x;
  if (r < rows && r < 96) rowoff[r] = r * v.pitch + (int)((v.lo + (uint32_t)(((v.by0 + r) * v.W + v.bx0) * v.C)) & 15u);
}
// all four taps of (sx, sy) lie inside the staged rectangle (hence inside the frame): no per-tap tests needed
__device__ __forceinline__ bool taps_inside(const StageView& v, int sx, int sy) {
  return sx >= v.bx0 && sx + 1 < v.bx1 && sy >= v.by0 && sy + 1 < v.by1;
}
__device__ __forceinline__ void fast_fetch3(const StageView& v, const int* rowoff, int sx, int sy, int fx, int fy, int out[3]) {
  const uint8_t* r0 = v.sm + rowoff[sy - v.by0] + (sx - v.bx0) * 3;
  const uint8_t* r1 = v.sm + rowoff[sy + 1 - v.by0] + (sx - v.bx0) * 3;
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    out[c] = ((int)r0[c] * w00 + (int)r0[3 + c] * w01 + (int)r1[c] * w10 + (int)r1[3 + c] * w11 + (1 << 14)) >> 15;
}
__device__ __forceinline__ int fast_fetch1(const StageView& v, const int* rowoff, int sx, int sy, int fx, int fy) {
  const uint8_t* r0 = v.sm + rowoff[sy - v.by0] + (sx - v.bx0);
  const uint8_t* r1 = v.sm + rowoff[sy + 1 - v.by0] + (sx - v.bx0);
  return bilerp_fix(r0[0], r0[1], r1[0], r1[1], fx, fy);
}
// 4 consecutive pixels (12 bytes) of a u8x3 frame at byte offset o: three words when aligned, bytes otherwise
__device__ __forceinline__ void load12(const uint8_t* p, bool vec, int n, uint32_t w[3]) {
  if (vec) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
    w[0] = q[0]; w[1] = q[1]; w[2] = q[2];
  } else {
    w[0] = w[1] = w[2] = 0u;
    for (int k = 0; k < 3 * n; ++k) w[k >> 2] |= (uint32_t)p[k] << ((k & 3) * 8);
  }
}
__device__ __forceinline__ int byte_of(const uint32_t w[3], int k) { return (int)((w[k >> 2] >> ((k & 3) * 8)) & 255u); }
// ---- bboxes-only chains (bbox_augmentation.py:31-88), one box of one level ----------------------------------
__device__ OADG_HANDLER void bbo_stage(const ChainArgs& A, ChainSmem& S, const Item& I, bool catch_up) {
  worker_sync();
  const int key = I.obj * 2 + (catch_up ? 1 : 0);
  if (S.bs_key == key) return;   // the job is still staged from this CTA's previous tile (uniform: S.bs_key is shared)
  worker_sync();
  if (threadIdx.x == 0) {
    S.bs_key = key;
    const BboJob J = A.bjobs[I.obj];
    const Chain C = A.chains[J.chain];
    const oadg_view_t& V = A.P.views[C.view];
    BboStage& bs = S.bs;
    const oadg_bbo_t& B = A.P.bbo[J.bbo];
    for (int i = 0; i < 6; ++i) bs.minv[i] = B.minv[i];
    for (int i = 0; i < 4; ++i) bs.rect[i] = J.rect[i];
    bs.W = V.W;
    bs.H = V.H;
    bs.gt = B.gt;
    const int level = catch_up ? J.level + 1 : J.level;   // a catch-up runs in the NEXT level's phase
    bs.X = chain_src(C, level);
    bs.Y = chain_dst(C, level);
    bs.x_map = chain_src_map(C, level);
    bs.pad = 0;
    bs.n_excl = catch_up ? J.next_count : 0;
    for (int k = 0; k < bs.n_excl && k < 16; ++k)
      for (int e = 0; e < 4; ++e) bs.excl[k][e] = A.bjobs[J.next_first + k].rect[e];
  }
  worker_sync();
}
// blend of one box, frames WITHOUT a tensor map (row pitch not a multiple of 16 bytes): the source rows of every
// sub-tile are staged by hand (16-byte vector loads) between two block barriers
__device__ OADG_HANDLER void bbo_r_tile_hand(const ChainArgs& A, ChainSmem& S, uint8_t* dyn, const Item& I, int l0, int l1) {
  const BboStage& bs = S.bs;
  const int t = threadIdx.x;
  if (A.debug & 8) {   // reference path: the shared per-pixel body
    const BboJob& J = A.bjobs[I.obj];
    const Chain& C = A.chains[J.chain];
    for (int k = l0; k < l1; ++k) {
      const int tx0 = (bs.rect[0] & ~3) + (k % I.tx) * kBboTileW, ty0 = bs.rect[1] + (k / I.tx) * kBboTileH;
      for (int q = t; q < kBboTileW * kBboTileH; q += kCT) {
        const int x = tx0 + (q & (kBboTileW - 1)), y = ty0 + q / kBboTileW;
        if (x >= bs.rect[0] && x < bs.rect[2] && y < bs.rect[3]) bbo_r_pixel(A.P, C, A.P.bbo[J.bbo], bs.X, bs.Y, x, y);
      }
    }
    return;
  }
  const int W = bs.W, H = bs.H;
  const bool vec = ((W * 3) & 3) == 0 && ((((uintptr_t)bs.X) | ((uintptr_t)bs.Y)) & 3) == 0;
  const int ax0 = bs.rect[0] & ~3, tx = I.tx;
  constexpr int kSubPerTile = kBboTileW / kSubW;   // 64 x 16 sub-tiles per tile, each with its own staged source
  constexpr int kRowsPerTile = kBboTileH / kSubH;
  for (int k2 = kSubPerTile * kRowsPerTile * l0; k2 < kSubPerTile * kRowsPerTile * l1; ++k2) {
    const int k = k2 / (kSubPerTile * kRowsPerTile), sub = k2 % (kSubPerTile * kRowsPerTile);
    const int tx0 = ax0 + (k % tx) * kBboTileW + (sub % kSubPerTile) * kSubW;
    const int ty0 = bs.rect[1] + (k / tx) * kBboTileH + (sub / kSubPerTile) * kSubH;
    const int x0 = imax(tx0, bs.rect[0]), x1 = imin(tx0 + kSubW, bs.rect[2]), y1 = imin(ty0 + kSubH, bs.rect[3]);
    if (x1 <= x0 || y1 <= ty0) continue;   // uniform: the support ends before this sub-tile
    int sr[4];
    const bool any_src = warp_src_rect(bs.minv, x0, ty0, x1, y1, W, H, sr);
    const StageView sv = make_view(dyn, bs.X, W, H, 3, sr);
    const bool staged = any_src && (size_t)(sr[3] - sr[1]) * sv.pitch <= (size_t)kHandStage && !(A.debug & 2);
    // this thread's 4 pixels of the running image are requested before the staging barriers
    const int y = ty0 + (t >> 4), xg = tx0 + (t & 15) * 4;
    const bool active = !(y >= y1 || xg >= x1 || xg + 4 <= x0);
    const size_t o = ((size_t)y * W + xg) * 3;
    const bool full = xg >= x0 && xg + 4 <= x1;
    uint32_t in_w[3] = {0u, 0u, 0u}, out_w[3] = {0u, 0u, 0u};
    if (active) {
      if (full) load12(bs.X + o, vec, 4, in_w);
      else
        for (int i = 0; i < 4; ++i)
          if (xg + i >= x0 && xg + i < x1)
            for (int c = 0; c < 3; ++c) in_w[(3 * i + c) >> 2] |= (uint32_t)bs.X[o + 3 * i + c] << (((3 * i + c) & 3) * 8);
    }
    worker_sync();  // the previous tile's gathers are done (and the profile slices are in place)
    const bool fast = staged && sr[3] - sr[1] <= 96;
    if (staged) {
      stage_rows(dyn, sv.pitch, bs.X, W, H, 3, sr[0], sr[1], sr[3] - sr[1]);
      fill_rowoff(S.rowoff[0], sv, sr[3] - sr[1]);
    }
    worker_sync();
    if (!active) continue;
    const float uy = A.prof_y[(size_t)bs.gt * A.P.max_h + y];
    const WarpRowTerm rt = warp_row_term(bs.minv, y);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = xg + i;
      int v[3] = {byte_of(in_w, 3 * i), byte_of(in_w, 3 * i + 1), byte_of(in_w, 3 * i + 2)};
      if (x >= x0 && x < x1) {
        const float ux = A.prof_x[(size_t)bs.gt * A.P.max_w + x];
        const float m = fmul(uy, ux);
        // m <= 2^-25: fl(1 - m) == 1 and fl(1 - 1) == 0, so img*1 + aug*0 == img exactly
        if (m > 2.98023223876953125e-8f) {
          int sx, sy, fx, fy, a[3];
          warp_coord_row(bs.minv, rt, x, sx, sy, fx, fy);
          if (fast && taps_inside(sv, sx, sy)) fast_fetch3(sv, S.rowoff[0], sx, sy, fx, fy, a);
          else if (staged) staged_fetch3(sv, sx, sy, fx, fy, a);
          else {
            WarpTap tp;
            tp.sx = sx; tp.sy = sy; tp.fx = fx; tp.fy = fy;
            if (any_src) warp_fetch3(LdRW(), bs.X, H, W, tp, a);
            else a[0] = a[1] = a[2] = 0;
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) v[c] = bbo_blend(m, v[c], a[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) out_w[(3 * i + c) >> 2] |= (uint32_t)v[c] << (((3 * i + c) & 3) * 8);
    }
    if (full && vec) {
      uint32_t* q = reinterpret_cast<uint32_t*>(bs.Y + o);
      q[0] = out_w[0]; q[1] = out_w[1]; q[2] = out_w[2];
    } else {
      for (int i = 0; i < 4; ++i)
        if (xg + i >= x0 && xg + i < x1)
          for (int c = 0; c < 3; ++c) bs.Y[o + 3 * i + c] = (uint8_t)byte_of(out_w, 3 * i + c);
    }
  }
}

// ---- TMA-staged affine gathers ------------------------------------------------------------------------------
// The source rectangle of a sub-tile (64 x 16 output pixels) is at most 85 px x 48 rows for every map the reference
// can draw (rotations up to 30 degrees, shears up to 0.3, translations): thread 0 computes it from the four corners
// (cv2's fixed-point map is monotone in x and in y) and issues up to three 16-row TMA boxes per plane into one of two
// stages; parts of a box outside the frame arrive as zeros (BORDER_CONSTANT 0), so the taps need no bounds tests.
// TMA needs 16-byte aligned box columns: the box starts at the byte column of the rectangle rounded down.
// S.gather_rect[stage] = {first staged byte column of the frame, source y0, rows, first staged column of the mask}:
// rows > 0 staged; 0 the rectangle misses the frame (every tap is 0); < 0 the rectangle exceeds the stage (the
// sub-tile then gathers from global memory).
__device__ __forceinline__ void gather_issue(ChainSmem& S, uint8_t* dyn, unsigned stage, const void* img_map,
                                             const void* mask_map, const double* minv, int x0, int y0, int x1, int y1,
                                             int W, int H) {
  int sxa = 1 << 30, sxb = -(1 << 30), sya = 1 << 30, syb = -(1 << 30);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int sx, sy, fx, fy;
    warp_coord(minv, (k & 1) ? x1 - 1 : x0, (k & 2) ? y1 - 1 : y0, sx, sy, fx, fy);
    sxa = imin(sxa, sx); sxb = imax(sxb, sx);
    sya = imin(sya, sy); syb = imax(syb, sy);
  }
  const int rows = syb + 2 - sya, wpx = sxb + 2 - sxa;
  const bool ok = rows <= kMaxSrcRows && wpx <= kMaxSrcPx && sxa > -32768 && sxb < 32767 && sya > -32768 && syb < 32767;
  const bool inside = sxa < W && sxa + wpx > 0 && sya < H && sya + rows > 0;
  int* r = S.gather_rect[stage];
  const int xb = (3 * sxa) & ~15, xm = sxa & ~15;
  r[0] = xb;
  r[1] = sya;
  r[2] = !ok ? -1 : (inside ? rows : 0);
  r[3] = xm;
  uint64_t* bar = &S.full[stage];
  if (ok && inside) {
    const int nb = (rows + kGatherBoxRows - 1) / kGatherBoxRows;
    uint8_t* img = dyn + stage * kStageBytes;
    uint8_t* msk = img + kStageImgBytes;
    tma::mbar_expect_tx(bar, (uint32_t)nb * (uint32_t)(kGatherBoxRows * (kGatherImgBoxBytes + (mask_map ? kGatherMaskBoxBytes : 0))));
    for (int b = 0; b < nb; ++b) {
      tma::load_2d(img + b * (kGatherBoxRows * kGatherImgBoxBytes), img_map, bar, xb, sya + b * kGatherBoxRows);
      if (mask_map)
        tma::load_2d(msk + b * (kGatherBoxRows * kGatherMaskBoxBytes), mask_map, bar, xm, sya + b * kGatherBoxRows);
    }
  } else {
    tma::mbar_arrive(bar);
  }
}
// Which taps carry weight is a property of the map: with m00 == 1, m01 == 0 and an integral m02 every X is a multiple
// of 32 (fx == 0: translations, y-shears), likewise fy == 0 for m10 == 0, m11 == 1 and an integral m12 (translations,
// x-shears); the weights (32-fx)(32-fy)*32 ... then reduce to the one- / two-tap forms below (same integers).
//   mode bit 0: fx == 0 everywhere, bit 1: fy == 0 everywhere
__device__ __forceinline__ int gather_mode(const double* m) {
  const bool fx0 = m[0] == 1.0 && m[1] == 0.0 && m[2] == floor(m[2]);
  const bool fy0 = m[3] == 0.0 && m[4] == 1.0 && m[5] == floor(m[5]);
  return (fx0 ? 1 : 0) | (fy0 ? 2 : 0);
}
// taps of one pixel from a staged frame rectangle: byte loads, no bounds tests
template <int kMode>
__device__ __forceinline__ void tma_fetch3(const uint8_t* st, int xb, int ry0, int sx, int sy, int fx, int fy, int out[3]) {
  const uint8_t* p = st + (sy - ry0) * kGatherImgBoxBytes + (sx * 3 - xb);
  if (kMode == 3) {
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = p[c];
  } else if (kMode == 1) {   // fx == 0: vertical pair
    const int w0 = 32 - fy;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = ((int)p[c] * w0 + (int)p[kGatherImgBoxBytes + c] * fy + 16) >> 5;
  } else if (kMode == 2) {   // fy == 0: horizontal pair
    const int w0 = 32 - fx;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = ((int)p[c] * w0 + (int)p[3 + c] * fx + 16) >> 5;
  } else {
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[c] = ((int)p[c] * w00 + (int)p[3 + c] * w01 + (int)p[kGatherImgBoxBytes + c] * w10 +
                (int)p[kGatherImgBoxBytes + 3 + c] * w11 + (1 << 14)) >> 15;
  }
}
template <int kMode>
__device__ __forceinline__ int tma_fetch1(const uint8_t* st, int xm, int ry0, int sx, int sy, int fx, int fy) {
  const uint8_t* p = st + (sy - ry0) * kGatherMaskBoxBytes + (sx - xm);
  if (kMode == 3) return p[0];
  if (kMode == 1) return ((int)p[0] * (32 - fy) + (int)p[kGatherMaskBoxBytes] * fy + 16) >> 5;
  if (kMode == 2) return ((int)p[0] * (32 - fx) + (int)p[1] * fx + 16) >> 5;
  return bilerp_fix(p[0], p[1], p[kGatherMaskBoxBytes], p[kGatherMaskBoxBytes + 1], fx, fy);
}
// Conversion-free forms of the blends (u8_to_f32 / f32_trunc_u8 / u8_to_f64 / f64_trunc_u8, oamix_math.h):
// bbox_augmentation.py:63-71 (bbo_blend, oamix_math.h) with the box's two float32 factors hoisted
__device__ __forceinline__ int bbo_blend_fast(float mask, float rest, int img, int aug) {
  return f32_trunc_u8(fadd(fmul(u8_to_f32(img), mask), fmul(u8_to_f32(aug), rest)));
}
// bbox_augmentation.py:264-272 (bg_blend, oamix_math.h)
__device__ __forceinline__ int bg_blend_fast(double keep, double rest, int img, int aug) {
  return f64_trunc_u8(dadd(dmul(keep, u8_to_f64(img)), dmul(rest, u8_to_f64(aug))));
}
__device__ __forceinline__ void col_coord(WarpRowTerm r, int2 cd, int& sx, int& sy, int& fx, int& fy) {
  const int X = (r.X0 + cd.x) >> 5, Y = (r.Y0 + cd.y) >> 5;
  sx = X >> 5;
  sy = Y >> 5;
  fx = X & 31;
  fy = Y & 31;
}

// the 4 pixels of one thread of a blend sub-tile whose source rectangle is staged (rows > 0)
struct BboPx {
  const uint8_t* st;   // staged frame rows
  int xb, ry0;         // first staged byte column, first staged row
  const int2* col;     // this thread's 4 column terms
  const float* uxs;    // ... and x-profile values
  WarpRowTerm rt;
  float uy;
  int lo, hi;          // pixels lo <= i < hi of the group lie inside the sub-tile
};
template <int kMode>
__device__ __forceinline__ void bbo_pixels4(const BboPx& P, const uint32_t in_w[3], uint32_t out_w[3]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int v[3] = {byte_of(in_w, 3 * i), byte_of(in_w, 3 * i + 1), byte_of(in_w, 3 * i + 2)};
    if (i >= P.lo && i < P.hi) {
      const float m = fmul(P.uy, P.uxs[i]);
      // m <= 2^-25: fl(1 - m) == 1 and fl(1 - 1) == 0, so img*1 + aug*0 == img exactly
      if (m > 2.98023223876953125e-8f) {
        int sx, sy, fx, fy, a[3];
        col_coord(P.rt, P.col[i], sx, sy, fx, fy);
        tma_fetch3<kMode>(P.st, P.xb, P.ry0, sx, sy, fx, fy, a);
        const float mask = fsub(1.0f, m), rest = fsub(1.0f, mask);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = bbo_blend_fast(mask, rest, v[c], a[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out_w[(3 * i + c) >> 2] |= (uint32_t)v[c] << (((3 * i + c) & 3) * 8);
  }
}

// blend of one box, one tile of 256 x 16 px: Y = uint8(X*(1-m) + warp(X)*m) inside the support
// (bbox_augmentation.py:57-71).  Returns the advanced gather-stage counter.
__device__ OADG_HANDLER unsigned bbo_r_tile(const ChainArgs& A, ChainSmem& S, uint8_t* dyn, const Item& I, int local,
                                            unsigned n_stage) {
  bbo_stage(A, S, I, false);
  const BboStage& bs = S.bs;
  if (bs.x_map < 0 || (A.debug & (2 | 8 | 32))) {
    bbo_r_tile_hand(A, S, dyn, I, local, local + 1);
    return n_stage;
  }
  const int t = threadIdx.x, W = bs.W, H = bs.H;
  const int tile_x0 = (bs.rect[0] & ~3) + (local % I.tx) * kBboTileW, ty0 = bs.rect[1] + (local / I.tx) * kBboTileH;
  const int tile_x1 = imin(tile_x0 + kBboTileW, bs.rect[2]), y1 = imin(ty0 + kBboTileH, bs.rect[3]);
  const int nsub = (tile_x1 - tile_x0 + kSubW - 1) / kSubW;
  int2* col = reinterpret_cast<int2*>(dyn + kColOff);
  float* uxs = reinterpret_cast<float*>(dyn + kUxOff);
  for (int c = t; c < tile_x1 - tile_x0; c += kCT) {   // per column: cv2's adelta / bdelta and the box's x-profile
    const int x = tile_x0 + c;
    col[c] = make_int2(cv_round(dmul(dmul(bs.minv[0], (double)x), 1024.0)), cv_round(dmul(dmul(bs.minv[3], (double)x), 1024.0)));
    uxs[c] = x >= bs.rect[0] ? A.prof_x[(size_t)bs.gt * A.P.max_w + x] : 0.f;
  }
  const void* map = static_cast<const char*>(A.maps) + (size_t)bs.x_map * kTensorMapBytes;
  if (t == 0) {
    tma::fence_tensormap_acquire(map);
    tma::fence_proxy_async();
    gather_issue(S, dyn, n_stage & 1u, map, nullptr, bs.minv, imax(tile_x0, bs.rect[0]), ty0, imin(tile_x0 + kSubW, tile_x1),
                 y1, W, H);
  }
  const int y = ty0 + (t >> 4);
  const bool row_ok = y < y1;
  const float uy = row_ok ? A.prof_y[(size_t)bs.gt * A.P.max_h + y] : 0.f;
  const WarpRowTerm rt = warp_row_term(bs.minv, y);
  const int mode = gather_mode(bs.minv);
  worker_sync();   // the column tables are in place
  for (int s = 0; s < nsub; ++s, ++n_stage) {
    const int sx0 = tile_x0 + s * kSubW;
    const int x0 = imax(sx0, bs.rect[0]), x1 = imin(sx0 + kSubW, tile_x1);
    if (t == 0 && s + 1 < nsub)
      gather_issue(S, dyn, (n_stage + 1) & 1u, map, nullptr, bs.minv, sx0 + kSubW, ty0, imin(sx0 + 2 * kSubW, tile_x1), y1, W, H);
    // this thread's 4 pixels of the running image are requested before it waits for the staged rows
    const int xg = sx0 + (t & 15) * 4;
    const bool active = row_ok && xg < x1 && xg + 4 > x0;
    const bool full = xg >= x0 && xg + 4 <= x1;
    const size_t o = ((size_t)y * W + xg) * 3;
    uint32_t in_w[3] = {0u, 0u, 0u}, out_w[3] = {0u, 0u, 0u};
    if (active) {
      if (full) load12(bs.X + o, true, 4, in_w);
      else
        for (int i = 0; i < 4; ++i)
          if (xg + i >= x0 && xg + i < x1)
            for (int c = 0; c < 3; ++c) in_w[(3 * i + c) >> 2] |= (uint32_t)bs.X[o + 3 * i + c] << (((3 * i + c) & 3) * 8);
    }
    tma::mbar_wait(&S.full[n_stage & 1u], (n_stage >> 1) & 1u);
    if (active) {
      const int* gr = S.gather_rect[n_stage & 1u];
      const int rows = gr[2];
      BboPx P4;
      P4.st = dyn + (n_stage & 1u) * kStageBytes;
      P4.xb = gr[0]; P4.ry0 = gr[1];
      P4.col = col + (xg - tile_x0); P4.uxs = uxs + (xg - tile_x0);
      P4.rt = rt; P4.uy = uy; P4.lo = x0 - xg; P4.hi = x1 - xg;
      if (rows > 0) {
        switch (mode) {
          case 0: bbo_pixels4<0>(P4, in_w, out_w); break;
          case 1: bbo_pixels4<1>(P4, in_w, out_w); break;
          case 2: bbo_pixels4<2>(P4, in_w, out_w); break;
          default: bbo_pixels4<3>(P4, in_w, out_w); break;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int x = xg + i;
          int v[3] = {byte_of(in_w, 3 * i), byte_of(in_w, 3 * i + 1), byte_of(in_w, 3 * i + 2)};
          if (x >= x0 && x < x1) {
            const float m = fmul(uy, uxs[x - tile_x0]);
            if (m > 2.98023223876953125e-8f) {
              int sx, sy, fx, fy, a[3] = {0, 0, 0};
              col_coord(rt, col[x - tile_x0], sx, sy, fx, fy);
              if (rows < 0) {
                WarpTap tp;
                tp.sx = imin(imax(sx, -32768), 32767); tp.sy = imin(imax(sy, -32768), 32767); tp.fx = fx; tp.fy = fy;
                warp_fetch3(LdRW(), bs.X, H, W, tp, a);
              }
#pragma unroll
              for (int c = 0; c < 3; ++c) v[c] = bbo_blend(m, v[c], a[c]);
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) out_w[(3 * i + c) >> 2] |= (uint32_t)v[c] << (((3 * i + c) & 3) * 8);
        }
      }
      if (full) {
        uint32_t* q = reinterpret_cast<uint32_t*>(bs.Y + o);
        q[0] = out_w[0]; q[1] = out_w[1]; q[2] = out_w[2];
      } else {
        for (int i = 0; i < 4; ++i)
          if (xg + i >= x0 && xg + i < x1)
            for (int c = 0; c < 3; ++c) bs.Y[o + 3 * i + c] = (uint8_t)byte_of(out_w, 3 * i + c);
      }
    }
    worker_sync();   // the stage (and its rectangle record) may be refilled
  }
  return n_stage;
}
// catch-up copy of a level l-1 support into the frame level l writes, minus the supports level l rewrites
__device__ OADG_HANDLER void bbo_c_segment(const ChainArgs& A, ChainSmem& S, const Item& I, int l0, int l1) {
  bbo_stage(A, S, I, true);
  const BboStage& bs = S.bs;
  const int t = threadIdx.x, W = bs.W;
  const bool vec = ((W * 3) & 3) == 0 && ((((uintptr_t)bs.X) | ((uintptr_t)bs.Y)) & 3) == 0;
  const int ax0 = bs.rect[0] & ~3, tx = I.tx;
  const BboJob* jobs = A.bjobs;
  const BboJob& J = A.bjobs[I.obj];
  if (A.debug & 16) {   // reference path: the shared per-pixel body
    for (int k = l0; k < l1; ++k) {
      const int tx0 = ax0 + (k % tx) * kBboCatchW, ty0 = bs.rect[1] + (k / tx) * kBboTileH;
      for (int q = t; q < kBboCatchW * kBboTileH; q += kCT) {
        const int x = tx0 + (q & (kBboCatchW - 1)), y = ty0 + q / kBboCatchW;
        if (x >= bs.rect[0] && x < bs.rect[2] && y < bs.rect[3]) bbo_c_pixel(jobs, J, W, bs.X, bs.Y, x, y);
      }
    }
    return;
  }
  const bool vec16 = ((W * 3) & 15) == 0 && ((((uintptr_t)bs.X) | ((uintptr_t)bs.Y)) & 15) == 0;
  for (int k = l0; k < l1; ++k) {
    const int tx0 = ax0 + (k % tx) * kBboCatchW, ty0 = bs.rect[1] + (k / tx) * kBboTileH;
    const int x0 = imax(tx0, bs.rect[0]), x1 = imin(tx0 + kBboCatchW, bs.rect[2]), y1 = imin(ty0 + kBboTileH, bs.rect[3]);
    if (x1 <= x0 || y1 <= ty0) continue;
    // the next level's supports that meet this tile (uniform); one that covers it leaves nothing to copy
    unsigned hit = 0;
    bool covered = false;
    const bool many = bs.n_excl > 16;
    for (int e = 0; e < bs.n_excl && !many; ++e) {
      const int32_t* r = bs.excl[e];
      if (r[0] < x1 && r[2] > x0 && r[1] < y1 && r[3] > ty0) {
        hit |= 1u << e;
        covered |= r[0] <= x0 && r[2] >= x1 && r[1] <= ty0 && r[3] >= y1;
      }
    }
    if (covered) continue;
    // 16-pixel chunks on the frame's 16-pixel grid (48 bytes = three aligned 16-byte vectors), one (chunk, row) per
    // thread.  skip = the pixels of the chunk that are not this tile's to copy (outside the support's columns, or
    // inside a next-level support: that level's blend writes them in this very phase, they must not be touched).
    const int c0 = x0 >> 4, nc = ((x1 - 1) >> 4) - c0 + 1, nrow = y1 - ty0;
    if (many || !vec16) {
      const int wpx = x1 - x0;
      for (int q = t; q < wpx * nrow; q += kCT) bbo_c_pixel(jobs, J, W, bs.X, bs.Y, x0 + q % wpx, ty0 + q / wpx);
      continue;
    }
    for (int idx = t; idx < nc * nrow; idx += kCT) {
      const int y = ty0 + idx / nc, xc = (c0 + idx % nc) * 16;
      unsigned skip = 0;
      if (xc < x0) skip |= (1u << (x0 - xc)) - 1u;
      if (xc + 16 > x1) skip |= 0xFFFFu & ~((1u << (x1 - xc)) - 1u);
      for (unsigned h = hit; h; h &= h - 1) {
        const int32_t* r = bs.excl[__ffs(h) - 1];
        if (y < r[1] || y >= r[3]) continue;
        const int a = imax(r[0] - xc, 0), b = imin(r[2] - xc, 16);
        if (b > a) skip |= ((1u << b) - 1u) & ~((1u << a) - 1u);
      }
      if (skip == 0xFFFFu) continue;
      const size_t o = ((size_t)y * W + xc) * 3;
      if (skip == 0u) {
        const uint4* q = reinterpret_cast<const uint4*>(bs.X + o);
        const uint4 a0 = q[0], a1 = q[1], a2 = q[2];
        uint4* d = reinterpret_cast<uint4*>(bs.Y + o);
        d[0] = a0; d[1] = a1; d[2] = a2;
        continue;
      }
      for (int i = 0; i < 16; ++i)
        if (!(skip >> i & 1u)) {
          bs.Y[o + 3 * i] = bs.X[o + 3 * i];
          bs.Y[o + 3 * i + 1] = bs.X[o + 3 * i + 1];
          bs.Y[o + 3 * i + 2] = bs.X[o + 3 * i + 2];
        }
    }
  }
}

// one pixel of a bg-only op with hoisted coordinate terms (same arithmetic as bg_pixel / eval_op)
__device__ __forceinline__ void bg_pixel_fast(const DevPlan& P, const Lane& L, const RegOp& R, int ax, int bx,
                                              const double* div255, int x, int y) {
  const int X = (cv_round(dmul(dadd(dmul(R.minv[1], (double)y), R.minv[2]), 1024.0)) + 16 + ax) >> 5;
  const int Y = (cv_round(dmul(dadd(dmul(R.minv[4], (double)y), R.minv[5]), 1024.0)) + 16 + bx) >> 5;
  WarpTap t;
  t.sx = imin(imax(X >> 5, -32768), 32767);
  t.sy = imin(imax(Y >> 5, -32768), 32767);
  t.fx = X & 31;
  t.fy = Y & 31;
  int px[3];
  warp_fetch3(LdRO(), L.in, L.H, L.W, t, px);
  const size_t mo = (size_t)L.view * P.mask_stride;
  const float M = P.maskf[mo + (size_t)y * L.W + x];
  const uint8_t* mu = P.masku + mo;
  const bool x0 = (unsigned)t.sx < (unsigned)L.W, x1 = t.fx != 0 && (unsigned)(t.sx + 1) < (unsigned)L.W;
  const bool y0 = (unsigned)t.sy < (unsigned)L.H, y1 = t.fy != 0 && (unsigned)(t.sy + 1) < (unsigned)L.H;
  const uint8_t* r0 = mu + (size_t)t.sy * L.W + t.sx;
  const uint8_t* r1 = r0 + L.W;
  const int wm = bilerp_fix((y0 && x0) ? ldb(r0) : 0, (y0 && x1) ? ldb(r0 + 1) : 0, (y1 && x0) ? ldb(r1) : 0,
                            (y1 && x1) ? ldb(r1 + 1) : 0, t.fx, t.fy);
  const size_t o = ((size_t)y * L.W + x) * 3;
  if (M != 0.f || wm != 0) {  // keep == 0 => 0*img + 1*aug == aug exactly
    const double am = div255[wm];  // wm / 255 in float64, tabulated (exactly the reference's quotient)
    const double keep = (double)M > am ? (double)M : am;
    const double rest = dsub(1.0, keep);
#pragma unroll
    for (int c = 0; c < 3; ++c)
      px[c] = (int)dadd(dmul(keep, (double)ldb(L.in + o + c)), dmul(rest, (double)px[c]));
  }
  uint8_t* q = L.out + o;
  q[0] = (uint8_t)px[0];
  q[1] = (uint8_t)px[1];
  q[2] = (uint8_t)px[2];
}
// bg-only op (bbox_augmentation.py:240-272) on a sub-tile of 64 x 16 px that one region covers: the frame and the
// uint8 union mask are both warped from staged shared-memory rows; 4 pixels per thread.
__device__ OADG_HANDLER void bg_subtile(const ChainArgs& A, ChainSmem& S, uint8_t* dyn, const Lane& L, const RegOp& R,
                                        int r_only, int x0, int y0, int x1, int y1, const double* div255) {
  const DevPlan& P = A.P;
  const int W = L.W, H = L.H, t = threadIdx.x;
  int sr[4];
  const bool any_src = warp_src_rect(R.minv, x0, y0, x1, y1, W, H, sr);
  const int pitch_i = any_src ? stage_pitch(sr[0], sr[2], 3) : 0, pitch_m = any_src ? stage_pitch(sr[0], sr[2], 1) : 0;
  const int rows = any_src ? sr[3] - sr[1] : 0;
  const bool staged = any_src && (size_t)rows * (pitch_i + pitch_m) <= (size_t)kHandStage;
  const uint8_t* mu = P.masku + (size_t)L.view * P.mask_stride;
  const StageView si = make_view(dyn, L.in, W, H, 3, sr);
  const StageView sm = make_view(dyn + (size_t)rows * pitch_i, mu, W, H, 1, sr);
  // this thread's 4 pixels: the frame bytes and the float mask are requested before the staging barriers
  const int y = y0 + (t >> 4), xg = x0 + (t & 15) * 4;
  const bool active = y < y1 && xg < x1;
  const int n = active ? imin(4, x1 - xg) : 0;
  const size_t o = ((size_t)y * W + xg) * 3;
  const bool vec = n == 4 && ((W * 3) & 3) == 0 && ((((uintptr_t)L.in) | ((uintptr_t)L.out)) & 3) == 0;
  uint32_t in_w[3] = {0u, 0u, 0u}, out_w[3] = {0u, 0u, 0u};
  float Mv[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    load12(L.in + o, vec, n, in_w);
    const float* mf = P.maskf + (size_t)L.view * P.mask_stride + (size_t)y * W + xg;
    if (n == 4 && (((uintptr_t)mf) & 15) == 0) {
      const float4 q = *reinterpret_cast<const float4*>(mf);
      Mv[0] = q.x; Mv[1] = q.y; Mv[2] = q.z; Mv[3] = q.w;
    } else {
      for (int i = 0; i < n; ++i) Mv[i] = mf[i];
    }
  }
  worker_sync();  // the previous sub-tile's gathers are done
  const bool fast = staged && rows <= 96;
  if (staged) {
    stage_rows(dyn, pitch_i, L.in, W, H, 3, sr[0], sr[1], rows);
    stage_rows(dyn + (size_t)rows * pitch_i, pitch_m, mu, W, H, 1, sr[0], sr[1], rows);
    fill_rowoff(S.rowoff[0], si, rows);
    fill_rowoff(S.rowoff[1], sm, rows);
  }
  worker_sync();
  if (!active) return;
  const WarpRowTerm rt = warp_row_term(R.minv, y);
  unsigned keep_mask = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int px[3] = {0, 0, 0};
    const bool mine = i < n && (r_only < 0 || region_of_pixel(L, xg + i, y) == r_only);
    if (mine) {
      keep_mask |= 1u << i;
      const int x = xg + i;
      int sx, sy, fx, fy, wm = 0;
      warp_coord_row(R.minv, rt, x, sx, sy, fx, fy);
      if (fast && taps_inside(si, sx, sy)) {
        fast_fetch3(si, S.rowoff[0], sx, sy, fx, fy, px);
        wm = fast_fetch1(sm, S.rowoff[1], sx, sy, fx, fy);
      } else if (staged) {
        staged_fetch3(si, sx, sy, fx, fy, px);
        wm = staged_fetch1(sm, sx, sy, fx, fy);
      } else if (any_src) {
        WarpTap tp;
        tp.sx = sx; tp.sy = sy; tp.fx = fx; tp.fy = fy;
        warp_fetch3(LdRO(), L.in, H, W, tp, px);
        const bool bx0 = (unsigned)sx < (unsigned)W, bx1 = fx != 0 && (unsigned)(sx + 1) < (unsigned)W;
        const bool by0 = (unsigned)sy < (unsigned)H, by1 = fy != 0 && (unsigned)(sy + 1) < (unsigned)H;
        const uint8_t* r0 = mu + (size_t)sy * W + sx;
        const uint8_t* r1 = r0 + W;
        wm = bilerp_fix((by0 && bx0) ? ldb(r0) : 0, (by0 && bx1) ? ldb(r0 + 1) : 0, (by1 && bx0) ? ldb(r1) : 0,
                        (by1 && bx1) ? ldb(r1 + 1) : 0, fx, fy);
      }
      const float M = Mv[i];
      if (M != 0.f || wm != 0) {  // keep == 0 => 0*img + 1*aug == aug exactly
        const double am = div255[wm];  // wm / 255 in float64, tabulated (exactly the reference's quotient)
        const double keep = (double)M > am ? (double)M : am;
        const double rest = dsub(1.0, keep);
#pragma unroll
        for (int c = 0; c < 3; ++c)
          px[c] = (int)dadd(dmul(keep, (double)byte_of(in_w, 3 * i + c)), dmul(rest, (double)px[c]));
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out_w[(3 * i + c) >> 2] |= (uint32_t)px[c] << (((3 * i + c) & 3) * 8);
  }
  if (vec && keep_mask == 15u) {
    uint32_t* q = reinterpret_cast<uint32_t*>(L.out + o);
    q[0] = out_w[0]; q[1] = out_w[1]; q[2] = out_w[2];
  } else {
    for (int i = 0; i < 4; ++i)
      if (keep_mask >> i & 1)
        for (int c = 0; c < 3; ++c) L.out[o + 3 * i + c] = (uint8_t)byte_of(out_w, 3 * i + c);
  }
}
// one pixel of a non-bg op from the staged lane record, region op and LUTs (same arithmetic as eval_op, oamix_body.h)
__device__ __forceinline__ void pixel_op_fast(const ChainArgs& A, const Lane& L, const RegOp& R, const uint8_t* luts,
                                              int r, int x, int y) {
  const int W = L.W, H = L.H;
  const size_t o = ((size_t)y * W + x) * 3;
  uint8_t* q = L.out + o;
  const int kind = R.kind;
  if (kind == OADG_OP_BBO_AFFINE) {
    const uint8_t* s = (L.scratch[r] >= 0 ? A.scratch + (size_t)L.scratch[r] * A.frame_bytes : L.in) + o;
    q[0] = s[0]; q[1] = s[1]; q[2] = s[2];
    return;
  }
  const int v0 = L.in[o], v1 = L.in[o + 1], v2 = L.in[o + 2];
  int out[3] = {v0, v1, v2};
  if (is_lut_kind(kind)) {
    const uint8_t* lut = luts + r * 768;
    out[0] = lut[v0]; out[1] = lut[256 + v1]; out[2] = lut[512 + v2];
  } else if (kind == OADG_OP_INVERT) {
    const int xs = x - R.p0, ys = y - R.p1;
    if ((unsigned)xs < (unsigned)W && (unsigned)ys < (unsigned)H) {
      const uint8_t* p = L.in + ((size_t)ys * W + xs) * 3;
      out[0] = (-(int)p[0]) & 255; out[1] = (-(int)p[1]) & 255; out[2] = (-(int)p[2]) & 255;
    } else {
      out[0] = out[1] = out[2] = 0;
    }
  } else if (kind == OADG_OP_COLOR) {
    const int deg = pil_luma(v0, v1, v2);
    out[0] = pil_blend(deg, v0, R.factor); out[1] = pil_blend(deg, v1, R.factor); out[2] = pil_blend(deg, v2, R.factor);
  } else if (kind == OADG_OP_SHARPNESS) {
    if (x == 0 || y == 0 || x == W - 1 || y == H - 1) {  // SMOOTH copies the border
      for (int c = 0; c < 3; ++c) out[c] = pil_blend(out[c], out[c], R.factor);
    } else {
      for (int c = 0; c < 3; ++c) {
        int nb[9];
        for (int dy = 0; dy < 3; ++dy)
          for (int dx = 0; dx < 3; ++dx) nb[dy * 3 + dx] = L.in[((size_t)(y + 1 - dy) * W + (x - 1 + dx)) * 3 + c];
        out[c] = pil_blend(pil_smooth9(nb), out[c], R.factor);
      }
    }
  }
  q[0] = (uint8_t)out[0]; q[1] = (uint8_t)out[1]; q[2] = (uint8_t)out[2];
}

// the 4 pixels of one thread of a bg-only sub-tile whose source rectangle is staged (rows > 0)
struct BgPx {
  const uint8_t* st;     // staged frame rows, followed by the staged mask rows
  int xb, ry0, xm;
  const int2* col;
  WarpRowTerm rt;
  int n;                 // pixels of the group inside the frame
  const double* div255;  // shared memory
};
template <int kMode>
__device__ __forceinline__ void bg_pixels4(const BgPx& P, const uint32_t in_w[3], const float Mv[4], uint32_t out_w[3]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int px[3] = {0, 0, 0};
    if (i < P.n) {
      int sx, sy, fx, fy;
      col_coord(P.rt, P.col[i], sx, sy, fx, fy);
      tma_fetch3<kMode>(P.st, P.xb, P.ry0, sx, sy, fx, fy, px);
      const int wm = tma_fetch1<kMode>(P.st + kStageImgBytes, P.xm, P.ry0, sx, sy, fx, fy);
      const float M = Mv[i];
      if (M != 0.f || wm != 0) {  // keep == 0 => 0*img + 1*aug == aug exactly
        const double am = P.div255[wm];  // wm / 255 in float64, tabulated (exactly the reference's quotient)
        const double keep = (double)M > am ? (double)M : am;
        const double rest = dsub(1.0, keep);
#pragma unroll
        for (int c = 0; c < 3; ++c) px[c] = bg_blend_fast(keep, rest, byte_of(in_w, 3 * i + c), px[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out_w[(3 * i + c) >> 2] |= (uint32_t)px[c] << (((3 * i + c) & 3) * 8);
  }
}

// ------------------------------------------------------------------------------------
// one tile (512 or 256 x 16 px) of one depth step of one lane (oa_mix.py:226-234), handled per 64 x 16 sub-tile:
//   STREAM  one table-lookup / bbo-copy region covers the sub-tile: 16-pixel runs move as three 16-byte vectors per
//           thread through the LUTs in shared memory;
//   BG      one bg-only region covers it (bbox_augmentation.py:240-272): frame and uint8 union mask are resampled
//           from TMA-staged rows (gather_issue), 4 pixels per thread, float64 blend;
//   PIXEL   everything else (a multi-level box edge crosses it, or invert / colour / sharpness): per pixel.
// Returns the advanced gather-stage counter.
// ------------------------------------------------------------------------------------
enum { kSubStream = 0, kSubBg = 1, kSubPixel = 2 };

__device__ OADG_HANDLER unsigned step_tile(const ChainArgs& A, ChainSmem& S, uint8_t* dyn, int local, int tx, int tw,
                                           unsigned n_stage) {
  const Lane& L = S.lane;
  const int W = L.W, H = L.H, t = threadIdx.x;
  const int x0 = (local % tx) * tw, y0 = (local / tx) * kStepTileH;
  const int x1 = min(x0 + tw, W), y1 = min(y0 + kStepTileH, H);
  const int nsub = (x1 - x0 + kSubW - 1) / kSubW;
  worker_sync();   // the previous tile is done with the sub-tile classes and the column table
  if (t < nsub) {
    const int sx0 = x0 + t * kSubW, sx1 = min(sx0 + kSubW, x1);
    const int region = tile_region(L, sx0, y0, sx1, y1);
    int cls = kSubPixel;
    if (region >= 0) {
      const int kind = S.rop[region].kind;
      if (kind_streams(kind)) cls = kSubStream | region << 4;
      else if (kind == OADG_OP_BG_AFFINE && !(A.debug & 1)) cls = kSubBg | region << 4;
    }
    S.sub_cls[t] = cls;
  }
  worker_sync();
  bool any_bg = false, any_pixel = false;
  for (int s = 0; s < nsub; ++s) {
    any_bg |= (S.sub_cls[s] & 15) == kSubBg;
    any_pixel |= (S.sub_cls[s] & 15) == kSubPixel;
  }
  if (t == 0) S.step_class = any_pixel ? 9 : (any_bg ? 8 : 7);
  const bool use_tma = L.in_map >= 0 && L.mask_map >= 0 && !(A.debug & 32);
  if (any_bg) {
    const uint8_t* mu = A.P.masku + (size_t)L.view * A.P.mask_stride;
    const float* mf = A.P.maskf + (size_t)L.view * A.P.mask_stride;
    for (int r = 0; r <= L.n_ml; ++r) {
      if (S.rop[r].kind != OADG_OP_BG_AFFINE) continue;
      const int want = kSubBg | r << 4;
      int first = -1;
      for (int s = nsub - 1; s >= 0; --s)
        if (S.sub_cls[s] == want) first = s;
      if (first < 0) continue;
      const RegOp& R = S.rop[r];
      if (!use_tma) {   // frames without a tensor map: hand-staged rows
        for (int s = first; s < nsub; ++s)
          if (S.sub_cls[s] == want)
            bg_subtile(A, S, dyn, L, R, -1, x0 + s * kSubW, y0, min(x0 + (s + 1) * kSubW, x1), y1, S.div255);
        continue;
      }
      int2* col = reinterpret_cast<int2*>(dyn + kColOff);
      for (int c = t; c < x1 - x0; c += kCT)
        col[c] = make_int2(cv_round(dmul(dmul(R.minv[0], (double)(x0 + c)), 1024.0)),
                           cv_round(dmul(dmul(R.minv[3], (double)(x0 + c)), 1024.0)));
      const void* imap = static_cast<const char*>(A.maps) + (size_t)L.in_map * kTensorMapBytes;
      const void* mmap = static_cast<const char*>(A.maps) + (size_t)L.mask_map * kTensorMapBytes;
      if (t == 0) {
        tma::fence_tensormap_acquire(imap);
        tma::fence_tensormap_acquire(mmap);
        tma::fence_proxy_async();
        gather_issue(S, dyn, n_stage & 1u, imap, mmap, R.minv, x0 + first * kSubW, y0, min(x0 + (first + 1) * kSubW, x1), y1, W, H);
      }
      const int y = y0 + (t >> 4);
      const bool row_ok = y < y1;
      const WarpRowTerm rt = warp_row_term(R.minv, y);
      const int mode = gather_mode(R.minv);
      worker_sync();   // the column table is in place
      for (int s = first; s >= 0;) {
        int nxt = -1;
        for (int k = nsub - 1; k > s; --k)
          if (S.sub_cls[k] == want) nxt = k;
        const int sx0 = x0 + s * kSubW, sx1 = min(sx0 + kSubW, x1);
        if (t == 0 && nxt >= 0)
          gather_issue(S, dyn, (n_stage + 1) & 1u, imap, mmap, R.minv, x0 + nxt * kSubW, y0, min(x0 + (nxt + 1) * kSubW, x1), y1, W, H);
        // this thread's 4 pixels: the frame bytes and the float mask are requested before it waits for the staged rows
        const int xg = sx0 + (t & 15) * 4;
        const bool active = row_ok && xg < sx1;
        const int n = active ? imin(4, sx1 - xg) : 0;
        const size_t o = ((size_t)y * W + xg) * 3;
        uint32_t in_w[3] = {0u, 0u, 0u}, out_w[3] = {0u, 0u, 0u};
        float Mv[4] = {0.f, 0.f, 0.f, 0.f};
        if (active) {
          load12(L.in + o, n == 4, n, in_w);
          const float* mp = mf + (size_t)y * W + xg;
          if (n == 4) {
            const float4 q = *reinterpret_cast<const float4*>(mp);
            Mv[0] = q.x; Mv[1] = q.y; Mv[2] = q.z; Mv[3] = q.w;
          } else {
            for (int i = 0; i < n; ++i) Mv[i] = mp[i];
          }
        }
        tma::mbar_wait(&S.full[n_stage & 1u], (n_stage >> 1) & 1u);
        if (active) {
          const int* gr = S.gather_rect[n_stage & 1u];
          const int rows = gr[2];
          if (rows > 0) {
            BgPx P4;
            P4.st = dyn + (n_stage & 1u) * kStageBytes;
            P4.xb = gr[0]; P4.ry0 = gr[1]; P4.xm = gr[3];
            P4.col = col + (xg - x0);
            P4.rt = rt; P4.n = n; P4.div255 = S.div255;
            switch (mode) {
              case 0: bg_pixels4<0>(P4, in_w, Mv, out_w); break;
              case 1: bg_pixels4<1>(P4, in_w, Mv, out_w); break;
              case 2: bg_pixels4<2>(P4, in_w, Mv, out_w); break;
              default: bg_pixels4<3>(P4, in_w, Mv, out_w); break;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              int px[3] = {0, 0, 0};
              if (i < n) {
                int sx, sy, fx, fy, wm = 0;
                col_coord(rt, col[xg + i - x0], sx, sy, fx, fy);
                if (rows < 0) {
                  WarpTap tp;
                  tp.sx = imin(imax(sx, -32768), 32767); tp.sy = imin(imax(sy, -32768), 32767); tp.fx = fx; tp.fy = fy;
                  warp_fetch3(LdRO(), L.in, H, W, tp, px);
                  const bool bx0 = (unsigned)tp.sx < (unsigned)W, bx1 = fx != 0 && (unsigned)(tp.sx + 1) < (unsigned)W;
                  const bool by0 = (unsigned)tp.sy < (unsigned)H, by1 = fy != 0 && (unsigned)(tp.sy + 1) < (unsigned)H;
                  const uint8_t* r0 = mu + (size_t)tp.sy * W + tp.sx;
                  const uint8_t* r1 = r0 + W;
                  wm = bilerp_fix((by0 && bx0) ? ldb(r0) : 0, (by0 && bx1) ? ldb(r0 + 1) : 0, (by1 && bx0) ? ldb(r1) : 0,
                                  (by1 && bx1) ? ldb(r1 + 1) : 0, fx, fy);
                }
                const float M = Mv[i];
                if (M != 0.f || wm != 0) {  // keep == 0 => 0*img + 1*aug == aug exactly
                  const double am = S.div255[wm];
                  const double keep = (double)M > am ? (double)M : am;
                  const double rest = dsub(1.0, keep);
#pragma unroll
                  for (int c = 0; c < 3; ++c)
                    px[c] = (int)dadd(dmul(keep, (double)byte_of(in_w, 3 * i + c)), dmul(rest, (double)px[c]));
                }
              }
#pragma unroll
              for (int c = 0; c < 3; ++c) out_w[(3 * i + c) >> 2] |= (uint32_t)px[c] << (((3 * i + c) & 3) * 8);
            }
          }
          if (n == 4) {
            uint32_t* q = reinterpret_cast<uint32_t*>(L.out + o);
            q[0] = out_w[0]; q[1] = out_w[1]; q[2] = out_w[2];
          } else {
            for (int k = 0; k < 3 * n; ++k) L.out[o + k] = (uint8_t)byte_of(out_w, k);
          }
        }
        worker_sync();   // the stage (and its rectangle record) may be refilled
        ++n_stage;
        s = nxt;
      }
    }
  }
  // table-lookup / bbo-copy sub-tiles: 16-pixel runs as vectors
  const bool vec = ((W * 3) & 15) == 0 &&
                   ((((uintptr_t)L.in) | ((uintptr_t)L.out) | ((uintptr_t)A.scratch) | A.frame_bytes) & 15) == 0;
  for (int x = x0 + (t & 15) * kChunkPx; x < x1; x += 16 * kChunkPx) {
    const int cls = S.sub_cls[(x - x0) / kSubW];
    if ((cls & 15) != kSubStream) continue;
    const int reg = cls >> 4;
    const int n = min(kChunkPx, x1 - x);
    for (int y = y0 + (t >> 4); y < y1; y += 16) {
      if (vec && n == kChunkPx) {
        Chunk c;
        chunk_load(stream_src(L, reg, A.scratch, A.frame_bytes) + ((size_t)y * W + x) * 3, n, vec, c);
        stream_chunk(L, reg, S.lut + reg * 768, A.scratch, A.frame_bytes, c, x, y, n, vec);
      } else {
        for (int i = 0; i < n; ++i) pixel_op_fast(A, L, S.rop[reg], S.lut, reg, x + i, y);
      }
    }
  }
  // everything else per pixel, consecutive lanes on consecutive pixels
  if (any_pixel) {
    for (int s = 0; s < nsub; ++s) {
      if ((S.sub_cls[s] & 15) != kSubPixel) continue;
      const int sx0 = x0 + s * kSubW;
      for (int q = t; q < kSubW * kStepTileH; q += kCT) {
        const int x = sx0 + (q & (kSubW - 1)), y = y0 + q / kSubW;
        if (x >= x1 || y >= y1) continue;
        const int r = region_of_pixel(L, x, y);
        if (S.rop[r].kind == OADG_OP_BG_AFFINE)
          bg_pixel_fast(A.P, L, S.rop[r], cv_round(dmul(dmul(S.rop[r].minv[0], (double)x), 1024.0)),
                        cv_round(dmul(dmul(S.rop[r].minv[3], (double)x), 1024.0)), S.div255, x, y);
        else
          pixel_op_fast(A, L, S.rop[r], S.lut, r, x, y);
      }
    }
  }
  return n_stage;
}
// stage a lane record, its LUTs and its region ops in shared memory (once per lane a CTA works on)
__device__ OADG_HANDLER void stage_lane(const ChainArgs& A, ChainSmem& S, int lane) {
  const int t = threadIdx.x;
  worker_sync();
  if (t < (int)(sizeof(Lane) / 4))
    reinterpret_cast<uint32_t*>(&S.lane)[t] = reinterpret_cast<const uint32_t*>(A.lanes + lane)[t];
  worker_sync();
  const Lane& L = S.lane;
#pragma unroll
  for (int r = 0; r < OADG_MAX_REGIONS; ++r) {
    if (r <= L.n_ml && L.lut[r] >= 0 && t < 192)
      reinterpret_cast<uint32_t*>(S.lut + r * 768)[t] =
          reinterpret_cast<const uint32_t*>(A.luts + (size_t)L.lut[r] * 768)[t];
  }
  if (t <= L.n_ml) {
    const oadg_op_t& op = A.P.ops[L.op_base + t];
    S.rop[t].kind = op.kind;
    S.rop[t].p0 = op.p0;
    S.rop[t].p1 = op.p1;
    S.rop[t].factor = op.factor;
#pragma unroll
    for (int i = 0; i < 6; ++i) S.rop[t].minv[i] = op.minv[i];
  }
  worker_sync();
}

// kStats: per-kind CTA time accounting with %globaltimer and global atomics (oadg_oamix_execute_profiled only; the
// production instantiation carries none of it)
//
// Warps 0-7 (256 threads) are the WORKERS: they run the tile handlers.  Warp 8 is the SCHEDULER: while the workers
// process tile n it claims tile n + 1 (ticket + item record into the other slot) and publishes tile n - 1 (release
// fence, dependency counters, tickets of the successors), so none of the queue's memory round trips sits between two
// tiles of the workers.  Hand-over is two named barriers per slot.
template <bool kStats>
__global__ void __launch_bounds__(kCtaThreads, kCtaPerSm)
oamix_chain_kernel(const ChainArgs Aparam) {
  __shared__ ChainSmem S;
  extern __shared__ __align__(128) uint8_t dyn[];   // kDynSmem bytes, see the layout above
  if (threadIdx.x == 0) {
    S.args = Aparam;
    S.bs_key = -1;
    S.done_seq = 0u;
    tma::mbar_init(&S.full[0], 1);
    tma::mbar_init(&S.full[1], 1);
    tma::mbar_fence_init();
  }
  if (threadIdx.x < 256) S.div255[threadIdx.x] = (double)threadIdx.x / 255.0;
  __syncthreads();
  const ChainArgs& A = S.args;
  // tickets of the items that start ready (host-assigned ranges, priority order)
  for (int k = blockIdx.x; k < (int)A.ring[2]; k += gridDim.x) {
    const unsigned long long e = A.ring[kRingHeader + k];
    const int it0 = (int)(unsigned)e;
    const unsigned long long first = e >> 32;
    const int nt = A.items[it0].ntiles;
    for (int t = threadIdx.x; t < nt; t += kCtaThreads)
      st_release_u64(A.tickets + first + t, ((unsigned long long)(unsigned)(it0 + 1) << 32) | (unsigned)t);
  }
  if (threadIdx.x >= kCT) {
    // ---------------------------------------------------------------- scheduler warp
    // Iteration n hands tile n to the workers.  A finished tile is published as soon as the scheduler sees it --
    // in particular while it waits for an unpublished ticket: the work it waits for may depend on that very tile.
    const int lane = threadIdx.x & 31;
    const int ahead_q = (A.debug >> 8) ? (A.debug >> 8) : 8;   // claim ahead when more than ahead_q / 4 tickets per CTA are unclaimed
    unsigned published = 0;
    auto publish_ready = [&]() {   // warp-uniform: publish the tiles the workers have finished; returns how many
      unsigned done = 0;
      if (lane == 0) done = ld_acquire_cta_u32(&S.done_seq);
      done = __shfl_sync(0xffffffffu, done, 0);
      unsigned cnt = 0;
      while (published < done) {
        publish_tile(A, S.slot[published & 1u].I);
        ++published;
        ++cnt;
      }
      return cnt;
    };
    for (unsigned n = 0;; ++n) {
      const int k = (int)(n & 1u);
      while (n >= 2 && published < n - 1) {   // slot k still holds tile n - 2: the workers are finishing it
        if (!publish_ready()) __nanosleep(400);
      }
      // Claiming AHEAD (while the workers are still on tile n - 1) hides the claim's round trips, but a ticket held
      // by a busy CTA is a tile nobody runs: when published tickets are scarce (fewer than two per CTA unclaimed),
      // the claim waits until the workers are done -- idle CTAs then find the tile at once.
      if (n >= 1) {
        for (;;) {
          int go = 1;
          if (lane == 0) {
            const long long avail = (long long)ld_relaxed_u64(A.ring) - (long long)ld_relaxed_u64(A.ring + 1);
            go = avail > ((long long)A.grid * ahead_q) / 4 || ld_acquire_cta_u32(&S.done_seq) >= n;
          }
          if (__shfl_sync(0xffffffffu, go, 0)) break;
          if (!publish_ready()) __nanosleep(300);
        }
      }
      int item = -1, tile = -1;
      unsigned long long T = 0, t_enter = 0;
      if (kStats && lane == 0) t_enter = globaltimer_ns();
      if (lane == 0) T = atomicAdd(A.ring + 1, 1ull);
      T = __shfl_sync(0xffffffffu, T, 0);
      if (T < (unsigned long long)A.n_tiles) {
        unsigned long long spin_t0 = 0;
        for (;;) {
          unsigned long long e = 0;
          if (lane == 0) e = ld_acquire_u64(A.tickets + T);   // acquire: the item's inputs are complete
          e = __shfl_sync(0xffffffffu, e, 0);
          if (e != 0ull) {
            item = (int)(unsigned)(e >> 32) - 1;
            tile = (int)(unsigned)e;
            break;
          }
          if (publish_ready()) continue;
          __nanosleep(200);   // the ticket is not published yet
          bool give_up = false;
          if (lane == 0) {
            if (kStats) atomicAdd(A.kind_ns + 16 + 10, 1ull);   // measurement aid: polls that found the ticket unpublished
            const unsigned long long now = globaltimer_ns();
            if (spin_t0 == 0) spin_t0 = now;
            else if (now - spin_t0 > 2000000000ull) {   // safety valve: sticky fault flag instead of a hung GPU
              atomicAdd(A.fault, 1u);
              give_up = true;
            }
          }
          if (__shfl_sync(0xffffffffu, (int)give_up, 0)) break;
        }
        if (kStats && lane == 0 && item >= 0) {   // slot 11 = claims whose ticket was there, slot 12 = the others
          const int kk = spin_t0 ? 12 : 11;
          atomicAdd(A.kind_ns + kk, globaltimer_ns() - t_enter);
          atomicAdd(A.kind_ns + 16 + kk, 1ull);
        }
      }
      if (item >= 0) {
        constexpr int kItemWords = (int)(sizeof(Item) / 4);
        uint32_t w = 0;
        if (lane < kItemWords) w = reinterpret_cast<const uint32_t*>(A.items + item)[lane];
        if (lane < kItemWords) reinterpret_cast<uint32_t*>(&S.slot[k].I)[lane] = w;
        const int perm_first = __shfl_sync(0xffffffffu, (int)w, (int)(offsetof(Item, perm_first) / 4));
        if (lane == 0) S.slot[k].tile = perm_first >= 0 ? A.perm[perm_first + tile] : tile;
      }
      if (lane == 0) S.slot[k].item = item;
      __syncwarp();
      bar_arrive_all(2 + k);          // the workers may start tile n
      if (item < 0) {                 // drained: publish what the workers still hold, then leave
        while (published < n)
          if (!publish_ready()) __nanosleep(200);
        break;
      }
      publish_ready();
    }
    return;
  }
  // ------------------------------------------------------------------ workers
  int staged_lane = -1;
  unsigned n_stage = 0;   // gather stages used so far by this CTA (stage = n & 1, mbarrier parity = (n >> 1) & 1)
  unsigned long long idle_t0 = 0;
  if (kStats) idle_t0 = globaltimer_ns();
  for (unsigned n = 0;; ++n) {
    const int k = (int)(n & 1u);
    bar_sync_all(2 + k);
    const int it = S.slot[k].item;
    if (it < 0) break;
    const Item I = S.slot[k].I;
    const int local = S.slot[k].tile;
    unsigned long long seg_t0 = 0;
    if (kStats) seg_t0 = globaltimer_ns();
    switch (I.kind) {
      case OADG_IT_PROFILE: profile_tile(A, S, dyn, I.obj); break;
      case OADG_IT_MASK: mask_tile(A, S, dyn, I.obj, local, I.tx); break;
      case OADG_IT_HIST: {
        const Lane& L = A.lanes[I.obj];
        unsigned long long lsum = 0;
        unsigned* hist = reinterpret_cast<unsigned*>(dyn);
        hist_begin(hist);
        hist_tile(L, hist, local, lsum);
        hist_flush(A, hist, L.hist_slot, lsum);
        break;
      }
      case OADG_IT_LUT: lut_tile(A, S, I.obj); break;
      case OADG_IT_COPY: {
        const Chain& C = A.chains[I.obj];
        const oadg_view_t& V = A.P.views[C.view];
        copy_segment(C, (size_t)V.H * V.W * 3, I.aux != 0, local, local + 1);
        break;
      }
      case OADG_IT_BBO_R: n_stage = bbo_r_tile(A, S, dyn, I, local, n_stage); break;
      case OADG_IT_BBO_C: bbo_c_segment(A, S, I, local, local + 1); break;
      case OADG_IT_STEP:
        if (staged_lane != I.obj) {
          stage_lane(A, S, I.obj);
          staged_lane = I.obj;
        }
        n_stage = step_tile(A, S, dyn, local, I.tx, I.aux, n_stage);
        break;
      default: break;
    }
    if (kStats && threadIdx.x == 0) {
      const unsigned long long t1 = globaltimer_ns();
      const int kk = I.kind == OADG_IT_STEP ? S.step_class : I.kind;   // 7 stream, 8 bg staged, 9 mixed / per pixel
      const unsigned long long dt = t1 - seg_t0;
      atomicAdd(A.kind_ns + kk, dt);
      atomicAdd(A.kind_ns + 16 + kk, 1ull);
      atomicMax(A.kind_ns + 32 + kk, dt);
      atomicAdd(A.kind_ns + 10, seg_t0 - idle_t0);   // slot 10: the workers waited for the scheduler's next tile
      atomicMax(A.item_ts + 2 * it, (1ull << 63) - seg_t0);
      atomicMax(A.item_ts + 2 * it + 1, t1);
      idle_t0 = t1;
    }
    worker_sync();   // every worker's stores of the tile are issued ...
    if (threadIdx.x == 0) st_release_cta_u32(&S.done_seq, n + 1u);   // ... finished: the scheduler publishes the tile
    if (kStats && threadIdx.x == 0) {   // slot 13: thread 0 waited for the CTA's slower warps at the end of the tile
      const unsigned long long t2 = globaltimer_ns();
      atomicAdd(A.kind_ns + 13, t2 - idle_t0);
      atomicAdd(A.kind_ns + 16 + 13, 1ull);
    }
  }
}

// branch mixing + object-aware mixing (oa_mix.py:236,281-309); grid = (tiles x, tiles y, views).
// A thread owns groups of 4 pixels (12 bytes = three 32-bit words of each of the up to five frames it reads); the
// object-aware targets that can be non-zero in the tile are listed once per tile (classify_mix_tile), most tiles
// have none.  Same per-pixel arithmetic as mix_chunk / mix_pixel (oamix_tile.h, oamix_body.h), which serve frames
// whose rows are not 4-byte periodic and the host arithmetic check.  kFused: the opt-in Normalize + Pad + CHW
// float32 epilogue (oadg_fused_out_t).
constexpr int kTileThreads = 256;
template <bool kFused>
__global__ void __launch_bounds__(kTileThreads, 4)
mix_kernel(DevPlan P, const MixJob* jobs) {
  __shared__ MixTile T;
  __shared__ float s_lut[kFused ? 768 : 1];
  __shared__ float s_keep[256];   // float((1 - m) * v): the float64 term of mix_finish for a pixel outside every target
  const MixJob J = jobs[blockIdx.z];
  const oadg_view_t& V = P.views[J.view];
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  if (x0 >= V.W || y0 >= V.H) return;
  const int x1 = min(x0 + kTileW, V.W), y1 = min(y0 + kTileH, V.H);
  const int t = threadIdx.x;
  if (t == 0) classify_mix_tile(P, J, x0, y0, x1, y1, T);
  if (kFused)
    for (int i = t; i < 768; i += kTileThreads) s_lut[i] = P.norm_lut[i];
  s_keep[t] = (float)dmul(dsub(1.0, V.m), (double)t);   // kTileThreads == 256
  __syncthreads();
  const int width = V.width;
  uintptr_t al = ((uintptr_t)J.src) | ((uintptr_t)J.out);
  for (int b = 0; b < width; ++b) al |= (uintptr_t)J.branch[b];
  const bool fast = ((V.W * 3) & 3) == 0 && (al & 3) == 0 && !T.overflow && width <= 4;
  if (!fast) {   // odd row pitch / many targets / wide mixtures: the per-pixel body
    for (int q = t; q < kTileW * kTileH; q += kTileThreads) {
      const int x = x0 + (q & (kTileW - 1)), y = y0 + q / kTileW;
      if (x < x1 && y < y1) mix_pixel(P, J, x, y);
    }
    return;
  }
  const float w0 = V.ws[0], w1 = V.ws[1], w2 = V.ws[2], w3 = V.ws[3];
  const double m = V.m;
  const float mf = (float)m;
  const int ntgt = T.n;
  const int gw = (x1 - x0 + 3) >> 2;                 // 4-pixel groups per tile row
  const size_t plane = kFused ? (size_t)J.Hp * J.Wp : 0;
#pragma unroll 1
  for (int q = t; q < gw * (y1 - y0); q += kTileThreads) {
    const int y = y0 + q / gw, x = x0 + (q % gw) * 4;
    const int n = min(4, x1 - x);
    const size_t o = ((size_t)y * V.W + x) * 3;
    uint32_t sw[3], bw[4][3], ow[3] = {0u, 0u, 0u};
    if (n == 4) {
      const uint32_t* ps = reinterpret_cast<const uint32_t*>(J.src + o);
      sw[0] = ps[0]; sw[1] = ps[1]; sw[2] = ps[2];
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (b < width) {
          const uint32_t* pb = reinterpret_cast<const uint32_t*>(J.branch[b] + o);
          bw[b][0] = pb[0]; bw[b][1] = pb[1]; bw[b][2] = pb[2];
        }
    } else {   // ragged right edge
      sw[0] = sw[1] = sw[2] = 0u;
      for (int k = 0; k < 3 * n; ++k) sw[k >> 2] |= (uint32_t)J.src[o + k] << ((k & 3) * 8);
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (b < width) {
          bw[b][0] = bw[b][1] = bw[b][2] = 0u;
          for (int k = 0; k < 3 * n; ++k) bw[b][k >> 2] |= (uint32_t)J.branch[b][o + k] << ((k & 3) * 8);
        }
    }
    if (ntgt == 0) {
      // no object-aware target reaches this tile: orig = aug = 0 and mask_sum = 0, so (mix_finish)
      //   out = clip(float((1 - m) * img) + float(m) * acc), with the float64 product tabulated per uint8 level
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const int wi = k >> 2, sel = 0x7650 + (k & 3);   // byte k of a frame word into an exact float: 2^23 + v
        float a = fadd(0.f, fmul(w0, fsub(__int_as_float(__byte_perm(bw[0][wi], 0x4B000000u, sel)), 8388608.0f)));
        if (width > 1) a = fadd(a, fmul(w1, fsub(__int_as_float(__byte_perm(bw[1][wi], 0x4B000000u, sel)), 8388608.0f)));
        if (width > 2) a = fadd(a, fmul(w2, fsub(__int_as_float(__byte_perm(bw[2][wi], 0x4B000000u, sel)), 8388608.0f)));
        if (width > 3) a = fadd(a, fmul(w3, fsub(__int_as_float(__byte_perm(bw[3][wi], 0x4B000000u, sel)), 8388608.0f)));
        float out = fadd(s_keep[__byte_perm(sw[wi], 0u, 0x4440 + (k & 3))], fmul(mf, a));
        if (!(out > 0.0f)) out = 0.0f;
        if (out > 255.0f) out = 255.0f;
        ow[wi] |= (uint32_t)f32_trunc_u8(out) << ((k & 3) * 8);
      }
    } else
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float acc[3];
      int img[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int k = 3 * i + c;
        float a = fadd(0.f, fmul(w0, u8_to_f32(byte_of(bw[0], k))));
        if (width > 1) a = fadd(a, fmul(w1, u8_to_f32(byte_of(bw[1], k))));
        if (width > 2) a = fadd(a, fmul(w2, u8_to_f32(byte_of(bw[2], k))));
        if (width > 3) a = fadd(a, fmul(w3, u8_to_f32(byte_of(bw[3], k))));
        acc[c] = a;
        img[c] = byte_of(sw, k);
      }
      float orig[3] = {0.f, 0.f, 0.f}, aug[3] = {0.f, 0.f, 0.f};
      MixMask ms = {0.f, 0.f};
      if (ntgt > 0 && i < n) {
        for (int tg = 0; tg < ntgt; ++tg) {
          const oadg_target_t& G = P.tgts[T.idx[tg]];
          float mask;
          if (G.kind == 0) mask = fg_mask(P, G.gt, x + i, y);
          else mask = (x + i >= G.box[0] && x + i < G.box[2] && y >= G.box[1] && y < G.box[3]) ? 1.f : 0.f;
          if (mask == 0.f) continue;   // exact: weight 0 adds +0 and leaves sum == max
          const float w = mix_target_weight(ms, mask);
#pragma unroll
          for (int c = 0; c < 3; ++c) mix_accumulate(orig[c], aug[c], G.m_oa, img[c], acc[c], w);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int k = 3 * i + c;
        ow[k >> 2] |= (uint32_t)mix_finish(orig[c], aug[c], m, img[c], acc[c], ms.sum) << ((k & 3) * 8);
      }
    }
    if (n == 4) {
      uint32_t* po = reinterpret_cast<uint32_t*>(J.out + o);
      po[0] = ow[0]; po[1] = ow[1]; po[2] = ow[2];
    } else {
      for (int k = 0; k < 3 * n; ++k) J.out[o + k] = (uint8_t)byte_of(ow, k);
    }
    if (kFused) {
      const size_t at = (size_t)y * J.Wp + x;
      const bool v4 = n == 4 && ((J.Wp | x) & 3) == 0 && (plane & 3) == 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int c = P.norm_rgb ? 2 - k : k;
        const float* lut = s_lut + k * 256;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          float* dst = which == 0 ? J.f32_out : J.f32_src;
          if (!dst) continue;
          const uint32_t* px = which == 0 ? ow : sw;
          float* row = dst + k * plane + at;
          if (v4 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            *reinterpret_cast<float4*>(row) = make_float4(lut[byte_of(px, c)], lut[byte_of(px, 3 + c)],
                                                          lut[byte_of(px, 6 + c)], lut[byte_of(px, 9 + c)]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < n) row[i] = lut[byte_of(px, 3 * i + c)];
          }
        }
      }
    }
  }
}
#define BE_TRY(expr)                       \
  do {                                     \
    cudaError_t _e = (expr);               \
    if (_e != cudaSuccess) return (int)_e; \
  } while (0)

// Plan + launch tables go up through a small process-wide ring of page-locked buffers: a copy from pageable memory
// may make the host wait for the stream's earlier kernels, which would serialise the loader loop with the GPU, and
// pinning memory costs milliseconds, so the buffers outlive the calling thread (a loader starts a new worker thread
// per epoch).  A slot also receives the launch's fault flag (a CTA gave up waiting for work: the dependency tables
// were inconsistent); it is reused once everything that touches it has finished (event), and a raised flag is
// reported by the next call that looks at the slot (or by oadg_oamix_poll_fault).  A slot is owned (mutex) by the
// call that fills it until its launch is enqueued.
struct PinSlot {
  std::mutex mu;
  void* p = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;
  unsigned* fault = nullptr;
  int dev = -1;
  bool busy = false;
};
constexpr int kPinSlots = 8;
PinSlot g_pin[kPinSlots];
std::atomic<unsigned> g_pin_next{0};

// collect the fault flags of finished launches (wait = true: of all launches so far); OADG_E_PLAN if any
int poll_faults(bool wait) {
  int rc = 0;
  for (PinSlot& sl : g_pin) {
    std::unique_lock<std::mutex> lk(sl.mu, std::try_to_lock);
    if (!lk.owns_lock()) {
      if (!wait) continue;
      lk.lock();
    }
    if (!sl.busy || !sl.ev) continue;
    if (wait) {
      BE_TRY(cudaEventSynchronize(sl.ev));
    } else {
      const cudaError_t q = cudaEventQuery(sl.ev);
      if (q == cudaErrorNotReady) continue;
      if (q != cudaSuccess) return (int)q;
    }
    sl.busy = false;
    if (sl.fault && *sl.fault) {
      *sl.fault = 0;
      rc = OADG_E_PLAN;
    }
  }
  return rc;
}

struct CudaBackend {
  cudaStream_t stream;
  int launches = 0;
  int n_sm = 0, ctas_per_sm = 0;
  // optional CUDA-event timing of the two launches (oadg_oamix_execute_profiled)
  bool profile = false;
  bool fused = false;   // the mix writes the Normalize + Pad + CHW epilogue (oadg_oamix_execute_fused)
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  int n_items = 0, n_tiles = 0;
  const unsigned long long* kind_ns_dev = nullptr;
  const unsigned long long* item_ts_dev = nullptr;
  const unsigned* fault_dev = nullptr;
  std::vector<Item> items_host;
  std::vector<int32_t> deps_host;
  PinSlot* slot = nullptr;

  int want_ctas = 0;   // oadg_oamix_execute_shared: resident CTAs per SM this launch may take (0 = all that fit)
  int grid() {
    // per device, queried once per process (one process drives one GPU; the tables cover the multi-device case)
    static int sm_of_device[64] = {0}, ctas_of_device[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    const bool cacheable = dev >= 0 && dev < 64;
    if (cacheable && ctas_of_device[dev] > 0) {
      n_sm = sm_of_device[dev];
      ctas_per_sm = ctas_of_device[dev];
    } else {
      if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
      int nb = 0, nb2 = 0;
      if (cudaFuncSetAttribute(oamix_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem) != cudaSuccess) return -1;
      if (cudaFuncSetAttribute(oamix_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem) != cudaSuccess) return -1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, oamix_chain_kernel<false>, kCtaThreads, kDynSmem) != cudaSuccess) return -1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, oamix_chain_kernel<true>, kCtaThreads, kDynSmem) != cudaSuccess) return -1;
      nb = nb2 < nb ? nb2 : nb;
      ctas_per_sm = nb < 1 ? 0 : (nb > kCtaPerSm ? kCtaPerSm : nb);
      if (const char* e = getenv("OADG_CTAS_PER_SM")) {   // experiments
        const int want = atoi(e);
        if (want >= 1 && want < ctas_per_sm) ctas_per_sm = want;
      }
      if (ctas_per_sm == 0) return -1;
      if (cacheable) {
        sm_of_device[dev] = n_sm;
        ctas_of_device[dev] = ctas_per_sm;
      }
    }
    const int c = (want_ctas >= 1 && want_ctas < ctas_per_sm) ? want_ctas : ctas_per_sm;
    return n_sm * c;
  }
  // cuTensorMapEncodeTiled costs microseconds and the same frames (workspace, a loader's frame pool) come back launch
  // after launch: encoded maps are kept per calling thread, keyed by everything that goes into them
  struct MapKey {
    const void* base;
    size_t inner;
    int rows, box;
    bool operator==(const MapKey& o) const { return base == o.base && inner == o.inner && rows == o.rows && box == o.box; }
  };
  struct MapEntry {
    MapKey key;
    bool valid = false;
    alignas(64) unsigned char map[kTensorMapBytes];
  };
  int make_map(void* dst, const void* base, size_t inner_bytes, int rows, int box_inner) {
    static const bool off = getenv("OADG_NO_TMA") != nullptr;
    if (off) return -1;
    constexpr int kSlots = 1024;
    static thread_local std::vector<MapEntry> cache(kSlots);
    const MapKey key{base, inner_bytes, rows, box_inner};
    const size_t h = ((reinterpret_cast<uintptr_t>(base) >> 8) * 0x9E3779B97F4A7C15ull + inner_bytes * 31 + (size_t)rows * 7 + box_inner) % kSlots;
    MapEntry& e = cache[h];
    if (!(e.valid && e.key == key)) {
      if (tma::encode_u8_2d(e.map, base, inner_bytes, (uint64_t)rows, inner_bytes, (uint32_t)box_inner, kGatherBoxRows) != 0) {
        e.valid = false;
        return -1;
      }
      e.key = key;
      e.valid = true;
    }
    memcpy(dst, e.map, kTensorMapBytes);
    return 0;
  }
  ~CudaBackend() {
    if (slot) slot->mu.unlock();
  }
  int upload(void* dst, const void* src, size_t bytes) {
    int rc = poll_faults(false);
    if (rc) return rc;
    {  // first call of the process: pin every slot now (milliseconds each), not one by one under a running loader
      static std::once_flag once;
      std::call_once(once, [] {
        for (PinSlot& s0 : g_pin) {
          std::lock_guard<std::mutex> lk(s0.mu);
          if (!s0.p && cudaHostAlloc(&s0.p, (size_t)2 << 20, cudaHostAllocPortable) == cudaSuccess) s0.cap = (size_t)2 << 20;
          if (!s0.fault && cudaHostAlloc((void**)&s0.fault, 64, cudaHostAllocPortable) == cudaSuccess) *s0.fault = 0;
        }
      });
    }
    PinSlot& sl = g_pin[g_pin_next.fetch_add(1u) % kPinSlots];
    sl.mu.lock();
    slot = &sl;   // released by the destructor, after the launch that uses the slot is enqueued
    int dev = 0;
    BE_TRY(cudaGetDevice(&dev));
    if (sl.ev && sl.dev != dev) {
      if (sl.busy) cudaEventSynchronize(sl.ev);
      cudaEventDestroy(sl.ev);
      sl.ev = nullptr;
      sl.busy = false;
    }
    if (sl.ev && sl.busy) {
      BE_TRY(cudaEventSynchronize(sl.ev));
      sl.busy = false;
      if (sl.fault && *sl.fault) {
        *sl.fault = 0;
        return OADG_E_PLAN;
      }
    }
    if (!sl.ev) BE_TRY(cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
    if (!sl.fault) {
      BE_TRY(cudaHostAlloc((void**)&sl.fault, 64, cudaHostAllocPortable));
      *sl.fault = 0;
    }
    sl.dev = dev;
    if (sl.cap < bytes) {
      if (sl.p) cudaFreeHost(sl.p);
      sl.p = nullptr;
      sl.cap = 0;
      const size_t cap = 2 * bytes > ((size_t)2 << 20) ? 2 * bytes : ((size_t)2 << 20);   // pinning costs milliseconds: be generous
      BE_TRY(cudaHostAlloc(&sl.p, cap, cudaHostAllocPortable));
      sl.cap = cap;
    }
    memcpy(sl.p, src, bytes);
    BE_TRY(cudaMemcpyAsync(dst, sl.p, bytes, cudaMemcpyHostToDevice, stream));
    BE_TRY(cudaEventRecord(sl.ev, stream));
    sl.busy = true;
    return 0;
  }
  int zero(void* dst, size_t bytes) {
    BE_TRY(cudaMemsetAsync(dst, 0, bytes, stream));
    return 0;
  }
  int zero2d(void* dst, size_t pitch, size_t width, size_t rows) {
    BE_TRY(cudaMemset2DAsync(dst, pitch, 0, width, rows, stream));
    return 0;
  }
  int chain(const ChainArgs& A, const ChainArgs& Hh, const PlanView&) {
    fault_dev = A.fault;
    if (profile) {
      for (auto& e : ev) BE_TRY(cudaEventCreate(&e));
      BE_TRY(cudaEventRecord(ev[0], stream));
      n_items = A.n_items;
      n_tiles = A.n_tiles;
      kind_ns_dev = A.kind_ns;
      item_ts_dev = A.item_ts;
      items_host.assign(Hh.items, Hh.items + A.n_items);
      int nd = 0;
      for (const Item& I : items_host) nd = I.dep_first + I.dep_count > nd ? I.dep_first + I.dep_count : nd;
      deps_host.assign(Hh.deps, Hh.deps + nd);
    }
    if (A.n_tiles > 0) {
      ChainArgs args = A;
      if (const char* dbg = getenv("OADG_DEBUG")) args.debug = atoi(dbg);
      void* params[1] = {(void*)&args};
      // cooperative launch: all CTAs are guaranteed co-resident, which the in-kernel dependency waits rely on
      const void* fn = profile ? (const void*)oamix_chain_kernel<true> : (const void*)oamix_chain_kernel<false>;
      BE_TRY(cudaLaunchCooperativeKernel(fn, dim3(A.grid), dim3(kCtaThreads), params, kDynSmem, stream));
      ++launches;
      if (slot) {   // the launch's fault flag travels to the upload slot; the slot's event now covers it
        BE_TRY(cudaMemcpyAsync(slot->fault, A.fault, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
        BE_TRY(cudaEventRecord(slot->ev, stream));
      }
    }
    if (profile) BE_TRY(cudaEventRecord(ev[1], stream));
    return 0;
  }
  int mix(const DevPlan& P, const MixJob* jobs, int n) {
    dim3 grid((P.max_w + kTileW - 1) / kTileW, (P.max_h + kTileH - 1) / kTileH, n);
    if (fused) mix_kernel<true><<<grid, kTileThreads, 0, stream>>>(P, jobs);
    else mix_kernel<false><<<grid, kTileThreads, 0, stream>>>(P, jobs);
    BE_TRY(cudaGetLastError());
    ++launches;
    if (profile) BE_TRY(cudaEventRecord(ev[2], stream));
    return 0;
  }
};

}  // namespace
}  // namespace oadg

using namespace oadg;

extern "C" int oadg_oamix_workspace_bytes(const void* plan_host, size_t plan_bytes, size_t* out_bytes) {
  if (!out_bytes) return OADG_E_ARG;
  PlanView pv;
  int rc = parse_plan(plan_host, plan_bytes, pv);
  if (rc) return rc;
  Layout L;
  make_layout(pv, L);
  *out_bytes = L.total;
  return 0;
}

extern "C" int oadg_oamix_poll_fault(int wait) { return poll_faults(wait != 0); }

// measurement aid: the work queue of the last profiled execution on this workspace (kind, obj, tiles, first claim and
// last publish in ns relative to the earliest claim, dependencies) -- filled by oadg_oamix_execute_profiled
struct ItemTrace {
  int32_t kind, obj, ntiles, dep_count;
  double t0_us, t1_us;
  int32_t deps[8];
};
static std::vector<ItemTrace> g_last_trace;   // debugging only (not thread safe)
extern "C" int oadg_oamix_last_trace(void* out, int cap) {
  const int n = (int)g_last_trace.size() < cap ? (int)g_last_trace.size() : cap;
  if (out && n > 0) memcpy(out, g_last_trace.data(), (size_t)n * sizeof(ItemTrace));
  return (int)g_last_trace.size();
}

extern "C" int oadg_oamix_execute_profiled(const void* plan_host, size_t plan_bytes, const uint8_t* const* src_dev,
                                           int n_img, uint8_t* const* dst_dev, void* workspace_dev,
                                           size_t workspace_bytes, float* ms_chain, float* ms_mix, int* n_items_out,
                                           int* n_tiles_out, unsigned long long* kind_stats, void* stream) {
  if (!ms_chain || !ms_mix) return OADG_E_ARG;
  CudaBackend be;
  be.stream = (cudaStream_t)stream;
  be.profile = true;
  int rc = execute_plan(be, plan_host, plan_bytes, src_dev, n_img, dst_dev, workspace_dev, workspace_bytes);
  cudaError_t e = cudaStreamSynchronize(be.stream);
  *ms_chain = *ms_mix = 0.f;
  if (n_items_out) *n_items_out = be.n_items;
  if (n_tiles_out) *n_tiles_out = be.n_tiles;
  if (rc == 0 && e == cudaSuccess && be.ev[2]) {
    cudaEventElapsedTime(ms_chain, be.ev[0], be.ev[1]);
    cudaEventElapsedTime(ms_mix, be.ev[1], be.ev[2]);
    unsigned long long ks[48] = {0};
    if (be.kind_ns_dev) e = cudaMemcpy(ks, be.kind_ns_dev, sizeof(ks), cudaMemcpyDeviceToHost);
    if (kind_stats) memcpy(kind_stats, ks, sizeof(ks));
    unsigned fault = 0;
    if (be.fault_dev) e = cudaMemcpy(&fault, be.fault_dev, sizeof(fault), cudaMemcpyDeviceToHost);
    if (fault != 0) rc = OADG_E_PLAN;   // a CTA gave up waiting: the dependency tables were inconsistent
    if (getenv("OADG_TRACE") && be.item_ts_dev && be.n_items > 0) {
      std::vector<unsigned long long> ts((size_t)be.n_items * 2);
      cudaMemcpy(ts.data(), be.item_ts_dev, ts.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      unsigned long long base = ~0ull;
      for (int k = 0; k < be.n_items; ++k) {
        const unsigned long long t0 = (1ull << 63) - ts[2 * k];
        if (ts[2 * k] && t0 < base) base = t0;
      }
      g_last_trace.assign(be.n_items, ItemTrace{});
      for (int k = 0; k < be.n_items; ++k) {
        ItemTrace& T = g_last_trace[k];
        const Item& I = be.items_host[k];
        T.kind = I.kind; T.obj = I.obj; T.ntiles = I.ntiles; T.dep_count = I.dep_count;
        T.t0_us = ts[2 * k] ? (double)((1ull << 63) - ts[2 * k] - base) * 1e-3 : -1.0;
        T.t1_us = ts[2 * k + 1] ? (double)(ts[2 * k + 1] - base) * 1e-3 : -1.0;
        for (int d = 0; d < 8; ++d) T.deps[d] = d < I.dep_count ? be.deps_host[I.dep_first + d] : -1;
      }
    }
  }
  for (auto& ev : be.ev)
    if (ev) cudaEventDestroy(ev);
  if (rc) return rc;
  return e == cudaSuccess ? 0 : (int)e;
}

extern "C" int oadg_oamix_execute_shared(const void* plan_host, size_t plan_bytes, const uint8_t* const* src_dev,
                                         int n_img, uint8_t* const* dst_dev, void* workspace_dev,
                                         size_t workspace_bytes, int ctas_per_sm, int* launches_out, void* stream) {
  CudaBackend be;
  be.stream = (cudaStream_t)stream;
  be.want_ctas = ctas_per_sm;
  int rc = execute_plan(be, plan_host, plan_bytes, src_dev, n_img, dst_dev, workspace_dev, workspace_bytes);
  if (launches_out) *launches_out = be.launches;
  return rc;
}

extern "C" int oadg_oamix_execute_fused(const void* plan_host, size_t plan_bytes, const uint8_t* const* src_dev,
                                        int n_img, uint8_t* const* dst_dev, const oadg_fused_out_t* fused,
                                        void* workspace_dev, size_t workspace_bytes, int* launches_out, void* stream) {
  if (!fused) return OADG_E_ARG;
  CudaBackend be;
  be.stream = (cudaStream_t)stream;
  be.fused = true;
  int rc = execute_plan(be, plan_host, plan_bytes, src_dev, n_img, dst_dev, workspace_dev, workspace_bytes, fused);
  if (launches_out) *launches_out = be.launches;
  return rc;
}

extern "C" int oadg_oamix_execute(const void* plan_host, size_t plan_bytes, const uint8_t* const* src_dev,
                                  int n_img, uint8_t* const* dst_dev, void* workspace_dev,
                                  size_t workspace_bytes, int* launches_out, void* stream) {
  CudaBackend be;
  be.stream = (cudaStream_t)stream;
  int rc = execute_plan(be, plan_host, plan_bytes, src_dev, n_img, dst_dev, workspace_dev, workspace_bytes);
  if (launches_out) *launches_out = be.launches;
  return rc;
}
