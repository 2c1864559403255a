// OA-Mix plan executor: the device side of OAMix.oamix (reference oa_mix.py:207-309).
//
// Data layout in HBM (all caller- or workspace-owned):
//   frames      u8 HWC, pitch 3*W, one per source image / branch ping-pong / bbo chain (S, T)
//   profiles    per gt box two float32 vectors ux[W], uy[H]; blurred mask(y,x) = uy[y]*ux[x]
//               (the reference materialises a 25 MB float HxWx3 mask per box, oa_mix.py:75-93)
//   hist / lut  per lane input 3x256 u32 histogram (+ luma sum), per LUT op 3x256 u8 table
//   plan        the host-sampled plan blob (oadg.h records) + work tables, one H2D copy
//
// Two launches per batch:
//   oamix_chain_kernel   ONE persistent launch (four independent 256-thread CTAs per SM) that drains the host-built
//                        work queue (oamix_exec.h): mask profiles, union masks, histograms, LUTs, the bboxes-only
//                        chains level by level and every depth step of every (view, branch) lane.  The queue is a
//                        list of work items cut into tiles, ordered so that dependencies come first; CTAs claim
//                        tiles with one atomic counter and an item starts as soon as the items it depends on are
//                        complete (per-item completion counters, no grid-wide barrier), so nothing returns to the
//                        host between the ~10-40 dependent stages.
//   mix_kernel           branch mixing + object-aware mixing of all views (oa_mix.py:236,281-309)
#include <stdlib.h>
#include <string.h>

#include "oadg_common.cuh"
#include "oamix_exec.h"
#include "oamix_tile.h"

namespace oadg {
namespace {

// i / 255 in float64 (the reference divides a uint8 array by the python int 255, bbox_augmentation.py:267)
__device__ const double g_div255[256] = {0.0 / 255.0, 1.0 / 255.0, 2.0 / 255.0, 3.0 / 255.0, 4.0 / 255.0, 5.0 / 255.0, 6.0 / 255.0, 7.0 / 255.0, 8.0 / 255.0, 9.0 / 255.0, 10.0 / 255.0, 11.0 / 255.0, 12.0 / 255.0, 13.0 / 255.0, 14.0 / 255.0, 15.0 / 255.0, 16.0 / 255.0, 17.0 / 255.0, 18.0 / 255.0, 19.0 / 255.0, 20.0 / 255.0, 21.0 / 255.0, 22.0 / 255.0, 23.0 / 255.0, 24.0 / 255.0, 25.0 / 255.0, 26.0 / 255.0, 27.0 / 255.0, 28.0 / 255.0, 29.0 / 255.0, 30.0 / 255.0, 31.0 / 255.0, 32.0 / 255.0, 33.0 / 255.0, 34.0 / 255.0, 35.0 / 255.0, 36.0 / 255.0, 37.0 / 255.0, 38.0 / 255.0, 39.0 / 255.0, 40.0 / 255.0, 41.0 / 255.0, 42.0 / 255.0, 43.0 / 255.0, 44.0 / 255.0, 45.0 / 255.0, 46.0 / 255.0, 47.0 / 255.0, 48.0 / 255.0, 49.0 / 255.0, 50.0 / 255.0, 51.0 / 255.0, 52.0 / 255.0, 53.0 / 255.0, 54.0 / 255.0, 55.0 / 255.0, 56.0 / 255.0, 57.0 / 255.0, 58.0 / 255.0, 59.0 / 255.0, 60.0 / 255.0, 61.0 / 255.0, 62.0 / 255.0, 63.0 / 255.0, 64.0 / 255.0, 65.0 / 255.0, 66.0 / 255.0, 67.0 / 255.0, 68.0 / 255.0, 69.0 / 255.0, 70.0 / 255.0, 71.0 / 255.0, 72.0 / 255.0, 73.0 / 255.0, 74.0 / 255.0, 75.0 / 255.0, 76.0 / 255.0, 77.0 / 255.0, 78.0 / 255.0, 79.0 / 255.0, 80.0 / 255.0, 81.0 / 255.0, 82.0 / 255.0, 83.0 / 255.0, 84.0 / 255.0, 85.0 / 255.0, 86.0 / 255.0, 87.0 / 255.0, 88.0 / 255.0, 89.0 / 255.0, 90.0 / 255.0, 91.0 / 255.0, 92.0 / 255.0, 93.0 / 255.0, 94.0 / 255.0, 95.0 / 255.0, 96.0 / 255.0, 97.0 / 255.0, 98.0 / 255.0, 99.0 / 255.0, 100.0 / 255.0, 101.0 / 255.0, 102.0 / 255.0, 103.0 / 255.0, 104.0 / 255.0, 105.0 / 255.0, 106.0 / 255.0, 107.0 / 255.0, 108.0 / 255.0, 109.0 / 255.0, 110.0 / 255.0, 111.0 / 255.0, 112.0 / 255.0, 113.0 / 255.0, 114.0 / 255.0, 115.0 / 255.0, 116.0 / 255.0, 117.0 / 255.0, 118.0 / 255.0, 119.0 / 255.0, 120.0 / 255.0, 121.0 / 255.0, 122.0 / 255.0, 123.0 / 255.0, 124.0 / 255.0, 125.0 / 255.0, 126.0 / 255.0, 127.0 / 255.0, 128.0 / 255.0, 129.0 / 255.0, 130.0 / 255.0, 131.0 / 255.0, 132.0 / 255.0, 133.0 / 255.0, 134.0 / 255.0, 135.0 / 255.0, 136.0 / 255.0, 137.0 / 255.0, 138.0 / 255.0, 139.0 / 255.0, 140.0 / 255.0, 141.0 / 255.0, 142.0 / 255.0, 143.0 / 255.0, 144.0 / 255.0, 145.0 / 255.0, 146.0 / 255.0, 147.0 / 255.0, 148.0 / 255.0, 149.0 / 255.0, 150.0 / 255.0, 151.0 / 255.0, 152.0 / 255.0, 153.0 / 255.0, 154.0 / 255.0, 155.0 / 255.0, 156.0 / 255.0, 157.0 / 255.0, 158.0 / 255.0, 159.0 / 255.0, 160.0 / 255.0, 161.0 / 255.0, 162.0 / 255.0, 163.0 / 255.0, 164.0 / 255.0, 165.0 / 255.0, 166.0 / 255.0, 167.0 / 255.0, 168.0 / 255.0, 169.0 / 255.0, 170.0 / 255.0, 171.0 / 255.0, 172.0 / 255.0, 173.0 / 255.0, 174.0 / 255.0, 175.0 / 255.0, 176.0 / 255.0, 177.0 / 255.0, 178.0 / 255.0, 179.0 / 255.0, 180.0 / 255.0, 181.0 / 255.0, 182.0 / 255.0, 183.0 / 255.0, 184.0 / 255.0, 185.0 / 255.0, 186.0 / 255.0, 187.0 / 255.0, 188.0 / 255.0, 189.0 / 255.0, 190.0 / 255.0, 191.0 / 255.0, 192.0 / 255.0, 193.0 / 255.0, 194.0 / 255.0, 195.0 / 255.0, 196.0 / 255.0, 197.0 / 255.0, 198.0 / 255.0, 199.0 / 255.0, 200.0 / 255.0, 201.0 / 255.0, 202.0 / 255.0, 203.0 / 255.0, 204.0 / 255.0, 205.0 / 255.0, 206.0 / 255.0, 207.0 / 255.0, 208.0 / 255.0, 209.0 / 255.0, 210.0 / 255.0, 211.0 / 255.0, 212.0 / 255.0, 213.0 / 255.0, 214.0 / 255.0, 215.0 / 255.0, 216.0 / 255.0, 217.0 / 255.0, 218.0 / 255.0, 219.0 / 255.0, 220.0 / 255.0, 221.0 / 255.0, 222.0 / 255.0, 223.0 / 255.0, 224.0 / 255.0, 225.0 / 255.0, 226.0 / 255.0, 227.0 / 255.0, 228.0 / 255.0, 229.0 / 255.0, 230.0 / 255.0, 231.0 / 255.0, 232.0 / 255.0, 233.0 / 255.0, 234.0 / 255.0, 235.0 / 255.0, 236.0 / 255.0, 237.0 / 255.0, 238.0 / 255.0, 239.0 / 255.0, 240.0 / 255.0, 241.0 / 255.0, 242.0 / 255.0, 243.0 / 255.0, 244.0 / 255.0, 245.0 / 255.0, 246.0 / 255.0, 247.0 / 255.0, 248.0 / 255.0, 249.0 / 255.0, 250.0 / 255.0, 251.0 / 255.0, 252.0 / 255.0, 253.0 / 255.0, 254.0 / 255.0, 255.0 / 255.0};

// The phase handlers are separate device functions: compiled into one monolithic kernel body, ptxas (12.9) produced
// wrong code for the mixed-tile path (caught by tests/test_gpu_oamix.py::test_single_op_plans_match_host_arithmetic).
#ifdef OADG_INLINE_HANDLERS
#define OADG_HANDLER __forceinline__
#else
#define OADG_HANDLER __noinline__
#endif

constexpr int kMaxQueueItems = 4096;   // work items per launch (a CTA keeps a bitmap of the exhausted ones)
constexpr int kCT = 256;     // threads per CTA of the chain kernel
constexpr int kCtaPerSm = 4; // independent CTAs per SM (64 registers per thread): tiles of different kinds overlap on an SM

struct RegOp {      // op parameters of one region, staged in shared memory for per-pixel tiles
  int32_t kind, p0, p1;
  float factor;
  double minv[6];
};

struct BboStage {    // one bbo job staged for the CTA (bbo_r_segment / bbo_c_segment)
  double minv[6];
  int32_t rect[4];
  int32_t W, H, gt, n_excl;
  const uint8_t* X;
  uint8_t* Y;
  int32_t excl[16][4];   // supports of the next level's boxes (catch-up exclusion)
};

struct ChainSmem {
  int next_tile;       // the CTA's next claimed tile ...
  int next_item;       // ... and the item it belongs to (-1: the queue is drained)
  unsigned epoch_seen; // value of the ready-epoch when this CTA last scanned the queue
  unsigned exhausted[kMaxQueueItems / 32];   // items this CTA knows to have no unclaimed tile left
  int rowoff[2][96];   // staged gathers: byte offset of every staged source row (frame rows, mask rows)
  int cand[16];        // mask tiles: the gt boxes whose support meets the tile
  int step_class;      // measurement aid: class of the last step tile (7 stream, 8 staged bg, 9 mixed / per pixel)
  int prof_ready;      // u.prof holds the profile slices of the staged blend job
  int bs_key;          // which bbo job `bs` (and, for blends, the profile slices in u.prof) currently holds; -1 = none
  ChainArgs args;      // the kernel arguments, copied once: the (non-inlined) handlers read them from shared memory
  BboStage bs;
  union {
    unsigned hist[2][768];   // histogram tiles: 2 privatised copies (4 warps share one)
    float prof[3072];        // bbo jobs: the support's slices of the two mask profiles
    struct {
      double K[1537];        // profile tiles: exclusive prefix sums of the gaussian kernel
      float p[1024];         //                low-res blurred profile
    } g;
  } u;
  __align__(16) uint8_t lut[OADG_MAX_REGIONS * 768];
  RegOp rop[OADG_MAX_REGIONS];
  Lane lane;
  double red[32];
  double ksum;
  uint8_t tab[3][256];
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}

// ---- the work queue ---------------------------------------------------------------------------------------
// Every item has a counter of claimed tiles and a counter of outstanding dependency tiles (`pending`, initialised by
// the host).  A CTA drains the item it is working on (one atomic per tile, issued before the tile is processed); when
// that item runs out it scans the queue from the front for the first item that is READY (pending == 0) and still has
// unclaimed tiles, skipping items whose inputs are not complete -- so a CTA only idles when nothing at all is ready.
// A finished tile is published with bar.sync + fence by one thread, which then decrements `pending` of the item's
// successors; a claimer reads `pending` with an acquire load, fences, and the CTA bar.syncs before touching the data
// (the pattern of a cooperative-groups grid sync, per item instead of per grid).  All CTAs are co-resident
// (cooperative launch), and a waiting CTA holds no tile, so the scheme cannot deadlock.
__device__ __forceinline__ int ld_acquire_s32(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// thread 0 only: find the next (item, tile); returns false when every item is exhausted
__device__ bool scan_for_work(const ChainArgs& A, unsigned* exhausted, int& scan_from, int& item, int& tile) {
  unsigned long long spin_t0 = 0;
  for (;;) {
    bool any_left = false;
    while (scan_from < A.n_items && (exhausted[scan_from >> 5] >> (scan_from & 31) & 1u)) ++scan_from;
    for (int k = scan_from; k < A.n_items; ++k) {
      if (exhausted[k >> 5] >> (k & 31) & 1u) continue;
      const unsigned nt = (unsigned)A.items[k].ntiles;
      const unsigned cl = ld_relaxed_u32(A.claimed + k);   // the two loads are independent: one L2 round trip
      const int pend = ld_acquire_s32(A.pending + k);
      if (cl >= nt) {
        exhausted[k >> 5] |= 1u << (k & 31);
        continue;
      }
      any_left = true;
      if (pend > 0) continue;   // inputs not complete yet: look further down the queue
      const unsigned t = atomicAdd(A.claimed + k, 1u);
      if (t < nt) {
        item = k;
        tile = (int)t;
        __threadfence();
        return true;
      }
      exhausted[k >> 5] |= 1u << (k & 31);
    }
    if (!any_left) return false;
    __nanosleep(256);   // nothing is ready: back off before polling the counters again
    // safety valve: a CTA that finds nothing ready for 2 s gives up (flag in stats slot 15) instead of hanging the GPU
    const unsigned long long now = globaltimer_ns();
    if (spin_t0 == 0) spin_t0 = now;
    else if (now - spin_t0 > 2000000000ull) {
      atomicAdd(A.kind_ns + 15, 1ull);
      return false;
    }
  }
}
__device__ __forceinline__ void publish_tile(const ChainArgs& A, const Item& I) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    for (int k = 0; k < I.succ_count; ++k) atomicSub(A.pending + A.succ[I.succ_first + k], 1);
  }
}

// ------------------------------------------------------------------------------------
// blurred-mask profile of one (gt box, axis) (oa_mix.py:78-91): indicator on the 1/sr canvas -> GaussianBlur
// (separable, BORDER_REFLECT_101, float32 kernel from getGaussianKernel) -> bilinear cv2.resize to full resolution.
// ------------------------------------------------------------------------------------
__device__ OADG_HANDLER void profile_tile(const ChainArgs& A, ChainSmem& S, int obj) {
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
  float* p = S.u.g.p;
  double* K = S.u.g.K;
  float* out = (axis == 0 ? A.prof_x + (size_t)g * P.max_w : A.prof_y + (size_t)g * P.max_h);
  const int tid = threadIdx.x;
  __syncthreads();  // the shared buffers may still be in use by the previous tile
  if (tid == 0) S.bs_key = -1;   // u.g overwrites the staged profile slices
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
    __syncthreads();
    if (tid == 0) {
      double t = 0;
      for (int w = 0; w < kCT / 32; ++w) t += S.red[w];
      S.ksum = 1.0 / t;
    }
    __syncthreads();
    const double ksum = S.ksum;
    // K[j] = sum of the float32 taps 0..j-1 in float64 (the blur of an indicator is a difference of prefix sums)
    for (int i = tid; i <= ks; i += kCT) {
      double x = (i - 1) - (ks - 1) * 0.5;
      K[i] = i == 0 ? 0.0 : (double)(float)(exp(s2 * x * x) * ksum);
    }
    __syncthreads();
    for (int off = 1; off <= ks; off <<= 1) {  // Hillis-Steele inclusive scan over K[0..ks]
      double v[8];
      int n = 0;
      for (int i = tid; i <= ks; i += kCT, ++n) v[n] = i >= off ? K[i] + K[i - off] : K[i];
      __syncthreads();
      n = 0;
      for (int i = tid; i <= ks; i += kCT, ++n) K[i] = v[n];
      __syncthreads();
    }
    const int r = ks / 2;
    auto range = [&](int a, int b) {  // sum of taps j in [a, b) clipped to [0, ks)
      a = max(a, 0);
      b = min(b, ks);
      return b > a ? K[b] - K[a] : 0.0;
    };
    if (r <= n_lo - 1) {
      // BORDER_REFLECT_101 with at most one reflection per side: tap j reads q = x + j - r, -q or 2(n_lo-1) - q
      for (int x = tid; x < n_lo; x += kCT) {
        const int s = r - x;  // j = q + s
        double acc = range(lo + s, hi + s);                                    // q in [lo, hi)
        acc += range(-hi + 1 + s, min(-lo, -1) + 1 + s);                       // q in [-hi+1, min(-lo,-1)]
        const int m2 = 2 * (n_lo - 1);
        acc += range(max(m2 - hi + 1, n_lo) + s, m2 - lo + 1 + s);             // q in [max(m2-hi+1,n_lo), m2-lo]
        p[x] = (float)acc;
      }
    } else {
      const int period = 2 * (n_lo - 1);
      for (int x = tid; x < n_lo; x += kCT) {
        double acc = 0.0;
        for (int j = 0; j < ks; ++j) {
          int q = x + j - r;
          if (n_lo == 1) q = 0;
          else {
            if (q < 0) q = -q;
            q %= period;
            if (q >= n_lo) q = period - q;
          }
          if (q >= lo && q < hi) acc += K[j + 1] - K[j];
        }
        p[x] = (float)acc;
      }
    }
  } else {
    for (int x = tid; x < n_lo; x += kCT) p[x] = (x >= lo && x < hi) ? 1.f : 0.f;
  }
  __syncthreads();
  // cv2.resize(f32, INTER_LINEAR): fx = (float)((dx+0.5)*scale - 0.5)
  const double scale = (double)n_lo / (double)n_hi;
  for (int d = tid; d < n_hi; d += kCT) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    float t = fsub(f, (float)s);
    if (s < 0) { s = 0; t = 0.f; }
    if (s >= n_lo - 1) { s = n_lo - 1; t = 0.f; }
    int s1 = min(s + 1, n_lo - 1);
    out[d] = fadd(fmul(p[s], fsub(1.f, t)), fmul(p[s1], t));
  }
}

// union of the blurred gt masks of a view (np.max(mask_bboxes, axis=0), bbox_augmentation.py:260) as float32 and as
// uint8(mask*255): written once per batch, read by every bg-only op.  Tile = 256 x 8 px, one column x 8 rows per
// thread; the boxes whose support meets the tile are listed once per tile, and a thread keeps the x-profile value of
// each listed box for its column in registers.
__device__ OADG_HANDLER void mask_tile(const ChainArgs& A, ChainSmem& S, int view, int local, int tx) {
  const DevPlan& P = A.P;
  const oadg_view_t& V = P.views[view];
  const int x0 = (local % tx) * kMaskTileW, y0 = (local / tx) * kMaskTileH;
  const int x1 = min(x0 + kMaskTileW, V.W), y1 = min(y0 + kMaskTileH, V.H);
  __syncthreads();
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
  __syncthreads();
  const int n = S.bs.n_excl;
  const int x = x0 + (threadIdx.x & 255);
  if (x >= x1) return;
  const size_t base = (size_t)view * P.mask_stride;
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
__device__ OADG_HANDLER void hist_tile(const Lane& L, ChainSmem& S, int local, unsigned long long& lsum) {
  const int tid = threadIdx.x;
  const size_t npx = (size_t)L.H * L.W;
  const size_t p0 = (size_t)local * kHistTilePx;
  const size_t p1 = p0 + kHistTilePx < npx ? p0 + kHistTilePx : npx;
  unsigned* my = S.u.hist[(tid >> 5) & 1];
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
__device__ OADG_HANDLER void hist_begin(ChainSmem& S) {
  __syncthreads();
  if (threadIdx.x == 0) S.bs_key = -1;   // u.hist overwrites the staged profile slices
  for (int i = threadIdx.x; i < 2 * 768; i += kCT) (&S.u.hist[0][0])[i] = 0;
  __syncthreads();
}
__device__ OADG_HANDLER void hist_flush(const ChainArgs& A, ChainSmem& S, int slot, unsigned long long& lsum) {
  __syncthreads();
  unsigned* dst = A.hist + (size_t)slot * 768;
  for (int i = threadIdx.x; i < 768; i += kCT) {
    unsigned s = 0;
#pragma unroll
    for (int w = 0; w < 2; ++w) s += S.u.hist[w][i];
    if (s) atomicAdd(dst + i, s);
  }
  lsum = warp_sum(lsum);
  if ((threadIdx.x & 31) == 0 && lsum) atomicAdd(A.luma + slot, lsum);
  lsum = 0;
  __syncthreads();
}

// one LUT op: PIL.ImageOps autocontrast / equalize from the finished histogram, or a closed-form table
__device__ OADG_HANDLER void lut_tile(const ChainArgs& A, ChainSmem& S, int job) {
  const LutJob J = A.lutjobs[job];
  const oadg_op_t& op = A.P.ops[J.op];
  uint8_t* out = A.luts + (size_t)op.lut * 768;
  const int tid = threadIdx.x;
  __syncthreads();
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
      __syncthreads();
      if (lane == 31) wsum[warp] = incl;
      if (lane == 0) msk[warp] = bal;
      __syncthreads();
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
constexpr int kSubW = 64, kSubH = 16;
constexpr int kDynSmem = 30 * 1024;

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
  __syncthreads();
  const int key = I.obj * 2 + (catch_up ? 1 : 0);
  if (S.bs_key == key) return;   // the job is still staged from this CTA's previous tile (uniform: S.bs_key is shared)
  __syncthreads();
  if (threadIdx.x == 0) {
    S.bs_key = key;
    S.prof_ready = 0;
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
    bs.n_excl = catch_up ? J.next_count : 0;
    for (int k = 0; k < bs.n_excl && k < 16; ++k)
      for (int e = 0; e < 4; ++e) bs.excl[k][e] = A.bjobs[J.next_first + k].rect[e];
  }
  __syncthreads();
}
// blend of one box: tiles [l0, l1) of 128 x 16 px; Y = uint8(X*(1-m) + warp(X)*m) inside the support
__device__ OADG_HANDLER void bbo_r_segment(const ChainArgs& A, ChainSmem& S, uint8_t* dyn, const Item& I, int l0, int l1) {
  bbo_stage(A, S, I, false);
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
  const int w = bs.rect[2] - bs.rect[0], h = bs.rect[3] - bs.rect[1];
  const bool prof_smem = w + h <= 3072;
  if (prof_smem && !S.prof_ready) {
    for (int i = t; i < w; i += kCT) S.u.prof[i] = A.prof_x[(size_t)bs.gt * A.P.max_w + bs.rect[0] + i];
    for (int i = t; i < h; i += kCT) S.u.prof[w + i] = A.prof_y[(size_t)bs.gt * A.P.max_h + bs.rect[1] + i];
    __syncthreads();
    if (t == 0) S.prof_ready = 1;
  }
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
    const bool staged = any_src && (size_t)(sr[3] - sr[1]) * sv.pitch <= (size_t)kDynSmem && !(A.debug & 2);
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
    __syncthreads();  // the previous tile's gathers are done (and the profile slices are in place)
    const bool fast = staged && sr[3] - sr[1] <= 96;
    if (staged) {
      stage_rows(dyn, sv.pitch, bs.X, W, H, 3, sr[0], sr[1], sr[3] - sr[1]);
      fill_rowoff(S.rowoff[0], sv, sr[3] - sr[1]);
    }
    __syncthreads();
    if (!active) continue;
    const float uy = prof_smem ? S.u.prof[w + y - bs.rect[1]] : A.prof_y[(size_t)bs.gt * A.P.max_h + y];
    const WarpRowTerm rt = warp_row_term(bs.minv, y);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = xg + i;
      int v[3] = {byte_of(in_w, 3 * i), byte_of(in_w, 3 * i + 1), byte_of(in_w, 3 * i + 2)};
      if (x >= x0 && x < x1) {
        const float ux = prof_smem ? S.u.prof[x - bs.rect[0]] : A.prof_x[(size_t)bs.gt * A.P.max_w + x];
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
  for (int k = l0; k < l1; ++k) {
    const int tx0 = ax0 + (k % tx) * kBboCatchW, ty0 = bs.rect[1] + (k / tx) * kBboTileH;
    const int x0 = imax(tx0, bs.rect[0]), x1 = imin(tx0 + kBboCatchW, bs.rect[2]), y1 = imin(ty0 + kBboTileH, bs.rect[3]);
    for (int y = ty0 + (t >> 4); y < y1; y += 16)
#pragma unroll
    for (int gq = 0; gq < kBboCatchW / 64; ++gq) {
      const int xg = tx0 + gq * 64 + (t & 15) * 4;
      if (xg >= x1 || xg + 4 <= x0) continue;
      const size_t o = ((size_t)y * W + xg) * 3;
      // does a next-level support touch this 4-px group?
      bool touch = bs.n_excl > 16, all_in = false;
      if (bs.n_excl <= 16)
        for (int e = 0; e < bs.n_excl; ++e) {
          const bool yy = y >= bs.excl[e][1] && y < bs.excl[e][3];
          touch |= yy && xg < bs.excl[e][2] && xg + 4 > bs.excl[e][0];
          all_in |= yy && xg >= bs.excl[e][0] && xg + 4 <= bs.excl[e][2];
        }
      if (all_in) continue;
      if (!touch && vec && xg >= x0 && xg + 4 <= x1) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(bs.X + o);
        const uint32_t a0 = q[0], a1 = q[1], a2 = q[2];
        uint32_t* d = reinterpret_cast<uint32_t*>(bs.Y + o);
        d[0] = a0; d[1] = a1; d[2] = a2;
        continue;
      }
      for (int i = 0; i < 4; ++i)
        if (xg + i >= x0 && xg + i < x1) bbo_c_pixel(jobs, J, W, bs.X, bs.Y, xg + i, y);
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
    const double am = __ldg(div255 + wm);  // wm / 255 in float64, tabulated (exactly the reference's quotient)
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
  const bool staged = any_src && (size_t)rows * (pitch_i + pitch_m) <= (size_t)kDynSmem;
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
  __syncthreads();  // the previous sub-tile's gathers are done
  const bool fast = staged && rows <= 96;
  if (staged) {
    stage_rows(dyn, pitch_i, L.in, W, H, 3, sr[0], sr[1], rows);
    stage_rows(dyn + (size_t)rows * pitch_i, pitch_m, mu, W, H, 1, sr[0], sr[1], rows);
    fill_rowoff(S.rowoff[0], si, rows);
    fill_rowoff(S.rowoff[1], sm, rows);
  }
  __syncthreads();
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
        const double am = __ldg(div255 + wm);  // wm / 255 in float64, tabulated (exactly the reference's quotient)
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

// ------------------------------------------------------------------------------------
// one 256 x 16 tile of one depth step of one lane (oa_mix.py:226-234).  Runs of 16 pixels that one table-lookup /
// bbo-copy region covers move as three 16-byte vectors per thread (LUTs in shared memory); everything else (bg-only
// gathers, invert / colour / sharpness, runs cut by a multi-level box edge) is evaluated per pixel with consecutive
// lanes on consecutive pixels (run_is_stream, oamix_tile.h, decides which pass owns a run).
// ------------------------------------------------------------------------------------
__device__ OADG_HANDLER void step_tile(const ChainArgs& A, ChainSmem& S, uint8_t* dyn, int local, int tx, int tw,
                                       const double* div255) {
  const Lane& L = S.lane;
  const int W = L.W, H = L.H, t = threadIdx.x;
  const int x0 = (local % tx) * tw, y0 = (local / tx) * kStepTileH;
  const int x1 = min(x0 + tw, W), y1 = min(y0 + kStepTileH, H);
  const int region = tile_region(L, x0, y0, x1, y1);
  const bool tile_stream = region >= 0 && kind_streams(L.kind[region]);
  const bool tile_pixel = region >= 0 && !tile_stream;
  if (t == 0) S.step_class = tile_stream ? 7 : ((tile_pixel && S.rop[region].kind == OADG_OP_BG_AFFINE) ? 8 : 9);
  const bool staged_bg = !(A.debug & 1);
  if (!tile_stream && staged_bg) {
    // bg-only regions of the tile: staged gathers, sub-tile by sub-tile (a mixed tile filters by region per pixel)
    for (int r = 0; r <= L.n_ml; ++r) {
      if (S.rop[r].kind != OADG_OP_BG_AFFINE) continue;
      if (tile_pixel && r != region) continue;
      for (int sy0 = y0; sy0 < y1; sy0 += kSubH)
        for (int sx0 = x0; sx0 < x1; sx0 += kSubW) {
          const int sx1 = min(sx0 + kSubW, x1), sy1 = min(sy0 + kSubH, y1);
          if (!tile_pixel) {  // skip sub-tiles that hold no pixel of region r
            const int sub = tile_region(L, sx0, sy0, sx1, sy1);
            if (sub >= 0 && sub != r) continue;
            if (r < L.n_ml && !rect_hit(L.box[r], sx0, sy0, sx1, sy1)) continue;
          }
          bg_subtile(A, S, dyn, L, S.rop[r], tile_pixel ? -1 : r, sx0, sy0, sx1, sy1, div255);
        }
    }
    if (tile_pixel && S.rop[region].kind == OADG_OP_BG_AFFINE) return;
  }
  if (!tile_pixel) {
    for (int x = x0 + (t & 15) * kChunkPx; x < x1; x += 16 * kChunkPx)
    for (int y = y0 + (t >> 4); y < y1; y += 16) {
      const int n = min(kChunkPx, W - x);
      int reg;
      if (run_is_stream(L, x, y, n, reg)) {
        const bool vec = ((W * 3) & 15) == 0 &&
                         ((((uintptr_t)L.in) | ((uintptr_t)L.out) | ((uintptr_t)A.scratch) | A.frame_bytes) & 15) == 0;
        if (reg >= 0 && kind_streams(L.kind[reg]) && vec && n == kChunkPx) {
          Chunk c;
          chunk_load(stream_src(L, reg, A.scratch, A.frame_bytes) + ((size_t)y * W + x) * 3, n, vec, c);
          stream_chunk(L, reg, S.lut + reg * 768, A.scratch, A.frame_bytes, c, x, y, n, vec);
        } else {
          for (int i = 0; i < n; ++i) {
            const int rr = region_of_pixel(L, x + i, y);
            pixel_op_fast(A, L, S.rop[rr], S.lut, rr, x + i, y);
          }
        }
      }
    }
  }
  if (!tile_stream && !L.all_streaming) {
    for (int x = x0 + (t & 255); x < x1; x += 256) {  // this thread's columns
    const int xc = x & ~(kChunkPx - 1), nc = min(kChunkPx, W - xc);  // the 16-pixel run this column belongs to
    int ax[OADG_MAX_REGIONS], bx[OADG_MAX_REGIONS];
#pragma unroll
    for (int r = 0; r < OADG_MAX_REGIONS; ++r) {
      ax[r] = bx[r] = 0;
      if (r <= L.n_ml && S.rop[r].kind == OADG_OP_BG_AFFINE) {
        ax[r] = cv_round(dmul(dmul(S.rop[r].minv[0], (double)x), 1024.0));
        bx[r] = cv_round(dmul(dmul(S.rop[r].minv[3], (double)x), 1024.0));
      }
    }
    const int yb = y0, ye = y1;
#pragma unroll 1
    for (int y = yb; y < ye; ++y) {
      int run_region;
      if (run_is_stream(L, xc, y, nc, run_region)) continue;  // the vector pass owns this run
      const int r = region_of_pixel(L, x, y);
      if (S.rop[r].kind == OADG_OP_BG_AFFINE) {
        if (staged_bg) continue;  // done above from staged rows
        const int axr = r == 0 ? ax[0] : (r == 1 ? ax[1] : ax[2]);
        const int bxr = r == 0 ? bx[0] : (r == 1 ? bx[1] : bx[2]);
        bg_pixel_fast(A.P, L, S.rop[r], axr, bxr, div255, x, y);
      } else {
        pixel_op_fast(A, L, S.rop[r], S.lut, r, x, y);
      }
    }
    }
  }
}

// stage a lane record, its LUTs and its region ops in shared memory (once per lane a CTA works on)
__device__ OADG_HANDLER void stage_lane(const ChainArgs& A, ChainSmem& S, int lane) {
  const int t = threadIdx.x;
  __syncthreads();
  if (t < (int)(sizeof(Lane) / 4))
    reinterpret_cast<uint32_t*>(&S.lane)[t] = reinterpret_cast<const uint32_t*>(A.lanes + lane)[t];
  __syncthreads();
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
  __syncthreads();
}

__global__ void __launch_bounds__(kCT, kCtaPerSm)
oamix_chain_kernel(const ChainArgs Aparam, const double* div255) {
  __shared__ ChainSmem S;
  extern __shared__ __align__(16) uint8_t dyn[];   // kDynSmem bytes: staged source rows of the affine gathers
  if (threadIdx.x == 0) {
    S.args = Aparam;
    S.bs_key = -1;
    S.prof_ready = 0;
  }
  __syncthreads();
  const ChainArgs& A = S.args;
  int staged_lane = -1, scan_from = 0, streak = 0;
  const int kStickyTiles = A.debug >> 8;
  if (threadIdx.x < kMaxQueueItems / 32) S.exhausted[threadIdx.x] = 0u;
  __syncthreads();
  if (threadIdx.x == 0) {
    int item = -1, tile = -1;
    const unsigned long long w0 = globaltimer_ns();
    S.epoch_seen = ld_relaxed_u32(A.epoch);
    if (!scan_for_work(A, S.exhausted, scan_from, item, tile)) item = -1;
    if (!(A.debug & 4)) atomicAdd(A.kind_ns + 10, globaltimer_ns() - w0);
    S.next_item = item;
    S.next_tile = tile;
  }
  __syncthreads();
  int it = S.next_item, tile = S.next_tile;
  while (it >= 0) {
    __syncthreads();  // every thread has read S.next_item / S.next_tile
    const Item I = A.items[it];
    unsigned claim = 0;
    bool prefetched = false;
    if (threadIdx.x == 0) {
      // A CTA never claims ahead: a tile claimed while the previous one is still being processed sits idle at the
      // end of an item, and the items of the critical path then finish one tile time later (measured over 24 bench
      // batches: claim-ahead 17.2-18.2 ms, re-scan after every tile 15.0 ms).  OADG_DEBUG bits 8..: n > 1 = claim ahead
      // for n - 1 consecutive tiles, then re-scan.
      prefetched = kStickyTiles > 1 && (++streak % kStickyTiles) != 0;
      if (prefetched) claim = atomicAdd(A.claimed + it, 1u);
    }
    const int l0 = I.perm_first >= 0 ? A.perm[I.perm_first + tile] : tile, l1 = l0 + 1;
    const unsigned long long seg_t0 = globaltimer_ns();
    switch (I.kind) {
      case OADG_IT_PROFILE: profile_tile(A, S, I.obj); break;
      case OADG_IT_MASK:
        for (int k = l0; k < l1; ++k) mask_tile(A, S, I.obj, k, I.tx);
        break;
      case OADG_IT_HIST: {
        const Lane& L = A.lanes[I.obj];
        unsigned long long lsum = 0;
        hist_begin(S);
        for (int k = l0; k < l1; ++k) hist_tile(L, S, k, lsum);
        hist_flush(A, S, L.hist_slot, lsum);
        break;
      }
      case OADG_IT_LUT: lut_tile(A, S, I.obj); break;
      case OADG_IT_COPY: {
        const Chain& C = A.chains[I.obj];
        const oadg_view_t& V = A.P.views[C.view];
        copy_segment(C, (size_t)V.H * V.W * 3, I.aux != 0, l0, l1);
        break;
      }
      case OADG_IT_BBO_R: bbo_r_segment(A, S, dyn, I, l0, l1); break;
      case OADG_IT_BBO_C: bbo_c_segment(A, S, I, l0, l1); break;
      case OADG_IT_STEP:
        if (staged_lane != I.obj) {
          stage_lane(A, S, I.obj);
          staged_lane = I.obj;
        }
        for (int k = l0; k < l1; ++k) step_tile(A, S, dyn, k, I.tx, I.aux, div255);
        break;
      default: break;
    }
    publish_tile(A, I);
    if (threadIdx.x == 0) {
      const unsigned long long t1 = globaltimer_ns();
      int item = it, nt = (int)claim;
      if (!prefetched || claim >= (unsigned)I.ntiles) {   // look for the first ready item with tiles left
        if (prefetched) S.exhausted[it >> 5] |= 1u << (it & 31);
        S.epoch_seen = ld_relaxed_u32(A.epoch);
        if (!scan_for_work(A, S.exhausted, scan_from, item, nt)) item = -1;
      }
      if (!(A.debug & 4)) {
        const int kk = I.kind == OADG_IT_STEP ? S.step_class : I.kind;   // 7 stream, 8 bg staged, 9 mixed / per pixel
        const unsigned long long dt = t1 - seg_t0;
        atomicAdd(A.kind_ns + kk, dt);
        atomicAdd(A.kind_ns + 16 + kk, 1ull);
        atomicMax(A.kind_ns + 32 + kk, dt);
        atomicAdd(A.kind_ns + 10, globaltimer_ns() - t1);   // slot 10: looking / waiting for ready work
        atomicMax(A.item_ts + 2 * it, (1ull << 63) - seg_t0);
        atomicMax(A.item_ts + 2 * it + 1, t1);
      }
      S.next_item = item;
      S.next_tile = nt;
    }
    __syncthreads();
    it = S.next_item;
    tile = S.next_tile;
  }
}

// branch mixing + object-aware mixing (oa_mix.py:236,281-309); grid = (.., .., views)
constexpr int kTileThreads = 256;
// (2 CTAs per SM measured best: 3 -> +13 %, 4 -> +5 % kernel time)
__global__ void __launch_bounds__(kTileThreads, 2)
mix_kernel(DevPlan P, const MixJob* jobs) {
  __shared__ MixTile T;
  const MixJob J = jobs[blockIdx.z];
  const oadg_view_t& V = P.views[J.view];
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  if (x0 >= V.W || y0 >= V.H) return;
  const int x1 = min(x0 + kTileW, V.W), y1 = min(y0 + kTileH, V.H);
  if (threadIdx.x == 0) classify_mix_tile(P, J, x0, y0, x1, y1, T);
  __syncthreads();
  uintptr_t al = ((uintptr_t)J.src) | ((uintptr_t)J.out);
  for (int b = 0; b < V.width; ++b) al |= (uintptr_t)J.branch[b];
  const bool vec = ((V.W * 3) & 15) == 0 && (al & 15) == 0;
  const int t = threadIdx.x;
  const int x = x0 + (t & 15) * kChunkPx;
  if (x >= x1) return;
  const int n = min(kChunkPx, x1 - x);
#pragma unroll 1
  for (int y = y0 + (t >> 4); y < y1; y += 16) mix_chunk(P, J, T, x, y, n, vec);
}

#define BE_TRY(expr)                       \
  do {                                     \
    cudaError_t _e = (expr);               \
    if (_e != cudaSuccess) return (int)_e; \
  } while (0)

struct CudaBackend {
  cudaStream_t stream;
  int launches = 0;
  int n_sm = 0, ctas_per_sm = 0;
  // optional CUDA-event timing of the two launches (oadg_oamix_execute_profiled)
  bool profile = false;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  int n_items = 0, n_tiles = 0;
  const unsigned long long* kind_ns_dev = nullptr;
  const unsigned long long* item_ts_dev = nullptr;
  std::vector<Item> items_host;
  std::vector<int32_t> deps_host;

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
      int nb = 0;
      if (cudaFuncSetAttribute(oamix_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem) != cudaSuccess) return -1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, oamix_chain_kernel, kCT, kDynSmem) != cudaSuccess) return -1;
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
  // Plan + launch tables go up through a small ring of page-locked buffers (per calling thread): a copy from
  // pageable memory may make the host wait for the stream's earlier kernels, which would serialise the loader
  // loop with the GPU.  A slot is reused once the copy that read it has finished (event).
  struct PinSlot {
    void* p = nullptr;
    size_t cap = 0;
    cudaEvent_t ev = nullptr;
    int dev = -1;
  };
  int upload(void* dst, const void* src, size_t bytes) {
    static thread_local PinSlot ring[4];
    static thread_local int next = 0;
    PinSlot& sl = ring[next];
    next = (next + 1) & 3;
    int dev = 0;
    BE_TRY(cudaGetDevice(&dev));
    if (sl.ev && sl.dev != dev) {
      cudaEventDestroy(sl.ev);
      sl.ev = nullptr;
    }
    if (sl.ev) BE_TRY(cudaEventSynchronize(sl.ev));
    else BE_TRY(cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
    sl.dev = dev;
    if (sl.cap < bytes) {
      if (sl.p) cudaFreeHost(sl.p);
      sl.p = nullptr;
      sl.cap = 0;
      BE_TRY(cudaHostAlloc(&sl.p, 2 * bytes, cudaHostAllocPortable));
      sl.cap = 2 * bytes;
    }
    memcpy(sl.p, src, bytes);
    BE_TRY(cudaMemcpyAsync(dst, sl.p, bytes, cudaMemcpyHostToDevice, stream));
    BE_TRY(cudaEventRecord(sl.ev, stream));
    return 0;
  }
  int zero(void* dst, size_t bytes) {
    BE_TRY(cudaMemsetAsync(dst, 0, bytes, stream));
    return 0;
  }
  int chain(const ChainArgs& A, const ChainArgs& Hh, const PlanView&) {
    const double* div255 = nullptr;
    BE_TRY(cudaGetSymbolAddress((void**)&div255, g_div255));
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
      void* params[2] = {(void*)&args, (void*)&div255};
      // cooperative launch: all CTAs are guaranteed co-resident, which the in-kernel grid barrier relies on
      BE_TRY(cudaLaunchCooperativeKernel((const void*)oamix_chain_kernel, dim3(A.grid), dim3(kCT), params, kDynSmem,
                                         stream));
      ++launches;
    }
    if (profile) BE_TRY(cudaEventRecord(ev[1], stream));
    return 0;
  }
  int mix(const DevPlan& P, const MixJob* jobs, int n) {
    dim3 grid((P.max_w + kTileW - 1) / kTileW, (P.max_h + kTileH - 1) / kTileH, n);
    mix_kernel<<<grid, kTileThreads, 0, stream>>>(P, jobs);
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
    if (ks[15] != 0) rc = OADG_E_PLAN;   // a CTA gave up waiting: the dependency tables were inconsistent
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

extern "C" int oadg_oamix_execute(const void* plan_host, size_t plan_bytes, const uint8_t* const* src_dev,
                                  int n_img, uint8_t* const* dst_dev, void* workspace_dev,
                                  size_t workspace_bytes, int* launches_out, void* stream) {
  CudaBackend be;
  be.stream = (cudaStream_t)stream;
  int rc = execute_plan(be, plan_host, plan_bytes, src_dev, n_img, dst_dev, workspace_dev, workspace_bytes);
  if (launches_out) *launches_out = be.launches;
  return rc;
}
