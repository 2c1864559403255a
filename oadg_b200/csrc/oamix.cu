// OA-Mix plan executor: the device side of OAMix.oamix (reference oa_mix.py:207-309).
//
// Data layout in HBM (all caller- or workspace-owned):
//   frames      u8 HWC, pitch 3*W, one per source image / branch ping-pong / bbo scratch
//   profiles    per gt box two float32 vectors ux[W], uy[H]; blurred mask(y,x) = uy[y]*ux[x]
//               (the reference materialises a 25 MB float HxWx3 mask per box, oa_mix.py:75-93)
//   hist / lut  per lane 3x256 u32 histogram (+ luma sum), per LUT op 3x256 u8 table
//   plan        the host-sampled plan blob (oadg.h records) + launch tables, one H2D copy
//
// Kernel chain per batch (oamix_exec.h):
//   profile_kernel                      1 launch   (all boxes of all views)
//   for depth d:  hist_kernel           <=1 launch (lanes whose step needs a histogram)
//                 lut_kernel            <=1 launch
//                 bbo_pass_kernel       1 per chained gt box (bboxes-only ops, ROI-limited, ping-pong frames)
//                 step_kernel           1 launch   (all (view, branch) lanes alive at depth d)
//   mix_kernel                          1 launch   (all views)
#include "oadg_common.cuh"
#include "oamix_exec.h"
#include "oamix_tile.h"

namespace oadg {
namespace {

// i / 255 in float64 (the reference divides a uint8 array by the python int 255, bbox_augmentation.py:267)
__device__ const double g_div255[256] = {0.0 / 255.0, 1.0 / 255.0, 2.0 / 255.0, 3.0 / 255.0, 4.0 / 255.0, 5.0 / 255.0, 6.0 / 255.0, 7.0 / 255.0, 8.0 / 255.0, 9.0 / 255.0, 10.0 / 255.0, 11.0 / 255.0, 12.0 / 255.0, 13.0 / 255.0, 14.0 / 255.0, 15.0 / 255.0, 16.0 / 255.0, 17.0 / 255.0, 18.0 / 255.0, 19.0 / 255.0, 20.0 / 255.0, 21.0 / 255.0, 22.0 / 255.0, 23.0 / 255.0, 24.0 / 255.0, 25.0 / 255.0, 26.0 / 255.0, 27.0 / 255.0, 28.0 / 255.0, 29.0 / 255.0, 30.0 / 255.0, 31.0 / 255.0, 32.0 / 255.0, 33.0 / 255.0, 34.0 / 255.0, 35.0 / 255.0, 36.0 / 255.0, 37.0 / 255.0, 38.0 / 255.0, 39.0 / 255.0, 40.0 / 255.0, 41.0 / 255.0, 42.0 / 255.0, 43.0 / 255.0, 44.0 / 255.0, 45.0 / 255.0, 46.0 / 255.0, 47.0 / 255.0, 48.0 / 255.0, 49.0 / 255.0, 50.0 / 255.0, 51.0 / 255.0, 52.0 / 255.0, 53.0 / 255.0, 54.0 / 255.0, 55.0 / 255.0, 56.0 / 255.0, 57.0 / 255.0, 58.0 / 255.0, 59.0 / 255.0, 60.0 / 255.0, 61.0 / 255.0, 62.0 / 255.0, 63.0 / 255.0, 64.0 / 255.0, 65.0 / 255.0, 66.0 / 255.0, 67.0 / 255.0, 68.0 / 255.0, 69.0 / 255.0, 70.0 / 255.0, 71.0 / 255.0, 72.0 / 255.0, 73.0 / 255.0, 74.0 / 255.0, 75.0 / 255.0, 76.0 / 255.0, 77.0 / 255.0, 78.0 / 255.0, 79.0 / 255.0, 80.0 / 255.0, 81.0 / 255.0, 82.0 / 255.0, 83.0 / 255.0, 84.0 / 255.0, 85.0 / 255.0, 86.0 / 255.0, 87.0 / 255.0, 88.0 / 255.0, 89.0 / 255.0, 90.0 / 255.0, 91.0 / 255.0, 92.0 / 255.0, 93.0 / 255.0, 94.0 / 255.0, 95.0 / 255.0, 96.0 / 255.0, 97.0 / 255.0, 98.0 / 255.0, 99.0 / 255.0, 100.0 / 255.0, 101.0 / 255.0, 102.0 / 255.0, 103.0 / 255.0, 104.0 / 255.0, 105.0 / 255.0, 106.0 / 255.0, 107.0 / 255.0, 108.0 / 255.0, 109.0 / 255.0, 110.0 / 255.0, 111.0 / 255.0, 112.0 / 255.0, 113.0 / 255.0, 114.0 / 255.0, 115.0 / 255.0, 116.0 / 255.0, 117.0 / 255.0, 118.0 / 255.0, 119.0 / 255.0, 120.0 / 255.0, 121.0 / 255.0, 122.0 / 255.0, 123.0 / 255.0, 124.0 / 255.0, 125.0 / 255.0, 126.0 / 255.0, 127.0 / 255.0, 128.0 / 255.0, 129.0 / 255.0, 130.0 / 255.0, 131.0 / 255.0, 132.0 / 255.0, 133.0 / 255.0, 134.0 / 255.0, 135.0 / 255.0, 136.0 / 255.0, 137.0 / 255.0, 138.0 / 255.0, 139.0 / 255.0, 140.0 / 255.0, 141.0 / 255.0, 142.0 / 255.0, 143.0 / 255.0, 144.0 / 255.0, 145.0 / 255.0, 146.0 / 255.0, 147.0 / 255.0, 148.0 / 255.0, 149.0 / 255.0, 150.0 / 255.0, 151.0 / 255.0, 152.0 / 255.0, 153.0 / 255.0, 154.0 / 255.0, 155.0 / 255.0, 156.0 / 255.0, 157.0 / 255.0, 158.0 / 255.0, 159.0 / 255.0, 160.0 / 255.0, 161.0 / 255.0, 162.0 / 255.0, 163.0 / 255.0, 164.0 / 255.0, 165.0 / 255.0, 166.0 / 255.0, 167.0 / 255.0, 168.0 / 255.0, 169.0 / 255.0, 170.0 / 255.0, 171.0 / 255.0, 172.0 / 255.0, 173.0 / 255.0, 174.0 / 255.0, 175.0 / 255.0, 176.0 / 255.0, 177.0 / 255.0, 178.0 / 255.0, 179.0 / 255.0, 180.0 / 255.0, 181.0 / 255.0, 182.0 / 255.0, 183.0 / 255.0, 184.0 / 255.0, 185.0 / 255.0, 186.0 / 255.0, 187.0 / 255.0, 188.0 / 255.0, 189.0 / 255.0, 190.0 / 255.0, 191.0 / 255.0, 192.0 / 255.0, 193.0 / 255.0, 194.0 / 255.0, 195.0 / 255.0, 196.0 / 255.0, 197.0 / 255.0, 198.0 / 255.0, 199.0 / 255.0, 200.0 / 255.0, 201.0 / 255.0, 202.0 / 255.0, 203.0 / 255.0, 204.0 / 255.0, 205.0 / 255.0, 206.0 / 255.0, 207.0 / 255.0, 208.0 / 255.0, 209.0 / 255.0, 210.0 / 255.0, 211.0 / 255.0, 212.0 / 255.0, 213.0 / 255.0, 214.0 / 255.0, 215.0 / 255.0, 216.0 / 255.0, 217.0 / 255.0, 218.0 / 255.0, 219.0 / 255.0, 220.0 / 255.0, 221.0 / 255.0, 222.0 / 255.0, 223.0 / 255.0, 224.0 / 255.0, 225.0 / 255.0, 226.0 / 255.0, 227.0 / 255.0, 228.0 / 255.0, 229.0 / 255.0, 230.0 / 255.0, 231.0 / 255.0, 232.0 / 255.0, 233.0 / 255.0, 234.0 / 255.0, 235.0 / 255.0, 236.0 / 255.0, 237.0 / 255.0, 238.0 / 255.0, 239.0 / 255.0, 240.0 / 255.0, 241.0 / 255.0, 242.0 / 255.0, 243.0 / 255.0, 244.0 / 255.0, 245.0 / 255.0, 246.0 / 255.0, 247.0 / 255.0, 248.0 / 255.0, 249.0 / 255.0, 250.0 / 255.0, 251.0 / 255.0, 252.0 / 255.0, 253.0 / 255.0, 254.0 / 255.0, 255.0 / 255.0};

// ------------------------------------------------------------------------------------
// blurred-mask profiles (oa_mix.py:78-91): indicator on the 1/sr canvas -> GaussianBlur
// (separable, BORDER_REFLECT_101, float32 kernel from getGaussianKernel) -> bilinear
// cv2.resize to full resolution.  grid = (n_gt, 2 axes).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
profile_kernel(DevPlan P, int sr, float* __restrict__ prof_x, float* __restrict__ prof_y) {
  extern __shared__ float sm[];  // [n_lo] blurred low-res profile, then [ksize] kernel
  const int g = blockIdx.x, axis = blockIdx.y;
  const oadg_gt_t G = P.gts[g];
  const oadg_view_t& V = P.views[G.view];
  const int n_hi = axis == 0 ? V.W : V.H;
  const int n_lo = n_hi / sr;
  const int lo = G.lo[axis], hi = G.lo[axis + 2];
  const int ks = axis == 0 ? G.kx : G.ky;
  const double sigma = axis == 0 ? G.sigma_x : G.sigma_y;
  float* p = sm;
  float* kern = sm + n_lo;
  float* out = (axis == 0 ? prof_x + (size_t)g * P.max_w : prof_y + (size_t)g * P.max_h);
  const int tid = threadIdx.x;
  __shared__ double red[8];
  __shared__ double ksum;
  if (n_lo <= 0) {
    for (int d = tid; d < n_hi; d += blockDim.x) out[d] = 0.f;
    return;
  }
  if (G.blur) {
    // cv::getGaussianKernel(ks, sigma, CV_32F): exp(-x^2/(2 sigma^2)) in double, normalised, cast
    const double s2 = -0.5 / (sigma * sigma);
    double part = 0.0;
    for (int i = tid; i < ks; i += blockDim.x) {
      double x = i - (ks - 1) * 0.5;
      part += exp(s2 * x * x);
    }
    part = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
      for (int w = 0; w < 8; ++w) t += red[w];
      ksum = 1.0 / t;
    }
    __syncthreads();
    for (int i = tid; i < ks; i += blockDim.x) {
      double x = i - (ks - 1) * 0.5;
      kern[i] = (float)(exp(s2 * x * x) * ksum);
    }
    __syncthreads();
    const int r = ks / 2;
    const int period = 2 * (n_lo - 1);
    for (int x = tid; x < n_lo; x += blockDim.x) {
      double acc = 0.0;
      for (int j = 0; j < ks; ++j) {
        int q = x + j - r;
        if (n_lo == 1) q = 0;
        else {
          if (q < 0) q = -q;
          q %= period;
          if (q >= n_lo) q = period - q;
        }
        if (q >= lo && q < hi) acc += (double)kern[j];
      }
      p[x] = (float)acc;
    }
  } else {
    for (int x = tid; x < n_lo; x += blockDim.x) p[x] = (x >= lo && x < hi) ? 1.f : 0.f;
  }
  __syncthreads();
  // cv2.resize(f32, INTER_LINEAR): fx = (float)((dx+0.5)*scale - 0.5)
  const double scale = (double)n_lo / (double)n_hi;
  for (int d = tid; d < n_hi; d += blockDim.x) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    float t = fsub(f, (float)s);
    if (s < 0) { s = 0; t = 0.f; }
    if (s >= n_lo - 1) { s = n_lo - 1; t = 0.f; }
    int s1 = min(s + 1, n_lo - 1);
    out[d] = fadd(fmul(p[s], fsub(1.f, t)), fmul(p[s1], t));
  }
}

// ------------------------------------------------------------------------------------
// union of the blurred gt masks of every view (np.max(mask_bboxes, axis=0), bbox_augmentation.py:260), as float32
// and as uint8(mask*255): computed once per batch, read by every bg-only op.  grid = (ceil(W/256), H, views)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mask_kernel(DevPlan P, float* __restrict__ maskf, uint8_t* __restrict__ masku) {
  const int view = blockIdx.z;
  const oadg_view_t& V = P.views[view];
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if (x >= V.W || y >= V.H) return;
  mask_pixel(P, view, x, y, maskf, masku);
}

// ------------------------------------------------------------------------------------
// per-channel histogram + luma sum of a lane's input frame (PIL Image.histogram())
// grid = (blocks, lanes_with_hist)
// ------------------------------------------------------------------------------------
constexpr int kHistThreads = 256;
__global__ void __launch_bounds__(kHistThreads)
hist_kernel(DevPlan P, const Lane* __restrict__ lanes, const int32_t* __restrict__ lane_ids,
            unsigned* __restrict__ hist, unsigned long long* __restrict__ luma) {
  __shared__ unsigned sh[8][768];
  const Lane L = lanes[lane_ids[blockIdx.y]];
  const oadg_view_t& V = P.views[L.view];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 8 * 768; i += kHistThreads) (&sh[0][0])[i] = 0;
  __syncthreads();
  const size_t npx = (size_t)V.H * V.W;
  unsigned long long lsum = 0;
  unsigned* my = sh[warp];
  if ((((uintptr_t)L.in) & 15) == 0) {
    // 16 px = 48 B = 3 x uint4 per iteration: the channel of byte k is k % 3 at a compile-time phase
    const size_t nchunk = npx / kChunkPx;
    for (size_t i = (size_t)blockIdx.x * kHistThreads + tid; i < nchunk; i += (size_t)gridDim.x * kHistThreads) {
      Chunk c;
      chunk_load(L.in + i * 48, kChunkPx, true, c);
#pragma unroll
      for (int px = 0; px < kChunkPx; ++px) {
        const int c0 = chunk_get(c, px * 3), c1 = chunk_get(c, px * 3 + 1), c2 = chunk_get(c, px * 3 + 2);
        atomicAdd(&my[c0], 1u);
        atomicAdd(&my[256 + c1], 1u);
        atomicAdd(&my[512 + c2], 1u);
        lsum += (unsigned)pil_luma(c0, c1, c2);
      }
    }
    for (size_t i = nchunk * kChunkPx + (size_t)blockIdx.x * kHistThreads + tid; i < npx;
         i += (size_t)gridDim.x * kHistThreads) {
      const uint8_t* p = L.in + i * 3;
      int c0 = ldb(p), c1 = ldb(p + 1), c2 = ldb(p + 2);
      atomicAdd(&my[c0], 1u);
      atomicAdd(&my[256 + c1], 1u);
      atomicAdd(&my[512 + c2], 1u);
      lsum += (unsigned)pil_luma(c0, c1, c2);
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * kHistThreads + tid; i < npx; i += (size_t)gridDim.x * kHistThreads) {
      const uint8_t* p = L.in + i * 3;
      int c0 = ldb(p), c1 = ldb(p + 1), c2 = ldb(p + 2);
      atomicAdd(&my[c0], 1u);
      atomicAdd(&my[256 + c1], 1u);
      atomicAdd(&my[512 + c2], 1u);
      lsum += (unsigned)pil_luma(c0, c1, c2);
    }
  }
  __syncthreads();
  unsigned* dst = hist + (size_t)L.hist_slot * 768;
  for (int i = tid; i < 768; i += kHistThreads) {
    unsigned s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sh[w][i];
    if (s) atomicAdd(dst + i, s);
  }
  lsum = warp_sum(lsum);
  if ((tid & 31) == 0 && lsum) atomicAdd(luma + L.hist_slot, lsum);
}

// one block per LUT op
__global__ void __launch_bounds__(256)
lut_kernel(DevPlan P, const LutJob* __restrict__ jobs, const unsigned* __restrict__ hist,
           const unsigned long long* __restrict__ luma, uint8_t* __restrict__ luts) {
  const LutJob J = jobs[blockIdx.x];
  const oadg_op_t& op = P.ops[J.op];
  uint8_t* out = luts + (size_t)op.lut * 768;
  const int tid = threadIdx.x;
  __shared__ uint8_t tab[3][256];
  if (op.kind == OADG_OP_AUTOCONTRAST || op.kind == OADG_OP_EQUALIZE) {
    // 3 channels x 256 entries; the sequential scans are tiny: one thread per channel
    if (tid < 3) {
      const unsigned* h = hist + (size_t)J.hist_slot * 768 + tid * 256;
      if (op.kind == OADG_OP_AUTOCONTRAST) lut_autocontrast_ch(h, tab[tid]);
      else lut_equalize_ch(h, tab[tid]);
    }
    __syncthreads();
    for (int i = tid; i < 768; i += 256) out[i] = (&tab[0][0])[i];
    return;
  }
  const oadg_view_t& V = P.views[J.view];
  const double lsum = J.hist_slot >= 0 ? (double)luma[J.hist_slot] : 0.0;
  const uint8_t v = lut_simple_at(op, tid, lsum, (double)((long long)V.H * V.W));
  out[tid] = v;
  out[256 + tid] = v;
  out[512 + tid] = v;
}

// ------------------------------------------------------------------------------------
// bboxes-only chains (bbox_augmentation.py:74-88): box j of every active chain.
// pass j: Y[roi_j] = blend(X, warp_j(X), m_j), Y[roi_{j-1} \\ roi_j] = X  with (X, Y) = (S, T) swapping per box.
// grid = (ceil(max_roi_w/32), ceil(max_roi_h/8), chains)
// ------------------------------------------------------------------------------------
// Each thread walks 8 rows of one column of the pass rectangle: the box record, the support rectangles and the
// x-dependent half of the fixed-point affine coordinates are loaded / computed once per thread.
constexpr int kBboRows = 8;
__global__ void __launch_bounds__(256)
bbo_pass_kernel(DevPlan P, const Chain* __restrict__ chains, int j) {
  const Chain C = chains[blockIdx.z];
  if (j >= C.n) return;
  int r[4];
  bbo_pass_rect(P, C, j, r);
  const int x = r[0] + blockIdx.x * 32 + threadIdx.x;
  const int yb = r[1] + (blockIdx.y * 8 + threadIdx.y) * kBboRows;
  if (x >= r[2] || yb >= r[3]) return;
  const uint8_t* X = (j & 1) ? C.T : C.S;
  uint8_t* Y = (j & 1) ? C.S : C.T;
  const oadg_bbo_t& B = P.bbo[C.bbo_first + j];
  const int g = B.gt;
  const oadg_view_t& V = P.views[C.view];
  const int W = V.W, H = V.H;
  int cur[4], prev[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) cur[i] = P.gts[g].supp[i];
  if (j > 0) {
    const int32_t* q = P.gts[P.bbo[C.bbo_first + j - 1].gt].supp;
#pragma unroll
    for (int i = 0; i < 4; ++i) prev[i] = q[i];
  }
  double minv[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) minv[i] = B.minv[i];
  const bool x_cur = x >= cur[0] && x < cur[2], x_prev = x >= prev[0] && x < prev[2];
  if (!x_cur && !x_prev) return;
  const int ax = cv_round(dmul(dmul(minv[0], (double)x), 1024.0));
  const int bx = cv_round(dmul(dmul(minv[3], (double)x), 1024.0));
  const float ux = x_cur ? __ldg(P.prof_x + (size_t)g * P.max_w + x) : 0.f;
  const float* py = P.prof_y + (size_t)g * P.max_h;
  const int y_end = min(yb + kBboRows, r[3]);
#pragma unroll 1
  for (int y = yb; y < y_end; ++y) {
    const size_t o = ((size_t)y * W + x) * 3;
    if (!(x_cur && y >= cur[1] && y < cur[3])) {
      if (x_prev && y >= prev[1] && y < prev[3]) {  // catch up the pixels only box j-1 touched
        Y[o] = X[o];
        Y[o + 1] = X[o + 1];
        Y[o + 2] = X[o + 2];
      }
      continue;
    }
    const float m = fmul(__ldg(py + y), ux);
    int v[3] = {X[o], X[o + 1], X[o + 2]};
    if (m != 0.f) {  // m == 0 => img*1 + aug*0 == img exactly
      const int Xf = (cv_round(dmul(dadd(dmul(minv[1], (double)y), minv[2]), 1024.0)) + 16 + ax) >> 5;
      const int Yf = (cv_round(dmul(dadd(dmul(minv[4], (double)y), minv[5]), 1024.0)) + 16 + bx) >> 5;
      WarpTap t;
      t.sx = imin(imax(Xf >> 5, -32768), 32767);
      t.sy = imin(imax(Yf >> 5, -32768), 32767);
      t.fx = Xf & 31;
      t.fy = Yf & 31;
      int a[3];
      warp_fetch3(LdRW(), X, H, W, t, a);
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = bbo_blend(m, v[c], a[c]);
    }
    Y[o] = (uint8_t)v[0];
    Y[o + 1] = (uint8_t)v[1];
    Y[o + 2] = (uint8_t)v[2];
  }
}

// ------------------------------------------------------------------------------------
// one depth step of every live lane (oa_mix.py:226-234), split by kind of work (run_is_stream, oamix_tile.h):
//
// step_kernel        the HBM-bound part: table-lookup ops and bbo-result copies.  Persistent CTAs, grid =
//                    (CTAs per lane, lanes) sized to fill the 148 SMs once; a CTA stages its Lane record and the
//                    lane's LUTs in shared memory once, then grid-strides over 16-pixel runs (48 B = 3 x 16-byte
//                    vectors), two runs in flight per thread.
// step_pixel_kernel  everything that needs per-pixel evaluation (bg-only gathers, invert / colour / sharpness,
//                    runs cut by a multi-level box edge), one CTA per 256 x 32 tile, consecutive lanes on
//                    consecutive pixels; launched only over lanes that have such an op.
// ------------------------------------------------------------------------------------
constexpr int kTileThreads = 256;

struct RegOp {      // op parameters of one region, staged in shared memory for per-pixel tiles
  int32_t kind, p0, p1;
  float factor;
  double minv[6];
};

__global__ void __launch_bounds__(kTileThreads, 4)
step_kernel(DevPlan P, const Lane* __restrict__ lanes, const uint8_t* __restrict__ scratch, size_t frame_bytes) {
  __shared__ __align__(16) uint8_t lut_s[OADG_MAX_REGIONS * 768];
  __shared__ Lane Ls;
  const int t = threadIdx.x;
  if (t < (int)(sizeof(Lane) / 4))
    reinterpret_cast<uint32_t*>(&Ls)[t] = __ldg(reinterpret_cast<const uint32_t*>(lanes + blockIdx.y) + t);
  __syncthreads();
  const Lane& L = Ls;
  bool any_lut = false;
#pragma unroll
  for (int r = 0; r < OADG_MAX_REGIONS; ++r) {
    if (r <= L.n_ml && L.lut[r] >= 0) {
      any_lut = true;
      if (t < 192)
        reinterpret_cast<uint32_t*>(lut_s + r * 768)[t] =
            __ldg(reinterpret_cast<const uint32_t*>(P.luts + (size_t)L.lut[r] * 768) + t);
    }
  }
  if (any_lut) __syncthreads();
  const int W = L.W, H = L.H;
  const int cpr = (W + kChunkPx - 1) / kChunkPx;
  const int total = cpr * H;
  const int stride = gridDim.x * kTileThreads;
  const bool vec = ((W * 3) & 15) == 0 &&
                   ((((uintptr_t)L.in) | ((uintptr_t)L.out) | ((uintptr_t)scratch) | frame_bytes) & 15) == 0;
  for (int u0 = blockIdx.x * kTileThreads + t; u0 < total; u0 += 2 * stride) {
    // two independent runs per iteration so that six 16-byte loads are in flight per thread
    int x[2], y[2], n[2], reg[2];
    bool vecrun[2];
    Chunk c[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int u = u0 + k * stride;
      n[k] = 0;
      vecrun[k] = false;
      if (u < total) {
        y[k] = u / cpr;
        x[k] = (u - y[k] * cpr) * kChunkPx;
        n[k] = min(kChunkPx, W - x[k]);
        const bool mine = run_is_stream(L, x[k], y[k], n[k], reg[k]);
        if (!mine) n[k] = 0;
        else if (reg[k] >= 0 && kind_streams(L.kind[reg[k]]) && vec && n[k] == kChunkPx) {
          vecrun[k] = true;
          chunk_load(stream_src(L, reg[k], scratch, frame_bytes) + ((size_t)y[k] * W + x[k]) * 3, n[k], vec, c[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (n[k] == 0) continue;
      if (vecrun[k]) stream_chunk(L, reg[k], lut_s + reg[k] * 768, scratch, frame_bytes, c[k], x[k], y[k], n[k], vec);
      else
        for (int i = 0; i < n[k]; ++i) stream_pixel(L, lut_s, scratch, frame_bytes, x[k] + i, y[k]);
    }
  }
}

// one pixel of a bg-only op with hoisted coordinate terms (same arithmetic as bg_pixel / eval_op)
__device__ __forceinline__ void bg_pixel_fast(const DevPlan& P, const Lane& L, const RegOp& R, int ax, int bx,
                                              const double* __restrict__ div255, int x, int y) {
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
  const float M = __ldg(P.maskf + mo + (size_t)y * L.W + x);
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

__global__ void __launch_bounds__(kTileThreads, 4)
step_pixel_kernel(DevPlan P, const Lane* __restrict__ lanes, const int32_t* __restrict__ lane_ids,
                  const uint8_t* __restrict__ scratch, size_t frame_bytes, const double* __restrict__ div255) {
  __shared__ RegOp rop[OADG_MAX_REGIONS];
  __shared__ Lane Ls;
  const int t = threadIdx.x;
  if (t < (int)(sizeof(Lane) / 4))
    reinterpret_cast<uint32_t*>(&Ls)[t] = __ldg(reinterpret_cast<const uint32_t*>(lanes + lane_ids[blockIdx.z]) + t);
  __syncthreads();
  const Lane& L = Ls;
  const int W = L.W, H = L.H;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  if (x0 >= W || y0 >= H) return;
  const int x1 = min(x0 + kTileW, W), y1 = min(y0 + kTileH, H);
  {  // a tile that one streaming region covers has nothing for this kernel
    const int region = tile_region(L, x0, y0, x1, y1);
    if (region >= 0 && kind_streams(L.kind[region])) return;
  }
  if (t <= L.n_ml) {
    const oadg_op_t& op = P.ops[L.op_base + t];
    rop[t].kind = op.kind;
    rop[t].p0 = op.p0;
    rop[t].p1 = op.p1;
    rop[t].factor = op.factor;
#pragma unroll
    for (int i = 0; i < 6; ++i) rop[t].minv[i] = op.minv[i];
  }
  __syncthreads();
  const int x = x0 + t;  // this thread's column is fixed
  if (x >= x1) return;
  const int xc = x & ~(kChunkPx - 1), nc = min(kChunkPx, W - xc);  // the 16-pixel run this column belongs to
  int ax[OADG_MAX_REGIONS], bx[OADG_MAX_REGIONS];
#pragma unroll
  for (int r = 0; r < OADG_MAX_REGIONS; ++r) {
    ax[r] = bx[r] = 0;
    if (r <= L.n_ml && rop[r].kind == OADG_OP_BG_AFFINE) {
      ax[r] = cv_round(dmul(dmul(rop[r].minv[0], (double)x), 1024.0));
      bx[r] = cv_round(dmul(dmul(rop[r].minv[3], (double)x), 1024.0));
    }
  }
#pragma unroll 1
  for (int y = y0; y < y1; ++y) {
    int run_region;
    if (run_is_stream(L, xc, y, nc, run_region)) continue;  // the stream kernel owns this run
    const int r = region_of_pixel(L, x, y);
    if (rop[r].kind == OADG_OP_BG_AFFINE) {
      const int axr = r == 0 ? ax[0] : (r == 1 ? ax[1] : ax[2]);
      const int bxr = r == 0 ? bx[0] : (r == 1 ? bx[1] : bx[2]);
      bg_pixel_fast(P, L, rop[r], axr, bxr, div255, x, y);
    } else {
      step_pixel(P, L, scratch, frame_bytes, x, y);
    }
  }
}

// branch mixing + object-aware mixing (oa_mix.py:236,281-309), same tiling; grid = (.., .., views)
__global__ void __launch_bounds__(kTileThreads, 2)
mix_kernel(DevPlan P, const MixJob* __restrict__ jobs) {
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

enum { kKProfile = 0, kKHist, kKLut, kKBboPass, kKMask, kKStep, kKMix, kKCopy, kKStepPx, kKinds };

struct CudaBackend {
  cudaStream_t stream;
  int launches = 0;
  // optional per-launch CUDA-event timing (oadg_oamix_execute_profiled)
  bool profile = false;
  struct Rec { int kind; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  cudaEvent_t pending = nullptr;
  void begin() {
    if (!profile) return;
    cudaEventCreate(&pending);
    cudaEventRecord(pending, stream);
  }
  void end(int kind) {
    ++launches;
    if (!profile) return;
    cudaEvent_t b;
    cudaEventCreate(&b);
    cudaEventRecord(b, stream);
    recs.push_back(Rec{kind, pending, b});
  }

  int upload(void* dst, const void* src, size_t bytes) {
    BE_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return 0;
  }
  int zero(void* dst, size_t bytes) {
    BE_TRY(cudaMemsetAsync(dst, 0, bytes, stream));
    return 0;
  }
  int copy(void* dst, const void* src, size_t bytes) {
    begin();
    BE_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream));
    end(kKCopy);
    return 0;
  }
  int profiles(const DevPlan& P, const PlanView& pv, float* px, float* py) {
    const oadg_plan_header_t& h = *pv.h;
    const int sr = 4;  // spatial_ratio of every reference config; the host rejects other values
    int max_lo = (h.max_w > h.max_h ? h.max_w : h.max_h) / sr + 1;
    int max_k = 1;
    for (int g = 0; g < h.n_gt; ++g) {
      max_k = pv.gts[g].kx > max_k ? pv.gts[g].kx : max_k;
      max_k = pv.gts[g].ky > max_k ? pv.gts[g].ky : max_k;
    }
    size_t smem = (size_t)(max_lo + max_k) * sizeof(float);
    if (smem > 200 * 1024) return OADG_E_LIMIT;
    if (smem > 48 * 1024)
      BE_TRY(cudaFuncSetAttribute(profile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    begin();
    profile_kernel<<<dim3(h.n_gt, 2), 256, smem, stream>>>(P, sr, px, py);
    BE_TRY(cudaGetLastError());
    end(kKProfile);
    return 0;
  }
  int masks(const DevPlan& P, int n_views, float* maskf, uint8_t* masku) {
    begin();
    mask_kernel<<<dim3((P.max_w + 255) / 256, P.max_h, n_views), 256, 0, stream>>>(P, maskf, masku);
    BE_TRY(cudaGetLastError());
    end(kKMask);
    return 0;
  }
  int hist(const DevPlan& P, const Lane* lanes, const int32_t* ids, int n, unsigned* hist, unsigned long long* luma) {
    begin();
    hist_kernel<<<dim3(kNumSMs * 2, n), kHistThreads, 0, stream>>>(P, lanes, ids, hist, luma);
    BE_TRY(cudaGetLastError());
    end(kKHist);
    return 0;
  }
  int lut(const DevPlan& P, const LutJob* jobs, int n, const unsigned* hist, const unsigned long long* luma,
          uint8_t* luts) {
    begin();
    lut_kernel<<<n, 256, 0, stream>>>(P, jobs, hist, luma, luts);
    BE_TRY(cudaGetLastError());
    end(kKLut);
    return 0;
  }
  int bbo_pass(const DevPlan& P, const Chain* chains, int n, int j, int roi_w, int roi_h) {
    begin();
    bbo_pass_kernel<<<dim3((roi_w + 31) / 32, (roi_h + 8 * kBboRows - 1) / (8 * kBboRows), n), dim3(32, 8), 0, stream>>>(
        P, chains, j);
    BE_TRY(cudaGetLastError());
    end(kKBboPass);
    return 0;
  }
  // lanes of one depth; lane_px_ids / n_px: the lanes that have a per-pixel op
  int step(const DevPlan& P, const Lane* lanes, int n, const int32_t* lane_px_ids, int n_px, const uint8_t* scratch,
           size_t frame_bytes) {
    int per_lane = (kNumSMs * 4 + n - 1) / n;  // persistent CTAs: fill the SMs once (4 CTAs of 256 threads per SM)
    if (per_lane < 1) per_lane = 1;
    begin();
    step_kernel<<<dim3(per_lane, n), kTileThreads, 0, stream>>>(P, lanes, scratch, frame_bytes);
    BE_TRY(cudaGetLastError());
    end(kKStep);
    if (n_px > 0) {
      const double* div255 = nullptr;
      BE_TRY(cudaGetSymbolAddress((void**)&div255, g_div255));
      dim3 grid((P.max_w + kTileW - 1) / kTileW, (P.max_h + kTileH - 1) / kTileH, n_px);
      begin();
      step_pixel_kernel<<<grid, kTileThreads, 0, stream>>>(P, lanes, lane_px_ids, scratch, frame_bytes, div255);
      BE_TRY(cudaGetLastError());
      end(kKStepPx);
    }
    return 0;
  }
  int mix(const DevPlan& P, const MixJob* jobs, int n) {
    dim3 grid((P.max_w + kTileW - 1) / kTileW, (P.max_h + kTileH - 1) / kTileH, n);
    begin();
    mix_kernel<<<grid, kTileThreads, 0, stream>>>(P, jobs);
    BE_TRY(cudaGetLastError());
    end(kKMix);
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

extern "C" int oadg_oamix_execute_profiled(const void* plan_host, size_t plan_bytes, const uint8_t* const* src_dev,
                                           int n_img, uint8_t* const* dst_dev, void* workspace_dev,
                                           size_t workspace_bytes, float* ms_by_kind, int* count_by_kind,
                                           void* stream) {
  if (!ms_by_kind || !count_by_kind) return OADG_E_ARG;
  CudaBackend be;
  be.stream = (cudaStream_t)stream;
  be.profile = true;
  int rc = execute_plan(be, plan_host, plan_bytes, src_dev, n_img, dst_dev, workspace_dev, workspace_bytes);
  cudaError_t e = cudaStreamSynchronize(be.stream);
  for (int k = 0; k < kKinds; ++k) {
    ms_by_kind[k] = 0.f;
    count_by_kind[k] = 0;
  }
  for (auto& r : be.recs) {
    float ms = 0.f;
    if (e == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ms_by_kind[r.kind] += ms;
      ++count_by_kind[r.kind];
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  if (rc) return rc;
  return e == cudaSuccess ? 0 : (int)e;
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
