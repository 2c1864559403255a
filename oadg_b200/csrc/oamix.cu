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
__global__ void __launch_bounds__(256)
bbo_pass_kernel(DevPlan P, const Chain* __restrict__ chains, int j) {
  const Chain C = chains[blockIdx.z];
  if (j >= C.n) return;
  int r[4];
  bbo_pass_rect(P, C, j, r);
  const int x = r[0] + blockIdx.x * 32 + threadIdx.x;
  const int y = r[1] + blockIdx.y * 8 + threadIdx.y;
  if (x >= r[2] || y >= r[3]) return;
  bbo_pixel(P, C, j, x, y);
}

// ------------------------------------------------------------------------------------
// one depth step of every live lane (oa_mix.py:226-234).  A CTA owns a 256 x 16 pixel tile and one Lane record.
//   streaming tile (one LUT / bbo-copy region covers it): one 16-pixel chunk (3 x 16-byte vectors) per thread,
//       loads issued before the region's 3 x 256 LUT is staged in shared memory
//   any other tile: 16 pixels per thread, consecutive lanes on consecutive pixels (gathers stay within a few lines)
// grid = (ceil(W/256), ceil(H/16), lanes)
// ------------------------------------------------------------------------------------
constexpr int kTileThreads = 256;

__global__ void __launch_bounds__(kTileThreads, 4)
step_kernel(DevPlan P, const Lane* __restrict__ lanes, const uint8_t* __restrict__ scratch, size_t frame_bytes) {
  __shared__ __align__(16) uint8_t lut_s[768];
  __shared__ Lane Ls;
  if (threadIdx.x < sizeof(Lane) / 4)
    reinterpret_cast<uint32_t*>(&Ls)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t*>(lanes + blockIdx.z) + threadIdx.x);
  __syncthreads();
  const Lane& L = Ls;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  if (x0 >= L.W || y0 >= L.H) return;
  const int x1 = min(x0 + kTileW, L.W), y1 = min(y0 + kTileH, L.H);
  const int region = tile_region(L, x0, y0, x1, y1);
  const int t = threadIdx.x;
  if (tile_streams(L, region)) {
    const bool vec = ((L.W * 3) & 15) == 0 &&
                     ((((uintptr_t)L.in) | ((uintptr_t)L.out) | ((uintptr_t)scratch) | frame_bytes) & 15) == 0;
    const int y = y0 + (t >> 4), x = x0 + (t & 15) * kChunkPx;
    const bool live = y < y1 && x < x1;
    const int n = live ? min(kChunkPx, x1 - x) : 0;
    Chunk in;
    if (live) chunk_load(stream_src(L, region, scratch, frame_bytes) + ((size_t)y * L.W + x) * 3, n, vec, in);
    if (L.lut[region] >= 0) {  // uniform over the CTA
      const uint32_t* src = reinterpret_cast<const uint32_t*>(P.luts + (size_t)L.lut[region] * 768);
      if (t < 192) reinterpret_cast<uint32_t*>(lut_s)[t] = __ldg(src + t);
      __syncthreads();
    }
    if (live) stream_chunk(L, region, lut_s, scratch, frame_bytes, in, x, y, n, vec);
    return;
  }
#pragma unroll 1
  for (int it = 0; it < kTileW * kTileH / kTileThreads; ++it) {
    const int idx = it * kTileThreads + t;
    const int x = x0 + (idx & (kTileW - 1)), y = y0 + idx / kTileW;
    if (x < x1 && y < y1) step_pixel(P, L, scratch, frame_bytes, x, y);
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
  const int y = y0 + (t >> 4), x = x0 + (t & 15) * kChunkPx;
  if (x >= x1 || y >= y1) return;
  mix_chunk(P, J, T, x, y, min(kChunkPx, x1 - x), vec);
}

#define BE_TRY(expr)                       \
  do {                                     \
    cudaError_t _e = (expr);               \
    if (_e != cudaSuccess) return (int)_e; \
  } while (0)

enum { kKProfile = 0, kKHist, kKLut, kKBboPass, kKMask, kKStep, kKMix, kKCopy, kKinds };

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
    bbo_pass_kernel<<<dim3((roi_w + 31) / 32, (roi_h + 7) / 8, n), dim3(32, 8), 0, stream>>>(P, chains, j);
    BE_TRY(cudaGetLastError());
    end(kKBboPass);
    return 0;
  }
  int step(const DevPlan& P, const Lane* lanes, int n, const uint8_t* scratch, size_t frame_bytes) {
    dim3 grid((P.max_w + kTileW - 1) / kTileW, (P.max_h + kTileH - 1) / kTileH, n);
    begin();
    step_kernel<<<grid, kTileThreads, 0, stream>>>(P, lanes, scratch, frame_bytes);
    BE_TRY(cudaGetLastError());
    end(kKStep);
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
