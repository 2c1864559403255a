// Spectral-residual saliency score per gt box: one CTA per box.
//
// Replaces reference oa_mix.py:107-110
//   saliency = cv2.saliency.StaticSaliencySpectralResidual_create()
//   (success, saliency_map) = saliency.computeSaliency(bbox_img)
//   saliency_score = np.mean((saliency_map * 255).astype("uint8"))
// following opencv_contrib staticSaliencySpectralResidual.cpp (computeSaliencyImpl):
// BGR2GRAY -> resize 64x64 INTER_LINEAR_EXACT -> complex DFT (f64) -> cartToPolar ->
// log -> 3x3 box blur -> exp(residual) -> polarToCart -> inverse DFT -> magnitude ->
// GaussianBlur 5x5 sigma 8 -> square -> /max -> f32 -> resize to the crop (INTER_LINEAR).
//
// Integer front end is exact; the FFT is our own radix-2 in f64; cartToPolar's
// float32 internals (sqrtf(fmaf(x,x,y*y)) magnitude and the polynomial fastAtan)
// are reproduced because they move the score by ~2e-3 otherwise.
// Spec: oracle/prims_np.py saliency_score_emul.
#include "oadg_common.cuh"

namespace oadg {
namespace {

constexpr int kRes = 64;
constexpr int kN = kRes * kRes;
constexpr int kThreads = 1024;

__device__ __forceinline__ int brev6(int v) { return (int)(__brev((unsigned)v) >> 26); }

// 64 independent 64-point FFTs over smem; element e of line l lives at l*ls + e*es.
// Thread mapping keeps consecutive threads on consecutive shared-memory words: along the element index for
// the row pass (es == 1) and along the line index for the column pass (ls == 1); the other order would put a
// whole warp on one bank (stride 64 doubles).
__device__ void fft64_lines(double* re, double* im, int es, int ls, const double* twc, const double* tws,
                            bool inverse) {
  const int tid = threadIdx.x;
  const bool lines_fast = ls == 1;
  for (int i = tid; i < kN; i += kThreads) {
    const int l = lines_fast ? (i & 63) : (i >> 6), e = lines_fast ? (i >> 6) : (i & 63), r = brev6(e);
    if (e < r) {
      int a = l * ls + e * es, b = l * ls + r * es;
      double t = re[a];
      re[a] = re[b];
      re[b] = t;
      t = im[a];
      im[a] = im[b];
      im[b] = t;
    }
  }
  __syncthreads();
  for (int s = 1; s <= 6; ++s) {
    const int half = 1 << (s - 1);
    const int tstep = 64 >> s;
    for (int i = tid; i < kN / 2; i += kThreads) {
      const int l = lines_fast ? (i & 63) : (i >> 5), b = lines_fast ? (i >> 6) : (i & 31);
      int grp = b / half, j = b - grp * half;
      int i0 = l * ls + (grp * 2 * half + j) * es;
      int i1 = i0 + half * es;
      double wr = twc[j * tstep];
      double wi = inverse ? -tws[j * tstep] : tws[j * tstep];
      double xr = re[i1], xi = im[i1];
      double tr = xr * wr - xi * wi;
      double ti = xr * wi + xi * wr;
      double ur = re[i0], ui = im[i0];
      re[i0] = ur + tr;
      im[i0] = ui + ti;
      re[i1] = ur - tr;
      im[i1] = ui - ti;
    }
    __syncthreads();
  }
}

// cv2.cartToPolar on 64F data runs in float32 on FMA hosts (probed against cv2 4.13).
__device__ __forceinline__ double cv_magnitude(double re, double im) {
  float x = (float)re, y = (float)im;
  return (double)sqrtf(fmaf(x, x, __fmul_rn(y, y)));
}
__device__ __forceinline__ double cv_fast_atan(double im, double re) {
  const float s = (float)(180.0 / 3.14159265358979323846);
  const float p1 = __fmul_rn(0.9997878412794807f, s), p3 = __fmul_rn(-0.3258083974640975f, s);
  const float p5 = __fmul_rn(0.1555786518463281f, s), p7 = __fmul_rn(-0.04432655554792128f, s);
  const float eps = 2.220446049250313e-16f;
  float x = (float)re, y = (float)im;
  float ax = fabsf(x), ay = fabsf(y), a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(fmaf(fmaf(fmaf(p7, c2, p5), c2, p3), c2, p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(fmaf(fmaf(fmaf(p7, c2, p5), c2, p3), c2, p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return (double)__fmul_rn(a, (float)(3.14159265358979323846 / 180.0));
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// INTER_LINEAR_EXACT tap for output index d of a 64-sample axis over n_in inputs.
__device__ __forceinline__ void exact_tap(int d, int n_in, int& i0, int& i1, int& a256) {
  int num = (2 * d + 1) * n_in - kRes;  // f = num / 128
  int q = num >> 7;                      // floor division by 128
  int t = num - (q << 7);
  if (q < 0) {
    q = 0;
    t = 0;
  }
  if (q >= n_in - 1) {
    q = n_in - 1;
    t = 0;
  }
  i0 = q;
  i1 = min(q + 1, n_in - 1);
  a256 = 2 * t;
}

__global__ void __launch_bounds__(kThreads)
saliency_kernel(const uint8_t* const* __restrict__ imgs, const int32_t* __restrict__ hw,
                const int32_t* __restrict__ boxes, double* __restrict__ scores) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* re = reinterpret_cast<double*>(smem_raw);
  double* im = re + kN;
  double* tmp = im + kN;
  __shared__ double twc[32], tws[32];
  __shared__ double red[kThreads / 32];
  __shared__ unsigned long long redu[kThreads / 32];

  const int tid = threadIdx.x;
  const int32_t* bx = boxes + (size_t)blockIdx.x * 5;
  const int img_i = bx[0], x1 = bx[1], y1 = bx[2], x2 = bx[3], y2 = bx[4];
  const int W = hw[img_i * 2 + 1];
  const uint8_t* img = imgs[img_i];
  const int cw = x2 - x1, ch = y2 - y1;

  if (tid < 32) {
    double s, c;
    sincospi(-2.0 * tid / 64.0, &s, &c);
    twc[tid] = c;
    tws[tid] = s;
  }
  // (a) gray + exact 64x64 resize
  for (int i = tid; i < kN; i += kThreads) {
    int oy = i >> 6, ox = i & 63;
    int xa, xb, ax, ya, yb, ay;
    exact_tap(ox, cw, xa, xb, ax);
    exact_tap(oy, ch, ya, yb, ay);
    auto gray = [&](int yy, int xx) {
      const uint8_t* p = img + ((size_t)(y1 + yy) * W + (x1 + xx)) * 3;
      return (int)((__ldg(p) * 3735 + __ldg(p + 1) * 19235 + __ldg(p + 2) * 9798 + (1 << 14)) >> 15);
    };
    int v0 = gray(ya, xa) * (256 - ax) + gray(ya, xb) * ax;
    int v1 = gray(yb, xa) * (256 - ax) + gray(yb, xb) * ax;
    int g = (v0 * (256 - ay) + v1 * ay + (1 << 15)) >> 16;
    re[i] = (double)g;
    im[i] = 0.0;
  }
  __syncthreads();
  // (b) forward 2-D DFT
  fft64_lines(re, im, 1, 64, twc, tws, false);
  fft64_lines(re, im, 64, 1, twc, tws, false);
  // (c) log-amplitude (tmp) and phase (im)
  for (int i = tid; i < kN; i += kThreads) {
    double r = re[i], q = im[i];
    tmp[i] = log(cv_magnitude(r, q));
    im[i] = cv_fast_atan(q, r);
  }
  __syncthreads();
  for (int i = tid; i < kN; i += kThreads) {
    int y = i >> 6, x = i & 63;
    double s = 0.0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      int yy = reflect101(y + dy, kRes);
      s += (tmp[yy * 64 + reflect101(x - 1, kRes)] + tmp[yy * 64 + x]) + tmp[yy * 64 + reflect101(x + 1, kRes)];
    }
    double nm = exp(tmp[i] - s * (1.0 / 9.0));
    double sn, cs;
    sincos(im[i], &sn, &cs);
    re[i] = nm * cs;
    im[i] = nm * sn;
  }
  __syncthreads();
  // (d) inverse DFT (unscaled) and magnitude
  fft64_lines(re, im, 1, 64, twc, tws, true);
  fft64_lines(re, im, 64, 1, twc, tws, true);
  for (int i = tid; i < kN; i += kThreads) tmp[i] = cv_magnitude(re[i], im[i]);
  __syncthreads();
  // (e) GaussianBlur 5x5, sigma 8, BORDER_REFLECT_101 (separable, f64)
  double k0, k1, k2;
  {
    double e1 = exp(-0.5 / 64.0 * 1.0), e2 = exp(-0.5 / 64.0 * 4.0);
    double sum = 1.0 + 2.0 * e1 + 2.0 * e2;
    k0 = 1.0 / sum;
    k1 = e1 / sum;
    k2 = e2 / sum;
  }
  for (int i = tid; i < kN; i += kThreads) {
    int y = i >> 6, x = i & 63;
    const double* row = tmp + y * 64;
    re[i] = row[x] * k0 + (row[reflect101(x - 1, kRes)] + row[reflect101(x + 1, kRes)]) * k1 +
            (row[reflect101(x - 2, kRes)] + row[reflect101(x + 2, kRes)]) * k2;
  }
  __syncthreads();
  double lmax = -1.0;
  bool has_nan = false;
  for (int i = tid; i < kN; i += kThreads) {
    int y = i >> 6, x = i & 63;
    double v = re[y * 64 + x] * k0 +
               (re[reflect101(y - 1, kRes) * 64 + x] + re[reflect101(y + 1, kRes) * 64 + x]) * k1 +
               (re[reflect101(y - 2, kRes) * 64 + x] + re[reflect101(y + 2, kRes) * 64 + x]) * k2;
    v = v * v;
    im[i] = v;
    if (v != v) has_nan = true;
    else lmax = fmax(lmax, v);
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) red[tid >> 5] = lmax;
  int any_nan = __syncthreads_or(has_nan ? 1 : 0);
  double vmax = red[0];
#pragma unroll
  for (int w = 1; w < kThreads / 32; ++w) vmax = fmax(vmax, red[w]);
  // (f) normalised float32 map (kept in smem over the `re` region)
  float* map = reinterpret_cast<float*>(re);
  __syncthreads();
  for (int i = tid; i < kN; i += kThreads) map[i] = (float)(im[i] / vmax);
  __syncthreads();
  // (g) bilinear resize to the crop (cv2 generic f32 path), *255, truncate, mean
  unsigned long long acc = 0;
  if (!any_nan) {
    const double sx = 64.0 / (double)cw, sy = 64.0 / (double)ch;
    const int total = cw * ch;
    for (int i = tid; i < total; i += kThreads) {
      int dy = i / cw, dx = i - dy * cw;
      float fx = (float)((dx + 0.5) * sx - 0.5), fy = (float)((dy + 0.5) * sy - 0.5);
      int ix = (int)floorf(fx), iy = (int)floorf(fy);
      fx = __fsub_rn(fx, (float)ix);
      fy = __fsub_rn(fy, (float)iy);
      if (ix < 0) { ix = 0; fx = 0.f; }
      if (ix >= kRes - 1) { ix = kRes - 1; fx = 0.f; }
      if (iy < 0) { iy = 0; fy = 0.f; }
      if (iy >= kRes - 1) { iy = kRes - 1; fy = 0.f; }
      int ix1 = min(ix + 1, kRes - 1), iy1 = min(iy + 1, kRes - 1);
      float ax = __fsub_rn(1.f, fx), ay = __fsub_rn(1.f, fy);
      float r0 = __fadd_rn(__fmul_rn(map[iy * 64 + ix], ax), __fmul_rn(map[iy * 64 + ix1], fx));
      float r1 = __fadd_rn(__fmul_rn(map[iy1 * 64 + ix], ax), __fmul_rn(map[iy1 * 64 + ix1], fx));
      float v = __fadd_rn(__fmul_rn(r0, ay), __fmul_rn(r1, fy));
      acc += (unsigned long long)(unsigned)(int)__fmul_rn(v, 255.f);
    }
  }
  acc = warp_sum(acc);
  if ((tid & 31) == 0) redu[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += redu[w];
    // a NaN map casts to 0 everywhere on the reference host (x86 cvttss2si low byte)
    scores[blockIdx.x] = any_nan ? 0.0 : (double)t / (double)((long long)cw * ch);
  }
}

}  // namespace
}  // namespace oadg

extern "C" int oadg_saliency_scores(const uint8_t* const* imgs_dev, const int32_t* hw_dev,
                                    const int32_t* boxes_dev, int n_boxes, double* scores_dev,
                                    void* stream) {
  if (n_boxes < 0) return OADG_E_ARG;
  if (n_boxes == 0) return 0;
  if (!imgs_dev || !hw_dev || !boxes_dev || !scores_dev) return OADG_E_ARG;
  const size_t smem = 3 * oadg::kN * sizeof(double);
  static bool attr_of[64] = {false};   // per device
  const int slot = oadg::device_slot();
  if (!attr_of[slot]) {
    OADG_CUDA_TRY(cudaFuncSetAttribute(oadg::saliency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    attr_of[slot] = true;
  }
  oadg::saliency_kernel<<<n_boxes, oadg::kThreads, smem, (cudaStream_t)stream>>>(imgs_dev, hw_dev, boxes_dev,
                                                                                   scores_dev);
  OADG_LAUNCH_CHECK();
  return 0;
}
