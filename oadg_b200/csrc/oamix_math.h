// Per-pixel arithmetic of the OA-Mix kernels, shared by the CUDA kernels and the
// host-side arithmetic check (tests/hostsim, test infrastructure only).
//
// Every function states the reference expression it reproduces.  Float/double
// expressions are written with explicit un-fused multiplies and adds
// (__fmul_rn/__dadd_rn on the device, -ffp-contract=off on the host) because the
// reference evaluates them as separate NumPy ufuncs: a fused multiply-add would
// change the value that gets truncated to uint8.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define OADG_HD __host__ __device__ __forceinline__
#else
#define OADG_HD inline
#endif

namespace oadg {

// ---- un-fused scalar arithmetic -------------------------------------------------
OADG_HD float fmul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
OADG_HD float fadd(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
OADG_HD float fsub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
// the one place where the reference's arithmetic IS fused: OpenCV's vectorised column filter (v_fma)
OADG_HD float ffma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
// Value of cv2.GaussianBlur on a window that lies entirely inside the box (all taps read 1.0), as OpenCV computes it
// in float32 (checked against cv2 4.13 on 700 sigma pairs): the row filter adds the taps in order (kernels of 3 and
// 5 taps take the symmetric form below instead), the symmetric column filter starts at the centre tap and adds k[c+j] * (v + v) outward with fused
// multiply-adds.  The result is 1 - 2^-24, 1 or 1 + 2^-23 depending on the kernel, and where the mask saturates the
// blend img*(1-m) + aug*m sits on an integer, so this last bit decides the truncation of EVERY pixel there.
OADG_HD float cv_saturated_col(const float* k, int ks, float v);
OADG_HD float cv_saturated_row(const float* k, int ks) {
  if (ks <= 5) return cv_saturated_col(k, ks, 1.0f);   // OpenCV's small symmetric row filter: same order as its column filter
  float s = k[0];
  for (int j = 1; j < ks; ++j) s = fadd(s, k[j]);
  return s;
}
OADG_HD float cv_saturated_col(const float* k, int ks, float v) {
  const int c = ks / 2;
  float s = fmul(k[c], v);
  const float v2 = fadd(v, v);
  for (int j = 1; j <= c; ++j) s = ffma(k[c + j], v2, s);
  return s;
}
OADG_HD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
OADG_HD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
OADG_HD double dsub(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
// cvRound(double): round half to even, as saturate_cast<int>(double) in OpenCV.
OADG_HD int cv_round(double v) {
#ifdef __CUDA_ARCH__
  return __double2int_rn(v);
#else
  return (int)lrint(v);
#endif
}
// ---- conversions between small non-negative integers and floats -------------------------------------------
// On the device the I2F / F2I / I2D instructions run on the quarter-rate pipe; for 0 <= v < 2^23 the same values
// come out of one logic op + one add: float(v) == as_float(0x4B000000 | v) - 2^23, and for a float 0 <= f < 2^23
// int(f) (truncation) == low bits of (f + 2^23) rounded toward -inf; likewise with 2^52 in float64.
OADG_HD float u8_to_f32(int v) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(__int_as_float(0x4B000000 | v), 8388608.0f);
#else
  return (float)v;
#endif
}
OADG_HD double u8_to_f64(int v) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(__hiloint2double(0x43300000, v), 4503599627370496.0);
#else
  return (double)v;
#endif
}
OADG_HD int f32_trunc_u8(float f) {   // 0 <= f < 2^23
#ifdef __CUDA_ARCH__
  return __float_as_int(__fadd_rd(f, 8388608.0f)) & 0x7FFFFF;
#else
  return (int)f;
#endif
}
OADG_HD int f64_trunc_u8(double d) {   // 0 <= d < 2^31
#ifdef __CUDA_ARCH__
  return __double2loint(__dadd_rd(d, 4503599627370496.0));
#else
  return (int)d;
#endif
}
OADG_HD int imin(int a, int b) { return a < b ? a : b; }
OADG_HD int imax(int a, int b) { return a > b ? a : b; }

// ---- cv2.warpAffine(8U, INTER_LINEAR, BORDER_CONSTANT 0) -------------------------
// OpenCV imgwarp.cpp WarpAffineInvoker: AB_BITS = 10, INTER_BITS = 5; the matrix is
// the INVERSE map (dst -> src) in doubles, inverted on the host exactly like
// cv::warpAffine does.  Reference call sites: augmix.py:92,116,136,156,177.
struct WarpRow {
  int X0, Y0;
};
OADG_HD WarpRow warp_row(const double* m, int y) {
  WarpRow r;
  r.X0 = cv_round(dmul(dadd(dmul(m[1], (double)y), m[2]), 1024.0)) + 16;
  r.Y0 = cv_round(dmul(dadd(dmul(m[4], (double)y), m[5]), 1024.0)) + 16;
  return r;
}
struct WarpTap {
  int sx, sy, fx, fy;
};
OADG_HD WarpTap warp_px(const double* m, WarpRow r, int x) {
  int X = (r.X0 + cv_round(dmul(dmul(m[0], (double)x), 1024.0))) >> 5;
  int Y = (r.Y0 + cv_round(dmul(dmul(m[3], (double)x), 1024.0))) >> 5;
  WarpTap t;
  t.sx = imin(imax(X >> 5, -32768), 32767);  // saturate_cast<short>
  t.sy = imin(imax(Y >> 5, -32768), 32767);
  t.fx = X & 31;
  t.fy = Y & 31;
  return t;
}
// remapBilinear fixed point: weights (32-fx)(32-fy)*32 ... sum to 32768, round at 1<<14.
OADG_HD int bilerp_fix(int v00, int v01, int v10, int v11, int fx, int fy) {
  int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
  int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  return (v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11 + (1 << 14)) >> 15;
}

// One warped u8x3 pixel from a tightly packed HWC image (out-of-frame taps = 0).
template <typename Ld>
OADG_HD void warp_fetch3(Ld ld, const uint8_t* img, int H, int W, WarpTap t, int out[3]) {
  // a tap whose weight is zero (fx == 0 / fy == 0: translations, one axis of the shears) is never read
  const bool x0 = (unsigned)t.sx < (unsigned)W, x1 = t.fx != 0 && (unsigned)(t.sx + 1) < (unsigned)W;
  const bool y0 = (unsigned)t.sy < (unsigned)H, y1 = t.fy != 0 && (unsigned)(t.sy + 1) < (unsigned)H;
  const uint8_t* r0 = img + ((size_t)t.sy * W + t.sx) * 3;
  const uint8_t* r1 = r0 + (size_t)W * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int v00 = (y0 && x0) ? ld(r0 + c) : 0;
    int v01 = (y0 && x1) ? ld(r0 + 3 + c) : 0;
    int v10 = (y1 && x0) ? ld(r1 + c) : 0;
    int v11 = (y1 && x1) ? ld(r1 + 3 + c) : 0;
    out[c] = bilerp_fix(v00, v01, v10, v11, t.fx, t.fy);
  }
}

// ---- blends -----------------------------------------------------------------------
// bbox_augmentation.py:63-71: mask = 1.0 - blur_bbox (f32);
//   img = uint8(img * mask + aug * (1.0 - mask))            all float32
OADG_HD int bbo_blend(float m, int img, int aug) {
  float mask = fsub(1.0f, m);
  float v = fadd(fmul((float)img, mask), fmul((float)aug, fsub(1.0f, mask)));
  return (int)v;  // astype(uint8) on [0,255]: truncation
}
// bbox_augmentation.py:264-272: augmented_mask = warp(uint8(mask*255)) / 255 (f64);
//   keep = maximum(mask_f32, augmented_mask) (f64); img = uint8(keep*img + (1-keep)*aug)
OADG_HD int mask_to_u8(float m) { return (int)fmul(m, 255.0f); }
OADG_HD int bg_blend(float mask, int warped_mask_u8, int img, int aug) {
  double am = (double)warped_mask_u8 / 255.0;
  double keep = (double)mask > am ? (double)mask : am;
  double v = dadd(dmul(keep, (double)img), dmul(dsub(1.0, keep), (double)aug));
  return (int)v;
}

// ---- PIL.ImageOps LUTs (augmix.py:64-75,103-105) ----------------------------------
// hist: 256 counts of one channel.  out: 256 entries.
OADG_HD void lut_autocontrast_ch(const unsigned* hist, uint8_t* out) {
  int lo = 0, hi = 255;
  for (lo = 0; lo < 256; ++lo)
    if (hist[lo]) break;
  for (hi = 255; hi >= 0; --hi)
    if (hist[hi]) break;
  if (hi <= lo) {
    for (int i = 0; i < 256; ++i) out[i] = (uint8_t)i;
    return;
  }
  double scale = 255.0 / (double)(hi - lo);
  double offset = dmul((double)(-lo), scale);
  for (int i = 0; i < 256; ++i) {
    int v = (int)dadd(dmul((double)i, scale), offset);
    out[i] = (uint8_t)imin(imax(v, 0), 255);
  }
}
OADG_HD void lut_equalize_ch(const unsigned* hist, uint8_t* out) {
  unsigned total = 0, last = 0;
  int nnz = 0;
  for (int i = 0; i < 256; ++i)
    if (hist[i]) {
      total += hist[i];
      last = hist[i];
      ++nnz;
    }
  unsigned step = nnz <= 1 ? 0u : (total - last) / 255u;
  if (!step) {
    for (int i = 0; i < 256; ++i) out[i] = (uint8_t)i;
    return;
  }
  unsigned n = step / 2;
  for (int i = 0; i < 256; ++i) {
    unsigned v = n / step;
    out[i] = (uint8_t)(v > 255u ? 255u : v);  // Image.point clips list entries to 8 bits
    n += hist[i];
  }
}
OADG_HD uint8_t lut_posterize_at(int i, int bits) { return (uint8_t)(i & ~((1 << (8 - bits)) - 1)); }
OADG_HD uint8_t lut_solarize_at(int i, int thr) { return (uint8_t)(i < thr ? i : 255 - i); }

// ---- PIL.ImageEnhance pieces (augmix.py:192-212) ----------------------------------
// convert('L') on the buffer handed to Image.fromarray(..., 'RGB'): channel 0 is "R".
OADG_HD int pil_luma(int c0, int c1, int c2) { return (19595 * c0 + 38470 * c1 + 7471 * c2 + 0x8000) >> 16; }
// Image.blend(degenerate, image, factor): out = deg + factor*(img - deg) in float32;
// factor in [0,1] -> plain truncation, otherwise clip to [0,255] first.
OADG_HD int pil_blend(int deg, int img, float factor) {
  float t = fadd((float)deg, fmul(factor, (float)(img - deg)));
  if (factor >= 0.0f && factor <= 1.0f) return (int)t;
  if (t <= 0.0f) return 0;
  if (t >= 255.0f) return 255;
  return (int)t;
}
// ImageFilter.SMOOTH interior pixel: kernel (1,1,1,1,5,1,1,1,1)/13, float32 sum + 0.5, clip.
OADG_HD int pil_smooth9(const int v[9]) {
  const float k1 = (float)(1.0 / 13.0), k5 = (float)(5.0 / 13.0);
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; ++i) s = fadd(s, fmul((float)v[i], i == 4 ? k5 : k1));
  s = fadd(s, 0.5f);
  if (s <= 0.0f) return 0;
  if (s >= 255.0f) return 255;
  return (int)s;
}

// ---- object-aware mixing (oa_mix.py:281-309), one channel of one pixel -----------
struct MixState {
  float orig, aug;  // per channel accumulators live in the caller; see mix_target()
};
// per-pixel target bookkeeping shared by the 3 channels
struct MixMask {
  float sum, mx;
};
// returns the weight (mask - overlap*0.5) for this target and advances sum/max
OADG_HD float mix_target_weight(MixMask& s, float mask) {
  s.sum = fadd(s.sum, mask);                   // mask_sum += mask
  s.mx = mask > s.mx ? mask : s.mx;            // np.max(mask_max_list)
  float overlap = fsub(s.sum, s.mx);           // mask_overlap = mask_sum - mask_max
  float w = fsub(mask, fmul(overlap, 0.5f));   // (mask - mask_overlap * 0.5)
  s.sum = s.mx;                                // mask_sum = mask_max
  return w;
}
// orig += (1.0 - m_oa) * img * w ;  aug += m_oa * img_aug * w        (float32)
OADG_HD void mix_accumulate(float& orig, float& aug, float m_oa, int img, float img_aug, float w) {
  orig = fadd(orig, fmul(fmul(fsub(1.0f, m_oa), u8_to_f32(img)), w));
  aug = fadd(aug, fmul(fmul(m_oa, img_aug), w));
}
// img_oamix = orig + aug; += (1.0-m)*img*(1.0-mask_sum) [f64 term]; += m*img_aug*(1.0-mask_sum) [f32];
// clip(0,255); uint8 truncation.   m is a python float (double).
OADG_HD int mix_finish(float orig, float aug, double m, int img, float img_aug, float mask_sum) {
  float out = fadd(orig, aug);
  float rest = fsub(1.0f, mask_sum);
  double t1 = dmul(dmul(dsub(1.0, m), u8_to_f64(img)), (double)rest);
  out = (float)dadd((double)out, t1);
  float t2 = fmul(fmul((float)m, img_aug), rest);
  out = fadd(out, t2);
  if (!(out > 0.0f)) out = 0.0f;
  if (out > 255.0f) out = 255.0f;
  return f32_trunc_u8(out);
}

}  // namespace oadg
