// Host-side OA-Mix plan sampler + packer (no CUDA): everything random about a batch of OAMix.oamix calls
// (reference oa_mix.py:207-262,281-298) drawn from the CALLER'S random stream in the reference's draw order
//  and written straight into the plan blob oadg_oamix_execute consumes.
//
// The stream is the caller's: oadg_rng_t carries the next_uint32 / next_double entry points of the generator
// (for the reference that is numpy's global legacy RandomState, whose MT19937 bit generator exports exactly these
// two functions), so a seeded run consumes np.random draw for draw like the reference and leaves it in the same
// state.  The distributions on top are numpy's *legacy* algorithms, restated here:
//   random_sample  = next_double
//   randint(lo,hi) = lo + masked rejection of next_uint32 on [0, hi-lo-1]      (_bounded_integers, use_masked)
//   dirichlet(1..) = standard_exponential draws (-log(1 - u)) normalised        (legacy_standard_gamma(shape=1))
//   beta(1,1)      = Johnk's algorithm                                          (legacy_beta, a,b <= 1)
// Float32 IoU tests reproduce NumPy's float32 arithmetic and pairwise summation order.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "oadg.h"

namespace {

struct Rng {
  const oadg_rng_t* r;
  double u01() const { return r->next_double(r->state); }
  // np.random.randint(lo, hi) with the default int64 dtype: range hi-lo-1 fits 32 bits here
  int64_t randint(int64_t lo, int64_t hi) const {
    const uint64_t rng = (uint64_t)(hi - 1 - lo);
    if (rng == 0) return lo;
    if (rng <= 0xFFFFFFFFull) {
      if (rng == 0xFFFFFFFFull) return lo + (int64_t)r->next_uint32(r->state);
      uint32_t mask = (uint32_t)rng;
      mask |= mask >> 1;
      mask |= mask >> 2;
      mask |= mask >> 4;
      mask |= mask >> 8;
      mask |= mask >> 16;
      uint32_t v;
      while ((v = (r->next_uint32(r->state) & mask)) > (uint32_t)rng) {
      }
      return lo + (int64_t)v;
    }
    // 64-bit ranges never occur for frame coordinates; numpy would draw next_uint64 here
    uint64_t mask = rng;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    mask |= mask >> 32;
    uint64_t v;
    do {
      const uint64_t hi32 = r->next_uint32(r->state), lo32 = r->next_uint32(r->state);
      v = ((hi32 << 32) | lo32) & mask;
    } while (v > rng);
    return lo + (int64_t)v;
  }
};

// NumPy's FLOAT_pairwise_sum (loops_utils.h.src) for a contiguous float32 vector
float pairwise_sum_f32(const float* a, size_t n) {
  if (n < 8) {
    float res = -0.0f;
    for (size_t i = 0; i < n; ++i) res += a[i];
    return res;
  }
  if (n <= 128) {
    float r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    size_t i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  size_t n2 = n / 2;
  n2 -= n2 % 8;
  return pairwise_sum_f32(a, n2) + pairwise_sum_f32(a + n2, n - n2);
}

// float32 IoU of one box against k boxes (core/evaluation/bbox_overlaps.py:5-65), eps 1e-6
void iou_1xk(const float b1[4], const float* b2, int k, float* out) {
  const float a1 = (b1[2] - b1[0]) * (b1[3] - b1[1]);
  for (int i = 0; i < k; ++i) {
    const float* q = b2 + 4 * i;
    const float a2 = (q[2] - q[0]) * (q[3] - q[1]);
    float w = fminf(b1[2], q[2]) - fmaxf(b1[0], q[0]);
    float h = fminf(b1[3], q[3]) - fmaxf(b1[1], q[1]);
    w = w > 0.f ? w : 0.f;
    h = h > 0.f ? h : 0.f;
    const float ov = w * h;
    float uni = a1 + a2 - ov;
    uni = uni > 1e-6f ? uni : 1e-6f;
    out[i] = ov / uni;
  }
}

struct Box {
  int64_t v[4];
};

// OAMix.get_random_regions (oa_mix.py:122-184).  num_lo/num_hi: randint(*num) when num_hi > num_lo, else the count.
void sample_regions(const Rng& R, int h, int w, const double scale[2], const double ratio[2], int num_lo, int num_hi,
                    const float* gt, int n_gt, const double* scores, bool with_gt, std::vector<Box>& boxes,
                    std::vector<double>& bscores) {
  const int64_t target = num_hi > num_lo ? R.randint(num_lo, num_hi) : num_lo;
  const double s_lo = scale[0], s_w = scale[1] - scale[0];
  const double r_lo = ratio[0], r_w = ratio[1] - ratio[0];
  boxes.clear();
  bscores.clear();
  std::vector<float> acc, ious;
  for (int it = 0; it < 50; ++it) {
    if ((int64_t)boxes.size() >= target) break;
    const int64_t x1 = R.randint(0, w), y1 = R.randint(0, h);
    const double area = (s_lo + s_w * R.u01()) * h * w;
    const double r = r_lo + r_w * R.u01();
    const int64_t bw = (int64_t)sqrt(area / r), bh = (int64_t)sqrt(area * r);
    if (x1 + bw > w || y1 + bh > h) continue;
    Box b;
    b.v[0] = x1;
    b.v[1] = y1;
    b.v[2] = x1 + bw < w ? x1 + bw : w;
    b.v[3] = y1 + bh < h ? y1 + bh : h;
    const float bf[4] = {(float)b.v[0], (float)b.v[1], (float)b.v[2], (float)b.v[3]};
    if (!boxes.empty()) {
      acc.resize(boxes.size() * 4);
      for (size_t i = 0; i < boxes.size(); ++i)
        for (int e = 0; e < 4; ++e) acc[i * 4 + e] = (float)boxes[i].v[e];
      ious.resize(boxes.size());
      iou_1xk(bf, acc.data(), (int)boxes.size(), ious.data());
      if (0.0f + pairwise_sum_f32(ious.data(), ious.size()) > 1e-6f) continue;
    }
    if (with_gt) {
      ious.resize(n_gt > 0 ? n_gt : 1);
      iou_1xk(bf, gt, n_gt, ious.data());
      double s = INFINITY;
      if (n_gt > 0 && 0.0f + pairwise_sum_f32(ious.data(), (size_t)n_gt) > 1e-6f) {
        for (int i = 0; i < n_gt; ++i) {
          const float* fb = gt + 4 * i;
          if (ious[i] == 0.0f || fb[2] - fb[0] < 1.f || fb[3] - fb[1] < 1.f) continue;
          if (scores[i] < s) s = scores[i];
        }
      }
      bscores.push_back(s);
    }
    boxes.push_back(b);
  }
}

void invert_affine(double m[6]) {  // forward 2x3 -> inverse map in doubles, as cv::warpAffine computes it
  double D = m[0] * m[4] - m[1] * m[3];
  D = D != 0 ? 1.0 / D : 0.0;
  const double A11 = m[4] * D, A22 = m[0] * D;
  m[0] = A11;
  m[1] *= -D;
  m[3] *= -D;
  m[4] = A22;
  const double b1 = -m[0] * m[2] - m[1] * m[5];
  const double b2 = -m[3] * m[2] - m[4] * m[5];
  m[2] = b1;
  m[5] = b2;
}

enum Geo { ROTATE, SHEAR_X, SHEAR_Y, TRANSLATE_X, TRANSLATE_Y };

// 2x3 forward matrix as augmix.py:83-188 builds it (float32 entries except rotate)
void forward_affine(Geo geo, double level, bool neg, double size_lvl_w, double size_lvl_h, bool has_center,
                    double cx, double cy, int img_w, int img_h, double m[6]) {
  if (geo == ROTATE) {
    int deg = (int)(level * 30 / 10);
    if (neg) deg = -deg;
    if (!has_center) {
      cx = img_w / 2.0;
      cy = img_h / 2.0;
    }
    // cv2.getRotationMatrix2D(center, deg, 1.0): Point2f centre, double math
    const double fx = (double)(float)cx, fy = (double)(float)cy;
    const double a = deg * (M_PI / 180.0);
    const double al = cos(a), be = sin(a);
    m[0] = al; m[1] = be; m[2] = (1 - al) * fx - be * fy;
    m[3] = -be; m[4] = al; m[5] = be * fx + (1 - al) * fy;
    return;
  }
  if (geo == SHEAR_X || geo == SHEAR_Y) {
    double l = level * 0.3 / 10.;
    if (neg) l = -l;
    if (geo == SHEAR_X) {
      const double tx = has_center ? -l * cy : 0.0;
      m[0] = 1.0; m[1] = (double)(float)(-l); m[2] = has_center ? (double)(float)(-tx) : 0.0;
      m[3] = 0.0; m[4] = 1.0; m[5] = 0.0;
    } else {
      const double ty = has_center ? -l * cx : 0.0;
      m[0] = 1.0; m[1] = 0.0; m[2] = 0.0;
      m[3] = (double)(float)(-l); m[4] = 1.0; m[5] = has_center ? (double)(float)(-ty) : 0.0;
    }
    return;
  }
  if (geo == TRANSLATE_X) {
    int64_t l = (int64_t)(level * (size_lvl_w / 3) / 10);
    if (neg) l = -l;
    m[0] = 1.0; m[1] = 0.0; m[2] = (double)(float)(-l);
    m[3] = 0.0; m[4] = 1.0; m[5] = 0.0;
    return;
  }
  int64_t l = (int64_t)(level * (size_lvl_h / 3) / 10);
  if (neg) l = -l;
  m[0] = 1.0; m[1] = 0.0; m[2] = 0.0;
  m[3] = 0.0; m[4] = 1.0; m[5] = (double)(float)(-l);
}

// the reference aug lists (oa_mix.py:15-29), same order => same np.random.choice index
enum Name { AUTOCONTRAST, EQUALIZE, POSTERIZE, SOLARIZE, INVERT, COLOR, CONTRAST, BRIGHTNESS, SHARPNESS,
            BBO_ROTATE, BBO_SHEAR_XY, BBO_TRANSLATE_XY, BG_ROTATE, BG_SHEAR_XY, BG_TRANSLATE_XY };
const Name kAugmix[] = {AUTOCONTRAST, EQUALIZE, POSTERIZE, SOLARIZE, BBO_ROTATE, BBO_SHEAR_XY, BBO_TRANSLATE_XY,
                        BG_ROTATE, BG_SHEAR_XY, BG_TRANSLATE_XY};
const Name kAugmixAll[] = {AUTOCONTRAST, EQUALIZE, POSTERIZE, SOLARIZE, INVERT, COLOR, CONTRAST, BRIGHTNESS, SHARPNESS,
                           BBO_ROTATE, BBO_SHEAR_XY, BBO_TRANSLATE_XY, BG_ROTATE, BG_SHEAR_XY, BG_TRANSLATE_XY};

inline size_t a8(size_t n) { return (n + 7) / 8 * 8; }

// python slice(lo, hi).indices(n) -> (start, max(stop, start))
inline void py_slice(int64_t lo, int64_t hi, int64_t n, int32_t& s, int32_t& e) {
  auto fix = [n](int64_t v) {
    if (v < 0) {
      v += n;
      if (v < 0) v = 0;
    } else if (v > n) {
      v = n;
    }
    return v;
  };
  const int64_t a = fix(lo), b = fix(hi);
  s = (int32_t)a;
  e = (int32_t)(b > a ? b : a);
}

struct ViewDraw {
  int h, w, n_gt;
  float ws[OADG_MAX_WIDTH];
  std::vector<Box> ml, oa;
  std::vector<double> oa_scores;
  int depth[OADG_MAX_WIDTH];
  std::vector<oadg_op_t> ops;          // OPS_PER_VIEW records (unused slots zero)
  std::vector<oadg_bbo_t> bbo;         // gt field = LOCAL gt index until packed
  std::vector<int> low;                // gt boxes with score <= thresh
  double m;
  std::vector<float> m_oa;
};

}  // namespace

extern "C" int oadg_oamix_sample_plan(const oadg_rng_t* rng, const oadg_sampler_cfg_t* cfg, int n_img,
                                      const int32_t* hw, const float* const* gt, const int32_t* n_gt,
                                      const double* const* scores, void* plan_out, size_t plan_cap,
                                      size_t* plan_bytes, int64_t* ml_boxes_out, int32_t* n_ml_out,
                                      int64_t* oa_boxes_out, int32_t* n_oa_out, int32_t* depth_sum_out) {
  if (!rng || !rng->next_uint32 || !rng->next_double || !cfg || n_img < 0 || !plan_bytes) return OADG_E_ARG;
  if (n_img > 0 && (!hw || !gt || !n_gt || !scores || !n_ml_out || !n_oa_out || !ml_boxes_out || !oa_boxes_out))
    return OADG_E_ARG;
  if (cfg->mixture_width < 1 || cfg->mixture_width > OADG_MAX_WIDTH || cfg->mixture_depth > OADG_MAX_DEPTH)
    return OADG_E_LIMIT;
  if (cfg->spatial_ratio != 4) return OADG_E_LIMIT;
  if (cfg->version != 0 && cfg->version != 1) return OADG_E_ARG;
  const Name* aug = cfg->version == 0 ? kAugmix : kAugmixAll;
  const int n_aug = cfg->version == 0 ? 10 : 15;
  const Rng R{rng};
  const double sev = (double)cfg->severity;
  const int OPS_PER_VIEW = OADG_MAX_WIDTH * OADG_MAX_DEPTH * OADG_MAX_REGIONS;
  std::vector<ViewDraw> views((size_t)n_img);
  std::vector<double> dummy_scores;

  auto level_sign = [&](double& level, bool& neg) {  // augmix.py:61 sample_level, then one sign draw
    level = 0.1 + (sev - 0.1) * R.u01();
    neg = R.u01() > 0.5;
  };

  for (int v = 0; v < n_img; ++v) {
    ViewDraw& D = views[v];
    D.h = hw[2 * v];
    D.w = hw[2 * v + 1];
    D.n_gt = n_gt[v];
    const float* g = gt[v];
    if (D.h <= 0 || D.w <= 0 || D.n_gt < 0 || (D.n_gt > 0 && (!g || !scores[v]))) return OADG_E_ARG;
    // ---- head: oa_mix.py:212-234
    {  // np.float32(np.random.dirichlet([1.0] * width))
      double val[OADG_MAX_WIDTH], acc = 0.0;
      for (int j = 0; j < cfg->mixture_width; ++j) {
        val[j] = -log(1.0 - R.u01());
        acc = acc + val[j];
      }
      const double inv = 1 / acc;
      for (int j = 0; j < OADG_MAX_WIDTH; ++j) D.ws[j] = j < cfg->mixture_width ? (float)(val[j] * inv) : 0.f;
    }
    std::vector<double> unused;
    sample_regions(R, D.h, D.w, cfg->random_box_scale, cfg->random_box_ratio, 1, 3, nullptr, 0, nullptr, false, D.ml,
                   unused);
    n_ml_out[v] = (int32_t)D.ml.size();
    if (D.ml.empty()) {  // the reference raises ValueError from np.stack([]) here (oa_mix.py:217)
      *plan_bytes = 0;
      return OADG_E_NOBOX;
    }
    if (D.ml.size() > 2) return OADG_E_LIMIT;
    for (size_t k = 0; k < D.ml.size(); ++k)
      for (int e = 0; e < 4; ++e) ml_boxes_out[((size_t)v * 2 + k) * 4 + e] = D.ml[k].v[e];
    std::vector<int64_t> gi((size_t)D.n_gt * 4);   // int(b[k]) truncation (bbox_augmentation.py:44)
    for (int k = 0; k < D.n_gt * 4; ++k) gi[k] = (int64_t)g[k];
    D.ops.assign(OPS_PER_VIEW, oadg_op_t{});
    const int n_reg = (int)D.ml.size() + 1;
    int depth_sum = 0;
    for (int b = 0; b < cfg->mixture_width; ++b) {
      const int depth = cfg->mixture_depth > 0 ? cfg->mixture_depth : (int)R.randint(1, 4);
      if (depth > OADG_MAX_DEPTH) return OADG_E_LIMIT;
      D.depth[b] = depth;
      depth_sum += depth;
      for (int d = 0; d < depth; ++d)
        for (int r = 0; r < n_reg; ++r) {
          oadg_op_t& op = D.ops[(b * OADG_MAX_DEPTH + d) * OADG_MAX_REGIONS + r];
          op.lut = -1;
          op.scratch = -1;
          const Name name = aug[R.randint(0, n_aug)];
          switch (name) {
            case AUTOCONTRAST: op.kind = OADG_OP_AUTOCONTRAST; break;
            case EQUALIZE: op.kind = OADG_OP_EQUALIZE; break;
            case POSTERIZE:
              op.kind = OADG_OP_POSTERIZE;
              op.p0 = 4 - (int)((0.1 + (sev - 0.1) * R.u01()) * 4 / 10);
              break;
            case SOLARIZE:
              op.kind = OADG_OP_SOLARIZE;
              op.p0 = 256 - (int)((0.1 + (sev - 0.1) * R.u01()) * 256 / 10);
              break;
            case COLOR:
            case CONTRAST:
            case BRIGHTNESS:
            case SHARPNESS:
              op.kind = name == COLOR ? OADG_OP_COLOR
                                      : (name == CONTRAST ? OADG_OP_CONTRAST
                                                          : (name == BRIGHTNESS ? OADG_OP_BRIGHTNESS : OADG_OP_SHARPNESS));
              op.factor = (float)((0.1 + (sev - 0.1) * R.u01()) * 1.8 / 10. + 0.1);
              break;
            case INVERT:
              op.kind = OADG_OP_INVERT;
              op.p0 = R.u01() > 0.5 ? 1 : -1;
              op.p1 = R.u01() > 0.5 ? 1 : -1;
              break;
            default: {
              const bool bg = name >= BG_ROTATE;
              const Name base = bg ? (Name)(name - BG_ROTATE + BBO_ROTATE) : name;
              Geo geo = ROTATE;
              if (base == BBO_SHEAR_XY) geo = R.u01() < 0.5 ? SHEAR_X : SHEAR_Y;
              else if (base == BBO_TRANSLATE_XY) geo = R.u01() < 0.5 ? TRANSLATE_X : TRANSLATE_Y;
              if (bg) {
                double level;
                bool neg;
                level_sign(level, neg);
                op.kind = OADG_OP_BG_AFFINE;
                forward_affine(geo, level, neg, D.w, D.h, false, 0, 0, D.w, D.h, op.minv);
                invert_affine(op.minv);
              } else {
                op.kind = OADG_OP_BBO_AFFINE;
                op.bbo_first = (int32_t)D.bbo.size();
                for (int k = 0; k < D.n_gt; ++k) {
                  const int64_t x1 = gi[4 * k], y1 = gi[4 * k + 1], x2 = gi[4 * k + 2], y2 = gi[4 * k + 3];
                  if (x2 - x1 < 1 || y2 - y1 < 1) continue;  // bbox_augmentation.py:45-47: skipped before any draw
                  double level;
                  bool neg;
                  level_sign(level, neg);
                  oadg_bbo_t B{};
                  B.gt = k;
                  forward_affine(geo, level, neg, (double)(x2 - x1 + 1), (double)(y2 - y1 + 1), true, (x1 + x2) / 2.,
                                 (y1 + y2) / 2., D.w, D.h, B.minv);
                  invert_affine(B.minv);
                  D.bbo.push_back(B);
                }
                op.bbo_count = (int32_t)D.bbo.size() - op.bbo_first;
              }
            }
          }
        }
    }
    for (int b = cfg->mixture_width; b < OADG_MAX_WIDTH; ++b) D.depth[b] = 0;
    if (depth_sum_out) depth_sum_out[v] = depth_sum;
    // ---- tail: object-aware targets and mixing coefficients (oa_mix.py:245-262,282,295-298)
    const double* sc = scores[v];
    for (int k = 0; k < D.n_gt; ++k)
      if (sc[k] <= (double)cfg->score_thresh) D.low.push_back(k);
    int want = (int)D.low.size();
    want = want < 1 ? 1 : (want > 5 ? 5 : want);
    sample_regions(R, D.h, D.w, cfg->oa_random_box_scale, cfg->oa_random_box_ratio, want, want, g, D.n_gt, sc, true, D.oa,
                   D.oa_scores);
    n_oa_out[v] = (int32_t)D.oa.size();
    for (size_t k = 0; k < D.oa.size() && k < 5; ++k)
      for (int e = 0; e < 4; ++e) oa_boxes_out[((size_t)v * 5 + k) * 4 + e] = D.oa[k].v[e];
    {  // np.random.beta(1.0, 1.0): Johnk's algorithm
      for (;;) {
        const double U = R.u01(), V = R.u01();
        const double X = pow(U, 1.0 / 1.0), Y = pow(V, 1.0 / 1.0);
        const double XpY = X + Y;
        if (XpY <= 1.0 && U + V > 0.0) {
          if (XpY > 0) {
            D.m = X / XpY;
          } else {
            double logX = log(U) / 1.0, logY = log(V) / 1.0;
            const double logM = logX > logY ? logX : logY;
            logX -= logM;
            logY -= logM;
            D.m = exp(logX - log(exp(logX) + exp(logY)));
          }
          break;
        }
      }
    }
    for (size_t k = 0; k < D.low.size(); ++k) {
      const double s = sc[D.low[k]];
      D.m_oa.push_back(s <= (double)cfg->score_thresh ? (float)(0.0 + 0.5 * R.u01()) : (float)(0.0 + 1.0 * R.u01()));
    }
    for (size_t k = 0; k < D.oa.size(); ++k) {
      const double s = D.oa_scores[k];
      D.m_oa.push_back(s <= (double)cfg->score_thresh ? (float)(0.0 + 0.5 * R.u01()) : (float)(0.0 + 1.0 * R.u01()));
    }
  }

  // ---- pack (layout of oadg_b200/plan.py: header | views | gts | ops | bbo | targets, 8-byte aligned sections)
  size_t n_gt_tot = 0, n_bbo = 0, n_tgt = 0;
  for (const ViewDraw& D : views) {
    n_gt_tot += (size_t)D.n_gt;
    n_bbo += D.bbo.size();
    n_tgt += D.low.size() + D.oa.size();
  }
  size_t off = a8(sizeof(oadg_plan_header_t));
  const size_t off_views = off;
  off = a8(off + (size_t)n_img * sizeof(oadg_view_t));
  const size_t off_gt = off;
  off = a8(off + n_gt_tot * sizeof(oadg_gt_t));
  const size_t off_ops = off;
  off = a8(off + (size_t)n_img * OPS_PER_VIEW * sizeof(oadg_op_t));
  const size_t off_bbo = off;
  off = a8(off + n_bbo * sizeof(oadg_bbo_t));
  const size_t off_tgt = off;
  off = a8(off + n_tgt * sizeof(oadg_target_t));
  *plan_bytes = off;
  if (!plan_out || plan_cap < off) return OADG_E_ARG;
  char* blob = static_cast<char*>(plan_out);
  memset(blob, 0, off);
  auto* H = reinterpret_cast<oadg_plan_header_t*>(blob);
  auto* PV = reinterpret_cast<oadg_view_t*>(blob + off_views);
  auto* PG = reinterpret_cast<oadg_gt_t*>(blob + off_gt);
  auto* PO = reinterpret_cast<oadg_op_t*>(blob + off_ops);
  auto* PB = reinterpret_cast<oadg_bbo_t*>(blob + off_bbo);
  auto* PT = reinterpret_cast<oadg_target_t*>(blob + off_tgt);
  int g0 = 0, b0 = 0, t0 = 0, max_h = 1, max_w = 1;
  const int sr = cfg->spatial_ratio;
  for (int v = 0; v < n_img; ++v) {
    const ViewDraw& D = views[v];
    max_h = D.h > max_h ? D.h : max_h;
    max_w = D.w > max_w ? D.w : max_w;
    const int h4 = D.h / sr, w4 = D.w / sr;
    for (int k = 0; k < D.n_gt; ++k) {  // blurred-mask source records (oa_mix.py:78-91)
      oadg_gt_t& G = PG[g0 + k];
      const float* fb = gt[v] + 4 * k;
      int32_t lo[4];
      for (int e = 0; e < 4; ++e) lo[e] = (int32_t)floorf(fb[e] / (float)sr);   // np.array(gt // sr, int32)
      const double sx = (lo[2] - lo[0]) * cfg->sigma_ratio / 3 * 2;
      const double sy = (lo[3] - lo[1]) * cfg->sigma_ratio / 3 * 2;
      const bool blur = !(sx <= 0 || sy <= 0);
      int32_t xs, xe, ys, ye;
      py_slice(lo[0], lo[2], w4, xs, xe);
      py_slice(lo[1], lo[3], h4, ys, ye);
      G.lo[0] = xs; G.lo[1] = ys; G.lo[2] = xe; G.lo[3] = ye;
      G.blur = blur ? 1 : 0;
      G.kx = blur ? ((int)nearbyint(sx * 8 + 1) | 1) : 1;   // cvRound(sigma*8+1)|1 for non-8U depth
      G.ky = blur ? ((int)nearbyint(sy * 8 + 1) | 1) : 1;
      G.view = v;
      G.sigma_x = blur ? sx : 1.0;
      G.sigma_y = blur ? sy : 1.0;
      G.supp[0] = G.supp[1] = G.supp[2] = G.supp[3] = 0;
      if (xe > xs && ye > ys && w4 > 0 && h4 > 0) {
        const int s_[2] = {xs, ys}, e_[2] = {xe, ye}, nlo[2] = {w4, h4}, nhi[2] = {D.w, D.h}, ks[2] = {G.kx, G.ky};
        for (int ax = 0; ax < 2; ++ax) {
          const int rad = blur ? ks[ax] / 2 : 0;
          const int a = s_[ax] - rad > 0 ? s_[ax] - rad : 0;
          const int bnd = e_[ax] - 1 + rad < nlo[ax] - 1 ? e_[ax] - 1 + rad : nlo[ax] - 1;
          const double up = (double)nhi[ax] / nlo[ax];
          const int d0 = (int)floor((a - 0.5) * up - 0.5) - 1;
          const int d1 = (int)ceil((bnd + 1.5) * up - 0.5) + 1;
          G.supp[ax] = d0 > 0 ? d0 : 0;
          G.supp[ax + 2] = d1 < nhi[ax] ? d1 : nhi[ax];
        }
      }
    }
    const int base = v * OPS_PER_VIEW;
    for (int i = 0; i < OPS_PER_VIEW; ++i) {
      oadg_op_t op = D.ops[i];
      if (op.kind == OADG_OP_BBO_AFFINE) op.bbo_first += b0;
      PO[base + i] = op;
    }
    for (size_t i = 0; i < D.bbo.size(); ++i) {
      PB[b0 + i] = D.bbo[i];
      PB[b0 + i].gt = g0 + D.bbo[i].gt;
    }
    int nt = 0;
    for (size_t k = 0; k < D.low.size(); ++k, ++nt) {
      oadg_target_t& T = PT[t0 + nt];
      T.kind = 0;
      T.gt = g0 + D.low[k];
      T.m_oa = D.m_oa[k];
    }
    for (size_t k = 0; k < D.oa.size(); ++k, ++nt) {
      oadg_target_t& T = PT[t0 + nt];
      T.kind = 1;
      T.gt = -1;
      py_slice(D.oa[k].v[0], D.oa[k].v[2], D.w, T.box[0], T.box[2]);
      py_slice(D.oa[k].v[1], D.oa[k].v[3], D.h, T.box[1], T.box[3]);
      T.m_oa = D.m_oa[D.low.size() + k];
    }
    oadg_view_t& V = PV[v];
    V.H = D.h;
    V.W = D.w;
    V.img = v;
    V.n_gt = D.n_gt;
    V.gt_first = g0;
    V.n_ml = (int32_t)D.ml.size();
    for (size_t k = 0; k < D.ml.size(); ++k)
      for (int e = 0; e < 4; ++e) V.ml_box[k][e] = (int32_t)D.ml[k].v[e];
    V.width = cfg->mixture_width;
    for (int b = 0; b < OADG_MAX_WIDTH; ++b) {
      V.depth[b] = D.depth[b];
      V.ws[b] = D.ws[b];
    }
    V.op_first = base;
    V.n_tgt = nt;
    V.tgt_first = t0;
    V.m = D.m;
    g0 += D.n_gt;
    b0 += (int)D.bbo.size();
    t0 += nt;
  }
  H->magic = 0x4F414447;
  H->abi = OADG_ABI_VERSION;
  H->n_views = n_img;
  H->n_gt = (int32_t)n_gt_tot;
  H->n_ops = n_img * OPS_PER_VIEW;
  H->n_bbo = (int32_t)n_bbo;
  H->n_tgt = (int32_t)n_tgt;
  H->max_h = max_h;
  H->max_w = max_w;
  H->off_views = (int32_t)off_views;
  H->off_gt = (int32_t)off_gt;
  H->off_ops = (int32_t)off_ops;
  H->off_bbo = (int32_t)off_bbo;
  H->off_tgt = (int32_t)off_tgt;
  H->total_bytes = (int32_t)off;
  return 0;
}
