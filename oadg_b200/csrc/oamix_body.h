// Per-pixel bodies of the OA-Mix kernels (device) -- also compiled for the host by the
// arithmetic check in tests/hostsim (test infrastructure only, never a product path).
#pragma once
#include "oadg.h"
#include "oamix_math.h"

// Every frame, profile, mask, histogram and LUT the chain kernel reads may have been written earlier in the SAME
// launch by another SM (phases are separated by a grid barrier), so device loads are plain ld.global (coherent at
// the barrier's fence), never the non-coherent __ldg path.
#define OADG_LDG(p) (*(p))

namespace oadg {

struct DevPlan {
  const oadg_view_t* views;
  const oadg_gt_t* gts;
  const oadg_op_t* ops;
  const oadg_bbo_t* bbo;
  const oadg_target_t* tgts;
  const float* prof_x;  // [n_gt][max_w]
  const float* prof_y;  // [n_gt][max_h]
  int max_w, max_h;
  const uint8_t* luts;  // [slot][3][256]
  // union of the blurred gt masks of each view, np.max(mask_bboxes, axis=0) (bbox_augmentation.py:260):
  const float* maskf;    // [view][max_h*max_w] float32 value
  const uint8_t* masku;  // [view][max_h*max_w] uint8(mask*255), the image bg-only ops warp (bbox_augmentation.py:264)
  size_t mask_stride;    // max_h*max_w
  // fused Normalize + Pad + CHW epilogue of the mix (oadg_fused_out_t): normalised float32 value of every uint8 level,
  // [3 output planes][256]; plane k reads frame channel (norm_rgb ? 2 - k : k)
  const float* norm_lut;
  int32_t norm_rgb, pad0;
};

struct Lane {            // one (view, branch) chain alive at the current depth
  int32_t view, branch;
  int32_t op_base;       // index of region 0's op for this (branch, depth)
  int32_t hist_slot;     // -1 if no histogram needed
  const uint8_t* in;
  uint8_t* out;
  // copied from the plan by the host so that a CTA needs one record, not a chain of dependent loads:
  int32_t H, W, n_ml;
  int32_t box[2][4];                 // multi-level boxes
  int32_t kind[OADG_MAX_REGIONS];    // op kind of region r (r = n_ml: outside)
  int32_t lut[OADG_MAX_REGIONS];     // LUT slot of region r or -1
  int32_t scratch[OADG_MAX_REGIONS]; // bbo result frame slot of region r or -1
  int32_t all_streaming;             // every region runs a table-lookup / bbo-copy op: no pixel kernel needed
  int32_t in_map, mask_map;          // tensor-map slots (TMA staging) of `in` and of the view's uint8 union mask, -1: none
};

struct Chain {           // one bboxes-only op being evaluated (bbox_augmentation.py:74-88)
  int32_t view, lane;
  const uint8_t* in;     // the lane input the chain starts from (read-only source of level 1)
  uint8_t* S;            // ping-pong frames of the running image: level l reads X_l and writes Y_l with
  uint8_t* T;            //   Y_l = (l odd ? T : S),  X_l = (l == 1 ? in : (l odd ? S : T))
  int32_t map_in, map_S, map_T;   // tensor-map slots of the three frames (-1: none, the blend stages its rows by hand)
  int32_t pad;
};

struct BboJob {          // one needed box of one chain
  int32_t chain, bbo;    // chain index, index into the oadg_bbo array
  int32_t level;         // 1-based dependency level (oamix_exec.h schedule_chain)
  int32_t next_first, next_count;   // the chain's jobs of level + 1 (contiguous in the job table)
  int32_t pad;
  int32_t rect[4];       // the box's mask support [x0,y0,x1,y1): the only pixels the box can change
};
OADG_HD const uint8_t* chain_src(const Chain& C, int level) { return level == 1 ? C.in : ((level & 1) ? C.S : C.T); }
OADG_HD uint8_t* chain_dst(const Chain& C, int level) { return (level & 1) ? C.T : C.S; }
OADG_HD int chain_src_map(const Chain& C, int level) { return level == 1 ? C.map_in : ((level & 1) ? C.map_S : C.map_T); }

// ---- work items of the chain kernel ------------------------------------------------------------------------
// The host turns a plan into a queue of work ITEMS cut into TILES (one CTA iteration each), ordered so that every item
// comes after the items it depends on; an item starts when all tiles of its dependencies are done.
enum {
  OADG_IT_PROFILE = 0,   // obj = gt*2 + axis            1 tile
  OADG_IT_MASK = 1,      // obj = view                   tiles of 256 x 32 px
  OADG_IT_HIST = 2,      // obj = lane                   tiles of kHistTilePx px (linear)
  OADG_IT_LUT = 3,       // obj = lut job                1 tile
  OADG_IT_COPY = 4,      // obj = chain                  tiles of kCopyTileBytes (linear); aux = 1: S as well as T
  OADG_IT_BBO_R = 5,     // obj = bbo job                blend of one box, tiles of 256 x 16 px over its support
  OADG_IT_BBO_C = 6,     // obj = bbo job (level l-1)    catch-up copy X_l -> Y_l of its support minus level l's
  OADG_IT_STEP = 7,      // obj = lane                   tiles of 256 x 16 px
  OADG_IT_KINDS = 8
};
struct Item {
  int32_t kind, obj;
  int32_t tile0, ntiles;   // tile0: first tile index in the work queue
  int32_t tx;              // tiles per row (2-D kinds)
  int32_t aux;
  int32_t dep_first, dep_count;   // the items (indices into the item table) that must be complete before this one starts
  int32_t succ_first, succ_count; // the items that wait for this one (entries of the successor table)
  int32_t perm_first;             // >= 0: claim t runs tile perm[perm_first + t] (long tiles first), -1: tile t
  int32_t pad;
};
// (the OADG_* macros let scripts/tile_sweep.sh build timing variants; the defaults are the shipped configuration:
// measured over 24 bench batches, 64-wide bbo / step tiles cost +9 % / +14 %, 256-wide ones save 3 % / 5 %, taller
// (32-row) tiles cost +9 %: the queue wants tiles of ~20-30 us)
#ifndef OADG_HIST_TILE_PX
#define OADG_HIST_TILE_PX 32768
#endif
#ifndef OADG_COPY_TILE_BYTES
#define OADG_COPY_TILE_BYTES 65536
#endif
#ifndef OADG_BBO_TILE_W
#define OADG_BBO_TILE_W 256
#endif
#ifndef OADG_STEP_TILE_W_PX
#define OADG_STEP_TILE_W_PX 256
#endif
#ifndef OADG_CATCH_TILE_W
#define OADG_CATCH_TILE_W 512
#endif
#ifndef OADG_BBO_TILE_H
#define OADG_BBO_TILE_H 16
#endif
#ifndef OADG_STEP_TILE_H
#define OADG_STEP_TILE_H 16
#endif
#ifndef OADG_STEP_TILE_W
#define OADG_STEP_TILE_W 512
#endif
#ifndef OADG_MASK_TILE_H
#define OADG_MASK_TILE_H 32
#endif
constexpr int kMaskTileW = 256, kMaskTileH = OADG_MASK_TILE_H;
constexpr int kHistTilePx = OADG_HIST_TILE_PX;
constexpr int kCopyTileBytes = OADG_COPY_TILE_BYTES;
constexpr int kBboTileW = OADG_BBO_TILE_W, kBboTileH = OADG_BBO_TILE_H;   // bbo tiles start at the support's x0 rounded down to a multiple of 4
constexpr int kStepTileW = OADG_STEP_TILE_W, kStepTileH = OADG_STEP_TILE_H;   // lanes with per-pixel ops use kStepTileWPx wide tiles (Item.aux)
constexpr int kStepTileWPx = OADG_STEP_TILE_W_PX;
constexpr int kBboCatchW = OADG_CATCH_TILE_W;                     // catch-up copies move 512 x 16 px per tile

struct MixJob {
  int32_t view, pad;
  const uint8_t* src;
  const uint8_t* branch[OADG_MAX_WIDTH];
  uint8_t* out;
  float* f32_out;        // fused epilogue: [3, Hp, Wp] of the generated view, or null
  float* f32_src;        // ... and of the source frame, or null
  int32_t Wp, Hp;
};

struct LutJob {
  int32_t op, hist_slot, view, pad;
};

struct LdRO {   // read-only for the whole launch
  OADG_HD int operator()(const uint8_t* p) const { return (int)OADG_LDG(p); }
};
struct LdRW {   // frames rewritten by the same chain between launches
  OADG_HD int operator()(const uint8_t* p) const { return (int)*p; }
};
OADG_HD int ldb(const uint8_t* p) { return (int)OADG_LDG(p); }

OADG_HD bool is_lut_kind(int k) {
  return k == OADG_OP_AUTOCONTRAST || k == OADG_OP_EQUALIZE || k == OADG_OP_POSTERIZE ||
         k == OADG_OP_SOLARIZE || k == OADG_OP_CONTRAST || k == OADG_OP_BRIGHTNESS;
}
OADG_HD bool needs_hist(int k) {
  return k == OADG_OP_AUTOCONTRAST || k == OADG_OP_EQUALIZE || k == OADG_OP_CONTRAST;
}

// blurred mask of gt g at (x, y): outer product of the two 1-D profiles (oa_mix.py:75-93)
OADG_HD float fg_mask(const DevPlan& P, int g, int x, int y) {
  const oadg_gt_t& G = P.gts[g];
  if (x < G.supp[0] || x >= G.supp[2] || y < G.supp[1] || y >= G.supp[3]) return 0.f;
  return fmul(OADG_LDG(P.prof_y + (size_t)g * P.max_h + y), OADG_LDG(P.prof_x + (size_t)g * P.max_w + x));
}
// np.max(mask_bboxes, axis=0) at (x, y)   (bbox_augmentation.py:260)
OADG_HD float union_mask(const DevPlan& P, const oadg_view_t& V, int x, int y) {
  float m = 0.f;
  for (int k = 0; k < V.n_gt; ++k) {
    float v = fg_mask(P, V.gt_first + k, x, y);
    m = v > m ? v : m;
  }
  return m;
}

// LUT entry i of a non-histogram op / the CONTRAST op
OADG_HD uint8_t lut_simple_at(const oadg_op_t& op, int i, double luma_sum, double npx) {
  if (op.kind == OADG_OP_POSTERIZE) return lut_posterize_at(i, op.p0);
  if (op.kind == OADG_OP_SOLARIZE) return lut_solarize_at(i, op.p0);
  if (op.kind == OADG_OP_BRIGHTNESS) return (uint8_t)pil_blend(0, i, op.factor);
  // CONTRAST: degenerate = int(mean(luma) + 0.5)  (PIL ImageEnhance.Contrast / ImageStat.mean)
  double mean = luma_sum / npx;
  int deg = (int)(mean + 0.5);
  return (uint8_t)pil_blend(deg, i, op.factor);
}

// ---- one pixel of one box of a bboxes-only chain (bbox_augmentation.py:57-71) ----------
// Boxes are grouped into dependency LEVELS (oamix_exec.h): all boxes of level l read the image as it stands after
// level l-1 (X_l) and write their blended supports into the other frame (Y_l); the supports level l-1 wrote are
// copied across where level l does not overwrite them (catch-up), so Y_l is the complete image after level l.
OADG_HD void bbo_r_pixel(const DevPlan& P, const Chain& C, const oadg_bbo_t& B, const uint8_t* X, uint8_t* Y, int x, int y) {
  const oadg_view_t& V = P.views[C.view];
  const size_t o = ((size_t)y * V.W + x) * 3;
  const float m = fmul(OADG_LDG(P.prof_y + (size_t)B.gt * P.max_h + y), OADG_LDG(P.prof_x + (size_t)B.gt * P.max_w + x));
  int v[3] = {X[o], X[o + 1], X[o + 2]};
  if (m != 0.f) {  // m == 0 => img*1 + aug*0 == img exactly
    double minv[6];
    for (int i = 0; i < 6; ++i) minv[i] = B.minv[i];
    WarpTap t = warp_px(minv, warp_row(minv, y), x);
    int a[3];
    warp_fetch3(LdRW(), X, V.H, V.W, t, a);
    for (int c = 0; c < 3; ++c) v[c] = bbo_blend(m, v[c], a[c]);
  }
  Y[o] = (uint8_t)v[0];
  Y[o + 1] = (uint8_t)v[1];
  Y[o + 2] = (uint8_t)v[2];
}
// catch-up of pixel (x, y) of a level l-1 support: copied unless a level-l box (jobs next_first..) rewrites it
OADG_HD void bbo_c_pixel(const BboJob* jobs, const BboJob& J, int W, const uint8_t* X, uint8_t* Y, int x, int y) {
  for (int k = 0; k < J.next_count; ++k) {
    const int32_t* r = jobs[J.next_first + k].rect;
    if (x >= r[0] && x < r[2] && y >= r[1] && y < r[3]) return;
  }
  const size_t o = ((size_t)y * W + x) * 3;
  Y[o] = X[o];
  Y[o + 1] = X[o + 1];
  Y[o + 2] = X[o + 2];
}

// ---- union mask of one view at one pixel: written once per batch by mask_kernel -------------------
OADG_HD void mask_pixel(const DevPlan& P, int view, int x, int y, float* maskf, uint8_t* masku) {
  const oadg_view_t& V = P.views[view];
  const float m = union_mask(P, V, x, y);
  const size_t o = (size_t)view * P.mask_stride + (size_t)y * V.W + x;
  maskf[o] = m;
  masku[o] = (uint8_t)mask_to_u8(m);
}

// ---- bg-only op at one pixel (bbox_augmentation.py:240-272) ----------------------------------------
OADG_HD void bg_pixel(const DevPlan& P, const oadg_view_t& V, const oadg_op_t& op, const uint8_t* in, int view,
                      int x, int y, const int img[3], int out[3]) {
  double minv[6];
  for (int i = 0; i < 6; ++i) minv[i] = op.minv[i];
  const WarpTap t = warp_px(minv, warp_row(minv, y), x);
  warp_fetch3(LdRO(), in, V.H, V.W, t, out);
  const size_t mo = (size_t)view * P.mask_stride;
  const float M = OADG_LDG(P.maskf + mo + (size_t)y * V.W + x);
  // cv2.warpAffine of uint8(mask*255) with the same matrix (bbox_augmentation.py:263-264)
  const uint8_t* mu = P.masku + mo;
  const bool x0 = (unsigned)t.sx < (unsigned)V.W, x1 = t.fx != 0 && (unsigned)(t.sx + 1) < (unsigned)V.W;
  const bool y0 = (unsigned)t.sy < (unsigned)V.H, y1 = t.fy != 0 && (unsigned)(t.sy + 1) < (unsigned)V.H;
  const uint8_t* r0 = mu + (size_t)t.sy * V.W + t.sx;
  const uint8_t* r1 = r0 + V.W;
  const int wm = bilerp_fix((y0 && x0) ? ldb(r0) : 0, (y0 && x1) ? ldb(r0 + 1) : 0, (y1 && x0) ? ldb(r1) : 0,
                            (y1 && x1) ? ldb(r1 + 1) : 0, t.fx, t.fy);
  if (M != 0.f || wm != 0)  // keep == 0 => 0*img + 1*aug == aug exactly
    for (int c = 0; c < 3; ++c) out[c] = bg_blend(M, wm, img[c], out[c]);
}

// ---- one op at one pixel (oa_mix.py:264-279 dispatch) ------------------------------------
OADG_HD void eval_op(const DevPlan& P, const oadg_view_t& V, const oadg_op_t& op, const uint8_t* in,
                     const uint8_t* scratch, size_t frame_bytes, int x, int y, int out[3]) {
  const int H = V.H, W = V.W;
  const size_t o = ((size_t)y * W + x) * 3;
  const int kind = op.kind;
  if (kind == OADG_OP_BBO_AFFINE) {
    const uint8_t* s = op.scratch >= 0 ? scratch + (size_t)op.scratch * frame_bytes : in;
    out[0] = ldb(s + o);
    out[1] = ldb(s + o + 1);
    out[2] = ldb(s + o + 2);
    return;
  }
  int v[3] = {ldb(in + o), ldb(in + o + 1), ldb(in + o + 2)};
  if (is_lut_kind(kind)) {
    const uint8_t* lut = P.luts + (size_t)op.lut * 768;
    out[0] = ldb(lut + v[0]);
    out[1] = ldb(lut + 256 + v[1]);
    out[2] = ldb(lut + 512 + v[2]);
  } else if (kind == OADG_OP_BG_AFFINE) {
    bg_pixel(P, V, op, in, (int)(&V - P.views), x, y, v, out);
  } else if (kind == OADG_OP_INVERT) {
    // -cv2.warpAffine(img, [[1,0,tx],[0,1,ty]]) : integer shift, zero fill, uint8 negation
    int xs = x - op.p0, ys = y - op.p1;
    if ((unsigned)xs < (unsigned)W && (unsigned)ys < (unsigned)H) {
      const uint8_t* p = in + ((size_t)ys * W + xs) * 3;
      out[0] = (-ldb(p)) & 255;
      out[1] = (-ldb(p + 1)) & 255;
      out[2] = (-ldb(p + 2)) & 255;
    } else {
      out[0] = out[1] = out[2] = 0;
    }
  } else if (kind == OADG_OP_COLOR) {
    int deg = pil_luma(v[0], v[1], v[2]);
    for (int c = 0; c < 3; ++c) out[c] = pil_blend(deg, v[c], op.factor);
  } else if (kind == OADG_OP_SHARPNESS) {
    if (x == 0 || y == 0 || x == W - 1 || y == H - 1) {  // SMOOTH copies the border
      for (int c = 0; c < 3; ++c) out[c] = pil_blend(v[c], v[c], op.factor);
    } else {
      for (int c = 0; c < 3; ++c) {
        int nb[9];
        for (int dy = 0; dy < 3; ++dy)
          for (int dx = 0; dx < 3; ++dx)
            nb[dy * 3 + dx] = ldb(in + ((size_t)(y + 1 - dy) * W + (x - 1 + dx)) * 3 + c);
        out[c] = pil_blend(pil_smooth9(nb), v[c], op.factor);
      }
    }
  } else {
    out[0] = v[0];
    out[1] = v[1];
    out[2] = v[2];
  }
}

// ---- one pixel of one depth step (oa_mix.py:226-234): the region picks the op ---------
OADG_HD void step_pixel(const DevPlan& P, const Lane& L, const uint8_t* scratch, size_t frame_bytes, int x, int y) {
  const oadg_view_t& V = P.views[L.view];
  int r = V.n_ml;
  for (int b = 0; b < V.n_ml; ++b)
    if (x >= V.ml_box[b][0] && x < V.ml_box[b][2] && y >= V.ml_box[b][1] && y < V.ml_box[b][3]) r = b;
  int out[3];
  eval_op(P, V, P.ops[L.op_base + r], L.in, scratch, frame_bytes, x, y, out);
  uint8_t* q = L.out + ((size_t)y * V.W + x) * 3;
  q[0] = (uint8_t)out[0];
  q[1] = (uint8_t)out[1];
  q[2] = (uint8_t)out[2];
}

// ---- one pixel of branch mixing + object-aware mixing (oa_mix.py:236,281-309) ---------
OADG_HD void mix_pixel(const DevPlan& P, const MixJob& J, int x, int y) {
  const oadg_view_t& V = P.views[J.view];
  const size_t o = ((size_t)y * V.W + x) * 3;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < V.width; ++i) {
    const uint8_t* b = J.branch[i] + o;
    for (int c = 0; c < 3; ++c) acc[c] = fadd(acc[c], fmul(V.ws[i], (float)ldb(b + c)));
  }
  const int img[3] = {ldb(J.src + o), ldb(J.src + o + 1), ldb(J.src + o + 2)};
  float orig[3] = {0.f, 0.f, 0.f}, aug[3] = {0.f, 0.f, 0.f};
  MixMask ms = {0.f, 0.f};
  for (int t = 0; t < V.n_tgt; ++t) {
    const oadg_target_t& T = P.tgts[V.tgt_first + t];
    float mask;
    if (T.kind == 0) mask = fg_mask(P, T.gt, x, y);
    else mask = (x >= T.box[0] && x < T.box[2] && y >= T.box[1] && y < T.box[3]) ? 1.f : 0.f;
    if (mask == 0.f) continue;  // exact: weight 0 adds +0 and leaves sum == max
    float w = mix_target_weight(ms, mask);
    for (int c = 0; c < 3; ++c) mix_accumulate(orig[c], aug[c], T.m_oa, img[c], acc[c], w);
  }
  uint8_t* q = J.out + o;
  int res[3];
  for (int c = 0; c < 3; ++c) q[c] = (uint8_t)(res[c] = mix_finish(orig[c], aug[c], V.m, img[c], acc[c], ms.sum));
  if (J.f32_out || J.f32_src) {   // fused Normalize + Pad + CHW epilogue
    const size_t plane = (size_t)J.Hp * J.Wp, at = (size_t)y * J.Wp + x;
    for (int k = 0; k < 3; ++k) {
      const int c = P.norm_rgb ? 2 - k : k;
      if (J.f32_out) J.f32_out[k * plane + at] = P.norm_lut[k * 256 + res[c]];
      if (J.f32_src) J.f32_src[k * plane + at] = P.norm_lut[k * 256 + img[c]];
    }
  }
}

}  // namespace oadg
