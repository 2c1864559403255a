// Host orchestration of an OA-Mix plan: validation, workspace layout, launch tables and the
// launch sequence, written against a small Backend so that the CUDA launcher (oamix.cu) and
// the host arithmetic check (tests/hostsim, test infrastructure only) share it verbatim.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "oamix_body.h"

#ifdef OADG_SCHED_TIMING
#include <chrono>
#include <stdio.h>
#define OADG_T(name) do { auto _n = std::chrono::steady_clock::now(); fprintf(stderr, "[sched] %-12s %7.1f us\n", name, std::chrono::duration<double, std::micro>(_n - _t0).count()); _t0 = _n; } while (0)
#else
#define OADG_T(name) do { } while (0)
#endif

namespace oadg {

constexpr int kMagic = 0x4F414447;
constexpr int kTensorMapBytes = 128;   // sizeof(CUtensorMap), 64-byte aligned in device memory
// The tile tickets (see oamix.cu).  Every tile of the launch gets a TICKET, in the order in which its item became
// ready; tickets[T] = (item + 1) << 32 | tile once ticket T is published, 0 before (the table is zeroed with the other
// counters).  A CTA claims a tile with ONE atomic on the claim cursor and reads its ticket.  The uploaded `ring` holds
// the header {[0] tickets published so far (the publish cursor), [1] the claim cursor, [2] number of initial items}
// and then, for every item that starts ready (in queue = priority order), first_ticket << 32 | item: the kernel's
// CTAs write those tickets themselves when they start.
constexpr int kRingHeader = 4;
// geometry of a staged gather source (one sub-tile of 64 x 16 output pixels, oamix.cu): TMA boxes of 16 rows
constexpr int kGatherBoxRows = 16, kGatherImgBoxBytes = 256, kGatherMaskBoxBytes = 96;

inline size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct PlanView {
  const oadg_plan_header_t* h;
  const oadg_view_t* views;
  const oadg_gt_t* gts;
  const oadg_op_t* ops;
  const oadg_bbo_t* bbo;
  const oadg_target_t* tgts;
};

inline int op_index(const oadg_view_t& V, int b, int d, int r) {
  return V.op_first + (b * OADG_MAX_DEPTH + d) * OADG_MAX_REGIONS + r;
}

inline int parse_plan(const void* blob, size_t bytes, PlanView& pv) {
  if (!blob || bytes < sizeof(oadg_plan_header_t)) return OADG_E_ARG;
  const auto* h = static_cast<const oadg_plan_header_t*>(blob);
  if (h->magic != kMagic || h->abi != OADG_ABI_VERSION) return OADG_E_PLAN;
  if ((size_t)h->total_bytes != bytes) return OADG_E_PLAN;
  auto fits = [&](int off, int n, size_t sz) {
    return off >= (int)sizeof(oadg_plan_header_t) && n >= 0 && (size_t)off + (size_t)n * sz <= bytes && off % 8 == 0;
  };
  if (!fits(h->off_views, h->n_views, sizeof(oadg_view_t)) || !fits(h->off_gt, h->n_gt, sizeof(oadg_gt_t)) ||
      !fits(h->off_ops, h->n_ops, sizeof(oadg_op_t)) || !fits(h->off_bbo, h->n_bbo, sizeof(oadg_bbo_t)) ||
      !fits(h->off_tgt, h->n_tgt, sizeof(oadg_target_t)))
    return OADG_E_PLAN;
  const char* b = static_cast<const char*>(blob);
  pv.h = h;
  pv.views = reinterpret_cast<const oadg_view_t*>(b + h->off_views);
  pv.gts = reinterpret_cast<const oadg_gt_t*>(b + h->off_gt);
  pv.ops = reinterpret_cast<const oadg_op_t*>(b + h->off_ops);
  pv.bbo = reinterpret_cast<const oadg_bbo_t*>(b + h->off_bbo);
  pv.tgts = reinterpret_cast<const oadg_target_t*>(b + h->off_tgt);
  if (h->max_h <= 0 || h->max_w <= 0) return OADG_E_PLAN;
  for (int v = 0; v < h->n_views; ++v) {
    const oadg_view_t& V = pv.views[v];
    if (V.H <= 0 || V.W <= 0 || V.H > h->max_h || V.W > h->max_w) return OADG_E_PLAN;
    if (V.width < 1 || V.width > OADG_MAX_WIDTH || V.n_ml < 0 || V.n_ml > 2) return OADG_E_LIMIT;
    if (V.n_gt < 0 || V.gt_first < 0 || V.gt_first + V.n_gt > h->n_gt) return OADG_E_PLAN;
    if (V.n_tgt < 0 || V.tgt_first < 0 || V.tgt_first + V.n_tgt > h->n_tgt) return OADG_E_PLAN;
    if (V.op_first < 0 || V.op_first + OADG_MAX_WIDTH * OADG_MAX_DEPTH * OADG_MAX_REGIONS > h->n_ops)
      return OADG_E_PLAN;
    for (int b2 = 0; b2 < V.n_ml; ++b2)
      if (V.ml_box[b2][0] < 0 || V.ml_box[b2][1] < 0 || V.ml_box[b2][2] > V.W || V.ml_box[b2][3] > V.H)
        return OADG_E_PLAN;
    for (int br = 0; br < V.width; ++br)
      if (V.depth[br] < 1 || V.depth[br] > OADG_MAX_DEPTH) return OADG_E_LIMIT;
    for (int br = 0; br < V.width; ++br)
      for (int d = 0; d < V.depth[br]; ++d)
        for (int r = 0; r <= V.n_ml; ++r) {
          const oadg_op_t& op = pv.ops[op_index(V, br, d, r)];
          if (op.kind < 0 || op.kind > OADG_OP_BBO_AFFINE) return OADG_E_PLAN;
          if (op.kind == OADG_OP_BBO_AFFINE &&
              (op.bbo_first < 0 || op.bbo_count < 0 || op.bbo_first + op.bbo_count > h->n_bbo))
            return OADG_E_PLAN;
          if (op.kind == OADG_OP_POSTERIZE && (op.p0 < 1 || op.p0 > 8)) return OADG_E_PLAN;
        }
    for (int t = 0; t < V.n_tgt; ++t) {
      const oadg_target_t& T = pv.tgts[V.tgt_first + t];
      if (T.kind == 0 && (T.gt < V.gt_first || T.gt >= V.gt_first + V.n_gt)) return OADG_E_PLAN;
    }
  }
  for (int g = 0; g < h->n_gt; ++g) {
    const oadg_gt_t& G = pv.gts[g];
    if (G.view < 0 || G.view >= h->n_views) return OADG_E_PLAN;
    const oadg_view_t& V = pv.views[G.view];
    if (G.supp[0] < 0 || G.supp[1] < 0 || G.supp[2] > V.W || G.supp[3] > V.H) return OADG_E_PLAN;
    if (G.blur && (G.kx < 1 || G.ky < 1 || !(G.sigma_x > 0) || !(G.sigma_y > 0))) return OADG_E_PLAN;
    if (G.lo[0] < 0 || G.lo[1] < 0 || G.lo[2] > V.W / 4 || G.lo[3] > V.H / 4) return OADG_E_PLAN;
  }
  for (int i = 0; i < h->n_bbo; ++i)
    if (pv.bbo[i].gt < 0 || pv.bbo[i].gt >= h->n_gt) return OADG_E_PLAN;
  return 0;
}


struct Layout {
  size_t frame_bytes;
  size_t off_plan, off_prof_x, off_prof_y, off_branch, off_scratch, off_zero, off_hist, off_luma, off_lut, off_tables;
  size_t off_maskf, off_masku;
  size_t zero_bytes;
  int any_bg;
  size_t tables_bytes;
  int n_lanes_total, n_chains, n_lut, n_hist, max_depth, max_items, max_deps, max_perm, n_bbo, max_maps;
  size_t max_tiles, off_tickets;
  size_t total;
};

// Everything is sized from the plan alone (upper bounds: dead boxes / chains are only pruned at execute time).
inline void make_layout(const PlanView& pv, Layout& L) {
  const oadg_plan_header_t& h = *pv.h;
  L.frame_bytes = align_up_sz((size_t)h.max_h * h.max_w * 3, 256);
  int max_depth = 0, lanes_total = 0, n_lut = 0, n_hist = 0, n_chains = 0, max_chain = 0;
  int any_bg = 0;
  for (int v = 0; v < h.n_views; ++v) {
    const oadg_view_t& V = pv.views[v];
    for (int b = 0; b < V.width; ++b) {
      max_depth = V.depth[b] > max_depth ? V.depth[b] : max_depth;
      lanes_total += V.depth[b];
      for (int d = 0; d < V.depth[b]; ++d) {
        bool hist = false;
        for (int r = 0; r <= V.n_ml; ++r) {
          const oadg_op_t& op = pv.ops[op_index(V, b, d, r)];
          if (is_lut_kind(op.kind)) ++n_lut;
          if (needs_hist(op.kind)) hist = true;
          if (op.kind == OADG_OP_BBO_AFFINE && op.bbo_count > 0) {
            ++n_chains;
            max_chain = op.bbo_count > max_chain ? op.bbo_count : max_chain;
          }
          if (op.kind == OADG_OP_BG_AFFINE) any_bg = 1;
        }
        if (hist) ++n_hist;
      }
    }
  }
  L.max_depth = max_depth;
  L.n_lanes_total = lanes_total;
  L.n_chains = n_chains;
  L.n_lut = n_lut;
  L.n_hist = n_hist;
  L.n_bbo = h.n_bbo;
  L.max_items = 2 * h.n_gt + h.n_views + 2 * lanes_total + n_lut + n_chains + 2 * h.n_bbo + 8;
  L.max_deps = 8 * L.max_items + n_chains * 4 * max_chain * max_chain + 64;
  L.max_perm = lanes_total * (((h.max_w + kStepTileWPx - 1) / kStepTileWPx) * ((h.max_h + kStepTileH - 1) / kStepTileH)) + 16;
  {  // tensor maps: source + union mask per view, two ping-pong frames per branch, two frames per chain
    int wsum = 0;
    for (int v = 0; v < h.n_views; ++v) wsum += pv.views[v].width;
    L.max_maps = 2 * h.n_views + 2 * wsum + 2 * n_chains + 1;
  }
  {  // upper bound of the launch's tiles (dead boxes / chains are only pruned at execute time)
    auto cdiv = [](long long a, long long b) { return (size_t)((a + b - 1) / b); };
    size_t tiles = 64;
    for (int v = 0; v < h.n_views; ++v) {
      const oadg_view_t& V = pv.views[v];
      tiles += 2 * (size_t)V.n_gt + cdiv(V.W, kMaskTileW) * cdiv(V.H, kMaskTileH);
      for (int b = 0; b < V.width; ++b)
        for (int d = 0; d < V.depth[b]; ++d) {
          tiles += cdiv(V.W, kStepTileWPx) * cdiv(V.H, kStepTileH) + cdiv((long long)V.W * V.H, kHistTilePx);
          for (int r = 0; r <= V.n_ml; ++r) {
            const oadg_op_t& op = pv.ops[op_index(V, b, d, r)];
            tiles += 1;
            if (op.kind != OADG_OP_BBO_AFFINE || op.bbo_count <= 0) continue;
            tiles += cdiv((long long)V.W * V.H * 3, kCopyTileBytes);
            for (int j = 0; j < op.bbo_count; ++j) {
              const int32_t* sp = pv.gts[pv.bbo[op.bbo_first + j].gt].supp;
              const int w = sp[2] - (sp[0] & ~3), hg = sp[3] - sp[1];
              if (w <= 0 || hg <= 0) continue;
              tiles += cdiv(w, kBboTileW) * cdiv(hg, kBboTileH) + cdiv(w, kBboCatchW) * cdiv(hg, kBboTileH);
            }
          }
        }
    }
    L.max_tiles = tiles;
  }
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o = align_up_sz(o + bytes, 256);
    return at;
  };
  L.tables_bytes = align_up_sz((size_t)lanes_total * sizeof(Lane), 16) +
                   align_up_sz((size_t)(n_lut > 0 ? n_lut : 1) * sizeof(LutJob), 16) +
                   align_up_sz((size_t)(n_chains > 0 ? n_chains : 1) * sizeof(Chain), 16) +
                   align_up_sz((size_t)(h.n_bbo > 0 ? h.n_bbo : 1) * sizeof(BboJob), 16) +
                   align_up_sz((size_t)L.max_items * sizeof(Item), 16) +
                   2 * align_up_sz((size_t)L.max_deps * sizeof(int32_t), 16) +
                   align_up_sz((size_t)L.max_items * sizeof(int32_t), 16) +
                   align_up_sz((size_t)L.max_perm * sizeof(int32_t), 16) +
                   align_up_sz((size_t)h.n_views * sizeof(MixJob), 16) + 3 * 256 * sizeof(float) +
                   align_up_sz((size_t)L.max_maps * kTensorMapBytes, 128) + 128 +
                   align_up_sz((size_t)(L.max_items + kRingHeader) * sizeof(unsigned long long), 16) + 256;
  // plan blob and launch tables are contiguous so that one H2D copy uploads both
  L.off_tables = align_up_sz((size_t)h.total_bytes, 16);
  L.off_plan = take(L.off_tables + L.tables_bytes);
  L.off_prof_x = take((size_t)h.n_gt * h.max_w * sizeof(float));
  L.off_prof_y = take((size_t)h.n_gt * h.max_h * sizeof(float));
  size_t n_branch_frames = 0;
  for (int v = 0; v < h.n_views; ++v) n_branch_frames += (size_t)pv.views[v].width * 2;
  L.off_branch = take(n_branch_frames * L.frame_bytes);
  L.off_scratch = take((size_t)n_chains * 2 * L.frame_bytes);  // S + T per chain
  // one memset: [grid barrier counter | histograms | luma sums]
  // one memset: [0] work-queue counter; bytes 64..447: busy ns [16], tile counts [16], longest tile ns [16] per item
  // kind (uint64); bytes 512..: tiles done per item | histograms | luma sums
  L.off_zero = take(512 + (size_t)L.max_items * (sizeof(unsigned) + 2 * sizeof(unsigned long long)));   // bytes 512..: tiles
  // claimed per item, then (measurement aid) per item 2^63 - first claim time and last publish time (globaltimer ns)
  L.off_hist = take((size_t)(n_hist > 0 ? n_hist : 1) * 768 * sizeof(unsigned));
  L.off_luma = take((size_t)(n_hist > 0 ? n_hist : 1) * sizeof(unsigned long long));
  L.off_tickets = take(L.max_tiles * sizeof(unsigned long long));
  L.zero_bytes = o - L.off_zero;
  L.off_lut = take((size_t)(n_lut > 0 ? n_lut : 1) * 768);
  L.any_bg = any_bg;
  const size_t mask_px = (size_t)h.max_h * h.max_w;
  L.off_maskf = take(any_bg ? (size_t)h.n_views * mask_px * sizeof(float) : 0);
  L.off_masku = take(any_bg ? (size_t)h.n_views * mask_px : 0);
  L.total = o;
}

// ---- bboxes-only chains: dead-box elimination and level scheduling -----------------------------------------
// Box j of a chain (bbox_augmentation.py:74-88) changes only the pixels of its mask support supp_j and reads the
// running image at supp_j and at the inverse-affine footprint fp_j of supp_j.
//   * box j is NEEDED when supp_j meets the region that keeps the chain's result, or a later needed box reads it;
//     every other box cannot influence a kept pixel and is skipped (its RNG draws were already consumed).
//   * level(j) = max over needed i < j of { level(i)+1 if (fp_j u supp_j) meets supp_i ;  level(i) if fp_i meets
//     supp_j }.  Boxes of one level have disjoint supports and read nothing a same-level box writes, so a level
//     runs as one read phase + one write phase.
struct IRect {
  int x0, y0, x1, y1;
};
inline bool irect_hit(const IRect& a, const IRect& b) {
  return a.x0 < b.x1 && a.x1 > b.x0 && a.y0 < b.y1 && a.y1 > b.y0;
}
inline bool irect_empty(const IRect& a) { return a.x1 <= a.x0 || a.y1 <= a.y0; }
inline IRect irect_union(const IRect& a, const IRect& b) {
  if (irect_empty(a)) return b;
  if (irect_empty(b)) return a;
  return IRect{a.x0 < b.x0 ? a.x0 : b.x0, a.y0 < b.y0 ? a.y0 : b.y0, a.x1 > b.x1 ? a.x1 : b.x1, a.y1 > b.y1 ? a.y1 : b.y1};
}
// conservative source footprint of `s` under the dst->src map (fixed-point rounding < 1 px, +1 tap, +1 margin)
inline IRect footprint(const double* m, const IRect& s, int W, int H) {
  if (irect_empty(s)) return IRect{0, 0, 0, 0};
  double xs[4], ys[4];
  const double cx[2] = {(double)s.x0, (double)(s.x1 - 1)}, cy[2] = {(double)s.y0, (double)(s.y1 - 1)};
  for (int i = 0; i < 4; ++i) {
    xs[i] = m[0] * cx[i & 1] + m[1] * cy[i >> 1] + m[2];
    ys[i] = m[3] * cx[i & 1] + m[4] * cy[i >> 1] + m[5];
  }
  double xa = xs[0], xb = xs[0], ya = ys[0], yb = ys[0];
  for (int i = 1; i < 4; ++i) {
    xa = xs[i] < xa ? xs[i] : xa; xb = xs[i] > xb ? xs[i] : xb;
    ya = ys[i] < ya ? ys[i] : ya; yb = ys[i] > yb ? ys[i] : yb;
  }
  auto clampd = [](double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); };
  IRect f;
  f.x0 = (int)clampd(xa - 3.0, 0.0, (double)W);
  f.y0 = (int)clampd(ya - 3.0, 0.0, (double)H);
  f.x1 = (int)clampd(xb + 4.0, 0.0, (double)W);
  f.y1 = (int)clampd(yb + 4.0, 0.0, (double)H);
  return f;
}

struct ChainSched {
  std::vector<int> box;     // needed boxes (offsets into the op's bbo slice), plan order
  std::vector<int> level;   // 1-based level of each needed box
  int n_levels = 0;
};
inline void schedule_chain(const PlanView& pv, const oadg_view_t& V, const oadg_op_t& op, const IRect* keep,
                           ChainSched& out) {
  const int n = op.bbo_count;
  std::vector<IRect> supp(n), reads(n), fp(n);
  for (int j = 0; j < n; ++j) {
    const oadg_bbo_t& B = pv.bbo[op.bbo_first + j];
    const int32_t* s = pv.gts[B.gt].supp;
    supp[j] = IRect{s[0], s[1], s[2], s[3]};
    fp[j] = footprint(B.minv, supp[j], V.W, V.H);
    reads[j] = irect_union(fp[j], supp[j]);
  }
  std::vector<char> need(n, 0);
  for (int j = n - 1; j >= 0; --j) {
    if (irect_empty(supp[j])) continue;
    bool nd = keep == nullptr || irect_hit(supp[j], *keep);
    for (int k = j + 1; k < n && !nd; ++k) nd = need[k] && irect_hit(reads[k], supp[j]);
    need[j] = nd;
  }
  std::vector<int> lvl(n, 0);
  out.box.clear();
  out.level.clear();
  out.n_levels = 0;
  for (int j = 0; j < n; ++j) {
    if (!need[j]) continue;
    int l = 1;
    for (int i = 0; i < j; ++i) {
      if (!need[i]) continue;
      if (irect_hit(reads[j], supp[i])) l = lvl[i] + 1 > l ? lvl[i] + 1 : l;
      else if (irect_hit(fp[i], supp[j])) l = lvl[i] > l ? lvl[i] : l;
    }
    lvl[j] = l;
    out.box.push_back(j);
    out.level.push_back(l);
    out.n_levels = l > out.n_levels ? l : out.n_levels;
  }
}

// Tiles are claimed dynamically from one queue, in item order: within one dependency depth the items that hold long
// tiles come first so that the short ones fill the gaps.
inline int item_priority(int kind, bool lane_all_streaming) {
  switch (kind) {
    case OADG_IT_PROFILE: return 0;
    case OADG_IT_LUT: return 1;
    case OADG_IT_STEP: return lane_all_streaming ? 6 : 2;
    case OADG_IT_HIST: return 3;
    case OADG_IT_BBO_R: return 4;
    case OADG_IT_COPY: return 5;
    case OADG_IT_MASK: return 7;
    default: return 8;
  }
}

struct ChainArgs {       // everything the chain kernel needs (device pointers)
  DevPlan P;
  const Lane* lanes;
  const LutJob* lutjobs;
  const Chain* chains;
  const BboJob* bjobs;
  const Item* items;
  const int32_t* deps;       // dependency lists of the items (host-side checks only)
  const int32_t* succ;       // successor lists of the items
  const int32_t* perm;       // tile orders of the depth-step items with per-pixel work (long tiles first)
  int32_t* pending;          // [n_items] dependency tiles still outstanding (uploaded with the tables)
  int32_t n_items, n_tiles, grid, debug;
  unsigned* epoch;           // bumped whenever an item becomes ready (zeroed before the launch)
  unsigned* claimed;         // [n_items] tiles claimed of every item (zeroed before the launch)
  unsigned long long* item_ts;   // [n_items][2] measurement aid: 2^63 - first claim time, last publish time
  unsigned* hist;
  unsigned long long* luma;
  uint8_t* luts;
  float* prof_x;
  float* prof_y;
  float* maskf;
  uint8_t* masku;
  const uint8_t* scratch;
  size_t frame_bytes;
  unsigned long long* kind_ns;   // [16] CTA-busy ns per item kind, [16] tiles per kind, [16] longest tile (measurement aid)
  const void* maps;              // tensor maps (kTensorMapBytes each) of the frames the affine gathers stage with TMA
  unsigned long long* ring;      // ticket header + the initially ready items (uploaded with the tables)
  unsigned long long* tickets;   // [n_tiles] tile tickets (zeroed before the launch)
  unsigned* fault;               // sticky: != 0 when a CTA gave up waiting for work (inconsistent dependency tables)
};

// Backend concept (all return 0 or an error code):
//   grid()                         CTAs the chain kernel will run (one per SM)
//   upload(dst, src_host, bytes)   zero(dst, bytes)
//   chain(args, host_tables...)    the work-queue interpreter (one persistent launch on the device)
//   mix(P, jobs, n)                zero2d(dst, pitch_bytes, width_bytes, rows)
//   make_map(dst, base, inner_bytes, rows, box_inner)   encode a tensor map into dst (host staging), != 0: unavailable
template <class Backend>
int execute_plan(Backend& be, const void* plan_host, size_t plan_bytes, const uint8_t* const* src, int n_img,
                 uint8_t* const* dst, void* workspace, size_t workspace_bytes, const oadg_fused_out_t* fused = nullptr) {
#ifdef OADG_SCHED_TIMING
  auto _t0 = std::chrono::steady_clock::now();
#endif
  PlanView pv;
  int rc = parse_plan(plan_host, plan_bytes, pv);
  if (rc) return rc;
  const oadg_plan_header_t& h = *pv.h;
  if (h.n_views == 0) return 0;
  if (!src || !dst || !workspace) return OADG_E_ARG;
  for (int v = 0; v < h.n_views; ++v)
    if (pv.views[v].img < 0 || pv.views[v].img >= n_img || !src[pv.views[v].img] || !dst[v]) return OADG_E_ARG;
  Layout L;
  make_layout(pv, L);
  if (workspace_bytes < L.total) return OADG_E_ARG;
  if (((uintptr_t)workspace & 255) != 0) return OADG_E_ARG;
  {  // the profile tile keeps the kernel's prefix sums (<= 1536 taps) and the low-res profile (<= 1024) in shared memory
    int max_k = 1;
    for (int g = 0; g < h.n_gt; ++g) {
      max_k = pv.gts[g].kx > max_k ? pv.gts[g].kx : max_k;
      max_k = pv.gts[g].ky > max_k ? pv.gts[g].ky : max_k;
    }
    if ((h.max_w > h.max_h ? h.max_w : h.max_h) / 4 > 1024 || max_k > 1536) return OADG_E_LIMIT;
  }
  char* ws = static_cast<char*>(workspace);
  const int G = be.grid();
  if (G < 1) return OADG_E_LIMIT;

  OADG_T("parse+layout");
  // host staging buffer: plan (with lut / scratch slots filled in) followed by the launch tables
  std::vector<char> stage(L.off_tables + L.tables_bytes, 0);
  memcpy(stage.data(), plan_host, plan_bytes);
  auto* ops = reinterpret_cast<oadg_op_t*>(stage.data() + h.off_ops);

  std::vector<size_t> branch_base(h.n_views);
  {
    size_t f = 0;
    for (int v = 0; v < h.n_views; ++v) {
      branch_base[v] = f;
      f += (size_t)pv.views[v].width * 2;
    }
  }
  auto branch_frame = [&](int v, int b, int pp) {
    return reinterpret_cast<uint8_t*>(ws + L.off_branch + (branch_base[v] + (size_t)b * 2 + pp) * L.frame_bytes);
  };

  size_t to = L.off_tables;
  auto carve = [&](size_t bytes) {
    size_t at = to;
    to = align_up_sz(to + bytes, 16);
    return at;
  };
  const size_t t_lanes = carve((size_t)L.n_lanes_total * sizeof(Lane));
  const size_t t_lut = carve((size_t)(L.n_lut > 0 ? L.n_lut : 1) * sizeof(LutJob));
  const size_t t_chain = carve((size_t)(L.n_chains > 0 ? L.n_chains : 1) * sizeof(Chain));
  const size_t t_bjob = carve((size_t)(L.n_bbo > 0 ? L.n_bbo : 1) * sizeof(BboJob));
  const size_t t_items = carve((size_t)L.max_items * sizeof(Item));
  const size_t t_deps = carve((size_t)L.max_deps * sizeof(int32_t));
  const size_t t_succ = carve((size_t)L.max_deps * sizeof(int32_t));
  const size_t t_pending = carve((size_t)L.max_items * sizeof(int32_t));
  const size_t t_perm = carve((size_t)L.max_perm * sizeof(int32_t));
  const size_t t_mix = carve((size_t)h.n_views * sizeof(MixJob));
  const size_t t_norm = carve(3 * 256 * sizeof(float));
  to = align_up_sz(to, 128);   // L.off_plan is 256-byte aligned: the maps are 128-byte aligned on the device too
  const size_t t_maps = carve((size_t)L.max_maps * kTensorMapBytes);
  const size_t t_ring = carve((size_t)(L.max_items + kRingHeader) * sizeof(unsigned long long));
  auto* lanes = reinterpret_cast<Lane*>(stage.data() + t_lanes);
  auto* lutjobs = reinterpret_cast<LutJob*>(stage.data() + t_lut);
  auto* chains = reinterpret_cast<Chain*>(stage.data() + t_chain);
  auto* bjobs = reinterpret_cast<BboJob*>(stage.data() + t_bjob);
  auto* items = reinterpret_cast<Item*>(stage.data() + t_items);
  auto* deps = reinterpret_cast<int32_t*>(stage.data() + t_deps);
  auto* succ = reinterpret_cast<int32_t*>(stage.data() + t_succ);
  auto* pending = reinterpret_cast<int32_t*>(stage.data() + t_pending);
  auto* perm = reinterpret_cast<int32_t*>(stage.data() + t_perm);
  auto* mixjobs = reinterpret_cast<MixJob*>(stage.data() + t_mix);
  auto* ring = reinterpret_cast<unsigned long long*>(stage.data() + t_ring);
  // tensor maps by (frame, bytes per pixel): encoded once per distinct frame of this plan
  struct MapKey {
    const void* base;
    int slot;
  };
  std::vector<MapKey> map_keys;
  int n_maps = 0;
  auto map_for = [&](const void* base, int W_px, int H_px, int bpp) -> int {
    for (const MapKey& k : map_keys)
      if (k.base == base) return k.slot;
    int slot = -1;
    if (n_maps < L.max_maps &&
        be.make_map(stage.data() + t_maps + (size_t)n_maps * kTensorMapBytes, base, (size_t)W_px * bpp, H_px,
                    bpp == 3 ? kGatherImgBoxBytes : kGatherMaskBoxBytes) == 0)
      slot = n_maps++;
    map_keys.push_back(MapKey{base, slot});
    return slot;
  };

  OADG_T("stage alloc");
  // ---- items with their earliest phase (dependencies are always scheduled before their consumers) ----------
  // `phase` (the length of the longest dependency chain below an item) only orders the work queue; what an item
  // actually waits for is its `deps` list (indices into `todo`).
  struct Todo {
    int kind, obj, phase, w, hgt, aux;   // w x hgt: pixel extent (2-D kinds) / w = linear size (1-D kinds)
    std::vector<int> deps;
  };
  std::vector<Todo> todo;
  todo.reserve(L.max_items);
  int n_phases = 0;
  auto add = [&](int kind, int obj, int phase, int w, int hgt, int aux) {
    todo.push_back(Todo{kind, obj, phase, w, hgt, aux, {}});
    n_phases = phase + 1 > n_phases ? phase + 1 : n_phases;
    return (int)todo.size() - 1;
  };
  auto dep = [&](int item, int on) {
    if (on >= 0) todo[item].deps.push_back(on);
  };
  const int prof_done = h.n_gt > 0 ? 1 : 0;
  std::vector<int> prof_item((size_t)h.n_gt * 2, -1);
  for (int g = 0; g < h.n_gt; ++g)
    for (int axis = 0; axis < 2; ++axis) prof_item[g * 2 + axis] = add(OADG_IT_PROFILE, g * 2 + axis, 0, 1, 1, 0);
  int mask_done = prof_done;
  std::vector<int> mask_item(h.n_views, -1);
  if (L.any_bg) {  // union mask of every view (views without gt boxes get zeros)
    for (int v = 0; v < h.n_views; ++v) {
      const oadg_view_t& V = pv.views[v];
      mask_item[v] = add(OADG_IT_MASK, v, prof_done, V.W, V.H, 0);
      for (int g = V.gt_first; g < V.gt_first + V.n_gt; ++g) {
        dep(mask_item[v], prof_item[2 * g]);
        dep(mask_item[v], prof_item[2 * g + 1]);
      }
    }
    mask_done = prof_done + 1;
  }

  int lane_n = 0, hist_n = 0, lut_n = 0, chain_n = 0, bjob_n = 0;
  std::vector<const uint8_t*> final_frame((size_t)h.n_views * OADG_MAX_WIDTH, nullptr);
  std::vector<int> step_phase((size_t)h.n_views * OADG_MAX_WIDTH, -1);  // phase of the lane's previous step
  std::vector<int> step_item((size_t)h.n_views * OADG_MAX_WIDTH, -1);   // ... and its item
  struct HistKey {
    const uint8_t* in;
    int slot, done, item;
  };
  std::vector<HistKey> hist_keys;
  ChainSched cs;
  for (int d = 0; d < L.max_depth; ++d) {
    hist_keys.clear();  // branch frames are recycled every other depth: a key is only valid within one depth
    for (int v = 0; v < h.n_views; ++v) {
      const oadg_view_t& V = pv.views[v];
      for (int b = 0; b < V.width; ++b) {
        if (V.depth[b] <= d) continue;
        const int lane_id = lane_n++;
        Lane& ln = lanes[lane_id];
        ln.view = v;
        ln.branch = b;
        ln.op_base = op_index(V, b, d, 0);
        ln.in = d == 0 ? src[V.img] : branch_frame(v, b, (d - 1) & 1);
        ln.out = branch_frame(v, b, d & 1);
        ln.hist_slot = -1;
        ln.H = V.H;
        ln.W = V.W;
        ln.n_ml = V.n_ml;
        ln.all_streaming = 1;
        ln.in_map = ln.mask_map = -1;
        for (int q = 0; q < 2; ++q)
          for (int e = 0; e < 4; ++e) ln.box[q][e] = q < V.n_ml ? V.ml_box[q][e] : 0;
        for (int r = 0; r < OADG_MAX_REGIONS; ++r) ln.kind[r] = ln.lut[r] = ln.scratch[r] = -1;
        final_frame[(size_t)v * OADG_MAX_WIDTH + b] = ln.out;
        const int ready = step_phase[(size_t)v * OADG_MAX_WIDTH + b] + 1;  // phase in which ln.in is complete
        const int prev_step = step_item[(size_t)v * OADG_MAX_WIDTH + b];   // the item that writes ln.in (-1: source)
        std::vector<int> step_deps;
        if (prev_step >= 0) step_deps.push_back(prev_step);
        int step_at = ready;
        bool hist = false;
        for (int r = 0; r <= V.n_ml; ++r) hist |= needs_hist(ops[ln.op_base + r].kind);
        int hist_done = 0, hist_item = -1;
        if (hist) {  // lanes of one view share the histogram of the source frame at depth 0
          int found = -1;
          for (size_t k = 0; k < hist_keys.size(); ++k)
            if (hist_keys[k].in == ln.in) found = (int)k;
          if (found < 0) {
            hist_keys.push_back(HistKey{ln.in, hist_n++, ready + 1, -1});
            found = (int)hist_keys.size() - 1;
            ln.hist_slot = hist_keys[found].slot;
            hist_keys[found].item = add(OADG_IT_HIST, lane_id, ready, V.W * V.H, 1, 0);
            dep(hist_keys[found].item, prev_step);
          }
          ln.hist_slot = hist_keys[found].slot;
          hist_done = hist_keys[found].done;
          hist_item = hist_keys[found].item;
        }
        for (int r = 0; r <= V.n_ml; ++r) {
          oadg_op_t& op = ops[ln.op_base + r];
          op.lut = -1;
          op.scratch = -1;
          if (is_lut_kind(op.kind)) {
            op.lut = lut_n;
            lutjobs[lut_n] = LutJob{ln.op_base + r, ln.hist_slot, v, 0};
            const int at = needs_hist(op.kind) ? hist_done : 0;
            const int li = add(OADG_IT_LUT, lut_n, at, 1, 1, 0);
            if (needs_hist(op.kind)) dep(li, hist_item);
            step_deps.push_back(li);
            step_at = at + 1 > step_at ? at + 1 : step_at;
            ++lut_n;
          } else if (op.kind == OADG_OP_BBO_AFFINE && op.bbo_count > 0) {
            IRect keep{0, 0, V.W, V.H};
            if (r < V.n_ml) keep = IRect{V.ml_box[r][0], V.ml_box[r][1], V.ml_box[r][2], V.ml_box[r][3]};
            schedule_chain(pv, V, op, &keep, cs);
            if (!cs.box.empty()) {
              const int c_id = chain_n++;
              Chain& c = chains[c_id];
              c.view = v;
              c.lane = lane_id;
              c.in = ln.in;
              c.S = reinterpret_cast<uint8_t*>(ws + L.off_scratch + (size_t)(2 * c_id) * L.frame_bytes);
              c.T = reinterpret_cast<uint8_t*>(ws + L.off_scratch + (size_t)(2 * c_id + 1) * L.frame_bytes);
              c.map_in = map_for(c.in, V.W, V.H, 3);
              c.map_S = map_for(c.S, V.W, V.H, 3);
              c.map_T = map_for(c.T, V.W, V.H, 3);
              c.pad = 0;
              const int NL = cs.n_levels;
              op.scratch = 2 * c_id + (NL & 1);  // the step reads Y of the last level: T when NL is odd, else S
              // T (and S when a second level exists) start as copies of the lane input
              const int copy_item = add(OADG_IT_COPY, c_id, ready, V.W * V.H * 3, 1, NL >= 2 ? 1 : 0);
              dep(copy_item, prev_step);
              const int r0 = (ready + 1 > prof_done ? ready + 1 : prof_done);
              // jobs of the chain sorted by level (stable): level l occupies [first[l], first[l+1])
              std::vector<int> first(NL + 2, 0);
              for (size_t k = 0; k < cs.box.size(); ++k) ++first[cs.level[k] + 1];
              for (int l = 1; l <= NL + 1; ++l) first[l] += first[l - 1];
              std::vector<int> fill(first.begin(), first.end());
              const int job0 = bjob_n;
              for (size_t k = 0; k < cs.box.size(); ++k) {
                const int l = cs.level[k];
                BboJob& J = bjobs[job0 + fill[l]++];
                J.chain = c_id;
                J.bbo = op.bbo_first + cs.box[k];
                J.level = l;
                J.next_first = job0 + first[l + 1];
                J.next_count = l < NL ? first[l + 2] - first[l + 1] : 0;
                J.pad = 0;
                const int32_t* s = pv.gts[pv.bbo[J.bbo].gt].supp;
                for (int e = 0; e < 4; ++e) J.rect[e] = s[e];
              }
              bjob_n += (int)cs.box.size();
              // items that run "in level l": the blends of level l and the catch-ups of the level l-1 boxes.  They
              // read the frame level l-1 wrote and write the frame level l-1 read, so each waits for ALL items of
              // level l-1 (level 1: for the initial copy); a blend also needs its box's two mask profiles.
              std::vector<std::vector<int>> lvl_items(NL + 1);
              for (int k = job0; k < bjob_n; ++k) {
                const BboJob& J = bjobs[k];
                const int w = J.rect[2] - (J.rect[0] & ~3), hg = J.rect[3] - J.rect[1];  // tiles on a 4-px grid
                const int ri = add(OADG_IT_BBO_R, k, r0 + J.level - 1, w, hg, 0);
                lvl_items[J.level].push_back(ri);
                const int g = pv.bbo[J.bbo].gt;
                dep(ri, prof_item[2 * g]);
                dep(ri, prof_item[2 * g + 1]);
                if (J.level < NL)   // caught up by the next level
                  lvl_items[J.level + 1].push_back(add(OADG_IT_BBO_C, k, r0 + J.level, w, hg, 0));
              }
              for (int l = 1; l <= NL; ++l)
                for (int it : lvl_items[l]) {
                  if (l == 1) dep(it, copy_item);
                  else
                    for (int pr : lvl_items[l - 1]) dep(it, pr);
                }
              for (int it : lvl_items[NL]) step_deps.push_back(it);
              const int done = r0 + NL;
              step_at = done > step_at ? done : step_at;
            }
          } else if (op.kind == OADG_OP_BG_AFFINE) {
            step_at = mask_done > step_at ? mask_done : step_at;
            step_deps.push_back(mask_item[v]);
            ln.in_map = map_for(ln.in, V.W, V.H, 3);
            ln.mask_map = map_for(ws + L.off_masku + (size_t)v * h.max_h * h.max_w, V.W, V.H, 1);
          }
        }
        for (int r = 0; r <= V.n_ml; ++r) {
          ln.kind[r] = ops[ln.op_base + r].kind;
          ln.lut[r] = ops[ln.op_base + r].lut;
          ln.scratch[r] = ops[ln.op_base + r].scratch;
          if (!(is_lut_kind(ln.kind[r]) || ln.kind[r] == OADG_OP_BBO_AFFINE)) ln.all_streaming = 0;
        }
        const int si = add(OADG_IT_STEP, lane_id, step_at, V.W, V.H, 0);
        for (int d2 : step_deps) dep(si, d2);
        step_phase[(size_t)v * OADG_MAX_WIDTH + b] = step_at;
        step_item[(size_t)v * OADG_MAX_WIDTH + b] = si;
      }
    }
  }
  OADG_T("todo build");
  if ((int)todo.size() > L.max_items || (int)todo.size() > 4096) return OADG_E_LIMIT;   // 4096: the kernel's item bitmap

  // ---- the work queue: items ordered by PRIORITY = the length of the longest dependency chain that still hangs on
  // an item (its "bottom level" in a list schedule with rough per-kind tile durations).  CTAs always take the first
  // READY item of the queue, so critical items are served as soon as their inputs are complete and the rest fills
  // the gaps.  The order need not be topological: an item whose inputs are missing is skipped, never waited for.
  const int n_items = (int)todo.size();
  std::vector<int> order(n_items), pos(n_items);
  {
    std::vector<double> dur(n_items, 0.0), bottom(n_items, 0.0), start(n_items, 0.0);
    const double n_cta = (double)G;
    for (int i = 0; i < n_items; ++i) {   // todo is topological (dependencies are created first)
      const Todo& t = todo[i];
      int tw = 1, th = 1;
      double us = 8.0;
      switch (t.kind) {
        case OADG_IT_MASK: tw = kMaskTileW; th = kMaskTileH; us = 5.6; break;
        case OADG_IT_HIST: tw = kHistTilePx; us = 20.0; break;
        case OADG_IT_LUT: us = 4.0; break;
        case OADG_IT_COPY: tw = kCopyTileBytes; us = 10.0; break;
        case OADG_IT_BBO_R: tw = kBboTileW; th = kBboTileH; us = 13.0; break;
        case OADG_IT_BBO_C: tw = kBboCatchW; th = kBboTileH; us = 7.0; break;
        case OADG_IT_STEP:
          tw = lanes[t.obj].all_streaming ? kStepTileW : kStepTileWPx;
          th = kStepTileH;
          us = lanes[t.obj].all_streaming ? 5.3 : 15.0;
          break;
        default: break;
      }
      const double tiles = (double)((t.w + tw - 1) / tw) * (double)((t.hgt + th - 1) / th);
      const double waves = tiles / n_cta < 1.0 ? 1.0 : tiles / n_cta;
      dur[i] = waves * us + 3.0;   // + hand-over latency between dependent items
      double s0 = 0.0;
      for (int d2 : t.deps) s0 = start[d2] + dur[d2] > s0 ? start[d2] + dur[d2] : s0;
      start[i] = s0;
    }
    for (int i = n_items - 1; i >= 0; --i) {   // longest path from the start of item i to the end of the queue
      bottom[i] += dur[i];
      for (int d2 : todo[i].deps) bottom[d2] = bottom[i] > bottom[d2] ? bottom[i] : bottom[d2];
    }
    for (int i = 0; i < n_items; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
      if (bottom[x] != bottom[y]) return bottom[x] > bottom[y];
      return start[x] < start[y];
    });
    for (int k = 0; k < n_items; ++k) pos[order[k]] = k;
  }
  OADG_T("priority sort");
  int n_tiles = 0, n_deps = 0, n_perm = 0;
  std::vector<uint8_t> tile_cls;
  for (int k = 0; k < n_items; ++k) {
    const Todo& t = todo[order[k]];
    Item& it = items[k];
    it.kind = t.kind;
    it.obj = t.obj;
    it.tile0 = n_tiles;
    it.aux = t.aux;
    int tw = 1, th = 1;
    switch (t.kind) {
      case OADG_IT_MASK: tw = kMaskTileW; th = kMaskTileH; break;
      case OADG_IT_HIST: tw = kHistTilePx; break;
      case OADG_IT_COPY: tw = kCopyTileBytes; break;
      case OADG_IT_BBO_R: tw = kBboTileW; th = kBboTileH; break;
      case OADG_IT_BBO_C: tw = kBboCatchW; th = kBboTileH; break;
      case OADG_IT_STEP:   // narrower tiles for lanes with per-pixel ops: their tiles are long
        tw = it.aux = lanes[t.obj].all_streaming ? kStepTileW : kStepTileWPx;
        th = kStepTileH;
        break;
      default: break;
    }
    it.tx = (t.w + tw - 1) / tw;
    it.ntiles = it.tx * ((t.hgt + th - 1) / th);
    if (it.ntiles < 0) it.ntiles = 0;
    n_tiles += it.ntiles;
    it.perm_first = -1;
    it.pad = 0;
    if (t.kind == OADG_IT_STEP && !lanes[t.obj].all_streaming && n_perm + it.ntiles <= L.max_perm) {
      // long tiles first: the item is complete when its LAST tile is, so the tail should be made of short ones
      const Lane& ln = lanes[t.obj];
      it.perm_first = n_perm;
      bool any_bg = false;
      for (int r = 0; r <= ln.n_ml; ++r) any_bg |= ln.kind[r] == OADG_OP_BG_AFFINE;
      tile_cls.resize((size_t)it.ntiles);
      int cls_count[4] = {0, 0, 0, 0};
      for (int ti = 0; ti < it.ntiles; ++ti) {   // class of every tile: 3 box edge + bg-only, 2 bg-only, 1 per pixel, 0 stream
        const int x0 = (ti % it.tx) * tw, y0 = (ti / it.tx) * th;
        const int x1 = x0 + tw < ln.W ? x0 + tw : ln.W, y1 = y0 + th < ln.H ? y0 + th : ln.H;
        int region = ln.n_ml, c = 0;
        bool edge = false;
        for (int bb = 0; bb < ln.n_ml; ++bb) {
          const int32_t* B = ln.box[bb];
          if (!(B[0] < x1 && B[2] > x0 && B[1] < y1 && B[3] > y0)) continue;
          if (B[0] <= x0 && B[2] >= x1 && B[1] <= y0 && B[3] >= y1) region = bb;
          else edge = true;
        }
        if (edge) c = any_bg ? 3 : 1;
        else if (ln.kind[region] == OADG_OP_BG_AFFINE) c = 2;
        else c = (is_lut_kind(ln.kind[region]) || ln.kind[region] == OADG_OP_BBO_AFFINE) ? 0 : 1;
        tile_cls[ti] = (uint8_t)c;
        ++cls_count[c];
      }
      int at[4];   // counting sort, classes in descending order, tiles of a class in ascending order
      at[3] = n_perm;
      at[2] = at[3] + cls_count[3];
      at[1] = at[2] + cls_count[2];
      at[0] = at[1] + cls_count[1];
      for (int ti = 0; ti < it.ntiles; ++ti) perm[at[tile_cls[ti]]++] = ti;
      n_perm += it.ntiles;
    }
    it.dep_first = n_deps;
    it.dep_count = 0;
    for (int d2 : t.deps) {
      if (n_deps >= L.max_deps) return OADG_E_LIMIT;
      deps[n_deps++] = pos[d2];
      ++it.dep_count;
    }
  }
  OADG_T("items+perm");
  {  // successor lists (the reverse of deps) and the number of dependency tiles every item starts with
    std::vector<int> cnt(n_items + 1, 0);
    for (int k = 0; k < n_items; ++k) {
      pending[k] = 0;
      for (int d2 = 0; d2 < items[k].dep_count; ++d2) {
        const int pr = deps[items[k].dep_first + d2];
        ++cnt[pr + 1];
        pending[k] += items[pr].ntiles;
      }
    }
    for (int k = 0; k < n_items; ++k) cnt[k + 1] += cnt[k];
    for (int k = 0; k < n_items; ++k) {
      items[k].succ_first = cnt[k];
      items[k].succ_count = 0;
    }
    for (int k = 0; k < n_items; ++k)
      for (int d2 = 0; d2 < items[k].dep_count; ++d2) {
        Item& pr = items[deps[items[k].dep_first + d2]];
        succ[pr.succ_first + pr.succ_count++] = k;
      }
  }

  if ((size_t)n_tiles > L.max_tiles) return OADG_E_LIMIT;
  {  // tickets of the items without dependencies, in queue (priority) order; the rest is published on the device
    unsigned long long n_init = 0, first = 0;
    for (int k = 0; k < n_items; ++k)
      if (pending[k] == 0) {
        ring[kRingHeader + n_init++] = (first << 32) | (unsigned)k;
        first += (unsigned)items[k].ntiles;
      }
    ring[0] = first;
    ring[1] = 0;
    ring[2] = n_init;
    ring[3] = 0;
  }

  for (int v = 0; v < h.n_views; ++v) {
    const oadg_view_t& V = pv.views[v];
    MixJob& J = mixjobs[v];
    J.view = v;
    J.src = src[V.img];
    J.out = dst[v];
    for (int b = 0; b < V.width; ++b) J.branch[b] = final_frame[(size_t)v * OADG_MAX_WIDTH + b];
    J.f32_out = J.f32_src = nullptr;
    J.Wp = J.Hp = 0;
    if (fused) {
      const int d = fused->size_divisor;
      J.Wp = (V.W + d - 1) / d * d;
      J.Hp = (V.H + d - 1) / d * d;
      J.f32_out = fused->view_f32_dev[v];
      J.f32_src = fused->src_f32_dev ? fused->src_f32_dev[V.img] : nullptr;
    }
  }
  if (fused) {
    // normalised value of every uint8 level per output plane, rounded exactly like mmcv.imnormalize's
    // cv2.subtract(float32 img, float64 mean) then cv2.multiply(float32 img, float64 1/std): float64 arithmetic,
    // float32 result, twice
    if (fused->size_divisor < 1 || !fused->view_f32_dev) return OADG_E_ARG;
    for (int v = 0; v < h.n_views; ++v)
      if (!fused->view_f32_dev[v]) return OADG_E_ARG;
    float* lut = reinterpret_cast<float*>(stage.data() + t_norm);
    for (int k = 0; k < 3; ++k) {
      const double mean = (double)fused->mean[k], stdinv = 1.0 / (double)fused->std[k];
      for (int i = 0; i < 256; ++i) {
        const float d1 = (float)((double)(float)i - mean);
        lut[k * 256 + i] = (float)((double)d1 * stdinv);
      }
    }
  }

  OADG_T("succ+ring+mix");
  // upload plan + tables in one copy; clear the barrier counter and the histograms in one memset
  rc = be.upload(ws + L.off_plan, stage.data(), stage.size());
  if (rc) return rc;
  rc = be.zero(ws + L.off_zero, L.zero_bytes);
  if (rc) return rc;
  const char* dplan = ws + L.off_plan;
  ChainArgs A;
  DevPlan& P = A.P;
  P.views = reinterpret_cast<const oadg_view_t*>(dplan + h.off_views);
  P.gts = reinterpret_cast<const oadg_gt_t*>(dplan + h.off_gt);
  P.ops = reinterpret_cast<const oadg_op_t*>(dplan + h.off_ops);
  P.bbo = reinterpret_cast<const oadg_bbo_t*>(dplan + h.off_bbo);
  P.tgts = reinterpret_cast<const oadg_target_t*>(dplan + h.off_tgt);
  P.prof_x = reinterpret_cast<const float*>(ws + L.off_prof_x);
  P.prof_y = reinterpret_cast<const float*>(ws + L.off_prof_y);
  P.max_w = h.max_w;
  P.max_h = h.max_h;
  P.luts = reinterpret_cast<const uint8_t*>(ws + L.off_lut);
  P.maskf = reinterpret_cast<const float*>(ws + L.off_maskf);
  P.masku = reinterpret_cast<const uint8_t*>(ws + L.off_masku);
  P.mask_stride = (size_t)h.max_h * h.max_w;
  P.norm_lut = reinterpret_cast<const float*>(dplan + t_norm);
  P.norm_rgb = fused ? (fused->to_rgb != 0) : 0;
  P.pad0 = 0;
  A.lanes = reinterpret_cast<const Lane*>(dplan + t_lanes);
  A.lutjobs = reinterpret_cast<const LutJob*>(dplan + t_lut);
  A.chains = reinterpret_cast<const Chain*>(dplan + t_chain);
  A.bjobs = reinterpret_cast<const BboJob*>(dplan + t_bjob);
  A.items = reinterpret_cast<const Item*>(dplan + t_items);
  A.deps = reinterpret_cast<const int32_t*>(dplan + t_deps);
  A.succ = reinterpret_cast<const int32_t*>(dplan + t_succ);
  A.perm = reinterpret_cast<const int32_t*>(dplan + t_perm);
  A.pending = reinterpret_cast<int32_t*>(const_cast<char*>(dplan) + t_pending);
  A.n_items = n_items;
  A.n_tiles = n_tiles;
  A.grid = G;
  A.epoch = reinterpret_cast<unsigned*>(ws + L.off_zero);
  A.claimed = reinterpret_cast<unsigned*>(ws + L.off_zero + 512);
  A.item_ts = reinterpret_cast<unsigned long long*>(ws + L.off_zero + 512 + align_up_sz((size_t)L.max_items * sizeof(unsigned), 8));
  A.hist = reinterpret_cast<unsigned*>(ws + L.off_hist);
  A.luma = reinterpret_cast<unsigned long long*>(ws + L.off_luma);
  A.luts = reinterpret_cast<uint8_t*>(ws + L.off_lut);
  A.prof_x = reinterpret_cast<float*>(ws + L.off_prof_x);
  A.prof_y = reinterpret_cast<float*>(ws + L.off_prof_y);
  A.maskf = reinterpret_cast<float*>(ws + L.off_maskf);
  A.masku = reinterpret_cast<uint8_t*>(ws + L.off_masku);
  A.scratch = reinterpret_cast<const uint8_t*>(ws + L.off_scratch);
  A.frame_bytes = L.frame_bytes;
  A.debug = 0;
  A.kind_ns = reinterpret_cast<unsigned long long*>(ws + L.off_zero + 64);
  A.fault = reinterpret_cast<unsigned*>(ws + L.off_zero + 16);
  A.maps = dplan + t_maps;
  A.ring = reinterpret_cast<unsigned long long*>(const_cast<char*>(dplan) + t_ring);
  A.tickets = reinterpret_cast<unsigned long long*>(ws + L.off_tickets);
  // host views of the same tables (the host arithmetic check interprets them directly)
  ChainArgs Hh = A;
  Hh.lanes = lanes;
  Hh.lutjobs = lutjobs;
  Hh.chains = chains;
  Hh.bjobs = bjobs;
  Hh.items = items;
  Hh.deps = deps;
  Hh.succ = succ;
  Hh.perm = perm;
  Hh.pending = pending;
  Hh.ring = ring;
  OADG_T("upload+zero");
  if ((rc = be.chain(A, Hh, pv))) return rc;
  if ((rc = be.mix(P, reinterpret_cast<const MixJob*>(dplan + t_mix), h.n_views))) return rc;
  if (fused) {   // Pad (transforms.py:573-640): zeros right of column W and below row H, in every plane
    for (int v = 0; v < h.n_views; ++v) {
      const oadg_view_t& V = pv.views[v];
      const MixJob& J = mixjobs[v];
      float* bufs[2] = {J.f32_out, J.f32_src};
      for (float* b : bufs) {
        if (!b) continue;
        for (int k = 0; k < 3; ++k) {
          float* plane = b + (size_t)k * J.Hp * J.Wp;
          if (J.Wp > V.W && (rc = be.zero2d(plane + V.W, (size_t)J.Wp * 4, (size_t)(J.Wp - V.W) * 4, (size_t)V.H))) return rc;
          if (J.Hp > V.H && (rc = be.zero(plane + (size_t)V.H * J.Wp, (size_t)(J.Hp - V.H) * J.Wp * 4))) return rc;
        }
      }
    }
  }
  return 0;
}

}  // namespace oadg
