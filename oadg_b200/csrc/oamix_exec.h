// Host orchestration of an OA-Mix plan: validation, workspace layout, launch tables and the
// launch sequence, written against a small Backend so that the CUDA launcher (oamix.cu) and
// the host arithmetic check (tests/hostsim, test infrastructure only) share it verbatim.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "oamix_body.h"

namespace oadg {

constexpr int kMagic = 0x4F414447;

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
  size_t off_plan, off_prof_x, off_prof_y, off_branch, off_scratch, off_hist, off_luma, off_lut, off_tables;
  size_t off_maskf, off_masku;
  int any_bg;
  size_t tables_bytes;
  int n_lanes_total, n_scratch, n_lut, n_hist, max_depth;
  size_t total;
};

// Everything is sized from the plan alone.
inline void make_layout(const PlanView& pv, Layout& L) {
  const oadg_plan_header_t& h = *pv.h;
  L.frame_bytes = align_up_sz((size_t)h.max_h * h.max_w * 3, 256);
  int max_depth = 0, lanes_total = 0, n_lut = 0, n_hist = 0;
  int bbo_at_depth[OADG_MAX_DEPTH] = {0};
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
          if (op.kind == OADG_OP_BBO_AFFINE && op.bbo_count > 0) ++bbo_at_depth[d];
          if (op.kind == OADG_OP_BG_AFFINE) any_bg = 1;
        }
        if (hist) ++n_hist;
      }
    }
  }
  int n_scratch = 0;
  for (int d = 0; d < OADG_MAX_DEPTH; ++d) n_scratch = bbo_at_depth[d] > n_scratch ? bbo_at_depth[d] : n_scratch;
  L.max_depth = max_depth;
  L.n_lanes_total = lanes_total;
  L.n_scratch = n_scratch * 2;  // S + T per concurrent chain
  L.n_lut = n_lut;
  L.n_hist = n_hist;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o = align_up_sz(o + bytes, 256);
    return at;
  };
  L.tables_bytes = align_up_sz((size_t)lanes_total * sizeof(Lane), 16) +
                   2 * align_up_sz((size_t)lanes_total * sizeof(int32_t), 16) +
                   align_up_sz((size_t)(n_lut > 0 ? n_lut : 1) * sizeof(LutJob), 16) +
                   align_up_sz((size_t)lanes_total * OADG_MAX_REGIONS * sizeof(Chain), 16) +
                   align_up_sz((size_t)h.n_views * sizeof(MixJob), 16) + 256;
  // plan blob and launch tables are contiguous so that one H2D copy uploads both
  L.off_tables = align_up_sz((size_t)h.total_bytes, 16);
  L.off_plan = take(L.off_tables + L.tables_bytes);
  L.off_prof_x = take((size_t)h.n_gt * h.max_w * sizeof(float));
  L.off_prof_y = take((size_t)h.n_gt * h.max_h * sizeof(float));
  size_t n_branch_frames = 0;
  for (int v = 0; v < h.n_views; ++v) n_branch_frames += (size_t)pv.views[v].width * 2;
  L.off_branch = take(n_branch_frames * L.frame_bytes);
  L.off_scratch = take((size_t)L.n_scratch * L.frame_bytes);
  L.off_hist = take((size_t)(n_hist > 0 ? n_hist : 1) * 768 * sizeof(unsigned));
  L.off_luma = take((size_t)(n_hist > 0 ? n_hist : 1) * sizeof(unsigned long long));
  L.off_lut = take((size_t)(n_lut > 0 ? n_lut : 1) * 768);
  L.any_bg = any_bg;
  const size_t mask_px = (size_t)h.max_h * h.max_w;
  L.off_maskf = take(any_bg ? (size_t)h.n_views * mask_px * sizeof(float) : 0);
  L.off_masku = take(any_bg ? (size_t)h.n_views * mask_px : 0);
  L.total = o;
}

// Backend concept (all return 0 or an error code):
//   upload(dst, src_host, bytes)  zero(dst, bytes)  copy(dst, src, bytes)
//   profiles(P, pv, prof_x, prof_y)           masks(P, n_views, maskf, masku)
//   hist(P, lanes, lane_ids, n, hist, luma)   lut(P, jobs, n, hist, luma, luts)
//   bbo_pass(P, chains, n, j, roi_w, roi_h)
//   step(P, lanes, n, pixel_lane_ids, n_pixel_lanes, scratch, frame_bytes)   mix(P, jobs, n)
template <class Backend>
int execute_plan(Backend& be, const void* plan_host, size_t plan_bytes, const uint8_t* const* src, int n_img,
                 uint8_t* const* dst, void* workspace, size_t workspace_bytes) {
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
  char* ws = static_cast<char*>(workspace);

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
  const size_t t_lane_ids = carve((size_t)L.n_lanes_total * sizeof(int32_t));
  const size_t t_px_ids = carve((size_t)L.n_lanes_total * sizeof(int32_t));
  const size_t t_lut = carve((size_t)(L.n_lut > 0 ? L.n_lut : 1) * sizeof(LutJob));
  const size_t t_chain = carve((size_t)L.n_lanes_total * OADG_MAX_REGIONS * sizeof(Chain));
  const size_t t_mix = carve((size_t)h.n_views * sizeof(MixJob));
  auto* lanes = reinterpret_cast<Lane*>(stage.data() + t_lanes);
  auto* lane_ids = reinterpret_cast<int32_t*>(stage.data() + t_lane_ids);
  auto* px_ids = reinterpret_cast<int32_t*>(stage.data() + t_px_ids);
  auto* lutjobs = reinterpret_cast<LutJob*>(stage.data() + t_lut);
  auto* chains = reinterpret_cast<Chain*>(stage.data() + t_chain);
  auto* mixjobs = reinterpret_cast<MixJob*>(stage.data() + t_mix);

  struct DepthInfo {
    int lane0 = 0, n_lanes = 0, hist0 = 0, n_hist = 0, lut0 = 0, n_lut = 0, chain0 = 0, n_chain = 0;
    int max_chain = 0, max_roi_w = 0, max_roi_h = 0, n_px = 0;
  };
  std::vector<DepthInfo> di(L.max_depth);
  int lane_n = 0, hist_n = 0, lut_n = 0, chain_n = 0, histid_n = 0;
  std::vector<const uint8_t*> final_frame((size_t)h.n_views * OADG_MAX_WIDTH, nullptr);
  for (int d = 0; d < L.max_depth; ++d) {
    DepthInfo& D = di[d];
    D.lane0 = lane_n;
    D.hist0 = histid_n;
    D.lut0 = lut_n;
    D.chain0 = chain_n;
    int scratch_used = 0;
    for (int v = 0; v < h.n_views; ++v) {
      const oadg_view_t& V = pv.views[v];
      for (int b = 0; b < V.width; ++b) {
        if (V.depth[b] <= d) continue;
        Lane& ln = lanes[lane_n];
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
        for (int q = 0; q < 2; ++q)
          for (int e = 0; e < 4; ++e) ln.box[q][e] = q < V.n_ml ? V.ml_box[q][e] : 0;
        for (int r = 0; r < OADG_MAX_REGIONS; ++r) ln.kind[r] = ln.lut[r] = ln.scratch[r] = -1;
        final_frame[(size_t)v * OADG_MAX_WIDTH + b] = ln.out;
        bool hist = false;
        for (int r = 0; r <= V.n_ml; ++r) hist |= needs_hist(ops[ln.op_base + r].kind);
        if (hist) {
          ln.hist_slot = hist_n++;
          lane_ids[histid_n++] = lane_n;
          ++D.n_hist;
        }
        for (int r = 0; r <= V.n_ml; ++r) {
          oadg_op_t& op = ops[ln.op_base + r];
          op.lut = -1;
          op.scratch = -1;
          if (is_lut_kind(op.kind)) {
            op.lut = lut_n;
            lutjobs[lut_n] = LutJob{ln.op_base + r, ln.hist_slot, v, 0};
            ++lut_n;
            ++D.n_lut;
          } else if (op.kind == OADG_OP_BBO_AFFINE && op.bbo_count > 0) {
            Chain& c = chains[chain_n++];
            c.view = v;
            c.n = op.bbo_count;
            c.bbo_first = op.bbo_first;
            c.lane = lane_n;  // lane whose input seeds S
            op.scratch = scratch_used + (op.bbo_count & 1);  // result frame: T after an odd number of boxes
            c.S = reinterpret_cast<uint8_t*>(ws + L.off_scratch + (size_t)scratch_used * L.frame_bytes);
            c.T = reinterpret_cast<uint8_t*>(ws + L.off_scratch + (size_t)(scratch_used + 1) * L.frame_bytes);
            scratch_used += 2;
            ++D.n_chain;
            D.max_chain = c.n > D.max_chain ? c.n : D.max_chain;
            for (int j = 0; j < c.n; ++j) {  // pass j covers the bounding rect of supports j and j-1
              const int32_t* a = pv.gts[pv.bbo[c.bbo_first + j].gt].supp;
              int x0 = a[0], y0 = a[1], x1 = a[2], y1 = a[3];
              if (j > 0) {
                const int32_t* q = pv.gts[pv.bbo[c.bbo_first + j - 1].gt].supp;
                if (q[2] > q[0] && q[3] > q[1]) {
                  if (x1 <= x0 || y1 <= y0) { x0 = q[0]; y0 = q[1]; x1 = q[2]; y1 = q[3]; }
                  else { x0 = x0 < q[0] ? x0 : q[0]; y0 = y0 < q[1] ? y0 : q[1]; x1 = x1 > q[2] ? x1 : q[2]; y1 = y1 > q[3] ? y1 : q[3]; }
                }
              }
              D.max_roi_w = (x1 - x0) > D.max_roi_w ? (x1 - x0) : D.max_roi_w;
              D.max_roi_h = (y1 - y0) > D.max_roi_h ? (y1 - y0) : D.max_roi_h;
            }
          }
        }
        for (int r = 0; r <= V.n_ml; ++r) {
          ln.kind[r] = ops[ln.op_base + r].kind;
          ln.lut[r] = ops[ln.op_base + r].lut;
          ln.scratch[r] = ops[ln.op_base + r].scratch;
          if (!(is_lut_kind(ln.kind[r]) || ln.kind[r] == OADG_OP_BBO_AFFINE)) ln.all_streaming = 0;
        }
        if (!ln.all_streaming) px_ids[D.lane0 + D.n_px++] = lane_n - D.lane0;  // index within this depth's lane slice
        ++lane_n;
        ++D.n_lanes;
      }
    }
  }
  for (int v = 0; v < h.n_views; ++v) {
    const oadg_view_t& V = pv.views[v];
    MixJob& J = mixjobs[v];
    J.view = v;
    J.src = src[V.img];
    J.out = dst[v];
    for (int b = 0; b < V.width; ++b) J.branch[b] = final_frame[(size_t)v * OADG_MAX_WIDTH + b];
  }

  // upload plan + tables in one copy
  rc = be.upload(ws + L.off_plan, stage.data(), stage.size());
  if (rc) return rc;
  const char* dplan = ws + L.off_plan;
  DevPlan P;
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
  const Lane* d_lanes = reinterpret_cast<const Lane*>(dplan + t_lanes);
  const int32_t* d_lane_ids = reinterpret_cast<const int32_t*>(dplan + t_lane_ids);
  const int32_t* d_px_ids = reinterpret_cast<const int32_t*>(dplan + t_px_ids);
  const LutJob* d_lut = reinterpret_cast<const LutJob*>(dplan + t_lut);
  const Chain* d_chain = reinterpret_cast<const Chain*>(dplan + t_chain);
  const MixJob* d_mix = reinterpret_cast<const MixJob*>(dplan + t_mix);
  unsigned* d_hist = reinterpret_cast<unsigned*>(ws + L.off_hist);
  unsigned long long* d_luma = reinterpret_cast<unsigned long long*>(ws + L.off_luma);
  uint8_t* d_luts = reinterpret_cast<uint8_t*>(ws + L.off_lut);
  const uint8_t* d_scratch = reinterpret_cast<const uint8_t*>(ws + L.off_scratch);

  if (h.n_gt > 0) {
    rc = be.profiles(P, pv, const_cast<float*>(P.prof_x), const_cast<float*>(P.prof_y));
    if (rc) return rc;
  }
  if (L.any_bg) {  // union mask of every view, once per batch (views without gt boxes get zeros)
    rc = be.masks(P, h.n_views, const_cast<float*>(P.maskf), const_cast<uint8_t*>(P.masku));
    if (rc) return rc;
  }
  if (hist_n > 0) {
    rc = be.zero(d_hist, (size_t)hist_n * 768 * sizeof(unsigned));
    if (rc) return rc;
    rc = be.zero(d_luma, (size_t)hist_n * sizeof(unsigned long long));
    if (rc) return rc;
  }
  for (int d = 0; d < L.max_depth; ++d) {
    const DepthInfo& D = di[d];
    if (D.n_lanes == 0) continue;
    if (D.n_hist > 0 && (rc = be.hist(P, d_lanes, d_lane_ids + D.hist0, D.n_hist, d_hist, d_luma))) return rc;
    if (D.n_lut > 0 && (rc = be.lut(P, d_lut + D.lut0, D.n_lut, d_hist, d_luma, d_luts))) return rc;
    if (D.n_chain > 0) {
      for (int c = 0; c < D.n_chain; ++c) {
        const Chain& C = chains[D.chain0 + c];
        const oadg_view_t& V = pv.views[C.view];
        if ((rc = be.copy(C.S, lanes[C.lane].in, (size_t)V.H * V.W * 3))) return rc;
        if ((rc = be.copy(C.T, lanes[C.lane].in, (size_t)V.H * V.W * 3))) return rc;
      }
      if (D.max_roi_w > 0 && D.max_roi_h > 0) {
        for (int j = 0; j < D.max_chain; ++j)
          if ((rc = be.bbo_pass(P, d_chain + D.chain0, D.n_chain, j, D.max_roi_w, D.max_roi_h))) return rc;
      }
    }
    if ((rc = be.step(P, d_lanes + D.lane0, D.n_lanes, d_px_ids + D.lane0, D.n_px, d_scratch, L.frame_bytes))) return rc;
  }
  return be.mix(P, d_mix, h.n_views);
}

}  // namespace oadg
