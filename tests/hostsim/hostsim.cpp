// HOST ARITHMETIC CHECK -- test infrastructure only, never a product path.
//
// Compiles the exact per-pixel bodies (csrc/oamix_body.h, oamix_math.h) and the exact host
// orchestration (csrc/oamix_exec.h) of the OA-Mix CUDA executor for the CPU, with plain
// loops standing in for the kernel grids.  It lets the dev container (no GPU) check the
// plan sampler + kernel arithmetic against the oracle before GPU time is spent.  The
// product (oadg_b200) never loads this library; it fails loudly without CUDA.
//
//   g++ -O2 -ffp-contract=off -shared -fPIC -I include -I oadg_b200/csrc tests/hostsim/hostsim.cpp
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oamix_exec.h"
#include "oamix_tile.h"

using namespace oadg;

namespace {

struct HostBackend {
  int launches = 0;
  int n_cta = 5;          // > 1: adversarial item order (see chain()); 1: queue order
  int n_phases = 0, n_items = 0, n_bbo_jobs = 0;
  int grid() { return n_cta; }
  int upload(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); return 0; }
  int zero(void* dst, size_t bytes) { memset(dst, 0, bytes); return 0; }
  int zero2d(void* dst, size_t pitch, size_t width, size_t rows) {
    for (size_t r = 0; r < rows; ++r) memset(static_cast<char*>(dst) + r * pitch, 0, width);
    return 0;
  }
  int make_map(void*, const void*, size_t, int, int) { return -1; }   // no TMA on the host: the per-pixel bodies gather directly

  static void profile_item(const ChainArgs& A, int obj) {
    const DevPlan& P = A.P;
    const int sr = 4;
    const int g = obj >> 1, axis = obj & 1;
    const oadg_gt_t& G = P.gts[g];
    const oadg_view_t& V = P.views[G.view];
    const int n_hi = axis == 0 ? V.W : V.H, n_lo = n_hi / sr;
    const int lo = G.lo[axis], hi = G.lo[axis + 2];
    const int ks = axis == 0 ? G.kx : G.ky;
    const double sigma = axis == 0 ? G.sigma_x : G.sigma_y;
    float* out = axis == 0 ? A.prof_x + (size_t)g * P.max_w : A.prof_y + (size_t)g * P.max_h;
    if (n_lo <= 0) {
      for (int d = 0; d < n_hi; ++d) out[d] = 0.f;
      return;
    }
    std::vector<float> p(n_lo), kern(ks > 0 ? ks : 1);
    auto taps = [](int n, double sg, std::vector<float>& k) {
      k.assign(n > 0 ? n : 1, 0.f);
      const double s2 = -0.5 / (sg * sg);
      double sum = 0;
      for (int i = 0; i < n; ++i) {
        double x = i - (n - 1) * 0.5;
        sum += exp(s2 * x * x);
      }
      const double ksum = 1.0 / sum;
      for (int i = 0; i < n; ++i) {
        double x = i - (n - 1) * 0.5;
        k[i] = (float)(exp(s2 * x * x) * ksum);
      }
    };
    if (G.blur) {
      taps(ks, sigma, kern);
      // a window that lies entirely inside the box: the x profile carries 1.0 and the y profile cv2's saturated value
      float sat = 1.f;
      if (axis == 1) {
        std::vector<float> kx;
        taps(G.kx, G.sigma_x, kx);
        sat = cv_saturated_col(kern.data(), ks, cv_saturated_row(kx.data(), G.kx));
      }
      const int r = ks / 2, period = 2 * (n_lo - 1);
      for (int x = 0; x < n_lo; ++x) {
        double acc = 0;
        int hits = 0;
        for (int j = 0; j < ks; ++j) {
          int q = x + j - r;
          if (n_lo == 1) q = 0;
          else {
            if (q < 0) q = -q;
            q %= period;
            if (q >= n_lo) q = period - q;
          }
          if (q >= lo && q < hi) {
            acc += (double)kern[j];
            ++hits;
          }
        }
        p[x] = hits == ks ? sat : (float)acc;
      }
    } else {
      for (int x = 0; x < n_lo; ++x) p[x] = (x >= lo && x < hi) ? 1.f : 0.f;
    }
    const double scale = (double)n_lo / (double)n_hi;
    for (int d = 0; d < n_hi; ++d) {
      float f = (float)((d + 0.5) * scale - 0.5);
      int s = (int)floorf(f);
      float t = fsub(f, (float)s);
      if (s < 0) { s = 0; t = 0.f; }
      if (s >= n_lo - 1) { s = n_lo - 1; t = 0.f; }
      int s1 = s + 1 < n_lo - 1 ? s + 1 : n_lo - 1;
      out[d] = ffma(t, fsub(p[s1], p[s]), p[s]);   // cv2.resize (float32, linear): a + t * (b - a), fused
    }
  }
  static void lut_item(const ChainArgs& A, int job) {
    const DevPlan& P = A.P;
    const LutJob& J = A.lutjobs[job];
    const oadg_op_t& op = P.ops[J.op];
    uint8_t* out = A.luts + (size_t)op.lut * 768;
    if (op.kind == OADG_OP_AUTOCONTRAST || op.kind == OADG_OP_EQUALIZE) {
      for (int c = 0; c < 3; ++c) {
        const unsigned* h = A.hist + (size_t)J.hist_slot * 768 + c * 256;
        if (op.kind == OADG_OP_AUTOCONTRAST) lut_autocontrast_ch(h, out + c * 256);
        else lut_equalize_ch(h, out + c * 256);
      }
    } else {
      const oadg_view_t& V = P.views[J.view];
      const double ls = J.hist_slot >= 0 ? (double)A.luma[J.hist_slot] : 0.0;
      for (int i = 0; i < 256; ++i) {
        uint8_t v = lut_simple_at(op, i, ls, (double)((long long)V.H * V.W));
        out[i] = out[256 + i] = out[512 + i] = v;
      }
    }
  }
  // mirrors step_tile (oamix.cu): the vector pass over streaming runs, then the per-pixel pass over the rest
  static void step_item_tile(const ChainArgs& A, const Lane& L, int local, int tx, int tw) {
    const DevPlan& P = A.P;
    const int x0 = (local % tx) * tw, y0 = (local / tx) * kStepTileH;
    const int x1 = imin(x0 + tw, L.W), y1 = imin(y0 + kStepTileH, L.H);
    uint8_t luts[OADG_MAX_REGIONS * 768];
    for (int r = 0; r <= L.n_ml; ++r)
      if (L.lut[r] >= 0) memcpy(luts + r * 768, P.luts + (size_t)L.lut[r] * 768, 768);
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; x += kChunkPx) {
        const int nn = imin(kChunkPx, L.W - x);
        int region;
        if (!run_is_stream(L, x, y, nn, region)) continue;
        if (region >= 0 && kind_streams(L.kind[region])) {
          Chunk in;
          chunk_load(stream_src(L, region, A.scratch, A.frame_bytes) + ((size_t)y * L.W + x) * 3, nn, true, in);
          stream_chunk(L, region, luts + region * 768, A.scratch, A.frame_bytes, in, x, y, nn, true);
        } else {
          for (int i = 0; i < nn; ++i) stream_pixel(L, luts, A.scratch, A.frame_bytes, x + i, y);
        }
      }
    if (L.all_streaming) return;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        const int xc = x & ~(kChunkPx - 1);
        int region;
        if (run_is_stream(L, xc, y, imin(kChunkPx, L.W - xc), region)) continue;
        step_pixel(P, L, A.scratch, A.frame_bytes, x, y);
      }
  }

  // one tile of one item (the host twin of the handlers in oamix.cu)
  int run_tile(const ChainArgs& A, const Item& I, int local) {
    const DevPlan& P = A.P;
    switch (I.kind) {
      case OADG_IT_PROFILE: profile_item(A, I.obj); break;
      case OADG_IT_MASK: {
        const oadg_view_t& V = P.views[I.obj];
        const int x0 = (local % I.tx) * kMaskTileW, y0 = (local / I.tx) * kMaskTileH;
        for (int y = y0; y < imin(y0 + kMaskTileH, V.H); ++y)
          for (int x = x0; x < imin(x0 + kMaskTileW, V.W); ++x) mask_pixel(P, I.obj, x, y, A.maskf, A.masku);
        break;
      }
      case OADG_IT_HIST: {
        const Lane& L = A.lanes[I.obj];
        const size_t npx = (size_t)L.H * L.W;
        const size_t p0 = (size_t)local * kHistTilePx, p1 = p0 + kHistTilePx < npx ? p0 + kHistTilePx : npx;
        unsigned* hh = A.hist + (size_t)L.hist_slot * 768;
        unsigned long long ls = 0;
        for (size_t i = p0; i < p1; ++i) {
          const uint8_t* q = L.in + i * 3;
          ++hh[q[0]];
          ++hh[256 + q[1]];
          ++hh[512 + q[2]];
          ls += (unsigned)pil_luma(q[0], q[1], q[2]);
        }
        A.luma[L.hist_slot] += ls;
        break;
      }
      case OADG_IT_LUT: lut_item(A, I.obj); break;
      case OADG_IT_COPY: {
        const Chain& C = A.chains[I.obj];
        const size_t nbytes = (size_t)P.views[C.view].H * P.views[C.view].W * 3;
        const size_t b0 = (size_t)local * kCopyTileBytes, b1 = b0 + kCopyTileBytes < nbytes ? b0 + kCopyTileBytes : nbytes;
        memcpy(C.T + b0, C.in + b0, b1 - b0);
        if (I.aux) memcpy(C.S + b0, C.in + b0, b1 - b0);
        break;
      }
      case OADG_IT_BBO_R:
      case OADG_IT_BBO_C: {
        const BboJob& J = A.bjobs[I.obj];
        const Chain& C = A.chains[J.chain];
        const int level = I.kind == OADG_IT_BBO_R ? J.level : J.level + 1;
        const uint8_t* X = chain_src(C, level);
        uint8_t* Y = chain_dst(C, level);
        const int tw = I.kind == OADG_IT_BBO_R ? kBboTileW : kBboCatchW;
        const int x0 = (J.rect[0] & ~3) + (local % I.tx) * tw, y0 = J.rect[1] + (local / I.tx) * kBboTileH;
        for (int y = y0; y < imin(y0 + kBboTileH, J.rect[3]); ++y)
          for (int x = imax(x0, J.rect[0]); x < imin(x0 + tw, J.rect[2]); ++x) {
            if (I.kind == OADG_IT_BBO_R) bbo_r_pixel(P, C, P.bbo[J.bbo], X, Y, x, y);
            else bbo_c_pixel(A.bjobs, J, P.views[C.view].W, X, Y, x, y);
          }
        if (I.kind == OADG_IT_BBO_R && local == 0) ++n_bbo_jobs;
        break;
      }
      case OADG_IT_STEP: step_item_tile(A, A.lanes[I.obj], local, I.tx, I.aux); break;
      default: return -203;
    }
    return 0;
  }

  // The device drains the queue with hundreds of CTAs; an item may start as soon as its dependency list is complete.
  // The host twin therefore runs the items in an ADVERSARIAL order that honours nothing but the dependency table:
  // it always picks the LAST item of the queue whose dependencies are done (n_cta > 1), or the FIRST one like the
  // device does (n_cta == 1).  A missing dependency shows up as a wrong image in the tests.
  std::vector<int32_t> dump;   // {n_items, then per item: kind, obj, ntiles, aux, streaming, dep_count, deps...}
  bool want_dump = false;
  int chain(const ChainArgs& Adev, const ChainArgs& A, const PlanView&) {
    (void)Adev;
    if (want_dump) {
      dump.clear();
      dump.push_back(A.n_items);
      for (int k = 0; k < A.n_items; ++k) {
        const Item& I = A.items[k];
        dump.push_back(I.kind); dump.push_back(I.obj); dump.push_back(I.ntiles); dump.push_back(I.aux);
        dump.push_back(I.kind == OADG_IT_STEP ? A.lanes[I.obj].all_streaming : 0);
        int cls_n[4] = {0, 0, 0, 0};   // step items: tiles per class (0 stream, 1 per pixel, 2 bg-only, 3 box edge + bg-only)
        if (I.kind == OADG_IT_STEP) {
          const Lane& ln = A.lanes[I.obj];
          const int tw = I.aux, th = kStepTileH;
          for (int ti = 0; ti < I.ntiles; ++ti) {
            const int x0 = (ti % I.tx) * tw, y0 = (ti / I.tx) * th;
            const int x1 = x0 + tw < ln.W ? x0 + tw : ln.W, y1 = y0 + th < ln.H ? y0 + th : ln.H;
            int region = ln.n_ml, c = 0;
            bool edge = false, any_bg = false;
            for (int bb = 0; bb < ln.n_ml; ++bb) {
              const int32_t* B = ln.box[bb];
              if (!(B[0] < x1 && B[2] > x0 && B[1] < y1 && B[3] > y0)) continue;
              if (B[0] <= x0 && B[2] >= x1 && B[1] <= y0 && B[3] >= y1) region = bb;
              else edge = true;
            }
            for (int r = 0; r <= ln.n_ml; ++r) any_bg |= ln.kind[r] == OADG_OP_BG_AFFINE;
            if (edge) c = any_bg ? 3 : 1;
            else if (ln.kind[region] == OADG_OP_BG_AFFINE) c = 2;
            else c = (is_lut_kind(ln.kind[region]) || ln.kind[region] == OADG_OP_BBO_AFFINE) ? 0 : 1;
            ++cls_n[c];
          }
        }
        for (int c = 0; c < 4; ++c) dump.push_back(cls_n[c]);
        dump.push_back(I.dep_count);
        for (int d = 0; d < I.dep_count; ++d) dump.push_back(A.deps[I.dep_first + d]);
      }
      ++launches;
      return 0;   // tables only: nothing is executed
    }
    n_phases = 0;
    n_items += A.n_items;
    std::vector<char> finished(A.n_items, 0);
    int tiles_seen = 0;
    for (int k = 0; k < A.n_items; ++k) {
      const Item& I = A.items[k];
      if (I.tile0 != tiles_seen || I.ntiles < 0) return -200;
      tiles_seen += I.ntiles;
      for (int d = 0; d < I.dep_count; ++d)
        if (A.deps[I.dep_first + d] < 0 || A.deps[I.dep_first + d] >= A.n_items || A.deps[I.dep_first + d] == k) return -201;
    }
    if (tiles_seen != A.n_tiles) return -202;
    for (int left = A.n_items; left > 0; --left) {
      int pick = -1;
      for (int k = 0; k < A.n_items; ++k) {
        const int cand = n_cta > 1 ? A.n_items - 1 - k : k;
        if (finished[cand]) continue;
        bool ready = true;
        const Item& I = A.items[cand];
        for (int d = 0; d < I.dep_count && ready; ++d) ready = finished[A.deps[I.dep_first + d]] != 0;
        if (ready) {
          pick = cand;
          break;
        }
      }
      if (pick < 0) return -204;
      const Item& I = A.items[pick];
      for (int t = I.ntiles - 1; t >= 0; --t) {   // tiles of an item are independent: any order
        const int tt = n_cta > 1 ? t : I.ntiles - 1 - t;
        const int rc = run_tile(A, I, I.perm_first >= 0 ? A.perm[I.perm_first + tt] : tt);
        if (rc) return rc;
      }
      finished[pick] = 1;
    }
    ++launches;
    return 0;
  }
  int mix(const DevPlan& P, const MixJob* jobs, int n) {
    if (want_dump) return 0;
    for (int k = 0; k < n; ++k) {
      const oadg_view_t& V = P.views[jobs[k].view];
      for (int y0 = 0; y0 < V.H; y0 += kTileH)
        for (int x0 = 0; x0 < V.W; x0 += kTileW) {
          const int x1 = imin(x0 + kTileW, V.W), y1 = imin(y0 + kTileH, V.H);
          MixTile T;
          classify_mix_tile(P, jobs[k], x0, y0, x1, y1, T);
          for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; x += kChunkPx) mix_chunk(P, jobs[k], T, x, y, imin(kChunkPx, x1 - x), true);
        }
    }
    ++launches;
    return 0;
  }
};

}  // namespace

extern "C" int hostsim_workspace_bytes(const void* plan, size_t bytes, size_t* out) {
  PlanView pv;
  int rc = parse_plan(plan, bytes, pv);
  if (rc) return rc;
  Layout L;
  make_layout(pv, L);
  *out = L.total;
  return 0;
}

extern "C" int hostsim_oamix_execute(const void* plan, size_t bytes, const uint8_t* const* src, int n_img,
                                     uint8_t* const* dst, int* launches_out) {
  size_t need = 0;
  int rc = hostsim_workspace_bytes(plan, bytes, &need);
  if (rc) return rc;
  void* ws = aligned_alloc(256, (need + 255) / 256 * 256 + 256);
  if (!ws) return -100;
  HostBackend be;
  rc = execute_plan(be, plan, bytes, src, n_img, dst, ws, need);
  if (launches_out) *launches_out = be.launches;
  free(ws);
  return rc;
}

// with the fused Normalize + Pad + CHW epilogue (host twin of oadg_oamix_execute_fused)
extern "C" int hostsim_oamix_execute_fused(const void* plan, size_t bytes, const uint8_t* const* src, int n_img,
                                           uint8_t* const* dst, const oadg_fused_out_t* fused) {
  size_t need = 0;
  int rc = hostsim_workspace_bytes(plan, bytes, &need);
  if (rc) return rc;
  void* ws = aligned_alloc(256, (need + 255) / 256 * 256 + 256);
  if (!ws) return -100;
  HostBackend be;
  rc = execute_plan(be, plan, bytes, src, n_img, dst, ws, need, fused);
  free(ws);
  return rc;
}

// same, with a chosen pretend grid; stats = {phases, items, bbo boxes executed}
extern "C" int hostsim_oamix_execute_ex(const void* plan, size_t bytes, const uint8_t* const* src, int n_img,
                                        uint8_t* const* dst, int n_cta, int* stats) {
  size_t need = 0;
  int rc = hostsim_workspace_bytes(plan, bytes, &need);
  if (rc) return rc;
  void* ws = aligned_alloc(256, (need + 255) / 256 * 256 + 256);
  if (!ws) return -100;
  HostBackend be;
  be.n_cta = n_cta;
  rc = execute_plan(be, plan, bytes, src, n_img, dst, ws, need);
  if (stats) {
    stats[0] = be.n_phases;
    stats[1] = be.n_items;
    stats[2] = be.n_bbo_jobs;
  }
  free(ws);
  return rc;
}

// measurement aid: the work queue (items, tiles, dependencies) the scheduler builds for a plan, without executing it
extern "C" int hostsim_oamix_dump(const void* plan, size_t bytes, const uint8_t* const* src, int n_img,
                                  uint8_t* const* dst, int32_t* out, int cap) {
  size_t need = 0;
  int rc = hostsim_workspace_bytes(plan, bytes, &need);
  if (rc) return rc;
  void* ws = aligned_alloc(256, (need + 255) / 256 * 256 + 256);
  if (!ws) return -100;
  HostBackend be;
  be.want_dump = true;
  rc = execute_plan(be, plan, bytes, src, n_img, dst, ws, need);
  free(ws);
  if (rc) return rc;
  const int n = (int)be.dump.size();
  if (out && n <= cap) memcpy(out, be.dump.data(), (size_t)n * sizeof(int32_t));
  return n;
}
