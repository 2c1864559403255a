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
  int upload(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); return 0; }
  int zero(void* dst, size_t bytes) { memset(dst, 0, bytes); return 0; }
  int copy(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); ++launches; return 0; }

  int profiles(const DevPlan& P, const PlanView& pv, float* px, float* py) {
    const int sr = 4;
    for (int g = 0; g < pv.h->n_gt; ++g)
      for (int axis = 0; axis < 2; ++axis) {
        const oadg_gt_t& G = P.gts[g];
        const oadg_view_t& V = P.views[G.view];
        const int n_hi = axis == 0 ? V.W : V.H, n_lo = n_hi / sr;
        const int lo = G.lo[axis], hi = G.lo[axis + 2];
        const int ks = axis == 0 ? G.kx : G.ky;
        const double sigma = axis == 0 ? G.sigma_x : G.sigma_y;
        float* out = axis == 0 ? px + (size_t)g * P.max_w : py + (size_t)g * P.max_h;
        if (n_lo <= 0) {
          for (int d = 0; d < n_hi; ++d) out[d] = 0.f;
          continue;
        }
        std::vector<float> p(n_lo), kern(ks > 0 ? ks : 1);
        if (G.blur) {
          const double s2 = -0.5 / (sigma * sigma);
          double sum = 0;
          for (int i = 0; i < ks; ++i) {
            double x = i - (ks - 1) * 0.5;
            sum += exp(s2 * x * x);
          }
          const double ksum = 1.0 / sum;
          for (int i = 0; i < ks; ++i) {
            double x = i - (ks - 1) * 0.5;
            kern[i] = (float)(exp(s2 * x * x) * ksum);
          }
          const int r = ks / 2, period = 2 * (n_lo - 1);
          for (int x = 0; x < n_lo; ++x) {
            double acc = 0;
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
          out[d] = fadd(fmul(p[s], fsub(1.f, t)), fmul(p[s1], t));
        }
      }
    ++launches;
    return 0;
  }
  int masks(const DevPlan& P, int n_views, float* maskf, uint8_t* masku) {
    for (int v = 0; v < n_views; ++v)
      for (int y = 0; y < P.views[v].H; ++y)
        for (int x = 0; x < P.views[v].W; ++x) mask_pixel(P, v, x, y, maskf, masku);
    ++launches;
    return 0;
  }
  int hist(const DevPlan& P, const Lane* lanes, const int32_t* ids, int n, unsigned* hist, unsigned long long* luma) {
    for (int k = 0; k < n; ++k) {
      const Lane& L = lanes[ids[k]];
      const oadg_view_t& V = P.views[L.view];
      unsigned* h = hist + (size_t)L.hist_slot * 768;
      unsigned long long ls = 0;
      for (size_t i = 0; i < (size_t)V.H * V.W; ++i) {
        const uint8_t* p = L.in + i * 3;
        ++h[p[0]];
        ++h[256 + p[1]];
        ++h[512 + p[2]];
        ls += (unsigned)pil_luma(p[0], p[1], p[2]);
      }
      luma[L.hist_slot] += ls;
    }
    ++launches;
    return 0;
  }
  int lut(const DevPlan& P, const LutJob* jobs, int n, const unsigned* hist, const unsigned long long* luma,
          uint8_t* luts) {
    for (int k = 0; k < n; ++k) {
      const LutJob& J = jobs[k];
      const oadg_op_t& op = P.ops[J.op];
      uint8_t* out = luts + (size_t)op.lut * 768;
      if (op.kind == OADG_OP_AUTOCONTRAST || op.kind == OADG_OP_EQUALIZE) {
        for (int c = 0; c < 3; ++c) {
          const unsigned* h = hist + (size_t)J.hist_slot * 768 + c * 256;
          if (op.kind == OADG_OP_AUTOCONTRAST) lut_autocontrast_ch(h, out + c * 256);
          else lut_equalize_ch(h, out + c * 256);
        }
      } else {
        const oadg_view_t& V = P.views[J.view];
        const double ls = J.hist_slot >= 0 ? (double)luma[J.hist_slot] : 0.0;
        for (int i = 0; i < 256; ++i) {
          uint8_t v = lut_simple_at(op, i, ls, (double)((long long)V.H * V.W));
          out[i] = out[256 + i] = out[512 + i] = v;
        }
      }
    }
    ++launches;
    return 0;
  }
  int bbo_pass(const DevPlan& P, const Chain* chains, int n, int j, int, int) {
    for (int k = 0; k < n; ++k) {
      const Chain& C = chains[k];
      if (j >= C.n) continue;
      int r[4];
      bbo_pass_rect(P, C, j, r);
      for (int y = r[1]; y < r[3]; ++y)
        for (int x = r[0]; x < r[2]; ++x) bbo_pixel(P, C, j, x, y);
    }
    ++launches;
    return 0;
  }
  // mirrors step_kernel (stream runs) + step_pixel_kernel (everything else), same work split predicate
  int step(const DevPlan& P, const Lane* lanes, int n, const int32_t* px_ids, int n_px, const uint8_t* scratch,
           size_t frame_bytes) {
    for (int k = 0; k < n; ++k) {
      const Lane& L = lanes[k];
      uint8_t luts[OADG_MAX_REGIONS * 768];
      for (int r = 0; r <= L.n_ml; ++r)
        if (L.lut[r] >= 0) memcpy(luts + r * 768, P.luts + (size_t)L.lut[r] * 768, 768);
      for (int y = 0; y < L.H; ++y)
        for (int x = 0; x < L.W; x += kChunkPx) {
          const int nn = imin(kChunkPx, L.W - x);
          int region;
          if (!run_is_stream(L, x, y, nn, region)) continue;
          if (region >= 0 && kind_streams(L.kind[region])) {
            Chunk in;
            chunk_load(stream_src(L, region, scratch, frame_bytes) + ((size_t)y * L.W + x) * 3, nn, true, in);
            stream_chunk(L, region, luts + region * 768, scratch, frame_bytes, in, x, y, nn, true);
          } else {
            for (int i = 0; i < nn; ++i) stream_pixel(L, luts, scratch, frame_bytes, x + i, y);
          }
        }
    }
    for (int k = 0; k < n_px; ++k) {
      const Lane& L = lanes[px_ids[k]];
      for (int y = 0; y < L.H; ++y)
        for (int x = 0; x < L.W; ++x) {
          const int xc = x & ~(kChunkPx - 1);
          int region;
          if (run_is_stream(L, xc, y, imin(kChunkPx, L.W - xc), region)) continue;
          step_pixel(P, L, scratch, frame_bytes, x, y);
        }
    }
    launches += n_px > 0 ? 2 : 1;
    return 0;
  }
  int mix(const DevPlan& P, const MixJob* jobs, int n) {
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
