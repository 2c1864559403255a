// Tile-level bodies of the OA-Mix step / mix kernels: a CTA owns a kTileW x kTileH pixel tile, classifies it
// once (which region/op covers it, which gt masks can be non-zero there) and then streams 16-pixel chunks
// (48 bytes = three 16-byte vectors) per thread.  Shared with tests/hostsim (test infrastructure only).
#pragma once
#include <string.h>

#include "oamix_body.h"

namespace oadg {

constexpr int kChunkPx = 16;  // 16 px * 3 B = 48 B = 3 x uint4: the smallest pixel run that is 16-byte periodic
constexpr int kTileW = 512;   // 32 lanes x 16 px
constexpr int kTileH = 32;    // 8 warps x 4 rows
constexpr int kMaxCand = 12;

struct Chunk {
  uint32_t w[12];
};

OADG_HD void chunk_load(const uint8_t* p, int n, bool vec, Chunk& c) {
#ifdef __CUDA_ARCH__
  if (vec && n == kChunkPx) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1), d = __ldg(q + 2);
    c.w[0] = a.x; c.w[1] = a.y; c.w[2] = a.z; c.w[3] = a.w;
    c.w[4] = b.x; c.w[5] = b.y; c.w[6] = b.z; c.w[7] = b.w;
    c.w[8] = d.x; c.w[9] = d.y; c.w[10] = d.z; c.w[11] = d.w;
    return;
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (i * 4 + b < n * 3) v |= (uint32_t)__ldg(p + i * 4 + b) << (8 * b);
    c.w[i] = v;
  }
#else
  (void)vec;
  memset(c.w, 0, sizeof(c.w));
  memcpy(c.w, p, (size_t)n * 3);
#endif
}

OADG_HD void chunk_store(uint8_t* p, int n, bool vec, const Chunk& c) {
#ifdef __CUDA_ARCH__
  if (vec && n == kChunkPx) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(c.w[0], c.w[1], c.w[2], c.w[3]);
    q[1] = make_uint4(c.w[4], c.w[5], c.w[6], c.w[7]);
    q[2] = make_uint4(c.w[8], c.w[9], c.w[10], c.w[11]);
    return;
  }
#pragma unroll
  for (int i = 0; i < 12; ++i)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (i * 4 + b < n * 3) p[i * 4 + b] = (uint8_t)(c.w[i] >> (8 * b));
#else
  (void)vec;
  memcpy(p, c.w, (size_t)n * 3);
#endif
}

// byte k (0..47) of a chunk; k must be a compile-time constant after unrolling for register residency
OADG_HD int chunk_get(const Chunk& c, int k) { return (int)((c.w[k >> 2] >> ((k & 3) * 8)) & 255u); }

struct RegionInfo {            // one region (multi-level box or the outside) as seen from a tile
  int32_t present;             // the region intersects the tile
  int32_t op;                  // global op index of the region for this lane step
};
struct TileInfo {
  int32_t uniform;             // region id covering the whole tile, or -1 when several regions meet in it
  int32_t any_bg;              // some present region runs a bg-only op: the tile uses the lane-per-pixel mapping
  RegionInfo R[OADG_MAX_REGIONS];
};

OADG_HD bool rect_hit(const int32_t* s, int x0, int y0, int x1, int y1) {
  return s[0] < x1 && s[2] > x0 && s[1] < y1 && s[3] > y0;
}

// classify the tile [x0,x1) x [y0,y1) of lane L
OADG_HD void classify_step_tile(const DevPlan& P, const Lane& L, int x0, int y0, int x1, int y1, TileInfo& T) {
  const oadg_view_t& V = P.views[L.view];
  T.uniform = V.n_ml;
  T.any_bg = 0;
  for (int r = 0; r < OADG_MAX_REGIONS; ++r) T.R[r].present = 0;
  T.R[V.n_ml].present = 1;
  for (int b = 0; b < V.n_ml; ++b) {
    const int32_t* B = V.ml_box[b];
    if (!rect_hit(B, x0, y0, x1, y1)) continue;
    T.R[b].present = 1;
    if (B[0] <= x0 && B[2] >= x1 && B[1] <= y0 && B[3] >= y1) {  // tile inside box b
      T.uniform = b;
      T.R[V.n_ml].present = 0;
    } else {
      T.uniform = -1;
    }
  }
  for (int r = 0; r <= V.n_ml; ++r) {
    if (!T.R[r].present) continue;
    T.R[r].op = L.op_base + r;
    if (P.ops[T.R[r].op].kind == OADG_OP_BG_AFFINE) T.any_bg = 1;
  }
}

OADG_HD int region_of_pixel(const oadg_view_t& V, int x, int y) {
  int r = V.n_ml;
  for (int b = 0; b < V.n_ml; ++b)
    if (x >= V.ml_box[b][0] && x < V.ml_box[b][2] && y >= V.ml_box[b][1] && y < V.ml_box[b][3]) r = b;
  return r;
}
// region covering the whole pixel run [x, x+n) of row y, or -1 when a box edge falls inside it
OADG_HD int region_of_run(const oadg_view_t& V, int x, int y, int n) {
  int r = V.n_ml;
  for (int b = 0; b < V.n_ml; ++b) {
    const int32_t* B = V.ml_box[b];
    if (y < B[1] || y >= B[3] || B[0] >= x + n || B[2] <= x) continue;
    if (B[0] <= x && B[2] >= x + n) r = b;
    else return -1;
  }
  return r;
}

// 16 pixels of one depth step.  `luts` holds the 3x256 table of region r at luts + r*768 when that region's op
// is a LUT op (shared memory on the device).
OADG_HD void step_chunk(const DevPlan& P, const Lane& L, const TileInfo& T, const uint8_t* luts,
                        const uint8_t* scratch, size_t frame_bytes, int x, int y, int n, bool vec) {
  const oadg_view_t& V = P.views[L.view];
  const size_t o = ((size_t)y * V.W + x) * 3;
  const int r = T.uniform >= 0 ? T.uniform : region_of_run(V, x, y, n);
  if (r >= 0) {
    const oadg_op_t& op = P.ops[T.R[r].op];
    const int kind = op.kind;
    Chunk out;
    if (is_lut_kind(kind)) {
      const uint8_t* lut = luts + r * 768;
      Chunk in;
      chunk_load(L.in + o, n, vec, in);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int i = 0; i < 12; ++i) {
        uint32_t v = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int b = 0; b < 4; ++b) {
          const int k = i * 4 + b;
          v |= (uint32_t)lut[(k % 3) * 256 + chunk_get(in, k)] << (8 * b);
        }
        out.w[i] = v;
      }
      chunk_store(L.out + o, n, vec, out);
      return;
    }
    if (kind == OADG_OP_BBO_AFFINE) {
      const uint8_t* s = op.scratch >= 0 ? scratch + (size_t)op.scratch * frame_bytes : L.in;
      chunk_load(s + o, n, vec, out);
      chunk_store(L.out + o, n, vec, out);
      return;
    }
  }
  // a box edge inside the run, or invert / color / sharpness: per pixel, byte stores
  for (int i = 0; i < n; ++i) step_pixel(P, L, scratch, frame_bytes, x + i, y);
}

// ---- mix ------------------------------------------------------------------------------------------
struct MixTile {
  int32_t n;          // object-aware targets that can be non-zero in the tile, in plan order
  int32_t overflow;   // more than kMaxCand: walk every target
  int32_t idx[kMaxCand];
};

OADG_HD void classify_mix_tile(const DevPlan& P, const MixJob& J, int x0, int y0, int x1, int y1, MixTile& T) {
  const oadg_view_t& V = P.views[J.view];
  T.n = 0;
  T.overflow = 0;
  for (int t = 0; t < V.n_tgt; ++t) {
    const oadg_target_t& G = P.tgts[V.tgt_first + t];
    const int32_t* s = G.kind == 0 ? P.gts[G.gt].supp : G.box;
    if (!rect_hit(s, x0, y0, x1, y1)) continue;
    if (T.n < kMaxCand) T.idx[T.n++] = V.tgt_first + t;
    else T.overflow = 1;
  }
}

OADG_HD void mix_chunk(const DevPlan& P, const MixJob& J, const MixTile& T, int x, int y, int n, bool vec) {
  const oadg_view_t& V = P.views[J.view];
  const size_t o = ((size_t)y * V.W + x) * 3;
  if (T.overflow) {
    for (int i = 0; i < n; ++i) mix_pixel(P, J, x + i, y);
    return;
  }
  Chunk src, out;
  chunk_load(J.src + o, n, vec, src);
  float acc[kChunkPx * 3];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int k = 0; k < kChunkPx * 3; ++k) acc[k] = 0.f;
  for (int b = 0; b < V.width; ++b) {
    Chunk br;
    chunk_load(J.branch[b] + o, n, vec, br);
    const float wgt = V.ws[b];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < kChunkPx * 3; ++k) acc[k] = fadd(acc[k], fmul(wgt, (float)chunk_get(br, k)));
  }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 0; i < kChunkPx; ++i) {
    float orig[3] = {0.f, 0.f, 0.f}, aug[3] = {0.f, 0.f, 0.f};
    MixMask ms = {0.f, 0.f};
    const int img[3] = {chunk_get(src, i * 3), chunk_get(src, i * 3 + 1), chunk_get(src, i * 3 + 2)};
    if (i < n) {
      for (int t = 0; t < T.n; ++t) {
        const oadg_target_t& G = P.tgts[T.idx[t]];
        float mask;
        if (G.kind == 0) mask = fg_mask(P, G.gt, x + i, y);
        else mask = (x + i >= G.box[0] && x + i < G.box[2] && y >= G.box[1] && y < G.box[3]) ? 1.f : 0.f;
        if (mask == 0.f) continue;
        const float w = mix_target_weight(ms, mask);
        for (int c = 0; c < 3; ++c) mix_accumulate(orig[c], aug[c], G.m_oa, img[c], acc[i * 3 + c], w);
      }
    }
    for (int c = 0; c < 3; ++c) {
      const int k = i * 3 + c;
      const uint32_t v = (uint32_t)mix_finish(orig[c], aug[c], V.m, img[c], acc[k], ms.sum);
      if ((k & 3) == 0) out.w[k >> 2] = v;
      else out.w[k >> 2] |= v << ((k & 3) * 8);
    }
  }
  chunk_store(J.out + o, n, vec, out);
}

}  // namespace oadg
