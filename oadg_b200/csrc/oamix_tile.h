// Tile-level bodies of the OA-Mix step / mix kernels.  A CTA of 256 threads owns a 256 x 32 pixel tile.
// Streaming tiles (one LUT / bbo-copy region covers the tile) move one 16-pixel chunk (48 bytes = three
// 16-byte vectors) per thread; every other tile (region borders, bg-only / invert / colour / sharpness ops)
// is evaluated pixel by pixel with consecutive lanes on consecutive pixels.  Shared with tests/hostsim
// (test infrastructure only).
#pragma once
#include <string.h>

#include "oamix_body.h"

namespace oadg {

constexpr int kChunkPx = 16;  // 16 px * 3 B = 48 B = 3 x uint4: the smallest pixel run that is 16-byte periodic
constexpr int kTileW = 256;   // mix tiles: 16 chunks
constexpr int kTileH = 32;    // 16 x 32 chunks: two per thread
constexpr int kMaxCand = 12;

struct Chunk {
  uint32_t w[12];
};

// On the device a chunk is always a full, 16-byte aligned run (three vector loads / stores); ragged right edges
// and unaligned frames take the per-pixel paths, so that a Chunk never has its address taken (it must stay in
// registers).  The host build (tests/hostsim) copies n pixels.
OADG_HD void chunk_load(const uint8_t* p, int n, bool vec, Chunk& c) {
#ifdef __CUDA_ARCH__
  (void)n;
  (void)vec;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1], d = q[2];  // plain loads: see OADG_LDG
  c.w[0] = a.x; c.w[1] = a.y; c.w[2] = a.z; c.w[3] = a.w;
  c.w[4] = b.x; c.w[5] = b.y; c.w[6] = b.z; c.w[7] = b.w;
  c.w[8] = d.x; c.w[9] = d.y; c.w[10] = d.z; c.w[11] = d.w;
#else
  (void)vec;
  memset(c.w, 0, sizeof(c.w));
  memcpy(c.w, p, (size_t)n * 3);
#endif
}

OADG_HD void chunk_store(uint8_t* p, int n, bool vec, const Chunk& c) {
#ifdef __CUDA_ARCH__
  (void)n;
  (void)vec;
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(c.w[0], c.w[1], c.w[2], c.w[3]);
  q[1] = make_uint4(c.w[4], c.w[5], c.w[6], c.w[7]);
  q[2] = make_uint4(c.w[8], c.w[9], c.w[10], c.w[11]);
#else
  (void)vec;
  memcpy(p, c.w, (size_t)n * 3);
#endif
}

// byte k (0..47) of a chunk; k must be a compile-time constant after unrolling for register residency
OADG_HD int chunk_get(const Chunk& c, int k) { return (int)((c.w[k >> 2] >> ((k & 3) * 8)) & 255u); }

OADG_HD bool rect_hit(const int32_t* s, int x0, int y0, int x1, int y1) {
  return s[0] < x1 && s[2] > x0 && s[1] < y1 && s[3] > y0;
}

// Region that covers the whole rectangle [x0,x1) x [y0,y1), or -1 when a multi-level box edge crosses it.
OADG_HD int tile_region(const Lane& L, int x0, int y0, int x1, int y1) {
  int region = L.n_ml;
  for (int b = 0; b < L.n_ml; ++b) {
    const int32_t* B = L.box[b];
    if (!rect_hit(B, x0, y0, x1, y1)) continue;
    if (B[0] <= x0 && B[2] >= x1 && B[1] <= y0 && B[3] >= y1) region = b;
    else return -1;
  }
  return region;
}
OADG_HD bool kind_streams(int kind) { return is_lut_kind(kind) || kind == OADG_OP_BBO_AFFINE; }
// Work split of one depth step: the 16-pixel run [x, x+n) of row y belongs to the STREAM kernel when one
// region covers it and that region's op is a table lookup / bbo-result copy (vector path), or when the lane has
// no other kind of op at all (box-edge runs are then done per pixel by the stream kernel itself).  Every other
// run belongs to the PIXEL kernel.  Returns the covering region (or -1) through `region`.
OADG_HD bool run_is_stream(const Lane& L, int x, int y, int n, int& region) {
  region = tile_region(L, x, y, x + n, y + 1);
  return (region >= 0 && kind_streams(L.kind[region])) || L.all_streaming;
}
OADG_HD int region_of_pixel(const Lane& L, int x, int y) {
  int r = L.n_ml;
  for (int b = 0; b < L.n_ml; ++b)
    if (x >= L.box[b][0] && x < L.box[b][2] && y >= L.box[b][1] && y < L.box[b][3]) r = b;
  return r;
}
// one pixel of a table-lookup / bbo-copy op (box-edge runs of all-streaming lanes); luts: region r at luts + r*768
OADG_HD void stream_pixel(const Lane& L, const uint8_t* luts, const uint8_t* scratch, size_t frame_bytes, int x, int y) {
  const int r = region_of_pixel(L, x, y);
  const size_t o = ((size_t)y * L.W + x) * 3;
  uint8_t* q = L.out + o;
  if (L.kind[r] == OADG_OP_BBO_AFFINE) {
    const uint8_t* s = (L.scratch[r] >= 0 ? scratch + (size_t)L.scratch[r] * frame_bytes : L.in) + o;
    q[0] = (uint8_t)ldb(s);
    q[1] = (uint8_t)ldb(s + 1);
    q[2] = (uint8_t)ldb(s + 2);
  } else {
    const uint8_t* lut = luts + r * 768;
    q[0] = lut[ldb(L.in + o)];
    q[1] = lut[256 + ldb(L.in + o + 1)];
    q[2] = lut[512 + ldb(L.in + o + 2)];
  }
}

// 16 pixels of a streaming tile.  `lut`: the region's 3x256 table (shared memory on the device).
OADG_HD void stream_chunk(const Lane& L, int region, const uint8_t* lut, const uint8_t* scratch, size_t frame_bytes,
                          const Chunk& in, int x, int y, int n, bool vec) {
  const size_t o = ((size_t)y * L.W + x) * 3;
  if (L.kind[region] == OADG_OP_BBO_AFFINE) {  // `in` already holds the bbo result (or the input when no box was valid)
    chunk_store(L.out + o, n, vec, in);
    return;
  }
  Chunk out;
#ifdef __CUDA_ARCH__
  // byte extraction / packing with PRMT, one LDS.U8 per byte; the channel of byte k is k % 3 (compile time)
  const uint8_t* lc[3] = {lut, lut + 256, lut + 512};
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const uint32_t w = in.w[i];
    const uint32_t r0 = lc[(4 * i) % 3][__byte_perm(w, 0, 0x4440)];
    const uint32_t r1 = lc[(4 * i + 1) % 3][__byte_perm(w, 0, 0x4441)];
    const uint32_t r2 = lc[(4 * i + 2) % 3][__byte_perm(w, 0, 0x4442)];
    const uint32_t r3 = lc[(4 * i + 3) % 3][__byte_perm(w, 0, 0x4443)];
    out.w[i] = __byte_perm(__byte_perm(r0, r1, 0x0040), __byte_perm(r2, r3, 0x0040), 0x5410);
  }
#else
  for (int i = 0; i < 12; ++i) {
    uint32_t v = 0;
    for (int b = 0; b < 4; ++b) {
      const int k = i * 4 + b;
      v |= (uint32_t)lut[(k % 3) * 256 + chunk_get(in, k)] << (8 * b);
    }
    out.w[i] = v;
  }
#endif
  chunk_store(L.out + o, n, vec, out);
}
// source frame of a streaming chunk: the lane input, or the bbo scratch frame
OADG_HD const uint8_t* stream_src(const Lane& L, int region, const uint8_t* scratch, size_t frame_bytes) {
  if (L.kind[region] == OADG_OP_BBO_AFFINE && L.scratch[region] >= 0)
    return scratch + (size_t)L.scratch[region] * frame_bytes;
  return L.in;
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

// fused Normalize + Pad + CHW epilogue (oadg_fused_out_t): the n pixels of a chunk as float32 into the 3 planes
OADG_HD void emit_f32_chunk(const DevPlan& P, const MixJob& J, const Chunk& px, float* dst, int x, int y, int n) {
  const size_t plane = (size_t)J.Hp * J.Wp, at = (size_t)y * J.Wp + x;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int k = 0; k < 3; ++k) {
    const int c = P.norm_rgb ? 2 - k : k;
    const float* lut = P.norm_lut + k * 256;
    float* row = dst + k * plane + at;
#ifdef __CUDA_ARCH__
    if (n == kChunkPx && ((J.Wp | x) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (plane & 3) == 0) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        reinterpret_cast<float4*>(row)[g] =
            make_float4(lut[chunk_get(px, (4 * g) * 3 + c)], lut[chunk_get(px, (4 * g + 1) * 3 + c)],
                        lut[chunk_get(px, (4 * g + 2) * 3 + c)], lut[chunk_get(px, (4 * g + 3) * 3 + c)]);
      continue;
    }
#pragma unroll
    for (int i = 0; i < kChunkPx; ++i)   // unrolled: the chunk must stay in registers (no dynamic byte index)
      if (i < n) row[i] = lut[chunk_get(px, i * 3 + c)];
#else
    for (int i = 0; i < n; ++i) row[i] = lut[chunk_get(px, i * 3 + c)];
#endif
  }
}

// 16 pixels of branch mixing + object-aware mixing, 4 pixels (12 bytes = 3 words) at a time
OADG_HD void mix_chunk(const DevPlan& P, const MixJob& J, const MixTile& T, int x, int y, int n, bool vec) {
  const oadg_view_t& V = P.views[J.view];
  const size_t o = ((size_t)y * V.W + x) * 3;
  bool slow = T.overflow || V.width > 4;
#ifdef __CUDA_ARCH__
  slow = slow || !vec || n != kChunkPx;
#endif
  if (slow) {
    for (int i = 0; i < n; ++i) mix_pixel(P, J, x + i, y);
    return;
  }
  Chunk src, br[4], out;
  chunk_load(J.src + o, n, vec, src);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int b = 0; b < 4; ++b)
    if (b < V.width) chunk_load(J.branch[b] + o, n, vec, br[b]);
  const float w0 = V.ws[0], w1 = V.ws[1], w2 = V.ws[2], w3 = V.ws[3];
  const int width = V.width;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int g = 0; g < 4; ++g) {     // pixels 4g .. 4g+3 = bytes 12g .. 12g+11 = words 3g .. 3g+2
    float acc[12];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 12; ++k) {
      const int kk = g * 12 + k;
      float a = fadd(0.f, fmul(w0, u8_to_f32(chunk_get(br[0], kk))));
      if (width > 1) a = fadd(a, fmul(w1, u8_to_f32(chunk_get(br[1], kk))));
      if (width > 2) a = fadd(a, fmul(w2, u8_to_f32(chunk_get(br[2], kk))));
      if (width > 3) a = fadd(a, fmul(w3, u8_to_f32(chunk_get(br[3], kk))));
      acc[k] = a;
    }
    uint32_t ow[3] = {0u, 0u, 0u};
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < 4; ++i) {
      const int px = g * 4 + i;
      float orig[3] = {0.f, 0.f, 0.f}, aug[3] = {0.f, 0.f, 0.f};
      MixMask ms = {0.f, 0.f};
      const int img[3] = {chunk_get(src, px * 3), chunk_get(src, px * 3 + 1), chunk_get(src, px * 3 + 2)};
      if (px < n) {
        for (int t = 0; t < T.n; ++t) {
          const oadg_target_t& G = P.tgts[T.idx[t]];
          float mask;
          if (G.kind == 0) mask = fg_mask(P, G.gt, x + px, y);
          else mask = (x + px >= G.box[0] && x + px < G.box[2] && y >= G.box[1] && y < G.box[3]) ? 1.f : 0.f;
          if (mask == 0.f) continue;
          const float w = mix_target_weight(ms, mask);
          for (int c = 0; c < 3; ++c) mix_accumulate(orig[c], aug[c], G.m_oa, img[c], acc[i * 3 + c], w);
        }
      }
      for (int c = 0; c < 3; ++c) {
        const int k = i * 3 + c;
        ow[k >> 2] |= (uint32_t)mix_finish(orig[c], aug[c], V.m, img[c], acc[k], ms.sum) << ((k & 3) * 8);
      }
    }
    out.w[g * 3] = ow[0];
    out.w[g * 3 + 1] = ow[1];
    out.w[g * 3 + 2] = ow[2];
  }
  chunk_store(J.out + o, n, vec, out);
  if (J.f32_out) emit_f32_chunk(P, J, out, J.f32_out, x, y, n);
  if (J.f32_src) emit_f32_chunk(P, J, src, J.f32_src, x, y, n);
}

}  // namespace oadg
