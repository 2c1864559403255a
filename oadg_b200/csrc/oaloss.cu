// OA-Loss: fused L2-normalise + all-pairs cosine similarity + masked InfoNCE.
//
// Replaces the reference torch op chain
//   ContrastiveLossPlus.forward   contrastive_loss_plus.py:31-50
//   supcontrast                   contrastive_loss.py:170-232   (mask construction)
//   supcontrast_mask              contrastive_loss.py:147-167   (normalise, matmul/T, masked log-softmax)
// in closed form: no N x N tensor ever reaches HBM.
//   P_ij = [y_i = y_j != bg, i != j]  or  [y_i = y_j = bg, j = pair(i)],  bg = max(labels)
//   z = f f^T / T,  lse_i = log sum_{k != i} exp(z_ik),  n_i = sum_j P_ij
//   loss = -(w/N) sum_{i: n_i > 0} ( (1/n_i) sum_j P_ij z_ij - lse_i )      if #fg > min_samples else 0
//   dL/dz_ij = c_i (P_ij - n_i softmax_ij), c_i = -(w/N)/n_i ;  dL/df = (G + G^T) f / T
//
// This file is the CUDA-core (FFMA, fp32) implementation; the similarity contraction
// also has a tcgen05 path (oaloss_tc.cu) selected at run time.
#include <stdlib.h>

#include "oadg_common.cuh"
#include "oaloss.h"
#include "oadg_peer.cuh"

namespace oadg {
namespace {

constexpr int kC = 256;  // embedding width (out_dim_cont, ..._oadg.py:34); other widths go through the generic loop

// ---- F.normalize twice (contrastive_loss_plus.py:41, contrastive_loss.py:155): warp per row
__global__ void __launch_bounds__(256)
normalize_kernel(const float* __restrict__ x, int n, int c, int normalized_input, float* __restrict__ fhat,
                 float* __restrict__ inv1, float* __restrict__ inv2) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = x + (size_t)row * c;
  float s = 0.f;
  for (int k = lane; k < c; k += 32) {
    float v = xr[k];
    s += v * v;
  }
  s = warp_sum(s);
  float i1 = normalized_input ? 1.f / fmaxf(sqrtf(s), 1e-12f) : 1.f;
  float s2 = 0.f;
  for (int k = lane; k < c; k += 32) {
    float v = xr[k] * i1;
    s2 += v * v;
  }
  s2 = warp_sum(s2);
  float i2 = 1.f / fmaxf(sqrtf(s2), 1e-12f);
  for (int k = lane; k < c; k += 32) fhat[(size_t)row * c + k] = (xr[k] * i1) * i2;
  if (lane == 0) {
    inv1[row] = i1;
    inv2[row] = i2;
  }
}

// ---- label prep: bg = max(labels), #fg, n_i per row.  Single block (N is a few thousand).
// n_i of a foreground row = (#rows with the same label) - 1: counted with a shared-memory histogram when the
// label range is small (class ids), by direct comparison otherwise.
constexpr int kLabelBins = 4096;
__global__ void __launch_bounds__(1024)
label_prep_kernel(const int64_t* __restrict__ labels, const int32_t* __restrict__ pair, int n, int min_samples,
                  int* __restrict__ meta, float* __restrict__ npos) {
  __shared__ long long smax[32], smin[32];
  __shared__ int scnt[32];
  __shared__ long long bg_s, lo_s;
  __shared__ int bins[kLabelBins];
  const int tid = threadIdx.x;
  long long m = LLONG_MIN, mn = LLONG_MAX;
  for (int i = tid; i < n; i += blockDim.x) {
    const long long v = labels[i];
    m = v > m ? v : m;
    mn = v < mn ? v : mn;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    long long t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
    t = __shfl_xor_sync(0xffffffffu, mn, o);
    mn = t < mn ? t : mn;
  }
  if ((tid & 31) == 0) {
    smax[tid >> 5] = m;
    smin[tid >> 5] = mn;
  }
  for (int i = tid; i < kLabelBins; i += blockDim.x) bins[i] = 0;
  __syncthreads();
  if (tid == 0) {
    long long t = smax[0], u = smin[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      t = smax[w] > t ? smax[w] : t;
      u = smin[w] < u ? smin[w] : u;
    }
    bg_s = t;
    lo_s = u;
  }
  __syncthreads();
  const long long bg = bg_s, lo = lo_s;
  const bool small_range = (bg - lo) < (long long)kLabelBins;
  if (small_range) {
    for (int i = tid; i < n; i += blockDim.x) atomicAdd(&bins[(int)(labels[i] - lo)], 1);
    __syncthreads();
  }
  int cnt = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    const long long yi = labels[i];
    float np = 0.f;
    if (yi != bg) {
      ++cnt;
      int same = 0;
      if (small_range) same = bins[(int)(yi - lo)];
      else
        for (int j = 0; j < n; ++j) same += (labels[j] == yi);
      np = (float)(same - 1);
    } else {
      int pj = pair[i];
      np = (pj >= 0 && pj < n && pj != i && labels[pj] == bg) ? 1.f : 0.f;
    }
    npos[i] = np;
  }
  cnt = (int)warp_sum((float)cnt);
  if ((tid & 31) == 0) scnt[tid >> 5] = cnt;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scnt[w];
    meta[0] = (int)bg;
    meta[1] = t;
    meta[2] = t > min_samples ? 1 : 0;  // contrastive_loss.py:211
    meta[4] = 0;                        // ticket of the row-reduce blocks
  }
}

// ---- similarity tile: 64 x 64 outputs, K = c, 256 threads, 4x4 per thread -------------
constexpr int kTM = 64, kTN = 64, kTK = 32;

struct TileAcc {
  float v[4][4];
};

// computes acc = F[i0:i0+64] . F[j0:j0+64]^T over K = c; As/Bs: [kTK][kTM+4]
__device__ __forceinline__ void sim_tile(const float* __restrict__ f, int n, int c, int i0, int j0, float (*As)[kTM + 4],
                                         float (*Bs)[kTN + 4], TileAcc& acc) {
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc.v[a][b] = 0.f;
  for (int k0 = 0; k0 < c; k0 += kTK) {
    // 64 rows x 32 k per operand = 2048 floats = 512 float4; 256 threads x 2
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      int idx = tid + it * 256;
      int r = idx >> 3, kq = (idx & 7) * 4;
      float4 va = make_float4(0, 0, 0, 0), vb = make_float4(0, 0, 0, 0);
      if (i0 + r < n) va = *reinterpret_cast<const float4*>(f + (size_t)(i0 + r) * c + k0 + kq);
      if (j0 + r < n) vb = *reinterpret_cast<const float4*>(f + (size_t)(j0 + r) * c + k0 + kq);
      As[kq + 0][r] = va.x; As[kq + 1][r] = va.y; As[kq + 2][r] = va.z; As[kq + 3][r] = va.w;
      Bs[kq + 0][r] = vb.x; Bs[kq + 1][r] = vb.y; Bs[kq + 2][r] = vb.z; Bs[kq + 3][r] = vb.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc.v[p][q] = fmaf(av[p], bv[q], acc.v[p][q]);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ bool is_pos(long long yi, long long yj, long long bg, int i, int j, int pair_i) {
  if (yi != yj || i == j) return false;
  return yi != bg ? true : (j == pair_i);
}

// forward: per (row tile, col tile) partial row statistics
__global__ void __launch_bounds__(256)
sim_fwd_kernel(const float* __restrict__ f, const int64_t* __restrict__ labels, const int32_t* __restrict__ pair,
               const int* __restrict__ meta, int n, int c, float inv_t, float* __restrict__ partial) {
  if (!meta[2]) return;
  __shared__ __align__(16) float As[kTK][kTM + 4];
  __shared__ __align__(16) float Bs[kTK][kTN + 4];
  const int i0 = blockIdx.y * kTM, j0 = blockIdx.x * kTN;
  TileAcc acc;
  sim_tile(f, n, c, i0, j0, As, Bs, acc);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const long long bg = (long long)meta[0];
  long long yj[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) yj[q] = (j0 + tx * 4 + q) < n ? labels[j0 + tx * 4 + q] : 0;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int i = i0 + ty * 4 + p;
    const bool row_ok = i < n;
    const long long yi = row_ok ? labels[i] : 0;
    const int pi = row_ok ? pair[i] : -1;
    float m = -INFINITY, ps = 0.f;
    float z[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + tx * 4 + q;
      z[q] = acc.v[p][q] * inv_t;
      if (row_ok && j < n) {
        m = fmaxf(m, z[q]);
        if (is_pos(yi, yj[q], bg, i, j, pi)) ps += z[q];
      }
    }
    // reduce over the 16 threads (tx) that share this row
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + tx * 4 + q;
      if (row_ok && j < n && j != i) s += expf(z[q] - m);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ps += __shfl_xor_sync(0xffffffffu, ps, o);
    }
    if (tx == 0 && row_ok) {
      float* out = partial + ((size_t)i * gridDim.x + blockIdx.x) * 3;   // [row][column tile][3]
      out[0] = m;
      out[1] = s;
      out[2] = ps;
    }
  }
}

// combine the column-tile partials, emit per-row stats and the scalar loss (single block,
// fixed summation order => deterministic)
constexpr int kRowReduceThreads = 256;
// one warp per row, rows strided over at most 1024 blocks (`red` holds one partial sum per block)
inline int row_reduce_blocks(int n) {
  const int b = (n + kRowReduceThreads / 32 - 1) / (kRowReduceThreads / 32);
  return b < 1 ? 1 : (b > 1024 ? 1024 : b);
}
__global__ void __launch_bounds__(kRowReduceThreads)
row_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ npos, int* __restrict__ meta,
                  int n, int row0, int n_total, int col_tiles, float loss_weight, RowStats* __restrict__ stats,
                  double* __restrict__ red, float* __restrict__ loss, const float* __restrict__ z, int ld,
                  const int64_t* __restrict__ labels, const int32_t* __restrict__ pair) {
  // rows [row0, row0 + n) of an n_total-row problem (single GPU: row0 = 0, n = n_total); partial / stats are
  // indexed by the local row, npos by the global row; the loss is this range's share of the mean over n_total.
  __shared__ double sred[kRowReduceThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kWarps = kRowReduceThreads / 32;
  if (!meta[2]) {
    for (int i = blockIdx.x * kRowReduceThreads + tid; i < n; i += gridDim.x * kRowReduceThreads)
      stats[i] = RowStats{0.f, 0.f, 0.f, 0.f};
    if (blockIdx.x == 0 && tid == 0) *loss = 0.f;
    return;
  }
  // one warp per row (lanes stride over the column-tile partials), rows strided over the grid; every sum has a
  // fixed order: lanes, then warps, then blocks
  double local = 0.0;
  for (int i = blockIdx.x * kWarps + warp; i < n; i += gridDim.x * kWarps) {
    float M = -INFINITY;
    const float* prow = partial + (size_t)i * col_tiles * 3;   // [row][column tile][3]: contiguous per row
    for (int t = lane; t < col_tiles; t += 32) M = fmaxf(M, prow[t * 3]);
    M = warp_max(M);
    float S = 0.f, Ps = 0.f;
    for (int t = lane; t < col_tiles; t += 32) {
      const float* p = prow + t * 3;
      S += p[1] * expf(p[0] - M);
      Ps += p[2];
    }
    S = warp_sum(S);
    Ps = warp_sum(Ps);
    if (lane == 0) {
      const float lse = M + logf(S);
      const float np = npos[row0 + i];
      // tcgen05 forward: a background row's only positive is its other view (contrastive_loss.py:216-221); the
      // similarity kernel leaves it out and the stored logit is read here (np > 0 <=> the pair is a valid bg row)
      if (z && np > 0.f && labels[row0 + i] == (int64_t)meta[0]) Ps = z[(size_t)i * ld + pair[row0 + i]];
      RowStats st;
      st.lse = lse;
      st.npos = np;
      st.coef = np > 0.f ? -(loss_weight / (float)n_total) / np : 0.f;
      st.u = st.coef * np * expf(-lse);
      stats[i] = st;
      if (np > 0.f) local += (double)(Ps / np - lse);
    }
  }
  __shared__ int last_s;
  if (lane == 0) sred[warp] = local;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kWarps; ++w) t += sred[w];
    red[blockIdx.x] = t;
    __threadfence();
    last_s = atomicAdd(&meta[4], 1) == (int)gridDim.x - 1;   // the last block: every partial sum is visible
  }
  __syncthreads();
  if (last_s) {   // block partials added in a fixed order by the whole block: thread, then lanes, then warps
    __threadfence();
    double t = 0.0;
    for (unsigned b2 = tid; b2 < gridDim.x; b2 += kRowReduceThreads) t += *((volatile double*)red + b2);
    t = warp_sum(t);
    __syncthreads();
    if (lane == 0) sred[warp] = t;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < kWarps; ++w) tot += sred[w];
      *loss = (float)(-(double)loss_weight * tot / (double)n_total);
      meta[4] = 0;
    }
  }
}

// backward: block (row tile I, column split s): for its column tiles J
//   A_ij = G_ij + G_ji ;  dF_I += A . F_J / T       (atomicAdd into dfhat)
constexpr int kBwdThreads = 256;
__global__ void __launch_bounds__(kBwdThreads)
sim_bwd_kernel(const float* __restrict__ f, const int64_t* __restrict__ labels, const int32_t* __restrict__ pair,
               const int* __restrict__ meta, const RowStats* __restrict__ stats, int n, int c, float inv_t,
               int col_tiles, float* __restrict__ dfhat) {
  if (!meta[2]) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float(*As)[kTM + 4] = reinterpret_cast<float(*)[kTM + 4]>(smem_raw);
  float(*Bs)[kTN + 4] = As + kTK;
  float(*At)[kTN + 1] = reinterpret_cast<float(*)[kTN + 1]>(Bs + kTK);  // [64 rows][64 cols] A tile
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int i0 = blockIdx.y * kTM;
  const long long bg = (long long)meta[0];
  // each thread owns dF rows (4 rows: r = tid>>6 .. ) x cols: 64 rows x c cols / 256 threads
  // mapping: thread t -> row group rg = t >> 4 (16 groups of 4 rows), col lane cl = t & 15 (cols cl, cl+16, ...)
  float dacc[4][kC / 16];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < kC / 16; ++q) dacc[p][q] = 0.f;
  for (int jt = blockIdx.x; jt < col_tiles; jt += gridDim.x) {
    const int j0 = jt * kTN;
    TileAcc acc;
    sim_tile(f, n, c, i0, j0, As, Bs, acc);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int i = i0 + ty * 4 + p;
      const bool row_ok = i < n;
      const long long yi = row_ok ? labels[i] : 0;
      const int pi = row_ok ? pair[i] : -1;
      const RowStats si = row_ok ? stats[i] : RowStats{0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j0 + tx * 4 + q;
        float a = 0.f;
        if (row_ok && j < n && j != i) {
          const float z = acc.v[p][q] * inv_t;
          const long long yjq = labels[j];
          const RowStats sj = stats[j];
          const float pij = is_pos(yi, yjq, bg, i, j, pi) ? 1.f : 0.f;
          const float pji = is_pos(yjq, yi, bg, j, i, pair[j]) ? 1.f : 0.f;
          a = si.coef * (pij - si.npos * expf(z - si.lse)) + sj.coef * (pji - sj.npos * expf(z - sj.lse));
        }
        At[ty * 4 + p][tx * 4 + q] = a * inv_t;
      }
    }
    __syncthreads();
    // dF_I[r][col] += sum_j At[r][j] * F[j0+j][col]
    for (int j = 0; j < kTN; ++j) {
      if (j0 + j >= n) break;
      const float* fr = f + (size_t)(j0 + j) * c;
      float a[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) a[p] = At[ty * 4 + p][j];
#pragma unroll
      for (int q = 0; q < kC / 16; ++q) {
        const float fv = __ldg(fr + tx + q * 16);
#pragma unroll
        for (int p = 0; p < 4; ++p) dacc[p][q] = fmaf(a[p], fv, dacc[p][q]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int i = i0 + ty * 4 + p;
    if (i >= n) continue;
#pragma unroll
    for (int q = 0; q < kC / 16; ++q) atomicAdd(dfhat + (size_t)i * c + tx + q * 16, dacc[p][q]);
  }
}

// chain rule through the two normalisations (warp per row), scaled by the upstream gradient
__global__ void __launch_bounds__(256)
normalize_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dpart, int n_part,
                     const float* __restrict__ inv1, const float* __restrict__ inv2, const int* __restrict__ meta,
                     const float* __restrict__ gscale, int n, int c, int normalized_input, float* __restrict__ gx) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  float* out = gx + (size_t)row * c;
  if (!meta[2]) {
    for (int k = lane; k < c; k += 32) out[k] = 0.f;
    return;
  }
  const float g0 = *gscale;
  const float* xr = x + (size_t)row * c;
  const float i1 = inv1[row], i2 = inv2[row];
  // dL/dfhat of this row = sum of the column-split partials, in a fixed order (c == 256: 8 values per lane)
  // (partials outer, columns inner: eight independent loads per step, same summation order per column)
  float gr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int p = 0; p < n_part; ++p) {
    const float* dp = dpart + ((size_t)p * n + row) * c + lane;
#pragma unroll
    for (int q = 0; q < 8; ++q) gr[q] += dp[32 * q];
  }
  // second normalisation: v = x*i1, u = v*i2 ; g1 = (g - (u.g) u) * i2
  float dot = 0.f;
  for (int k = lane, q = 0; k < c; k += 32, ++q) dot += (xr[k] * i1 * i2) * gr[q];
  dot = warp_sum(dot);
  if (!normalized_input) {
    for (int k = lane, q = 0; k < c; k += 32, ++q) out[k] = g0 * (gr[q] - dot * (xr[k] * i2)) * i2;
    return;
  }
  // first normalisation: u1 = x*i1 ; gx = (g1 - (u1.g1) u1) * i1
  float dot1 = 0.f;
  for (int k = lane, q = 0; k < c; k += 32, ++q) {
    float u = xr[k] * i1 * i2;
    float g1 = (gr[q] - dot * u) * i2;
    dot1 += (xr[k] * i1) * g1;
  }
  dot1 = warp_sum(dot1);
  for (int k = lane, q = 0; k < c; k += 32, ++q) {
    float u = xr[k] * i1 * i2;
    float g1 = (gr[q] - dot * u) * i2;
    out[k] = g0 * (g1 - dot1 * (xr[k] * i1)) * i1;
  }
}

}  // namespace

}  // namespace oadg

using namespace oadg;

// ---- cross-rank path, packed buffers: a rank's contribution to the all-gather is ONE buffer of n_rows rows of
// kPackW floats = [fhat (c) | the row's int64 label as two float-sized words | 0 | 0]; the gathered buffer is read in
// place (row stride kPackW), so no pack / slice / contiguous kernels sit around the collectives.
constexpr int kPackPad = 4;   // keeps rows 16-byte aligned
__global__ void __launch_bounds__(256)
normalize_pack_kernel(const float* __restrict__ x, const int64_t* __restrict__ labels, int n_labels, int n, int c,
                      int normalized_input, float* __restrict__ send, float* __restrict__ inv1, float* __restrict__ inv2) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = x + (size_t)row * c;
  float s = 0.f;
  for (int k = lane; k < c; k += 32) {
    float v = xr[k];
    s += v * v;
  }
  s = warp_sum(s);
  float i1 = normalized_input ? 1.f / fmaxf(sqrtf(s), 1e-12f) : 1.f;
  float s2 = 0.f;
  for (int k = lane; k < c; k += 32) {
    float v = xr[k] * i1;
    s2 += v * v;
  }
  s2 = warp_sum(s2);
  float i2 = 1.f / fmaxf(sqrtf(s2), 1e-12f);
  float* out = send + (size_t)row * (c + kPackPad);
  for (int k = lane; k < c; k += 32) out[k] = (xr[k] * i1) * i2;
  if (lane == 0) {
    inv1[row] = i1;
    inv2[row] = i2;
    // rows beyond the label list (random proposals) take the last label (contrastive_loss_plus.py:44-47)
    const long long y = labels[row < n_labels ? row : n_labels - 1];
    out[c] = __uint_as_float((unsigned)((unsigned long long)y & 0xffffffffull));
    out[c + 1] = __uint_as_float((unsigned)((unsigned long long)y >> 32));
    out[c + 2] = 0.f;
    out[c + 3] = 0.f;
  }
}
// The same rows, stored straight into EVERY rank's gather buffer over NVLink (row index rank * n + row of the buffer at
// rows_offset), then this rank's flag word raised everywhere: the all-gather without a collective kernel.
__global__ void __launch_bounds__(256)
normalize_pack_peers_kernel(const float* __restrict__ x, const int64_t* __restrict__ labels, int n_labels, int n, int c,
                            int normalized_input, const oadg_peers_t P, size_t rows_offset, size_t flag_offset,
                            size_t counter_offset, unsigned seq, float* __restrict__ inv1, float* __restrict__ inv2) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row < n) {
    const float* xr = x + (size_t)row * c;
    float s = 0.f;
    for (int k = lane; k < c; k += 32) {
      float v = xr[k];
      s += v * v;
    }
    s = warp_sum(s);
    float i1 = normalized_input ? 1.f / fmaxf(sqrtf(s), 1e-12f) : 1.f;
    float s2 = 0.f;
    for (int k = lane; k < c; k += 32) {
      float v = xr[k] * i1;
      s2 += v * v;
    }
    s2 = warp_sum(s2);
    float i2 = 1.f / fmaxf(sqrtf(s2), 1e-12f);
    const size_t at = ((size_t)P.rank * n + row) * (c + kPackPad);
    for (int k = lane; k < c; k += 32) {
      const float v = (xr[k] * i1) * i2;
      for (int r = 0; r < P.world; ++r) (reinterpret_cast<float*>(static_cast<char*>(P.base[r]) + rows_offset) + at)[k] = v;
    }
    if (lane == 0) {
      inv1[row] = i1;
      inv2[row] = i2;
      const long long y = labels[row < n_labels ? row : n_labels - 1];
      const float4 t = make_float4(__uint_as_float((unsigned)((unsigned long long)y & 0xffffffffull)),
                                   __uint_as_float((unsigned)((unsigned long long)y >> 32)), 0.f, 0.f);
      for (int r = 0; r < P.world; ++r)
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(static_cast<char*>(P.base[r]) + rows_offset) + at + c) = t;
    }
  }
  peer_signal(P, flag_offset, counter_offset, seq, (unsigned)n);
}
__global__ void __launch_bounds__(256)
unpack_labels_kernel(const float* __restrict__ recv, int n_total, int c, int64_t* __restrict__ labels_all) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n_total) return;
  const float* r = recv + (size_t)i * (c + kPackPad) + c;
  labels_all[i] = (int64_t)((unsigned long long)__float_as_uint(r[0]) | ((unsigned long long)__float_as_uint(r[1]) << 32));
}
// gathered tails [world][n_rows + 1][4] (row statistics, then {the rank's loss part, 0, 0, 0}) -> contiguous statistics
// of all rows + the total loss (summed in rank order: every rank gets the same bits)
__global__ void __launch_bounds__(256)
finish_packed_kernel(const float4* __restrict__ tail_all, int world, int n_rows, float4* __restrict__ stats_all,
                     float* __restrict__ loss_out) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j < world * n_rows) stats_all[j] = tail_all[(size_t)(j / n_rows) * (n_rows + 1) + j % n_rows];
  if (j == 0) {
    float t = 0.f;
    for (int r = 0; r < world; ++r) t += tail_all[(size_t)r * (n_rows + 1) + n_rows].x;
    *loss_out = t;
  }
}

// The product computes the similarity on the tensor cores (tcgen05, oaloss_tc.cu).  The CUDA-core FFMA kernels above
// are kept as a TEST BUILD only (-DOADG_LOSS_FFMA, tests/test_gpu_oaloss.py builds it as a second library): an
// independent implementation of the same closed form to cross-check the tcgen05 path; there is no runtime switch.
// The tcgen05 forward shifts every logit by the diagonal 1 / T instead of a running row maximum (rows are unit vectors)
// and evaluates exp with ex2.approx.ftz: below this temperature a row whose other similarities are all low could have
// every term flushed to zero (2 / T * log2(e) > 126).  The reference's configs use 0.06 / 0.07.
static constexpr float kMinTemperatureTc = 0.025f;
static constexpr int loss_tc_enabled() {
#ifdef OADG_LOSS_FFMA
  return 0;
#else
  return 1;
#endif
}

extern "C" int oadg_supcon_workspace_bytes(int n, int c, size_t* out_bytes) {
  if (!out_bytes || n < 0 || c <= 0) return OADG_E_ARG;
  LossWs w = carve_loss_ws(nullptr, n > 0 ? n : 1, c);
  *out_bytes = w.bytes;
  return 0;
}

extern "C" int oadg_supcon_forward(const float* feats_dev, const int64_t* labels_dev, const int32_t* pair_dev, int n,
                                   int c, float temperature, float loss_weight, int min_samples,
                                   int normalized_input, float* loss_dev, void* workspace_dev,
                                   size_t workspace_bytes, int* launches_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!loss_dev || n < 0) return OADG_E_ARG;
  if (n == 0) {
    OADG_CUDA_TRY(cudaMemsetAsync(loss_dev, 0, sizeof(float), stream));
    if (launches_out) *launches_out = 0;
    return 0;
  }
  if (!feats_dev || !labels_dev || !pair_dev || !workspace_dev) return OADG_E_ARG;
  if (c != kC) return OADG_E_LIMIT;
  if (!(temperature > 0.f)) return OADG_E_ARG;
  if (loss_tc_enabled() && temperature < kMinTemperatureTc) return OADG_E_LIMIT;
  if (((uintptr_t)feats_dev & 15) || ((uintptr_t)workspace_dev & 255)) return OADG_E_ARG;
  LossWs w = carve_loss_ws(workspace_dev, n, c);
  if (workspace_bytes < w.bytes) return OADG_E_ARG;
  const int col_tiles = (n + kTN - 1) / kTN, row_tiles = (n + kTM - 1) / kTM;
  int launches = 0;
  normalize_kernel<<<(n + 7) / 8, 256, 0, stream>>>(feats_dev, n, c, normalized_input, w.fhat, w.inv1, w.inv2);
  OADG_LAUNCH_CHECK();
  label_prep_kernel<<<1, 1024, 0, stream>>>(labels_dev, pair_dev, n, min_samples, w.meta, w.npos);
  OADG_LAUNCH_CHECK();
  int red_tiles = col_tiles;
  if (loss_tc_enabled()) {
    int rc = launch_sim_fwd_tc(w, labels_dev, pair_dev, n, 0, n, 1.f / temperature, stream, &launches);
    if (rc) return rc;
    red_tiles = 2 * ((n + 127) / 128);   // the tcgen05 forward writes one partial per 64-column half tile
  } else {
    sim_fwd_kernel<<<dim3(col_tiles, row_tiles), 256, 0, stream>>>(w.fhat, labels_dev, pair_dev, w.meta, n, c,
                                                                   1.f / temperature, w.partial);
    OADG_LAUNCH_CHECK();
    ++launches;
  }
  row_reduce_kernel<<<row_reduce_blocks(n), kRowReduceThreads, 0, stream>>>(
      w.partial, w.npos, w.meta, n, 0, n, red_tiles, loss_weight, w.stats, w.red, loss_dev,
      loss_tc_enabled() ? w.z : nullptr, w.ld, labels_dev, pair_dev);
  OADG_LAUNCH_CHECK();
  launches += 3;
  if (launches_out) *launches_out = launches;
  return 0;
}

extern "C" int oadg_supcon_backward(const float* feats_dev, const int64_t* labels_dev, const int32_t* pair_dev, int n,
                                    int c, float temperature, float loss_weight, int normalized_input,
                                    const float* grad_loss_dev, float* grad_feats_dev, void* workspace_dev,
                                    size_t workspace_bytes, int* launches_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  (void)loss_weight;
  if (n < 0) return OADG_E_ARG;
  if (n == 0) {
    if (launches_out) *launches_out = 0;
    return 0;
  }
  if (!feats_dev || !labels_dev || !pair_dev || !workspace_dev || !grad_loss_dev || !grad_feats_dev) return OADG_E_ARG;
  if (c != kC) return OADG_E_LIMIT;
  LossWs w = carve_loss_ws(workspace_dev, n, c);
  if (workspace_bytes < w.bytes) return OADG_E_ARG;
  const int col_tiles = (n + kTN - 1) / kTN, row_tiles = (n + kTM - 1) / kTM;
  int launches = 0;
  const float* dpart = w.dfhat;
  int n_part = 1;
  if (loss_tc_enabled()) {
    int rc = launch_sim_bwd_tc(w, w.stats, labels_dev, pair_dev, n, 0, n, 1.f / temperature, stream, &launches);
    if (rc) return rc;
    dpart = w.dpart;
    n_part = kBwdSplits;
  } else {
    OADG_CUDA_TRY(cudaMemsetAsync(w.dfhat, 0, (size_t)n * c * sizeof(float), stream));
    int splits = (2 * kNumSMs + row_tiles - 1) / row_tiles;
    if (splits > col_tiles) splits = col_tiles;
    if (splits < 1) splits = 1;
    const size_t smem = (size_t)2 * kTK * (kTM + 4) * 4 + (size_t)kTM * (kTN + 1) * 4;
    static bool attr_of[64] = {false};
    const int slot = device_slot();
    if (!attr_of[slot]) {
      OADG_CUDA_TRY(cudaFuncSetAttribute(sim_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_of[slot] = true;
    }
    sim_bwd_kernel<<<dim3(splits, row_tiles), kBwdThreads, smem, stream>>>(w.fhat, labels_dev, pair_dev, w.meta,
                                                                           w.stats, n, c, 1.f / temperature, col_tiles,
                                                                           w.dfhat);
    OADG_LAUNCH_CHECK();
    ++launches;
  }
  normalize_bwd_kernel<<<(n + 7) / 8, 256, 0, stream>>>(feats_dev, dpart, n_part, w.inv1, w.inv2, w.meta, grad_loss_dev,
                                                       n, c, normalized_input, grad_feats_dev);
  OADG_LAUNCH_CHECK();
  ++launches;
  if (launches_out) *launches_out = launches;
  return 0;
}


// ---- cross-rank variant: anchors = this rank's rows, contrasts = the all-gathered rows -----------------------
extern "C" int oadg_supcon_normalize(const float* feats_dev, int n_rows, int n_total, int c, int normalized_input,
                                     float* fhat_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_rows <= 0 || n_total < n_rows || !feats_dev || !fhat_out_dev || !workspace_dev) return OADG_E_ARG;
  if (c != kC) return OADG_E_LIMIT;
  if ((uintptr_t)workspace_dev & 255) return OADG_E_ARG;
  LossWs w = carve_loss_ws(workspace_dev, n_total, c);
  if (workspace_bytes < w.bytes) return OADG_E_ARG;
  normalize_kernel<<<(n_rows + 7) / 8, 256, 0, stream>>>(feats_dev, n_rows, c, normalized_input, fhat_out_dev, w.inv1,
                                                         w.inv2);
  OADG_LAUNCH_CHECK();
  return 0;
}

extern "C" int oadg_supcon_forward_gathered(const float* fhat_all_dev, const int64_t* labels_all_dev,
                                            const int32_t* pair_all_dev, int n_total, int row0, int n_rows, int c,
                                            float temperature, float loss_weight, int min_samples,
                                            float* loss_part_dev, float* stats_local_dev, void* workspace_dev,
                                            size_t workspace_bytes, int* launches_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!fhat_all_dev || !labels_all_dev || !pair_all_dev || !loss_part_dev || !stats_local_dev || !workspace_dev)
    return OADG_E_ARG;
  if (n_rows <= 0 || row0 < 0 || row0 + n_rows > n_total) return OADG_E_ARG;
  if (c != kC) return OADG_E_LIMIT;
  if (!(temperature > 0.f) || ((uintptr_t)fhat_all_dev & 15) || ((uintptr_t)workspace_dev & 255)) return OADG_E_ARG;
  if (!loss_tc_enabled()) return OADG_E_LIMIT;  // the cross-rank path exists for the tcgen05 kernels only
  if (temperature < kMinTemperatureTc) return OADG_E_LIMIT;
  LossWs w = carve_loss_ws(workspace_dev, n_total, c);
  if (workspace_bytes < w.bytes) return OADG_E_ARG;
  int launches = 0;
  w.fhat = const_cast<float*>(fhat_all_dev);
  label_prep_kernel<<<1, 1024, 0, stream>>>(labels_all_dev, pair_all_dev, n_total, min_samples, w.meta, w.npos);
  OADG_LAUNCH_CHECK();
  int rc = launch_sim_fwd_tc(w, labels_all_dev, pair_all_dev, n_total, row0, n_rows, 1.f / temperature, stream, &launches);
  if (rc) return rc;
  row_reduce_kernel<<<row_reduce_blocks(n_rows), kRowReduceThreads, 0, stream>>>(
      w.partial, w.npos, w.meta, n_rows, row0, n_total, 2 * ((n_total + 127) / 128), loss_weight,
      reinterpret_cast<RowStats*>(stats_local_dev), w.red, loss_part_dev, w.z, w.ld, labels_all_dev, pair_all_dev);
  OADG_LAUNCH_CHECK();
  launches += 2;
  if (launches_out) *launches_out = launches;
  return 0;
}

extern "C" int oadg_supcon_backward_gathered(const float* feats_local_dev, const int64_t* labels_all_dev,
                                             const int32_t* pair_all_dev, const float* stats_all_dev, int n_total,
                                             int row0, int n_rows, int c, float temperature, int normalized_input,
                                             const float* grad_loss_dev, float* grad_feats_dev, void* workspace_dev,
                                             size_t workspace_bytes, int* launches_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!feats_local_dev || !labels_all_dev || !pair_all_dev || !stats_all_dev || !grad_loss_dev || !grad_feats_dev ||
      !workspace_dev)
    return OADG_E_ARG;
  if (n_rows <= 0 || row0 < 0 || row0 + n_rows > n_total) return OADG_E_ARG;
  if (c != kC) return OADG_E_LIMIT;
  if (!loss_tc_enabled()) return OADG_E_LIMIT;
  LossWs w = carve_loss_ws(workspace_dev, n_total, c);
  if (workspace_bytes < w.bytes) return OADG_E_ARG;
  int launches = 0;
  int rc = launch_sim_bwd_tc(w, reinterpret_cast<const RowStats*>(stats_all_dev), labels_all_dev, pair_all_dev, n_total,
                             row0, n_rows, 1.f / temperature, stream, &launches);
  if (rc) return rc;
  normalize_bwd_kernel<<<(n_rows + 7) / 8, 256, 0, stream>>>(feats_local_dev, w.dpart, kBwdSplits, w.inv1, w.inv2, w.meta,
                                                            grad_loss_dev, n_rows, c, normalized_input, grad_feats_dev);
  OADG_LAUNCH_CHECK();
  ++launches;
  if (launches_out) *launches_out = launches;
  return 0;
}

// ---- cross-rank variant with packed buffers (see normalize_pack_kernel) ----------------------------------------------
extern "C" int oadg_supcon_pack_width(int c) { return c + kPackPad; }

extern "C" int oadg_supcon_gather_pack(const float* feats_dev, const int64_t* labels_dev, int n_labels, int n_rows,
                                       int n_total, int c, int normalized_input, float* send_dev, void* workspace_dev,
                                       size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_rows <= 0 || n_total < n_rows || n_labels < 1 || n_labels > n_rows || !feats_dev || !labels_dev || !send_dev ||
      !workspace_dev)
    return OADG_E_ARG;
  if (c != kC) return OADG_E_LIMIT;
  if (((uintptr_t)workspace_dev & 255) || ((uintptr_t)send_dev & 15)) return OADG_E_ARG;
  LossWs w = carve_loss_ws(workspace_dev, n_total, c);
  if (workspace_bytes < w.bytes) return OADG_E_ARG;
  normalize_pack_kernel<<<(n_rows + 7) / 8, 256, 0, stream>>>(feats_dev, labels_dev, n_labels, n_rows, c,
                                                              normalized_input, send_dev, w.inv1, w.inv2);
  OADG_LAUNCH_CHECK();
  return 0;
}

extern "C" int oadg_supcon_gather_pack_peers(const float* feats_dev, const int64_t* labels_dev, int n_labels, int n_rows,
                                             int n_total, int c, int normalized_input, const oadg_peers_t* peers,
                                             size_t rows_offset, size_t flag_offset, size_t counter_offset, uint32_t seq,
                                             void* workspace_dev, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_rows <= 0 || n_labels < 1 || n_labels > n_rows || !feats_dev || !labels_dev || !peers || !workspace_dev)
    return OADG_E_ARG;
  if (peers->world < 1 || peers->world > OADG_PEER_MAX || peers->rank < 0 || peers->rank >= peers->world ||
      n_total != peers->world * n_rows)
    return OADG_E_ARG;
  for (int r = 0; r < peers->world; ++r)
    if (!peers->base[r]) return OADG_E_ARG;
  if (c != kC) return OADG_E_LIMIT;
  if (((uintptr_t)workspace_dev & 255) || (rows_offset & 15) || (flag_offset & 3) || (counter_offset & 3)) return OADG_E_ARG;
  LossWs w = carve_loss_ws(workspace_dev, n_total, c);
  if (workspace_bytes < w.bytes) return OADG_E_ARG;
  normalize_pack_peers_kernel<<<(n_rows + 7) / 8, 256, 0, stream>>>(feats_dev, labels_dev, n_labels, n_rows, c,
                                                                    normalized_input, *peers, rows_offset, flag_offset,
                                                                    counter_offset, seq, w.inv1, w.inv2);
  OADG_LAUNCH_CHECK();
  return 0;
}

extern "C" int oadg_supcon_forward_packed(const float* recv_dev, const int32_t* pair_all_dev, int n_total, int row0,
                                          int n_rows, int c, float temperature, float loss_weight, int min_samples,
                                          float* tail_dev, void* workspace_dev, size_t workspace_bytes,
                                          int* launches_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!recv_dev || !pair_all_dev || !tail_dev || !workspace_dev) return OADG_E_ARG;
  if (n_rows <= 0 || row0 < 0 || row0 + n_rows > n_total) return OADG_E_ARG;
  if (c != kC) return OADG_E_LIMIT;
  if (!(temperature > 0.f) || ((uintptr_t)recv_dev & 15) || ((uintptr_t)tail_dev & 15) || ((uintptr_t)workspace_dev & 255))
    return OADG_E_ARG;
  if (!loss_tc_enabled()) return OADG_E_LIMIT;  // the cross-rank path exists for the tcgen05 kernels only
  if (temperature < kMinTemperatureTc) return OADG_E_LIMIT;
  LossWs w = carve_loss_ws(workspace_dev, n_total, c);
  if (workspace_bytes < w.bytes) return OADG_E_ARG;
  int launches = 0;
  w.fhat = const_cast<float*>(recv_dev);
  w.fhat_ld = c + kPackPad;
  unpack_labels_kernel<<<(n_total + 255) / 256, 256, 0, stream>>>(recv_dev, n_total, c, w.labels_all);
  OADG_LAUNCH_CHECK();
  label_prep_kernel<<<1, 1024, 0, stream>>>(w.labels_all, pair_all_dev, n_total, min_samples, w.meta, w.npos);
  OADG_LAUNCH_CHECK();
  int rc = launch_sim_fwd_tc(w, w.labels_all, pair_all_dev, n_total, row0, n_rows, 1.f / temperature, stream, &launches);
  if (rc) return rc;
  row_reduce_kernel<<<row_reduce_blocks(n_rows), kRowReduceThreads, 0, stream>>>(
      w.partial, w.npos, w.meta, n_rows, row0, n_total, 2 * ((n_total + 127) / 128), loss_weight,
      reinterpret_cast<RowStats*>(tail_dev), w.red, tail_dev + (size_t)n_rows * 4, w.z, w.ld, w.labels_all, pair_all_dev);
  OADG_LAUNCH_CHECK();
  launches += 3;
  if (launches_out) *launches_out = launches;
  return 0;
}

extern "C" int oadg_supcon_finish_packed(const float* tail_all_dev, int world, int n_rows, int c, float* loss_dev,
                                         void* workspace_dev, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!tail_all_dev || !loss_dev || !workspace_dev || world < 1 || n_rows <= 0) return OADG_E_ARG;
  if (((uintptr_t)tail_all_dev & 15) || ((uintptr_t)workspace_dev & 255)) return OADG_E_ARG;
  LossWs w = carve_loss_ws(workspace_dev, world * n_rows, c);
  if (workspace_bytes < w.bytes) return OADG_E_ARG;
  finish_packed_kernel<<<(world * n_rows + 255) / 256, 256, 0, stream>>>(
      reinterpret_cast<const float4*>(tail_all_dev), world, n_rows, reinterpret_cast<float4*>(w.stats_all), loss_dev);
  OADG_LAUNCH_CHECK();
  return 0;
}

extern "C" int oadg_supcon_backward_packed(const float* feats_local_dev, const int32_t* pair_all_dev, int n_total,
                                           int row0, int n_rows, int c, float temperature, int normalized_input,
                                           const float* grad_loss_dev, float* grad_feats_dev, void* workspace_dev,
                                           size_t workspace_bytes, int* launches_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!feats_local_dev || !pair_all_dev || !grad_loss_dev || !grad_feats_dev || !workspace_dev) return OADG_E_ARG;
  if (n_rows <= 0 || row0 < 0 || row0 + n_rows > n_total) return OADG_E_ARG;
  if (c != kC) return OADG_E_LIMIT;
  if (!loss_tc_enabled()) return OADG_E_LIMIT;
  LossWs w = carve_loss_ws(workspace_dev, n_total, c);
  if (workspace_bytes < w.bytes) return OADG_E_ARG;
  int launches = 0;
  int rc = launch_sim_bwd_tc(w, w.stats_all, w.labels_all, pair_all_dev, n_total, row0, n_rows, 1.f / temperature, stream,
                             &launches);
  if (rc) return rc;
  normalize_bwd_kernel<<<(n_rows + 7) / 8, 256, 0, stream>>>(feats_local_dev, w.dpart, kBwdSplits, w.inv1, w.inv2, w.meta,
                                                            grad_loss_dev, n_rows, c, normalized_input, grad_feats_dev);
  OADG_LAUNCH_CHECK();
  ++launches;
  if (launches_out) *launches_out = launches;
  return 0;
}
