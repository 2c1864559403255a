// Shared declarations of the OA-Loss kernels (oaloss.cu: CUDA-core path + API, oaloss_tc.cu: tcgen05 path).
#pragma once
#include "oadg_common.cuh"

namespace oadg {

constexpr int kBwdSplits = 8;   // column splits of the tcgen05 backward (deterministic partial sums)

struct RowStats {  // per row, kept for backward
  float lse;       // log sum_{k != i} exp(z_ik)
  float coef;      // -(w/N)/n_i or 0
  float npos;      // n_i
  float u;         // coef * n_i * exp(-lse): the softmax part of dL/dz_ij is -exp(z_ij) * (u_i + u_j)
};

struct LossWs {
  float* fhat;       // [n, c] doubly normalised embeddings
  float* inv1;       // [n] 1/max(||x||, eps)
  float* inv2;       // [n] 1/max(||x/||x||||, eps)
  RowStats* stats;   // [n]
  int* meta;         // [0] bg label (low 32 bits), [1] n_fg, [2] active flag, [4] ticket of the row-reduce blocks
  double* red;       // [<= 1024] per-block partial sums of the row reduce
  float* npos;       // [n]
  float* partial;    // [col_tiles][n][3]  (max, sumexp, possum)
  float* dfhat;      // [n, c] gradient wrt fhat
  float* f_hi;       // [n, c] TF32 split of fhat for the tcgen05 path
  float* f_lo;
  float* ft_hi;      // [c, ld] transposed split (B operand of the backward GEMM)
  float* ft_lo;
  float* z;          // [n, ld] logits kept by the tcgen05 forward for its backward (L2-resident, 17 MB at n=2088)
  float* dpart;      // [kBwdSplits][n, c] partial gradients of the tcgen05 backward
  int64_t* labels_all;   // [n] labels of the gathered rows (cross-rank path: unpacked from the gathered buffer)
  RowStats* stats_all;   // [n] row statistics of every rank (cross-rank path)
  int fhat_ld;       // row stride of fhat in floats (c, or the packed width of a gathered buffer)
  int ld;            // n rounded up to 32
  size_t bytes;
};

inline LossWs carve_loss_ws(void* base, int n, int c) {
  LossWs w;
  char* p = static_cast<char*>(base);
  size_t o = 0;
  auto take = [&](size_t b) {
    size_t at = o;
    o = align_up(o + b, 256);
    return at;
  };
  // partial row statistics: one per 64-column tile (FFMA test build) or per half of a 128-column tile (tcgen05)
  const int col_tiles = (n + 63) / 64 > 2 * ((n + 127) / 128) ? (n + 63) / 64 : 2 * ((n + 127) / 128);
  size_t o_f = take((size_t)n * c * 4), o_i1 = take((size_t)n * 4), o_i2 = take((size_t)n * 4);
  size_t o_st = take((size_t)n * sizeof(RowStats)), o_meta = take(64), o_np = take((size_t)n * 4);
  size_t o_pa = take((size_t)col_tiles * n * 3 * 4), o_df = take((size_t)n * c * 4);
  size_t o_hi = take((size_t)n * c * 4), o_lo = take((size_t)n * c * 4);
  const int ld = (n + 31) / 32 * 32;
  size_t o_th = take((size_t)c * ld * 4), o_tl = take((size_t)c * ld * 4);
  size_t o_z = take((size_t)n * ld * 4), o_dp = take((size_t)kBwdSplits * n * c * 4);
  size_t o_red = take(1024 * sizeof(double));
  size_t o_la = take((size_t)n * sizeof(int64_t)), o_sa = take((size_t)n * sizeof(RowStats));
  w.fhat = reinterpret_cast<float*>(p + o_f);
  w.inv1 = reinterpret_cast<float*>(p + o_i1);
  w.inv2 = reinterpret_cast<float*>(p + o_i2);
  w.stats = reinterpret_cast<RowStats*>(p + o_st);
  w.meta = reinterpret_cast<int*>(p + o_meta);
  w.npos = reinterpret_cast<float*>(p + o_np);
  w.partial = reinterpret_cast<float*>(p + o_pa);
  w.dfhat = reinterpret_cast<float*>(p + o_df);
  w.f_hi = reinterpret_cast<float*>(p + o_hi);
  w.f_lo = reinterpret_cast<float*>(p + o_lo);
  w.ft_hi = reinterpret_cast<float*>(p + o_th);
  w.ft_lo = reinterpret_cast<float*>(p + o_tl);
  w.z = reinterpret_cast<float*>(p + o_z);
  w.dpart = reinterpret_cast<float*>(p + o_dp);
  w.red = reinterpret_cast<double*>(p + o_red);
  w.labels_all = reinterpret_cast<int64_t*>(p + o_la);
  w.stats_all = reinterpret_cast<RowStats*>(p + o_sa);
  w.fhat_ld = c;
  w.ld = ld;
  w.bytes = o;
  return w;
}


// oaloss_tc.cu
// rows [row0, row0 + n_rows) of the n-row problem are the anchors (single GPU: row0 = 0, n_rows = n)
int launch_sim_fwd_tc(const LossWs& w, const int64_t* labels, const int32_t* pair, int n, int row0, int n_rows,
                      float inv_t, cudaStream_t stream, int* launches);
int launch_sim_bwd_tc(const LossWs& w, const RowStats* stats_all, const int64_t* labels, const int32_t* pair, int n,
                      int row0, int n_rows, float inv_t, cudaStream_t stream, int* launches);

}  // namespace oadg
