// Jensen-Shannon consistency term of OA-Loss between the two views' class distributions
// (reference mmdet/models/losses/oadg/cross_entropy_loss_plus.py:264-319 `jsdv1_3_2aug`): one fused pass computes
//   p = softmax(a_i) (or (sigmoid, 1 - sigmoid) for the single-logit RPN head), q likewise from view 2,
//   m = clamp((p + q) / 2, 1e-7, 1),   row = 1/2 sum_c [ p (log p - log m) + q (log q - log m) ],
//   loss = sum_i row_i        (the reference divides by len() of a tensor it has just reshaped to [1, n, C], i.e. by 1,
//                            :300-310 -- not by the rows of a view, whatever its docstring says)
// AND d loss / d logits of both views (the reference's autograd through kl_div, log, clamp, softmax / sigmoid), so the
// backward is a scale.  One thread per row pair, three light passes over the <= 32 logits of a row (L1-resident);
// block partial sums are added in block order by the last block (deterministic).
#include "oadg_common.cuh"

namespace oadg {
namespace {

constexpr int kJsdBlocks = 296, kJsdThreads = 256;

__device__ __forceinline__ float xlogx_minus(float p, float lm) { return p > 0.f ? p * (logf(p) - lm) : 0.f; }

__global__ void __launch_bounds__(kJsdThreads)
jsd2_kernel(const float* __restrict__ pred, int n, int c, float* __restrict__ loss, float* __restrict__ grad,
            float* __restrict__ partial, unsigned* __restrict__ ticket) {
  __shared__ double sred[kJsdThreads / 32];
  double local = 0.0;
  for (int i = blockIdx.x * kJsdThreads + threadIdx.x; i < n; i += gridDim.x * kJsdThreads) {
    const float* a = pred + (size_t)i * c;
    const float* b = pred + (size_t)(n + i) * c;
    float* ga = grad + (size_t)i * c;
    float* gb = grad + (size_t)(n + i) * c;
    if (c == 1) {   // RPN objectness: classes (s, 1 - s)
      const float sa = 1.f / (1.f + expf(-a[0])), sb = 1.f / (1.f + expf(-b[0]));
      float row = 0.f, dsa = 0.f, dsb = 0.f;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float p = k ? 1.f - sa : sa, q = k ? 1.f - sb : sb;
        const float mp = 0.5f * (p + q), m = fminf(fmaxf(mp, 1e-7f), 1.f), lm = logf(m);
        row += 0.5f * (xlogx_minus(p, lm) + xlogx_minus(q, lm));
        const float through = (mp >= 1e-7f && mp <= 1.f) ? 0.25f * (p + q) / m : 0.f;
        const float gp = p > 0.f ? 0.5f * (logf(p) + 1.f - lm) - through : 0.f;
        const float gq = q > 0.f ? 0.5f * (logf(q) + 1.f - lm) - through : 0.f;
        dsa += k ? -gp : gp;
        dsb += k ? -gq : gq;
      }
      ga[0] = sa * (1.f - sa) * dsa;
      gb[0] = sb * (1.f - sb) * dsb;
      local += (double)row;
      continue;
    }
    float ma = -INFINITY, mb = -INFINITY;
    for (int k = 0; k < c; ++k) {
      ma = fmaxf(ma, a[k]);
      mb = fmaxf(mb, b[k]);
    }
    float za = 0.f, zb = 0.f;
    for (int k = 0; k < c; ++k) {
      za += expf(a[k] - ma);
      zb += expf(b[k] - mb);
    }
    const float ia = 1.f / za, ib = 1.f / zb;
    float row = 0.f, dot_a = 0.f, dot_b = 0.f;   // dot = sum_c p_c g_c
    for (int k = 0; k < c; ++k) {
      const float p = expf(a[k] - ma) * ia, q = expf(b[k] - mb) * ib;
      const float mp = 0.5f * (p + q), m = fminf(fmaxf(mp, 1e-7f), 1.f), lm = logf(m);
      row += 0.5f * (xlogx_minus(p, lm) + xlogx_minus(q, lm));
      const float through = (mp >= 1e-7f && mp <= 1.f) ? 0.25f * (p + q) / m : 0.f;
      const float gp = p > 0.f ? 0.5f * (logf(p) + 1.f - lm) - through : 0.f;
      const float gq = q > 0.f ? 0.5f * (logf(q) + 1.f - lm) - through : 0.f;
      dot_a += p * gp;
      dot_b += q * gq;
    }
    for (int k = 0; k < c; ++k) {
      const float p = expf(a[k] - ma) * ia, q = expf(b[k] - mb) * ib;
      const float mp = 0.5f * (p + q), m = fminf(fmaxf(mp, 1e-7f), 1.f), lm = logf(m);
      const float through = (mp >= 1e-7f && mp <= 1.f) ? 0.25f * (p + q) / m : 0.f;
      const float gp = p > 0.f ? 0.5f * (logf(p) + 1.f - lm) - through : 0.f;
      const float gq = q > 0.f ? 0.5f * (logf(q) + 1.f - lm) - through : 0.f;
      ga[k] = p * (gp - dot_a);
      gb[k] = q * (gq - dot_b);
    }
    local += (double)row;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kJsdThreads / 32; ++w) t += sred[w];
    partial[blockIdx.x] = (float)t;
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {   // the last block: every partial sum is visible
      __threadfence();
      double tot = 0.0;
      for (unsigned b2 = 0; b2 < gridDim.x; ++b2) tot += (double)*((volatile float*)partial + b2);
      *loss = (float)tot;
      *ticket = 0u;
    }
  }
}

}  // namespace
}  // namespace oadg

// pred_dev [2 n, c] float32 (view 1 rows, then view 2 rows), c <= 32; loss_dev [1]; grad_dev [2 n, c];
// scratch_dev: kJsdBlocks floats + one zero-initialised uint32 (the caller keeps it; the kernel leaves the counter 0)
extern "C" int oadg_jsd2_scratch_bytes(void) { return (int)(oadg::kJsdBlocks * sizeof(float) + 64); }

extern "C" int oadg_jsd2_forward(const float* pred_dev, int n, int c, float* loss_dev, float* grad_dev,
                                 void* scratch_dev, void* stream) {
  using namespace oadg;
  if (!pred_dev || !loss_dev || !grad_dev || !scratch_dev || n < 1) return OADG_E_ARG;
  if (c < 1 || c > 32) return OADG_E_LIMIT;
  float* partial = static_cast<float*>(scratch_dev);
  unsigned* ticket = reinterpret_cast<unsigned*>(partial + kJsdBlocks);
  int blocks = (n + kJsdThreads - 1) / kJsdThreads;
  if (blocks > kJsdBlocks) blocks = kJsdBlocks;
  jsd2_kernel<<<blocks, kJsdThreads, 0, (cudaStream_t)stream>>>(pred_dev, n, c, loss_dev, grad_dev, partial, ticket);
  OADG_LAUNCH_CHECK();
  return 0;
}
