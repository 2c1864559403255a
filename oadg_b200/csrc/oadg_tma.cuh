// TMA (cp.async.bulk.tensor) + mbarrier PTX wrappers and the host-side tensor-map encoder (sm_100a).
// Used by the OA-Mix chain kernel to stage the source rectangles of its affine gathers: a frame is a 2-D
// uint8 tensor [H rows][3W bytes]; out-of-frame parts of a box arrive as zeros, which is exactly
// cv2.warpAffine's BORDER_CONSTANT 0 (reference augmix.py:92,116,136,156,177).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace oadg {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 2-D tile load: coordinates are element indices {inner, outer}; may be negative / past the end (zero fill)
__device__ __forceinline__ void load_2d(void* dst, const void* map, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
// A tensor map that lives in GLOBAL memory (written by a host copy before the launch) is read through the tensormap
// proxy: every CTA acquires it once before its first use (CUDA programming guide, "tensor map in global memory").
__device__ __forceinline__ void fence_tensormap_acquire(const void* map) {
  asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(map) : "memory");
}
// orders this thread's earlier generic-proxy accesses (and what it acquired) before its later async-proxy ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return (EncodeTiledFn)p;
    return (EncodeTiledFn) nullptr;
  }();
  return fn;
}
// uint8 tensor [rows][inner_bytes] with row pitch `pitch` bytes, box [box_rows][box_inner]; 0 on success.
// Requirements (cuTensorMapEncodeTiled): base 16-byte aligned, pitch a multiple of 16, box_inner a multiple of 16
// and <= 256, box_rows <= 256.
inline int encode_u8_2d(void* map_out, const void* base, uint64_t inner_bytes, uint64_t rows, uint64_t pitch,
                        uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch & 15) != 0 || (box_inner & 15) != 0 || box_inner > 256 ||
      box_rows > 256 || inner_bytes == 0 || rows == 0)
    return -1;
  cuuint64_t dims[2] = {inner_bytes, rows};
  cuuint64_t strides[1] = {pitch};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

}  // namespace tma
}  // namespace oadg
