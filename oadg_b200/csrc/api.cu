// Library-level entry points of libOADG.so (see include/oadg.h).
#include "oadg_common.cuh"

extern "C" int oadg_abi_version(void) { return OADG_ABI_VERSION; }

extern "C" void oadg_struct_sizes(int32_t out[6]) {
  out[0] = (int32_t)sizeof(oadg_plan_header_t);
  out[1] = (int32_t)sizeof(oadg_view_t);
  out[2] = (int32_t)sizeof(oadg_gt_t);
  out[3] = (int32_t)sizeof(oadg_op_t);
  out[4] = (int32_t)sizeof(oadg_bbo_t);
  out[5] = (int32_t)sizeof(oadg_target_t);
}

extern "C" const char* oadg_error_string(int code) {
  if (code == 0) return "ok";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case OADG_E_ARG: return "OADG_E_ARG: null pointer, bad size or misaligned buffer";
    case OADG_E_PLAN: return "OADG_E_PLAN: malformed plan blob, or a launch whose work queue starved (incomplete views)";
    case OADG_E_LIMIT: return "OADG_E_LIMIT: exceeds a compiled limit";
    case OADG_E_NOBOX: return "OADG_E_NOBOX: no random box could be placed";
    case OADG_E_ROWS: return "OADG_E_ROWS: fewer rows than the two-view layout needs";
    default: return "unknown libOADG error";
  }
}

// Plain asynchronous copy on a stream.  The Python host side uses it for its page-locked staging buffers: a
// torch `copy_` costs 15-20 us of dispatch and host-allocator bookkeeping per call, this is one driver call.
extern "C" int oadg_memcpy_async(void* dst, const void* src, size_t bytes, int to_device, void* stream) {
  if (bytes == 0) return 0;
  if (!dst || !src) return OADG_E_ARG;
  return (int)cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                              (cudaStream_t)stream);
}
