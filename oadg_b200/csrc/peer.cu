// One-sided exchange between the ranks of a node for the gathered OA-Loss (include/oadg.h, "OA-Loss across ranks
// without a collective library on the critical path"): exportable buffers + CUDA IPC mapping on the host side, and on
// the device side a scatter kernel that stores a block of the caller's buffer into every peer's buffer over NVLink and
// raises a flag word there, plus the kernel that waits for the flags of all sources.
//
// Memory ordering: every thread's peer stores are followed by a system-scope fence before the block barrier; thread 0
// takes a ticket; the block with the last ticket fences again and writes the flags with st.release.sys.  The waiter
// polls with ld.acquire.sys, and what consumes the data are LATER kernels of the waiter's stream.
#include "oadg_common.cuh"
#include "oadg_peer.cuh"

namespace oadg {
namespace {

__global__ void __launch_bounds__(256)
peer_scatter_kernel(const oadg_peers_t P, size_t offset, size_t n_vec, size_t flag_offset, size_t counter_offset,
                    unsigned seq, unsigned tag) {
  const uint4* src = reinterpret_cast<const uint4*>(static_cast<const char*>(P.base[P.rank]) + offset);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * 256) {
    const uint4 v = src[i];
    for (int r = 0; r < P.world; ++r)
      if (r != P.rank) reinterpret_cast<uint4*>(static_cast<char*>(P.base[r]) + offset)[i] = v;
  }
  peer_signal(P, flag_offset, counter_offset, seq, tag);
}

__global__ void __launch_bounds__(32)
peer_wait_kernel(const unsigned* flags, int world, unsigned seq, unsigned tag, unsigned long long timeout_ns,
                 unsigned* fault_host) {
  const int r = threadIdx.x;
  if (r >= world) return;
  unsigned long long t0 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned spins = 0;
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
    if ((int)(v - seq) >= 0) {
      // the sender's view of the exchange must be ours (same row count on every rank), unless it is already a step ahead
      if (v == seq && flags[OADG_PEER_MAX + r] != tag) {
        *fault_host = 2u;
        __threadfence_system();
      }
      return;
    }
    __nanosleep(spins < 64 ? 20 : 200);
    if ((++spins & 1023u) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        *fault_host = 1u;   // page-locked, mapped: the host sees it without a synchronisation
        __threadfence_system();
        return;
      }
    }
  }
}

}  // namespace
}  // namespace oadg

using namespace oadg;

extern "C" int oadg_peer_alloc(size_t bytes, void** ptr_out) {
  if (!ptr_out || !bytes) return OADG_E_ARG;
  void* p = nullptr;
  OADG_CUDA_TRY(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  *ptr_out = p;
  return 0;
}
extern "C" int oadg_peer_free(void* ptr) {
  OADG_CUDA_TRY(cudaFree(ptr));
  return 0;
}
extern "C" int oadg_peer_export(void* ptr, unsigned char handle_out[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  if (!ptr || !handle_out) return OADG_E_ARG;
  cudaIpcMemHandle_t h;
  OADG_CUDA_TRY(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle_out, &h, 64);
  return 0;
}
extern "C" int oadg_peer_import(const unsigned char handle[64], void** ptr_out) {
  if (!handle || !ptr_out) return OADG_E_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  OADG_CUDA_TRY(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int oadg_peer_release(void* ptr) {
  OADG_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return 0;
}
extern "C" int oadg_peer_fault_alloc(uint32_t** fault_host_out) {
  if (!fault_host_out) return OADG_E_ARG;
  void* p = nullptr;
  OADG_CUDA_TRY(cudaHostAlloc(&p, 64, cudaHostAllocMapped | cudaHostAllocPortable));
  memset(p, 0, 64);
  *fault_host_out = static_cast<uint32_t*>(p);
  return 0;
}

extern "C" int oadg_peer_scatter(const oadg_peers_t* peers, size_t offset, size_t bytes, size_t flag_offset,
                                 size_t counter_offset, uint32_t seq, uint32_t tag, void* stream) {
  if (!peers || peers->world < 1 || peers->world > OADG_PEER_MAX || peers->rank < 0 || peers->rank >= peers->world)
    return OADG_E_ARG;
  if ((bytes & 15) || (offset & 15) || (flag_offset & 3) || (counter_offset & 3) || !bytes) return OADG_E_ARG;
  for (int r = 0; r < peers->world; ++r)
    if (!peers->base[r]) return OADG_E_ARG;
  const size_t n_vec = bytes / 16;
  int blocks = (int)((n_vec + 255) / 256);
  if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
  peer_scatter_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*peers, offset, n_vec, flag_offset, counter_offset, seq,
                                                                tag);
  OADG_LAUNCH_CHECK();
  return 0;
}

extern "C" int oadg_peer_wait(const uint32_t* flags_dev, int world, uint32_t seq, uint32_t tag, uint32_t timeout_ms,
                              uint32_t* fault_host, void* stream) {
  if (!flags_dev || !fault_host || world < 1 || world > OADG_PEER_MAX) return OADG_E_ARG;
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags_dev, world, seq, tag,
                                                       (unsigned long long)timeout_ms * 1000000ull, fault_host);
  OADG_LAUNCH_CHECK();
  return 0;
}
