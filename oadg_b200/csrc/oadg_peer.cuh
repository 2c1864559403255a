// Device side of the one-sided peer exchange (peer.cu, oaloss.cu): "all my stores are out, raise my flag everywhere".
#pragma once
#include "oadg_common.cuh"

namespace oadg {

// Called by EVERY thread of EVERY block of a kernel after its last store to peer memory.  flag word [rank] at
// flag_offset of every rank's buffer := seq, written by the block that finishes last.
__device__ __forceinline__ void peer_signal(const oadg_peers_t& P, size_t flag_offset, size_t counter_offset, unsigned seq) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* counter = reinterpret_cast<unsigned*>(static_cast<char*>(P.base[P.rank]) + counter_offset);
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {
      __threadfence_system();
      *counter = 0u;
      for (int r = 0; r < P.world; ++r) {
        unsigned* flag = reinterpret_cast<unsigned*>(static_cast<char*>(P.base[r]) + flag_offset) + P.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
      }
    }
  }
}

}  // namespace oadg
