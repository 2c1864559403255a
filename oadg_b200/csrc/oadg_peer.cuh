// Device side of the one-sided peer exchange (peer.cu, oaloss.cu): "all my stores are out, raise my flag everywhere".
#pragma once
#include "oadg_common.cuh"

namespace oadg {

// Called by EVERY thread of EVERY block of a kernel after its last store to peer memory.  The block that finishes last
// writes, into every rank's buffer, tag word [OADG_PEER_MAX + rank] := tag (what the sender thinks the exchange looks
// like: its row count) and then flag word [rank] := seq.
__device__ __forceinline__ void peer_signal(const oadg_peers_t& P, size_t flag_offset, size_t counter_offset, unsigned seq,
                                            unsigned tag) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* counter = reinterpret_cast<unsigned*>(static_cast<char*>(P.base[P.rank]) + counter_offset);
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {
      __threadfence_system();
      *counter = 0u;
      for (int r = 0; r < P.world; ++r) {
        unsigned* flag = reinterpret_cast<unsigned*>(static_cast<char*>(P.base[r]) + flag_offset) + P.rank;
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(flag + OADG_PEER_MAX), "r"(tag) : "memory");
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
      }
    }
  }
}

}  // namespace oadg
