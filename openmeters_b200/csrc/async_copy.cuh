// async_copy.cuh — 16-byte asynchronous global -> shared copies (cp.async.cg, LDGSTS) used by the staging rings of the
// specialised kernels: issued one iteration ahead, committed as one group, waited for before the barrier that publishes
// the ring.  Under the CPU emulator (OMB_EMU) the copy is synchronous and commit / wait are no-ops.
#pragma once
#include "common.h"

namespace omb {

__device__ __forceinline__ void async_copy16(float* dst_smem, const float* src_gmem) {
#ifdef OMB_EMU
  for (int i = 0; i < 4; ++i) dst_smem[i] = src_gmem[i];
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src_gmem));
#endif
}
__device__ __forceinline__ void async_commit() {
#ifndef OMB_EMU
  asm volatile("cp.async.commit_group;\n" ::);
#endif
}
__device__ __forceinline__ void async_wait_all() {
#ifndef OMB_EMU
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
}

// Copies lane samples [s0, s1) (multiples of 4) into a power-of-two ring at (sample index mod ring length); all `n_threads`
// threads of the CTA take part.
__device__ __forceinline__ void ring_fetch_pow2(float* ring, int ring_mask, const float* x, uint64_t s0, uint64_t s1, int n_threads) {
  for (uint64_t s = s0 + 4ull * threadIdx.x; s < s1; s += 4ull * (uint64_t)n_threads) async_copy16(ring + ((int)s & ring_mask), x + s);
}

}  // namespace omb
