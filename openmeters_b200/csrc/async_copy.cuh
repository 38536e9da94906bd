// async_copy.cuh — asynchronous global <-> shared copies of the specialised kernels' staging rings and output columns.
//  * 16-byte per-thread copies (cp.async.cg, LDGSTS): issued one iteration ahead, committed as one group, waited for before
//    the barrier that publishes the ring;
//  * bulk copies by one elected thread (cp.async.bulk, the TMA engine's 1-D mode, UBLKCP) with mbarrier completion — second half.
// Under the CPU emulator (OMB_EMU) copies are synchronous and commit / wait are no-ops.
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

// ---------------------------------------------------------------------------------------------------------------------
// Bulk asynchronous copies (the TMA engine's 1-D mode: cp.async.bulk, SASS UBLKCP) with mbarrier completion.
// ONE elected thread moves a contiguous run of bytes; no per-thread address arithmetic, no LSU issue slots:
//   global -> shared : `bulk_g2s`, completion counted in bytes on an mbarrier (`mbar_expect_tx` + `mbar_wait`)
//   shared -> global : `bulk_s2g` + `bulk_commit` + `bulk_wait_read` (the source may be overwritten afterwards)
// Addresses and sizes are multiples of 16 bytes.  Under the CPU emulator the copies are synchronous memcpys and the
// mbarrier is modelled in its own 64 bits (pending bytes | armed flag | phase counter), so a thread that reaches
// `mbar_wait` before the elected thread has issued the copy yields until the phase completes, as on the device.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* mbar, unsigned arrivals) {
#ifndef OMB_EMU
  const unsigned a = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(arrivals));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
  (void)arrivals;  // every use has one arriving thread (the one that issues the copies)
  *mbar = 0;
#endif
}
#ifdef OMB_EMU
// emulated mbarrier word: bits 0..31 phase counter, bit 32 armed (the arrival happened), bits 33..63 pending bytes
__device__ __forceinline__ void mbar_emu_settle(uint64_t* mbar) {
  if (((*mbar >> 32) & 1u) && (*mbar >> 33) == 0) *mbar = (uint32_t)(*mbar) + 1u;  // armed and nothing pending: next phase
}
#endif
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, unsigned bytes) {  // one arrival + `bytes` pending
#ifndef OMB_EMU
  const unsigned a = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
#else
  *mbar += ((uint64_t)bytes << 33) | (1ull << 32);
  mbar_emu_settle(mbar);
#endif
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {  // one arrival, no bytes (release semantics at CTA scope)
#ifndef OMB_EMU
  const unsigned a = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
#else
  *mbar += 1ull << 32;
  mbar_emu_settle(mbar);
#endif
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, unsigned parity) {
#ifndef OMB_EMU
  const unsigned a = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "OMB_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra OMB_MBAR_DONE;\n"
      "bra OMB_MBAR_WAIT;\n"
      "OMB_MBAR_DONE:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
#else
  while (((uint32_t)(*mbar) & 1u) == parity) omb_emu::yield();
#endif
}
// Same, for a thread that expects to wait long (a consumer warp polling a producer): sleeps between polls so the spin does not
// take issue slots from the warps it is waiting for.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* mbar, unsigned parity, unsigned sleep_ns) {
#ifndef OMB_EMU
  const unsigned a = (unsigned)__cvta_generic_to_shared(mbar);
  unsigned done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(sleep_ns);
  }
#else
  (void)sleep_ns;
  while (((uint32_t)(*mbar) & 1u) == parity) omb_emu::yield();
#endif
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* mbar) {
#ifdef OMB_EMU
  memcpy(dst_smem, src_gmem, bytes);
  *mbar -= (uint64_t)bytes << 33;
  mbar_emu_settle(mbar);
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem), m = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src_gmem),
               "r"(bytes), "r"(m) : "memory");
#endif
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
#ifdef OMB_EMU
  memcpy(dst_gmem, src_smem, bytes);
#else
  const unsigned s = (unsigned)__cvta_generic_to_shared(src_smem);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(s), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void bulk_commit() {
#ifndef OMB_EMU
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void bulk_wait_read() {  // every committed shared -> global copy has finished READING shared memory
#ifndef OMB_EMU
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
// Generic-proxy writes to shared memory (ordinary stores) must be made visible to the async proxy before a bulk copy reads them.
__device__ __forceinline__ void fence_async_smem() {
#ifndef OMB_EMU
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// Lane samples [s0, s1) (multiples of 4 floats) into a power-of-two ring at (sample index mod ring length) as one or two bulk
// copies issued by the calling thread; returns the number of bytes in flight (what mbar_expect_tx must have announced).
__device__ __forceinline__ unsigned ring_bytes(uint64_t s0, uint64_t s1) { return s0 < s1 ? (unsigned)((s1 - s0) * sizeof(float)) : 0u; }
__device__ __forceinline__ void ring_fetch_bulk(float* ring, int ring_mask, const float* x, uint64_t s0, uint64_t s1, uint64_t* mbar) {
  if (s0 >= s1) return;
  const int p0 = (int)s0 & ring_mask;
  const uint64_t n = s1 - s0, room = (uint64_t)(ring_mask + 1 - p0);
  if (n <= room) {
    bulk_g2s(ring + p0, x + s0, (unsigned)(n * sizeof(float)), mbar);
  } else {
    bulk_g2s(ring + p0, x + s0, (unsigned)(room * sizeof(float)), mbar);
    bulk_g2s(ring, x + s0 + room, (unsigned)((n - room) * sizeof(float)), mbar);
  }
}

}  // namespace omb
