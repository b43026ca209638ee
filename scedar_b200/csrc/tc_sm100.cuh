// tcgen05 / TMEM helpers shared by the tensor-core Gram kernels (sm_100a):
// shared-memory matrix descriptors, MMA issue for kind::i8, commit, TMEM
// allocation and loads.  (gram_tc.cu keeps its own private copies of the f16
// forms it was tuned with.)
#pragma once

#include <stdint.h>

#include "ptx_sm100.cuh"

namespace scedar {
namespace tc {

// K-major operand tile in 64B-swizzled shared memory: rows of 64 bytes (one
// K slab), 8-row swizzle atoms 512 bytes apart.  Only the 14-bit start-address
// field varies between tiles, so a tile at byte offset `off` (a multiple of 16)
// from this one has descriptor `desc + (off >> 4)`.
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(512 >> 4) << 32;     // stride byte offset between 8-row atoms
  d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
  d |= (uint64_t)4 << 61;              // SWIZZLE_64B
  return d;
}

// The same for 128B-swizzled tiles: rows of 128 bytes, 8-row atoms 1024 bytes
// apart.  A K step of 32 bytes inside the row advances the start address by 32.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}

// instruction descriptor of kind::i8: D = S32, A and B unsigned 8-bit, both
// K-major, M = 128 (cta_group::1)
__host__ __device__ constexpr uint32_t idesc_u8(int n) {
  return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on `bar` when the MMAs issued so far have retired
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// whole warp; writes the TMEM base address to the shared-memory word at `slot`
__device__ __forceinline__ void tmem_alloc512(uint32_t slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot),
               "r"(512)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free512(uint32_t tmem_base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
               : "memory");
}
// 32 lanes x 4 consecutive columns -> 4 registers per thread; NOT complete
// until tmem_wait() has been executed
__device__ __forceinline__ void tmem_ld4_async(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
// the registers are in/out operands so that no use of them can be scheduled
// above the wait (one call per loaded array; the first one does the waiting)
__device__ __forceinline__ void tmem_wait4(uint32_t (&v)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3])
               :
               : "memory");
}

}  // namespace tc
}  // namespace scedar
