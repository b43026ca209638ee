// Thin wrappers over the sm_100a PTX used by both Gram kernels: mbarrier,
// TMA (cp.async.bulk.tensor) and the host-side tensor-map encoder.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

namespace scedar {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// make freshly initialised barriers visible to the async (TMA / tensor) proxy
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a descriptor or protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  const long long t0 = clock64();
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 8000000000LL) {
      printf("scedar: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  } while (!ok);
}
// The same wait with a suspend-time hint: the hardware parks the thread for up to
// `ns` nanoseconds per attempt (it wakes as soon as the phase completes), so a
// long wait costs a handful of polls instead of a spin.
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  const long long t0 = clock64();
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(ns)
        : "memory");
    if (!ok && clock64() - t0 > 8000000000LL) {
      printf("scedar: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  } while (!ok);
}
// The same wait for register-starved loops (the FP64 Gram consumers hold 128
// accumulator registers): 32-bit clock, no printf call on the time-out path.
__device__ __forceinline__ void mbar_wait_lite(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t t0 = (uint32_t)clock();
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && (uint32_t)clock() - t0 > 3000000000u) __trap();   // ~1.5 s: a protocol bug
  } while (!ok);
}
// 2-D tiled TMA load: box at (c0 = inner element, c1 = row) -> shared memory,
// completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// Row-major [rows, cols] matrix, boxes of box_rows x box_cols elements whose
// inner extent in bytes equals the swizzle span (128 or 64).
int encode_tensor_map(CUtensorMap* m, CUtensorMapDataType dtype, int elem_bytes, const void* ptr,
                      int64_t rows, int64_t cols, int box_rows, int box_cols,
                      CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B);

}  // namespace scedar
