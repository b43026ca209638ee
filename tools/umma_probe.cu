// Probe for the INT8 tensor path (tcgen05.mma.kind::i8, INT32 accumulators in
// TMEM) that the error-free-split Gram kernel (scedar_b200/csrc/gram_oz.cu) is
// built on.  Two questions, both answered by measurement on a B200:
//   check   is the operand plumbing right?  TMA (64-byte swizzle) of a uint8
//           matrix -> K-major smem descriptors -> kind::i8 with every
//           signed/unsigned operand combination, cta_group 1 and 2, against a
//           CPU int64 dot product.
//   rate    cycles per MMA for M = 128 x cta_group, N in {32..256}, kind i8 and
//           f16, issued in the slice-pair pattern of the split Gram (S slices,
//           one accumulator per diagonal s + t) -- is N = 64 paced by the
//           tensor pipe (N/2 cycles) or by the shared-memory operand reads?
// Build: make -C tools umma_probe      Run (GPU box): tools/umma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../scedar_b200/csrc/ptx_sm100.cuh"

using namespace scedar;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static int encode_u8_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int box_rows,
                         int box_cols) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<encode_fn>(fn)(
      m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return 1;
  }
  return 0;
}

// ---- device helpers (same forms as gram_tc.cu) -------------------------------
__device__ __forceinline__ uint64_t umma_desc64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;   // 8-row atoms of 64-byte rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;            // SWIZZLE_64B
  return d;
}
template <int CG, int KIND>   // KIND 0 = i8, 1 = f16
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc,
                                     uint32_t acc) {
  if (CG == 1 && KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  if (CG == 2 && KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  if (CG == 1 && KIND == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  if (CG == 2 && KIND == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
  else
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(bar),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ uint32_t ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void csync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}
__device__ __forceinline__ void tma_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                         uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
template <int CG>
__device__ __forceinline__ uint32_t tmem_setup(uint32_t slot_addr, volatile uint32_t* slot, int warp) {
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_addr),
                   "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_addr),
                   "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) csync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *slot;
}
template <int CG>
__device__ __forceinline__ void tmem_teardown(uint32_t tmem_base, int warp) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) csync();
  if (warp == 0) {
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                   : "memory");
  }
}

// idesc of kind::i8: D = S32 (2 << 4), A/B format 0 = u8, 1 = s8, K-major both
static uint32_t idesc_i8(int M, int N, int a_signed, int b_signed) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
static uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- check: one 128*CG x N tile, K = 64 bytes (two MMAs) -----------------------
// matrix g: rows [0, 256) are A rows, rows [256, 256 + N) are B rows, 64 columns
template <int CG>
__global__ void __launch_bounds__(128, 1)
check_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             int N, uint32_t idesc, int32_t* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - raw);
  const uint32_t sa = base, sb = base + 8192;
  const uint32_t bars = base + 8192 + 16384;
  const uint32_t full = bars, done = bars + 8;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(g + 8192 + 16384 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    mbar_fence_init();
  }
  const uint32_t tmem_base = tmem_setup<CG>(bars + 64, slot, warp);
  const int nb = N / CG;   // B rows staged by this CTA
  if (threadIdx.x == 0) {
    if (CG == 1) {
      mbar_expect_tx(full, 8192 + nb * 64);
      tma_load_2d(sa, &map_a, 0, 0, full);
      tma_load_2d(sb, &map_b, 0, 256, full);
    } else {
      if (rank == 0) mbar_expect_tx(full, 2 * (8192 + nb * 64));
      const uint32_t lbar = mapa(full, 0);
      tma_pair(sa, &map_a, 0, (int)rank * 128, lbar);
      tma_pair(sb, &map_b, 0, 256 + (int)rank * nb, lbar);
    }
    if (rank == 0) {
      mbar_wait(full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kk = 0; kk < 2; ++kk)
        umma<CG, 0>(tmem_base, umma_desc64(sa + kk * 32), umma_desc64(sb + kk * 32), idesc, kk != 0);
      commit<CG>(done);
    }
  }
  __syncthreads();
  mbar_wait(done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = (int)rank * 128 + warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int c = 0; c < 32; ++c) out[row * N + c0 + c] = (int32_t)v[c];
  }
  tmem_teardown<CG>(tmem_base, warp);
}

// ---- rate: S slices of A and B resident in smem, the slice-pair pattern -------
// mode 0: pairs (s, t), s + t < S, accumulator s + t (needs S * N <= 512)
// mode 1: plain GEMM pattern: A slice (i % S), B slice (i % S), accumulators 0/1
template <int CG, int KIND>
__global__ void __launch_bounds__(128, 1)
rate_kernel(int N, int S, int iters, int mode, uint32_t idesc, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - raw);
  const int nb = N / CG;
  const uint32_t a_bytes = 8192, b_bytes = (uint32_t)nb * 64;
  const uint32_t sa = base, sb = base + S * a_bytes;
  const uint32_t bars = sb + S * ((b_bytes + 1023u) & ~1023u);
  const uint32_t done = bars;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(g + (bars - base) + 64);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = CG == 2 ? ctarank() : 0;
  // small integers / small halves: no NaNs, no denormal slow paths
  for (uint32_t i = threadIdx.x; i < (bars - base) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(g)[i] = KIND == 0 ? 0x01020301u : 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    mbar_fence_init();
  }
  const uint32_t tmem_base = tmem_setup<CG>(bars + 64, slot, warp);
  const uint32_t bstride = (b_bytes + 1023u) & ~1023u;
  if (threadIdx.x == 0 && rank == 0) {
    const long long t0 = clock64();
    long long n_mma = 0;
    for (int it = 0; it < iters; ++it) {
      for (int kk = 0; kk < 2; ++kk) {
        if (mode == 0) {
          for (int s = 0; s < S; ++s)
            for (int t = 0; s + t < S; ++t) {
              umma<CG, KIND>(tmem_base + (s + t) * N, umma_desc64(sa + s * a_bytes + kk * 32),
                             umma_desc64(sb + t * bstride + kk * 32), idesc, 1);
              ++n_mma;
            }
        } else {
          for (int s = 0; s < S; ++s) {
            umma<CG, KIND>(tmem_base + (s & 1) * N, umma_desc64(sa + s * a_bytes + kk * 32),
                           umma_desc64(sb + s * bstride + kk * 32), idesc, 1);
            ++n_mma;
          }
        }
      }
    }
    commit<CG>(done);
    mbar_wait(done, 0);
    const long long t1 = clock64();
    out[2 * (blockIdx.x / CG)] = (unsigned long long)(t1 - t0);
    out[2 * (blockIdx.x / CG) + 1] = (unsigned long long)n_mma;
  }
  if (CG == 2 && rank == 1 && threadIdx.x == 0) mbar_wait(done, 0);
  tmem_teardown<CG>(tmem_base, warp);
}

// ---- rate2: a tight, warp-uniform issue loop (elected lane, descriptors =
// base + constant, fully unrolled) -- what gram_oz.cu really issues.  cta_group 1.
//   mode 0  SS: A and B from shared memory, one N
//   mode 1  TS: A from TMEM (8 columns at 480), B from shared memory, one N
//   mode 2  SS, the S = 7 stacked schedule of gram_oz.cu (10 instructions / K step)
//   mode 3  TS + one tcgen05.cp 128x256b per 4 MMAs (re-staging A in TMEM)
//   mode 4  SS, every instruction N = 256 (7 per K step: the ideal packing)
__device__ __forceinline__ bool elect1() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b,
                                        uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
               "r"(a_tmem), "l"(b), "r"(idesc)
               : "memory");
}
__device__ __forceinline__ void umma_ts_acc(uint32_t d_tmem, uint32_t a_tmem, uint64_t b,
                                            uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
               "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tmem_cp(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__global__ void __launch_bounds__(128, 1)
rate2_kernel(int N, int mode, int iters, unsigned long long* out, int random_fill = 0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - raw);
  // A: 7 slices x 8 KB (128 rows x 64 B); B: 7 slices x 4 KB stacked (64 rows x 64 B)
  const uint32_t sa = base, sb = base + 7 * 8192;
  const uint32_t bars = sb + 16 * 4096;
  const uint32_t done = bars;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(g + (bars - base) + 64);
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < (bars - base) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(g)[i] =
        random_fill ? (i * 2654435761u ^ (i >> 7) * 40503u ^ 0x9e3779b9u) * 2246822519u : 0x01020301u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    mbar_fence_init();
  }
  const uint32_t tmem_base = tmem_setup<1>(bars + 64, slot, warp);
  if (warp == 0) {
    const bool el = elect1();
    const uint64_t da = umma_desc64(sa), db = umma_desc64(sb);
    const uint32_t idN = (2u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t id64 = (2u << 4) | (8u << 17) | (8u << 24), id128 = (2u << 4) | (16u << 17) | (8u << 24),
                   id192 = (2u << 4) | (24u << 17) | (8u << 24), id256 = (2u << 4) | (32u << 17) | (8u << 24);
    const uint32_t a_t = tmem_base + 480;
    __syncwarp();
    const long long t0 = clock64();
    long long n_mma = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        if (mode == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (el) umma<1, 0>(tmem_base + (q & 1) * 224, da + (uint64_t)(((q % 7) * 8192 + kk * 32) >> 4),
                               db + (uint64_t)((kk * 32) >> 4), idN, 1);
          n_mma += 8;
        } else if (mode == 1) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (el) umma_ts(tmem_base + (q & 1) * 224, a_t, db + (uint64_t)((kk * 32) >> 4), idN);
          n_mma += 8;
        } else if (mode == 2) {
          // slice s x slices [0, 7 - s): 4+3, 3+3, 3+2, 4, 3, 2, 1
          if (el) {
            umma<1, 0>(tmem_base + 0, da + (uint64_t)((0 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id256, 1);
            umma<1, 0>(tmem_base + 256, da + (uint64_t)((0 * 8192 + kk * 32) >> 4), db + (uint64_t)((4 * 4096 + kk * 32) >> 4), id192, 1);
            umma<1, 0>(tmem_base + 64, da + (uint64_t)((1 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id192, 1);
            umma<1, 0>(tmem_base + 256, da + (uint64_t)((1 * 8192 + kk * 32) >> 4), db + (uint64_t)((3 * 4096 + kk * 32) >> 4), id192, 1);
            umma<1, 0>(tmem_base + 128, da + (uint64_t)((2 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id192, 1);
            umma<1, 0>(tmem_base + 320, da + (uint64_t)((2 * 8192 + kk * 32) >> 4), db + (uint64_t)((3 * 4096 + kk * 32) >> 4), id128, 1);
            umma<1, 0>(tmem_base + 192, da + (uint64_t)((3 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id256, 1);
            umma<1, 0>(tmem_base + 256, da + (uint64_t)((4 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id192, 1);
            umma<1, 0>(tmem_base + 320, da + (uint64_t)((5 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id128, 1);
            umma<1, 0>(tmem_base + 384, da + (uint64_t)((6 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id64, 1);
          }
          n_mma += 10;
        } else if (mode == 3) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if ((q & 3) == 0 && el) tmem_cp(a_t, da + (uint64_t)(((q % 7) * 8192 + kk * 32) >> 4));
            if (el) umma_ts(tmem_base + (q & 1) * 224, a_t, db + (uint64_t)((kk * 32) >> 4), idN);
          }
          n_mma += 8;
        } else if (mode == 5) {
          // the stacked S = 7 schedule with the row operand staged in TMEM: one
          // tcgen05.cp per slice and K step (slot 448 + 8 s), MMAs read it there
          if (el) {
#pragma unroll
            for (int sl = 0; sl < 7; ++sl) {
              const uint32_t at = tmem_base + 448 + 8 * sl;
              tmem_cp(at, da + (uint64_t)((sl * 8192 + kk * 32) >> 4));
              const int run = 7 - sl, first = run <= 4 ? run : (run + 1) / 2;
              const uint32_t id1 = (2u << 4) | ((uint32_t)(first * 8) << 17) | (8u << 24);
              umma_ts(tmem_base + sl * 64, at, db + (uint64_t)((kk * 32) >> 4), id1);
              if (run > first) {
                const uint32_t id2 = (2u << 4) | ((uint32_t)((run - first) * 8) << 17) | (8u << 24);
                umma_ts(tmem_base + (sl + first) * 64, at, db + (uint64_t)((first * 4096 + kk * 32) >> 4), id2);
              }
            }
          }
          n_mma += 10;
        } else {
#pragma unroll
          for (int q = 0; q < 7; ++q)
            if (el) umma<1, 0>(tmem_base + (q & 1) * 256, da + (uint64_t)((q * 8192 + kk * 32) >> 4),
                               db + (uint64_t)(((q % 4) * 4096 + kk * 32) >> 4), id256, 1);
          n_mma += 7;
        }
      }
      __syncwarp();
    }
    if (el) commit<1>(done);
    __syncwarp();
    mbar_wait(done, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) {
      out[2 * blockIdx.x] = (unsigned long long)(t1 - t0);
      out[2 * blockIdx.x + 1] = (unsigned long long)n_mma;
    }
  }
  tmem_teardown<1>(tmem_base, warp);
}

static void run_rate2(int N, int mode, const char* what) {
  const int units = 148;
  const int smem = 1024 + 7 * 8192 + 16 * 4096 + 256;
  unsigned long long* out;
  CK(cudaMalloc(&out, sizeof(unsigned long long) * 2 * units));
  CK(cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) {
    rate2_kernel<<<units, 128, smem>>>(N, mode, 300, out);
    CK(cudaDeviceSynchronize());
  }
  std::vector<unsigned long long> h(2 * units);
  CK(cudaMemcpy(h.data(), out, sizeof(unsigned long long) * 2 * units, cudaMemcpyDeviceToHost));
  std::vector<double> per;
  for (int u = 0; u < units; ++u) per.push_back((double)h[2 * u] / (double)h[2 * u + 1]);
  std::sort(per.begin(), per.end());
  printf("rate2 %-44s N=%3d: %.1f cycles per MMA (ideal N/2 = %d)\n", what, N,
         per[per.size() / 2], N / 2);
  cudaFree(out);
}

// ---- contend: what does a busy tensor pipe cost the other warps of the SM? ----
// Warp 0 issues the stacked S = 7 schedule (mode 2 of rate2) back to back while 8
// more warps -- the epilogue warps of gram_oz.cu -- run a loop of one kind of
// instruction, 4 independent chains per thread, and time themselves.
//   work 0 none  1 DFMA  2 FFMA  3 IMAD  4 LDS (broadcast)  5 SHFL  6 STG  7 DMUL+DSETP
template <int work>
__global__ void __launch_bounds__(288, 1)
contend_kernel(int with_mma, int iters, int witers, double* sink,
               unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - raw);
  const uint32_t sa = base, sb = base + 7 * 8192;
  const uint32_t bars = sb + 16 * 4096;
  const uint32_t done = bars;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(g + (bars - base) + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (bars - base) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(g)[i] = 0x01020301u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    mbar_fence_init();
  }
  const uint32_t tmem_base = tmem_setup<1>(bars + 64, slot, warp);
  if (warp == 0) {
    const bool el = elect1();
    const uint64_t da = umma_desc64(sa), db = umma_desc64(sb);
    const uint32_t id64 = (2u << 4) | (8u << 17) | (8u << 24), id128 = (2u << 4) | (16u << 17) | (8u << 24),
                   id192 = (2u << 4) | (24u << 17) | (8u << 24), id256 = (2u << 4) | (32u << 17) | (8u << 24);
    __syncwarp();
    const long long t0 = clock64();
    long long n_mma = 0;
    if (with_mma) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          if (el) {
            umma<1, 0>(tmem_base + 0, da + (uint64_t)((0 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id256, 1);
            umma<1, 0>(tmem_base + 256, da + (uint64_t)((0 * 8192 + kk * 32) >> 4), db + (uint64_t)((4 * 4096 + kk * 32) >> 4), id192, 1);
            umma<1, 0>(tmem_base + 64, da + (uint64_t)((1 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id192, 1);
            umma<1, 0>(tmem_base + 256, da + (uint64_t)((1 * 8192 + kk * 32) >> 4), db + (uint64_t)((3 * 4096 + kk * 32) >> 4), id192, 1);
            umma<1, 0>(tmem_base + 128, da + (uint64_t)((2 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id192, 1);
            umma<1, 0>(tmem_base + 320, da + (uint64_t)((2 * 8192 + kk * 32) >> 4), db + (uint64_t)((3 * 4096 + kk * 32) >> 4), id128, 1);
            umma<1, 0>(tmem_base + 192, da + (uint64_t)((3 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id256, 1);
            umma<1, 0>(tmem_base + 256, da + (uint64_t)((4 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id192, 1);
            umma<1, 0>(tmem_base + 320, da + (uint64_t)((5 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id128, 1);
            umma<1, 0>(tmem_base + 384, da + (uint64_t)((6 * 8192 + kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), id64, 1);
          }
          n_mma += 10;
        }
        __syncwarp();
      }
      if (el) commit<1>(done);
      __syncwarp();
      mbar_wait(done, 0);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) {
      out[4 * blockIdx.x] = (unsigned long long)(t1 - t0);
      out[4 * blockIdx.x + 1] = (unsigned long long)n_mma;
    }
  } else {
    double d0 = 1.0 + lane, d1 = 2.0, d2 = 3.0, d3 = 4.0;
    const double m = 1.0000001, c = 1e-9;
    float f0 = 1.f + lane, f1 = 2.f, f2 = 3.f, f3 = 4.f;
    int i0 = lane, i1 = 2, i2 = 3, i3 = 4;
    double* my = sink + ((size_t)blockIdx.x * 288 + threadIdx.x) * 4;
    const long long t0 = clock64();
    for (int it = 0; it < witers; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (work == 1) {
          d0 = fma(d0, m, c); d1 = fma(d1, m, c); d2 = fma(d2, m, c); d3 = fma(d3, m, c);
        } else if (work == 2) {
          f0 = fmaf(f0, 1.0001f, 1e-5f); f1 = fmaf(f1, 1.0001f, 1e-5f);
          f2 = fmaf(f2, 1.0001f, 1e-5f); f3 = fmaf(f3, 1.0001f, 1e-5f);
        } else if (work == 3) {
          i0 = i0 * 3 + 1; i1 = i1 * 3 + 1; i2 = i2 * 3 + 1; i3 = i3 * 3 + 1;
        } else if (work == 4) {
          unsigned long long x, y;
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(sa + (uint32_t)((u + i0) & 63) * 16));
          i0 += (int)(x & 1);
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(sa + (uint32_t)((u + i1) & 63) * 16));
          i1 += (int)(y & 1);
        } else if (work == 5) {
          i0 = __shfl_xor_sync(0xffffffffu, i0, 16); i1 = __shfl_xor_sync(0xffffffffu, i1, 8);
          i2 = __shfl_xor_sync(0xffffffffu, i2, 4); i3 = __shfl_xor_sync(0xffffffffu, i3, 2);
        } else if (work == 6) {
          __stcs(my + (u & 3), d0);
        } else if (work == 7) {
          d0 = __dmul_rn(d0, m); d1 = fmin(d1, d0); d2 = __dmul_rn(d2, m); d3 = fmin(d3, d2);
        } else if (work == 8) {
          d0 += (double)(long long)i0; d1 += (double)(long long)(i0 + u);
          d2 += (double)(((long long)i0 << 20) + u); d3 += (double)(((long long)i0 << 21) + u);
        } else if (work == 9) {
          d0 = __dadd_rn(d0, m); d1 = __dadd_rn(d1, c); d2 = __dadd_rn(d2, m); d3 = __dadd_rn(d3, c);
        }
      }
    }
    const long long t1 = clock64();
    if (d0 + d1 + d2 + d3 + f0 + f1 + f2 + f3 + i0 + i1 + i2 + i3 == 12345.678) my[0] = 1.0;
    if (threadIdx.x == 32) {
      out[4 * blockIdx.x + 2] = (unsigned long long)(t1 - t0);
      out[4 * blockIdx.x + 3] = (unsigned long long)witers * 8;
    }
  }
  tmem_teardown<1>(tmem_base, warp);
}

template <int work>
static void run_contend(const char* what, int per_round) {
  const int units = 148;
  const int smem = 1024 + 7 * 8192 + 16 * 4096 + 256;
  unsigned long long* out;
  double* sink;
  CK(cudaMalloc(&out, sizeof(unsigned long long) * 4 * units));
  CK(cudaMalloc(&sink, sizeof(double) * 4 * 288 * units));
  CK(cudaFuncSetAttribute(contend_kernel<work>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  double res[2][2];
  for (int with = 0; with < 2; ++with) {
    for (int rep = 0; rep < 2; ++rep) {
      contend_kernel<work><<<units, 288, smem>>>(with, 3000, 4000, sink, out);
      CK(cudaDeviceSynchronize());
    }
    std::vector<unsigned long long> h(4 * units);
    CK(cudaMemcpy(h.data(), out, sizeof(unsigned long long) * 4 * units, cudaMemcpyDeviceToHost));
    std::vector<double> pm, pw;
    for (int u = 0; u < units; ++u) {
      pm.push_back(h[4 * u + 1] ? (double)h[4 * u] / (double)h[4 * u + 1] : 0.0);
      pw.push_back((double)h[4 * u + 2] / (double)h[4 * u + 3]);
    }
    std::sort(pm.begin(), pm.end());
    std::sort(pw.begin(), pw.end());
    res[with][0] = pm[units / 2];
    res[with][1] = pw[units / 2];
  }
  printf("contend %-26s: cycles per round of %d instr/thread (8 warps): alone %.1f, under MMA %.1f "
         "(x%.2f) | cycles per MMA under this work %.1f (alone 90.9)\n",
         what, per_round, res[0][1], res[1][1], res[1][1] / res[0][1], res[1][0]);
  cudaFree(out);
  cudaFree(sink);
}

// ---- check (TS): the row operand is copied shared memory -> TMEM with
// tcgen05.cp.128x256b (one 32-byte K step of the 64B-swizzled tile) and the MMA
// reads it from there; must give the same products as the SS form
__global__ void __launch_bounds__(128, 1)
check_ts_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                int N, uint32_t idesc, int32_t* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - raw);
  const uint32_t sa = base, sb = base + 8192;
  const uint32_t bars = base + 8192 + 16384;
  const uint32_t full = bars, done = bars + 8;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(g + 8192 + 16384 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    mbar_fence_init();
  }
  const uint32_t tmem_base = tmem_setup<1>(bars + 64, slot, warp);
  if (threadIdx.x == 0) {
    mbar_expect_tx(full, 8192 + N * 64);
    tma_load_2d(sa, &map_a, 0, 0, full);
    tma_load_2d(sb, &map_b, 0, 256, full);
    mbar_wait(full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int kk = 0; kk < 2; ++kk) {
      tmem_cp(tmem_base + 480, umma_desc64(sa + kk * 32));
      umma_ts_acc(tmem_base, tmem_base + 480, umma_desc64(sb + kk * 32), idesc, kk != 0);
    }
    commit<1>(done);
  }
  __syncthreads();
  mbar_wait(done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int c = 0; c < 32; ++c) out[row * N + c0 + c] = (int32_t)v[c];
  }
  tmem_teardown<1>(tmem_base, warp);
}

template <int CG>
static int run_check(int N, int a_signed, int b_signed) {
  const int rows = 256 + 256, cols = 64;
  std::vector<uint8_t> h((size_t)rows * cols);
  srand(1234 + N + 2 * a_signed + b_signed);
  for (auto& v : h) v = (uint8_t)(rand() & 0xff);
  uint8_t* d;
  int32_t* out;
  CK(cudaMalloc(&d, h.size()));
  CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&out, sizeof(int32_t) * 256 * N));
  CK(cudaMemset(out, 0xff, sizeof(int32_t) * 256 * N));
  CUtensorMap ma, mb;
  if (encode_u8_map(&ma, d, rows, cols, 128, 64)) return 1;
  if (encode_u8_map(&mb, d, rows, cols, N / CG, 64)) return 1;
  const int smem = 1024 + 8192 + 16384 + 256;
  auto kern = check_kernel<CG>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CG);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = CG;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, kern, ma, mb, N, idesc_i8(128 * CG, N, a_signed, b_signed), out));
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> got((size_t)256 * N);
  CK(cudaMemcpy(got.data(), out, got.size() * 4, cudaMemcpyDeviceToHost));
  long bad = 0;
  const int M = 128 * CG;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      long long acc = 0;
      for (int k = 0; k < cols; ++k) {
        const uint8_t ab = h[(size_t)i * cols + k], bb = h[(size_t)(256 + j) * cols + k];
        const int av = a_signed ? (int)(int8_t)ab : (int)ab;
        const int bv = b_signed ? (int)(int8_t)bb : (int)bb;
        acc += (long long)av * bv;
      }
      if ((long long)got[(size_t)i * N + j] != acc) {
        if (bad < 5)
          printf("  mismatch (%d,%d): got %d want %lld\n", i, j, got[(size_t)i * N + j], acc);
        ++bad;
      }
    }
  printf("check cg=%d N=%d A=%s B=%s: %s (%ld of %d wrong)\n", CG, N, a_signed ? "s8" : "u8",
         b_signed ? "s8" : "u8", bad ? "FAIL" : "ok", bad, M * N);
  cudaFree(d);
  cudaFree(out);
  return bad != 0;
}

static int run_check_ts(int N) {
  const int rows = 256 + 256, cols = 64;
  std::vector<uint8_t> h((size_t)rows * cols);
  srand(99 + N);
  for (auto& v : h) v = (uint8_t)(rand() & 0xff);
  uint8_t* d;
  int32_t* out;
  CK(cudaMalloc(&d, h.size()));
  CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&out, sizeof(int32_t) * 128 * N));
  CK(cudaMemset(out, 0xff, sizeof(int32_t) * 128 * N));
  CUtensorMap ma, mb;
  if (encode_u8_map(&ma, d, rows, cols, 128, 64)) return 1;
  if (encode_u8_map(&mb, d, rows, cols, N, 64)) return 1;
  const int smem = 1024 + 8192 + 16384 + 256;
  CK(cudaFuncSetAttribute(check_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  check_ts_kernel<<<1, 128, smem>>>(ma, mb, N, idesc_i8(128, N, 0, 0), out);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> got((size_t)128 * N);
  CK(cudaMemcpy(got.data(), out, got.size() * 4, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < N; ++j) {
      long long acc = 0;
      for (int k = 0; k < cols; ++k)
        acc += (long long)h[(size_t)i * cols + k] * (long long)h[(size_t)(256 + j) * cols + k];
      if ((long long)got[(size_t)i * N + j] != acc) {
        if (bad < 5) printf("  TS mismatch (%d,%d): got %d want %lld\n", i, j, got[(size_t)i * N + j], acc);
        ++bad;
      }
    }
  printf("check TS (A via tcgen05.cp -> TMEM) N=%d u8 x u8: %s (%ld of %d wrong)\n", N,
         bad ? "FAIL" : "ok", bad, 128 * N);
  cudaFree(d);
  cudaFree(out);
  return bad != 0;
}

template <int CG, int KIND>
static void run_rate(int N, int S, int mode) {
  const int units = 148 / CG;
  const int nb = N / CG;
  const int smem = 1024 + S * 8192 + S * ((nb * 64 + 1023) & ~1023) + 256;
  if (smem > 227 * 1024) return;
  if (mode == 0 && S * N > 512) return;
  if (mode == 1 && 2 * N > 512) return;
  unsigned long long* out;
  CK(cudaMalloc(&out, sizeof(unsigned long long) * 2 * units));
  auto kern = rate_kernel<CG, KIND>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(units * CG);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = CG;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  const uint32_t idesc = KIND == 0 ? idesc_i8(128 * CG, N, 1, 0) : idesc_f16(128 * CG, N);
  const int iters = 200;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaLaunchKernelEx(&cfg, kern, N, S, iters, mode, idesc, out));
    CK(cudaDeviceSynchronize());
  }
  std::vector<unsigned long long> h(2 * units);
  CK(cudaMemcpy(h.data(), out, sizeof(unsigned long long) * 2 * units, cudaMemcpyDeviceToHost));
  std::vector<double> per;
  for (int u = 0; u < units; ++u) per.push_back((double)h[2 * u] / (double)h[2 * u + 1]);
  std::sort(per.begin(), per.end());
  const double med = per[per.size() / 2];
  // MACs per cycle per SM: each SM contracts 128 x N x (32 bytes of K) per MMA
  const double k_el = KIND == 0 ? 32.0 : 16.0;
  printf("rate kind=%s cg=%d N=%3d S=%d mode=%d: %.1f cyc/MMA (min %.1f max %.1f; floor N/2 = %d) "
         "-> %.0f MAC/cyc/SM, smem operand bytes/cyc/SM %.0f\n",
         KIND == 0 ? "i8 " : "f16", CG, N, S, mode, med, per.front(), per.back(), N / 2,
         128.0 * N * k_el / med, (128.0 * 32 + nb * 32.0) / med);
  cudaFree(out);
}

// ---- sustained: the stacked S = 7 schedule from RESIDENT operands for ~0.2 s on all
// SMs -- what the tensor pipe alone delivers under the board's power cap.
static void run_sustained() {
  const int units = 148;
  const int smem = 1024 + 7 * 8192 + 16 * 4096 + 256;
  unsigned long long* out;
  CK(cudaMalloc(&out, sizeof(unsigned long long) * 2 * units));
  CK(cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int rnd = 0; rnd < 2; ++rnd)
  for (int iters : {300, 20000, 150000}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      rate2_kernel<<<units, 128, smem>>>(0, 2, iters, out, rnd);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::vector<unsigned long long> h(2 * units);
      CK(cudaMemcpy(h.data(), out, sizeof(unsigned long long) * 2 * units, cudaMemcpyDeviceToHost));
      const double ksteps = (double)iters * 2.0;    // 32-byte K steps (10 MMAs each)
      const double cyc = (double)h[0] / ksteps;
      // executed INT8 MACs per K step per SM: 28 products x 128 x 64 x 32
      const double tops = 2.0 * 28 * 128 * 64 * 32 * ksteps * units / (ms * 1e-3) * 1e-12;
      printf("sustained S=7 schedule, %s operands, %6d iterations: %8.3f ms, %.1f cycles per K step, "
             "%.0f ns per K step, SM clock %.0f MHz, %.0f TOP/s executed\n",
             rnd ? "random-byte" : "constant", iters, ms, cyc, ms * 1e6 / ksteps,
             cyc / (ms * 1e6 / ksteps) * 1e3, tops);
    }
  }
  cudaFree(out);
}

static void run_contend_all() {
  run_contend<0>("no work", 0);
  run_contend<1>("DFMA", 4);
  run_contend<9>("DADD", 4);
  run_contend<7>("2 DMUL + 2 fmin", 4);
  run_contend<8>("4 x (I2F.F64.S64 + DADD)", 8);
  run_contend<2>("FFMA", 4);
  run_contend<3>("IMAD", 4);
  run_contend<4>("LDS.128 broadcast", 2);
  run_contend<5>("SHFL", 4);
  run_contend<6>("STG.64 (.cs)", 1);
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  if (argc > 1 && std::string(argv[1]) == "sustained") {
    run_sustained();
    return 0;
  }
  if (argc > 1 && std::string(argv[1]) == "contend") {
    run_contend_all();
    return 0;
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device: %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  int fails = 0;
  for (int as = 0; as < 2; ++as)
    for (int bs = 0; bs < 2; ++bs) {
      fails += run_check<1>(64, as, bs);
      fails += run_check<2>(64, as, bs);
    }
  fails += run_check<1>(128, 1, 0);
  fails += run_check<2>(128, 1, 0);
  fails += run_check<1>(256, 1, 0);
  fails += run_check<2>(256, 1, 0);
  fails += run_check_ts(64);
  fails += run_check_ts(256);
  fails += run_check<1>(32, 1, 0);
  fails += run_check<2>(32, 1, 0);
  const int Ns[] = {32, 64, 96, 128, 160, 256};
  for (int N : Ns) {
    run_rate<1, 0>(N, 7, 0);
    run_rate<1, 0>(N, 6, 0);
    run_rate<2, 0>(N, 7, 0);
    run_rate<2, 0>(N, 6, 0);
    run_rate<1, 0>(N, 4, 1);
    run_rate<2, 0>(N, 4, 1);
    run_rate<1, 1>(N, 4, 1);
    run_rate<2, 1>(N, 4, 1);
  }
  for (int N : {64, 128, 192, 256}) run_rate2(N, 0, "SS (A, B from shared memory)");
  for (int N : {64, 128, 192, 256}) run_rate2(N, 1, "TS (A from TMEM)");
  for (int N : {64, 128, 192, 256}) run_rate2(N, 3, "TS + tcgen05.cp 128x256b per 4 MMAs");
  run_rate2(0, 2, "SS, stacked S = 7 schedule (10 per K step)");
  run_rate2(256, 4, "SS, all N = 256 (7 per K step)");
  run_rate2(0, 5, "TS stacked S = 7 schedule + 7 tcgen05.cp / K step");
  run_contend_all();
  // three accumulators (two-pass split): S = 3 pattern at N = 128 / 160
  run_rate<1, 0>(128, 3, 0);
  run_rate<2, 0>(128, 3, 0);
  run_rate<1, 0>(160, 3, 0);
  run_rate<2, 0>(160, 3, 0);
  printf("probe done: %d check failures\n", fails);
  return fails ? 1 : 0;
}
