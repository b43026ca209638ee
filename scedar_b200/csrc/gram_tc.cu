// K2 (fast path) -- Gram contraction on the 5th-generation tensor cores:
// tcgen05.mma (kind::f16, FP32 accumulators in TMEM) fed by TMA into
// 128B-swizzled shared memory, with the metric epilogue and the streaming
// top-k fused.  Selected by SCEDAR_PRECISION_FAST.
//
// Precision scheme ("3xFP16"): every operand is split into two halves
// x = hi + lo (|x - hi - lo| <= 2^-22 |x|); the contraction accumulates
// hi.hi + hi.lo + lo.hi in FP32, i.e. a product of FP32 quality from three
// tensor-core passes over the same K slab (the north star's split-precision
// fast path; FP16 halves rather than BF16 because the operands are scaled into
// FP16's range and need mantissa, not exponent).
//
// Operands.  The row tile (A) and the column tile (B) are different splits of
// the same cells, so that the accumulator IS the ranking key / the metric:
//   cosine, correlation   A = u,            B = -u            acc = -cos
//   euclidean             A = [s y, 1],     B = [-2 s y, s^2|y|^2]
//                                           acc = s^2 (|y_c|^2 - 2 y_r.y_c)
// (u = unit-length row; y = x - column means: distances do not change under a
// translation, and cells far from the origin would otherwise spend the FP16
// mantissa on their common offset -- blobs shifted to be non-negative made
// 68,579 x 2,000 fall back to FP64 almost everywhere; s = a power of two that
// brings max |y|^2 under 2^15; the extra K column is one of the padding
// columns of the K slab).
// The epilogue therefore needs no per-column loads: d^2 = acc / s^2 + |x_r|^2,
// d = 1 + acc, and for the top-k a bare compare of the accumulator against
// the row's threshold.
//
// Stated tolerance of the fast distance matrix (tests assert it): cosine and
// correlation |D_fast - D_fp64| <= 1e-5; euclidean |D_fast^2 - D_fp64^2| <=
// 2e-6 * max |x|^2, i.e. |D_fast - D_fp64| <= 1e-4 * max(D) except between
// near-duplicate cells, where the square root turns a 1e-6 error of d^2 into
// up to 1.5e-3 * max |x| of d.  The fast kNN is EXACT: the
// tensor-core pass only proposes candidates (32 per cell and column split),
// which are re-ranked with float64 distances computed like the FP64 path and
// verified against the candidate threshold with a margin that is the larger of
// 8 x the error observed on the re-ranked candidates and an a-priori bound of
// the FP32 accumulation (margin_kernel); cells whose margin is too small are
// recomputed by the FP64 kernel.
//
// Structure (one CTA per SM, 192 threads):
//   warp 0   TMA producer: per 32-wide K slab 4 boxes (A_hi, A_lo: 128 rows;
//            B_hi, B_lo: 256 rows, 64-byte swizzle) -> 48 KB stage, 4 stages,
//            mbarrier full/empty
//   warp 1   MMA issuer (one elected lane): per slab 2 k-steps x 3 products of
//            UMMA 128 x 256 x 16 into one of two 256-column TMEM accumulators;
//            tcgen05.commit frees the stage / publishes the accumulator
//   CG = 2   (rows of >= 256 K columns) a cluster of two CTAs owns two row
//            tiles: tcgen05.mma.cta_group::2 (M = 256) issued by the leader
//            CTA reads each CTA's A rows and HALF of the B tile from each CTA's
//            shared memory (32 KB stage, 6 stages); both producers' TMA bytes
//            are counted on the leader's full barrier
//            (cp.async.bulk.tensor.cta_group::2), the commits are multicast
//            to both CTAs' empty / accumulator-full barriers, and both
//            epilogues arrive on the leader's accumulator-empty barrier
//   warps 2-5 epilogue, one thread per tile row, TMEM lane quarter = warp % 4:
//            STORE  tcgen05.ld 32 columns -> metric in float64 -> coalesced
//                   store of the D tile through shared memory
//            TOPK   double-buffered tcgen05.ld of 32 columns; the hot loop is
//                   one FSETP per candidate (accumulator < row threshold,
//                   OR-reduced) and one vote per 32 x 32 block; a hit (rare
//                   once the lists are warm) re-reads its column from TMEM and
//                   is inserted warp-cooperatively (ballot + shuffle) into the
//                   row's sorted 32-entry list in shared memory
// The CTA owns a 128-row tile and a range of 256-column tiles, so per-row
// candidate lists never leave the SM.
//
// Replaces the same reference arithmetic as gram_f64.cu (np.dot, cosine /
// correlation scaling, sklearn euclidean, scedar/eda/sdm.py:1216, 1245-1266).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "topk_list.cuh"

namespace scedar {
namespace {

constexpr int TM = 128, TN = 256, TK = 32;
constexpr int A_BYTES = TM * TK * 2;
// CG = CTAs per MMA (tcgen05 cta_group): with CG = 2 a cluster of two CTAs owns
// two row tiles and each CTA stages only half of the B tile
constexpr int b_bytes(int cg) { return TN / cg * TK * 2; }
constexpr int stage_bytes(int cg) { return 2 * A_BYTES + 2 * b_bytes(cg); }  // 48 KB / 32 KB
constexpr int n_stages(int cg) { return cg == 1 ? 4 : 6; }                   // 192 KB ring
constexpr int KC = 32;       // candidates kept per row and column split
constexpr int STG_LD = 33;   // staging row stride of the STORE epilogue (conflict-free)
constexpr int EPI_BYTES = TM * STG_LD * 8;   // == 2 * TM * (KC + 1) * 4 (the TOPK lists)
constexpr int BAR_BYTES = 256;
constexpr int TC_SMEM = 1024 + n_stages(1) * stage_bytes(1) + EPI_BYTES + BAR_BYTES;
static_assert(n_stages(1) * stage_bytes(1) == n_stages(2) * stage_bytes(2), "one ring size");
constexpr int TC_THREADS = 192;       // producer + MMA + one epilogue group
constexpr int TC_THREADS_AR = 320;    // ... + a second epilogue group (A-resident mode)
constexpr int RR_CH = 64, RR_LD = 65;   // re-rank: genes per chunk, padded row stride
constexpr int RR_WARPS = 4;
constexpr int RR_SMEM = RR_WARPS * (32 * RR_LD + RR_CH) * 8;
constexpr uint32_t idesc(int cg) {
  return (1u << 4)                              // accumulator FP32
         | ((uint32_t)(TN >> 3) << 17)          // N
         | ((uint32_t)((TM * cg) >> 4) << 24);  // M (both CTAs' rows); A,B = F16 K-major
}
constexpr unsigned FULL = 0xffffffffu;

struct TcArgs {
  const double* stat;   // sqrt(1/|x|^2) or |x|^2 (euclid), float64
  const double* scale;  // device: [0] = s, [1] = 1/s^2 (both 1 unless euclid)
  const double* n2c;    // euclid: |x - mean|^2 of every cell (centred copy), float64
  int64_t n;
  int kblocks;
  int64_t row_begin, row_end;
  int euclid;
  int splits;           // column splits S; grid = groups * S persistent CTAs
  int groups;           // CTA (g, s) walks row tiles g, g + groups, ...
  int row_tiles;
  double* out;          // STORE: D stripe
  int64_t ldd;
  int sym;              // STORE over the whole matrix: column tiles wholly above
                        // the diagonal are skipped and written as transposes
  int kc;               // TOPK: list entries that count (<= KC): the row threshold is
                        // entry kc - 1; shorter lists for small k mean fewer inserts
  float* part_key;      // TOPK: [splits, rows, KC]
  int32_t* part_idx;
};

// K-major operand tile in 64B-swizzled shared memory: rows of 64 bytes (one
// K slab of 32 halves), 8-row swizzle atoms 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(512 >> 4) << 32;     // stride byte offset between 8-row atoms
  d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
  d |= (uint64_t)4 << 61;              // SWIZZLE_64B
  return d;
}
template <int CG>
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t acc) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc(1)), "r"(acc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc(2)), "r"(acc)
        : "memory");
}
// arrive on `bar` when the MMAs issued so far retire; with CG = 2 on the
// barrier at the same offset in both CTAs of the pair
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     bar)
                 : "memory");
  else
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(bar),
        "h"((uint16_t)3)
        : "memory");
}
// one lane of a converged warp (the whole warp runs the MMA / TMA loops so
// that ptxas keeps the descriptor arithmetic in uniform registers; only the
// asynchronous instructions themselves are issued by the elected lane)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `addr` (a shared::cta address) in CTA `rank`
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;"
               ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
// CG = 2 TMA load: the bytes are counted on the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0,
                                                 int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// 32 lanes x 32 columns -> 32 registers per thread; NOT complete until
// tmem_wait() has been executed on the same registers
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// the registers are in/out operands so that no use of them can be scheduled
// above the wait
__device__ __forceinline__ void tmem_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]),
                 "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]),
                 "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]),
                 "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]),
                 "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  tmem_ld32_async(taddr, v);
  tmem_wait(v);
}
// one column (dynamic address) of the warp's 32 lanes, complete on return
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(v)
      : "r"(taddr)
      : "memory");
  return v;
}

constexpr int LKC = KC + 1;   // list row stride in shared memory (conflict-free both ways)

// TOPK: one 32-row x 32-column block of accumulators against the per-row
// thresholds.  Hot path: 32 compares OR-ed into one predicate + one vote.
// A column with hits re-reads its 32 accumulators from TMEM and the hits are
// inserted one by one by the whole warp (ballot + shuffle).  (A per-lane
// insertion path for cold lists was measured and was slower.)
__device__ __forceinline__ void scan_block(const uint32_t (&v)[32], uint32_t taddr, int64_t jbase,
                                           int64_t n, float& tau, float* lkey, int* lidx,
                                           int row0, int lane, int kc) {
  bool any = false;
#pragma unroll
  for (int c = 0; c < 32; ++c) any |= __uint_as_float(v[c]) < tau;
  if (!__any_sync(FULL, any)) return;
  unsigned hm = 0;
#pragma unroll
  for (int c = 0; c < 32; ++c) hm |= (__uint_as_float(v[c]) < tau) ? (1u << c) : 0u;
  unsigned cm = __reduce_or_sync(FULL, hm);   // columns with a hit in some row
  while (cm) {
    const int c = __ffs(cm) - 1;
    cm &= cm - 1;
    const float key = __uint_as_float(tmem_ld1(taddr + c));
    const int64_t gj = jbase + c;
    const int nj = (int)gj;
    const bool hit = key < tau && gj < n;
    unsigned m = __ballot_sync(FULL, hit);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float nk = __shfl_sync(FULL, key, src);
      // warp-cooperative insert into the sorted list of row row0 + src
      float* lk_row = lkey + (row0 + src) * LKC;
      int* li_row = lidx + (row0 + src) * LKC;
      float lk = lk_row[lane];
      int li = li_row[lane];
      const int pos = __popc(__ballot_sync(FULL, lk < nk || (lk == nk && li < nj)));
      const float pk = __shfl_up_sync(FULL, lk, 1);
      const int pi = __shfl_up_sync(FULL, li, 1);
      if (lane > pos) {
        lk = pk;
        li = pi;
      } else if (lane == pos) {
        lk = nk;
        li = nj;
      }
      if (lane >= pos) {
        lk_row[lane] = lk;
        li_row[lane] = li;
      }
      const float worst = __shfl_sync(FULL, lk, kc - 1);
      if (lane == src) tau = worst;
    }
    __syncwarp();
  }
}

// AR ("A resident", CG = 1, at most two K slabs -- e.g. 50 PCs): a tile is then
// only 12 MMAs, and the single epilogue group (one warp per scheduler, a serial
// chain of TMEM loads and compares) is what bounds the tile rate.  In this mode
// the row tile's A panel is loaded once and stays in shared memory, the ring
// carries only B (3 stages of 32 KB), and the freed shared memory holds a
// second set of candidate lists for a SECOND epilogue group: group g drains
// accumulator g (every other tile), so each has two MMA periods per tile.
template <bool TOPK, int CG, bool AR>
__global__ void __launch_bounds__(AR ? TC_THREADS_AR : TC_THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi,
               const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi,
               const __grid_constant__ CUtensorMap map_b_lo, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  static_assert(!(AR && CG == 2), "A-resident mode is single-CTA");
  constexpr int B_BYTES = b_bytes(CG);
  constexpr int EG = AR ? 2 : 1;                   // epilogue groups
  constexpr int TSTAGES = AR ? 3 : n_stages(CG);
  constexpr int STAGE_BYTES = AR ? 2 * B_BYTES : stage_bytes(CG);
  constexpr int RING_OFF = AR ? 4 * A_BYTES : 0;   // resident A: 2 slabs x (hi, lo)
  constexpr int RING_BYTES = AR ? RING_OFF + TSTAGES * STAGE_BYTES : n_stages(1) * stage_bytes(1);
  static_assert(1024 + RING_BYTES + EG * EPI_BYTES + BAR_BYTES <= TC_SMEM, "smem budget");
  uint8_t* epi = gbase + RING_BYTES;
  const uint32_t bars = base + RING_BYTES + EG * EPI_BYTES;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;
  // barrier slots (8 bytes each): full[<=6], empty[<=6], tfull[2], tempty[2], afull
  const uint32_t full0 = bars, empty0 = bars + 48, tfull0 = bars + 96, tempty0 = bars + 112,
                 afull = bars + 128;
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(epi + EG * EPI_BYTES + 144);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull0 + 8 * s, 1);
      mbar_init(tempty0 + 8 * s, 128 * CG);   // the epilogue threads of the whole pair
    }
    mbar_init(afull, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(const_cast<uint32_t*>(tmem_slot))),
                   "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(const_cast<uint32_t*>(tmem_slot))),
                   "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();   // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // 1-D grid of groups x splits CTAs, split fastest: CTA (g, s) sweeps the
  // column tiles of split s for row tiles g, g + groups, ... (see tc_schedule;
  // with groups == row_tiles every CTA owns one row tile and the CTAs resident
  // at one time cover few row tiles, so their A panels stay in L2).
  const int unit = (int)(blockIdx.x / CG);     // a CTA (CG = 1) or a CTA pair
  const int split = unit % a.splits;
  const int group = unit / a.splits;
  const int64_t rt0 = a.row_begin / TM;
  const int64_t CT = (a.n + TN - 1) / TN;
  const int KB = a.kblocks;
  // column tiles of row tile rt handled by this CTA: [J0, J1).  Symmetric
  // STORE: only the tiles that touch the lower triangle (256 J < 128 (I + 1)),
  // split evenly; the heavy (bottom) row tiles are scheduled first.
  // a.row_tiles counts row units: CG row tiles each
  auto row_tile_of = [&](int rt) -> int64_t {
    return rt0 + (int64_t)CG * (a.sym ? a.row_tiles - 1 - rt : rt) + cta_rank;
  };
  auto j_range = [&](int64_t I, int64_t& J0, int64_t& J1) {
    int64_t L = CT;
    if (a.sym) {
      const int64_t ja = (I + 2) / 2;
      L = ja < CT ? ja : CT;
    }
    J0 = L * split / a.splits;
    J1 = L * (split + 1) / a.splits;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int rt = group; rt < a.row_tiles; rt += a.groups) {
        const int64_t I = row_tile_of(rt);
        const int i0 = (int)(I * TM);
        int64_t J0, J1;
        j_range(I, J0, J1);
        if (AR) {
          // one row tile per CTA in this mode (host: groups == row_tiles)
          mbar_expect_tx(afull, KB * 2 * A_BYTES);
          for (int kb = 0; kb < KB; ++kb) {
            tma_load_2d(base + kb * 2 * A_BYTES, &map_a_hi, kb * TK, i0, afull);
            tma_load_2d(base + kb * 2 * A_BYTES + A_BYTES, &map_a_lo, kb * TK, i0, afull);
          }
        }
        for (int64_t J = J0; J < J1; ++J) {
          for (int kb = 0; kb < KB; ++kb, ++it) {
            const uint32_t s = it % TSTAGES, ph = (it / TSTAGES) & 1;
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            const uint32_t st = base + RING_OFF + s * STAGE_BYTES;
            if (AR) {
              mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
              tma_load_2d(st, &map_b_hi, kb * TK, (int)(J * TN), full0 + 8 * s);
              tma_load_2d(st + B_BYTES, &map_b_lo, kb * TK, (int)(J * TN), full0 + 8 * s);
            } else if (CG == 1) {
              mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
              tma_load_2d(st, &map_a_hi, kb * TK, i0, full0 + 8 * s);
              tma_load_2d(st + A_BYTES, &map_a_lo, kb * TK, i0, full0 + 8 * s);
              tma_load_2d(st + 2 * A_BYTES, &map_b_hi, kb * TK, (int)(J * TN), full0 + 8 * s);
              tma_load_2d(st + 2 * A_BYTES + B_BYTES, &map_b_lo, kb * TK, (int)(J * TN),
                          full0 + 8 * s);
            } else {
              // the leader's barrier counts both CTAs' bytes; each CTA brings its
              // own A rows and its half of the B tile
              if (leader) mbar_expect_tx(full0 + 8 * s, 2 * STAGE_BYTES);
              const uint32_t lbar = map_to_cta(full0 + 8 * s, 0);
              const int jb = (int)(J * TN) + (int)cta_rank * (TN / 2);
              tma_load_2d_pair(st, &map_a_hi, kb * TK, i0, lbar);
              tma_load_2d_pair(st + A_BYTES, &map_a_lo, kb * TK, i0, lbar);
              tma_load_2d_pair(st + 2 * A_BYTES, &map_b_hi, kb * TK, jb, lbar);
              tma_load_2d_pair(st + 2 * A_BYTES + B_BYTES, &map_b_lo, kb * TK, jb, lbar);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      const bool el = elect_one();
      uint32_t it = 0, t = 0;
      int64_t tiles = 0;
      for (int rt = group; rt < a.row_tiles; rt += a.groups) {
        int64_t J0, J1;
        j_range(row_tile_of(rt), J0, J1);
        tiles += J1 - J0;
      }
      if (AR && tiles > 0) mbar_wait(afull, 0);
      for (int64_t tt = 0; tt < tiles; ++tt, ++t) {
        const uint32_t acc = t & 1, aph = (t >> 1) & 1;
        mbar_wait(tempty0 + 8 * acc, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const uint32_t s = it % TSTAGES, ph = (it / TSTAGES) & 1;
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t st = base + RING_OFF + s * STAGE_BYTES;
          const uint32_t sa = AR ? base + kb * 2 * A_BYTES : st;       // A_hi, then A_lo
          const uint32_t sb = AR ? st : st + 2 * A_BYTES;              // B_hi, then B_lo
#pragma unroll
          for (int kk = 0; kk < TK / 16; ++kk) {
            const uint64_t ah = umma_desc(sa + kk * 32);
            const uint64_t al = umma_desc(sa + A_BYTES + kk * 32);
            const uint64_t bh = umma_desc(sb + kk * 32);
            const uint64_t bl = umma_desc(sb + B_BYTES + kk * 32);
            if (el) {
              umma_f16<CG>(d_tmem, ah, bh, (kb | kk) != 0);
              umma_f16<CG>(d_tmem, ah, bl, 1);
              umma_f16<CG>(d_tmem, al, bh, 1);
            }
          }
          if (el) umma_commit<CG>(empty0 + 8 * s);   // stage reusable once these MMAs retire
          __syncwarp();
        }
        if (el) umma_commit<CG>(tfull0 + 8 * acc);   // accumulator complete
        __syncwarp();
      }
    }
  } else {
    // ---- epilogue warps 2..5: TMEM lane quarter q, one thread per tile row --
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int eg = (warp - 2) >> 2;         // epilogue group: drains accumulator eg when EG == 2
    const int e = ((warp - 2) & 3) * 32 + lane;   // 0..127 inside the epilogue group
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t t = 0;
    for (int rt = group; rt < a.row_tiles; rt += a.groups) {
    const int64_t I = row_tile_of(rt);
    const int64_t i0 = I * TM;
    int64_t J0, J1;
    j_range(I, J0, J1);
    const int64_t gi = i0 + r;
    const bool row_ok = gi >= a.row_begin && gi < a.row_end;
    if (TOPK) {
      float* lkey = reinterpret_cast<float*>(epi + eg * EPI_BYTES);               // [TM][LKC]
      int* lidx = reinterpret_cast<int*>(epi + eg * EPI_BYTES + TM * LKC * 4);     // [TM][LKC]
      for (int rr = 0; rr < 32; ++rr) {
        lkey[(q * 32 + rr) * LKC + lane] = __int_as_float(0x7f800000);
        lidx[(q * 32 + rr) * LKC + lane] = 0x7fffffff;
      }
      __syncwarp();
      // rows outside the stripe never hit
      float tau = row_ok ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
      for (int64_t J = J0; J < J1; ++J, ++t) {
        const uint32_t acc = t & 1, aph = (t >> 1) & 1;
        if (EG == 2 && (int)acc != eg) continue;   // the other group's tile
        mbar_wait(tfull0 + 8 * acc, aph);
        tc_fence_after();
        const int64_t j0 = J * TN;
        const uint32_t tile = lane_base + acc * TN;
        uint32_t v0[32], v1[32];
        tmem_ld32(tile, v0);
#pragma unroll
        for (int ch = 0; ch < TN / 32; ch += 2) {
          tmem_ld32_async(tile + (ch + 1) * 32, v1);
          scan_block(v0, tile + ch * 32, j0 + ch * 32, a.n, tau, lkey, lidx, q * 32, lane, a.kc);
          tmem_wait(v1);
          if (ch + 2 < TN / 32) tmem_ld32_async(tile + (ch + 2) * 32, v0);
          scan_block(v1, tile + (ch + 1) * 32, j0 + (ch + 1) * 32, a.n, tau, lkey, lidx, q * 32,
                     lane, a.kc);
          if (ch + 2 < TN / 32) tmem_wait(v0);
        }
        tc_fence_before();
        if (CG == 1)
          mbar_arrive(tempty0 + 8 * acc);
        else
          mbar_arrive_cluster(map_to_cta(tempty0 + 8 * acc, 0));   // the leader issues the MMAs
      }
      __syncwarp();
      const int64_t rows = a.row_end - a.row_begin;
      for (int rr = 0; rr < 32; ++rr) {
        const int64_t g2 = i0 + q * 32 + rr;
        if (g2 < a.row_begin || g2 >= a.row_end) continue;
        const int64_t o = ((int64_t)(split * EG + eg) * rows + (g2 - a.row_begin)) * KC;
        // entries beyond kc were kept sorted but never thresholded: not a
        // complete set, so they are not handed on
        const bool keep = lane < a.kc;
        a.part_key[o + lane] = keep ? lkey[(q * 32 + rr) * LKC + lane] : __int_as_float(0x7f800000);
        a.part_idx[o + lane] = keep ? lidx[(q * 32 + rr) * LKC + lane] : 0x7fffffff;
      }
    } else {
      double* stg = reinterpret_cast<double*>(epi);
      const double inv_s2 = a.scale[1];
      const double n2_row = (a.euclid && gi < a.n) ? a.n2c[gi] : 0.0;
      for (int64_t J = J0; J < J1; ++J, ++t) {
        const uint32_t acc = t & 1, aph = (t >> 1) & 1;
        mbar_wait(tfull0 + 8 * acc, aph);
        tc_fence_after();
        const int64_t j0 = J * TN;
        for (int c0 = 0; c0 < TN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(lane_base + acc * TN + c0, v);
          // metric in float64 from the FP32 accumulator, staged for coalescing
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int64_t gj = j0 + c0 + c;
            const double pv = (double)__uint_as_float(v[c]);
            double d;
            if (a.euclid) {
              double t2 = fma(pv, inv_s2, n2_row);
              t2 = t2 < 0.0 ? 0.0 : t2;
              d = sqrt(t2);
            } else {
              d = 1.0 + pv;
              d = d < 0.0 ? 0.0 : d;
            }
            stg[r * STG_LD + c] = gi == gj ? 0.0 : d;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int rr = 0; rr < 32; ++rr) {
            const int row = rr * 4 + (e >> 5), col = e & 31;
            const int64_t gi2 = i0 + row, gj = j0 + c0 + col;
            if (gi2 >= a.row_begin && gi2 < a.row_end && gj < a.n && (!a.sym || gi2 >= gj))
              a.out[(gi2 - a.row_begin) * a.ldd + gj] = stg[row * STG_LD + col];
          }
          if (a.sym) {
            // transposed copy: column gj of the block is row gj of D, entries
            // strictly below the diagonal only (1 KB contiguous per column)
            const int64_t gi2 = i0 + e;
            for (int col = 0; col < 32; ++col) {
              const int64_t gj = j0 + c0 + col;
              if (gj < gi2 && gi2 < a.n) a.out[gj * a.ldd + gi2] = stg[e * STG_LD + col];
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        tc_fence_before();
        if (CG == 1)
          mbar_arrive(tempty0 + 8 * acc);
        else
          mbar_arrive_cluster(map_to_cta(tempty0 + 8 * acc, 0));   // the leader issues the MMAs
      }
    }
    if (TOPK) __syncwarp();   // the lists are re-initialised for the next row tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();   // nobody leaves while the peer may still read its smem / TMEM
  if (warp == 1) {
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(512)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(512)
                   : "memory");
  }
}

// euclid: column means of the working copy, deterministic two-stage sum.
// part: [CM_CHUNKS, p_pad]
constexpr int CM_CHUNKS = 64;
__global__ void __launch_bounds__(256)
colsum_kernel(const double* __restrict__ xw, int64_t n, int64_t p_pad,
              double* __restrict__ part) {
  const int64_t c = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;                 // 8 row lanes
  const int64_t rows_per = (n + CM_CHUNKS - 1) / CM_CHUNKS;
  const int64_t r0 = blockIdx.y * rows_per, r1 = r0 + rows_per < n ? r0 + rows_per : n;
  double acc = 0.0;
  if (c < p_pad)
    for (int64_t r = r0 + ty; r < r1; r += 8) acc += xw[r * p_pad + c];
  __shared__ double sm[8][33];
  sm[ty][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ty == 0 && c < p_pad) {
    double t = 0.0;
    for (int q = 0; q < 8; ++q) t += sm[q][threadIdx.x];
    part[(int64_t)blockIdx.y * p_pad + c] = t;
  }
}
__global__ void __launch_bounds__(256)
colmean_kernel(const double* __restrict__ part, int64_t n, int64_t p_pad,
               double* __restrict__ mu) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p_pad) return;
  double t = 0.0;
  for (int q = 0; q < CM_CHUNKS; ++q) t += part[(int64_t)q * p_pad + c];
  mu[c] = n > 0 ? t / (double)n : 0.0;
}
// |x - mu|^2 of every cell (one warp per cell)
__global__ void __launch_bounds__(256)
n2c_kernel(const double* __restrict__ xw, const double* __restrict__ mu, int64_t n,
           int64_t n256, int64_t p, int64_t p_pad, double* __restrict__ n2c) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n256) return;
  double acc = 0.0;
  if (row < n)
    for (int64_t c = lane; c < p; c += 32) {
      const double v = xw[row * p_pad + c] - mu[c];
      acc = fma(v, v, acc);
    }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
  if (lane == 0) n2c[row] = acc;
}

// euclid: s = 2^h with max |y|^2 s^2 < 2^15 (FP16 holds the norm column and
// every element); scale[0] = s, scale[1] = 1 / s^2.  One block.
__global__ void __launch_bounds__(1024)
fast_scale_kernel(const double* __restrict__ stat, int64_t n, int euclid,
                  double* __restrict__ scale) {
  __shared__ double part[32];
  double m = 0.0;
  if (euclid)
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, stat[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 32; ++w) m = fmax(m, part[w]);
    int h = 0;
    if (euclid && m > 0.0 && m < 1e300) {
      int ex;
      frexp(m, &ex);  // m < 2^ex
      const int room = 15 - ex;
      h = room >= 0 ? room / 2 : -((-room + 1) / 2);
    }
    scale[0] = ldexp(1.0, h);
    scale[1] = ldexp(1.0, -2 * h);
  }
}

__device__ __forceinline__ void split_half(double v, __half& h, __half& l) {
  h = __double2half(v);
  l = __double2half(v - (double)__half2float(h));
}

// the A- and B-role splits of every cell: ah, al, bh, bl [n256, p64]
__global__ void __launch_bounds__(256)
split_kernel(const double* __restrict__ xw, const double* __restrict__ stat,
             const double* __restrict__ scale, const double* __restrict__ mu,
             const double* __restrict__ n2c, int64_t n, int64_t p, int64_t p_pad,
             int64_t p64, int euclid, __half* __restrict__ ah, __half* __restrict__ al,
             __half* __restrict__ bh, __half* __restrict__ bl) {
  const int64_t row = blockIdx.x;
  const bool real = row < n;
  double mul = 0.0, n2s = 0.0;
  if (real) {
    if (euclid) {
      mul = scale[0];
      n2s = n2c[row] * scale[0] * scale[0];
    } else {
      mul = stat[row];   // sqrt(1/|x|^2): unit length (0 for a zero vector)
    }
  }
  const double bmul = euclid ? -2.0 : -1.0;
  for (int64_t c = threadIdx.x; c < p64; c += blockDim.x) {
    double va = 0.0, vb = 0.0;
    if (real && c < p) {
      va = (xw[row * p_pad + c] - (euclid ? mu[c] : 0.0)) * mul;
      vb = va * bmul;
    } else if (real && euclid && c == p) {   // the augmented K column
      va = 1.0;
      vb = n2s;
    }
    __half h, l;
    split_half(va, h, l);
    ah[row * p64 + c] = h;
    al[row * p64 + c] = l;
    split_half(vb, h, l);
    bh[row * p64 + c] = h;
    bl[row * p64 + c] = l;
  }
}

// Merge the per-split candidate lists of a row (top 32*R by approximate key),
// recompute the candidates' distances in float64 (same accumulation order and
// epilogue as the FP64 Gram kernel), sort, emit k, and record what the margin
// check needs.  Error and margin live in the space where the approximate key
// is linear: squared distance for euclidean, distance otherwise.
template <int R>
__global__ void __launch_bounds__(32 * RR_WARPS)
rerank_kernel(const double* __restrict__ xw, const double* __restrict__ stat,
              const double* __restrict__ scale, const double* __restrict__ n2c, int64_t n,
              int64_t p_pad, int euclid,
              const float* __restrict__ part_key, const int32_t* __restrict__ part_idx,
              int splits, int kc, int64_t rows, int64_t row_begin, int k, int exclude,
              int64_t* __restrict__ idx_out, double* __restrict__ dist_out,
              unsigned int* __restrict__ err_bits, double* __restrict__ tau_out,
              double* __restrict__ dk_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int K = k + 1;
  const int64_t gi = row_begin + row;
  const double inv_s2 = scale[1];
  const double si = stat[gi];
  // every cell outside the union of the split lists has a key >= the last
  // entry of its split's (full) list
  WarpList<R> L;
  L.init();
  double t1 = pos_inf();
  for (int s = 0; s < splits; ++s) {
    const int64_t o = ((int64_t)s * rows + row) * KC;
    const float kf = part_key[o + lane];
    const int j = part_idx[o + lane];
    const double last = __shfl_sync(FULL, (double)kf, kc - 1);
    const int last_j = __shfl_sync(FULL, j, kc - 1);
    if (last_j != 0x7fffffff) t1 = fmin(t1, last);
    L.offer((double)kf, j, j != 0x7fffffff, 32 * R, lane);
  }
  // ... and every listed cell that is not re-ranked has a key >= the last
  // entry of the merged list
  {
    const double last = __shfl_sync(FULL, L.d[R - 1], 31);
    const int last_j = __shfl_sync(FULL, L.i[R - 1], 31);
    if (last_j != 0x7fffffff) t1 = fmin(t1, last);
  }
  const double n2c_i = euclid ? n2c[gi] : 0.0;
  auto to_metric = [&](double key) {   // approximate key -> margin space
    return euclid ? fma(key, inv_s2, n2c_i) : 1.0 + key;
  };
  // exact distance of my candidates: the warp streams 32 candidate rows
  // through shared memory in 64-gene chunks (coalesced 512-byte row segments),
  // then every lane walks its own row in gene order with one FMA chain -- the
  // accumulation order of the FP64 Gram kernel
  extern __shared__ double rr_smem[];
  double* buf = rr_smem + (threadIdx.x >> 5) * (32 * RR_LD + RR_CH);
  double* xs = buf + 32 * RR_LD;
  const double* __restrict__ xi = xw + gi * p_pad;
  double dist[R], lin[R];
  float err = 0.f;
#pragma unroll
  for (int rg = 0; rg < R; ++rg) {
    const int j = L.i[rg];
    const bool valid = j != 0x7fffffff;
    double acc = 0.0;
    for (int64_t c0 = 0; c0 < p_pad; c0 += RR_CH) {
      const int64_t left = p_pad - c0;
      const bool in = 2 * lane < left;          // p_pad is a multiple of 16
#pragma unroll 8
      for (int m = 0; m < 32; ++m) {
        const int jm = __shfl_sync(FULL, j, m);
        double2 w = make_double2(0.0, 0.0);
        if (jm != 0x7fffffff && in)
          w = *reinterpret_cast<const double2*>(xw + (int64_t)jm * p_pad + c0 + 2 * lane);
        buf[m * RR_LD + 2 * lane] = w.x;
        buf[m * RR_LD + 2 * lane + 1] = w.y;
      }
      if (rg == 0 || p_pad > RR_CH) {
        double2 u = make_double2(0.0, 0.0);
        if (in) u = *reinterpret_cast<const double2*>(xi + c0 + 2 * lane);
        xs[2 * lane] = u.x;
        xs[2 * lane + 1] = u.y;
      }
      __syncwarp();
      const int lim = left < RR_CH ? (int)left : RR_CH;
      for (int c = 0; c < lim; ++c) acc = fma(xs[c], buf[lane * RR_LD + c], acc);
      __syncwarp();
    }
    dist[rg] = pos_inf();
    lin[rg] = pos_inf();
    if (valid) {
      const bool lower = gi >= j;
      const double sj = stat[j];
      const double s_hi = lower ? si : sj, s_lo = lower ? sj : si;
      dist[rg] = finish_entry(acc, s_hi, s_lo, gi == j, euclid);
      if (euclid) {
        double t2 = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, acc), s_hi), s_lo);
        lin[rg] = gi == j ? 0.0 : (t2 < 0.0 ? 0.0 : t2);
      } else {
        lin[rg] = dist[rg];
      }
      // observed approximation error (feeds the global margin); the forced
      // zero of the diagonal is not an approximation error
      if (gi != j) err = fmaxf(err, (float)fabs(to_metric(L.d[rg]) - lin[rg]));
    }
  }
  for (int o = 16; o > 0; o >>= 1) err = fmaxf(err, __shfl_xor_sync(FULL, err, o));
  if (lane == 0) atomicMax(err_bits, __float_as_uint(err * (1.f + 1e-6f)));
  // rank by (exact distance, index)
  int rank[R];
#pragma unroll
  for (int rg = 0; rg < R; ++rg) rank[rg] = 0;
#pragma unroll
  for (int sg = 0; sg < R; ++sg) {
    for (int m = 0; m < 32; ++m) {
      const double dm = __shfl_sync(FULL, dist[sg], m);
      const int jm = __shfl_sync(FULL, L.i[sg], m);
      // empty slots all carry (+inf, INT_MAX): their position breaks the tie so
      // that every entry gets its own rank (and every output slot is written)
#pragma unroll
      for (int rg = 0; rg < R; ++rg) {
        const bool same = dm == dist[rg] && jm == L.i[rg];
        if (before(dm, jm, dist[rg], L.i[rg]) || (same && sg * 32 + m < rg * 32 + lane))
          ++rank[rg];
      }
    }
  }
  // scatter through shared memory into sorted order (entry e -> lane e%32, reg e/32)
  __syncwarp();
  int* sidx = reinterpret_cast<int*>(buf + 32 * R);
#pragma unroll
  for (int rg = 0; rg < R; ++rg) {
    buf[rank[rg]] = dist[rg];
    buf[32 * R + 32 * R + rank[rg]] = lin[rg];
    sidx[rank[rg]] = L.i[rg];
  }
  __syncwarp();
  WarpList<R> S;
  double slin[R];
#pragma unroll
  for (int rg = 0; rg < R; ++rg) {
    S.d[rg] = buf[rg * 32 + lane];
    S.i[rg] = sidx[rg * 32 + lane];
    slin[rg] = buf[32 * R + 32 * R + rg * 32 + lane];
  }
  S.emit(K, exclude, (int)gi, idx_out + row * k, dist_out + row * k, lane);
  double dk = 0.0;
#pragma unroll
  for (int rg = 0; rg < R; ++rg) {
    const double x = __shfl_sync(FULL, slin[rg], (K - 1) & 31);
    if (rg == (K - 1) >> 5) dk = x;
  }
  if (lane == 0) {
    tau_out[row] = t1 == pos_inf() ? pos_inf() : to_metric(t1);
    dk_out[row] = dk;
  }
}

__global__ void __launch_bounds__(256)
margin_kernel(const double* __restrict__ tau, const double* __restrict__ dk,
              const unsigned int* __restrict__ err_bits, const double* __restrict__ scale,
              int euclid, int kblocks, int64_t rows, int64_t row_begin,
              unsigned char* __restrict__ tile_flag, unsigned char* __restrict__ row_flag,
              int* __restrict__ n_flagged) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  // A cell that was not re-ranked has approximate metric >= tau, hence exact
  // metric >= tau - eps; it cannot enter the top k+1 if dk < tau - eps.  eps is
  // the larger of (a) 8 x the largest error OBSERVED on the re-ranked candidates
  // of this launch and (b) an a-priori bound that does not depend on what was
  // re-ranked: every K slab of 32 adds 6 tensor-core products (K = 16 each) to
  // the FP32 accumulator, each addition rounding at 2^-24 of a partial sum that
  // is bounded by B in margin space, plus the dropped lo.lo term and the FP16
  // split of the operands (2^-21 B).  B: |cos| <= 1 for unit vectors (cosine,
  // correlation); for euclidean the key is s^2 (|y_c|^2 - 2 y_r.y_c) with
  // s^2 |y|^2 < 2^15, i.e. |partial| <= 3 * 2^15 / s^2 in squared-distance units.
  const double d = dk[row];
  const double B = euclid ? 3.0 * 32768.0 * scale[1] : 1.0;
  const double eps_prior = ((6.0 * (double)kblocks + 2.0) * 5.9604644775390625e-8 +
                            4.76837158203125e-7) * B;
  const double eps_seen = 8.0 * (double)__uint_as_float(*err_bits);
  const double eps = (eps_seen > eps_prior ? eps_seen : eps_prior) + 1e-12 * (d > 1.0 ? d : 1.0);
  const bool bad = !(d + eps < tau[row]);
  row_flag[row] = bad;
  if (bad) {
    tile_flag[(row_begin + row) / 128 - row_begin / 128] = 1;
    atomicAdd(n_flagged, 1);
  }
}

int encode_map(CUtensorMap* m, const __half* ptr, int64_t rows, int64_t cols, int box_rows) {
  return encode_tensor_map(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, rows, cols, box_rows, TK,
                           CU_TENSOR_MAP_SWIZZLE_64B);
}

struct FastMaps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;   // boxes of 128 (A) / 256 (B) rows
  CUtensorMap b_hi_half, b_lo_half;     // B in boxes of 128 rows (CTA pairs)
};

// CTAs per MMA.  Measured on B200 (tools/fast_knn_bench.py): CTA pairs
// (cta_group::2, half the B tile per CTA) are 1-8 % faster for long rows
// (100,000 x 2,000: D 91.9 -> 88.1 ms; 68,579 x 2,000 kNN 61.7 -> 56.8 ms) and
// 5-13 % slower for one- or two-slab rows (1,000,000 x 50: 394 -> 431 ms), so
// pairs are used from 256 K columns up.  SCEDAR_B200_TC_CG = 1 | 2 forces one.
int tc_cta_group(int64_t p64) {
  const char* e = getenv("SCEDAR_B200_TC_CG");
  if (e && (e[0] == '1' || e[0] == '2')) return e[0] - '0';
  return p64 >= 256 ? 2 : 1;
}

template <bool TOPK, int CG, bool AR = false>
int launch_tc(const FastMaps& m, const TcArgs& a, cudaStream_t s) {
  auto kern = gram_tc_kernel<TOPK, CG, AR>;
  SCEDAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  const unsigned grid = (unsigned)((int64_t)a.groups * a.splits * CG);
  if (CG == 1) {
    kern<<<grid, AR ? TC_THREADS_AR : TC_THREADS, TC_SMEM, s>>>(m.a_hi, m.a_lo, m.b_hi, m.b_lo, a);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TC_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    SCEDAR_CUDA(cudaLaunchKernelEx(&cfg, kern, m.a_hi, m.a_lo, m.b_hi_half, m.b_lo_half, a));
  }
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

// Schedule: S column splits x G groups of row tiles; CTA (g, s) walks the row
// tiles g, g + G, ...  Measured on B200 (profiles/README.md): a persistent
// lockstep grid (S * G = #SMs) is no faster when the operands exceed L2 -- the
// kernel is bound by L2 -> shared-memory bandwidth, not by HBM -- and slower
// when they fit (148 CTAs hammer the same L2 lines), so every row tile gets
// its own CTAs (G = RT) and the hardware scheduler de-synchronises them.
// STORE: at least 8 column splits; TOPK: as few as fill the SMs twice, every
// split warms up its own candidate lists.
void tc_schedule(int64_t RT, int64_t n, int64_t min_splits, int* splits, int* groups) {
  const int64_t CT = (n + TN - 1) / TN;
  int64_t s = (2 * (int64_t)num_sms() + RT) / (RT + 1);
  if (s < min_splits) s = min_splits;
  if (s > CT) s = CT;
  if (s > 32) s = 32;
  if (s < 1) s = 1;
  *splits = (int)s;
  *groups = (int)RT;
}

}  // namespace

int encode_tensor_map(CUtensorMap* m, CUtensorMapDataType dtype, int elem_bytes, const void* ptr,
                      int64_t rows, int64_t cols, int box_rows, int box_cols,
                      CUtensorMapSwizzle swizzle) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  // resolved through the runtime: the library must load (and be inspected) on
  // machines without a driver, so it cannot link against libcuda
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) !=
            cudaSuccess || !fn)
      return fail(SCEDAR_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    encode = reinterpret_cast<encode_fn>(fn);
  }
  CUresult r = encode(m, dtype, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SCEDAR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " +
                                     std::to_string((int)r));
  return SCEDAR_OK;
}

// fast-path working set of a cells handle (created on first use)
struct FastState {
  __half* ah = nullptr;   // A role, hi / lo halves: [n256, p64]
  __half* al = nullptr;
  __half* bh = nullptr;   // B role
  __half* bl = nullptr;
  double* scale = nullptr;  // [2] on the device: s, 1/s^2
  double* mu = nullptr;     // euclid: column means [p_pad]
  double* n2c = nullptr;    // euclid: |x - mu|^2 [n256]
  int64_t n256 = 0, p64 = 0;   // p64: K extent, a multiple of the slab width TK
  FastMaps maps;
};

void fast_state_free(void* p, cudaStream_t s) {
  FastState* f = static_cast<FastState*>(p);
  if (!f) return;
  if (f->ah) cudaFreeAsync(f->ah, s);
  if (f->al) cudaFreeAsync(f->al, s);
  if (f->bh) cudaFreeAsync(f->bh, s);
  if (f->bl) cudaFreeAsync(f->bl, s);
  if (f->scale) cudaFreeAsync(f->scale, s);
  if (f->mu) cudaFreeAsync(f->mu, s);
  if (f->n2c) cudaFreeAsync(f->n2c, s);
  delete f;
}

static int ensure_fast(const scedar_cells* c, cudaStream_t s, FastState** out) {
  scedar_cells* mc = const_cast<scedar_cells*>(c);
  if (mc->fast) {
    *out = static_cast<FastState*>(mc->fast);
    return SCEDAR_OK;
  }
  const int euclid = c->metric == SCEDAR_METRIC_EUCLIDEAN;
  FastState* f = new FastState();
  f->n256 = round_up(c->n_pad, 256);
  // euclidean needs one spare K column for the norm term
  f->p64 = round_up(c->p + (euclid ? 1 : 0), TK);
  if (f->p64 < TK) f->p64 = TK;
  const size_t hb = sizeof(__half) * f->n256 * f->p64;
  cudaError_t e = cudaMallocAsync(&f->ah, hb, s);
  if (e == cudaSuccess) e = cudaMallocAsync(&f->al, hb, s);
  if (e == cudaSuccess) e = cudaMallocAsync(&f->bh, hb, s);
  if (e == cudaSuccess) e = cudaMallocAsync(&f->bl, hb, s);
  if (e == cudaSuccess) e = cudaMallocAsync(&f->scale, 2 * sizeof(double), s);
  if (e == cudaSuccess) e = cudaMallocAsync(&f->mu, sizeof(double) * c->p_pad, s);
  if (e == cudaSuccess) e = cudaMallocAsync(&f->n2c, sizeof(double) * f->n256, s);
  if (e != cudaSuccess) {
    cudaGetLastError();
    fast_state_free(f, s);
    return fail(SCEDAR_ERR_NOMEM, std::string("fast-path buffers: ") + cudaGetErrorString(e));
  }
  if (euclid) {
    Scratch part;
    int prc = part.alloc(sizeof(double) * CM_CHUNKS * c->p_pad, s);
    if (prc != SCEDAR_OK) {
      fast_state_free(f, s);
      return prc;
    }
    colsum_kernel<<<dim3((unsigned)((c->p_pad + 31) / 32), CM_CHUNKS), 256, 0, s>>>(
        c->xw, c->n, c->p_pad, part.as<double>());
    colmean_kernel<<<(unsigned)((c->p_pad + 255) / 256), 256, 0, s>>>(part.as<double>(), c->n,
                                                                     c->p_pad, f->mu);
    n2c_kernel<<<(unsigned)((f->n256 + 7) / 8), 256, 0, s>>>(c->xw, f->mu, c->n, f->n256, c->p,
                                                            c->p_pad, f->n2c);
    count_launch(3);
  }
  fast_scale_kernel<<<1, 1024, 0, s>>>(euclid ? f->n2c : c->stat, c->n, euclid, f->scale);
  split_kernel<<<(unsigned)f->n256, 256, 0, s>>>(c->xw, c->stat, f->scale, f->mu, f->n2c, c->n,
                                                 c->p, c->p_pad,
                                                 f->p64, euclid, f->ah, f->al, f->bh, f->bl);
  count_launch(2);
  int rc = SCEDAR_OK;
  if (cudaGetLastError() != cudaSuccess) rc = fail(SCEDAR_ERR_CUDA, "split kernel launch failed");
  if (rc == SCEDAR_OK) rc = encode_map(&f->maps.a_hi, f->ah, f->n256, f->p64, TM);
  if (rc == SCEDAR_OK) rc = encode_map(&f->maps.a_lo, f->al, f->n256, f->p64, TM);
  if (rc == SCEDAR_OK) rc = encode_map(&f->maps.b_hi, f->bh, f->n256, f->p64, TN);
  if (rc == SCEDAR_OK) rc = encode_map(&f->maps.b_lo, f->bl, f->n256, f->p64, TN);
  if (rc == SCEDAR_OK) rc = encode_map(&f->maps.b_hi_half, f->bh, f->n256, f->p64, TN / 2);
  if (rc == SCEDAR_OK) rc = encode_map(&f->maps.b_lo_half, f->bl, f->n256, f->p64, TN / 2);
  if (rc != SCEDAR_OK) {
    fast_state_free(f, s);
    return rc;
  }
  mc->fast = f;
  *out = f;
  return SCEDAR_OK;
}

static TcArgs tc_args(const scedar_cells* c, const FastState* f, int64_t row_begin,
                      int64_t row_end) {
  TcArgs a{};
  a.stat = c->stat;
  a.scale = f->scale;
  a.n2c = f->n2c;
  a.n = c->n;
  a.kblocks = (int)(f->p64 / TK);
  a.row_begin = row_begin;
  a.row_end = row_end;
  a.euclid = c->metric == SCEDAR_METRIC_EUCLIDEAN;
  return a;
}

int launch_fast_stripe(const scedar_cells* c, int64_t row_begin, int64_t row_end, double* d_out,
                       int64_t ldd, cudaStream_t s, int symmetric) {
  if (row_end <= row_begin) return SCEDAR_OK;
  if (symmetric && (row_begin != 0 || row_end != c->n))
    return fail(SCEDAR_ERR_ARG, "the symmetric fast path covers the whole matrix");
  FastState* f = nullptr;
  SCEDAR_CHECK(ensure_fast(c, s, &f));
  TcArgs a = tc_args(c, f, row_begin, row_end);
  a.out = d_out;
  a.ldd = ldd;
  a.sym = symmetric;
  const int64_t RT = (row_end + TM - 1) / TM - row_begin / TM;
  const int cg = tc_cta_group(f->p64);
  const int64_t units = (RT + cg - 1) / cg;
  tc_schedule(units, c->n, 8, &a.splits, &a.groups);
  a.row_tiles = (int)units;
  return cg == 2 ? launch_tc<false, 2>(f->maps, a, s) : launch_tc<false, 1>(f->maps, a, s);
}

// k + 1 + 8 spare candidates must be re-ranked: up to 64 per cell
int fast_knn_max_k() { return 2 * KC - 9; }

int launch_fast_knn(const scedar_cells* c, int k, int exclude, int64_t row_begin,
                    int64_t row_end, int64_t* idx_out, double* dist_out, cudaStream_t s) {
  const int64_t rows = row_end - row_begin;
  if (rows <= 0 || k == 0) return SCEDAR_OK;
  FastState* f = nullptr;
  SCEDAR_CHECK(ensure_fast(c, s, &f));
  TcArgs a = tc_args(c, f, row_begin, row_end);
  // lists hold 32 candidates per split: k + 1 + 8 <= 32 is served by any
  // number of splits, larger k re-ranks 64 drawn from >= 4 splits
  const bool wide = k + 1 + 8 > KC;
  // k + 1 results + 8 spare candidates per list (fewer list entries = fewer
  // inserts).  On data with near-ties at the threshold (log counts) some cells
  // then fail the margin check; they are recomputed row by row below, which
  // costs nothing measurable (100,000 x 2,000: 112 ms with 8 spare, 114 ms with
  // 16; with the tile-granular fallback 8 spare cost 158-182 ms).
  int spare = 8;
  if (const char* e = getenv("SCEDAR_B200_TC_SPARE")) spare = std::max(1, atoi(e));
  a.kc = wide ? KC : std::min(KC, k + 1 + spare);
  if (getenv("SCEDAR_B200_TC_KC32")) a.kc = KC;
  const int64_t RT = (row_end + TM - 1) / TM - row_begin / TM;
  const int cg = tc_cta_group(f->p64);
  const int64_t units = (RT + cg - 1) / cg;
  // one or two K slabs: A-resident kernel with two epilogue groups, each of
  // which keeps (and emits) its own lists
  const bool a_res = cg == 1 && a.kblocks <= 2 && !getenv("SCEDAR_B200_TC_NO_AR");
  const int lists_per_split = a_res ? 2 : 1;
  tc_schedule(units, c->n, wide ? 4 / lists_per_split : 1, &a.splits, &a.groups);
  a.row_tiles = (int)units;
  const int n_lists = a.splits * lists_per_split;
  Scratch pk, pi, tau, dk, misc, flags, rflags;
  SCEDAR_CHECK(rflags.alloc((size_t)rows + 16, s));
  SCEDAR_CHECK(pk.alloc(sizeof(float) * n_lists * rows * KC, s));
  SCEDAR_CHECK(pi.alloc(sizeof(int32_t) * n_lists * rows * KC, s));
  SCEDAR_CHECK(tau.alloc(sizeof(double) * rows, s));
  SCEDAR_CHECK(dk.alloc(sizeof(double) * rows, s));
  SCEDAR_CHECK(misc.alloc(16, s));
  SCEDAR_CHECK(flags.alloc(RT + 16, s));
  SCEDAR_CUDA(cudaMemsetAsync(misc.ptr, 0, 16, s));
  SCEDAR_CUDA(cudaMemsetAsync(flags.ptr, 0, RT + 16, s));
  a.part_key = pk.as<float>();
  a.part_idx = pi.as<int32_t>();
  if (a_res)
    SCEDAR_CHECK((launch_tc<true, 1, true>(f->maps, a, s)));
  else
    SCEDAR_CHECK(cg == 2 ? (launch_tc<true, 2>(f->maps, a, s)) : (launch_tc<true, 1>(f->maps, a, s)));
  unsigned int* err_bits = misc.as<unsigned int>();
  int* n_flagged = reinterpret_cast<int*>(misc.as<unsigned int>() + 1);
  const unsigned rr_grid = (unsigned)((rows + RR_WARPS - 1) / RR_WARPS);
  if (wide) {
    SCEDAR_CUDA(cudaFuncSetAttribute(rerank_kernel<2>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, RR_SMEM));
    rerank_kernel<2><<<rr_grid, 32 * RR_WARPS, RR_SMEM, s>>>(
        c->xw, c->stat, f->scale, f->n2c, c->n, c->p_pad, a.euclid, a.part_key, a.part_idx, n_lists, a.kc,
        rows, row_begin, k, exclude, idx_out, dist_out, err_bits, tau.as<double>(),
        dk.as<double>());
  } else {
    SCEDAR_CUDA(cudaFuncSetAttribute(rerank_kernel<1>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, RR_SMEM));
    rerank_kernel<1><<<rr_grid, 32 * RR_WARPS, RR_SMEM, s>>>(
        c->xw, c->stat, f->scale, f->n2c, c->n, c->p_pad, a.euclid, a.part_key, a.part_idx, n_lists, a.kc,
        rows, row_begin, k, exclude, idx_out, dist_out, err_bits, tau.as<double>(),
        dk.as<double>());
  }
  margin_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(
      tau.as<double>(), dk.as<double>(), err_bits, f->scale, a.euclid, a.kblocks, rows, row_begin,
      flags.as<unsigned char>(), rflags.as<unsigned char>(), n_flagged);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch(2);
  // rows whose candidate margin is too thin are recomputed by the FP64 kernel
  int h_flagged = 0;
  SCEDAR_CUDA(cudaMemcpyAsync(&h_flagged, n_flagged, sizeof(int), cudaMemcpyDeviceToHost, s));
  SCEDAR_CUDA(cudaStreamSynchronize(s));
  if (h_flagged == 0) return SCEDAR_OK;
  if ((int64_t)h_flagged * 8 <= rows && !getenv("SCEDAR_B200_TC_TILE_FALLBACK")) {
    // a few cells: recompute exactly those -- gather their working-copy rows,
    // contract them against all cells with the FP64 kernel (entries
    // bit-identical to the whole matrix), select, and scatter the lists into
    // the cells' own slots.  128 x n x p flops per 128 flagged cells instead of
    // per tile that happens to contain one.
    const int64_t m = h_flagged, m_pad = round_up(m, 128), n = c->n;
    Scratch ids, gbuf, dbuf;
    SCEDAR_CHECK(ids.alloc(sizeof(int64_t) * m_pad, s));
    SCEDAR_CHECK(gbuf.alloc(sizeof(double) * m_pad * c->p_pad, s));
    SCEDAR_CHECK(launch_flag_to_ids(rflags.as<unsigned char>(), rows, row_begin,
                                    ids.as<int64_t>(), m, m_pad, s));
    SCEDAR_CUDA(cudaMemsetAsync(gbuf.ptr, 0, sizeof(double) * m_pad * c->p_pad, s));
    SCEDAR_CHECK(launch_gather_rows(c->xw, c->p_pad, c->p_pad, ids.as<int64_t>(), m,
                                    gbuf.as<double>(), c->p_pad, s));
    int64_t chunk = (int64_t)(2e9 / (8.0 * (double)n)) / 128 * 128;
    chunk = std::max<int64_t>(128, std::min(chunk, m_pad));
    SCEDAR_CHECK(dbuf.alloc(sizeof(double) * (size_t)chunk * (size_t)n, s));
    for (int64_t c0 = 0; c0 < m; c0 += chunk) {
      const int64_t mc = std::min(chunk, m - c0);
      SCEDAR_CHECK(launch_gram_gathered(c, gbuf.as<double>() + c0 * c->p_pad,
                                        ids.as<int64_t>() + c0, mc, round_up(mc, 128),
                                        dbuf.as<double>(), n, s));
      SCEDAR_CHECK(launch_topk_from_d(dbuf.as<double>(), n, n, 0, mc, k, exclude, idx_out,
                                      dist_out, s, ids.as<int64_t>() + c0, row_begin));
    }
    return SCEDAR_OK;
  }
  std::vector<unsigned char> hf((size_t)RT);
  SCEDAR_CUDA(cudaMemcpyAsync(hf.data(), flags.ptr, (size_t)RT, cudaMemcpyDeviceToHost, s));
  SCEDAR_CUDA(cudaStreamSynchronize(s));
  const int64_t t0 = row_begin / TM;
  for (int64_t t = 0; t < RT; ++t) {
    if (!hf[(size_t)t]) continue;
    int64_t t1 = t;
    while (t1 + 1 < RT && hf[(size_t)(t1 + 1)]) ++t1;
    const int64_t r0 = std::max<int64_t>(row_begin, (t0 + t) * TM);
    const int64_t r1 = std::min<int64_t>(row_end, (t0 + t1 + 1) * TM);
    SCEDAR_CHECK(knn_stripe_fp64(c, k, exclude, r0, r1, idx_out + (r0 - row_begin) * k,
                                 dist_out + (r0 - row_begin) * k, s));
    t = t1;
  }
  return SCEDAR_OK;
}

}  // namespace scedar
