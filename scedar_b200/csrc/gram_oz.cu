// K2 (FP64-grade path on the tcgen05 tensor cores) -- Gram contraction X.Xt as
// an error-free split over the INT8 tensor pipe (tcgen05.mma.kind::i8, exact
// INT32 accumulation in TMEM, operands streamed by TMA), with the float64 metric
// epilogue of gram_f64.cu fused.  Selected by SCEDAR_PRECISION_FP64_I8 (7
// slices) / SCEDAR_PRECISION_FP64_I8S6 (6 slices).
//
// Replaces the same reference arithmetic as gram_f64.cu: np.dot(x, x.T)
// (scedar/eda/sdm.py:1245-1247), the cosine / correlation scaling (1249-1266),
// sklearn's euclidean formula (1216) and num_correct_dist_mat's clean-up
// (215-221).  tcgen05 has no f64 kind; the DMMA kernel sits at 0.92 of a
// 37 TFLOP/s pipe, the INT8 pipe is two orders of magnitude wider.
//
// Split (Ozaki scheme, fixed point per row).  Row i of the working copy is
// scaled by 2^-e_i (2^e_i > max|x_i|) and truncated to F = 7 + 8(S-1)
// fractional bits: q = floor(x 2^(F - e_i)), a two's-complement integer that is
// cut into S bytes, most significant first.  The top byte is stored
// offset-binary (a_0 + 128), so EVERY slice is an unsigned byte and the slices
// of the column operand can be stacked along N: one MMA of N = 64 m columns
// multiplies slice s of the row tile with slices t0..t0+m-1 of the column tile
// and lands on the accumulators of the diagonals s+t0 .. s+t0+m-1, which sit
// side by side in TMEM (diagonal d = columns [64 d, 64 d + 64)).  Only the
// pairs with s + t < S are formed (S(S+1)/2 products in 8 (S=6) / 10 (S=7)
// instructions per 32-byte K step); the dropped pairs are below 2^-(8S-2) of
// the row scales.  The offset of the top bytes contributes 128 (R_d[i] +
// R_d[j]) (+ 2^14 K on diagonal 0) with R_d the byte sums of a row; the
// epilogue subtracts it exactly in 64-bit integers and joins the two halves of
// the weighted sum into ONE 64-bit integer W with 17 fraction bits below the
// unit of the upper half (oz_weighted; the cut, 2^-48 of the row scales, is below
// what the dropped diagonals are worth), converted to float64 once:
// P = 2^(e_i + e_j - 30 - 17) W.  Integer input (counts) therefore gives the
// exact dot product, identical rows give P_ij == P_ii == P_jj bit for bit (the
// statistic comes from the same integers through the same function), and for
// S = 7 the truncation (<= 6 p 2^-56 of the row scales, a few 1e-15 of
// |x_i||x_j| on log counts) is of the size of DGEMM's own rounding.  K chunks
// (4096 columns per INT32 accumulation) are summed as integers too.
//
// Kernel (persistent, one CTA per SM, 320 threads):
//   warp 0     TMA producer, 128-byte K slabs (128B swizzle; with 64-byte box rows
//              the TMA unit of one SM delivers ~52 B/cycle, with 128-byte rows
//              73-105 -- tools/tma_probe.cu -- and this kernel needs 47): the S
//              slices of the row tile (16 KB each) sit in S single slots that
//              are re-filled slice by slice as the MMAs of a slice retire, the
//              S slices of the column tile (8 KB each, stacked) are double
//              buffered; full / empty mbarriers per slot
//   warp 1     MMA issuer (elected lane), slice-major inside a slab, descriptors
//              = slot base + constant
//   warps 2-9  epilogue: TMEM lane quarter = warp % 4, column half = (warp-2)/4.
//              Phase 1 drains the accumulators (S x tcgen05.ld.x4 per 4 columns,
//              the next 4 in flight) into 32 integer weighted sums per thread
//              and hands TMEM back; phase 2 (under the next tile's MMAs) converts
//              them (one I2F each), applies the metric and stores: the mirrored
//              tile straight from the registers (lanes = consecutive entries of
//              a row of D), the tile itself in 32-byte runs (st.global.v4.f64),
//              and -- when the handle carries block-minima buffers -- the minimum
//              of every run of 32 entries of both, for the selection that follows
// TMEM holds S x 64 columns (448 of 512 for S = 7): no second accumulator set,
// so phase 1 (~3 k cycles of a ~75 k cycle tile at p = 2000) is not overlapped.
// What bounds it (SCEDAR_B200_OZ_DEBUG=1 prints per-tile cycle counters;
// tools/umma_probe.cu, profiles/r02_umma_probe.log, r02_umma_contend.log):
// * issued alone, from operands resident in shared memory, the 10-instruction K
//   step runs at 909 cycles (ideal 896); in the kernel it takes ~1,130.
// * while the tensor pipe is busy the other warps of the SM pay ~12x for every
//   FP64 instruction and ~4x for every shared-memory load (and the MMAs lose
//   8 % to them); integer work, shuffles and global stores are unaffected.  The
//   first round-2 epilogue (FP64 scaling, column constants from L1) took 53-69 k
//   cycles per tile in phase 2 and 12.7 k in phase 1 and had become the critical
//   path; the integer form above takes 37 k / 3.2 k and waits for the MMAs.
// * the cycles saved do not show as time: the launch runs under the board's
//   power cap, the SM clock settles where the executed INT8 MACs fit the power
//   budget (~1.65 GHz before, ~1.45 GHz after, same 214-225 ms at 100k x 2000).
//   Executed MACs are 0.93 of twice the sustained bf16 rate of
//   MEASURED_PEAKS.json; what is left is S (6 slices: 0.79 of the time at
//   2.7e-12 instead of 4e-14).
// cta_group::2 would halve the column operand's shared-memory reads, but splits
// the N columns of an instruction between the two CTAs' shared memories, which
// breaks the stacking of slices along N.
// Tiles (128 rows x 64 columns) come from a host-built list in L2-friendly
// order, dealt round-robin to the CTAs; the list also carries what to store
// (tile / mirrored tile), so the whole-matrix, row-stripe, peer-store and
// row-range forms are one kernel.
// Roofline: INT8 tensor pipe; executed MACs = S(S+1)/2 x the algorithmic ones.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tc_sm100.cuh"
#include "topk_list.cuh"

namespace scedar {
namespace {

constexpr int OM = 128;            // tile rows = TMEM lanes
constexpr int ON = 64;             // tile columns (per slice diagonal)
constexpr int OKB = 128;           // K bytes per slab (one 128B-swizzle row: the TMA unit of an
                                   // SM moves ~1 box row per cycle, so 64-byte rows cap it at
                                   // ~52 B/cycle -- tools/tma_probe.cu -- against the 47 this
                                   // kernel needs; 128-byte rows reach 73-105)
constexpr int A_SLAB = OM * OKB;   // 16 KB: one slice of the row tile, one K slab
constexpr int B_SLAB = ON * OKB;   // 8 KB
// shared memory: the S row-tile slices of ONE slab, each slot re-filled as soon
// as the MMAs of its slice have retired (the MMA loop runs slice-major), and
// TWO slabs of column-tile slices (all S are needed from the first slice on)
constexpr int OZ_BAR_BYTES = 256;
constexpr int oz_smem(int S) { return S * A_SLAB + 2 * S * B_SLAB + OZ_BAR_BYTES; }
static_assert(oz_smem(6) <= 227 * 1024 && oz_smem(7) <= 227 * 1024, "shared memory budget");
constexpr int OZ_EPI_WARPS = 8;
constexpr int OZ_EPI_THREADS = 32 * OZ_EPI_WARPS;
constexpr int OZ_THREADS = 64 + OZ_EPI_THREADS;
constexpr int OZ_KCHUNK = 4096;    // K columns accumulated in INT32 at a time: 7 * 255^2 * 4096
                                   // < 2^31; longer rows are contracted chunk by chunk, the
                                   // integer chunk sums added in the epilogue's registers
constexpr int OZ_CHUNK_SLABS = OZ_KCHUNK / OKB;
constexpr int kOzPeers = 16;
constexpr int TILE_DIRECT = 1, TILE_MIRROR = 2;

struct OzTile {
  int32_t I, J;      // row tile (128 rows), column tile (64 columns)
  int32_t flags;     // TILE_DIRECT: store D[I rows][J cols]; TILE_MIRROR: store the transpose
  int32_t pad;
};

struct OzArgs {
  const double* sc;        // 2^(e_i - 15)
  const long long* hi;     // offset corrections of the two halves of the weighted sum,
  const long long* lo;     // [row][chunk]
  int n_chunks;            // K chunks of <= 4096 columns
  long long c0_last;       // c0 of the last (possibly shorter) chunk
  const double* stat;      // sqrt(1/P_ii) (0 for zero rows) or P_ii, from the same integers
  long long c0;            // 2^14 K_chunk << 16: the offset x offset term of diagonal 0
  int64_t n, n_pl;         // cells; rows per slice plane
  int64_t row_lo, row_hi;  // rows of D this launch may store
  int kslabs;
  int euclid;
  const OzTile* tiles;
  int n_tiles;
  int64_t ldd;
  unsigned long long* dbg;                // SCEDAR_B200_OZ_DEBUG: 8 cycle counters per CTA
  int vec_ok;                             // rows of D are 32-byte aligned (ldd % 4 == 0)
  int world;                              // output stripes (1 for a single buffer)
  int64_t stripe_begin[kOzPeers + 1];     // row ranges of the stripes
  double* stripe_ptr[kOzPeers];
  // block minima for the top-k that follows (optional): bmin[r][row - begin_r][b]
  // = min of D[row][32 b .. 32 b + 31]; same stripe table as D
  double* bmin_ptr[kOzPeers];
  int64_t ldb;
  int wshift;   // fraction bits of the integer weighted sums (oz_wshift)
  int park_ns;  // suspend-time hint of the mbarrier waits (0: plain try_wait loop)
};

__device__ __forceinline__ void oz_wait(uint32_t bar, uint32_t parity, int park_ns) {
  if (park_ns > 0)
    mbar_wait_parked(bar, parity, (uint32_t)park_ns);
  else
    mbar_wait(bar, parity);
}

// Weighted sum of the diagonal accumulators minus the offset terms, as ONE
// 64-bit integer.  hi (unit 1) and lo (unit 2^-8(S-3)) are exact: diagonal d
// holds d + 1 products of at most 255^2 * 4096 < 2^28 each, so |hi| < 2^45 and
// |lo| < 2^56.  W = hi 2^ws + round(lo 2^(ws - 8(S-3))) keeps ws fraction bits
// below the hi unit (ws = 17 for a single K chunk: |W| < 2^62).  The cut is
// <= 2^-18 hi units = 2^-48 of the row scales, against up to 6 p 2^-56 from the
// diagonals the contraction drops; integer input whose dot products are
// integers of the hi unit's scale stays exact.  ls = 8 (S - 3) - ws >= 1.
template <int S>
__device__ __forceinline__ long long oz_weighted(const uint32_t (&acc)[S], long long hi_corr,
                                                 long long lo_corr, int ws, int ls) {
  long long hi = ((long long)acc[0] << 16) + ((long long)acc[1] << 8) + (long long)acc[2];
  long long lo = 0;
#pragma unroll
  for (int d = 3; d < S; ++d) lo = (lo << 8) + (long long)acc[d];
  hi -= hi_corr;
  lo -= lo_corr;
  return (hi << ws) + ((lo + (1ll << (ls - 1))) >> ls);
}
// fraction bits for n_chunks K chunks summed as integers: |sum| < 2^63
__host__ __device__ inline int oz_wshift(int n_chunks) {
  int ws = 17;
  while (n_chunks > 1) {
    --ws;
    n_chunks = (n_chunks + 1) >> 1;
  }
  return ws;
}

// ---- K1': slices, scales, offset corrections and the Gram diagonal -----------
template <int S>
__global__ void __launch_bounds__(256)
oz_slice_kernel(const double* __restrict__ xw, int64_t n, int64_t p, int64_t p_pad, int64_t p_oz,
                int64_t n_pl, int64_t row_begin, int euclid, long long c0,
                uint8_t* __restrict__ planes, double* __restrict__ sc, long long* __restrict__ hi,
                long long* __restrict__ lo, double* __restrict__ stat) {
  constexpr int F = 7 + 8 * (S - 1);
  const int64_t row = row_begin + blockIdx.x;
  const bool real = row < n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ double smax[8];
  __shared__ long long sred[8][2 * S];
  const double* __restrict__ xr = xw + row * p_pad;
  double m = 0.0;
  if (real)
    for (int64_t k = tid; k < p; k += 256) m = fmax(m, fabs(xr[k]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) smax[warp] = m;
  __syncthreads();
  m = smax[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmax(m, smax[w]);
  int e = 0;
  if (m > 0.0) frexp(m, &e);   // m = f 2^e, f in [0.5, 1): |x| < 2^e
  const int n_chunks = (int)((p_oz + OZ_KCHUNK - 1) / OZ_KCHUNK);
  const double sci = scalbn(1.0, e - 15);
  long long w_ii = 0;   // thread 0: the weighted sum of (row, row), as the tile epilogue forms it
  const int ws = oz_wshift(n_chunks), ls = 8 * (S - 3) - ws;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int64_t k_begin = (int64_t)ch * OZ_KCHUNK;
    const int64_t k_end = k_begin + OZ_KCHUNK < p_oz ? k_begin + OZ_KCHUNK : p_oz;
    long long R[S], dg[S];
#pragma unroll
    for (int s = 0; s < S; ++s) R[s] = dg[s] = 0;
    for (int64_t k4 = k_begin + (int64_t)tid * 4; k4 < k_end; k4 += 1024) {
      uint32_t pack[S];
#pragma unroll
      for (int s = 0; s < S; ++s) pack[s] = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t k = k4 + u;
        const double x = (real && k < p) ? xr[k] : 0.0;
        const long long q = __double2ll_rd(scalbn(x, F - e));
        int dgt[S];
        const int top = (int)(q >> (8 * (S - 1)));   // [-128, 127]
        dgt[0] = top + 128;
        R[0] += top;
#pragma unroll
        for (int s = 1; s < S; ++s) {
          dgt[s] = (int)((q >> (8 * (S - 1 - s))) & 255);
          R[s] += dgt[s];
        }
#pragma unroll
        for (int s = 0; s < S; ++s) pack[s] |= (uint32_t)dgt[s] << (8 * u);
        // what the tensor core accumulates for the pair (row, row)
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int t = 0; s + t < S; ++t) dg[s + t] += (long long)(dgt[s] * dgt[t]);
      }
#pragma unroll
      for (int s = 0; s < S; ++s)
        *reinterpret_cast<uint32_t*>(planes + ((int64_t)s * n_pl + row) * p_oz + k4) = pack[s];
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      for (int o = 16; o > 0; o >>= 1) {
        R[s] += __shfl_xor_sync(0xffffffffu, R[s], o);
        dg[s] += __shfl_xor_sync(0xffffffffu, dg[s], o);
      }
      if (lane == 0) {
        sred[warp][s] = R[s];
        sred[warp][S + s] = dg[s];
      }
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t acc[S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        long long r = 0, g = 0;
        for (int w = 0; w < 8; ++w) {
          r += sred[w][s];
          g += sred[w][S + s];
        }
        R[s] = r;
        acc[s] = (uint32_t)g;
      }
      const long long hi_i = 128 * ((R[0] << 16) + (R[1] << 8) + R[2]);
      long long lo_i = 0;
#pragma unroll
      for (int d = 3; d < S; ++d) lo_i = (lo_i << 8) + R[d];
      lo_i *= 128;
      hi[row * n_chunks + ch] = hi_i;
      lo[row * n_chunks + ch] = lo_i;
      const long long c0c = ((long long)16384 * (k_end - k_begin)) << 16;
      w_ii += oz_weighted<S>(acc, 2 * hi_i + c0c, 2 * lo_i, ws, ls);
    }
    __syncthreads();
  }
  if (tid == 0) {
    sc[row] = sci;
    const double p_ii = __dmul_rn((double)w_ii, __longlong_as_double((long long)(1023 - ws) << 52));
    const double P = __dmul_rn(p_ii, __dmul_rn(sci, sci));
    double sv = 0.0;
    if (real) sv = euclid ? P : __dsqrt_rn(P != 0.0 ? __ddiv_rn(1.0, P) : 0.0);
    stat[row] = sv;
  }
}

// ---- K2: the tile kernel -------------------------------------------------------
// 32-byte store of four consecutive doubles (evict-first: D is write-once)
__device__ __forceinline__ void st_cs_v4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c),
               "d"(d)
               : "memory");
}

template <int S>
__global__ void __launch_bounds__(OZ_THREADS, 1)
gram_oz_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const OzArgs a) {
  constexpr int A_STAGE = S * A_SLAB, B_STAGE = S * B_SLAB;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  if (raw & 1023u) __trap();        // the 128B-swizzled slabs need 1024-byte alignment
  const uint32_t base = raw;
  uint8_t* gbase = smem_raw;
  const uint32_t ring_a = base, ring_b = base + A_STAGE;
  const uint32_t bars = ring_b + 2 * B_STAGE;
  // barrier slots (8 bytes each): fullA[S], emptyA[S], fullB[2], emptyB[2], tfull, tempty
  const uint32_t full_a = bars, empty_a = bars + 64, full_b = bars + 128, empty_b = bars + 144,
                 tfull = bars + 160, tempty = bars + 168;
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(gbase + (bars - base) + 192);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KS = a.kslabs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_a + 8 * s, 1);
      mbar_init(empty_a + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_b + 8 * s, 1);
      mbar_init(empty_b + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, OZ_EPI_WARPS);
    mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc512(bars + 192);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- TMA producer.  Slab `it` of this CTA's stream needs slice slots
    // A_0..A_S-1 (use number `it` of each slot) and column buffer it & 1.  The
    // requests go out in the order the slots come free: the slices of slab
    // it + 1 one by one while slab it is contracted, then the columns of slab
    // it + 2 -- each a full slab period ahead of its use.
    if (lane == 0) {
      int my_tiles = 0;
      for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) ++my_tiles;
      const int total = my_tiles * KS;
      int c_i0 = 0, c_j0 = 0, c_tile = -1;
      auto coords = [&](int it, int& i0, int& j0, int& ks) {
        const int tl_i = it / KS;
        ks = it - tl_i * KS;
        if (tl_i != c_tile) {
          const OzTile tl = a.tiles[blockIdx.x + (long long)tl_i * gridDim.x];
          c_i0 = tl.I * OM;
          c_j0 = tl.J * ON;
          c_tile = tl_i;
        }
        i0 = c_i0;
        j0 = c_j0;
      };
      auto load_a = [&](int it) {
        int i0, j0, ks;
        coords(it, i0, j0, ks);
#pragma unroll
        for (int sl = 0; sl < S; ++sl) {
          oz_wait(empty_a + 8 * sl, (uint32_t)((it & 1) ^ 1), a.park_ns);
          mbar_expect_tx(full_a + 8 * sl, A_SLAB);
          tma_load_2d(ring_a + sl * A_SLAB, &map_a, ks * OKB, (int)(sl * a.n_pl) + i0,
                      full_a + 8 * sl);
        }
      };
      auto load_b = [&](int it) {
        int i0, j0, ks;
        coords(it, i0, j0, ks);
        const uint32_t buf = (uint32_t)(it & 1), ph = (uint32_t)((it >> 1) & 1);
        oz_wait(empty_b + 8 * buf, ph ^ 1, a.park_ns);
        mbar_expect_tx(full_b + 8 * buf, B_STAGE);
        const uint32_t st = ring_b + buf * B_STAGE;
#pragma unroll
        for (int sl = 0; sl < S; ++sl)
          tma_load_2d(st + sl * B_SLAB, &map_b, ks * OKB, (int)(sl * a.n_pl) + j0,
                      full_b + 8 * buf);
      };
      if (total > 0) load_b(0);
      if (total > 1) load_b(1);
      if (total > 0) load_a(0);
      for (int it = 0; it < total; ++it) {
        if (it + 1 < total) load_a(it + 1);
        if (it + 2 < total) load_b(it + 2);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: slice-major inside a slab, so that a row-tile slice is
    // done with (and its slot re-filled) after its four K steps
    const bool el = tc::elect_one();
    uint32_t it = 0, tcount = 0;
    long long w_tempty = 0, w_full = 0;
    const long long t_begin = clock64();
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
      for (int ks0 = 0; ks0 < KS; ks0 += OZ_CHUNK_SLABS, ++tcount) {
        // a new K chunk (or tile): the epilogue has drained the accumulators
        long long c0 = a.dbg ? clock64() : 0;
        oz_wait(tempty, (tcount & 1) ^ 1, a.park_ns);
        if (a.dbg) w_tempty += clock64() - c0;
        tc::fence_after();
        const int kc_end = KS - ks0 < OZ_CHUNK_SLABS ? KS - ks0 : OZ_CHUNK_SLABS;
        for (int kc = 0; kc < kc_end; ++kc, ++it) {
          const uint32_t buf = it & 1;
          c0 = a.dbg ? clock64() : 0;
          oz_wait(full_b + 8 * buf, (it >> 1) & 1, a.park_ns);
          if (a.dbg) w_full += clock64() - c0;
          const uint64_t db = tc::desc_sw128(ring_b + buf * B_STAGE);
#pragma unroll
          for (int sl = 0; sl < S; ++sl) {
            c0 = a.dbg ? clock64() : 0;
            oz_wait(full_a + 8 * sl, it & 1, a.park_ns);
            if (a.dbg) w_full += clock64() - c0;
            tc::fence_after();
            const uint64_t da = tc::desc_sw128(ring_a + sl * A_SLAB);
#pragma unroll
            for (int kk = 0; kk < OKB / 32; ++kk) {
              // slice sl of the row tile x slices [0, S - sl) of the column tile in
              // one instruction (N = 64 m <= 256) or two of about equal length
              // (N = 64 alone would take 54 cycles instead of 32, N >= 128 runs at
              // N / 2: tools/umma_probe.cu).  Slice 0 goes first in a chunk: it
              // covers every diagonal, so its first K step zero-initialises them.
              constexpr int kFullRun = 4;
              const int run = S - sl;
              const int first = run <= kFullRun ? run : (run + 1) / 2;
              const uint32_t acc = (kc | kk | sl) != 0;
              if (el) {
                tc::mma_i8(tmem_base + (uint32_t)(sl * ON), da + (uint64_t)((kk * 32) >> 4),
                           db + (uint64_t)((kk * 32) >> 4), tc::idesc_u8(ON * first), acc);
                if (run > first)
                  tc::mma_i8(tmem_base + (uint32_t)((sl + first) * ON),
                             da + (uint64_t)((kk * 32) >> 4),
                             db + (uint64_t)((first * B_SLAB + kk * 32) >> 4),
                             tc::idesc_u8(ON * (run - first)), acc);
              }
            }
            if (el) tc::commit(empty_a + 8 * sl);   // slot reusable once these MMAs retire
            __syncwarp();
          }
          if (el) tc::commit(empty_b + 8 * buf);
          __syncwarp();
        }
        if (el) tc::commit(tfull);              // this chunk's accumulators are complete
        __syncwarp();
      }
    }
    if (a.dbg && lane == 0) {
      a.dbg[8 * blockIdx.x + 0] = (unsigned long long)w_tempty;
      a.dbg[8 * blockIdx.x + 1] = (unsigned long long)w_full;
      a.dbg[8 * blockIdx.x + 2] = (unsigned long long)(clock64() - t_begin);
      a.dbg[8 * blockIdx.x + 7] = tcount;
    }
  } else {
    // ---- epilogue warps 2..9 -----------------------------------------------------
    // While the tensor pipe is busy, FP64 instructions of the same SM issue ~12x
    // slower and shared-memory loads ~4x slower; integer work, shuffles and
    // global stores are unaffected (tools/umma_probe.cu `contend`,
    // profiles/r02_umma_contend.log).  So the epilogue keeps everything it can
    // in integers: the weighted sum stays a 64-bit integer until one conversion
    // per entry, the power-of-two scales are folded into the per-row / per-column
    // factors (exact), clamps and minima compare bit patterns (the distances are
    // non-negative), and the constants of the tile's columns live in registers
    // -- lane c holds column c's -- and travel by shuffle.
    const int q = warp & 3;                   // TMEM lane quarter this warp may read
    const int h = (warp - 2) >> 2;            // column half
    const int r = q * 32 + lane;              // tile row
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr unsigned kFull = 0xffffffffu;
    const int ws = a.wshift;                  // fraction bits kept below the `hi` unit
    const int ls = 8 * (S - 3) - ws;
    uint32_t tcount = 0;
    long long e_wait = 0, e_p1 = 0, e_p2 = 0;
    OzTile tl_next{0, 0, 0, 0};
    if ((int)blockIdx.x < a.n_tiles) tl_next = a.tiles[blockIdx.x];
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
      const OzTile tl = tl_next;
      if (t + (int)gridDim.x < a.n_tiles) tl_next = a.tiles[t + gridDim.x];
      const int64_t i0 = (int64_t)tl.I * OM, j0 = (int64_t)tl.J * ON;
      const int64_t gi = i0 + r;
      const bool row_ok = gi < a.n;
      // this thread's row: scale, statistic; this lane's column (j0 + 32 h + lane,
      // inside the padded planes): the same plus the corrections of chunk 0
      const int64_t gjl = j0 + h * 32 + lane;
      double sci = 0.0, sti = 0.0;
      if (row_ok) {
        sci = __ldg(a.sc + gi);
        sti = __ldg(a.stat + gi);
      }
      const double scl = __ldg(a.sc + gjl), stl = __ldg(a.stat + gjl);
      long long chi = __ldg(a.hi + gjl * a.n_chunks), clo = __ldg(a.lo + gjl * a.n_chunks);
      // factors: cosine / correlation d = 1 - (W f_hi) f_lo with f = stat sc (and the
      // 2^-ws of W on the row side); euclidean d^2 = (W (-2 sc_i 2^-ws)) sc_j + s_hi + s_lo
      const double wsc = __longlong_as_double((long long)(1023 - ws) << 52);
      const double rowf = a.euclid ? __dmul_rn(__dmul_rn(-2.0, sci), wsc)
                                   : __dmul_rn(__dmul_rn(sti, sci), wsc);
      const double colf = a.euclid ? scl : __dmul_rn(stl, scl);
      // output stripes of the tile's rows and of its (mirrored) columns
      int od = 0, om = 0;
      while (od + 1 < a.world && i0 >= a.stripe_begin[od + 1]) ++od;
      while (om + 1 < a.world && j0 >= a.stripe_begin[om + 1]) ++om;
      // this thread's run of 32 entries of row gi, and column 0 of its mirror
      double* __restrict__ drow =
          a.stripe_ptr[od] + (gi - a.stripe_begin[od]) * a.ldd + j0 + h * 32;
      double* __restrict__ mcol =
          a.stripe_ptr[om] + (j0 + h * 32 - a.stripe_begin[om]) * a.ldd + gi;
      const bool direct = (tl.flags & TILE_DIRECT) && gi >= a.row_lo && gi < a.row_hi;
      const bool mirror = (tl.flags & TILE_MIRROR) && row_ok;
      // block minima: this thread's 32 entries are block (j0 + 32 h) / 32 of row gi;
      // column c of the mirrored tile is row j0 + 32 h + c, whose block
      // (i0 + 32 q) / 32 is the minimum over this warp's 32 rows
      double* __restrict__ bm_row = nullptr;
      double* __restrict__ bm_col = nullptr;
      if (a.bmin_ptr[0]) {
        bm_row = a.bmin_ptr[od] + (gi - a.stripe_begin[od]) * a.ldb + (j0 >> 5) + h;
        bm_col = a.bmin_ptr[om] + (j0 + h * 32 - a.stripe_begin[om]) * a.ldb + (i0 >> 5) + q;
      }
      const bool want_cmin = bm_col != nullptr && (tl.flags & TILE_MIRROR);
      constexpr long long kInfBits = 0x7ff0000000000000LL;
      long long rmin = kInfBits;

      // phase 1, once per K chunk: drain the accumulators into the weighted sums
      // (64-bit integers, 32 per thread), four columns at a time, the next four in
      // flight while these are combined; then hand TMEM back so the next chunk's /
      // tile's MMAs run under the rest
      long long pw[32];
      long long ec0 = 0, ec1 = 0;
      auto phase1 = [&](int ch, auto first_chunk) {
        constexpr bool kFirst = decltype(first_chunk)::value;
        if (!kFirst) {
          chi = __ldg(a.hi + gjl * a.n_chunks + ch);
          clo = __ldg(a.lo + gjl * a.n_chunks + ch);
        }
        const long long c0c = ch == a.n_chunks - 1 ? a.c0_last : a.c0;
        const long long hic = (row_ok ? __ldg(a.hi + gi * a.n_chunks + ch) : 0) + c0c;
        const long long loc = row_ok ? __ldg(a.lo + gi * a.n_chunks + ch) : 0;
        ec0 = a.dbg ? clock64() : 0;
        oz_wait(tfull, tcount & 1, a.park_ns);
        ec1 = a.dbg ? clock64() : 0;
        if (a.dbg) e_wait += ec1 - ec0;
        tc::fence_after();
        uint32_t va[S][4], vb[S][4];
        auto drain = [&](uint32_t (&v)[S][4], int c0) {
#pragma unroll
          for (int d = 0; d < S; ++d)
            tc::tmem_ld4_async(lane_base + (uint32_t)(d * ON + h * 32 + c0), v[d]);
        };
        auto settle = [&](uint32_t (&v)[S][4]) {
#pragma unroll
          for (int d = 0; d < S; ++d) tc::tmem_wait4(v[d]);
        };
        auto combine = [&](const uint32_t (&v)[S][4], int c0) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int c = c0 + cc;
            uint32_t acc[S];
#pragma unroll
            for (int d = 0; d < S; ++d) acc[d] = v[d][cc];
            const long long w =
                oz_weighted<S>(acc, hic + __shfl_sync(kFull, chi, c),
                               loc + __shfl_sync(kFull, clo, c), ws, ls);
            pw[c] = kFirst ? w : pw[c] + w;
          }
        };
        drain(va, 0);
        settle(va);
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 8) {
          drain(vb, c0 + 4);
          combine(va, c0);
          settle(vb);
          if (c0 + 8 < 32) drain(va, c0 + 8);
          combine(vb, c0 + 4);
          if (c0 + 8 < 32) settle(va);
        }
        tc::fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty);   // the accumulators may be overwritten
        if (a.dbg) e_p1 += clock64() - ec1;
        ++tcount;
      };
      phase1(0, std::true_type{});
      for (int ch = 1; ch < a.n_chunks; ++ch) phase1(ch, std::false_type{});
      long long ec2 = a.dbg ? clock64() : 0;
      // phase 2 (under the next tile's MMAs): metric epilogue and the two stores.
      // The mirrored tile is written straight from the registers -- lanes are
      // consecutive entries of a row of D; the tile itself in 32-byte runs
      const int cols_ok = (int)(a.n - (j0 + h * 32) < 32 ? a.n - (j0 + h * 32) : 32);  // may be <= 0
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 4) {
        double dv[4];
        long long db[4];       // bit patterns: non-negative doubles order like integers
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = c0 + cc;
          const int64_t gj = j0 + h * 32 + c;
          const double fj = __shfl_sync(kFull, colf, c);
          const double wv = (double)pw[c];
          const bool lower = gi >= gj;
          double d;
          if (a.euclid) {
            const double stj = __shfl_sync(kFull, stl, c);
            const double t2 = __dadd_rn(__fma_rn(__dmul_rn(wv, rowf), fj, lower ? sti : stj),
                                        lower ? stj : sti);
            d = __double_as_longlong(t2) < 0 ? 0.0 : __dsqrt_rn(t2);
          } else {
            const double f_hi = lower ? rowf : fj, f_lo = lower ? fj : rowf;
            d = __dsub_rn(1.0, __dmul_rn(__dmul_rn(wv, f_hi), f_lo));
          }
          long long bits = __double_as_longlong(d);
          bits = (bits < 0 || gi == gj) ? 0 : bits;    // clamp at 0 (also -0); zero diagonal
          db[cc] = bits;
          dv[cc] = __longlong_as_double(bits);
          if (c < cols_ok) {
            if (mirror && gj >= a.row_lo && gj < a.row_hi) __stcs(mcol + (int64_t)c * a.ldd, dv[cc]);
            rmin = bits < rmin ? bits : rmin;
          }
        }
        if (want_cmin) {
          // minima of these four columns over the warp's 32 rows (= one block of
          // four rows of D each): a halving exchange -- after two steps every
          // lane holds one column's partial minimum, three more finish it
          long long w[4];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) w[cc] = row_ok ? db[cc] : kInfBits;
          const bool b4 = lane & 16, b3 = lane & 8;
          auto imin = [](long long x, long long y) { return x < y ? x : y; };
          long long k0 = b4 ? w[2] : w[0], k1 = b4 ? w[3] : w[1];
          k0 = imin(k0, __shfl_xor_sync(kFull, b4 ? w[0] : w[2], 16));
          k1 = imin(k1, __shfl_xor_sync(kFull, b4 ? w[1] : w[3], 16));
          long long m = imin(b3 ? k1 : k0, __shfl_xor_sync(kFull, b3 ? k0 : k1, 8));
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) m = imin(m, __shfl_xor_sync(kFull, m, o));
          const int c = c0 + (b4 ? 2 : 0) + (b3 ? 1 : 0);
          const int64_t gj = j0 + h * 32 + c;
          if ((lane & 7) == 0 && c < cols_ok && gj >= a.row_lo && gj < a.row_hi &&
              i0 + q * 32 < a.n)
            bm_col[(int64_t)c * a.ldb] = __longlong_as_double(m);
        }
        if (direct) {
          if (a.vec_ok && c0 + 3 < cols_ok) {
            st_cs_v4(drow + c0, dv[0], dv[1], dv[2], dv[3]);
          } else {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
              if (c0 + cc < cols_ok) __stcs(drow + c0 + cc, dv[cc]);
          }
        }
      }
      if (bm_row && direct && cols_ok > 0) *bm_row = __longlong_as_double(rmin);
      if (a.dbg) e_p2 += clock64() - ec2;
    }
    if (a.dbg && threadIdx.x == 64) {
      a.dbg[8 * blockIdx.x + 3] = (unsigned long long)e_p1;
      a.dbg[8 * blockIdx.x + 4] = (unsigned long long)e_p2;
      a.dbg[8 * blockIdx.x + 5] = (unsigned long long)e_wait;
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_free512(tmem_base);
}

int oz_group_rows() {
  // row tiles per L2 group: 16 panels of S x 128 x p bytes stay in L2 while the
  // column panels stream once per group
  const char* e = getenv("SCEDAR_B200_OZ_GROUP");
  const int g = e ? atoi(e) : 16;
  return g >= 1 ? g : 16;
}

}  // namespace

// Tile lists live on the device, keyed by matrix size, launch form and range:
// they depend on nothing else, so the handles of successive steps (and the row
// pieces of a pipelined upload, every step) share them.
struct OzTileList {
  int device;
  int64_t n;
  int mode;
  int64_t ka, kb, kc, kd;
  OzTile* dev;
  int count;
};

// working set of the INT8 split of a cells handle (created on first use)
struct OzState {
  int S = 0;
  uint8_t* planes = nullptr;    // [S][n_pl][p_oz]
  double* sc = nullptr;         // [n_pl]
  long long* hi = nullptr;
  long long* lo = nullptr;
  double* stat = nullptr;
  int64_t n_pl = 0, p_oz = 0;
  int n_chunks = 1;
  bool sliced_all = false;
  const struct OzTileList* current = nullptr;   // list of the launch being prepared
  std::vector<OzTile> host_tiles;               // scratch of the list builders
  CUtensorMap map_a, map_b;
};

void oz_state_free(void* p, cudaStream_t s) {
  OzState* o = static_cast<OzState*>(p);
  if (!o) return;
  if (o->planes) cudaFreeAsync(o->planes, s);
  if (o->sc) cudaFreeAsync(o->sc, s);
  if (o->hi) cudaFreeAsync(o->hi, s);
  if (o->lo) cudaFreeAsync(o->lo, s);
  if (o->stat) cudaFreeAsync(o->stat, s);
  delete o;
}

bool oz_supported(const scedar_cells* c) { return c->p > 0; }

namespace {

int oz_encode(CUtensorMap* m, const uint8_t* ptr, int64_t rows, int64_t cols, int box_rows) {
  return encode_tensor_map(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, ptr, rows, cols, box_rows, OKB,
                           CU_TENSOR_MAP_SWIZZLE_128B);
}

int oz_alloc(const scedar_cells* c, int S, cudaStream_t s, OzState** out) {
  scedar_cells* mc = const_cast<scedar_cells*>(c);
  OzState* o = static_cast<OzState*>(mc->oz);
  if (o && o->S == S) {
    *out = o;
    return SCEDAR_OK;
  }
  if (o) {
    oz_state_free(o, s);
    mc->oz = nullptr;
  }
  if (!oz_supported(c))
    return fail(SCEDAR_ERR_UNSUPPORTED, "the INT8 split path needs at least one feature");
  o = new OzState();
  o->S = S;
  o->n_pl = c->n_pad;
  o->p_oz = round_up(c->p, OKB);
  cudaError_t e = cudaMallocAsync(&o->planes, (size_t)S * o->n_pl * o->p_oz, s);
  if (e == cudaSuccess) e = cudaMallocAsync(&o->sc, sizeof(double) * o->n_pl, s);
  o->n_chunks = (int)((o->p_oz + OZ_KCHUNK - 1) / OZ_KCHUNK);
  if (e == cudaSuccess)
    e = cudaMallocAsync(&o->hi, sizeof(long long) * o->n_pl * o->n_chunks, s);
  if (e == cudaSuccess)
    e = cudaMallocAsync(&o->lo, sizeof(long long) * o->n_pl * o->n_chunks, s);
  if (e == cudaSuccess) e = cudaMallocAsync(&o->stat, sizeof(double) * o->n_pl, s);
  if (e != cudaSuccess) {
    cudaGetLastError();
    oz_state_free(o, s);
    return fail(SCEDAR_ERR_NOMEM, std::string("INT8 slice planes: ") + cudaGetErrorString(e));
  }
  int rc = oz_encode(&o->map_a, o->planes, (int64_t)S * o->n_pl, o->p_oz, OM);
  if (rc == SCEDAR_OK) rc = oz_encode(&o->map_b, o->planes, (int64_t)S * o->n_pl, o->p_oz, ON);
  if (rc != SCEDAR_OK) {
    oz_state_free(o, s);
    return rc;
  }
  mc->oz = o;
  *out = o;
  return SCEDAR_OK;
}

// offset x offset term of diagonal 0 for a chunk of `cols` K columns
long long oz_c0(int64_t cols) { return ((long long)16384 * cols) << 16; }

// slices / scales / statistic of the cells [row_begin, row_end) (padding rows allowed)
int oz_slice_rows(const scedar_cells* c, OzState* o, int64_t row_begin, int64_t row_end,
                  cudaStream_t s) {
  if (row_end <= row_begin) return SCEDAR_OK;
  const int euclid = c->metric == SCEDAR_METRIC_EUCLIDEAN;
  const unsigned grid = (unsigned)(row_end - row_begin);
  if (o->S == 6)
    oz_slice_kernel<6><<<grid, 256, 0, s>>>(c->xw, c->n, c->p, c->p_pad, o->p_oz, o->n_pl,
                                            row_begin, euclid, 0, o->planes, o->sc, o->hi,
                                            o->lo, o->stat);
  else
    oz_slice_kernel<7><<<grid, 256, 0, s>>>(c->xw, c->n, c->p, c->p_pad, o->p_oz, o->n_pl,
                                            row_begin, euclid, 0, o->planes, o->sc, o->hi,
                                            o->lo, o->stat);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int ensure_oz(const scedar_cells* c, int S, cudaStream_t s, OzState** out) {
  SCEDAR_CHECK(oz_alloc(c, S, s, out));
  if (!(*out)->sliced_all) {
    SCEDAR_CHECK(oz_slice_rows(c, *out, 0, c->n_pad, s));
    (*out)->sliced_all = true;
  }
  return SCEDAR_OK;
}

std::vector<OzTileList>& oz_lists() {
  static std::vector<OzTileList> lists;
  if (lists.capacity() < 256) lists.reserve(256);   // `current` points into the vector
  return lists;
}

bool oz_have_tiles(const scedar_cells* c, OzState* o, int mode, int64_t ka, int64_t kb,
                   int64_t kc, int64_t kd) {
  for (const auto& l : oz_lists())
    if (l.device == c->device && l.n == c->n && l.mode == mode && l.ka == ka && l.kb == kb &&
        l.kc == kc && l.kd == kd) {
      o->current = &l;
      return true;
    }
  return false;
}

// keep the freshly built host list on the device under its key (unless it is there)
int oz_put_tiles(const scedar_cells* c, OzState* o, int mode, int64_t ka, int64_t kb, int64_t kc,
                 int64_t kd, cudaStream_t s) {
  if (oz_have_tiles(c, o, mode, ka, kb, kc, kd)) return SCEDAR_OK;
  auto& lists = oz_lists();
  if (lists.size() >= 256) {          // many different shapes: start over
    for (auto& l : lists)
      if (l.dev) cudaFreeAsync(l.dev, s);
    lists.clear();
  }
  OzTileList l{c->device, c->n, mode, ka, kb, kc, kd, nullptr, (int)o->host_tiles.size()};
  if (l.count) {
    SCEDAR_CUDA(cudaMallocAsync(&l.dev, sizeof(OzTile) * l.count, s));
    // the copy from pageable memory is staged before the call returns
    SCEDAR_CUDA(cudaMemcpyAsync(l.dev, o->host_tiles.data(), sizeof(OzTile) * l.count,
                                cudaMemcpyHostToDevice, s));
  }
  lists.push_back(l);
  o->current = &lists.back();
  return SCEDAR_OK;
}

int oz_launch(const scedar_cells* c, OzState* o, OzArgs& a, cudaStream_t s, int reserve_sms = 0) {
  a.sc = o->sc;
  a.hi = o->hi;
  a.lo = o->lo;
  a.stat = o->stat;
  a.n_chunks = o->n_chunks;
  a.c0 = oz_c0(std::min<int64_t>(o->p_oz, OZ_KCHUNK));
  a.c0_last = oz_c0(o->p_oz - (int64_t)(o->n_chunks - 1) * OZ_KCHUNK);
  a.n = c->n;
  a.n_pl = o->n_pl;
  a.kslabs = (int)(o->p_oz / OKB);
  a.euclid = c->metric == SCEDAR_METRIC_EUCLIDEAN;
  if (!o->current) return fail(SCEDAR_ERR_CUDA, "no tile list");
  a.tiles = o->current->dev;
  a.n_tiles = o->current->count;
  if (a.n_tiles == 0) return SCEDAR_OK;
  Scratch dbg;
  const bool debug = getenv("SCEDAR_B200_OZ_DEBUG") != nullptr;
  a.dbg = nullptr;
  if (debug) {
    SCEDAR_CHECK(dbg.alloc(sizeof(unsigned long long) * 8 * num_sms(), s));
    SCEDAR_CUDA(cudaMemsetAsync(dbg.ptr, 0, sizeof(unsigned long long) * 8 * num_sms(), s));
    a.dbg = dbg.as<unsigned long long>();
  }
  // block minima ride along when the handle holds one buffer per output stripe
  for (int r = 0; r < kOzPeers; ++r) a.bmin_ptr[r] = nullptr;
  a.ldb = 0;
  if (c->bmin_count == a.world && c->bmin_count > 0) {
    for (int r = 0; r < a.world; ++r) a.bmin_ptr[r] = c->bmin[r];
    a.ldb = c->bmin_ld;
    const_cast<scedar_cells*>(c)->bmin_filled = true;
  }
  a.wshift = oz_wshift(o->n_chunks);
  // parked waits (a suspend-time hint on mbarrier.try_wait) instead of spinning:
  // ~1 % under the power cap (224-229 -> 221-227 ms at 100k x 2000); 0 = spin
  a.park_ns = getenv("SCEDAR_B200_OZ_PARK_NS") ? atoi(getenv("SCEDAR_B200_OZ_PARK_NS")) : 2000;
  a.vec_ok = a.ldd % 4 == 0;
  for (int r = 0; r < a.world; ++r)
    if (reinterpret_cast<uintptr_t>(a.stripe_ptr[r]) % 32 != 0) a.vec_ok = 0;
  int grid = std::min(a.n_tiles, std::max(1, num_sms() - reserve_sms));
  if (const char* e = getenv("SCEDAR_B200_OZ_GRID")) grid = std::max(1, std::min(grid, atoi(e)));
  if (o->S == 6) {
    SCEDAR_CUDA(cudaFuncSetAttribute(gram_oz_kernel<6>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, oz_smem(6)));
    gram_oz_kernel<6><<<grid, OZ_THREADS, oz_smem(6), s>>>(o->map_a, o->map_b, a);
  } else {
    SCEDAR_CUDA(cudaFuncSetAttribute(gram_oz_kernel<7>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, oz_smem(7)));
    gram_oz_kernel<7><<<grid, OZ_THREADS, oz_smem(7), s>>>(o->map_a, o->map_b, a);
  }
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  if (debug) {
    std::vector<unsigned long long> h(8 * (size_t)grid);
    SCEDAR_CUDA(cudaMemcpyAsync(h.data(), dbg.ptr, sizeof(unsigned long long) * h.size(),
                                cudaMemcpyDeviceToHost, s));
    SCEDAR_CUDA(cudaStreamSynchronize(s));
    double sum[8] = {0};
    for (int b = 0; b < grid; ++b)
      for (int q = 0; q < 8; ++q) sum[q] += (double)h[8 * b + q];
    const double tiles = sum[7] > 0 ? sum[7] : 1;
    fprintf(stderr,
            "oz debug: S=%d tiles/CTA %.1f | per tile cycles: total %.0f, MMA warp waits: "
            "tempty %.0f full %.0f | epilogue: wait tfull %.0f phase1 %.0f phase2 %.0f | "
            "ideal MMA %.0f\n",
            o->S, tiles / grid, sum[2] / tiles, sum[0] / tiles, sum[1] / tiles, sum[5] / tiles,
            sum[3] / tiles, sum[4] / tiles,
            (double)a.kslabs * (OKB / 32) * (o->S * (o->S + 1) / 2) * 32.0);
  }
  return SCEDAR_OK;
}

// Tiles of the row tiles [rt0, rt1) in L2-friendly order: groups of G row tiles,
// inside a group column block after column block.  `want(I, J)` says whether this
// launch computes the tile and what it stores; with `lower` only the blocks at
// or left of the diagonal (J / 2 <= I) are visited, every `j_step`-th starting
// at `j_first` (in units of 128-column block pairs).
template <typename Want>
void oz_list(std::vector<OzTile>& out, int64_t rt0, int64_t rt1, int64_t n, bool lower,
             int64_t jr_first, int64_t jr_step, Want want) {
  out.clear();
  const int64_t CT = (n + ON - 1) / ON;
  const int G = oz_group_rows();
  for (int64_t g0 = rt0; g0 < rt1; g0 += G) {
    const int64_t g1 = std::min<int64_t>(rt1, g0 + G);
    const int64_t jr_end = lower ? std::min<int64_t>(g1, (CT + 1) / 2) : (CT + 1) / 2;
    for (int64_t Jr = jr_first; Jr < jr_end; Jr += jr_step)
      for (int64_t J = 2 * Jr; J < std::min<int64_t>(2 * Jr + 2, CT); ++J)
        for (int64_t I = lower ? std::max<int64_t>(g0, Jr) : g0; I < g1; ++I) {
          const int f = want(I, J);
          if (f) out.push_back(OzTile{(int32_t)I, (int32_t)J, f, 0});
        }
  }
}

}  // namespace

int launch_oz_stat(const scedar_cells* c, int S, const double** stat_out, cudaStream_t s) {
  OzState* o = nullptr;
  SCEDAR_CHECK(ensure_oz(c, S, s, &o));
  *stat_out = o->stat;
  return SCEDAR_OK;
}

// rows [row_begin, row_end) of D -> d_out [rows, n]; tile pairs that lie wholly
// inside the stripe's own square block are computed once and mirrored
int launch_oz_stripe(const scedar_cells* c, int S, int64_t row_begin, int64_t row_end,
                     double* d_out, int64_t ldd, cudaStream_t s) {
  if (row_end <= row_begin) return SCEDAR_OK;
  if (row_begin % OM != 0 || (row_end % OM != 0 && row_end != c->n))
    return fail(SCEDAR_ERR_UNSUPPORTED, "INT8 split stripes must be 128-row aligned");
  OzState* o = nullptr;
  SCEDAR_CHECK(ensure_oz(c, S, s, &o));
  const int64_t n = c->n;
  const int64_t rt0 = row_begin / OM, rt1 = (row_end + OM - 1) / OM;
  // row tiles wholly inside the stripe (the stripe may start / end mid-tile)
  auto inside = [&](int64_t T) {
    return T * OM >= row_begin && std::min<int64_t>((T + 1) * OM, n) <= row_end;
  };
  if (!oz_have_tiles(c, o, 0, row_begin, row_end, oz_group_rows(), 0))
  oz_list(o->host_tiles, rt0, rt1, n, false, 0, 1, [&](int64_t I, int64_t J) -> int {
    const int64_t Jr = J / 2;
    if (inside(I) && inside(Jr)) {
      if (Jr > I) return 0;                     // produced as a mirror
      return Jr == I ? TILE_DIRECT : TILE_DIRECT | TILE_MIRROR;
    }
    return TILE_DIRECT;
  });
  SCEDAR_CHECK(oz_put_tiles(c, o, 0, row_begin, row_end, oz_group_rows(), 0, s));
  OzArgs a{};
  a.ldd = ldd;
  a.world = 1;
  // rows outside [row_begin, row_end) are never stored: the stripe table maps
  // row `row_begin` to d_out and the per-row guards below keep the rest out
  a.stripe_begin[0] = row_begin;
  a.stripe_begin[1] = row_end;
  a.stripe_ptr[0] = d_out;
  a.row_lo = row_begin;
  a.row_hi = row_end;
  return oz_launch(c, o, a, s);
}

int launch_oz_symmetric(const scedar_cells* c, int S, double* d_out, int64_t ldd,
                        cudaStream_t s) {
  return launch_oz_stripe(c, S, 0, c->n, d_out, ldd, s);
}

// Peer-store form: every tile pair of the whole matrix is computed by one rank
// -- the owner of one of its two row tiles, alternating with the parity of the
// pair so the ranks' tile counts balance -- and stored into both owners'
// stripes (its own HBM or the peer's over NVLink).  Only owned tiles are
// enumerated: no CTA is launched to find out that it has nothing to do.
int launch_oz_stripe_p2p(const scedar_cells* c, int S, int rank, int world,
                         const int64_t* stripe_begin, double* const* stripe_ptr, int64_t ldd,
                         cudaStream_t s) {
  if (world < 1 || world > kOzPeers) return fail(SCEDAR_ERR_ARG, "world must be in 1..16");
  if (rank < 0 || rank >= world) return fail(SCEDAR_ERR_ARG, "rank out of range");
  const int64_t n = c->n;
  for (int r = 0; r <= world; ++r) {
    const int64_t b = stripe_begin[r];
    if (!(b % OM == 0 || b == n) || (r > 0 && b < stripe_begin[r - 1]))
      return fail(SCEDAR_ERR_ARG, "stripes must be ascending and 128-row aligned");
  }
  if (stripe_begin[0] != 0 || stripe_begin[world] != n)
    return fail(SCEDAR_ERR_ARG, "stripes must cover [0, n)");
  OzState* o = nullptr;
  SCEDAR_CHECK(ensure_oz(c, S, s, &o));
  const int64_t RT = (n + OM - 1) / OM;
  auto owner = [&](int64_t T) {
    int r = 0;
    while (r + 1 < world && T * OM >= stripe_begin[r + 1]) ++r;
    return r;
  };
  if (!oz_have_tiles(c, o, 1, rank, world, stripe_begin[rank], oz_group_rows())) {
    // The pair of row tiles {u, v} is computed by the owner of the higher one
    // when u + v is even, of the lower one when odd (balanced: every rank gets
    // the same number of pairs), and it is ORIENTED so that the computing
    // rank's own tile supplies the 128 rows: the tile itself (32-byte runs per
    // row) then lands in the rank's own HBM and only the mirrored tile (1 KB
    // contiguous per row of D) crosses NVLink.  The rank's row operands are
    // thus always its own stripe -- groups of G of its row tiles stay in L2
    // while the partners' column blocks stream once per group.
    o->host_tiles.clear();
    const int64_t CT = (n + ON - 1) / ON;
    const int64_t u0 = stripe_begin[rank] / OM, u1 = (stripe_begin[rank + 1] + OM - 1) / OM;
    // Groups hold own row tiles of ONE parity: for a partner v of another rank the
    // rule above picks every other u, so a mixed group would bring a column block
    // in for half of its row tiles only.
    const int G = oz_group_rows();
    std::vector<int64_t> us;
    const bool by_parity = getenv("SCEDAR_B200_OZ_P2P_MIXED") == nullptr;
    const bool rotate = getenv("SCEDAR_B200_OZ_P2P_NOROTATE") == nullptr;
    for (int par = 0; par < 2; ++par)
      for (int64_t u = u0; u < u1; ++u)
        if (by_parity ? (u & 1) == par : par == 0) us.push_back(u);
    for (size_t g0 = 0; g0 < us.size(); g0 += G) {
      const size_t g1 = std::min(us.size(), g0 + (size_t)G);
      // every rank starts its sweep over the partners at its own stripe: at any
      // time the ranks then store mirrored tiles into DIFFERENT peers (all of them
      // sweeping from row tile 0 would aim every rank's NVLink traffic at one
      // owner at a time)
      for (int64_t vi = 0; vi < RT; ++vi) {
        const int64_t v = rotate ? (u0 + vi) % RT : vi;
        for (int64_t J = 2 * v; J < std::min<int64_t>(2 * v + 2, CT); ++J)
          for (size_t gi = g0; gi < g1; ++gi) {
            const int64_t u = us[gi];
            if (u == v) {
              o->host_tiles.push_back(OzTile{(int32_t)u, (int32_t)J, TILE_DIRECT, 0});
              continue;
            }
            const int64_t hi = std::max(u, v), lo = std::min(u, v);
            const int by = owner(((hi + lo) & 1) == 0 ? hi : lo);
            // both tiles in this rank's stripe: the lower-triangle orientation
            if (by != rank || (owner(v) == rank && v > u)) continue;
            o->host_tiles.push_back(OzTile{(int32_t)u, (int32_t)J, TILE_DIRECT | TILE_MIRROR, 0});
          }
      }
    }
  }
  SCEDAR_CHECK(oz_put_tiles(c, o, 1, rank, world, stripe_begin[rank], oz_group_rows(), s));
  OzArgs a{};
  a.ldd = ldd;
  a.world = world;
  for (int r = 0; r < world; ++r) {
    a.stripe_begin[r] = stripe_begin[r];
    a.stripe_ptr[r] = stripe_ptr[r];
  }
  a.stripe_begin[world] = stripe_begin[world];
  a.row_lo = 0;
  a.row_hi = n;
  return oz_launch(c, o, a, s);
}

// Rows [row_begin, row_end) of the lower triangle of the WHOLE matrix d_full
// (row 0 first), mirrored: needs only the cells [0, row_end), so D can be
// produced range by range while later cells are still being uploaded.
int launch_oz_lower_rows(const scedar_cells* c, int S, int64_t row_begin, int64_t row_end,
                         double* d_full, int64_t ldd, cudaStream_t s) {
  if (row_end <= row_begin) return SCEDAR_OK;
  if (row_begin % OM != 0) return fail(SCEDAR_ERR_ARG, "row_begin must be 128-aligned");
  OzState* o = nullptr;
  SCEDAR_CHECK(oz_alloc(c, S, s, &o));
  const int64_t n = c->n;
  // this range's cells were just filled: slice them (the last range also the padding)
  const int64_t pad_end = row_end == n ? c->n_pad : round_up(row_end, OM);
  SCEDAR_CHECK(oz_slice_rows(c, o, row_begin, std::min<int64_t>(pad_end, c->n_pad), s));
  if (row_begin == 0 && row_end == n) o->sliced_all = true;
  const int64_t rt0 = row_begin / OM, rt1 = (row_end + OM - 1) / OM;
  if (!oz_have_tiles(c, o, 2, row_begin, row_end, oz_group_rows(), 0))
  oz_list(o->host_tiles, rt0, rt1, std::min<int64_t>(n, rt1 * OM), true, 0, 1, [&](int64_t I, int64_t J) -> int {
    const int64_t Jr = J / 2;
    if (Jr > I) return 0;
    return Jr == I ? TILE_DIRECT : TILE_DIRECT | TILE_MIRROR;
  });
  SCEDAR_CHECK(oz_put_tiles(c, o, 2, row_begin, row_end, oz_group_rows(), 0, s));
  OzArgs a{};
  a.ldd = ldd;
  a.world = 1;
  a.stripe_begin[0] = 0;
  a.stripe_begin[1] = n;
  a.stripe_ptr[0] = d_full;
  a.row_lo = 0;
  a.row_hi = n;
  return oz_launch(c, o, a, s);
}

// Row-range form of the peer-store contraction, for cells that arrive piece by
// piece (the upload of x and its all-gather run behind it): the tiles (I, J) with
// I in the range and J at or left of the diagonal, every `world`-th column block
// taken by this rank (balanced inside every row tile, whatever part of the
// matrix has arrived so far); tile and mirrored tile go to their owners' stripes,
// local or over NVLink.  `reserve_sms` SMs are left to the collective that
// gathers the next piece.
int launch_oz_p2p_rows(const scedar_cells* c, int S, int rank, int world,
                       const int64_t* stripe_begin, double* const* stripe_ptr, int64_t ldd,
                       int64_t row_begin, int64_t row_end, int reserve_sms, cudaStream_t s) {
  if (world < 1 || world > kOzPeers) return fail(SCEDAR_ERR_ARG, "world must be in 1..16");
  if (rank < 0 || rank >= world) return fail(SCEDAR_ERR_ARG, "rank out of range");
  if (row_end <= row_begin) return SCEDAR_OK;
  if (row_begin % OM != 0) return fail(SCEDAR_ERR_ARG, "row_begin must be 128-aligned");
  const int64_t n = c->n;
  for (int r = 0; r <= world; ++r) {
    const int64_t b = stripe_begin[r];
    if (!(b % OM == 0 || b == n) || (r > 0 && b < stripe_begin[r - 1]))
      return fail(SCEDAR_ERR_ARG, "stripes must be ascending and 128-row aligned");
  }
  if (stripe_begin[0] != 0 || stripe_begin[world] != n)
    return fail(SCEDAR_ERR_ARG, "stripes must cover [0, n)");
  OzState* o = nullptr;
  SCEDAR_CHECK(oz_alloc(c, S, s, &o));
  const int64_t pad_end = row_end == n ? c->n_pad : round_up(row_end, OM);
  SCEDAR_CHECK(oz_slice_rows(c, o, row_begin, std::min<int64_t>(pad_end, c->n_pad), s));
  if (row_begin == 0 && row_end == n) o->sliced_all = true;
  const int64_t rt0 = row_begin / OM, rt1 = (row_end + OM - 1) / OM;
  if (!oz_have_tiles(c, o, 3, row_begin, row_end, rank * 64 + world, oz_group_rows()))
    oz_list(o->host_tiles, rt0, rt1, std::min<int64_t>(n, rt1 * OM), true, rank, world,
            [&](int64_t I, int64_t J) -> int {
              return J / 2 == I ? TILE_DIRECT : TILE_DIRECT | TILE_MIRROR;
            });
  SCEDAR_CHECK(
      oz_put_tiles(c, o, 3, row_begin, row_end, rank * 64 + world, oz_group_rows(), s));
  OzArgs a{};
  a.ldd = ldd;
  a.world = world;
  for (int r = 0; r < world; ++r) {
    a.stripe_begin[r] = stripe_begin[r];
    a.stripe_ptr[r] = stripe_ptr[r];
  }
  a.stripe_begin[world] = stripe_begin[world];
  a.row_lo = 0;
  a.row_hi = n;
  return oz_launch(c, o, a, s, reserve_sms);
}

// Piece form of the peer-store contraction, for cells that arrive in `pieces` of
// row tiles taken from EVERY stripe at once (piece_of_tile[T] = the piece that
// brings row tile T; a rank's host shard is its own stripe, so piece q is the
// q-th part of every stripe).  Launch q computes the tile pairs {u, v} whose
// later tile arrives with piece q, under the ownership and orientation of
// launch_oz_stripe_p2p: a rank contracts its own row tiles against everybody's,
// stores the tile into its own stripe and only the mirrored tile over NVLink,
// and every rank gets the same share of every launch.  Called for q = 0, 1, ...
// on every rank it yields the matrix of launch_oz_stripe_p2p bit for bit.
int launch_oz_p2p_pieces(const scedar_cells* c, int S, int rank, int world,
                         const int64_t* stripe_begin, double* const* stripe_ptr, int64_t ldd,
                         const int32_t* piece_of_tile, int piece, int n_pieces, int reserve_sms,
                         cudaStream_t s) {
  if (world < 1 || world > kOzPeers) return fail(SCEDAR_ERR_ARG, "world must be in 1..16");
  if (rank < 0 || rank >= world) return fail(SCEDAR_ERR_ARG, "rank out of range");
  if (piece < 0 || piece >= n_pieces) return fail(SCEDAR_ERR_ARG, "piece out of range");
  const int64_t n = c->n;
  for (int r = 0; r <= world; ++r) {
    const int64_t b = stripe_begin[r];
    if (!(b % OM == 0 || b == n) || (r > 0 && b < stripe_begin[r - 1]))
      return fail(SCEDAR_ERR_ARG, "stripes must be ascending and 128-row aligned");
  }
  if (stripe_begin[0] != 0 || stripe_begin[world] != n)
    return fail(SCEDAR_ERR_ARG, "stripes must cover [0, n)");
  const int64_t RT = (n + OM - 1) / OM;
  for (int64_t T = 0; T < RT; ++T)
    if (piece_of_tile[T] < 0 || piece_of_tile[T] >= n_pieces)
      return fail(SCEDAR_ERR_ARG, "piece_of_tile entries must be in [0, n_pieces)");
  OzState* o = nullptr;
  SCEDAR_CHECK(oz_alloc(c, S, s, &o));
  // slice the row tiles this piece brought (runs of consecutive tiles; the last
  // tile of the matrix takes the padding rows with it)
  for (int64_t T = 0; T < RT;) {
    if (piece_of_tile[T] != piece) {
      ++T;
      continue;
    }
    int64_t T1 = T;
    while (T1 < RT && piece_of_tile[T1] == piece) ++T1;
    const int64_t e = T1 == RT ? c->n_pad : T1 * OM;
    SCEDAR_CHECK(oz_slice_rows(c, o, T * OM, std::min<int64_t>(e, c->n_pad), s));
    T = T1;
  }
  if (n_pieces == 1) o->sliced_all = true;
  auto owner = [&](int64_t T) {
    int r = 0;
    while (r + 1 < world && T * OM >= stripe_begin[r + 1]) ++r;
    return r;
  };
  // key: the layout is a function of (n, world, n_pieces) for the callers in
  // this package; a different table under the same key would need its own mode
  long long key = 1469598103934665603ll;
  for (int64_t T = 0; T < RT; ++T) key = (key ^ piece_of_tile[T]) * 1099511628211ll;
  if (!oz_have_tiles(c, o, 4, rank * 64 + world, piece * 64 + n_pieces, key, oz_group_rows())) {
    o->host_tiles.clear();
    const int64_t CT = (n + ON - 1) / ON;
    const int64_t u0 = stripe_begin[rank] / OM, u1 = (stripe_begin[rank + 1] + OM - 1) / OM;
    const int G = oz_group_rows();
    std::vector<int64_t> us;
    for (int par = 0; par < 2; ++par)
      for (int64_t u = u0; u < u1; ++u)
        if ((u & 1) == par && piece_of_tile[u] <= piece) us.push_back(u);
    for (size_t g0 = 0; g0 < us.size(); g0 += G) {
      const size_t g1 = std::min(us.size(), g0 + (size_t)G);
      for (int64_t vi = 0; vi < RT; ++vi) {
        const int64_t v = (u0 + vi) % RT;
        if (piece_of_tile[v] > piece) continue;
        for (int64_t J = 2 * v; J < std::min<int64_t>(2 * v + 2, CT); ++J)
          for (size_t gi = g0; gi < g1; ++gi) {
            const int64_t u = us[gi];
            if (std::max(piece_of_tile[u], piece_of_tile[v]) != piece) continue;
            if (u == v) {
              o->host_tiles.push_back(OzTile{(int32_t)u, (int32_t)J, TILE_DIRECT, 0});
              continue;
            }
            const int64_t hi = std::max(u, v), lo = std::min(u, v);
            const int by = owner(((hi + lo) & 1) == 0 ? hi : lo);
            if (by != rank || (owner(v) == rank && v > u)) continue;
            o->host_tiles.push_back(OzTile{(int32_t)u, (int32_t)J, TILE_DIRECT | TILE_MIRROR, 0});
          }
      }
    }
  }
  SCEDAR_CHECK(oz_put_tiles(c, o, 4, rank * 64 + world, piece * 64 + n_pieces, key,
                            oz_group_rows(), s));
  OzArgs a{};
  a.ldd = ldd;
  a.world = world;
  for (int r = 0; r < world; ++r) {
    a.stripe_begin[r] = stripe_begin[r];
    a.stripe_ptr[r] = stripe_ptr[r];
  }
  a.stripe_begin[world] = stripe_begin[world];
  a.row_lo = 0;
  a.row_hi = n;
  return oz_launch(c, o, a, s, reserve_sms);
}

}  // namespace scedar
