// K2 (FP64 path) -- Gram contraction X.Xt on the float64 tensor-core pipe
// (mma.sync m8n8k4 f64, "DMMA") with the metric epilogue fused, and the fused
// streaming top-k (K3) that consumes tiles without writing D.
//
// tcgen05.mma has no f64 kind (CUDA 12.9: f16/tf32/i8/f8f6f4/mx*), so the path
// that must match the reference's float64 to 1e-10 runs on DMMA; the tcgen05
// path is the split-BF16 fast path in gram_tc.cu.
//
// Reference arithmetic replaced (scedar/eda/sdm.py):
//   np.dot(x, x.T)                                   1245-1247
//   inv = sqrt(1/ss or 0); (P*inv).T*inv; 1 - S       1249-1266
//   sklearn euclidean: (-2P + |x_i|^2) + |x_j|^2, max 0, sqrt      1216
//   num_correct_dist_mat: <0 -> 0, diag 0, upper := lower           215-221
// The lower triangle survives the mirror, so entry (i, j) is evaluated as the
// reference evaluates (max(i,j), min(i,j)): scale by the statistic of the
// larger index first, then the smaller -- bit-identical on both sides of the
// diagonal by construction.  Products are kept out of FMA contraction with
// __dmul_rn/__dadd_rn so exact ties in integer data stay exact ties.
//
// Tiling: CTA 128x128; 8 consumer warps (2 x 4), warp tile 64x32 = 8x4 DMMA
// fragments, plus one producer warp whose elected lane streams K slabs of 16
// doubles (one 128-byte line per row) with TMA (cp.async.bulk.tensor, 128B
// swizzle) through a 4-stage ring of full/empty mbarriers -- no CTA-wide
// barrier and no per-thread copy instructions in the main loop.  Fragment rows
// are mapped to tile rows by perm8() so the swizzled 64-bit fragment loads of
// a half-warp (4 rows x 4 k) hit 16 distinct 8-byte banks without padding.
// Roofline: FP64 tensor pipe; algorithmic flops 2*rows*n*p per stripe
// (n(n+1)p for the symmetric variant).  HBM traffic is bounded by the tile
// swizzle (16 row panels stay in L2 while the column panels stream once).
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "topk_list.cuh"

namespace scedar {
namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int WARPS_M = 2, WARPS_N = 4;       // consumer warp grid
constexpr int NCONS = WARPS_M * WARPS_N * 32; // 256 threads issue LDS + DMMA
constexpr int NT = NCONS + 32;                // + one TMA producer warp
constexpr int MF = 8, NF = 4;                 // 8x8 fragments per warp tile
constexpr int WTM = MF * 8, WTN = NF * 8;     // warp tile 64 x 32
constexpr int PANEL_DOUBLES = BM * BK;        // one TMA box: 128 rows x 16 doubles
constexpr int STAGE_DOUBLES = 2 * PANEL_DOUBLES;
constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;  // 32 KB: A panel + B panel
constexpr int STAGES = 4;
constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
constexpr int LDT = 129;  // odd stride: conflict-free row and column sweeps
constexpr int TILE_DOUBLES = BM * LDT;
constexpr int TILE_BYTES = TILE_DOUBLES * 8;
constexpr int MAIN_BYTES = PIPE_BYTES > TILE_BYTES ? PIPE_BYTES : TILE_BYTES;
constexpr int BAR_BYTES = 128;
constexpr int KFUSED = 32;  // entries kept per row by the fused top-k
constexpr int LIST_BYTES = BM * KFUSED * (8 + 4);
constexpr int SMEM_BASE = 1024 + MAIN_BYTES + BAR_BYTES;  // + LIST_BYTES for TOPK
constexpr int GROUP_ROWS = 16;  // row tiles per L2 swizzle group (32 MB of row panels at p = 2000;
                                // 32 tiles thrash L2: 160 GB of DRAM reads instead of 65 at n = 100k)
constexpr int kMaxPeers = 16;

enum Mode { MODE_STORE = 0, MODE_DIAG = 2, MODE_TOPK = 3, MODE_RAW = 4 };  // RAW: STORE of the bare dot products

struct GramArgs {
  const double* xw;
  const double* stat;
  int64_t n, n_pad, p_pad;
  int64_t row_begin, row_end;  // rows of D this launch owns
  double* out;                 // D stripe (STORE) or stat (DIAG)
  int64_t ldd;
  int euclid;
  int mirror;  // STORE: tiles above the diagonal of the stripe's own square
               // block are not computed but written as transposes
  const int64_t* row_ids;  // STORE over gathered rows: the A panels come from a second
                           // matrix (its own tensor map) whose row r is cell
                           // row_ids[r] (-1: padding); D rows stay in gathered order
  int lower;   // STORE: `out` is the WHOLE matrix (row 0 first); only tiles with
               // J <= I are computed and every one is mirrored -- the rows of D
               // can then be produced range by range, in ascending order, while
               // later cells are still being uploaded
  // STORE over peer memory (world > 1): every off-diagonal tile pair of the
  // WHOLE matrix is computed once, by the owner of one of its two row tiles,
  // and its transpose is stored straight into the other owner's stripe
  int world;
  int64_t stripe_begin[kMaxPeers + 1];  // tile-aligned row ranges of the ranks
  double* stripe_ptr[kMaxPeers];        // their D stripes, mapped on this GPU
  // TOPK
  int K, splits;
  double* part_d;
  int32_t* part_i;
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// Fragment row g of an 8-row group lives in tile row perm8(g): with the
// 128-byte TMA swizzle (16-byte chunk index XOR row%8) the 16 lanes of a
// half-warp (4 fragment rows x 4 k) then read 16 distinct 8-byte banks.
__device__ __forceinline__ int perm8(int g) { return ((g & 3) << 1) | (g >> 2); }

// Consumer side of the pipeline: acc += A(i0.., :) . B(j0.., :)^T over the whole
// padded K range.  `it` counts K slabs across tiles (stage = it % STAGES).
// DIAG_ONLY (the |x|^2 pass): only the 8x8 fragments on the tile's diagonal are
// contracted -- 4 of the 32 fragments of 4 of the 8 warps -- with the same
// DMMA sequence per element as the full tile, so the statistic stays
// bit-identical to the Gram diagonal while the pass becomes load-bound.
template <bool DIAG_ONLY>
__device__ __forceinline__ void gram_consume(double (&acc)[MF][NF][2], const double* pipe,
                                             uint32_t full0, uint32_t empty0, int KT,
                                             uint32_t& it) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N, g = lane >> 2, t = lane & 3;
  const int pa = perm8(g);
  // swizzled byte offset of k = 4*kk + t inside a 16-double row: off0 ^ (32 * kk)
  const uint32_t off0 = (uint32_t)(((((t >> 1) ^ pa) << 1) | (t & 1)) * 8);
  // 32-bit shared-window addresses: the accumulators leave no room for 64-bit
  // pointer arithmetic in the main loop
  const uint32_t pipe32 = smem_u32(pipe);
  const uint32_t Arow = pipe32 + (uint32_t)((wm * WTM + pa) * BK * 8);
  const uint32_t Brow = pipe32 + (uint32_t)((PANEL_DOUBLES + (wn * WTN + pa) * BK) * 8);
#pragma unroll
  for (int mf = 0; mf < MF; ++mf)
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;

  for (int kt = 0; kt < KT; ++kt, ++it) {
    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
    mbar_wait_lite(full0 + 8 * s, ph);   // the TMA boxes of this slab have landed
    const uint32_t As = Arow + s * STAGE_BYTES;
    const uint32_t Bs = Brow + s * STAGE_BYTES;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[MF], b[NF];
      const uint32_t ok = off0 ^ (uint32_t)(32 * kk);
#pragma unroll
      for (int mf = 0; mf < MF; ++mf) a[mf] = lds_f64(As + ok + mf * 8 * BK * 8);
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) b[nf] = lds_f64(Bs + ok + nf * 8 * BK * 8);
      if (DIAG_ONLY) {
        // rows wm*64 + mf*8 meet columns wn*32 + nf*8 when mf == (wn & 1) * 4 + nf
        // in the warps with wn / 2 == wm
        if ((wn >> 1) == wm) {
#pragma unroll
          for (int nf = 0; nf < NF; ++nf) {
            if (wn & 1)
              dmma(acc[4 + nf][nf], a[4 + nf], b[nf]);
            else
              dmma(acc[nf][nf], a[nf], b[nf]);
          }
        }
      } else {
#pragma unroll
        for (int mf = 0; mf < MF; ++mf)
#pragma unroll
          for (int nf = 0; nf < NF; ++nf) dmma(acc[mf][nf], a[mf], b[nf]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * s);   // this warp is done with the slab
  }
}

// Producer side (one lane of the producer warp): stream the K slabs of row
// panels i0 and j0 through the stage ring.
__device__ __forceinline__ void gram_produce(const CUtensorMap* amap, const CUtensorMap* map,
                                             uint32_t pipe_u32, uint32_t full0, uint32_t empty0,
                                             int KT, int64_t i0, int64_t j0, uint32_t& it) {
  for (int kt = 0; kt < KT; ++kt, ++it) {
    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
    mbar_wait(empty0 + 8 * s, ph ^ 1);
    mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
    const uint32_t st = pipe_u32 + s * STAGE_BYTES;
    tma_load_2d(st, amap, kt * BK, (int)i0, full0 + 8 * s);
    tma_load_2d(st + PANEL_DOUBLES * 8, map, kt * BK, (int)j0, full0 + 8 * s);
  }
}

__device__ __forceinline__ void consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory");
}

// accumulators -> metric epilogue -> smem tile T[r][c] (stride LDT)
template <bool RAW>
__device__ __forceinline__ void tile_to_smem(const double (&acc)[8][4][2], double* T,
                                             const double* __restrict__ stat,
                                             int64_t i0, int64_t j0, int euclid,
                                             const int64_t* __restrict__ row_ids = nullptr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N, g = lane >> 2, t = lane & 3;
  double sc[4][2];
  if (!RAW) {
#pragma unroll
    for (int nf = 0; nf < 4; ++nf) {
      sc[nf][0] = __ldg(stat + j0 + wn * 32 + nf * 8 + perm8(2 * t));
      sc[nf][1] = __ldg(stat + j0 + wn * 32 + nf * 8 + perm8(2 * t + 1));
    }
  }
#pragma unroll
  for (int mf = 0; mf < 8; ++mf) {
    const int r = wm * 64 + mf * 8 + perm8(g);   // fragment row -> tile row
    // cell index of the row (gathered rows: through row_ids) and its statistic,
    // fetched row by row: keeping all eight in registers spills
    int64_t gi = i0 + r;
    double srm = 0.0;
    if (!RAW) {
      if (row_ids) gi = row_ids[gi];
      if (gi >= 0) srm = __ldg(stat + gi);
    }
#pragma unroll
    for (int nf = 0; nf < 4; ++nf) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = wn * 32 + nf * 8 + perm8(2 * t + e);
        double v = acc[mf][nf][e];
        if (!RAW) {
          const int64_t gj = j0 + c;
          const bool lower = gi >= gj;
          v = finish_entry(v, lower ? srm : sc[nf][e], lower ? sc[nf][e] : srm, gi == gj,
                           euclid);
        }
        T[r * LDT + c] = v;
      }
    }
  }
}

// ---- warp-resident sorted list (one entry per lane), order (d, idx) --------
__device__ __forceinline__ void warp_insert(double& ld, int& li, double d, int j, int lane) {
  const unsigned less = __ballot_sync(0xffffffffu, before(ld, li, d, j));
  const int pos = __popc(less);
  const double pd = __shfl_up_sync(0xffffffffu, ld, 1);
  const int pi = __shfl_up_sync(0xffffffffu, li, 1);
  if (lane > pos) {
    ld = pd;
    li = pi;
  } else if (lane == pos) {
    ld = d;
    li = j;
  }
}

template <int MODE>
__global__ void __launch_bounds__(NT, 1)
gram_f64_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap amap,
                const GramArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // the TMA swizzle needs 1 KB alignment
  double* smem = reinterpret_cast<double*>(smem_raw + (base - raw));  // pipeline / epilogue tile
  const uint32_t full0 = base + MAIN_BYTES, empty0 = full0 + 8 * STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp == NCONS / 32;
  const int64_t CT = a.n_pad / BN;
  const int KT = (int)(a.p_pad / BK);

  // ---- tile coordinates (uniform over the CTA) -----------------------------
  int64_t I, J = 0, J0 = 0, J1 = 1;
  bool mirrored = false;
  double* mirror_out = a.out;       // stripe that receives the transposed tile
  int64_t mirror_begin = a.row_begin;
  if (MODE == MODE_TOPK) {
    // this CTA owns row tile I and the column tiles [J0, J1)
    I = a.row_begin / BM + blockIdx.x;
    J0 = CT * blockIdx.y / a.splits;
    J1 = CT * (blockIdx.y + 1) / a.splits;
  } else if (MODE == MODE_DIAG) {
    I = J0 = a.row_begin / BM + blockIdx.x;
    J1 = J0 + 1;
  } else {
    const int64_t rt0 = a.row_begin / BM;
    const int64_t RT = (a.row_end + BM - 1) / BM - rt0;
    const int64_t per_group = (int64_t)GROUP_ROWS * CT;
    const int64_t grp = blockIdx.x / per_group, rem = blockIdx.x % per_group;
    const int64_t left = RT - grp * GROUP_ROWS;
    const int64_t rows_here = left < GROUP_ROWS ? left : GROUP_ROWS;
    I = rt0 + grp * GROUP_ROWS + rem % rows_here;
    J = rem / rows_here;
    if (a.lower) {
      if (J > I) return;
      mirrored = J != I;
      mirror_begin = 0;
    } else if (a.world > 1 && I != J) {
      // balanced ownership of the pair {I, J}: the higher row tile computes it
      // when I+J is even, the lower one when odd
      const int64_t lo = I < J ? I : J, hi = I < J ? J : I;
      const int64_t row_tile = ((lo + hi) & 1) == 0 ? hi : lo;
      if (I != row_tile) return;
      mirrored = true;
      int o = 0;
      while (o + 1 < a.world && J * BN >= a.stripe_begin[o + 1]) ++o;
      mirror_out = a.stripe_ptr[o];
      mirror_begin = a.stripe_begin[o];
    } else if (a.mirror && I != J) {
      // both tiles lie wholly inside this stripe's rows: D[I,J] == D[J,I]^T is
      // produced once, by the lower tile
      const int64_t ie = (I + 1) * BM < a.n ? (I + 1) * BM : a.n;
      const int64_t je = (J + 1) * BN < a.n ? (J + 1) * BN : a.n;
      mirrored = I * BM >= a.row_begin && ie <= a.row_end &&
                 J * BN >= a.row_begin && je <= a.row_end;
      if (mirrored && J > I) return;
    }
    J0 = J;
    J1 = J + 1;
  }
  const int64_t i0 = I * BM;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);                // the producer's expect_tx arrive
      mbar_init(empty0 + 8 * s, NCONS / 32);      // one arrive per consumer warp
    }
    mbar_fence_init();
  }
  double* Ld = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(smem) + MAIN_BYTES + BAR_BYTES);
  int* Li = reinterpret_cast<int*>(Ld + BM * KFUSED);
  if (MODE == MODE_TOPK) {
    for (int e = tid; e < BM * KFUSED; e += NT) {
      Ld[e] = __longlong_as_double(0x7ff0000000000000LL);  // +inf
      Li[e] = 0x7fffffff;
    }
  }
  __syncthreads();

  uint32_t it = 0;  // K slabs streamed so far (same count on both sides)
  for (J = J0; J < J1; ++J) {
    const int64_t j0 = J * BN;
    if (producer) {
      if (lane == 0) gram_produce(&amap, &xmap, base, full0, empty0, KT, i0, j0, it);
      __syncwarp();
    } else {
      double acc[MF][NF][2];
      gram_consume<MODE == MODE_DIAG>(acc, smem, full0, empty0, KT, it);
      consumer_sync();  // every slab is consumed: the epilogue tile may alias the ring

      if (MODE == MODE_DIAG) {
        tile_to_smem<true>(acc, smem, nullptr, i0, j0, 0);
        consumer_sync();
        if (tid < BM) {
          const double ss = smem[tid * LDT + tid];
          double sv;
          if (a.euclid) {
            sv = ss;  // |x|^2
          } else {
            // sqrt(1/ss), 0 for a zero vector (sdm.py:1259-1261)
            sv = __dsqrt_rn(ss != 0.0 ? __ddiv_rn(1.0, ss) : 0.0);
          }
          a.out[i0 + tid] = sv;
        }
      } else {
        tile_to_smem<MODE == MODE_RAW>(acc, smem, a.stat, i0, j0, a.euclid, a.row_ids);
        consumer_sync();
        if (MODE == MODE_TOPK) {
          // warp w scans rows w, w+8, ...; lane l holds entry l of the row's list
          const int K = a.K;
          for (int r = warp; r < BM; r += NCONS / 32) {
            const int64_t gi = i0 + r;
            if (gi < a.row_begin || gi >= a.row_end) continue;
            double ld = Ld[r * KFUSED + lane];
            int li = Li[r * KFUSED + lane];
            double td = __shfl_sync(0xffffffffu, ld, K - 1);
            int tj = __shfl_sync(0xffffffffu, li, K - 1);
            bool dirty = false;
#pragma unroll
            for (int q = 0; q < BN / 32; ++q) {
              const int c = lane + 32 * q;
              const int64_t gj = j0 + c;
              const double v = smem[r * LDT + c];
              unsigned m = __ballot_sync(0xffffffffu, gj < a.n && before(v, (int)gj, td, tj));
              while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const double d = __shfl_sync(0xffffffffu, v, src);
                const int j = (int)(j0 + src + 32 * q);
                if (before(d, j, td, tj)) {
                  warp_insert(ld, li, d, j, lane);
                  td = __shfl_sync(0xffffffffu, ld, K - 1);
                  tj = __shfl_sync(0xffffffffu, li, K - 1);
                  dirty = true;
                }
              }
            }
            if (dirty) {
              Ld[r * KFUSED + lane] = ld;
              Li[r * KFUSED + lane] = li;
            }
          }
        } else {
          // rows of the tile -> rows of D (256-byte coalesced warp stores)
          for (int r = warp; r < BM; r += NCONS / 32) {
            const int64_t gi = i0 + r;
            if (gi < a.row_begin || gi >= a.row_end) continue;
            double* __restrict__ dst = a.out + (gi - (a.lower ? 0 : a.row_begin)) * a.ldd + j0;
#pragma unroll
            for (int q = 0; q < BN / 32; ++q) {
              const int c = lane + 32 * q;
              if (j0 + c < a.n) __stcs(dst + c, smem[r * LDT + c]);   // D is write-once: evict first
            }
          }
          if (mirrored) {
            // mirrored tile: column c of T is row j0+c of D
            for (int c = warp; c < BN; c += NCONS / 32) {
              const int64_t gj = j0 + c;
              if (gj >= a.n) continue;
              double* __restrict__ dst = mirror_out + (gj - mirror_begin) * a.ldd + i0;
#pragma unroll
              for (int q = 0; q < BM / 32; ++q) {
                const int r = lane + 32 * q;
                if (i0 + r < a.n) __stcs(dst + r, smem[r * LDT + c]);
              }
            }
          }
        }
      }
    }
    // the next tile's slabs overwrite the epilogue tile: everyone waits
    if (MODE == MODE_TOPK) __syncthreads();
  }

  if (MODE == MODE_TOPK) {
    // partial lists out: part[split][row - row_begin][K]
    const int K = a.K;
    const int64_t rows = a.row_end - a.row_begin;
    for (int e = tid; e < BM * K; e += NT) {
      const int r = e / K, l = e % K;
      const int64_t gi = i0 + r;
      if (gi < a.row_begin || gi >= a.row_end) continue;
      const int64_t o = ((int64_t)blockIdx.y * rows + (gi - a.row_begin)) * K + l;
      a.part_d[o] = Ld[r * KFUSED + l];
      a.part_i[o] = Li[r * KFUSED + l];
    }
  }
}

template <int MODE>
int launch(const scedar_cells* c, const GramArgs& a, dim3 grid, cudaStream_t s,
           const CUtensorMap* amap = nullptr) {
  const int smem_bytes = SMEM_BASE + (MODE == MODE_TOPK ? LIST_BYTES : 0);
  SCEDAR_CUDA(cudaFuncSetAttribute(gram_f64_kernel<MODE>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  gram_f64_kernel<MODE><<<grid, NT, smem_bytes, s>>>(c->xmap, amap ? *amap : c->xmap, a);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

GramArgs base_args(const scedar_cells* c) {
  GramArgs a{};
  a.xw = c->xw;
  a.stat = c->stat;
  a.n = c->n;
  a.n_pad = c->n_pad;
  a.p_pad = c->p_pad;
  a.euclid = c->metric == SCEDAR_METRIC_EUCLIDEAN;
  return a;
}

}  // namespace

int launch_gram_diag(const scedar_cells* c, cudaStream_t s) {
  return launch_gram_diag_rows(c, 0, c->n_pad, s);
}

// statistic of the cells [row_begin, row_end) only (128-aligned, padded rows allowed)
int launch_gram_diag_rows(const scedar_cells* c, int64_t row_begin, int64_t row_end,
                          cudaStream_t s) {
  if (row_end <= row_begin) return SCEDAR_OK;
  GramArgs a = base_args(c);
  a.row_begin = row_begin;
  a.row_end = row_end;
  a.out = c->stat;
  return launch<MODE_DIAG>(c, a, dim3((unsigned)((row_end - row_begin + BM - 1) / BM)), s);
}

int launch_gram_lower_rows(const scedar_cells* c, int64_t row_begin, int64_t row_end,
                           double* d_full, int64_t ldd, cudaStream_t s) {
  if (row_end <= row_begin) return SCEDAR_OK;
  GramArgs a = base_args(c);
  a.row_begin = row_begin;
  a.row_end = row_end;
  a.out = d_full;
  a.ldd = ldd;
  a.lower = 1;
  const int64_t RT = (row_end + BM - 1) / BM - row_begin / BM;
  // column tiles beyond the last row tile of the range are never needed
  const int64_t CT = (row_end + BN - 1) / BN;
  const int64_t tiles = RT * CT;
  if (tiles > 0x7fffffffLL) return fail(SCEDAR_ERR_UNSUPPORTED, "range too large for one launch");
  a.n_pad = CT * BN;   // the tile swizzle walks CT column tiles
  return launch<MODE_STORE>(c, a, dim3((unsigned)tiles), s);
}

int launch_gram_stripe(const scedar_cells* c, int64_t row_begin, int64_t row_end,
                       double* d_out, int64_t ldd, cudaStream_t s) {
  if (row_end <= row_begin) return SCEDAR_OK;
  GramArgs a = base_args(c);
  a.row_begin = row_begin;
  a.row_end = row_end;
  a.out = d_out;
  a.ldd = ldd;
  a.mirror = 1;
  const int64_t RT = (row_end + BM - 1) / BM - row_begin / BM;
  const int64_t tiles = RT * (c->n_pad / BN);
  if (tiles > 0x7fffffffLL) return fail(SCEDAR_ERR_UNSUPPORTED, "stripe too large for one launch");
  return launch<MODE_STORE>(c, a, dim3((unsigned)tiles), s);
}

int launch_gram_stripe_p2p(const scedar_cells* c, int rank, int world,
                           const int64_t* stripe_begin, double* const* stripe_ptr,
                           int64_t ldd, cudaStream_t s) {
  if (world < 1 || world > kMaxPeers) return fail(SCEDAR_ERR_ARG, "world must be in 1..16");
  if (rank < 0 || rank >= world) return fail(SCEDAR_ERR_ARG, "rank out of range");
  for (int r = 0; r <= world; ++r) {
    const int64_t b = stripe_begin[r];
    const bool aligned = b % BM == 0 || b == c->n;
    if (!aligned || (r > 0 && b < stripe_begin[r - 1]))
      return fail(SCEDAR_ERR_ARG, "stripes must be ascending and 128-row aligned");
  }
  if (stripe_begin[0] != 0 || stripe_begin[world] != c->n)
    return fail(SCEDAR_ERR_ARG, "stripes must cover [0, n)");
  const int64_t row_begin = stripe_begin[rank], row_end = stripe_begin[rank + 1];
  if (row_end <= row_begin) return SCEDAR_OK;
  GramArgs a = base_args(c);
  a.row_begin = row_begin;
  a.row_end = row_end;
  a.out = stripe_ptr[rank];
  a.ldd = ldd;
  a.mirror = 1;  // world == 1 degenerates to the local mirror
  a.world = world;
  for (int r = 0; r < world; ++r) {
    a.stripe_begin[r] = stripe_begin[r];
    a.stripe_ptr[r] = stripe_ptr[r];
  }
  a.stripe_begin[world] = stripe_begin[world];
  const int64_t RT = (row_end + BM - 1) / BM - row_begin / BM;
  const int64_t tiles = RT * (c->n_pad / BN);
  if (tiles > 0x7fffffffLL) return fail(SCEDAR_ERR_UNSUPPORTED, "stripe too large for one launch");
  return launch<MODE_STORE>(c, a, dim3((unsigned)tiles), s);
}

// D rows of an arbitrary set of cells: g [m_pad, p_pad] holds the working-copy
// rows of the cells row_ids[0..m) (zero rows / -1 beyond m, m_pad a multiple of
// 128); d_out [m, n] row r = distances of cell row_ids[r] to every cell, each
// entry bit-identical to the same entry of the whole matrix.
int launch_gram_gathered(const scedar_cells* c, const double* g, const int64_t* row_ids,
                         int64_t m, int64_t m_pad, double* d_out, int64_t ldd, cudaStream_t s) {
  if (m <= 0) return SCEDAR_OK;
  CUtensorMap gmap;
  SCEDAR_CHECK(encode_tensor_map(&gmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, g, m_pad, c->p_pad,
                                 128, 16));
  GramArgs a = base_args(c);
  a.row_begin = 0;
  a.row_end = m;
  a.out = d_out;
  a.ldd = ldd;
  a.row_ids = row_ids;
  const int64_t RT = (m + BM - 1) / BM;
  const int64_t tiles = RT * (c->n_pad / BN);
  if (tiles > 0x7fffffffLL) return fail(SCEDAR_ERR_UNSUPPORTED, "too many rows for one launch");
  return launch<MODE_STORE>(c, a, dim3((unsigned)tiles), s, &gmap);
}

// Bare Gram matrix X.Xt of the prepared rows [n, n] (no metric epilogue): lower
// tiles contracted once and mirrored.  The PCA producer calls it on the
// transposed cells (features x samples), where it is the scatter matrix XtX.
int launch_gram_raw_symmetric(const scedar_cells* c, double* out, int64_t ldd, cudaStream_t s) {
  if (c->n <= 0) return SCEDAR_OK;
  GramArgs a = base_args(c);
  a.row_begin = 0;
  a.row_end = c->n;
  a.out = out;
  a.ldd = ldd;
  a.mirror = 1;
  const int64_t RT = (c->n + BM - 1) / BM;
  const int64_t tiles = RT * (c->n_pad / BN);
  if (tiles > 0x7fffffffLL) return fail(SCEDAR_ERR_UNSUPPORTED, "matrix too large for one launch");
  return launch<MODE_RAW>(c, a, dim3((unsigned)tiles), s);
}

// Bare cross products A.Bt: the rows of `ca` (its working copy, all n_a of them)
// against the rows of `cb`; out [n_a, n_b].  Both must have the same number of
// columns.  Tall-skinny use: n_b <= a few hundred (the projection on principal axes).
int launch_gram_raw_cross(const scedar_cells* ca, const scedar_cells* cb, double* out,
                          int64_t ldd, cudaStream_t s) {
  if (ca->n <= 0 || cb->n <= 0) return SCEDAR_OK;
  if (ca->p_pad != cb->p_pad) return fail(SCEDAR_ERR_ARG, "operands differ in their column count");
  GramArgs a = base_args(cb);
  a.row_begin = 0;
  a.row_end = ca->n;
  a.out = out;
  a.ldd = ldd;
  const int64_t RT = (ca->n + BM - 1) / BM;
  const int64_t tiles = RT * (cb->n_pad / BN);
  if (tiles > 0x7fffffffLL) return fail(SCEDAR_ERR_UNSUPPORTED, "too many rows for one launch");
  return launch<MODE_RAW>(cb, a, dim3((unsigned)tiles), s, &ca->xmap);
}

int launch_gram_symmetric(const scedar_cells* c, double* d_out, int64_t ldd,
                          cudaStream_t s) {
  // the stripe [0, n) is the whole matrix: every upper tile is a mirror
  return launch_gram_stripe(c, 0, c->n, d_out, ldd, s);
}

int gram_topk_splits(const scedar_cells* c, int64_t rows) {
  // enough CTAs for ~2 waves of the SMs, at most one split per column tile
  const int64_t row_tiles = (rows + BM - 1) / BM + 1;
  const int64_t CT = c->n_pad / BN;
  int64_t s = (2 * (int64_t)num_sms() + row_tiles - 1) / row_tiles;
  if (s < 1) s = 1;
  if (s > CT) s = CT;
  if (s > 64) s = 64;
  return (int)s;
}

int launch_gram_topk(const scedar_cells* c, int64_t row_begin, int64_t row_end,
                     int K, int splits, double* part_d, int32_t* part_i,
                     cudaStream_t s) {
  if (row_end <= row_begin) return SCEDAR_OK;
  if (K < 1 || K > KFUSED) return fail(SCEDAR_ERR_UNSUPPORTED, "fused top-k keeps at most 32 entries");
  GramArgs a = base_args(c);
  a.row_begin = row_begin;
  a.row_end = row_end;
  a.K = K;
  a.splits = splits;
  a.part_d = part_d;
  a.part_i = part_i;
  const int64_t RT = (row_end + BM - 1) / BM - row_begin / BM;
  return launch<MODE_TOPK>(c, a, dim3((unsigned)RT, (unsigned)splits), s);
}

}  // namespace scedar
