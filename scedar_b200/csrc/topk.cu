// K3 -- row-wise k-nearest-neighbour selection (HBM-bound) and the full stable
// column sort.
//
// Order everywhere is (distance, index) ascending: the order of
// np.argsort(kind="mergesort") on a column of D (scedar/eda/sdm.py:1298-1305),
// i.e. exact ties go to the lower cell index.
//
//   topk_from_d : one warp streams one row of a materialised D stripe
//                 (8*n bytes per row, coalesced 256-byte warp loads, 8 loads in
//                 flight per lane) and keeps the k+1 best in registers -- a
//                 sorted list striped over the lanes, threshold-gated so almost
//                 every element costs one compare; insertions are warp-shuffle
//                 shifts.  Replaces sdm.py:764 and NearestNeighbors
//                 (metric="precomputed") sdm.py:1070-1079.
//   topk_merge  : merges the per-column-split partial lists of the fused
//                 Gram->top-k kernel and applies the same exclusion rule.
//   col_sort    : bitonic sort of (distance, index) pairs per row, shared
//                 memory for spans <= 4096 and global passes above
//                 (np.sort / np.argsort(kind="mergesort", axis=0), 1289-1305).
#include "common.cuh"
#include "topk_list.cuh"

namespace scedar {
namespace {

template <int R>
__global__ void __launch_bounds__(256)
topk_from_d_kernel(const double* __restrict__ d, int64_t ldd, int64_t n_cols,
                   int64_t row_begin, int64_t rows, int k, int exclude,
                   int64_t* __restrict__ idx_out, double* __restrict__ dist_out,
                   const int64_t* __restrict__ row_ids, int64_t out_base) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int K = k + 1;
  const double* __restrict__ src = d + row * ldd;
  WarpList<R> L;
  L.init();
  constexpr int U = 8;
  for (int64_t base = 0; base < n_cols; base += 32 * U) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t c = base + u * 32 + lane;
      v[u] = c < n_cols ? __ldcs(src + c) : pos_inf();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t c = base + u * 32 + lane;
      L.offer(v[u], (int)c, c < n_cols, K, lane);
    }
  }
  // gathered rows: row r of d is cell row_ids[r]; its lists go to that cell's slot
  const int64_t self = row_ids ? row_ids[row] : row_begin + row;
  const int64_t orow = row_ids ? self - out_base : row;
  L.emit(K, exclude, (int)self, idx_out + orow * k, dist_out + orow * k, lane);
}

// Selection from D with the block minima the INT8 split Gram kernel leaves
// behind (min of every aligned run of 32 entries of a row): the k+1 smallest
// entries of a row lie in runs whose minimum is <= the (k+1)-th smallest run
// minimum t (k+1 different entries are <= t, so the (k+1)-th smallest entry is
// too, and every entry that small sits in a run with minimum <= t -- ties
// included).  Pass 1 finds t from the n/32 minima, pass 2 reads only those
// runs of D: 8*n/32 + ~(k+1)*256 bytes per row instead of 8*n.
template <int R>
__global__ void __launch_bounds__(256)
topk_blockmin_kernel(const double* __restrict__ d, int64_t ldd, int64_t n_cols,
                     const double* __restrict__ bmin, int64_t ldb, int64_t row_begin,
                     int64_t rows, int k, int exclude, int64_t* __restrict__ idx_out,
                     double* __restrict__ dist_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int K = k + 1;
  const int64_t nb = (n_cols + 31) / 32;
  const double* __restrict__ mrow = bmin + row * ldb;
  double t;
  {
    WarpList<R> M;
    M.init();
    constexpr int U = 4;
    for (int64_t base = 0; base < nb; base += 32 * U) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t b = base + u * 32 + lane;
        v[u] = b < nb ? __ldg(mrow + b) : pos_inf();
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t b = base + u * 32 + lane;
        M.offer(v[u], (int)b, b < nb, K, lane);
      }
    }
    t = M.td;      // +inf while fewer than k+1 runs: everything is read
  }
  const double* __restrict__ src = d + row * ldd;
  WarpList<R> L;
  L.init();
  for (int64_t base = 0; base < nb; base += 32) {
    const int64_t b = base + lane;
    const double m = b < nb ? __ldg(mrow + b) : pos_inf();
    unsigned take = __ballot_sync(0xffffffffu, b < nb && m <= t);
    while (take) {
      // up to four runs in flight
      int64_t c[4];
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        c[u] = -1;
        if (take) {
          c[u] = (base + __ffs(take) - 1) * 32 + lane;
          take &= take - 1;
        }
        v[u] = (c[u] >= 0 && c[u] < n_cols) ? __ldcs(src + c[u]) : pos_inf();
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) L.offer(v[u], (int)c[u], c[u] >= 0 && c[u] < n_cols, K, lane);
    }
  }
  L.emit(K, exclude, (int)(row_begin + row), idx_out + row * k, dist_out + row * k, lane);
}

template <int R>
__global__ void __launch_bounds__(256)
topk_merge_kernel(const double* __restrict__ part_d, const int32_t* __restrict__ part_i,
                  int splits, int64_t rows, int K, int64_t row_begin, int k, int exclude,
                  int64_t* __restrict__ idx_out, double* __restrict__ dist_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  WarpList<R> L;
  L.init();
  for (int s = 0; s < splits; ++s) {
    const int64_t o = ((int64_t)s * rows + row) * K;
    for (int e0 = 0; e0 < K; e0 += 32) {
      const int e = e0 + lane;
      const bool ok = e < K;
      const double v = ok ? part_d[o + e] : pos_inf();
      const int j = ok ? part_i[o + e] : 0x7fffffff;
      L.offer(v, j, ok, K, lane);
    }
  }
  L.emit(K, exclude, (int)(row_begin + row), idx_out + row * k, dist_out + row * k, lane);
}

// ---------------------------------------------------------------------------
// full sort
// ---------------------------------------------------------------------------
constexpr int SORT_BLK = 4096;  // elements sorted in shared memory per CTA
constexpr int SORT_NT = 512;

__device__ __forceinline__ void cmp_swap(double& ad, int& ai, double& bd, int& bi, bool up) {
  // ascending span (up): the smaller goes first
  if (before(bd, bi, ad, ai) == up) {
    const double t = ad;
    ad = bd;
    bd = t;
    const int u = ai;
    ai = bi;
    bi = u;
  }
}

// k_lo == 2: full local sort of each SORT_BLK span (directions follow the
// global index so spans alternate as the bitonic network needs);
// k_lo == k_hi > SORT_BLK: the j < SORT_BLK tail of merge stage k.
__global__ void __launch_bounds__(SORT_NT)
bitonic_local_kernel(double* __restrict__ kd, int* __restrict__ ki, int64_t N2,
                     int64_t k_lo, int64_t k_hi) {
  __shared__ double sd[SORT_BLK];
  __shared__ int si[SORT_BLK];
  const int64_t spans = N2 / SORT_BLK;
  const int64_t row = blockIdx.x / spans, span = blockIdx.x % spans;
  const int64_t g0 = span * SORT_BLK;
  double* __restrict__ gd = kd + row * N2 + g0;
  int* __restrict__ gi = ki + row * N2 + g0;
  for (int e = threadIdx.x; e < SORT_BLK; e += SORT_NT) {
    sd[e] = gd[e];
    si[e] = gi[e];
  }
  __syncthreads();
  for (int64_t k = k_lo; k <= k_hi; k <<= 1) {
    int64_t jstart = k >> 1;
    if (jstart >= SORT_BLK) jstart = SORT_BLK >> 1;
    for (int64_t j = jstart; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < SORT_BLK / 2; t += SORT_NT) {
        const int lo = (int)(((t & ~(j - 1)) << 1) | (t & (j - 1)));
        const int hi = lo | (int)j;
        const bool up = ((g0 + lo) & k) == 0;
        cmp_swap(sd[lo], si[lo], sd[hi], si[hi], up);
      }
      __syncthreads();
    }
  }
  for (int e = threadIdx.x; e < SORT_BLK; e += SORT_NT) {
    gd[e] = sd[e];
    gi[e] = si[e];
  }
}

__global__ void __launch_bounds__(256)
bitonic_global_kernel(double* __restrict__ kd, int* __restrict__ ki, int64_t N2,
                      int64_t rows, int64_t k, int64_t j) {
  const int64_t half = N2 >> 1;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * half) return;
  const int64_t row = gid / half, t = gid % half;
  const int64_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
  const int64_t hi = lo | j;
  const bool up = (lo & k) == 0;
  double* __restrict__ d = kd + row * N2;
  int* __restrict__ i = ki + row * N2;
  double ad = d[lo], bd = d[hi];
  int ai = i[lo], bi = i[hi];
  if (before(bd, bi, ad, ai) == up) {
    d[lo] = bd;
    d[hi] = ad;
    i[lo] = bi;
    i[hi] = ai;
  }
}

__global__ void __launch_bounds__(256)
sort_fill_kernel(const double* __restrict__ d, int64_t ldd, int64_t n, int64_t row0,
                 int64_t rows, int64_t N2, double* __restrict__ kd, int* __restrict__ ki) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * N2) return;
  const int64_t r = gid / N2, c = gid % N2;
  // column j of a symmetric D == row j: read rows (coalesced)
  kd[gid] = c < n ? d[(row0 + r) * ldd + c] : pos_inf();
  ki[gid] = c < n ? (int)c : 0x7fffffff;
}

__global__ void __launch_bounds__(256)
sort_emit_kernel(const double* __restrict__ kd, const int* __restrict__ ki, int64_t n,
                 int64_t row0, int64_t rows, int64_t N2, double* __restrict__ sorted_out,
                 int64_t* __restrict__ arg_out) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * n) return;
  const int64_t r = gid / n, c = gid % n;
  if (sorted_out) sorted_out[(row0 + r) * n + c] = kd[r * N2 + c];
  if (arg_out) arg_out[(row0 + r) * n + c] = ki[r * N2 + c];
}

}  // namespace

int launch_topk_from_d(const double* d, int64_t ldd, int64_t n_cols, int64_t row_begin,
                       int64_t row_end, int k, int exclude, int64_t* idx_out,
                       double* dist_out, cudaStream_t s, const int64_t* row_ids,
                       int64_t out_base) {
  const int64_t rows = row_end - row_begin;
  if (rows <= 0 || k == 0) return SCEDAR_OK;
  const int K = k + 1;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (K <= 32)
    topk_from_d_kernel<1><<<grid, 256, 0, s>>>(d, ldd, n_cols, row_begin, rows, k, exclude, idx_out, dist_out, row_ids, out_base);
  else if (K <= 64)
    topk_from_d_kernel<2><<<grid, 256, 0, s>>>(d, ldd, n_cols, row_begin, rows, k, exclude, idx_out, dist_out, row_ids, out_base);
  else if (K <= 32 * kMaxR)
    topk_from_d_kernel<4><<<grid, 256, 0, s>>>(d, ldd, n_cols, row_begin, rows, k, exclude, idx_out, dist_out, row_ids, out_base);
  else
    return fail(SCEDAR_ERR_UNSUPPORTED, "k+1 > 128: use scedar_col_sort");
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int launch_topk_blockmin(const double* d, int64_t ldd, int64_t n_cols, const double* bmin,
                         int64_t ldb, int64_t row_begin, int64_t row_end, int k, int exclude,
                         int64_t* idx_out, double* dist_out, cudaStream_t s) {
  const int64_t rows = row_end - row_begin;
  if (rows <= 0 || k == 0) return SCEDAR_OK;
  const int K = k + 1;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (K <= 32)
    topk_blockmin_kernel<1><<<grid, 256, 0, s>>>(d, ldd, n_cols, bmin, ldb, row_begin, rows, k, exclude, idx_out, dist_out);
  else if (K <= 64)
    topk_blockmin_kernel<2><<<grid, 256, 0, s>>>(d, ldd, n_cols, bmin, ldb, row_begin, rows, k, exclude, idx_out, dist_out);
  else if (K <= 32 * kMaxR)
    topk_blockmin_kernel<4><<<grid, 256, 0, s>>>(d, ldd, n_cols, bmin, ldb, row_begin, rows, k, exclude, idx_out, dist_out);
  else
    return fail(SCEDAR_ERR_UNSUPPORTED, "k+1 > 128: use scedar_col_sort");
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int launch_topk_merge(const double* part_d, const int32_t* part_i, int splits, int64_t rows,
                      int K, int64_t row_begin, int k, int exclude, int64_t* idx_out,
                      double* dist_out, cudaStream_t s) {
  if (rows <= 0 || k == 0) return SCEDAR_OK;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (K <= 32)
    topk_merge_kernel<1><<<grid, 256, 0, s>>>(part_d, part_i, splits, rows, K, row_begin, k, exclude, idx_out, dist_out);
  else
    return fail(SCEDAR_ERR_UNSUPPORTED, "merge of more than 32 entries per row");
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int launch_col_sort(const double* d, int64_t ldd, int64_t total_rows, int64_t n,
                    double* sorted_out, int64_t* argsorted_out, cudaStream_t s) {
  if (n == 0 || total_rows == 0) return SCEDAR_OK;
  int64_t N2 = SORT_BLK;
  while (N2 < n) N2 <<= 1;
  // rows per pass: keep the (8+4)-byte pair workspace under ~2 GiB
  int64_t chunk = (int64_t(2) << 30) / (N2 * 12);
  if (chunk < 1) chunk = 1;
  if (chunk > total_rows) chunk = total_rows;
  Scratch wd, wi;
  SCEDAR_CHECK(wd.alloc(sizeof(double) * chunk * N2, s));
  SCEDAR_CHECK(wi.alloc(sizeof(int) * chunk * N2, s));
  double* kd = wd.as<double>();
  int* ki = wi.as<int>();
  for (int64_t row0 = 0; row0 < total_rows; row0 += chunk) {
    const int64_t rows = (total_rows - row0) < chunk ? (total_rows - row0) : chunk;
    sort_fill_kernel<<<(unsigned)((rows * N2 + 255) / 256), 256, 0, s>>>(d, ldd, n, row0, rows, N2, kd, ki);
    const unsigned local_grid = (unsigned)(rows * (N2 / SORT_BLK));
    bitonic_local_kernel<<<local_grid, SORT_NT, 0, s>>>(kd, ki, N2, 2, SORT_BLK);
    for (int64_t k = 2 * SORT_BLK; k <= N2; k <<= 1) {
      for (int64_t j = k >> 1; j >= SORT_BLK; j >>= 1)
        bitonic_global_kernel<<<(unsigned)((rows * (N2 / 2) + 255) / 256), 256, 0, s>>>(kd, ki, N2, rows, k, j);
      bitonic_local_kernel<<<local_grid, SORT_NT, 0, s>>>(kd, ki, N2, k, k);
    }
    sort_emit_kernel<<<(unsigned)((rows * n + 255) / 256), 256, 0, s>>>(kd, ki, n, row0, rows, N2, sorted_out, argsorted_out);
    SCEDAR_CUDA(cudaGetLastError());
  }
  return SCEDAR_OK;
}

}  // namespace scedar
