// Consumers of the kNN lists and of D that sit directly behind the hot path
// (SURVEY.md section 8f, ranks 1-3).  All HBM-bound, all on the device so the
// lists / the distance matrix never make a host round trip between stages.
//
//   conn_csr        kNN lists -> CSR connectivity matrix, zero distances stored
//                   as -inf, columns sorted inside a row
//                   (s_knn_connectivity_matrix, scedar/eda/sdm.py:934-947: the
//                   coo_matrix(...).tocsr() of the concatenated lists)
//   affinity        CSR data -> edge weights (max - w) * aff_scale with
//                   -inf -> 0 first (knn_conn_mat_to_aff_graph, sdm.py:1085-1093)
//   masked_kth      value at rank r of every live row of D restricted to the
//                   live columns: np.sort(curr_dist_mat, axis=0)[i_k, :] of the
//                   shrinking sub-matrix of RareSampleDetection
//                   (scedar/knn/detection.py:100-121) without re-slicing or
//                   re-sorting D; 8 n bytes per live row per iteration
//   rare_update     the keep / drop decision of one detection iteration
//                   (detection.py:63-66, 111-115)
//   gather_rows     x[kept, :] for the no-pdist variant (detection.py:71-74)
//   impute_iter     one iteration of FeatureImputation._impute_features_runner
//                   (scedar/knn/imputation.py:56-101): present mask, neighbour
//                   present count, per-gene median over the k neighbours
#include <algorithm>

#include "common.cuh"
#include "topk_list.cuh"

namespace scedar {
namespace {

constexpr int CONN_MAXK = 128;

__device__ __forceinline__ double neg_inf() {
  return __longlong_as_double(0xfff0000000000000LL);
}
__device__ __forceinline__ double quiet_nan() {
  return __longlong_as_double(0x7ff8000000000000LL);
}

// one warp per cell: order its k (neighbour, distance) pairs by neighbour index
__global__ void __launch_bounds__(256)
conn_csr_kernel(const int64_t* __restrict__ idx, const double* __restrict__ dist, int64_t n,
                int k, int64_t* __restrict__ indptr, int32_t* __restrict__ indices,
                double* __restrict__ data) {
  __shared__ int s_idx[8][CONN_MAXK];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * 8 + w;
  if (row > n) return;
  if (lane == 0) indptr[row] = row * k;  // row == n writes the end pointer
  if (row == n) return;
  for (int e = lane; e < k; e += 32) s_idx[w][e] = (int)idx[row * k + e];
  __syncwarp();
  for (int e = lane; e < k; e += 32) {
    const int me = s_idx[w][e];
    int rank = 0;
    for (int m = 0; m < k; ++m) {
      const int o = s_idx[w][m];
      rank += (o < me) || (o == me && m < e);
    }
    const double v = dist[row * k + e];
    indices[row * k + rank] = me;
    data[row * k + rank] = v == 0.0 ? neg_inf() : v;
  }
}

// any k: one block per cell, the cell's neighbour indices in shared memory when
// they fit (k <= 12,288), else read from global memory
__global__ void __launch_bounds__(256)
conn_csr_big_kernel(const int64_t* __restrict__ idx, const double* __restrict__ dist, int64_t n,
                    int k, int in_smem, int64_t* __restrict__ indptr,
                    int32_t* __restrict__ indices, double* __restrict__ data) {
  extern __shared__ int s_big[];
  const int64_t row = blockIdx.x;
  if (threadIdx.x == 0) indptr[row] = row * k;  // row == n writes the end pointer
  if (row == n) return;
  const int64_t* __restrict__ mine = idx + row * k;
  if (in_smem) {
    for (int e = threadIdx.x; e < k; e += blockDim.x) s_big[e] = (int)mine[e];
    __syncthreads();
  }
  for (int e = threadIdx.x; e < k; e += blockDim.x) {
    const int me = in_smem ? s_big[e] : (int)mine[e];
    int rank = 0;
    for (int m = 0; m < k; ++m) {
      const int o = in_smem ? s_big[m] : (int)mine[m];
      rank += (o < me) || (o == me && m < e);
    }
    const double v = dist[row * k + e];
    indices[row * k + rank] = me;
    data[row * k + rank] = v == 0.0 ? neg_inf() : v;
  }
}

// max over max(w, 0) with -inf -> 0: all candidates are >= 0, so the IEEE bit
// pattern orders like the value
__global__ void __launch_bounds__(256)
affinity_max_kernel(const double* __restrict__ w, int64_t nnz,
                    unsigned long long* __restrict__ max_bits) {
  double m = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz;
       e += (int64_t)gridDim.x * blockDim.x) {
    double v = w[e];
    if (v == neg_inf()) v = 0.0;
    m = v > m ? v : m;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(max_bits, (unsigned long long)__double_as_longlong(m));
}

__global__ void __launch_bounds__(256)
affinity_kernel(const double* __restrict__ w, int64_t nnz, double aff_scale,
                const unsigned long long* __restrict__ max_bits, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const double mx = __longlong_as_double((long long)*max_bits);
  double v = w[e];
  if (v == neg_inf()) v = 0.0;
  out[e] = __dmul_rn(__dsub_rn(mx, v), aff_scale);
}

// one warp per row of the stripe; columns and rows filtered by the live mask
template <int R>
__global__ void __launch_bounds__(256)
masked_kth_kernel(const double* __restrict__ d, int64_t ldd, int64_t n_cols, int64_t row_begin,
                  int64_t rows, const uint8_t* __restrict__ alive, int rank,
                  double* __restrict__ kth_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (!alive[row_begin + row]) {
    if (lane == 0) kth_out[row] = pos_inf();
    return;
  }
  const int K = rank + 1;
  const double* __restrict__ src = d + row * ldd;
  WarpList<R> L;
  L.init();
  constexpr int U = 8;
  for (int64_t base = 0; base < n_cols; base += 32 * U) {
    double v[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t c = base + u * 32 + lane;
      ok[u] = c < n_cols && alive[c];
      v[u] = c < n_cols ? __ldcs(src + c) : pos_inf();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t c = base + u * 32 + lane;
      L.offer(v[u], (int)c, ok[u], K, lane);
    }
  }
  L.refresh(K);
  if (lane == 0) kth_out[row] = L.td;
}

// Any rank: one block per row, most-significant-byte-first radix selection on
// the order-preserving integer image of the distances (8 passes over the row,
// the first from HBM, the rest from L2).  Same result as the list kernel: the
// value at 0-based `rank` among the live columns, +inf when fewer are alive.
__device__ __forceinline__ unsigned long long ord_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__global__ void __launch_bounds__(256)
masked_kth_radix_kernel(const double* __restrict__ d, int64_t ldd, int64_t n_cols,
                        int64_t row_begin, const uint8_t* __restrict__ alive, int rank,
                        double* __restrict__ kth_out) {
  __shared__ unsigned hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ unsigned s_rank;
  __shared__ int s_short;
  const int64_t row = blockIdx.x;
  if (!alive[row_begin + row]) {
    if (threadIdx.x == 0) kth_out[row] = pos_inf();
    return;
  }
  const double* __restrict__ src = d + row * ldd;
  if (threadIdx.x == 0) {
    s_prefix = 0;
    s_rank = (unsigned)rank;
    s_short = 0;
  }
  for (int pass = 7; pass >= 0; --pass) {
    hist[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    const unsigned long long mask = pass == 7 ? 0ULL : (~0ULL << (8 * (pass + 1)));
    for (int64_t c = threadIdx.x; c < n_cols; c += blockDim.x) {
      if (!alive[c]) continue;
      const unsigned long long key = ord_key(src[c]);
      if ((key & mask) == prefix) atomicAdd(&hist[(unsigned)(key >> (8 * pass)) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned r = s_rank, cum = 0;
      int b = 0;
      for (; b < 256; ++b) {
        if (cum + hist[b] > r) break;
        cum += hist[b];
      }
      if (b == 256) {
        s_short = 1;   // fewer live columns than rank + 1
        b = 255;
      }
      s_rank = r - cum;
      s_prefix = prefix | ((unsigned long long)b << (8 * pass));
    }
    __syncthreads();
    if (s_short) break;
  }
  if (threadIdx.x == 0) {
    const unsigned long long key = s_prefix;
    const unsigned long long bits = (key >> 63) ? (key & 0x7fffffffffffffffULL) : ~key;
    kth_out[row] = s_short ? pos_inf() : __longlong_as_double((long long)bits);
  }
}

// counters[0] = cells kept, counters[1] = bits of max(kth) over the kept cells
__global__ void __launch_bounds__(256)
rare_update_kernel(const double* __restrict__ kth, int64_t ldk, int64_t n, double cutoff,
                   int iter, uint8_t* __restrict__ alive, int32_t* __restrict__ last_kept,
                   unsigned long long* __restrict__ counters) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  if (i < n && alive[i]) {
    keep = kth[i * ldk] <= cutoff;
    if (keep)
      last_kept[i] = iter;
    else
      alive[i] = 0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(counters, (unsigned long long)__popc(m));
}

__global__ void __launch_bounds__(256)
masked_max_kernel(const double* __restrict__ v, int64_t ldv, int64_t n,
                  const uint8_t* __restrict__ alive, unsigned long long* __restrict__ max_bits) {
  double m = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    if (alive && !alive[e]) continue;
    const double t = v[e * ldv];
    m = t > m ? t : m;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(max_bits, (unsigned long long)__double_as_longlong(m));
}

// ---- ordered compaction of the live mask: indices of the live cells --------
constexpr int CMP_NT = 256, CMP_PER = 4;           // 1024 cells per block
__global__ void __launch_bounds__(CMP_NT)
compact_count_kernel(const uint8_t* __restrict__ alive, int64_t n,
                     int64_t* __restrict__ block_count) {
  const int64_t base = (int64_t)blockIdx.x * CMP_NT * CMP_PER;
  int c = 0;
  for (int u = 0; u < CMP_PER; ++u) {
    const int64_t i = base + u * CMP_NT + threadIdx.x;
    c += i < n && alive[i];
  }
  __shared__ int ws[CMP_NT / 32];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < CMP_NT / 32; ++w) t += ws[w];
    block_count[blockIdx.x] = t;
  }
}
// one block: exclusive scan of the block counts (in place)
__global__ void __launch_bounds__(1024)
compact_scan_kernel(int64_t* __restrict__ block_count, int64_t blocks) {
  __shared__ int64_t part[1024];
  const int64_t per = (blocks + 1023) / 1024;
  const int64_t b0 = threadIdx.x * per, b1 = b0 + per < blocks ? b0 + per : blocks;
  int64_t sum = 0;
  for (int64_t b = b0; b < b1; ++b) sum += block_count[b];
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t run = 0;
    for (int t = 0; t < 1024; ++t) {
      const int64_t v = part[t];
      part[t] = run;
      run += v;
    }
  }
  __syncthreads();
  int64_t run = part[threadIdx.x];
  for (int64_t b = b0; b < b1; ++b) {
    const int64_t v = block_count[b];
    block_count[b] = run;
    run += v;
  }
}
__global__ void __launch_bounds__(CMP_NT)
compact_write_kernel(const uint8_t* __restrict__ alive, int64_t n,
                     const int64_t* __restrict__ block_offset, int64_t* __restrict__ sel) {
  __shared__ int ws[CMP_NT / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * CMP_NT * CMP_PER;
  int64_t out = block_offset[blockIdx.x];
  for (int u = 0; u < CMP_PER; ++u) {
    const int64_t i = base + u * CMP_NT + threadIdx.x;
    const bool a = i < n && alive[i];
    const unsigned m = __ballot_sync(0xffffffffu, a);
    if (lane == 0) ws[w] = __popc(m);
    __syncthreads();
    int before_w = 0, total = 0;
    for (int q = 0; q < CMP_NT / 32; ++q) {
      before_w += q < w ? ws[q] : 0;
      total += ws[q];
    }
    if (a) sel[out + before_w + __popc(m & ((1u << lane) - 1))] = i;
    out += total;
    __syncthreads();
  }
}

// out[r, :] = x[sel[r], :], 128-bit accesses when the strides allow
__global__ void __launch_bounds__(256)
gather_rows_kernel(const double* __restrict__ x, int64_t ldx, int64_t p,
                   const int64_t* __restrict__ sel, double* __restrict__ out, int64_t ldo,
                   int vec) {
  const int64_t r = blockIdx.x;
  const double* __restrict__ src = x + sel[r] * ldx;
  double* __restrict__ dst = out + r * ldo;
  if (vec) {
    const double2* __restrict__ s2 = reinterpret_cast<const double2*>(src);
    double2* __restrict__ d2 = reinterpret_cast<double2*>(dst);
    for (int64_t c = threadIdx.x; c < p / 2; c += blockDim.x) d2[c] = s2[c];
    if ((p & 1) && threadIdx.x == 0) dst[p - 1] = src[p - 1];
  } else {
    for (int64_t c = threadIdx.x; c < p; c += blockDim.x) dst[c] = src[c];
  }
}

// One imputation iteration (Jacobi: reads x_curr only, writes all of x_next).
// Block = one cell x IMP_TF genes; the k neighbour values of a gene sit in a
// shared-memory column private to the thread, so no barrier is needed.
constexpr int IMP_TF = 128;
__global__ void __launch_bounds__(IMP_TF)
impute_iter_kernel(const double* __restrict__ x_curr, int64_t ldx, int64_t n, int64_t p,
                   const int64_t* __restrict__ knn, int k, int n_do_i, double min_present,
                   int iter, double* __restrict__ x_next, int64_t ldn,
                   int32_t* __restrict__ idc, int64_t ldi, int64_t s0) {
  extern __shared__ double vals[];  // [k][IMP_TF]
  const int t = threadIdx.x;
  const int64_t s = s0 + blockIdx.y;
  const int64_t f = (int64_t)blockIdx.x * IMP_TF + t;
  if (f >= p) return;
  const double own = x_curr[s * ldx + f];
  int present = 0;
  bool has_nan = false;
  for (int j = 0; j < k; ++j) {
    const double v = x_curr[knn[s * k + j] * ldx + f];
    vals[j * IMP_TF + t] = v;
    present += v >= min_present;
    has_nan |= v != v;
  }
  double res = own;
  // absent in this cell, present in >= n_do_i of its neighbours
  if (!(own >= min_present) && present >= n_do_i) {
    // np.median over all k neighbour values: element(s) at sorted position
    // (k-1)/2 and k/2, found by counting
    const int m1 = (k - 1) >> 1, m2 = k >> 1;
    double a1 = 0.0, a2 = 0.0;
    for (int a = 0; a < k; ++a) {
      const double va = vals[a * IMP_TF + t];
      int lt = 0, eq = 0;
      for (int b = 0; b < k; ++b) {
        const double vb = vals[b * IMP_TF + t];
        lt += vb < va;
        eq += vb == va;
      }
      if (lt <= m1 && m1 < lt + eq) a1 = va;
      if (lt <= m2 && m2 < lt + eq) a2 = va;
    }
    res = m1 == m2 ? a1 : __ddiv_rn(__dadd_rn(a1, a2), 2.0);
    if (has_nan) res = quiet_nan();
    idc[s * ldi + f] = iter;
  }
  x_next[s * ldn + f] = res;
}

// counters[0] += entries with idc > 0, counters[1] += cells with any
__global__ void __launch_bounds__(256)
impute_stats_kernel(const int32_t* __restrict__ idc, int64_t ldi, int64_t n, int64_t p,
                    unsigned long long* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  int c = 0;
  for (int64_t f = lane; f < p; f += 32) c += idc[row * ldi + f] > 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0 && c) {
    atomicAdd(counters, (unsigned long long)c);
    atomicAdd(counters + 1, 1ULL);
  }
}

}  // namespace

int launch_conn_csr(const int64_t* idx, const double* dist, int64_t n, int k, int64_t* indptr,
                    int32_t* indices, double* data, cudaStream_t s) {
  if (k > CONN_MAXK) {
    const int in_smem = k <= 12288;
    conn_csr_big_kernel<<<(unsigned)(n + 1), 256, in_smem ? sizeof(int) * k : 0, s>>>(
        idx, dist, n, k, in_smem, indptr, indices, data);
  } else {
    conn_csr_kernel<<<(unsigned)((n + 1 + 7) / 8), 256, 0, s>>>(idx, dist, n, k, indptr, indices,
                                                               data);
  }
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int launch_affinity(const double* w, int64_t nnz, double aff_scale, double* out,
                    cudaStream_t s) {
  if (nnz <= 0) return SCEDAR_OK;
  Scratch mx;
  SCEDAR_CHECK(mx.alloc(8, s));
  SCEDAR_CUDA(cudaMemsetAsync(mx.ptr, 0, 8, s));
  const int64_t blocks = std::min<int64_t>((nnz + 255) / 256, 8 * (int64_t)num_sms());
  affinity_max_kernel<<<(unsigned)blocks, 256, 0, s>>>(w, nnz, mx.as<unsigned long long>());
  affinity_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, s>>>(
      w, nnz, aff_scale, mx.as<unsigned long long>(), out);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch(2);
  return SCEDAR_OK;
}

int launch_masked_kth(const double* d, int64_t ldd, int64_t n_cols, int64_t row_begin,
                      int64_t row_end, const uint8_t* alive, int rank, double* kth_out,
                      cudaStream_t s) {
  const int64_t rows = row_end - row_begin;
  if (rows <= 0) return SCEDAR_OK;
  const int K = rank + 1;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (K <= 32)
    masked_kth_kernel<1><<<grid, 256, 0, s>>>(d, ldd, n_cols, row_begin, rows, alive, rank, kth_out);
  else if (K <= 64)
    masked_kth_kernel<2><<<grid, 256, 0, s>>>(d, ldd, n_cols, row_begin, rows, alive, rank, kth_out);
  else if (K <= 32 * kMaxR)
    masked_kth_kernel<4><<<grid, 256, 0, s>>>(d, ldd, n_cols, row_begin, rows, alive, rank, kth_out);
  else
    masked_kth_radix_kernel<<<(unsigned)rows, 256, 0, s>>>(d, ldd, n_cols, row_begin, alive, rank,
                                                          kth_out);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int launch_rare_update(const double* kth, int64_t ldk, int64_t n, double cutoff, int iter,
                       uint8_t* alive, int32_t* last_kept, unsigned long long* counters,
                       cudaStream_t s) {
  if (n <= 0) return SCEDAR_OK;
  rare_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(kth, ldk, n, cutoff, iter, alive,
                                                               last_kept, counters);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int launch_masked_max(const double* v, int64_t ldv, int64_t n, const uint8_t* alive,
                      unsigned long long* max_bits, cudaStream_t s) {
  if (n <= 0) return SCEDAR_OK;
  const int64_t blocks = std::min<int64_t>((n + 255) / 256, 8 * (int64_t)num_sms());
  masked_max_kernel<<<(unsigned)blocks, 256, 0, s>>>(v, ldv, n, alive, max_bits);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int launch_compact_alive(const uint8_t* alive, int64_t n, int64_t* sel, cudaStream_t s) {
  if (n <= 0) return SCEDAR_OK;
  const int64_t blocks = (n + CMP_NT * CMP_PER - 1) / (CMP_NT * CMP_PER);
  Scratch bc;
  SCEDAR_CHECK(bc.alloc(sizeof(int64_t) * blocks, s));
  compact_count_kernel<<<(unsigned)blocks, CMP_NT, 0, s>>>(alive, n, bc.as<int64_t>());
  compact_scan_kernel<<<1, 1024, 0, s>>>(bc.as<int64_t>(), blocks);
  compact_write_kernel<<<(unsigned)blocks, CMP_NT, 0, s>>>(alive, n, bc.as<int64_t>(), sel);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch(3);
  return SCEDAR_OK;
}

namespace {
// ids[0..m) += id0 (local row -> cell index), ids[m..cap) = -1 (padding)
__global__ void __launch_bounds__(256)
ids_fix_kernel(int64_t* __restrict__ ids, int64_t m, int64_t cap, int64_t id0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  ids[i] = i < m ? ids[i] + id0 : -1;
}
}  // namespace

// flagged rows of a stripe -> ascending cell indices (m of them, known to the
// host), padded with -1 up to ids_cap
int launch_flag_to_ids(const uint8_t* flag, int64_t n, int64_t id0, int64_t* ids, int64_t m,
                       int64_t ids_cap, cudaStream_t s) {
  if (ids_cap <= 0) return SCEDAR_OK;
  SCEDAR_CHECK(launch_compact_alive(flag, n, ids, s));
  ids_fix_kernel<<<(unsigned)((ids_cap + 255) / 256), 256, 0, s>>>(ids, m, ids_cap, id0);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int launch_gather_rows(const double* x, int64_t ldx, int64_t p, const int64_t* sel, int64_t m,
                       double* out, int64_t ldo, cudaStream_t s) {
  if (m <= 0 || p <= 0) return SCEDAR_OK;
  const int vec = (ldx % 2 == 0) && (ldo % 2 == 0) &&
                  (reinterpret_cast<uintptr_t>(x) % 16 == 0) &&
                  (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  gather_rows_kernel<<<(unsigned)m, 256, 0, s>>>(x, ldx, p, sel, out, ldo, vec);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int launch_impute_iter(const double* x_curr, int64_t ldx, int64_t n, int64_t p,
                       const int64_t* knn, int k, int n_do_i, double min_present, int iter,
                       int64_t row_begin, int64_t row_end, double* x_next, int64_t ldn,
                       int32_t* idc, int64_t ldi, unsigned long long* counters,
                       cudaStream_t s) {
  if (row_end <= row_begin || p <= 0) return SCEDAR_OK;
  const size_t smem = sizeof(double) * (size_t)k * IMP_TF;
  if (smem > 200 * 1024) return fail(SCEDAR_ERR_UNSUPPORTED, "imputation: k must be <= 200");
  if (n > 65535LL * 65535LL) return fail(SCEDAR_ERR_UNSUPPORTED, "imputation: too many cells");
  SCEDAR_CUDA(cudaFuncSetAttribute(impute_iter_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // grid.y is limited to 65535: walk the cells in slabs
  for (int64_t s0 = row_begin; s0 < row_end; s0 += 65535) {
    const int64_t ns = std::min<int64_t>(65535, row_end - s0);
    dim3 grid((unsigned)((p + IMP_TF - 1) / IMP_TF), (unsigned)ns);
    impute_iter_kernel<<<grid, IMP_TF, smem, s>>>(x_curr, ldx, n, p, knn, k, n_do_i, min_present,
                                                 iter, x_next, ldn, idc, ldi, s0);
    count_launch();
  }
  SCEDAR_CUDA(cudaGetLastError());
  if (counters) {
    SCEDAR_CUDA(cudaMemsetAsync(counters, 0, 16, s));
    impute_stats_kernel<<<(unsigned)((row_end - row_begin + 7) / 8), 256, 0, s>>>(
        idc + row_begin * ldi, ldi, row_end - row_begin, p, counters);
    SCEDAR_CUDA(cudaGetLastError());
    count_launch();
  }
  return SCEDAR_OK;
}

}  // namespace scedar
