// C ABI of libscedar_b200 (declared in include/scedar_b200.h): argument
// checking, scratch management and the sequencing of the kernels.  No compute
// happens on the host and there is no CPU fallback.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace scedar {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

// Keep freed scratch in the device's default memory pool instead of returning
// it to the driver at every synchronisation (the default threshold is 0).
static void keep_pool_warm() {
  static thread_local int done_for = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ULL;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done_for = dev;
}

int Scratch::alloc(size_t bytes, cudaStream_t s) {
  keep_pool_warm();
  stream = s;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMallocAsync(&ptr, bytes, s);
  if (e != cudaSuccess) {
    ptr = nullptr;
    cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? SCEDAR_ERR_NOMEM : SCEDAR_ERR_CUDA,
                std::string("cudaMallocAsync(") + std::to_string(bytes) +
                    "): " + cudaGetErrorString(e));
  }
  return SCEDAR_OK;
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches += n; }

int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

namespace {

int check_metric(int metric) {
  if (metric != SCEDAR_METRIC_COSINE && metric != SCEDAR_METRIC_CORRELATION &&
      metric != SCEDAR_METRIC_EUCLIDEAN)
    return fail(SCEDAR_ERR_ARG, "metric must be cosine (0), correlation (1) or euclidean (2)");
  return SCEDAR_OK;
}

int check_precision(int precision) {
  if (precision >= SCEDAR_PRECISION_FP64 && precision <= SCEDAR_PRECISION_FP64_I8S6)
    return SCEDAR_OK;
  return fail(SCEDAR_ERR_ARG, "precision must be 0 (fp64), 1 (fast), 2 (fp64_i8) or 3 (fp64_i8s6)");
}

// slices of the INT8 split path for `precision`, 0 when the DMMA kernel runs
// (not requested, or rows longer than the split's INT32 accumulators allow)
int oz_slices(const scedar_cells* c, int precision) {
  if (precision != SCEDAR_PRECISION_FP64_I8 && precision != SCEDAR_PRECISION_FP64_I8S6) return 0;
  if (!oz_supported(c)) return 0;
  return precision == SCEDAR_PRECISION_FP64_I8 ? 7 : 6;
}

int check_rows(const scedar_cells* c, int64_t row_begin, int64_t row_end) {
  if (!c) return fail(SCEDAR_ERR_ARG, "cells handle is NULL");
  if (row_begin < 0 || row_end < row_begin || row_end > c->n)
    return fail(SCEDAR_ERR_ARG, "row range must satisfy 0 <= row_begin <= row_end <= n");
  return SCEDAR_OK;
}

int new_cells(int64_t n, int64_t p, int metric, cudaStream_t s, scedar_cells** out) {
  if (!out) return fail(SCEDAR_ERR_ARG, "out is NULL");
  *out = nullptr;
  SCEDAR_CHECK(check_metric(metric));
  if (n < 0 || p < 0) return fail(SCEDAR_ERR_ARG, "n and p must be >= 0");
  if (n >= 0x7fffffffLL - 256) return fail(SCEDAR_ERR_UNSUPPORTED, "n must fit in int32");
  scedar_cells* c = new (std::nothrow) scedar_cells();
  if (!c) return fail(SCEDAR_ERR_NOMEM, "host allocation failed");
  c->n = n;
  c->p = p;
  c->metric = metric;
  c->n_pad = std::max<int64_t>(round_up(n, kRowPad), kRowPad);
  c->p_pad = std::max<int64_t>(round_up(p, kColPad), kColPad);
  cudaGetDevice(&c->device);
  // stream-ordered pool allocation: no device-wide sync, and re-preparing
  // cells of the same size reuses the pooled block
  c->stream = s;
  keep_pool_warm();
  cudaError_t e = cudaMallocAsync(&c->xw, sizeof(double) * c->n_pad * c->p_pad, s);
  if (e == cudaSuccess) e = cudaMallocAsync(&c->stat, sizeof(double) * c->n_pad, s);
  if (e != cudaSuccess) {
    cudaGetLastError();
    scedar_cells_free(c);
    return fail(e == cudaErrorMemoryAllocation ? SCEDAR_ERR_NOMEM : SCEDAR_ERR_CUDA,
                std::string("allocating the working copy: ") + cudaGetErrorString(e));
  }
  int rc = encode_tensor_map(&c->xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, c->xw, c->n_pad,
                             c->p_pad, 128, 16);
  if (rc != SCEDAR_OK) {
    scedar_cells_free(c);
    return rc;
  }
  *out = c;
  return SCEDAR_OK;
}

}  // namespace
// FP64 kNN of a row range.  Three ways to get there, chosen by cost:
//   symmetric  whole matrix and D fits in free HBM: lower tiles contracted once
//              (half the flops), mirrored into a scratch D, row-wise top-k from
//              it -- 2.5x faster than the fused kernel at 60,000 x 2,000
//   chunked    D in row chunks of a few GB through scratch + top-k per chunk
//              (all n^2 p products, but at the STORE kernel's rate; only worth
//              it for long rows, p >= 512)
//   fused      Gram tile -> epilogue -> per-row lists in shared memory, D never
//              reaches HBM (k + 1 <= 32); the only form that needs no scratch
// SCEDAR_B200_KNN_FP64 = symmetric | chunked | fused forces one (testing).
int knn_stripe_fp64(const scedar_cells* c, int k, int exclude, int64_t row_begin,
                    int64_t row_end, int64_t* idx_out, double* dist_out, cudaStream_t s,
                    int oz_S) {
  const int64_t rows = row_end - row_begin;
  if (k == 0 || rows <= 0) return SCEDAR_OK;
  const int K = k + 1;
  if (K > 128) return fail(SCEDAR_ERR_UNSUPPORTED, "k+1 > 128: use scedar_col_sort");
  const int64_t n = c->n;
  const char* forced = getenv("SCEDAR_B200_KNN_FP64");
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
    cudaGetLastError();
    free_b = 0;
  }
  // memory the stream-ordered pool already holds is reusable as well
  cudaMemPool_t pool;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long reserved = 0, used = 0;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) ==
            cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess &&
        reserved > used)
      free_b += (size_t)(reserved - used);
  }
  const double d_bytes = 8.0 * (double)n * (double)n;
  enum { SYM, CHUNKED, FUSED } mode;
  if (rows == n && d_bytes <= 0.45 * (double)free_b)
    mode = SYM;
  else if (K > 32 || ((double)free_b > 6e9 && c->p_pad >= 512))
    mode = CHUNKED;   // measured: 1.2-1.4x over fused at p = 2000, 2x slower at p = 50
  else
    mode = FUSED;
  // the INT8 split kernel has no fused top-k: D goes through scratch
  if (oz_S && mode == FUSED) mode = CHUNKED;
  if (forced && K <= 32 && !oz_S) {
    if (!strcmp(forced, "fused")) mode = FUSED;
    if (!strcmp(forced, "chunked")) mode = CHUNKED;
    if (!strcmp(forced, "symmetric") && rows == n) mode = SYM;
  }
  if (mode == SYM) {
    Scratch dbuf;
    if (dbuf.alloc((size_t)d_bytes, s) == SCEDAR_OK) {
      SCEDAR_CHECK(oz_S ? launch_oz_symmetric(c, oz_S, dbuf.as<double>(), n, s)
                        : launch_gram_symmetric(c, dbuf.as<double>(), n, s));
      return launch_topk_from_d(dbuf.as<double>(), n, n, 0, n, k, exclude, idx_out, dist_out, s);
    }
    mode = CHUNKED;   // fragmentation: fall through with a smaller request
  }
  if (mode == CHUNKED) {
    // row chunks of up to 4 GB (a quarter of what is free at most)
    double budget = std::min(4e9, 0.25 * (double)free_b);
    if (budget < 64e6) budget = 64e6;
    int64_t chunk = (int64_t)(budget / (8.0 * (double)std::max<int64_t>(n, 1)));
    chunk = std::max<int64_t>(128, chunk / 128 * 128);
    chunk = std::min(chunk, round_up(rows, 128));
    Scratch dbuf;
    if (dbuf.alloc(sizeof(double) * (size_t)chunk * (size_t)n, s) == SCEDAR_OK) {
      // the split kernel takes 128-aligned stripes
      if (oz_S && row_begin % 128 != 0) oz_S = 0;
      for (int64_t r0 = row_begin; r0 < row_end; r0 += chunk) {
        const int64_t r1 = std::min(row_end, r0 + chunk);
        SCEDAR_CHECK(oz_S ? launch_oz_stripe(c, oz_S, r0, r1, dbuf.as<double>(), n, s)
                          : launch_gram_stripe(c, r0, r1, dbuf.as<double>(), n, s));
        SCEDAR_CHECK(launch_topk_from_d(dbuf.as<double>(), n, n, r0, r1, k, exclude,
                                        idx_out + (r0 - row_begin) * k,
                                        dist_out + (r0 - row_begin) * k, s));
      }
      return SCEDAR_OK;
    }
    if (K > 32 || oz_S) return SCEDAR_ERR_NOMEM;
  }
  const int splits = gram_topk_splits(c, rows);
  Scratch pd, pi;
  SCEDAR_CHECK(pd.alloc(sizeof(double) * splits * rows * K, s));
  SCEDAR_CHECK(pi.alloc(sizeof(int32_t) * splits * rows * K, s));
  SCEDAR_CHECK(launch_gram_topk(c, row_begin, row_end, K, splits, pd.as<double>(),
                                pi.as<int32_t>(), s));
  return launch_topk_merge(pd.as<double>(), pi.as<int32_t>(), splits, rows, K, row_begin, k,
                           exclude, idx_out, dist_out, s);
}

}  // namespace scedar

using namespace scedar;

extern "C" {

int scedar_abi_version(void) { return SCEDAR_ABI_VERSION; }
const char* scedar_last_error(void) { return g_last_error.c_str(); }
int scedar_max_k(void) { return 127; }
long long scedar_launch_count(void) { return g_launches.load(); }

int scedar_cells_from_dense(const double* x, int64_t n, int64_t p, int64_t ldx, int metric,
                            void* stream, scedar_cells_t** out) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n > 0 && p > 0 && !x) return fail(SCEDAR_ERR_ARG, "x is NULL");
  if (ldx < p) return fail(SCEDAR_ERR_ARG, "ldx must be >= p");
  SCEDAR_CHECK(new_cells(n, p, metric, s, out));
  scedar_cells* c = *out;
  int rc = launch_prep_dense(x, n, p, ldx, metric, c->xw, c->n_pad, c->p_pad, s);
  if (rc == SCEDAR_OK) rc = launch_gram_diag(c, s);
  if (rc != SCEDAR_OK) {
    scedar_cells_free(c);
    *out = nullptr;
  }
  return rc;
}

int scedar_cells_from_csr(const void* indptr, int indptr_is_64, const int32_t* indices,
                          const double* data, int64_t n, int64_t p, int metric, void* stream,
                          scedar_cells_t** out) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (metric == SCEDAR_METRIC_CORRELATION)
    return fail(SCEDAR_ERR_ARG, "correlation is not defined for sparse cells (sdm.py:1196-1204)");
  if (!indptr) return fail(SCEDAR_ERR_ARG, "indptr is NULL");
  SCEDAR_CHECK(new_cells(n, p, metric, s, out));
  scedar_cells* c = *out;
  int rc = launch_prep_csr(indptr, indptr_is_64, indices, data, n, p, c->xw, c->n_pad, c->p_pad, s);
  if (rc == SCEDAR_OK) rc = launch_gram_diag(c, s);
  if (rc != SCEDAR_OK) {
    scedar_cells_free(c);
    *out = nullptr;
  }
  return rc;
}

int scedar_cells_alloc(int64_t n, int64_t p, int metric, void* stream, scedar_cells_t** out) {
  return new_cells(n, p, metric, static_cast<cudaStream_t>(stream), out);
}

int scedar_cells_fill_rows(scedar_cells_t* c, const double* x_rows, int64_t ldx,
                           int64_t row_begin, int64_t row_end, void* stream) {
  return scedar_cells_fill_rows_ex(c, x_rows, ldx, row_begin, row_end, 1, stream);
}

int scedar_cells_stat_rows(scedar_cells_t* c, int64_t row_begin, int64_t row_end, void* stream) {
  if (!c) return fail(SCEDAR_ERR_ARG, "cells handle is NULL");
  if (row_begin < 0 || row_end < row_begin || row_end > c->n)
    return fail(SCEDAR_ERR_ARG, "row range must satisfy 0 <= row_begin <= row_end <= n");
  if (row_begin % kRowPad != 0 || (row_end % kRowPad != 0 && row_end != c->n))
    return fail(SCEDAR_ERR_ARG, "row ranges must be 128-aligned (the last one may end at n)");
  if (row_end == row_begin) return SCEDAR_OK;
  return launch_gram_diag_rows(c, row_begin, row_end == c->n ? c->n_pad : row_end,
                               static_cast<cudaStream_t>(stream));
}

int scedar_cells_fill_rows_ex(scedar_cells_t* c, const double* x_rows, int64_t ldx,
                              int64_t row_begin, int64_t row_end, int with_stat, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!c) return fail(SCEDAR_ERR_ARG, "cells handle is NULL");
  if (row_begin < 0 || row_end < row_begin || row_end > c->n)
    return fail(SCEDAR_ERR_ARG, "row range must satisfy 0 <= row_begin <= row_end <= n");
  if (row_begin % kRowPad != 0 || (row_end % kRowPad != 0 && row_end != c->n))
    return fail(SCEDAR_ERR_ARG, "row ranges must be 128-aligned (the last one may end at n)");
  if (row_end == row_begin) return SCEDAR_OK;
  if (!x_rows && c->p > 0) return fail(SCEDAR_ERR_ARG, "x_rows is NULL");
  if (ldx < c->p) return fail(SCEDAR_ERR_ARG, "ldx must be >= p");
  // the range that ends at n also zero-fills the padding cells
  const int64_t pad_end = row_end == c->n ? c->n_pad : row_end;
  SCEDAR_CHECK(launch_prep_dense(x_rows, row_end - row_begin, c->p, ldx, c->metric,
                                 c->xw + row_begin * c->p_pad, pad_end - row_begin, c->p_pad, s));
  // the Gram diagonal (the statistic of the DMMA / 3xFP16 kernels) may be left
  // to one later scedar_cells_stat_rows pass: the INT8 split kernels take theirs
  // from the slices
  if (!with_stat) return SCEDAR_OK;
  return launch_gram_diag_rows(c, row_begin, pad_end, s);
}

int scedar_pdist_lower_rows(const scedar_cells_t* c, int64_t row_begin, int64_t row_end,
                            double* d, int64_t ldd, void* stream) {
  SCEDAR_CHECK(check_rows(c, row_begin, row_end));
  if (row_begin % kRowPad != 0)
    return fail(SCEDAR_ERR_ARG, "row_begin must be 128-aligned");
  if (row_end == row_begin) return SCEDAR_OK;
  if (!d) return fail(SCEDAR_ERR_ARG, "d is NULL");
  if (ldd < c->n) return fail(SCEDAR_ERR_ARG, "ldd must be >= n");
  return launch_gram_lower_rows(c, row_begin, row_end, d, ldd, static_cast<cudaStream_t>(stream));
}

int scedar_pdist_lower_rows_prec(const scedar_cells_t* c, int precision, int64_t row_begin,
                                 int64_t row_end, double* d, int64_t ldd, void* stream) {
  SCEDAR_CHECK(check_rows(c, row_begin, row_end));
  SCEDAR_CHECK(check_precision(precision));
  if (precision == SCEDAR_PRECISION_FAST)
    return fail(SCEDAR_ERR_UNSUPPORTED, "row ranges of the lower triangle are FP64-grade only");
  const int S = oz_slices(c, precision);
  if (!S) return scedar_pdist_lower_rows(c, row_begin, row_end, d, ldd, stream);
  if (row_begin % kRowPad != 0)
    return fail(SCEDAR_ERR_ARG, "row_begin must be 128-aligned");
  if (row_end == row_begin) return SCEDAR_OK;
  if (!d) return fail(SCEDAR_ERR_ARG, "d is NULL");
  if (ldd < c->n) return fail(SCEDAR_ERR_ARG, "ldd must be >= n");
  return launch_oz_lower_rows(c, S, row_begin, row_end, d, ldd, static_cast<cudaStream_t>(stream));
}

int scedar_cells_free(scedar_cells_t* c) {
  if (!c) return SCEDAR_OK;
  if (c->fast) fast_state_free(c->fast, c->stream);
  if (c->oz) oz_state_free(c->oz, c->stream);
  if (c->xw) cudaFreeAsync(c->xw, c->stream);
  if (c->stat) cudaFreeAsync(c->stat, c->stream);
  delete c;
  return SCEDAR_OK;
}

int64_t scedar_cells_n(const scedar_cells_t* c) { return c ? c->n : -1; }
int64_t scedar_cells_p(const scedar_cells_t* c) { return c ? c->p : -1; }
int scedar_cells_metric(const scedar_cells_t* c) { return c ? c->metric : -1; }

int scedar_pdist_stripe(const scedar_cells_t* c, int precision, int64_t row_begin,
                        int64_t row_end, double* d_out, int64_t ldd, void* stream) {
  SCEDAR_CHECK(check_rows(c, row_begin, row_end));
  SCEDAR_CHECK(check_precision(precision));
  if (row_end == row_begin) return SCEDAR_OK;
  if (!d_out) return fail(SCEDAR_ERR_ARG, "d_out is NULL");
  if (ldd < c->n) return fail(SCEDAR_ERR_ARG, "ldd must be >= n");
  if (precision == SCEDAR_PRECISION_FAST)
    return launch_fast_stripe(c, row_begin, row_end, d_out, ldd, static_cast<cudaStream_t>(stream));
  // the split kernel takes 128-aligned stripes; ragged ones run on DMMA
  const bool aligned = row_begin % kRowPad == 0 && (row_end % kRowPad == 0 || row_end == c->n);
  if (const int S = aligned ? oz_slices(c, precision) : 0)
    return launch_oz_stripe(c, S, row_begin, row_end, d_out, ldd, static_cast<cudaStream_t>(stream));
  return launch_gram_stripe(c, row_begin, row_end, d_out, ldd, static_cast<cudaStream_t>(stream));
}

int scedar_pdist_symmetric(const scedar_cells_t* c, int precision, double* d_out, int64_t ldd,
                           void* stream) {
  if (!c) return fail(SCEDAR_ERR_ARG, "cells handle is NULL");
  SCEDAR_CHECK(check_precision(precision));
  if (c->n == 0) return SCEDAR_OK;
  if (!d_out) return fail(SCEDAR_ERR_ARG, "d_out is NULL");
  if (ldd < c->n) return fail(SCEDAR_ERR_ARG, "ldd must be >= n");
  if (precision == SCEDAR_PRECISION_FAST)
    return launch_fast_stripe(c, 0, c->n, d_out, ldd, static_cast<cudaStream_t>(stream), 1);
  if (const int S = oz_slices(c, precision))
    return launch_oz_symmetric(c, S, d_out, ldd, static_cast<cudaStream_t>(stream));
  return launch_gram_symmetric(c, d_out, ldd, static_cast<cudaStream_t>(stream));
}

int scedar_pdist_stripe_p2p(const scedar_cells_t* c, int precision, int rank, int world,
                            const int64_t* stripe_begin, double* const* stripe_ptr,
                            int64_t ldd, void* stream) {
  if (!c) return fail(SCEDAR_ERR_ARG, "cells handle is NULL");
  SCEDAR_CHECK(check_precision(precision));
  if (precision == SCEDAR_PRECISION_FAST)
    return fail(SCEDAR_ERR_UNSUPPORTED, "peer-store stripes are FP64-grade only");
  if (!stripe_begin || !stripe_ptr) return fail(SCEDAR_ERR_ARG, "NULL stripe table");
  if (ldd < c->n) return fail(SCEDAR_ERR_ARG, "ldd must be >= n");
  if (const int S = oz_slices(c, precision))
    return launch_oz_stripe_p2p(c, S, rank, world, stripe_begin, stripe_ptr, ldd,
                                static_cast<cudaStream_t>(stream));
  return launch_gram_stripe_p2p(c, rank, world, stripe_begin, stripe_ptr, ldd,
                                static_cast<cudaStream_t>(stream));
}

int scedar_pdist_p2p_rows(const scedar_cells_t* c, int precision, int rank, int world,
                          const int64_t* stripe_begin, double* const* stripe_ptr, int64_t ldd,
                          int64_t row_begin, int64_t row_end, int reserve_sms, void* stream) {
  SCEDAR_CHECK(check_rows(c, row_begin, row_end));
  SCEDAR_CHECK(check_precision(precision));
  if (!stripe_begin || !stripe_ptr) return fail(SCEDAR_ERR_ARG, "NULL stripe table");
  if (ldd < c->n) return fail(SCEDAR_ERR_ARG, "ldd must be >= n");
  if (reserve_sms < 0) return fail(SCEDAR_ERR_ARG, "reserve_sms must be >= 0");
  const int S = oz_slices(c, precision);
  if (!S)
    return fail(SCEDAR_ERR_UNSUPPORTED,
                "row ranges of the peer-store contraction run on the INT8 split kernel "
                "(precision fp64_i8 / fp64_i8s6)");
  return launch_oz_p2p_rows(c, S, rank, world, stripe_begin, stripe_ptr, ldd, row_begin, row_end,
                            reserve_sms, static_cast<cudaStream_t>(stream));
}

int scedar_pdist_p2p_pieces(const scedar_cells_t* c, int precision, int rank, int world,
                            const int64_t* stripe_begin, double* const* stripe_ptr, int64_t ldd,
                            const int32_t* piece_of_tile, int piece, int n_pieces,
                            int reserve_sms, void* stream) {
  if (!c) return fail(SCEDAR_ERR_ARG, "cells handle is NULL");
  SCEDAR_CHECK(check_precision(precision));
  if (!stripe_begin || !stripe_ptr || !piece_of_tile)
    return fail(SCEDAR_ERR_ARG, "NULL stripe / piece table");
  if (ldd < c->n) return fail(SCEDAR_ERR_ARG, "ldd must be >= n");
  if (reserve_sms < 0) return fail(SCEDAR_ERR_ARG, "reserve_sms must be >= 0");
  if (c->n == 0) return SCEDAR_OK;
  const int S = oz_slices(c, precision);
  if (!S)
    return fail(SCEDAR_ERR_UNSUPPORTED,
                "the piece form of the peer-store contraction runs on the INT8 split kernel "
                "(precision fp64_i8 / fp64_i8s6)");
  return launch_oz_p2p_pieces(c, S, rank, world, stripe_begin, stripe_ptr, ldd, piece_of_tile,
                              piece, n_pieces, reserve_sms, static_cast<cudaStream_t>(stream));
}

// ---- PCA producer pieces ------------------------------------------------------
int scedar_cells_col_means(const scedar_cells_t* c, double* mean_out, void* stream) {
  if (!c) return fail(SCEDAR_ERR_ARG, "cells handle is NULL");
  if (c->p > 0 && !mean_out) return fail(SCEDAR_ERR_ARG, "mean_out is NULL");
  return launch_col_means(c, mean_out, static_cast<cudaStream_t>(stream));
}

int scedar_cells_shift_cols(scedar_cells_t* c, const double* shift, void* stream) {
  if (!c) return fail(SCEDAR_ERR_ARG, "cells handle is NULL");
  if (c->p > 0 && !shift) return fail(SCEDAR_ERR_ARG, "shift is NULL");
  return launch_shift_cols(c, shift, static_cast<cudaStream_t>(stream));
}

int scedar_cells_transposed(const scedar_cells_t* c, const double* shift, void* stream,
                            scedar_cells_t** out) {
  if (!c) return fail(SCEDAR_ERR_ARG, "cells handle is NULL");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SCEDAR_CHECK(new_cells(c->p, c->n, c->metric, s, out));
  int rc = launch_transpose_cells(c, shift, *out, s);
  if (rc != SCEDAR_OK) {
    scedar_cells_free(*out);
    *out = nullptr;
  }
  return rc;
}

int scedar_gram_raw(const scedar_cells_t* c, double* out, int64_t ldd, void* stream) {
  if (!c) return fail(SCEDAR_ERR_ARG, "cells handle is NULL");
  if (c->n == 0) return SCEDAR_OK;
  if (!out) return fail(SCEDAR_ERR_ARG, "out is NULL");
  if (ldd < c->n) return fail(SCEDAR_ERR_ARG, "ldd must be >= n");
  return launch_gram_raw_symmetric(c, out, ldd, static_cast<cudaStream_t>(stream));
}

int scedar_gram_raw_cross(const scedar_cells_t* a, const scedar_cells_t* b, double* out,
                          int64_t ldd, void* stream) {
  if (!a || !b) return fail(SCEDAR_ERR_ARG, "cells handle is NULL");
  if (a->p != b->p) return fail(SCEDAR_ERR_ARG, "operands differ in their column count");
  if (a->n == 0 || b->n == 0) return SCEDAR_OK;
  if (!out) return fail(SCEDAR_ERR_ARG, "out is NULL");
  if (ldd < b->n) return fail(SCEDAR_ERR_ARG, "ldd must be >= the row count of b");
  return launch_gram_raw_cross(a, b, out, ldd, static_cast<cudaStream_t>(stream));
}

// ---- peer-mapped stripes (one process per GPU, same node) ------------------
int scedar_stripe_alloc(size_t bytes, void** dev_ptr) {
  if (!dev_ptr) return fail(SCEDAR_ERR_ARG, "dev_ptr is NULL");
  SCEDAR_CUDA(cudaMalloc(dev_ptr, bytes ? bytes : 16));
  return SCEDAR_OK;
}
int scedar_stripe_free(void* dev_ptr) {
  if (dev_ptr) SCEDAR_CUDA(cudaFree(dev_ptr));
  return SCEDAR_OK;
}
int scedar_stripe_export(void* dev_ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  if (!dev_ptr || !handle64) return fail(SCEDAR_ERR_ARG, "NULL argument");
  cudaIpcMemHandle_t h;
  SCEDAR_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
  memcpy(handle64, &h, sizeof(h));
  return SCEDAR_OK;
}
int scedar_stripe_open(const void* handle64, void** peer_ptr) {
  if (!peer_ptr || !handle64) return fail(SCEDAR_ERR_ARG, "NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  SCEDAR_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SCEDAR_OK;
}
int scedar_stripe_close(void* peer_ptr) {
  if (peer_ptr) SCEDAR_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return SCEDAR_OK;
}

int scedar_knn_from_d(const double* d, int64_t ldd, int64_t n_cols, int64_t row_begin,
                      int64_t row_end, int k, int exclude, int64_t* idx_out, double* dist_out,
                      void* stream) {
  if (row_begin < 0 || row_end < row_begin || row_end > n_cols)
    return fail(SCEDAR_ERR_ARG, "row range must satisfy 0 <= row_begin <= row_end <= n_cols");
  if (k < 0 || k > n_cols - 1) return fail(SCEDAR_ERR_ARG, "k should >= 0 and <= n_samples-1");
  if (exclude != SCEDAR_EXCLUDE_FIRST && exclude != SCEDAR_EXCLUDE_SELF)
    return fail(SCEDAR_ERR_ARG, "exclude must be 0 (first) or 1 (self)");
  if (k == 0 || row_end == row_begin) return SCEDAR_OK;
  if (!d || !idx_out || !dist_out) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  if (ldd < n_cols) return fail(SCEDAR_ERR_ARG, "ldd must be >= n_cols");
  return launch_topk_from_d(d, ldd, n_cols, row_begin, row_end, k, exclude, idx_out, dist_out,
                            static_cast<cudaStream_t>(stream));
}

int scedar_cells_set_blockmin(scedar_cells_t* c, double* const* stripe_bmin, int world,
                              int64_t ldb) {
  if (!c) return fail(SCEDAR_ERR_ARG, "cells is NULL");
  if (world < 0 || world > 16) return fail(SCEDAR_ERR_ARG, "world must be in 0..16");
  if (world > 0) {
    if (!stripe_bmin) return fail(SCEDAR_ERR_ARG, "NULL buffer table");
    if (ldb < (c->n + 31) / 32) return fail(SCEDAR_ERR_ARG, "ldb must be >= ceil(n / 32)");
    for (int r = 0; r < world; ++r)
      if (!stripe_bmin[r]) return fail(SCEDAR_ERR_ARG, "NULL block-minima buffer");
  }
  for (int r = 0; r < 16; ++r) c->bmin[r] = r < world ? stripe_bmin[r] : nullptr;
  c->bmin_count = world;
  c->bmin_ld = world > 0 ? ldb : 0;
  c->bmin_filled = false;
  return SCEDAR_OK;
}

int scedar_knn_from_d_blockmin(const scedar_cells_t* c, const double* d, int64_t ldd,
                               const double* bmin, int64_t ldb, int64_t row_begin,
                               int64_t row_end, int k, int exclude, int64_t* idx_out,
                               double* dist_out, void* stream) {
  if (!c) return fail(SCEDAR_ERR_ARG, "cells is NULL");
  SCEDAR_CHECK(check_rows(c, row_begin, row_end));
  if (k < 0 || k > c->n - 1) return fail(SCEDAR_ERR_ARG, "k should >= 0 and <= n_samples-1");
  if (exclude != SCEDAR_EXCLUDE_FIRST && exclude != SCEDAR_EXCLUDE_SELF)
    return fail(SCEDAR_ERR_ARG, "exclude must be 0 (first) or 1 (self)");
  if (k == 0 || row_end == row_begin) return SCEDAR_OK;
  if (!d || !bmin || !idx_out || !dist_out) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  if (ldd < c->n) return fail(SCEDAR_ERR_ARG, "ldd must be >= n");
  if (ldb < (c->n + 31) / 32) return fail(SCEDAR_ERR_ARG, "ldb must be >= ceil(n / 32)");
  if (!c->bmin_filled)
    return fail(SCEDAR_ERR_ARG,
                "no INT8 split launch has written block minima for this handle");
  return launch_topk_blockmin(d, ldd, c->n, bmin, ldb, row_begin, row_end, k, exclude, idx_out,
                              dist_out, static_cast<cudaStream_t>(stream));
}

int scedar_knn_stripe(const scedar_cells_t* c, int precision, int k, int exclude,
                      int64_t row_begin, int64_t row_end, int64_t* idx_out, double* dist_out,
                      void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SCEDAR_CHECK(check_rows(c, row_begin, row_end));
  SCEDAR_CHECK(check_precision(precision));
  if (k < 0 || k > c->n - 1) return fail(SCEDAR_ERR_ARG, "k should >= 0 and <= n_samples-1");
  if (exclude != SCEDAR_EXCLUDE_FIRST && exclude != SCEDAR_EXCLUDE_SELF)
    return fail(SCEDAR_ERR_ARG, "exclude must be 0 (first) or 1 (self)");
  if (k == 0 || row_end == row_begin) return SCEDAR_OK;
  if (!idx_out || !dist_out) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  // the tensor-core pass proposes 32 candidates per cell; larger k runs FP64
  if (precision == SCEDAR_PRECISION_FAST && k <= fast_knn_max_k())
    return launch_fast_knn(c, k, exclude, row_begin, row_end, idx_out, dist_out, s);
  return knn_stripe_fp64(c, k, exclude, row_begin, row_end, idx_out, dist_out, s,
                         oz_slices(c, precision));
}

int scedar_col_sort(const double* d, int64_t ldd, int64_t rows, int64_t n_cols,
                    double* sorted_out, int64_t* argsorted_out, void* stream) {
  if (rows < 0 || n_cols < 0 || ldd < n_cols)
    return fail(SCEDAR_ERR_ARG, "need rows >= 0, n_cols >= 0 and ldd >= n_cols");
  if (n_cols >= 0x7fffffffLL) return fail(SCEDAR_ERR_UNSUPPORTED, "n_cols must fit in int32");
  if (rows == 0 || n_cols == 0 || (!sorted_out && !argsorted_out)) return SCEDAR_OK;
  if (!d) return fail(SCEDAR_ERR_ARG, "d is NULL");
  return launch_col_sort(d, ldd, rows, n_cols, sorted_out, argsorted_out,
                         static_cast<cudaStream_t>(stream));
}

int scedar_num_correct_dist_mat(double* d, int64_t ldd, int64_t n, int has_upper_bound,
                                double upper_bound, int32_t* flags_out, void* stream) {
  if (n < 0 || ldd < n) return fail(SCEDAR_ERR_ARG, "dmat must be a 2D (n_samples, n_samples) array");
  if (!flags_out) return fail(SCEDAR_ERR_ARG, "flags_out is NULL");
  if (n > 0 && !d) return fail(SCEDAR_ERR_ARG, "d is NULL");
  return launch_ncdm(d, ldd, n, has_upper_bound, upper_bound, flags_out,
                     static_cast<cudaStream_t>(stream));
}

// ---- consumers --------------------------------------------------------------
int scedar_knn_connectivity_csr(const int64_t* idx, const double* dist, int64_t n, int k,
                                int64_t* indptr_out, int32_t* indices_out, double* data_out,
                                void* stream) {
  if (n < 0 || k < 0) return fail(SCEDAR_ERR_ARG, "n and k must be >= 0");
  if (!indptr_out) return fail(SCEDAR_ERR_ARG, "indptr_out is NULL");
  if (n * (int64_t)k > 0 && (!idx || !dist || !indices_out || !data_out))
    return fail(SCEDAR_ERR_ARG, "NULL buffer");
  return launch_conn_csr(idx, dist, n, k, indptr_out, indices_out, data_out,
                         static_cast<cudaStream_t>(stream));
}

int scedar_knn_affinity(const double* data, int64_t nnz, double aff_scale, double* weights_out,
                        void* stream) {
  if (nnz < 0) return fail(SCEDAR_ERR_ARG, "nnz must be >= 0");
  if (nnz > 0 && (!data || !weights_out)) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  return launch_affinity(data, nnz, aff_scale, weights_out, static_cast<cudaStream_t>(stream));
}

int scedar_masked_kth(const double* d, int64_t ldd, int64_t n_cols, int64_t row_begin,
                      int64_t row_end, const uint8_t* alive, int rank, double* kth_out,
                      void* stream) {
  if (row_begin < 0 || row_end < row_begin || row_end > n_cols)
    return fail(SCEDAR_ERR_ARG, "row range must satisfy 0 <= row_begin <= row_end <= n_cols");
  if (rank < 0) return fail(SCEDAR_ERR_ARG, "rank must be >= 0");
  if (ldd < n_cols) return fail(SCEDAR_ERR_ARG, "ldd must be >= n_cols");
  if (row_end == row_begin) return SCEDAR_OK;
  if (!d || !alive || !kth_out) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  return launch_masked_kth(d, ldd, n_cols, row_begin, row_end, alive, rank, kth_out,
                           static_cast<cudaStream_t>(stream));
}

int scedar_rare_update(const double* kth, int64_t ldk, int64_t n, double cutoff, int iter,
                       uint8_t* alive, int32_t* last_kept, unsigned long long* n_kept_out,
                       void* stream) {
  if (n < 0 || ldk < 1) return fail(SCEDAR_ERR_ARG, "need n >= 0 and ldk >= 1");
  if (n == 0) return SCEDAR_OK;
  if (!kth || !alive || !last_kept || !n_kept_out) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  return launch_rare_update(kth, ldk, n, cutoff, iter, alive, last_kept, n_kept_out,
                            static_cast<cudaStream_t>(stream));
}

int scedar_masked_max(const double* v, int64_t ldv, int64_t n, const uint8_t* alive,
                      double* max_out, void* stream) {
  if (n < 0 || ldv < 1) return fail(SCEDAR_ERR_ARG, "need n >= 0 and ldv >= 1");
  if (n == 0) return SCEDAR_OK;
  if (!v || !max_out) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  return launch_masked_max(v, ldv, n, alive, reinterpret_cast<unsigned long long*>(max_out),
                           static_cast<cudaStream_t>(stream));
}

int scedar_compact_alive(const uint8_t* alive, int64_t n, int64_t* sel_out, void* stream) {
  if (n < 0) return fail(SCEDAR_ERR_ARG, "n must be >= 0");
  if (n == 0) return SCEDAR_OK;
  if (!alive || !sel_out) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  return launch_compact_alive(alive, n, sel_out, static_cast<cudaStream_t>(stream));
}

int scedar_gather_rows(const double* x, int64_t ldx, int64_t p, const int64_t* sel, int64_t m,
                       double* out, int64_t ldo, void* stream) {
  if (m < 0 || p < 0 || ldx < p || ldo < p)
    return fail(SCEDAR_ERR_ARG, "need m >= 0, p >= 0, ldx >= p and ldo >= p");
  if (m == 0 || p == 0) return SCEDAR_OK;
  if (!x || !sel || !out) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  return launch_gather_rows(x, ldx, p, sel, m, out, ldo, static_cast<cudaStream_t>(stream));
}

int scedar_squareform_rows(const double* d, int64_t ldd, int64_t n, int64_t row_begin,
                           int64_t row_end, double* condensed_out, int32_t* flags_out,
                           void* stream) {
  if (n < 0 || row_begin < 0 || row_end < row_begin || row_end > n)
    return fail(SCEDAR_ERR_ARG, "row range must satisfy 0 <= row_begin <= row_end <= n");
  if (ldd < n) return fail(SCEDAR_ERR_ARG, "ldd must be >= n");
  if (!flags_out) return fail(SCEDAR_ERR_ARG, "flags_out is NULL");
  if (row_end == row_begin) return SCEDAR_OK;
  // the last row of D has no entry above the diagonal, but its diagonal is checked
  const int64_t seg = row_end * (2 * n - row_end - 1) / 2 - row_begin * (2 * n - row_begin - 1) / 2;
  if (!d || (!condensed_out && seg > 0)) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  return launch_squareform_rows(d, ldd, n, row_begin, row_end, condensed_out, flags_out,
                                static_cast<cudaStream_t>(stream));
}

int scedar_d_gather_block(const double* d, int64_t ldd, int64_t n_rows, int64_t n_cols,
                          const int64_t* row_sel, int64_t m, const int64_t* col_sel, int64_t c,
                          double* out, int64_t ldo, int32_t* flags_out, void* stream) {
  if (n_rows < 0 || n_cols < 0 || m < 0 || c < 0 || ldd < n_cols || ldo < c)
    return fail(SCEDAR_ERR_ARG, "need non-negative sizes, ldd >= n_cols and ldo >= c");
  if (!flags_out) return fail(SCEDAR_ERR_ARG, "flags_out is NULL");
  if (m == 0 || c == 0) return SCEDAR_OK;
  if (!d || !row_sel || !col_sel || !out) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  return launch_gather_block(d, ldd, n_rows, n_cols, row_sel, m, col_sel, c, out, ldo,
                             flags_out, static_cast<cudaStream_t>(stream));
}

int scedar_impute_iter(const double* x_curr, int64_t ldx, int64_t n, int64_t p,
                       const int64_t* knn, int k, int n_do_i, double min_present_val, int iter,
                       int64_t row_begin, int64_t row_end, double* x_next, int64_t ldn,
                       int32_t* idc, int64_t ldi, unsigned long long* stats_out, void* stream) {
  if (row_begin < 0 || row_end < row_begin || row_end > n)
    return fail(SCEDAR_ERR_ARG, "row range must satisfy 0 <= row_begin <= row_end <= n");
  if (n < 0 || p < 0 || ldx < p || ldn < p || ldi < p)
    return fail(SCEDAR_ERR_ARG, "need n >= 0, p >= 0 and row strides >= p");
  if (k < 1 || k >= n) return fail(SCEDAR_ERR_ARG, "k should be >= 1 and < n_samples");
  if (n_do_i > k || n_do_i < 1) return fail(SCEDAR_ERR_ARG, "n_do should be <= k and >= 1");
  if (n == 0 || p == 0) return SCEDAR_OK;
  if (!x_curr || !knn || !x_next || !idc) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  if (x_curr == x_next) return fail(SCEDAR_ERR_ARG, "x_next must not alias x_curr");
  return launch_impute_iter(x_curr, ldx, n, p, knn, k, n_do_i, min_present_val, iter, row_begin,
                            row_end, x_next, ldn, idc, ldi, stats_out,
                            static_cast<cudaStream_t>(stream));
}

// ---- one-shot forms --------------------------------------------------------
// every argument is checked before the first CUDA call, so a bad request costs
// no device work (and is reported as ERR_ARG even without a device)
static int check_oneshot(int64_t n, int64_t p, int metric, int precision, int64_t row_begin,
                         int64_t row_end) {
  SCEDAR_CHECK(check_metric(metric));
  SCEDAR_CHECK(check_precision(precision));
  if (n < 0 || p < 0) return fail(SCEDAR_ERR_ARG, "n and p must be >= 0");
  if (row_begin < 0 || row_end < row_begin || row_end > n)
    return fail(SCEDAR_ERR_ARG, "row range must satisfy 0 <= row_begin <= row_end <= n");
  return SCEDAR_OK;
}

int scedar_pdist_dense(const double* x, int64_t n, int64_t p, int metric, int precision,
                       int64_t row_begin, int64_t row_end, double* d_out, void* stream) {
  SCEDAR_CHECK(check_oneshot(n, p, metric, precision, row_begin, row_end));
  if (row_end > row_begin && !d_out) return fail(SCEDAR_ERR_ARG, "d_out is NULL");
  scedar_cells_t* c = nullptr;
  SCEDAR_CHECK(scedar_cells_from_dense(x, n, p, p, metric, stream, &c));
  int rc = scedar_pdist_stripe(c, precision, row_begin, row_end, d_out, n, stream);
  if (rc == SCEDAR_OK && cudaStreamSynchronize(static_cast<cudaStream_t>(stream)) != cudaSuccess)
    rc = fail(SCEDAR_ERR_CUDA, "stream synchronize failed");
  scedar_cells_free(c);
  return rc;
}

int scedar_pdist_csr(const void* indptr, int indptr_is_64, const int32_t* indices,
                     const double* data, int64_t n, int64_t p, int metric, int precision,
                     int64_t row_begin, int64_t row_end, double* d_out, void* stream) {
  SCEDAR_CHECK(check_oneshot(n, p, metric, precision, row_begin, row_end));
  if (row_end > row_begin && !d_out) return fail(SCEDAR_ERR_ARG, "d_out is NULL");
  scedar_cells_t* c = nullptr;
  SCEDAR_CHECK(scedar_cells_from_csr(indptr, indptr_is_64, indices, data, n, p, metric, stream, &c));
  int rc = scedar_pdist_stripe(c, precision, row_begin, row_end, d_out, n, stream);
  if (rc == SCEDAR_OK && cudaStreamSynchronize(static_cast<cudaStream_t>(stream)) != cudaSuccess)
    rc = fail(SCEDAR_ERR_CUDA, "stream synchronize failed");
  scedar_cells_free(c);
  return rc;
}

int scedar_knn_dense(const double* x, int64_t n, int64_t p, int metric, int precision, int k,
                     int exclude, int64_t row_begin, int64_t row_end, int64_t* idx_out,
                     double* dist_out, void* stream) {
  SCEDAR_CHECK(check_oneshot(n, p, metric, precision, row_begin, row_end));
  if (k < 0 || k > n - 1) return fail(SCEDAR_ERR_ARG, "k should >= 0 and <= n_samples-1");
  if (exclude != SCEDAR_EXCLUDE_FIRST && exclude != SCEDAR_EXCLUDE_SELF)
    return fail(SCEDAR_ERR_ARG, "exclude must be 0 (first) or 1 (self)");
  if (k > 0 && row_end > row_begin && (!idx_out || !dist_out))
    return fail(SCEDAR_ERR_ARG, "NULL buffer");
  scedar_cells_t* c = nullptr;
  SCEDAR_CHECK(scedar_cells_from_dense(x, n, p, p, metric, stream, &c));
  int rc = scedar_knn_stripe(c, precision, k, exclude, row_begin, row_end, idx_out, dist_out, stream);
  if (rc == SCEDAR_OK && cudaStreamSynchronize(static_cast<cudaStream_t>(stream)) != cudaSuccess)
    rc = fail(SCEDAR_ERR_CUDA, "stream synchronize failed");
  scedar_cells_free(c);
  return rc;
}

// ---- host-buffer forms -----------------------------------------------------
int scedar_pdist_dense_host(const double* x, int64_t n, int64_t p, int metric, int precision,
                            double* d_out, int device) {
  SCEDAR_CHECK(check_metric(metric));
  SCEDAR_CHECK(check_precision(precision));
  if (n < 0 || p < 0) return fail(SCEDAR_ERR_ARG, "n and p must be >= 0");
  if (n == 0) return SCEDAR_OK;
  if (!d_out || (p > 0 && !x)) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  SCEDAR_CUDA(cudaSetDevice(device));
  cudaStream_t s0 = nullptr, s1 = nullptr;
  double* xd = nullptr;
  double* dbuf[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  scedar_cells_t* c = nullptr;
  int rc = SCEDAR_OK;
  auto cleanup = [&]() {
    if (c) scedar_cells_free(c);
    if (xd) cudaFree(xd);
    for (int b = 0; b < 2; ++b) {
      if (dbuf[b]) cudaFree(dbuf[b]);
      if (done[b]) cudaEventDestroy(done[b]);
    }
    if (s0) cudaStreamDestroy(s0);
    if (s1) cudaStreamDestroy(s1);
  };
#define HOST_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t e__ = (expr);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      rc = fail(e__ == cudaErrorMemoryAllocation ? SCEDAR_ERR_NOMEM : SCEDAR_ERR_CUDA, \
                std::string(#expr) + ": " + cudaGetErrorString(e__));          \
      cleanup();                                                               \
      return rc;                                                               \
    }                                                                          \
  } while (0)
  HOST_CUDA(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
  HOST_CUDA(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  HOST_CUDA(cudaMalloc(&xd, sizeof(double) * std::max<int64_t>(n * p, 1)));
  HOST_CUDA(cudaMemcpyAsync(xd, x, sizeof(double) * n * p, cudaMemcpyHostToDevice, s0));
  rc = scedar_cells_from_dense(xd, n, p, p, metric, s0, &c);
  if (rc != SCEDAR_OK) {
    cleanup();
    return rc;
  }
  // stripes of <= 1 GiB, double buffered: compute on s0, copy back on s1
  int64_t chunk = (int64_t(1) << 30) / (8 * n);
  chunk = std::max<int64_t>(128, chunk / 128 * 128);
  chunk = std::min(chunk, round_up(n, 128));
  for (int b = 0; b < 2; ++b) {
    HOST_CUDA(cudaMalloc(&dbuf[b], sizeof(double) * chunk * n));
    HOST_CUDA(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
  }
  cudaEvent_t ready = nullptr;
  HOST_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  int b = 0;
  for (int64_t r0 = 0; r0 < n; r0 += chunk, b ^= 1) {
    const int64_t r1 = std::min(n, r0 + chunk);
    // the copy that last used this buffer must have drained
    cudaStreamWaitEvent(s0, done[b], 0);
    rc = launch_gram_stripe(c, r0, r1, dbuf[b], n, s0);
    if (rc != SCEDAR_OK) break;
    cudaEventRecord(ready, s0);
    cudaStreamWaitEvent(s1, ready, 0);
    cudaMemcpyAsync(d_out + r0 * n, dbuf[b], sizeof(double) * (r1 - r0) * n,
                    cudaMemcpyDeviceToHost, s1);
    cudaEventRecord(done[b], s1);
  }
  cudaError_t e0 = cudaStreamSynchronize(s0), e1 = cudaStreamSynchronize(s1);
  cudaEventDestroy(ready);
  if (rc == SCEDAR_OK && (e0 != cudaSuccess || e1 != cudaSuccess))
    rc = fail(SCEDAR_ERR_CUDA, std::string("pdist stripes: ") +
                                   cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
  cleanup();
  return rc;
}

int scedar_knn_dense_host(const double* x, int64_t n, int64_t p, int metric, int precision, int k,
                          int exclude, int64_t* idx_out, double* dist_out, int device) {
  SCEDAR_CHECK(check_metric(metric));
  SCEDAR_CHECK(check_precision(precision));
  if (n < 0 || p < 0) return fail(SCEDAR_ERR_ARG, "n and p must be >= 0");
  if (k < 0 || k > n - 1) return fail(SCEDAR_ERR_ARG, "k should >= 0 and <= n_samples-1");
  if (exclude != SCEDAR_EXCLUDE_FIRST && exclude != SCEDAR_EXCLUDE_SELF)
    return fail(SCEDAR_ERR_ARG, "exclude must be 0 (first) or 1 (self)");
  if (n == 0 || k == 0) return SCEDAR_OK;
  if (!idx_out || !dist_out || (p > 0 && !x)) return fail(SCEDAR_ERR_ARG, "NULL buffer");
  SCEDAR_CUDA(cudaSetDevice(device));
  cudaStream_t s = nullptr;
  double* xd = nullptr;
  int64_t* id = nullptr;
  double* dd = nullptr;
  scedar_cells_t* c = nullptr;
  int rc = SCEDAR_OK;
  auto cleanup = [&]() {
    if (c) scedar_cells_free(c);
    if (xd) cudaFree(xd);
    if (id) cudaFree(id);
    if (dd) cudaFree(dd);
    if (s) cudaStreamDestroy(s);
  };
  HOST_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  HOST_CUDA(cudaMalloc(&xd, sizeof(double) * std::max<int64_t>(n * p, 1)));
  HOST_CUDA(cudaMalloc(&id, sizeof(int64_t) * n * k));
  HOST_CUDA(cudaMalloc(&dd, sizeof(double) * n * k));
  HOST_CUDA(cudaMemcpyAsync(xd, x, sizeof(double) * n * p, cudaMemcpyHostToDevice, s));
  rc = scedar_cells_from_dense(xd, n, p, p, metric, s, &c);
  if (rc == SCEDAR_OK) rc = scedar_knn_stripe(c, precision, k, exclude, 0, n, id, dd, s);
  if (rc != SCEDAR_OK) {
    cleanup();
    return rc;
  }
  HOST_CUDA(cudaMemcpyAsync(idx_out, id, sizeof(int64_t) * n * k, cudaMemcpyDeviceToHost, s));
  HOST_CUDA(cudaMemcpyAsync(dist_out, dd, sizeof(double) * n * k, cudaMemcpyDeviceToHost, s));
  HOST_CUDA(cudaStreamSynchronize(s));
  cleanup();
  return rc;
#undef HOST_CUDA
}

}  // extern "C"
