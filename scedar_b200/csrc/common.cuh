// Shared declarations of the scedar_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "scedar_b200.h"

namespace scedar {

// ---- tile geometry of the working copy -----------------------------------
constexpr int kRowPad = 128;  // cells padded to a multiple of the Gram tile
constexpr int kColPad = 16;   // genes padded to a multiple of the K slab

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// ---- error plumbing --------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define SCEDAR_CUDA(expr)                                                     \
  do {                                                                        \
    cudaError_t e__ = (expr);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      return ::scedar::fail(                                                  \
          e__ == cudaErrorMemoryAllocation ? SCEDAR_ERR_NOMEM                 \
                                           : SCEDAR_ERR_CUDA,                 \
          std::string(#expr) + ": " + cudaGetErrorString(e__));               \
    }                                                                         \
  } while (0)

#define SCEDAR_CHECK(call)            \
  do {                                \
    int rc__ = (call);                \
    if (rc__ != SCEDAR_OK) return rc__; \
  } while (0)

// stream-ordered scratch, released on scope exit
struct Scratch {
  void* ptr = nullptr;
  cudaStream_t stream = nullptr;
  ~Scratch() {
    if (ptr) cudaFreeAsync(ptr, stream);
  }
  int alloc(size_t bytes, cudaStream_t s);
  template <typename T>
  T* as() const { return static_cast<T*>(ptr); }
};

int num_sms();
void count_launch(int n = 1);  // kernels launched by this library (bench evidence)

}  // namespace scedar

// The prepared cells handle (opaque in the C header).
struct scedar_cells {
  int64_t n = 0, p = 0;          // logical shape
  int64_t n_pad = 0, p_pad = 0;  // padded shape of xw
  int metric = 0;
  int device = 0;
  double* xw = nullptr;    // [n_pad, p_pad] working copy (centred if correlation)
  double* stat = nullptr;  // [n_pad] sqrt(1/|x|^2) (0 for zero rows) or |x|^2
  cudaStream_t stream = nullptr;  // creation stream: the frees are ordered on it
  void* fast = nullptr;  // FastState of the tensor-core path (gram_tc.cu), lazy
  void* oz = nullptr;    // OzState of the INT8 split path (gram_oz.cu), lazy
  // optional by-product of the INT8 split launches: per output stripe, the
  // minimum of every aligned run of 32 entries of each row of D
  // ([rows of the stripe][bmin_ld]); consumed by launch_topk_blockmin
  double* bmin[16] = {nullptr};
  int bmin_count = 0;
  int64_t bmin_ld = 0;
  bool bmin_filled = false;
  CUtensorMap xmap;      // TMA view of xw: boxes of 128 rows x 16 doubles, 128B swizzle
};

namespace scedar {

// ---- kernel launchers (each in its own .cu) --------------------------------
// prep.cu
int launch_prep_dense(const double* x, int64_t n, int64_t p, int64_t ldx,
                      int metric, double* xw, int64_t n_pad, int64_t p_pad,
                      cudaStream_t s);
int launch_prep_csr(const void* indptr, int indptr_is_64,
                    const int32_t* indices, const double* data, int64_t n,
                    int64_t p, double* xw, int64_t n_pad, int64_t p_pad,
                    cudaStream_t s);

// gram_f64.cu
int launch_gram_diag(const scedar_cells* c, cudaStream_t s);
int launch_gram_diag_rows(const scedar_cells* c, int64_t row_begin, int64_t row_end,
                          cudaStream_t s);
int launch_gram_lower_rows(const scedar_cells* c, int64_t row_begin, int64_t row_end,
                           double* d_full, int64_t ldd, cudaStream_t s);
int launch_gram_stripe(const scedar_cells* c, int64_t row_begin,
                       int64_t row_end, double* d_out, int64_t ldd,
                       cudaStream_t s);
int launch_gram_stripe_p2p(const scedar_cells* c, int rank, int world,
                           const int64_t* stripe_begin, double* const* stripe_ptr,
                           int64_t ldd, cudaStream_t s);
int launch_gram_symmetric(const scedar_cells* c, double* d_out, int64_t ldd,
                          cudaStream_t s);
int launch_gram_gathered(const scedar_cells* c, const double* g, const int64_t* row_ids,
                         int64_t m, int64_t m_pad, double* d_out, int64_t ldd,
                         cudaStream_t s);
int launch_gram_raw_symmetric(const scedar_cells* c, double* out, int64_t ldd, cudaStream_t s);
int launch_gram_raw_cross(const scedar_cells* ca, const scedar_cells* cb, double* out,
                          int64_t ldd, cudaStream_t s);
// pca.cu (column means, column shift, transposed working copy)
int launch_col_means(const scedar_cells* c, double* mean_out, cudaStream_t s);
int launch_shift_cols(scedar_cells* c, const double* shift, cudaStream_t s);
int launch_transpose_cells(const scedar_cells* c, const double* shift, scedar_cells* out,
                           cudaStream_t s);
// fused Gram -> top-K; part_d/part_i: [splits, rows, K] partial lists
int gram_topk_splits(const scedar_cells* c, int64_t rows);
int launch_gram_topk(const scedar_cells* c, int64_t row_begin, int64_t row_end,
                     int K, int splits, double* part_d, int32_t* part_i,
                     cudaStream_t s);

// gram_tc.cu (tcgen05 fast path)
void fast_state_free(void* fast, cudaStream_t s);
int fast_knn_max_k();
int launch_fast_stripe(const scedar_cells* c, int64_t row_begin, int64_t row_end,
                       double* d_out, int64_t ldd, cudaStream_t s,
                       int symmetric = 0);
int launch_fast_knn(const scedar_cells* c, int k, int exclude, int64_t row_begin,
                    int64_t row_end, int64_t* idx_out, double* dist_out,
                    cudaStream_t s);
// gram_oz.cu (FP64-grade Gram as an error-free split on the INT8 tensor pipe;
// S = 6 or 7 slices)
void oz_state_free(void* oz, cudaStream_t s);
bool oz_supported(const scedar_cells* c);
int launch_oz_stripe(const scedar_cells* c, int S, int64_t row_begin, int64_t row_end,
                     double* d_out, int64_t ldd, cudaStream_t s);
int launch_oz_symmetric(const scedar_cells* c, int S, double* d_out, int64_t ldd,
                        cudaStream_t s);
int launch_oz_stripe_p2p(const scedar_cells* c, int S, int rank, int world,
                         const int64_t* stripe_begin, double* const* stripe_ptr,
                         int64_t ldd, cudaStream_t s);
int launch_oz_lower_rows(const scedar_cells* c, int S, int64_t row_begin, int64_t row_end,
                         double* d_full, int64_t ldd, cudaStream_t s);
int launch_oz_p2p_rows(const scedar_cells* c, int S, int rank, int world,
                       const int64_t* stripe_begin, double* const* stripe_ptr, int64_t ldd,
                       int64_t row_begin, int64_t row_end, int reserve_sms, cudaStream_t s);
int launch_oz_p2p_pieces(const scedar_cells* c, int S, int rank, int world,
                         const int64_t* stripe_begin, double* const* stripe_ptr, int64_t ldd,
                         const int32_t* piece_of_tile, int piece, int n_pieces, int reserve_sms,
                         cudaStream_t s);
// capi.cu: the FP64 kNN of a row range (fused when k+1 <= 32); oz_S = 6 / 7
// computes the distances with the INT8 split kernel instead of DMMA
int knn_stripe_fp64(const scedar_cells* c, int k, int exclude, int64_t row_begin,
                    int64_t row_end, int64_t* idx_out, double* dist_out,
                    cudaStream_t s, int oz_S = 0);

// topk.cu
int launch_topk_from_d(const double* d, int64_t ldd, int64_t n_cols,
                       int64_t row_begin, int64_t row_end, int k, int exclude,
                       int64_t* idx_out, double* dist_out, cudaStream_t s,
                       const int64_t* row_ids = nullptr, int64_t out_base = 0);
int launch_topk_blockmin(const double* d, int64_t ldd, int64_t n_cols, const double* bmin,
                         int64_t ldb, int64_t row_begin, int64_t row_end, int k, int exclude,
                         int64_t* idx_out, double* dist_out, cudaStream_t s);
int launch_topk_merge(const double* part_d, const int32_t* part_i, int splits,
                      int64_t rows, int K, int64_t row_begin, int k,
                      int exclude, int64_t* idx_out, double* dist_out,
                      cudaStream_t s);
int launch_col_sort(const double* d, int64_t ldd, int64_t rows, int64_t n,
                    double* sorted_out, int64_t* argsorted_out,
                    cudaStream_t s);

// consumers.cu (kNN graph, rare-sample detection, imputation)
int launch_conn_csr(const int64_t* idx, const double* dist, int64_t n, int k,
                    int64_t* indptr, int32_t* indices, double* data,
                    cudaStream_t s);
int launch_affinity(const double* w, int64_t nnz, double aff_scale, double* out,
                    cudaStream_t s);
int launch_masked_kth(const double* d, int64_t ldd, int64_t n_cols,
                      int64_t row_begin, int64_t row_end, const uint8_t* alive,
                      int rank, double* kth_out, cudaStream_t s);
int launch_rare_update(const double* kth, int64_t ldk, int64_t n, double cutoff,
                       int iter, uint8_t* alive, int32_t* last_kept,
                       unsigned long long* counters, cudaStream_t s);
int launch_masked_max(const double* v, int64_t ldv, int64_t n,
                      const uint8_t* alive, unsigned long long* max_bits,
                      cudaStream_t s);
int launch_compact_alive(const uint8_t* alive, int64_t n, int64_t* sel,
                         cudaStream_t s);
int launch_gather_rows(const double* x, int64_t ldx, int64_t p,
                       const int64_t* sel, int64_t m, double* out, int64_t ldo,
                       cudaStream_t s);
int launch_flag_to_ids(const uint8_t* flag, int64_t n, int64_t id0, int64_t* ids,
                       int64_t m, int64_t ids_cap, cudaStream_t s);
int launch_impute_iter(const double* x_curr, int64_t ldx, int64_t n, int64_t p,
                       const int64_t* knn, int k, int n_do_i, double min_present,
                       int iter, int64_t row_begin, int64_t row_end,
                       double* x_next, int64_t ldn, int32_t* idc, int64_t ldi,
                       unsigned long long* counters, cudaStream_t s);

// hclust.cu (condensed D, sub-matrix gather)
int launch_squareform_rows(const double* d, int64_t ldd, int64_t n, int64_t row_begin,
                           int64_t row_end, double* out, int32_t* flags, cudaStream_t s);
int launch_gather_block(const double* d, int64_t ldd, int64_t n_rows, int64_t n_cols,
                        const int64_t* row_sel, int64_t m, const int64_t* col_sel, int64_t c,
                        double* out, int64_t ldo, int32_t* flags, cudaStream_t s);

// ncdm.cu
int launch_ncdm(double* d, int64_t ldd, int64_t n, int has_ub, double ub,
                int32_t* flags, cudaStream_t s);

}  // namespace scedar
