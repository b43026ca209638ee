// Data side of the clustering consumers of D.
// Both kernels are pure data movement, HBM-bound, and keep D on the device
// between the Gram kernel and the (sequential, CPU) linkage.
//
//   squareform_rows   rows [row_begin, row_end) of the strict upper triangle of
//                     a symmetric n x n D, concatenated: the condensed vector
//                     scipy.spatial.distance.squareform(dmat) hands to
//                     sch.linkage in HClustTree.hclust_linkage
//                     (scedar/eda/sdm.py:1666-1668).  Row i owns the contiguous
//                     run [base(i), base(i+1)) of the vector, base(i) =
//                     i(2n-i-1)/2, so a row stripe of D yields one contiguous
//                     segment and D never has to be whole anywhere.  The grid
//                     walks the OUTPUT front to back (aligned, contiguous
//                     stores; reads of a few adjacent rows), which keeps the
//                     DRAM access pattern that of a plain copy: 5.86 TB/s at
//                     n = 40,000 (0.89 of the measured copy peak) against
//                     4.1-5.6 TB/s for one block per row.
//                     Algorithmic bytes: 8 read + 8 written per pair i < j,
//                     i.e. 8n(n-1) for the whole matrix -- half of what a host
//                     copy of D alone costs.
//   gather_block      out[r, c] = d[row_sel[r], col_sel[c]]: the sub-matrix
//                     slicing d[a][:, b] of SampleDistanceMatrix.ind_x
//                     (sdm.py:684-692) and of MIRAC's cluster encoders
//                     (scedar/cluster/mirac.py:316-331).
#include <algorithm>

#include "common.cuh"

namespace scedar {
namespace {

__device__ __forceinline__ int64_t sqf_base(int64_t i, int64_t n) {
  return i * (2 * n - i - 1) / 2;
}

// Output-linear form: the blocks walk the condensed vector front to back, one
// item of 256*U consecutive entries at a time, so at any moment the grid
// writes one contiguous window of the output and reads a few adjacent rows of
// D -- the access pattern of a plain copy.
template <int U>
__global__ void __launch_bounds__(256)
squareform_linear_kernel(const double* __restrict__ d, int64_t ldd, int64_t n, int64_t row_begin,
                         int64_t row_end, double* __restrict__ out,
                         int32_t* __restrict__ flags) {
  const int64_t base0 = sqf_base(row_begin, n);
  const int64_t total = sqf_base(row_end, n) - base0;
  constexpr int64_t C = 256 * U;
  int bad = 0;
  for (int64_t t0 = (int64_t)blockIdx.x * C; t0 < total; t0 += (int64_t)gridDim.x * C) {
    const int64_t k0 = t0 + threadIdx.x;
    double v[U];
    if (k0 < total) {
      // row of the first entry: largest i with base(i) <= K (exact in double:
      // every operand is below 2^53), then fixed up against rounding of sqrt
      const int64_t K = k0 + base0;
      const double b = (double)(2 * n - 1);
      int64_t i = (int64_t)((b - sqrt(b * b - 8.0 * (double)K)) * 0.5);
      i = i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
      while (sqf_base(i, n) > K) --i;
      while (sqf_base(i + 1, n) <= K) ++i;
      int64_t rs = sqf_base(i, n);  // first condensed index of row i
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t Ku = K + u * 256;
        if (k0 + u * 256 < total) {
          while (Ku >= rs + (n - 1 - i)) {
            rs += n - 1 - i;
            ++i;
          }
          v[u] = d[(i - row_begin) * ldd + (Ku - rs + i + 1)];
        } else {
          v[u] = 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k0 + u * 256 < total) {
          if (v[u] != v[u]) bad |= SCEDAR_SQF_NAN;
          out[k0 + u * 256] = v[u];
        }
      }
    }
  }
  for (int64_t i = row_begin + (int64_t)blockIdx.x * 256 + threadIdx.x; i < row_end;
       i += (int64_t)gridDim.x * 256)
    if (!(d[(i - row_begin) * ldd + i] == 0.0)) bad |= SCEDAR_SQF_BAD_DIAG;
  if (bad) atomicOr(flags, bad);
}

constexpr int GB_COLS = 1024;  // columns of one block's chunk

__global__ void __launch_bounds__(256)
gather_block_kernel(const double* __restrict__ d, int64_t ldd, int64_t n_rows, int64_t n_cols,
                    const int64_t* __restrict__ row_sel, int64_t m,
                    const int64_t* __restrict__ col_sel, int64_t c, double* __restrict__ out,
                    int64_t ldo, int32_t* __restrict__ flags) {
  int bad = 0;
  for (int64_t r = blockIdx.x; r < m; r += gridDim.x) {
    const int64_t gr = row_sel[r];
    const bool row_ok = gr >= 0 && gr < n_rows;
    if (!row_ok) bad = 1;
    const double* __restrict__ src = d + (row_ok ? gr : 0) * ldd;
    for (int64_t q0 = (int64_t)blockIdx.y * GB_COLS; q0 < c; q0 += (int64_t)gridDim.y * GB_COLS) {
      const int64_t q1 = q0 + GB_COLS < c ? q0 + GB_COLS : c;
      for (int64_t q = q0 + threadIdx.x; q < q1; q += 256) {
        const int64_t gc = col_sel[q];
        const bool ok = row_ok && gc >= 0 && gc < n_cols;
        if (!ok) bad = 1;
        // an out-of-range selection is reported through the flag; its slot
        // gets a NaN so it cannot pass for a distance
        out[r * ldo + q] = ok ? src[gc] : __longlong_as_double(0x7ff8000000000000LL);
      }
    }
  }
  if (bad) atomicOr(flags, SCEDAR_GATHER_BAD_INDEX);
}

}  // namespace

int launch_squareform_rows(const double* d, int64_t ldd, int64_t n, int64_t row_begin,
                           int64_t row_end, double* out, int32_t* flags, cudaStream_t s) {
  const int64_t rows = row_end - row_begin;
  if (rows <= 0) return SCEDAR_OK;
  // 256*4 entries per item; two waves of blocks (measured: 2.19 ms at 16 blocks
  // per SM against 2.48 ms at 8 for n = 40,000)
  const int64_t total = row_end * (2 * n - row_end - 1) / 2 - row_begin * (2 * n - row_begin - 1) / 2;
  const int64_t items = std::max<int64_t>(1, (total + 1023) / 1024);
  const int64_t grid = std::min<int64_t>(items, (int64_t)num_sms() * 16);
  squareform_linear_kernel<4><<<(unsigned)grid, 256, 0, s>>>(d, ldd, n, row_begin, row_end, out,
                                                             flags);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int launch_gather_block(const double* d, int64_t ldd, int64_t n_rows, int64_t n_cols,
                        const int64_t* row_sel, int64_t m, const int64_t* col_sel, int64_t c,
                        double* out, int64_t ldo, int32_t* flags, cudaStream_t s) {
  if (m <= 0 || c <= 0) return SCEDAR_OK;
  const int64_t chunks = (c + GB_COLS - 1) / GB_COLS;
  const int64_t want = (int64_t)num_sms() * 8;
  const int64_t gx = std::min<int64_t>(m, want);
  const int64_t gy = std::max<int64_t>(1, std::min<int64_t>(chunks, want / gx));
  gather_block_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, s>>>(
      d, ldd, n_rows, n_cols, row_sel, m, col_sel, c, out, ldo, flags);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

}  // namespace scedar
