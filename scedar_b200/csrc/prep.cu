// K1 -- per-cell preparation pass (HBM-bound).
//
// Builds the padded working copy xw[n_pad, p_pad] the Gram kernels read:
//   * correlation: every row centred on its mean      (scedar/eda/sdm.py:1286)
//   * cosine / euclidean: a plain copy                (sdm.py:1244-1247)
//   * CSR input: rows densified                        (sdm.py:1245)
// Pad rows / columns are zero so the contraction needs no edge handling
// (fma(0, 0, acc) == acc exactly).
//
// Algorithmic traffic: 8*n*p bytes read + 8*n_pad*p_pad bytes written.
// Dense rows are read with 128-bit loads when the row stride allows it; the
// second read of a row (after the mean is known) hits L1/L2, not HBM.
#include "common.cuh"

namespace scedar {
namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// G threads cooperate on one row; blockDim.x / G rows per block.
template <int G, bool VEC>
__global__ void __launch_bounds__(256)
prep_dense_kernel(const double* __restrict__ x, int64_t n, int64_t p,
                  int64_t ldx, int centre, double* __restrict__ xw,
                  int64_t n_pad, int64_t p_pad) {
  constexpr int ROWS = 256 / G;
  __shared__ double red[ROWS][G / 32];
  const int sub = threadIdx.x / G;   // row slot inside the block
  const int t = threadIdx.x % G;     // thread inside the row group
  const int64_t row = (int64_t)blockIdx.x * ROWS + sub;
  if (row >= n_pad) return;          // whole group exits together
  double* __restrict__ dst = xw + row * p_pad;

  if (row >= n) {                    // padding cell: all zero
    for (int64_t c = 2 * t; c < p_pad; c += 2 * G)
      *reinterpret_cast<double2*>(dst + c) = make_double2(0.0, 0.0);
    return;
  }
  const double* __restrict__ src = x + row * ldx;

  double mean = 0.0;
  if (centre) {
    // per-thread strided partial sums, then a shuffle tree: error growth like
    // pairwise summation (NumPy's mean uses pairwise blocks of 8/128)
    double s0 = 0.0, s1 = 0.0;
    if (VEC) {
      const int64_t p2 = p >> 1;
      const double2* __restrict__ s2 = reinterpret_cast<const double2*>(src);
      for (int64_t c = t; c < p2; c += G) {
        double2 v = __ldg(s2 + c);
        s0 += v.x;
        s1 += v.y;
      }
      if ((p & 1) && t == 0) s0 += __ldg(src + p - 1);
    } else {
      for (int64_t c = t; c < p; c += G) s0 += __ldg(src + c);
    }
    double s = warp_sum(s0 + s1);
    if (G > 32) {
      if ((t & 31) == 0) red[sub][t >> 5] = s;
      __syncthreads();
      s = 0.0;
#pragma unroll
      for (int w = 0; w < G / 32; ++w) s += red[sub][w];
    }
    mean = s / (double)p;
  }

  // second pass: (re)read, centre, write the padded row with 128-bit stores
  for (int64_t c = 2 * t; c < p_pad; c += 2 * G) {
    double a = 0.0, b = 0.0;
    if (VEC && c + 1 < p) {
      double2 v = __ldg(reinterpret_cast<const double2*>(src + c));
      a = v.x - mean;
      b = v.y - mean;
    } else {
      if (c < p) a = __ldg(src + c) - mean;
      if (c + 1 < p) b = __ldg(src + c + 1) - mean;
    }
    *reinterpret_cast<double2*>(dst + c) = make_double2(a, b);
  }
}

template <typename IP>
__global__ void __launch_bounds__(256)
prep_csr_scatter_kernel(const IP* __restrict__ indptr,
                        const int32_t* __restrict__ indices,
                        const double* __restrict__ data, int64_t n, int64_t p,
                        double* __restrict__ xw, int64_t p_pad) {
  // one warp per cell; xw was zero-filled by a memset on the same stream
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const int64_t b = (int64_t)indptr[row], e = (int64_t)indptr[row + 1];
  double* __restrict__ dst = xw + row * p_pad;
  for (int64_t t = b + lane; t < e; t += 32) {
    const int32_t col = indices[t];
    if (col >= 0 && col < p) atomicAdd(dst + col, data[t]);  // sums duplicates
  }
}

}  // namespace

int launch_prep_dense(const double* x, int64_t n, int64_t p, int64_t ldx,
                      int metric, double* xw, int64_t n_pad, int64_t p_pad,
                      cudaStream_t s) {
  const int centre = metric == SCEDAR_METRIC_CORRELATION;
  const bool vec = (ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (p_pad <= 512) {
    const unsigned grid = (unsigned)((n_pad + 7) / 8);
    if (vec)
      prep_dense_kernel<32, true><<<grid, 256, 0, s>>>(x, n, p, ldx, centre, xw, n_pad, p_pad);
    else
      prep_dense_kernel<32, false><<<grid, 256, 0, s>>>(x, n, p, ldx, centre, xw, n_pad, p_pad);
  } else {
    const unsigned grid = (unsigned)n_pad;
    if (vec)
      prep_dense_kernel<256, true><<<grid, 256, 0, s>>>(x, n, p, ldx, centre, xw, n_pad, p_pad);
    else
      prep_dense_kernel<256, false><<<grid, 256, 0, s>>>(x, n, p, ldx, centre, xw, n_pad, p_pad);
  }
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int launch_prep_csr(const void* indptr, int indptr_is_64,
                    const int32_t* indices, const double* data, int64_t n,
                    int64_t p, double* xw, int64_t n_pad, int64_t p_pad,
                    cudaStream_t s) {
  SCEDAR_CUDA(cudaMemsetAsync(xw, 0, sizeof(double) * n_pad * p_pad, s));
  if (n == 0) return SCEDAR_OK;
  const unsigned grid = (unsigned)((n * 32 + 255) / 256);
  if (indptr_is_64)
    prep_csr_scatter_kernel<int64_t><<<grid, 256, 0, s>>>(
        static_cast<const int64_t*>(indptr), indices, data, n, p, xw, p_pad);
  else
    prep_csr_scatter_kernel<int32_t><<<grid, 256, 0, s>>>(
        static_cast<const int32_t*>(indptr), indices, data, n, p, xw, p_pad);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

}  // namespace scedar
