// Data-layout passes of the PCA producer (SURVEY.md section 8f rank 4b;
// scedar/eda/sdm.py:1313-1334 -> sklearn PCA / TruncatedSVD): column means,
// column shift, and the transposed (features x samples) working copy on which
// the FP64 Gram kernel yields the scatter matrix XtX.  All HBM-bound.
#include "common.cuh"

namespace scedar {
namespace {

constexpr int CM_CHUNKS = 64;

// deterministic two-stage column sums of the real rows: part [CM_CHUNKS, p_pad]
__global__ void __launch_bounds__(256)
pca_colsum_kernel(const double* __restrict__ xw, int64_t n, int64_t p_pad,
                  double* __restrict__ part) {
  const int64_t c = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;
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
pca_colmean_kernel(const double* __restrict__ part, int64_t n, int64_t p, int64_t p_pad,
                   double* __restrict__ mean) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p) return;
  double t = 0.0;
  for (int q = 0; q < CM_CHUNKS; ++q) t += part[(int64_t)q * p_pad + c];
  mean[c] = n > 0 ? t / (double)n : 0.0;
}

__global__ void __launch_bounds__(256)
pca_shift_kernel(double* __restrict__ xw, int64_t n, int64_t p, int64_t p_pad,
                 const double* __restrict__ shift) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = blockIdx.y;
  if (c < p && r < n) xw[r * p_pad + c] -= shift[c];
}

// out [f][i] = in [i][f] - shift[f] for i < n, f < p, zero elsewhere (32 x 32
// tiles through shared memory: coalesced on both sides)
__global__ void __launch_bounds__(256)
pca_transpose_kernel(const double* __restrict__ in, int64_t n, int64_t p, int64_t ld_in,
                     const double* __restrict__ shift, double* __restrict__ out,
                     int64_t rows_out, int64_t ld_out) {
  __shared__ double t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t i0 = (int64_t)blockIdx.x * 32, f0 = (int64_t)blockIdx.y * 32;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t i = i0 + ty + 8 * q, f = f0 + tx;
    double v = 0.0;
    if (i < n && f < p) v = in[i * ld_in + f] - (shift ? shift[f] : 0.0);
    t[ty + 8 * q][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t f = f0 + ty + 8 * q, i = i0 + tx;
    if (f < rows_out && i < ld_out) out[f * ld_out + i] = t[tx][ty + 8 * q];
  }
}

}  // namespace

int launch_col_means(const scedar_cells* c, double* mean_out, cudaStream_t s) {
  if (c->p <= 0) return SCEDAR_OK;
  Scratch part;
  SCEDAR_CHECK(part.alloc(sizeof(double) * CM_CHUNKS * c->p_pad, s));
  pca_colsum_kernel<<<dim3((unsigned)((c->p_pad + 31) / 32), CM_CHUNKS), 256, 0, s>>>(
      c->xw, c->n, c->p_pad, part.as<double>());
  pca_colmean_kernel<<<(unsigned)((c->p + 255) / 256), 256, 0, s>>>(part.as<double>(), c->n, c->p,
                                                                   c->p_pad, mean_out);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch(2);
  return SCEDAR_OK;
}

int launch_shift_cols(scedar_cells* c, const double* shift, cudaStream_t s) {
  if (c->p <= 0 || c->n <= 0) return SCEDAR_OK;
  if (c->n > 0x7fffffffLL / 2) return fail(SCEDAR_ERR_UNSUPPORTED, "too many rows");
  pca_shift_kernel<<<dim3((unsigned)((c->p + 255) / 256), (unsigned)c->n), 256, 0, s>>>(
      c->xw, c->n, c->p, c->p_pad, shift);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

int launch_transpose_cells(const scedar_cells* c, const double* shift, scedar_cells* out,
                           cudaStream_t s) {
  if (out->n != c->p || out->p != c->n)
    return fail(SCEDAR_ERR_ARG, "the transposed handle must be [n_features, n_samples]");
  // every entry of the padded output is written (zeros in the padding)
  const dim3 grid((unsigned)((out->p_pad + 31) / 32), (unsigned)((out->n_pad + 31) / 32));
  pca_transpose_kernel<<<grid, 256, 0, s>>>(c->xw, c->n, c->p, c->p_pad, shift, out->xw,
                                            out->n_pad, out->p_pad);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

}  // namespace scedar
