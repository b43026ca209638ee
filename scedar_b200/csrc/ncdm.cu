// num_correct_dist_mat (scedar/eda/sdm.py:187-222) for a user-supplied matrix,
// in place, HBM-bound: every 32x32 tile pair (lower tile + its mirror) is read
// once and written once with coalesced rows; the transposition goes through
// shared memory.  Algorithmic traffic: 16*n^2 bytes (read + write).
//
// Reference order of operations: warn if any |diag| > 1e-5; warn if the
// triangles differ (|upper - lower| <= 1e-3*|lower| fails); negatives -> 0;
// diagonal -> 0; optional clamp to upper_bound; upper triangle := lower.
#include "common.cuh"

namespace scedar {
namespace {

constexpr int TS = 32;

__device__ __forceinline__ double clean(double v, int has_ub, double ub) {
  v = v < 0.0 ? 0.0 : v;
  if (has_ub && v > ub) v = ub;
  return v;
}

__global__ void __launch_bounds__(256)
ncdm_kernel(double* __restrict__ d, int64_t ldd, int64_t n, int has_ub, double ub,
            int32_t* __restrict__ flags) {
  __shared__ double lo[TS][TS + 1];  // tile (bi, bj), bi >= bj: the lower side
  __shared__ double up[TS][TS + 1];  // tile (bj, bi): the upper side
  // enumerate lower-triangular tile pairs
  const int64_t nt = (n + TS - 1) / TS;
  int64_t bi = (int64_t)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
  while (bi * (bi + 1) / 2 > (int64_t)blockIdx.x) --bi;
  while ((bi + 1) * (bi + 2) / 2 <= (int64_t)blockIdx.x) ++bi;
  const int64_t bj = (int64_t)blockIdx.x - bi * (bi + 1) / 2;
  if (bi >= nt) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int bad = 0;

  for (int r = ty; r < TS; r += 8) {
    const int64_t gi = bi * TS + r, gj = bj * TS + tx;
    lo[r][tx] = (gi < n && gj < n) ? d[gi * ldd + gj] : 0.0;
    const int64_t ui = bj * TS + r, uj = bi * TS + tx;
    up[r][tx] = (ui < n && uj < n) ? d[ui * ldd + uj] : 0.0;
  }
  __syncthreads();

  // lower side out: element (gi, gj) with gi > gj keeps its cleaned value;
  // on a diagonal tile the part above the diagonal takes the mirrored one
  for (int r = ty; r < TS; r += 8) {
    const int64_t gi = bi * TS + r, gj = bj * TS + tx;
    if (gi >= n || gj >= n) continue;
    double out;
    if (gi > gj) {
      const double l = lo[r][tx], u = up[tx][r];  // u = d[gj][gi]
      if (!(fabs(u - l) <= 1e-3 * fabs(l)) && !(u == l) && !(isnan(u) && isnan(l)))
        bad |= SCEDAR_NCDM_ASYMMETRIC;
      out = clean(l, has_ub, ub);
    } else if (gi == gj) {
      if (!(fabs(lo[r][tx]) <= 1e-5)) bad |= SCEDAR_NCDM_BAD_DIAG;
      out = (has_ub && 0.0 > ub) ? ub : 0.0;
    } else {
      out = clean(lo[tx][r], has_ub, ub);  // diagonal tile only: mirror
    }
    d[gi * ldd + gj] = out;
  }
  if (bi != bj) {
    for (int r = ty; r < TS; r += 8) {
      const int64_t ui = bj * TS + r, uj = bi * TS + tx;
      if (ui >= n || uj >= n) continue;
      d[ui * ldd + uj] = clean(lo[tx][r], has_ub, ub);
    }
  }
  if (bad) atomicOr(flags, bad);
}

}  // namespace

int launch_ncdm(double* d, int64_t ldd, int64_t n, int has_ub, double ub, int32_t* flags,
                cudaStream_t s) {
  SCEDAR_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t), s));
  if (n == 0) return SCEDAR_OK;
  const int64_t nt = (n + TS - 1) / TS;
  const int64_t pairs = nt * (nt + 1) / 2;
  if (pairs > 0x7fffffffLL) return fail(SCEDAR_ERR_UNSUPPORTED, "matrix too large");
  ncdm_kernel<<<(unsigned)pairs, 256, 0, s>>>(d, ldd, n, has_ub, ub, flags);
  SCEDAR_CUDA(cudaGetLastError());
  count_launch();
  return SCEDAR_OK;
}

}  // namespace scedar
