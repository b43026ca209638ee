// Warp-resident sorted (distance, index) lists and the metric epilogue shared
// by the top-k kernels and the Gram kernels.
#pragma once

#include "common.cuh"

namespace scedar {

constexpr int kMaxR = 4;  // registers per lane -> up to 128 entries per row

__device__ __forceinline__ bool before(double d, int j, double td, int tj) {
  return d < td || (d == td && j < tj);
}
__device__ __forceinline__ double pos_inf() {
  return __longlong_as_double(0x7ff0000000000000LL);
}

// Sorted list of 32*R (distance, index) entries; entry e lives in register
// e/32 of lane e%32.
template <int R>
struct WarpList {
  double d[R];
  int i[R];
  double td;  // threshold = entry K-1
  int tj;

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      d[r] = pos_inf();
      i[r] = 0x7fffffff;
    }
    td = pos_inf();
    tj = 0x7fffffff;
  }
  __device__ __forceinline__ void refresh(int K) {
    const int kr = (K - 1) >> 5, kl = (K - 1) & 31;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double x = __shfl_sync(0xffffffffu, d[r], kl);
      const int y = __shfl_sync(0xffffffffu, i[r], kl);
      if (r == kr) {
        td = x;
        tj = y;
      }
    }
  }
  // warp-uniform (nd, nj); caller has checked before(nd, nj, td, tj)
  __device__ __forceinline__ void insert(double nd, int nj, int lane) {
    int pos = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
      pos += __popc(__ballot_sync(0xffffffffu, before(d[r], i[r], nd, nj)));
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
      double pd = __shfl_up_sync(0xffffffffu, d[r], 1);
      int pi = __shfl_up_sync(0xffffffffu, i[r], 1);
      if (r > 0) {
        const double cd = __shfl_sync(0xffffffffu, d[r - 1], 31);
        const int ci = __shfl_sync(0xffffffffu, i[r - 1], 31);
        if (lane == 0) {
          pd = cd;
          pi = ci;
        }
      }
      const int e = r * 32 + lane;
      if (e > pos) {
        d[r] = pd;
        i[r] = pi;
      } else if (e == pos) {
        d[r] = nd;
        i[r] = nj;
      }
    }
  }
  // offer a per-lane candidate (v, j); `ok` masks lanes without one
  __device__ __forceinline__ void offer(double v, int j, bool ok, int K, int lane) {
    unsigned m = __ballot_sync(0xffffffffu, ok && before(v, j, td, tj));
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const double nd = __shfl_sync(0xffffffffu, v, src);
      const int nj = __shfl_sync(0xffffffffu, j, src);
      if (before(nd, nj, td, tj)) {
        insert(nd, nj, lane);
        refresh(K);
      }
    }
  }
  // write the k survivors of the k+1 kept: drop the row itself when present
  // (exclude == SELF), otherwise rank 0
  __device__ __forceinline__ void emit(int K, int exclude, int self, int64_t* idx_out,
                                       double* dist_out, int lane) {
    int drop = 0;
    if (exclude == SCEDAR_EXCLUDE_SELF) {
      int found = 0x7fffffff;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int e = r * 32 + lane;
        const unsigned hit = __ballot_sync(0xffffffffu, e < K && i[r] == self);
        if (hit && found == 0x7fffffff) found = r * 32 + __ffs(hit) - 1;
      }
      if (found != 0x7fffffff) drop = found;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int e = r * 32 + lane;
      if (e < K && e != drop) {
        const int o = e < drop ? e : e - 1;
        idx_out[o] = i[r] == 0x7fffffff ? -1 : (int64_t)i[r];
        dist_out[o] = d[r];
      }
    }
  }
};


// Metric epilogue in the reference's operation order (see gram_f64.cu): P is
// the raw dot product, s_hi / s_lo the statistics of the larger / smaller
// cell index.
__device__ __forceinline__ double finish_entry(double P, double s_hi, double s_lo,
                                               bool diag, int euclid) {
  double d;
  if (euclid) {
    double t2 = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, P), s_hi), s_lo);
    t2 = t2 < 0.0 ? 0.0 : t2;
    d = __dsqrt_rn(t2);
  } else {
    d = __dsub_rn(1.0, __dmul_rn(__dmul_rn(P, s_hi), s_lo));
    d = d < 0.0 ? 0.0 : d;
  }
  return diag ? 0.0 : d;
}

}  // namespace scedar
