// How fast can ONE SM pull operand tiles from L2 with TMA, as a function of the
// box shape?  gram_oz.cu measured the same ~34 bytes/cycle/SM for two different
// ring geometries (S = 6 and S = 7) whatever the number of CTAs running, i.e. a
// per-SM limit of the load path, with 64-byte box rows (64B swizzle).  This
// probe streams boxes of `inner` bytes x `rows` rows through a ring of stages
// (a consumer thread releases each stage as soon as it is full) and reports
// bytes per SM cycle.  Build: make -C tools tma_probe    Run: tools/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../scedar_b200/csrc/ptx_sm100.cuh"

using namespace scedar;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static int encode_u8(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int box_rows,
                     int box_cols, CUtensorMapSwizzle sw) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<encode_fn>(fn)(
      m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return 1;
  }
  return 0;
}

// each stage = `boxes` boxes of box_bytes; `stages` stages; `iters` stage loads
__global__ void __launch_bounds__(64, 1)
stream_kernel(const __grid_constant__ CUtensorMap map, int box_bytes, int boxes, int stages,
              int iters, int inner, int rows, int64_t mat_rows, int64_t mat_cols, int span_rows,
              unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t stage_bytes = (uint32_t)box_bytes * boxes;
  const uint32_t bars = base + stage_bytes * stages;
  const uint32_t full0 = bars, empty0 = bars + 256;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const int kcols = (int)(mat_cols / inner);
  if (threadIdx.x == 0) {
    // producer: CTA b walks boxes of its own row span, all K columns
    const long long t0 = clock64();
    // CTAs start at different K columns; plain 32-bit counters (a 64-bit division
    // per box would make the issuing thread the bottleneck)
    int kc = (int)((blockIdx.x * 37u) % (unsigned)kcols), rb = 0;
    const int nrb = span_rows / rows;
    const int row_base = (int)(((int64_t)blockIdx.x * span_rows) % (mat_rows - span_rows));
    for (int it = 0; it < iters; ++it) {
      const uint32_t s = it % stages, ph = (it / stages) & 1;
      mbar_wait(empty0 + 8 * s, ph ^ 1);
      mbar_expect_tx(full0 + 8 * s, stage_bytes);
      for (int b = 0; b < boxes; ++b) {
        tma_load_2d(base + s * stage_bytes + b * box_bytes, &map, kc * inner, row_base + rb * rows,
                    full0 + 8 * s);
        if (++kc == kcols) {
          kc = 0;
          if (++rb == nrb) rb = 0;
        }
      }
    }
    // wait for the last stages to land
    for (int it = iters; it < iters + stages; ++it) {
      const uint32_t s = it % stages, ph = (it / stages) & 1;
      mbar_wait(empty0 + 8 * s, ph ^ 1);
    }
    out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < iters; ++it) {
      const uint32_t s = it % stages, ph = (it / stages) & 1;
      mbar_wait(full0 + 8 * s, ph);
      mbar_arrive(empty0 + 8 * s);
    }
  }
}

static void run(const uint8_t* d, int64_t mat_rows, int64_t mat_cols, int inner, int rows, int boxes,
                int stages, int grid, int span_rows, const char* note) {
  CUtensorMap map;
  const CUtensorMapSwizzle sw = inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : inner == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : CU_TENSOR_MAP_SWIZZLE_32B;
  if (encode_u8(&map, d, mat_rows, mat_cols, rows, inner, sw)) return;
  const int box_bytes = inner * rows;
  const int smem = 1024 + box_bytes * boxes * stages + 512;
  if (smem > 227 * 1024) {
    printf("skip %s: smem %d\n", note, smem);
    return;
  }
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  unsigned long long* out;
  CK(cudaMalloc(&out, sizeof(unsigned long long) * grid));
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    stream_kernel<<<grid, 64, smem, 0>>>(map, box_bytes, boxes, stages, iters, inner, rows,
                                         mat_rows, mat_cols, span_rows, out);
    CK(cudaDeviceSynchronize());
  }
  std::vector<unsigned long long> h(grid);
  CK(cudaMemcpy(h.data(), out, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end());
  const double bytes = (double)box_bytes * boxes * iters;
  printf("inner %3d B x %3d rows, %2d boxes/stage (%3d KB), %d stages, grid %3d, %-18s: "
         "%.1f B/cycle/SM (median; slowest SM %.1f)\n",
         inner, rows, boxes, box_bytes * boxes / 1024, stages, grid, note,
         bytes / (double)h[grid / 2], bytes / (double)h[grid - 1]);
  cudaFree(out);
}

int main() {
  CK(cudaSetDevice(0));
  // a slice-plane-like matrix: 7 x 100,096 rows x 2,048 bytes (1.4 GB)
  const int64_t mat_rows = 7LL * 100096, mat_cols = 2048;
  uint8_t* d;
  CK(cudaMalloc(&d, (size_t)mat_rows * mat_cols));
  CK(cudaMemset(d, 1, (size_t)mat_rows * mat_cols));
  // L2-resident: every CTA re-reads a span of 256 rows (512 KB per CTA, 76 MB in all)
  for (int grid : {148, 37}) {
    run(d, mat_rows, mat_cols, 64, 128, 7, 3, grid, 256, "L2 hits");
    run(d, mat_rows, mat_cols, 64, 128, 14, 2, grid, 256, "L2 hits");
    run(d, mat_rows, mat_cols, 128, 128, 7, 2, grid, 256, "L2 hits");
    run(d, mat_rows, mat_cols, 128, 128, 4, 3, grid, 256, "L2 hits");
    run(d, mat_rows, mat_cols, 128, 64, 7, 3, grid, 256, "L2 hits");
    run(d, mat_rows, mat_cols, 128, 128, 1, 12, grid, 256, "L2 hits");
    run(d, mat_rows, mat_cols, 64, 128, 1, 24, grid, 256, "L2 hits");
    run(d, mat_rows, mat_cols, 32, 128, 7, 6, grid, 256, "L2 hits");
  }
  // streaming from HBM: spans of 4,096 rows per CTA (8 MB per CTA, 1.2 GB in all)
  run(d, mat_rows, mat_cols, 64, 128, 7, 3, 148, 4096, "HBM stream");
  run(d, mat_rows, mat_cols, 128, 128, 7, 2, 148, 4096, "HBM stream");
  cudaFree(d);
  return 0;
}
