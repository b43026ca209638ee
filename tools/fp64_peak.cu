// Measures the float64 peak of this GPU two ways (register-resident loops, no
// memory traffic): plain DFMA and DMMA (mma.sync m8n8k4 f64).  The larger of
// the two is the roofline denominator for the FP64 Gram kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cuda_runtime.h>
#include <cstdio>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile(
          "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
          : "+d"(c[i][0]), "+d"(c[i][1])
          : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// outer-product shaped: MFxNF accumulator fragments from MF A- and NF B-operands
// (the register pattern of a real Gram / GEMM warp tile)
template <int MF, int NF>
__global__ void dmma_outer_kernel(double* out, int iters, double a0, double b0) {
  double c[MF][NF][2], a[MF], b[NF];
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    a[i] = a0 + 1e-9 * (threadIdx.x + i);
#pragma unroll
    for (int j = 0; j < NF; ++j) c[i][j][0] = c[i][j][1] = threadIdx.x + i + j;
  }
#pragma unroll
  for (int j = 0; j < NF; ++j) b[j] = b0 + 1e-9 * (threadIdx.x + 7 * j);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
      for (int j = 0; j < NF; ++j)
        asm volatile(
            "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
            : "+d"(c[i][j][0]), "+d"(c[i][j][1])
            : "d"(a[i]), "d"(b[j]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the same with the operands re-loaded from shared memory every k-step, as a
// real kernel must: VEC=1 -> one LDS.64 per fragment, VEC=2 -> one LDS.128 per
// two k-steps
template <int VEC>
__global__ void dmma_lds_kernel(double* out, int iters) {
  constexpr int MF = 8, NF = 4;
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 12 * 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warp = threadIdx.x >> 5;
  double c[MF][NF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
    const double* base = sm + ((it + warp) & 7) * 1024;
    if (VEC == 1) {
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        double a[MF], b[NF];
#pragma unroll
        for (int i = 0; i < MF; ++i) a[i] = base[(i * 8 + g) * 12 + kk * 4 + t];
#pragma unroll
        for (int j = 0; j < NF; ++j) b[j] = base[768 + (j * 8 + g) * 4 + t];
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
          for (int j = 0; j < NF; ++j)
            asm volatile(
                "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(c[i][j][0]), "+d"(c[i][j][1])
                : "d"(a[i]), "d"(b[j]));
      }
    } else {
      double2 a[MF], b[NF];
#pragma unroll
      for (int i = 0; i < MF; ++i)
        a[i] = *reinterpret_cast<const double2*>(base + (i * 8 + g) * 12 + 2 * t);
#pragma unroll
      for (int j = 0; j < NF; ++j)
        b[j] = *reinterpret_cast<const double2*>(base + 768 + (j * 8 + g) * 8 + 2 * t);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
          for (int j = 0; j < NF; ++j)
            asm volatile(
                "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(c[i][j][0]), "+d"(c[i][j][1])
                : "d"(kk ? a[i].y : a[i].x), "d"(kk ? b[j].y : b[j].x));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int VEC>
void run_lds(double* out, int sms, const char* name) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms, threads = 256, iters = 2048, smem = 12 * 1024 * 8;
  cudaFuncSetAttribute(dmma_lds_kernel<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    dmma_lds_kernel<VEC><<<blocks, threads, smem>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  double flops = 2.0 * 256 * 32 * 2 * iters * (double)blocks * (threads / 32);
  printf("{\"probe\": \"%s\", \"ms\": %.3f, \"tflops\": %.2f}\n", name, best,
         flops / best * 1e-9);
}

template <int MF, int NF>
void run_outer(double* out, int sms, int threads, int blocks_per_sm, const char* name) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * blocks_per_sm, iters = 2048;
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    dmma_outer_kernel<MF, NF><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  double flops = 2.0 * 256 * MF * NF * iters * (double)blocks * (threads / 32);
  printf("{\"probe\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", name,
         threads / 32 * blocks_per_sm, best, flops / best * 1e-9);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int which = 0; which < 2; ++which) {
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      else dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 2 && ms < best) best = ms;
    }
    double flops = which == 0 ? 2.0 * 16 * iters * (double)blocks * threads
                              : 2.0 * 256 * 16 * iters * (double)blocks * (threads / 32);
    printf("{\"probe\": \"%s\", \"ms\": %.3f, \"tflops\": %.2f, \"sms\": %d}\n",
           which == 0 ? "dfma" : "dmma_m8n8k4", best, flops / best * 1e-9, sms);
  }
  run_outer<4, 4>(out, sms, 256, 2, "dmma_outer_4x4_16warps");
  run_outer<8, 4>(out, sms, 256, 1, "dmma_outer_8x4_8warps");
  run_outer<8, 4>(out, sms, 128, 1, "dmma_outer_8x4_4warps");
  run_outer<4, 4>(out, sms, 256, 8, "dmma_outer_4x4_64warps");
  run_lds<1>(out, sms, "dmma_8x4_8warps_lds64_per_kstep");
  run_lds<2>(out, sms, "dmma_8x4_8warps_lds128_per_2ksteps");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
