// fp64 pipe peak: the measured denominator for the sweep kernel's fp64 roofline
// (MEASURED_PEAKS.json carries only HBM and bf16 figures). kind 0 = DFMA, kind 1 = DMMA m8n8k4.
#include "ffp_internal.cuh"

namespace ffp {

__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double seed, double* sink) {
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = seed + k + threadIdx.x * 1e-3;
  const double x = 1.0000001, y = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = fma(a[k], x, y);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  if (s == 12345.678) sink[0] = s;
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double seed, double* sink) {
  double c[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = seed + k;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[k][0]), "+d"(c[k][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
  if (s == 12345.678) sink[0] = s;
}

// kind 2: both instruction streams in one loop. If DMMA had its own pipe the combined rate would
// approach the sum of the two peaks; if it shares the fp64 pipe it stays at one peak.
__global__ void __launch_bounds__(256) mixed_peak_kernel(int iters, double seed, double* sink) {
  double c[4][2], a[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) c[k][0] = c[k][1] = seed + k;
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = seed + k + threadIdx.x * 1e-3;
  const double x = 1.0000001, y = 1e-9, am = 1.0 + threadIdx.x * 1e-9, bm = 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[k][0]), "+d"(c[k][1])
                   : "d"(am), "d"(bm));
#pragma unroll
      for (int j = 0; j < 4; ++j) a[4 * k + j] = fma(a[4 * k + j], x, y);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1];
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  if (s == 12345.678) sink[0] = s;
}

// kind 3: register-only 9x8 outer-product accumulation (the contraction's instruction mix with no
// memory operations); kind 4: the same with the operands re-read from shared memory every step.
template <bool LDS_>
__global__ void __launch_bounds__(128, 2) outer_peak_kernel(int iters, double seed, double* sink) {
  __shared__ double sh[2][32];
  if (threadIdx.x < 64) sh[threadIdx.x >> 5][threadIdx.x & 31] = seed * 1e-3 + threadIdx.x * 1e-6;
  __syncthreads();
  double acc[9][8];
#pragma unroll
  for (int r = 0; r < 9; ++r)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[r][q] = seed + r + q;
  double a[2][9], b[2][8];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
#pragma unroll
    for (int r = 0; r < 9; ++r) a[k][r] = 1.0 + 1e-9 * (r + k + (threadIdx.x & 3));
#pragma unroll
    for (int q = 0; q < 8; ++q) b[k][q] = 1e-9 * (q + k + 1);
  }
  const int lm = (threadIdx.x >> 3) & 3, ln = threadIdx.x & 7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (LDS_) {
        const volatile double* va = &sh[k][0];
#pragma unroll
        for (int r = 0; r < 9; ++r) a[k][r] = va[(lm + 4 * r) & 31];
#pragma unroll
        for (int q = 0; q < 8; ++q) b[k][q] = va[(ln + 8 * (q >> 1) + (q & 1)) & 31];
      }
#pragma unroll
      for (int r = 0; r < 9; ++r)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[r][q] = fma(a[k][r], b[k][q], acc[r][q]);
    }
  }
  double s = 0;
#pragma unroll
  for (int r = 0; r < 9; ++r)
#pragma unroll
    for (int q = 0; q < 8; ++q) s += acc[r][q];
  if (s == 12345.678) sink[0] = s;
}

int run_fp64_peak(int kind, int iters, double* tflops, double* ms_out) {
  int dev = 0, sms = 0;
  FFP_CUDA(cudaGetDevice(&dev));
  FFP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* sink = nullptr;
  FFP_CUDA(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1;
  FFP_CUDA(cudaEventCreate(&e0));
  FFP_CUDA(cudaEventCreate(&e1));
  const int grid = sms * 8;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    FFP_CUDA(cudaEventRecord(e0));
    if (kind == 0) dfma_peak_kernel<<<grid, 256>>>(iters, 1.0, sink);
    else if (kind == 1) dmma_peak_kernel<<<grid, 256>>>(iters, 1.0, sink);
    else if (kind == 2) mixed_peak_kernel<<<grid, 256>>>(iters, 1.0, sink);
    else if (kind == 3) outer_peak_kernel<false><<<sms * 2, 128>>>(iters, 1.0, sink);
    else outer_peak_kernel<true><<<sms * 2, 128>>>(iters, 1.0, sink);
    FFP_CUDA(cudaEventRecord(e1));
    FFP_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    FFP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  g_launches += 4;
  FFP_CUDA(cudaGetLastError());
  // DFMA: 16 fma/thread/iter; DMMA: 8 mma/warp/iter, 8*8*4 fma each
  const double fma_count = kind == 0   ? (double)grid * 256 * 16.0 * iters
                           : kind == 1 ? (double)grid * 8 * 8.0 * 256.0 * iters
                           : kind == 2 ? (double)grid * (256 * 16.0 + 8 * 4.0 * 256.0) * iters
                                       : (double)sms * 2 * 128 * 144.0 * iters;
  *tflops = 2.0 * fma_count / (best * 1e-3) / 1e12;
  *ms_out = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return 0;
}

}  // namespace ffp
