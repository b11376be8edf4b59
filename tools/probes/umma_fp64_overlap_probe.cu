// Round-2 probe: do fp64 FMAs and tcgen05.mma kind::i8 overlap on one SM? The tensor sweep needs the producers' fp64
// work (sincos, digit split) to run NEXT to the MMAs. One thread issues the sweep's 28-product stage back to back
// (M=128, N=64, K=32, SWIZZLE_32B planes) while NW other warps run independent DFMA chains (ILP 8); both rates are
// measured alone and together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_fp64_overlap_probe umma_fp64_overlap_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
constexpr int NP = 7, KT = 32, N = 64;
constexpr int A_PLANE = 128 * KT, B_PLANE = N * KT;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
}

// mode bit 0: MMAs on, bit 1: side warps on, bits 2..: kind of side work
__global__ void __launch_bounds__(32 * 18, 1) overlap_kernel(int mode, int iters, int nw, long long* cyc, double* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  __shared__ volatile int stop;
  unsigned char* sA = smem;
  unsigned char* sB = smem + NP * A_PLANE;
  for (int i = threadIdx.x; i < NP * (A_PLANE + B_PLANE); i += blockDim.x) smem[i] = (unsigned char)((i * 7 + 3) & 3);
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    if (mode & 1) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
      for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < NP; ++i)
#pragma unroll
          for (int j = 0; j < NP - i; ++j) {
            const uint32_t acc = (it > 0 || i > 0) ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm + (uint32_t)((i + j) * N)),
                         "l"(desc(s32(sA + i * A_PLANE))), "l"(desc(s32(sB + j * B_PLANE))), "r"(idesc), "r"(acc));
          }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"((uint64_t)s32(&bar)));
      uint32_t ok = 0;
      long long spins = 0;
      while (!ok && spins < (1LL << 28)) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(&bar)), "r"(0u) : "memory");
        ++spins;
      }
    } else {
      while (clock64() - t0 < 3000000LL) {}
    }
    const long long t1 = clock64();
    stop = 1;
    if (blockIdx.x == 0) cyc[0] = t1 - t0;
  } else if (warp >= 2 && warp < 2 + nw && (mode & 2)) {
    long long cnt = 0;
    const long long t0 = clock64();
    const int kind = mode >> 2;  // 0 DFMA, 1 IMAD (data-dependent multiplier), 2 64x64 -> high 64 multiply, 3 FFMA, 4 int -> double
    if (kind == 1) {
      unsigned a[8];
      for (int q = 0; q < 8; ++q) a[q] = threadIdx.x * 2654435761u + q;
      while (!stop) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
          for (int q = 0; q < 8; ++q) a[q] = a[q] * (a[(q + 1) & 7] | 1u) + 12345u;
        cnt += 128;
      }
      if (a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7] == 12345u) sink[0] = 1.0;
    } else if (kind == 2) {
      unsigned long long a[4];
      for (int q = 0; q < 4; ++q) a[q] = 0x9e3779b97f4a7c15ULL * (threadIdx.x + q + 1);
      while (!stop) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) a[q] = __umul64hi(a[q], a[(q + 1) & 3] | 0x8000000000000001ULL) + 0x1234567ULL;
        cnt += 64;
      }
      if (a[0] + a[1] + a[2] + a[3] == 12345ULL) sink[0] = 1.0;
    } else if (kind == 3) {
      float a[8];
      for (int q = 0; q < 8; ++q) a[q] = 1.0f + 1e-6f * (threadIdx.x + q);
      while (!stop) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
          for (int q = 0; q < 8; ++q) a[q] = fmaf(a[q], 1.0000001f, 1e-7f);
        cnt += 128;
      }
      if (a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7] == 12345.678f) sink[0] = a[0];
    } else if (kind == 4) {
      int a[8];
      double acc = 0.0;
      for (int q = 0; q < 8; ++q) a[q] = threadIdx.x + q;
      while (!stop) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
          for (int q = 0; q < 8; ++q) { a[q] = a[q] * 1664525 + (int)__double_as_longlong((double)a[q]); }
        cnt += 128;
      }
      if (a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7] == 12345) sink[0] = acc;
    } else {
      double a[8];
      for (int q = 0; q < 8; ++q) a[q] = 1.0 + 1e-9 * (threadIdx.x + q);
      while (!stop) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
          for (int q = 0; q < 8; ++q) a[q] = fma(a[q], 1.0000001, 1e-9);
        cnt += 128;
      }
      if (a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7] == 12345.678) sink[0] = a[0];
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { cyc[2 + warp] = cnt; cyc[32 + warp] = t1 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "n"(512));
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 3000;
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t sm = (size_t)NP * (A_PLANE + B_PLANE) + 1024;
  CK(cudaFuncSetAttribute(overlap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  long long* d_cyc; double* d_sink;
  CK(cudaMalloc(&d_cyc, 64 * 8)); CK(cudaMalloc(&d_sink, 8));
  const char* names[] = {"DFMA", "IMAD", "MUL64HI", "FFMA", "I2F.F64"};
  for (int nw : {8, 16}) {
    for (int kind = 0; kind < 5; ++kind)
    for (int mode : {2 | (kind << 2), 3 | (kind << 2)}) {
      CK(cudaMemset(d_cyc, 0, 64 * 8));
      overlap_kernel<<<sms, 32 * 18, sm>>>(mode, iters, nw, d_cyc, d_sink);
      CK(cudaDeviceSynchronize());
      long long h[64];
      CK(cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
      long long ops = 0, wt = 0;
      for (int w = 0; w < nw; ++w) { ops += h[4 + w]; wt = h[34 + w] > wt ? h[34 + w] : wt; }
      printf("%2d side warps, MMAs %s, side work %-8s: ", nw, (mode & 1) ? "on " : "off", names[kind]);
      if (mode & 1) printf("%8.1f cycles per 28-MMA stage; ", (double)h[0] / iters);
      if (mode & 2) printf("%.2f side warp-instructions per cycle per SM", (double)ops / (double)wt);
      printf("\n");
    }
  }
  return 0;
}
