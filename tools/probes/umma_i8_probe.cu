// Throughput probe for the round-2 plan: back-to-back tcgen05.mma kind::i8
// (M=128, N=256, K=32, int32 accumulators in TMEM, A/B from shared memory through SWIZZLE_128B K-major
// descriptors) issued by one thread per CTA, one CTA per SM. Operand values are irrelevant; int32 wrap-around is
// allowed. Prints POP/s (2 x MAC/s). Standalone: not linked into libfastfp_b200.so.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_probe umma_i8_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N>
__global__ void __launch_bounds__(128, 1) umma_i8_kernel(int iters, int* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  unsigned char* sA = smem;                 // 128 rows x 128 bytes (K-major, one 128B-swizzle atom wide)
  unsigned char* sB = smem + 128 * 128;     // N rows x 128 bytes
  for (int i = threadIdx.x; i < (128 + N) * 128; i += blockDim.x) smem[i] = (unsigned char)((i * 7 + 3) & 3);
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async proxy reads
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    // shared-memory matrix descriptor: start address >> 4 | LBO(16 B, unused for swizzled K-major) << 16 |
    // SBO (8 rows x 128 B = 1024 B) >> 4 << 32 | version 1 << 46 | SWIZZLE_128B (2) << 61
    auto desc = [](uint32_t addr) {
      return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
             ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    };
    const uint64_t da = desc(s32(sA)), db = desc(s32(sB));
    // instruction descriptor: C = S32 (2) at [4,6), A/B = signed 8-bit (1) at [7,10)/[10,13), K-major both,
    // N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    for (int it = 0; it < iters; ++it) {
      const uint32_t acc = tm + (uint32_t)((it & 1) * (N <= 256 ? 256 : 0));
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // four K=32 steps across the 128-byte swizzle row
        const uint32_t scale = (it > 1 || k > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(acc),
            "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(idesc), "r"(scale));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"((uint64_t)s32(&bar)));
    uint32_t ok = 0;
    long long spins = 0;
    while (!ok && spins < (1LL << 28)) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}"
                   : "=r"(ok) : "r"(s32(&bar)), "r"(0u) : "memory");
      ++spins;
    }
    if (out) out[blockIdx.x] = ok ? 1 : -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "n"(512));
}

template <int N>
int run(int iters, int sms) {
  const size_t sm = (size_t)(128 + N) * 128 + 1024;
  CK(cudaFuncSetAttribute(umma_i8_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  int* d_out = nullptr;
  CK(cudaMalloc(&d_out, sms * sizeof(int)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    umma_i8_kernel<N><<<sms, 128, sm>>>(iters, d_out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  int h0 = 0; CK(cudaMemcpy(&h0, d_out, sizeof(int), cudaMemcpyDeviceToHost));
  const double macs = (double)sms * iters * 4.0 * 128.0 * N * 32.0;
  printf("tcgen05.mma kind::i8 M=128 N=%d K=32: %d iters x 4 per CTA, %d CTAs: %.3f ms, %.1f TOP/s (completed=%d)\n", N, iters,
         sms, best, 2.0 * macs / (best * 1e-3) / 1e12, h0);
  cudaFree(d_out);
  return 0;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 20000;
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (run<256>(iters, sms)) return 1;
  if (run<128>(iters, sms)) return 1;
  return 0;
}
