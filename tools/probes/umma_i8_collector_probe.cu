// Round-2 probe: does the A-operand collector of tcgen05.mma (.collector::a::fill / ::use / ::lastuse) remove the
// shared-memory re-reads of the G digit planes? The sweep's stage issues plane i of G against planes j = 0..6-i of
// [s c] (28 MMAs, M=128, N=64, K=32, kind::i8); without the collector every MMA reads 4 KB (A) + 2 KB (B) from shared
// memory = 48 cycles at 128 B/clk against 32 cycles of tensor time. With i-major order the A plane can stay in the
// collector for its 7-i uses. Prints cycles per 28-MMA stage for both forms and compares the accumulators bit for bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_collector_probe umma_i8_collector_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0: plain, 1: fill / use / lastuse, 2: fill / use (never released explicitly)
__device__ __forceinline__ void mma(uint32_t acc, uint64_t da, uint64_t db, uint32_t idesc, uint32_t scale, int pos, int last) {
  if (MODE == 0 || (pos == 0 && last == 0)) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(acc), "l"(da), "l"(db), "r"(idesc), "r"(scale));
  } else if (pos == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(acc), "l"(da), "l"(db), "r"(idesc), "r"(scale));
  } else if (pos < last || MODE == 2) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::use [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(acc), "l"(da), "l"(db), "r"(idesc), "r"(scale));
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(acc), "l"(da), "l"(db), "r"(idesc), "r"(scale));
  }
}

template <int MODE, int NSTW_MAX = 16>
__global__ void __launch_bounds__(32 * 22, 1) coll_kernel(int iters, int nstw, long long* cyc, unsigned long long* sum) {
  constexpr int N = 64, NP = 7, ROWS = 96, KB = 32, A_PLANE = ROWS * KB, B_PLANE = N * KB;
  constexpr int A_BYTES = (NP - 1) * A_PLANE + 128 * KB, B_BYTES = NP * B_PLANE;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  __shared__ volatile int stop;
  unsigned char* sA = smem;
  unsigned char* sB = smem + ((A_BYTES + 1023) / 1024) * 1024;
  unsigned char* sW = sB + ((B_BYTES + 1023) / 1024) * 1024;
  for (int i = threadIdx.x; i < A_BYTES + B_BYTES + 2048; i += blockDim.x)
    smem[i] = (unsigned char)(signed char)(((i * 2654435761u) >> 13) % 7 - 3);   // signed digits -3..3
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop = 0;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 32 * 4) {  // warp 4 lane 0 issues
    auto desc = [](uint32_t addr) {
      return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) |
             ((uint64_t)6 << 61);
    };
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < NP; ++i)
#pragma unroll
        for (int j = 0; j < NP - i; ++j)
          mma<MODE>(tm + (uint32_t)((i + j) * N), desc(s32(sA + i * A_PLANE)), desc(s32(sB + j * B_PLANE)), idesc,
                    (it > 0 || i > 0) ? 1u : 0u, j, NP - 1 - i);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"((uint64_t)s32(&bar)));
    uint32_t ok = 0;
    long long spins = 0;
    while (!ok && spins < (1LL << 28)) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}"
                   : "=r"(ok) : "r"(s32(&bar)), "r"(0u) : "memory");
      ++spins;
    }
    const long long t1 = clock64();
    stop = 1;
    if (blockIdx.x == 0) cyc[0] = t1 - t0;
  } else if (warp >= 6 && warp < 6 + nstw) {
    uint32_t* w = reinterpret_cast<uint32_t*>(sW) + (warp - 6) * 32 + lane;
    uint32_t v = lane;
    while (!stop) {
#pragma unroll
      for (int q = 0; q < 14; ++q) { w[(q & 3) * 1024] = v; v = v * 3 + 1; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  if (warp < 4) {  // checksum of the 7 accumulators (lanes < ROWS only: the rest multiplied the overlapped plane tails)
    unsigned long long h = 0;
    for (int c = 0; c < NP * N; c += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tm + ((uint32_t)(32 * warp) << 16) + (uint32_t)c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (32 * warp + lane < ROWS)
        for (int q = 0; q < 8; ++q) h = h * 1000003ULL + v[q] + (unsigned long long)(c + q) * (32 * warp + lane + 1);
    }
    if (blockIdx.x == 0) atomicAdd(sum, h);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "n"(512));
}

template <int MODE>
int run(int iters, int sms, int nstw, unsigned long long* checksum) {
  const size_t sm = 96 * 1024;
  CK(cudaFuncSetAttribute(coll_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  long long* d_cyc = nullptr; unsigned long long* d_sum = nullptr;
  CK(cudaMalloc(&d_cyc, 8 * sizeof(long long))); CK(cudaMalloc(&d_sum, sizeof(unsigned long long)));
  CK(cudaMemset(d_sum, 0, sizeof(unsigned long long)));
  coll_kernel<MODE><<<sms, 32 * 22, sm>>>(iters, nstw, d_cyc, d_sum);
  CK(cudaDeviceSynchronize());
  long long hc = 0;
  CK(cudaMemcpy(&hc, d_cyc, sizeof(hc), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(checksum, d_sum, sizeof(*checksum), cudaMemcpyDeviceToHost));
  printf("mode %d (%s) store-warps=%2d iters=%5d: %8.1f cycles per 28-MMA stage, %5.1f per MMA, checksum %016llx\n", MODE,
         MODE == 0 ? "plain" : MODE == 1 ? "fill/use/lastuse" : "fill/use", nstw, iters, (double)hc / iters, (double)hc / iters / 28, *checksum);
  cudaFree(d_cyc); cudaFree(d_sum);
  return 0;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  unsigned long long c0 = 0, c1 = 0, c2 = 0;
  for (int it : {3, iters}) {
    for (int nstw : {0, 16}) {
      if (run<0>(it, sms, nstw, &c0)) return 1;
      if (run<1>(it, sms, nstw, &c1)) return 1;
      if (run<2>(it, sms, nstw, &c2)) return 1;
      printf("   accumulators identical: fill/use/lastuse %s, fill/use %s\n", c0 == c1 ? "yes" : "NO", c0 == c2 ? "yes" : "NO");
    }
  }
  return 0;
}
