// Round-2 shape probe: cycles per staged block of the digit-plane product as the sweep kernel issues it --
// NP planes per operand (A: G digit planes, ROWS real rows inside the 128-row operand, plane stride = ROWS x 32
// bytes so the tail of each 128-row read runs into the next plane: those output lanes are never read; B: the sin/cos
// digit planes, N rows), SWIZZLE_32B K-major tiles of 32 TOAs, one tcgen05.mma kind::i8 (M=128, N, K=32) per plane
// pair (i, j) with i + j <= MAXG, accumulator i + j at TMEM column (i + j) * N. One issuing thread per CTA, one CTA
// per SM, optionally NSTW extra warps streaming 4-byte shared-memory stores (the producers' traffic) next to it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_shape_probe umma_i8_shape_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N, int NP, int ROWS, int MAXG>
__global__ void __launch_bounds__(32 * 18, 1) shape_kernel(int iters, int nstw, long long* cyc, int* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  constexpr int KB = 32, A_PLANE = ROWS * KB, B_PLANE = N * KB;
  constexpr int A_BYTES = (NP - 1) * A_PLANE + 128 * KB, B_BYTES = NP * B_PLANE;
  unsigned char* sA = smem;
  unsigned char* sB = smem + ((A_BYTES + 1023) / 1024) * 1024;
  unsigned char* sW = sB + ((B_BYTES + 1023) / 1024) * 1024;  // 16 KB scratch the store warps write
  for (int i = threadIdx.x; i < A_BYTES + B_BYTES + 2048; i += blockDim.x) smem[i] = (unsigned char)((i * 7 + 3) & 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tm = tmem_base;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    auto desc = [](uint32_t addr) {  // SWIZZLE_32B: SBO = 8 rows x 32 B = 256 B, layout type 6
      return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) |
             ((uint64_t)6 << 61);
    };
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < NP; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          if (i + j > MAXG) continue;
          const int g = i + j;
          const uint32_t acc = tm + (uint32_t)((g * N) % 512);
          const uint32_t scale = (it > 0 || i > 0) ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(acc),
              "l"(desc(s32(sA + i * A_PLANE))), "l"(desc(s32(sB + j * B_PLANE))), "r"(idesc), "r"(scale));
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
    const long long t1 = clock64();
    stop = 1;
    if (blockIdx.x == 0) cyc[0] = t1 - t0;
    if (out) out[blockIdx.x] = ok ? 1 : -1;
  } else if (warp >= 2 && warp < 2 + nstw) {
    // producer-like traffic: conflict-free 4-byte stores, 14 per "task", until the MMA thread is done
    uint32_t* w = reinterpret_cast<uint32_t*>(sW) + (warp - 2) * 32 + lane;
    uint32_t v = lane;
    long long cnt = 0;
    while (!stop) {
#pragma unroll
      for (int q = 0; q < 14; ++q) { w[(q & 3) * 1024] = v; v = v * 3 + 1; }
      ++cnt;
    }
    if (blockIdx.x == 0 && lane == 0) cyc[1 + warp] = cnt;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "n"(512));
}

template <int N, int NP, int ROWS, int MAXG>
int run(int iters, int sms, int nstw) {
  const size_t sm = 96 * 1024;
  CK(cudaFuncSetAttribute(shape_kernel<N, NP, ROWS, MAXG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  int* d_out = nullptr; long long* d_cyc = nullptr;
  CK(cudaMalloc(&d_out, sms * sizeof(int))); CK(cudaMalloc(&d_cyc, 32 * sizeof(long long)));
  CK(cudaMemset(d_cyc, 0, 32 * sizeof(long long)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    shape_kernel<N, NP, ROWS, MAXG><<<sms, 32 * 18, sm>>>(iters, nstw, d_cyc, d_out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  int h0 = 0; long long hc[32];
  CK(cudaMemcpy(&h0, d_out, sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hc, d_cyc, sizeof(hc), cudaMemcpyDeviceToHost));
  int nmma = 0;
  for (int i = 0; i < NP; ++i) for (int j = 0; j < NP; ++j) if (i + j <= MAXG) ++nmma;
  const double macs = (double)sms * iters * nmma * 128.0 * N * 32.0;
  long long stores = 0;
  for (int w = 0; w < nstw; ++w) stores += hc[3 + w];
  printf("N=%3d planes=%d rows=%3d maxg=%d (%2d MMAs/stage) store-warps=%2d: %8.1f cycles/stage, %5.1f cycles/MMA, %6.1f TOP/s, "
         "%.0f store-tasks/stage (completed=%d)\n", N, NP, ROWS, MAXG, nmma, nstw, (double)hc[0] / iters,
         (double)hc[0] / iters / nmma, 2.0 * macs / (best * 1e-3) / 1e12, (double)stores / iters, h0);
  cudaFree(d_out); cudaFree(d_cyc);
  return 0;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 4000;
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  for (int nstw : {0, 8, 16}) {
    if (run<64, 7, 80, 6>(iters, sms, nstw)) return 1;    // the plan: radix 256, 7 planes, 28 MMAs, 7 accumulators
    if (run<64, 7, 80, 7>(iters, sms, nstw)) return 1;    // + the i + j = 7 products (8 accumulators = all of TMEM)
  }
  if (run<64, 7, 128, 6>(iters, sms, 0)) return 1;        // same with full 128-row planes (no overlapped tails)
  if (run<64, 8, 80, 7>(iters, sms, 0)) return 1;         // round-1 plan: radix 128, 8 planes, 36 MMAs
  if (run<128, 7, 80, 6>(iters, sms, 0)) return 1;        // N = 128 (TMEM could not hold it: rate reference only)
  if (run<256, 7, 80, 6>(iters, sms, 0)) return 1;
  if (run<32, 7, 80, 6>(iters, sms, 0)) return 1;
  if (run<48, 7, 80, 6>(iters, sms, 0)) return 1;
  return 0;
}
