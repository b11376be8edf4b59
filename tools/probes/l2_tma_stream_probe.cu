// How fast can every SM stream L2-resident data into shared memory with TMA bulk copies? (Round-2 plan: the G digit
// planes are 8 bytes per element like today's fp64 G, but a chunk would be consumed ~5x faster.) One CTA per SM,
// 4-stage ring of 32 KB stages, each CTA walks a 64 MB buffer (L2-resident after the first pass) from its own
// starting offset. Prints aggregate TB/s and bytes per cycle per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_tma_stream_probe l2_tma_stream_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
constexpr int STAGE = 32 * 1024, NST = 4;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32, 1) stream_kernel(const unsigned char* __restrict__ buf, size_t bytes, int nchunk, int* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[NST];
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x != 0) return;
  const size_t nst = bytes / STAGE;
  size_t pos = ((size_t)blockIdx.x * 977) % nst;
  auto issue = [&](int c) {
    const int s = c % NST;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(STAGE) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + s * STAGE)),
                 "l"(buf + ((pos + c) % nst) * STAGE), "r"(STAGE), "r"(s32(&full[s])) : "memory");
  };
  for (int c = 0; c < NST - 1 && c < nchunk; ++c) issue(c);
  int acc = 0;
  for (int c = 0; c < nchunk; ++c) {
    if (c + NST - 1 < nchunk) issue(c + NST - 1);  // the stage consumed in iteration c-1
    const int s = c % NST;
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}"
                   : "=r"(ok) : "r"(s32(&full[s])), "r"((uint32_t)((c / NST) & 1)) : "memory");
    acc += smem[s * STAGE + (c & 1023)];
  }
  if (acc == 123456789) sink[0] = acc;
}
int main(int argc, char** argv) {
  const int nchunk = argc > 1 ? atoi(argv[1]) : 4000;
  int dev = 0, sms = 0, khz = 0;
  CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)); CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const size_t bytes = 64ull << 20;
  unsigned char* d; int* sink;
  CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 1, bytes)); CK(cudaMalloc(&sink, 4));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NST * STAGE));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    stream_kernel<<<sms, 32, NST * STAGE>>>(d, bytes, nchunk, sink);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  const double tot = (double)sms * nchunk * STAGE;
  printf("TMA bulk L2 -> smem, %d CTAs x %d x 32 KB: %.3f ms, %.2f TB/s aggregate, %.1f bytes per cycle per SM at %d MHz\n", sms, nchunk,
         best, tot / (best * 1e-3) / 1e12, tot / sms / (best * 1e-3 * khz * 1e3), khz / 1000);
  return 0;
}
