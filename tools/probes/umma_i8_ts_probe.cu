// Round-2 probe: can the A operand (the G digit planes) of tcgen05.mma kind::i8 come from TENSOR MEMORY instead of
// shared memory? The sweep's M=128, N=64 MMA is bound by its 4 KB + 2 KB of shared-memory operand reads (48 cycles);
// with the plane copied once per stage into TMEM (tcgen05.cp 128x256b: 128 rows x 32 bytes -> 128 lanes x 8 columns)
// the 8 - i MMAs that share plane i read only their 2 KB B operand.
//   part 1 (correctness): one stage, 7 x 7 planes; SS-mode products into accumulators 0..6, TS-mode products into a
//           second accumulator set, compared column by column on the device.
//   part 2 (rate): stages of 7 plane copies + 28 TS-mode MMAs issued back to back (copy of plane i of the next stage
//           right after the last MMA that reads plane i), cycles per stage.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_ts_probe umma_i8_ts_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int NP = 7, KT = 32, N = 32;  // N = 32 here: two accumulator sets of 7 x 32 columns + 56 columns of A planes fit in 512
constexpr int A_PLANE = 128 * KT, B_PLANE = N * KT;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline int swz32(int r, int c) { return (r >> 3) * 256 + (r & 7) * 32 + ((((c >> 4) ^ ((r & 7) >> 2)) & 1) << 4) + (c & 15); }
__device__ __forceinline__ uint64_t desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;\n" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"((uint64_t)s32(bar)) : "memory");
}
__device__ __forceinline__ bool wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  long long spins = 0;
  while (!ok && spins < (1LL << 26)) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    ++spins;
  }
  return ok != 0;
}

__global__ void __launch_bounds__(128, 1) ts_kernel(int iters, int* out, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  unsigned char* sA = smem;                 // [NP][128 x 32] swizzled
  unsigned char* sB = smem + NP * A_PLANE;  // [NP][N x 32] swizzled
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // deterministic signed byte patterns
  for (int u = tid; u < NP * 128 * KT; u += 128) {
    const int p = u / (128 * KT), r = (u / KT) % 128, c = u % KT;
    sA[p * A_PLANE + swz32(r, c)] = (unsigned char)(int8_t)(((r * 7 + c * 13 + p * 5) % 251) - 125);
  }
  for (int u = tid; u < NP * N * KT; u += 128) {
    const int p = u / (N * KT), r = (u / KT) % N, c = u % KT;
    sB[p * B_PLANE + swz32(r, c)] = (unsigned char)(int8_t)(((r * 11 + c * 3 + p * 17) % 241) - 120);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tm = tmem_base;
  const uint32_t acc_ss = tm, acc_ts = tm + NP * N, ta = tm + 2 * NP * N;  // 224 + 224 + 56 columns
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t phase = 0;
  if (tid == 0) {
    for (int i = 0; i < NP; ++i) cp_128x256b(ta + 8 * i, desc(s32(sA + i * A_PLANE)));
    for (int i = 0; i < NP; ++i)
      for (int j = 0; j < NP - i; ++j) {
        mma_ss(acc_ss + (i + j) * N, desc(s32(sA + i * A_PLANE)), desc(s32(sB + j * B_PLANE)), idesc, i > 0 ? 1u : 0u);
        mma_ts(acc_ts + (i + j) * N, ta + 8 * i, desc(s32(sB + j * B_PLANE)), idesc, i > 0 ? 1u : 0u);
      }
    commit(&bar);
    out[1] = wait(&bar, phase) ? 1 : -1;
  }
  phase ^= 1;
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  // compare the two accumulator sets: thread = row
  int bad = 0, nonzero = 0;
  for (int g = 0; g < NP; ++g)
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t a[8], b[8];
      const uint32_t lane_off = ((uint32_t)(32 * warp) << 16);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]) : "r"(acc_ss + lane_off + g * N + c0));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]) : "r"(acc_ts + lane_off + g * N + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int q = 0; q < 8; ++q) { bad += a[q] != b[q]; nonzero += a[q] != 0; }
    }
  atomicAdd(&out[2], bad);
  atomicAdd(&out[3], nonzero);
  if (tid == 5) { out[4] = 0; }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  // ---- part 2: rate of the TS pipeline
  if (tid == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int i = 0; i < NP; ++i) {
        for (int j = 0; j < NP - i; ++j)
          mma_ts(acc_ts + (i + j) * N, ta + 8 * i, desc(s32(sB + j * B_PLANE)), idesc, 1u);
        cp_128x256b(ta + 8 * i, desc(s32(sA + i * A_PLANE)));  // plane i of the next stage, behind its last reader
      }
    }
    commit(&bar);
    const bool ok = wait(&bar, phase);
    const long long t1 = clock64();
    cyc[0] = t1 - t0;
    out[5] = ok ? 1 : -1;
    // reference: the same stage in SS mode
    const long long t2 = clock64();
    for (int it = 0; it < iters; ++it)
      for (int i = 0; i < NP; ++i)
        for (int j = 0; j < NP - i; ++j)
          mma_ss(acc_ss + (i + j) * N, desc(s32(sA + i * A_PLANE)), desc(s32(sB + j * B_PLANE)), idesc, 1u);
    commit(&bar);
    const bool ok2 = wait(&bar, phase ^ 1);
    cyc[1] = clock64() - t2;
    out[6] = ok2 ? 1 : -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "n"(512));
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  int* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_cyc, 64));
  CK(cudaMemset(d_out, 0, 64)); CK(cudaMemset(d_cyc, 0, 64));
  const size_t sm = (size_t)NP * (A_PLANE + B_PLANE) + 1024;
  CK(cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  ts_kernel<<<1, 128, sm>>>(iters, d_out, d_cyc);
  CK(cudaDeviceSynchronize());
  int h[16]; long long c[8];
  CK(cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(c, d_cyc, 64, cudaMemcpyDeviceToHost));
  printf("A-from-TMEM probe (M=128, N=%d, 7 planes): barrier %d, mismatching accumulator entries %d of %d (non-zero SS entries %d)\n",
         N, h[1], h[2], 128 * NP * N, h[3]);
  printf("rate: TS pipeline (7 plane copies + 28 MMAs per stage) %.1f cycles/stage (done=%d); SS mode %.1f cycles/stage (done=%d)\n",
         (double)c[0] / iters, h[5], (double)c[1] / iters, h[6]);
  return 0;
}
