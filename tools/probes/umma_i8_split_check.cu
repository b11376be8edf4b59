// Correctness probe for the round-2 plan: one 128 x 64 output tile of Y = G S^T
// (G: 128 x K doubles with per-row scales, S: 64 x K doubles in [-1,1]) computed as 36 INT8 tensor-core products
// of 7-bit digit planes (tcgen05.mma kind::i8, int32 accumulators in TMEM, one accumulator per digit weight),
// recombined in fp64 and compared with a long-double product on the host. Exercises the parts a real kernel
// depends on: the SWIZZLE_128B K-major shared-memory layout for 1-byte operands and its descriptors, K-stepping
// inside a swizzle row, accumulation across K blocks, the TMEM lane/column layout seen by tcgen05.ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_split_check umma_i8_split_check.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int M = 128, N = 64, NS = 8;  // rows, columns, digit planes; KB = K bytes per staged block (128, 64 or 32)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of (row r, K byte c) inside one K-major swizzled tile whose rows are KB bytes long (SWIZZLE_128B /
// 64B / 32B for KB = 128 / 64 / 32): 8-row groups of 8 KB bytes, the 16-byte chunk index XORed with the low row bits
template <int KB>
__host__ __device__ inline int swz(int r, int c) {
  return (r >> 3) * (8 * KB) + (r & 7) * KB + (((c >> 4) ^ ((r & 7) >> (KB == 128 ? 0 : KB == 64 ? 1 : 2))) << 4) + (c & 15);
}

template <int KB>
__global__ void __launch_bounds__(128, 1) split_kernel(const int8_t* __restrict__ Ap, const int8_t* __restrict__ Bp, int K,
                                                        const double* __restrict__ row_scale, double* __restrict__ Y,
                                                        int* __restrict__ status, int b_unsigned,
                                                        const double* __restrict__ row_offset) {
  constexpr int A_PLANE = M * KB, B_PLANE = N * KB;
  extern __shared__ unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                    // [NS][128 rows x 128 B], swizzled
  unsigned char* sB = smem + NS * A_PLANE;     // [NS][64 rows x 128 B], swizzled
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tm = tmem_base;
  auto desc = [](uint32_t addr) {  // SBO = one 8-row group; layout type 2 / 4 / 6 = SWIZZLE_128B / 64B / 32B
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)((8 * KB) >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(KB == 128 ? 2 : KB == 64 ? 4 : 6) << 61);
  };
  // B format bit: 1 = signed 8-bit, 0 = unsigned 8-bit (the digits of S/2 + 1/2)
  const uint32_t idesc = (2u << 4) | (1u << 7) | ((b_unsigned ? 0u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) |
                         ((uint32_t)(M >> 4) << 24);
  const int nkb = K / KB;
  int ok_all = 1;
  for (int kb = 0; kb < nkb; ++kb) {
    // stage the digit planes of this K block: 16-byte chunks, swizzled
    for (int u = tid; u < NS * M * (KB / 16); u += 128) {
      const int p = u / (M * (KB / 16)), r = (u / (KB / 16)) % M, ch = u % (KB / 16);
      const uint4 v = *reinterpret_cast<const uint4*>(Ap + ((size_t)(p * M + r) * K + (size_t)kb * KB + ch * 16));
      *reinterpret_cast<uint4*>(sA + p * A_PLANE + swz<KB>(r, ch * 16)) = v;
    }
    for (int u = tid; u < NS * N * (KB / 16); u += 128) {
      const int p = u / (N * (KB / 16)), r = (u / (KB / 16)) % N, ch = u % (KB / 16);
      const uint4 v = *reinterpret_cast<const uint4*>(Bp + ((size_t)(p * N + r) * K + (size_t)kb * KB + ch * 16));
      *reinterpret_cast<uint4*>(sB + p * B_PLANE + swz<KB>(r, ch * 16)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    if (tid == 0) {
      for (int i = 0; i < NS; ++i)
        for (int j = 0; i + j < NS; ++j) {
          const int g = i + j;  // digit weight 2^(-7 (g + 2)): one accumulator per weight
          const uint64_t da = desc(s32(sA + i * A_PLANE)), db = desc(s32(sB + j * B_PLANE));
#pragma unroll
          for (int k4 = 0; k4 < KB / 32; ++k4) {
            const uint32_t accumulate = (kb > 0 || i > 0 || k4 > 0) ? 1u : 0u;  // first write of accumulator g: i == 0
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm + (uint32_t)(g * N)),
                "l"(da + (uint64_t)(2 * k4)), "l"(db + (uint64_t)(2 * k4)), "r"(idesc), "r"(accumulate));
          }
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"((uint64_t)s32(&bar)));
      uint32_t ok = 0;
      long long spins = 0;
      while (!ok && spins < (1LL << 26)) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(s32(&bar)), "r"((uint32_t)(kb & 1)) : "memory");
        ++spins;
      }
      if (!ok) ok_all = 0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();  // the MMAs of this block have read the staged planes
    asm volatile("tcgen05.fence::after_thread_sync;\n");
  }
  if (tid == 0) status[0] = ok_all;
  // epilogue: thread (warp, lane) owns row 32 * warp + lane; accumulator g occupies columns [g N, (g + 1) N)
  const int row = 32 * warp + lane;
  const double rs = row_scale[row];
  for (int c0 = 0; c0 < N; c0 += 8) {
    double y[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int g = NS - 1; g >= 0; --g) {  // smallest weight first
      uint32_t v[8];
      const uint32_t ta = tm + ((uint32_t)(32 * warp) << 16) + (uint32_t)(g * N + c0);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(ta));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const double wgt = exp2(-7.0 * (g + 2));
#pragma unroll
      for (int q = 0; q < 8; ++q) y[q] = fma((double)(int)v[q], wgt, y[q]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) Y[(size_t)row * N + c0 + q] = (y[q] - row_offset[row]) * rs;  // offset: (1/2) sum_k g_k
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "n"(512));
}

static void digits_unsigned(long double y, int* d) {  // y in [0, 1] -> 8 base-128 digits (the first may be 128)
  long double q = floorl(y * ldexpl(1.0L, 56) + 0.5L);
  for (int i = NS - 1; i >= 1; --i) { const long double hi = floorl(q / 128.0L); d[i] = (int)(q - hi * 128.0L); q = hi; }
  d[0] = (int)q;
}

static void digits(long double x, int* d) {  // |x| <= 1/2 -> 8 signed 7-bit digits, x ~ sum d_i 2^(-7 (i + 1))
  long double r = x;
  for (int i = 0; i < NS; ++i) {
    const long double w = ldexpl(1.0L, 7 * (i + 1));
    const long double di = rintl(r * w);
    d[i] = (int)di;
    r -= di / w;
  }
}

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 1024;
  const int b_unsigned = argc > 2 ? atoi(argv[2]) : 0;
  const int KBr = argc > 3 ? atoi(argv[3]) : 128;  // K bytes per staged block = swizzle width: 128, 64 or 32
  if ((KBr != 128 && KBr != 64 && KBr != 32) || K % KBr) { printf("usage: K (multiple of the stage) [unsigned 0|1] [stage 128|64|32]\n"); return 1; }
  std::mt19937_64 rng(7);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::uniform_real_distribution<double> ud(-1.0, 1.0), us(-2.0, 2.0);
  std::vector<double> G((size_t)M * K), S((size_t)N * K), rs(M), roff(M, 0.0);
  std::vector<int8_t> Ap((size_t)NS * M * K), Bp((size_t)NS * N * K);
  for (int m = 0; m < M; ++m) {
    const double sc = pow(10.0, us(rng));
    double mx = 0;
    for (int k = 0; k < K; ++k) { G[(size_t)m * K + k] = nd(rng) * sc * pow(10.0, 0.5 * ud(rng)); mx = fmax(mx, fabs(G[(size_t)m * K + k])); }
    const int e = (int)ceil(log2(mx)) + 1;   // |G / 2^e| <= 1/2
    rs[m] = ldexp(1.0, e + 1);               // times the 2^1 of S
    for (int k = 0; k < K; ++k) {
      int d[NS];
      digits(ldexpl((long double)G[(size_t)m * K + k], -e), d);
      for (int i = 0; i < NS; ++i) Ap[((size_t)i * M + m) * K + k] = (int8_t)d[i];
    }
    if (b_unsigned) {  // (1/2) sum_k g_k with g = G / 2^e: what the +1/2 offset of every S entry adds to this row
      long double acc = 0;
      for (int k = 0; k < K; ++k) acc += ldexpl((long double)G[(size_t)m * K + k], -e);
      roff[m] = (double)(0.5L * acc);
    }
  }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      S[(size_t)n * K + k] = ud(rng);
      int d[NS];
      if (b_unsigned) digits_unsigned((long double)S[(size_t)n * K + k] * 0.5L + 0.5L, d);
      else digits((long double)S[(size_t)n * K + k] * 0.5L, d);
      for (int i = 0; i < NS; ++i) Bp[((size_t)i * N + n) * K + k] = (int8_t)(uint8_t)d[i];
    }
  int8_t *dA, *dB; double *dR, *dY, *dO; int* dS;
  CK(cudaMalloc(&dO, M * 8));
  CK(cudaMemcpy(dO, roff.data(), M * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dA, Ap.size())); CK(cudaMalloc(&dB, Bp.size())); CK(cudaMalloc(&dR, M * 8)); CK(cudaMalloc(&dY, (size_t)M * N * 8));
  CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, Ap.data(), Ap.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bp.data(), Bp.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dR, rs.data(), M * 8, cudaMemcpyHostToDevice));
  const size_t sm = (size_t)NS * (M + N) * KBr + 1024;
  if (KBr == 128) {
    CK(cudaFuncSetAttribute(split_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    split_kernel<128><<<1, 128, sm>>>(dA, dB, K, dR, dY, dS, b_unsigned, dO);
  } else if (KBr == 64) {
    CK(cudaFuncSetAttribute(split_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    split_kernel<64><<<1, 128, sm>>>(dA, dB, K, dR, dY, dS, b_unsigned, dO);
  } else {
    CK(cudaFuncSetAttribute(split_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    split_kernel<32><<<1, 128, sm>>>(dA, dB, K, dR, dY, dS, b_unsigned, dO);
  }
  CK(cudaDeviceSynchronize());
  std::vector<double> Y((size_t)M * N);
  int st = 0;
  CK(cudaMemcpy(Y.data(), dY, Y.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  double worst = 0, worst64 = 0, sum = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      long double t = 0, sc = 0; double p64 = 0;
      for (int k = 0; k < K; ++k) {
        const long double a = G[(size_t)m * K + k], b = S[(size_t)n * K + k];
        t += a * b; sc += fabsl(a * b); p64 = fma(G[(size_t)m * K + k], S[(size_t)n * K + k], p64);
      }
      const double e = (double)(fabsl((long double)Y[(size_t)m * N + n] - t) / sc) / 2.220446049250313e-16;
      const double e64 = (double)(fabsl((long double)p64 - t) / sc) / 2.220446049250313e-16;
      worst = fmax(worst, e); worst64 = fmax(worst64, e64); sum += e;
    }
  printf("K=%d, stage/swizzle %d B, B digits %s, 36 INT8 products per block, barriers completed=%d: split product max err %.3f mean %.3f, plain fp64 fma chain max err %.3f "
         "[eps * sum|G||S|]\n", K, KBr, b_unsigned ? "unsigned (offset 1/2)" : "signed", st, worst, sum / (M * N), worst64);
  return worst < 4.0 ? 0 : 2;
}
