// The frequency-sweep kernel on the 5th-generation tensor cores: Y = G [s c] as an error-free product of 8-bit
// digit planes (tcgen05.mma kind::i8, exact int32 accumulation in tensor memory), everything else in fp64.
//
// Replaces the body of FastFp.calculate_Fp under jax.vmap (reference fastfp/fastfp.py:69-92, examples/run_fp.py:63)
// like fp_sweep_kernel.cuh does, for packs with n <= 16384 TOAs per pulsar, a diagonal N and up to 639 basis columns
// (128 operand rows -- 127 columns + the C^-1 r row -- per pass over the TOAs; wider bases take one pass per row group
// of 128, the b-sums adding up in the epilogue); the fp64 DMMA kernel stays as the path for everything else. Why: fp64 has no tcgen05 kind, and the DMMA
// formulation is pinned at 0.68 of the fp64 pipe it has to share with the sincos generation.
// The INT8 tensor path is a different unit altogether, and integer accumulation is exact.
//
// Number format (radix 256, 7 planes per operand, validated at the level of the statistic by
// tests/test_split_precision_emulation.py):
//   G row j (and the extra row w = C^-1 r):  g = G / 2^e_j with |g| < 1/4,  Qg = rint(g 2^55),
//        Qg = sum_i d_i 256^(6-i), balanced digits d_i in [-128, 127]  (pack time, i8_planes_kernel)
//   s, c in [-1, 1]:  Q = rint(x 2^54) = sum_j v_j 256^(6-j), balanced digits (producers: an exponent add and one
//        F2I.S64.F64; the bytes of Q + 0x80..80 are the digits + 128)
//   Y_j = 2^(e_j - 13) sum_{g=0..6} 256^-g  sum_{i+j=g} sum_k d_i(k) v_j(k):   28 plane products, one int32
//        accumulator per weight g (|d v| <= 2^14, 7 products per TOA: exact for n <= 18 724 TOAs)
//
// One 768-thread CTA per SM, static round-robin over (pulsar, 32-frequency tile) work items, per stage of 32 TOAs:
//   warp 0   lane 0: TMA -- one bulk copy of the stage's G planes (7 x rows x 32 bytes, already in the SWIZZLE_32B
//            K-major operand layout) and one of its (t, 1/N) vectors, mbarrier rings
//   warps 1-3 lane 0: 28 tcgen05.mma (M = 128: rows of G, N = 64: 32 frequencies x {sin, cos}, K = 32 TOAs) into the
//            7 accumulators (7 x 64 = 448 of the 512 TMEM columns), split by accumulator over three issuing threads;
//            tcgen05.commit frees the stage
//   warps 8-23 (producers, two groups alternating stages): sincos_cw of ((2 pi) f) t (fastfp.py:78-79 phase order),
//            the digit split, 14 conflict-free 4-byte stores per thread into the B-operand planes, and the fp64
//            sums s N^-1 s, s N^-1 c (c N^-1 c = sum 1/N - s N^-1 s)
//   warps 4-7 (epilogue, one TMEM lane quarter each): tcgen05.ld, fp64 recombination, b = Y_s.Y_s, Y_s.Y_c, Y_c.Y_c
//            over the basis rows, (s|r), (c|r) from the w row, pivoted 2x2 solve (jnp.linalg.solve at fastfp.py:90)
// Shared memory is the bound: an M=128, N=64 MMA reads 6 KB of operands, 48 cycles at 128 B/clk (measured,
// tools/probes/umma_i8_shape_probe.cu), i.e. 1344 cycles per stage against ~3490 for the same work on the DMMA path.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/fastfp_b200.h"
#include "ffp_internal.cuh"
#include "ffp_sincos.cuh"

namespace ffp {
namespace i8 {

constexpr int NPL = 7;                 // digit planes per operand
constexpr int KT = 32;                 // TOAs per stage = K bytes per operand row (one SWIZZLE_32B atom wide)
constexpr int NF = 32;                 // frequencies per work item
constexpr int NBR = 2 * NF;            // rows of the B operand: row = 32 * {0: sin, 1: cos} + frequency
constexpr int S_PLANE = NBR * KT;      // 2048 bytes
constexpr int S_STAGE = NPL * S_PLANE; // 14336 bytes
constexpr int V_STAGE = KT * 16;       // (t, 1/N) per TOA
constexpr int SST = 6, VST = 8;        // ring depths (even: a slot is always served by the same producer group)
constexpr int RS = 640;                // row-scale stride per pulsar: up to 5 row groups of 128 operand rows (m <= 639)
constexpr int NISSUE = 3;               // MMA-issuing threads (control warps 1-3), each owning a set of accumulators
// Warp layout: warps 0-3 control (TMA, MMA issue), 4-7 epilogue, 8.. producers. A stage is always produced by 8
// warps (thread = one frequency x four TOAs); with NPW = 16 two groups of 8 alternate stages, with NPW = 8 one group
// takes every stage and each producer thread gets twice the registers.
template <int NPW>
struct Roles {
  static_assert(NPW == 8 || NPW == 16, "8 or 16 producer warps");
  static constexpr int THREADS = 256 + 32 * NPW;
  static constexpr int NG = NPW / 8;                                   // producer groups
  static constexpr int REGS_LAUNCH = NPW == 16 ? 80 : 128;             // 65536 / THREADS, multiple of 8
  static constexpr int REGS_CTRL = 56, REGS_EPI = 136, REGS_PROD = NPW == 16 ? 72 : 152;
  static_assert(128 * REGS_CTRL + 128 * REGS_EPI + 32 * NPW * REGS_PROD <= THREADS * REGS_LAUNCH, "register pool");
};
static_assert(SST % 2 == 0 && VST % 2 == 0, "ring depths must be even");

struct Args {
  const unsigned char* planes;   // per pulsar, per stage: [V_STAGE bytes (t, 1/N)] then per row group [NPL][rows_g][32] swizzled
  const double* rowscale;        // [P][RS]  2^(e_j - 13), 0 for rows that do not exist
  const PulsarMeta* meta;
  const int* pidx;
  const double* freqs;
  int64_t F;
  double* terms;                 // [P][F]
  double* inner;                 // optional [P][F][5]: (s|s), (s|c), (c|c), (s|r), (c|r); null = terms only
  double* Z;                     // nmfp stage A: [P][ceil(F/32)][mvpad/4][8][32] z'_s, z'_c tiles (B-fragment order)
  double* A;                     // nmfp stage A: [P][ceil(F/32)][5][32] a_ss, a_sc, a_cc, a_sr, a_cr
  int mvpad;
  int ntile, nwork;
  int gslot, gst;                // G ring: bytes per slot (7 x rows_max x 32), number of slots
#ifdef FFP_I8_TRACE
  long long* trace;              // [TRACE_EV][TRACE_K] clock64 of CTA 0's first stages (tools/i8_trace.py)
#endif
};
#ifdef FFP_I8_TRACE
constexpr int TRACE_EV = 8, TRACE_K = 512;
#define FFP_TRACE(ev, k) do { if (ar.trace && blockIdx.x == 0 && (k) < (uint32_t)TRACE_K) ar.trace[(ev) * TRACE_K + (k)] = clock64(); } while (0)
#else
#define FFP_TRACE(ev, k) do { } while (0)
#endif

// byte offset of (row r, K byte c) in a K-major tile with 32-byte rows, SWIZZLE_32B: 8-row groups of 256 bytes, the
// 16-byte chunk index XORed with bit 2 of the row (validated by tools/probes/umma_i8_split_check.cu)
__host__ __device__ inline int swz32(int r, int c) {
  return (r >> 3) * 256 + (r & 7) * 32 + ((((c >> 4) ^ ((r & 7) >> 2)) & 1) << 4) + (c & 15);
}

// ---- pack time: digit planes of G (+ the w row) --------------------------------------------------------------
// e_j from the row maximum: |G_j / 2^e_j| < 1/4; rs_j = 2^(e_j - 13). bad[p] is set when a row holds a non-finite
// value (singular Sigma, NaN data): the integer planes could not carry it, so such a pack stays on the fp64 path.
__global__ void i8_rowscale_kernel(const double* __restrict__ packets, const PulsarMeta* __restrict__ meta,
                                   double* __restrict__ rowscale, int* __restrict__ rowexp, int* __restrict__ bad) {
  const PulsarMeta pm = meta[blockIdx.y];
  const int r = blockIdx.x;
  if (r >= RS || pm.i8_nst == 0) return;  // (pulsars the tensor sweep does not take have no stages)
  double* rs = rowscale + (size_t)blockIdx.y * RS + r;
  int* re = rowexp + (size_t)blockIdx.y * RS + r;
  if (r > pm.m) {
    if (threadIdx.x == 0) { *rs = 0.0; *re = 0; }
    return;
  }
  const int CI = pm.ci, mp = pm.mpad, pkw = CI * (4 + mp);
  const double* pk0 = packets + pm.pk_off;
  double mx = 0.0;
  bool nonfinite = false;
  for (int i = threadIdx.x; i < pm.n; i += blockDim.x) {
    const double* pk = pk0 + (size_t)(i / CI) * pkw;
    const int il = i % CI;
    const double v = r < pm.m ? pk[4 * CI + g_frag_index(il, r, mp >> 3)] : pk[4 * il + 2];  // G row or w
    const double a = fabs(v);
    if (!(a <= 1.7e308)) nonfinite = true;
    mx = fmax(mx, a);
  }
  __shared__ double smx[32];
  __shared__ int snf;
  if (threadIdx.x == 0) snf = 0;
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) smx[threadIdx.x >> 5] = mx;
  if (nonfinite) atomicOr(&snf, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, smx[w]);
    int e = 0;
    if (mx > 0.0) {
      int x;
      frexp(mx, &x);  // mx = f 2^x, f in [1/2, 1)  ->  |G| < 2^x
      e = x + 2;
    }
    *re = e;
    *rs = mx > 0.0 ? scalbn(1.0, e - 13) : 0.0;
    if (snf) atomicOr(&bad[blockIdx.y], 1);
  }
}

// one CTA per (stage, pulsar): the (t, 1/N) vectors and the 7 x rows x 32 digit bytes of the stage
__global__ void i8_planes_kernel(const double* __restrict__ packets, const PulsarMeta* __restrict__ meta,
                                 const int* __restrict__ rowexp, unsigned char* __restrict__ planes) {
  const PulsarMeta pm = meta[blockIdx.y];
  const int st = blockIdx.x;
  if (st >= pm.i8_nst) return;
  const int rows = pm.i8_rows, CI = pm.ci, mp = pm.mpad, pkw = CI * (4 + mp);
  const double* pk0 = packets + pm.pk_off;
  unsigned char* out = planes + pm.i8_off + (size_t)st * (V_STAGE + NPL * rows * KT);
  double2* v = reinterpret_cast<double2*>(out);
  for (int kk = threadIdx.x; kk < KT; kk += blockDim.x) {
    const int i = st * KT + kk;
    double2 tv = make_double2(0.0, 0.0);  // padded TOAs: weight 0 (and zero digits below)
    if (i < pm.n) {
      const double* pk = pk0 + (size_t)(i / CI) * pkw;
      tv = make_double2(pk[4 * (i % CI)], pk[4 * (i % CI) + 1]);
    }
    v[kk] = tv;
  }
  // row groups of 128 operand rows (one accumulator set each, one pass over the TOAs per group): group-major, then
  // plane-major, so that a (stage, group) is one contiguous TMA copy
  unsigned char* g = out + V_STAGE;
  const int* re = rowexp + (size_t)blockIdx.y * RS;
  for (int e = threadIdx.x; e < rows * KT; e += blockDim.x) {
    const int r = e / KT, kk = e - r * KT;
    const int i = st * KT + kk;
    double val = 0.0;
    if (i < pm.n && r <= pm.m) {
      const double* pk = pk0 + (size_t)(i / CI) * pkw;
      const int il = i % CI;
      val = r < pm.m ? pk[4 * CI + g_frag_index(il, r, mp >> 3)] : pk[4 * il + 2];
    }
    // Qg = rint(val 2^(55 - e)): exact scaling, |Qg| <= 2^53; balanced base-256 digits = bytes of Qg + 0x80..80, - 128
    const long long Q = __double2ll_rn(scalbn(val, 55 - re[r]));
    const unsigned long long U = (unsigned long long)(Q + 0x0080808080808080LL) ^ 0x0080808080808080ULL;
    const int grp = r >> 7, rl = r & 127, rows_g = min(128, rows - 128 * grp);
    unsigned char* gg = g + (size_t)grp * (NPL * 128 * KT);
    const int off = swz32(rl, kk);
#pragma unroll
    for (int p = 0; p < NPL; ++p) gg[(size_t)p * rows_g * KT + off] = (unsigned char)(U >> (8 * (6 - p)));
  }
}

// ---- device helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  // K-major, SWIZZLE_32B (layout type 6), stride between 8-row groups 256 bytes, descriptor version 1
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)6 << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"((uint64_t)smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int R>
__device__ __forceinline__ void reg_set_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void reg_set_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

// x in [-1, 1] -> the 7 bytes of Q + 0x80..80, Q = rint(x 2^54), every byte XORed with 0x80: balanced signed digits,
// most significant in byte 6. The scaling is an integer add on the exponent field (zero and subnormals land below
// 2^-900 and round to 0; the fast path is entered for finite data only) and the rounding ONE conversion instruction
// (F2I.S64.F64): conversions keep their rate while tcgen05 MMAs are in flight, fp64 multiplies and FMAs do not
// (tools/probes/umma_fp64_overlap_probe.cu), and an all-integer extraction costs ~30 instructions per value.
__device__ __forceinline__ uint2 digits7(double x) {
  const long long Q = __double2ll_rn(__hiloint2double(__double2hiint(x) + (54 << 20), __double2loint(x)));
  const long long U = Q + 0x0080808080808080LL;
  return make_uint2((uint32_t)U ^ 0x80808080u, (uint32_t)((unsigned long long)U >> 32) ^ 0x00808080u);
}

// Every wait of this kernel is bounded (wait_wd below): a protocol error would otherwise hang the GPU; a timed-out wait
// reports where it was stuck and traps, which the host sees as a launch failure instead of a hung device.
__device__ __forceinline__ void wait_timeout(int tag, uint32_t k) {
  printf("[fastfp_b200 i8 sweep] mbarrier wait timed out: tag %d, stage/item %u, block %d, warp %d\n", tag, k,
         (int)blockIdx.x, (int)(threadIdx.x >> 5));
  __trap();
}
// mbarrier.try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes or HINT_NS pass
// (without the hint the time slice is a few tens of nanoseconds, and __nanosleep between polls was measured not to
// lengthen it: the four epilogue warps alone executed a quarter of the kernel's instructions polling for their item).
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      " selp.b32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));  // volatile: stays inside the branch that asks for it
  return t;
}
// SLEEP_NS = 0: the waiter is on the critical path and polls with the default time slice; otherwise each poll may
// suspend the warp for up to SLEEP_NS, so that waiting warps do not spend issue slots the producers need. Bounded: no
// legitimate wait is longer than one item (< 1 ms); after 5 s on the global timer (looked at every 256 polls) the CTA
// reports where it was stuck and traps.
template <int SLEEP_NS>
__device__ __forceinline__ void wait_wd(uint64_t* bar, uint32_t parity, int tag, uint32_t k) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t n = 0;
  unsigned long long t0 = 0;
  while (!(SLEEP_NS ? mbar_try_wait_hint(bar, parity, (uint32_t)SLEEP_NS) : mbar_try_wait(bar, parity))) {
    if ((++n & 255u) == 0) {
      const unsigned long long t = global_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 5000000000ULL) wait_timeout(tag, k);
    }
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// The products of one 32-TOA stage that issuer Q owns: accumulator g = i + j collects plane i of G times plane j of
// [s c]; the accumulators are dealt to the three issuers as {6, 1}, {5, 2}, {4, 3, 0} (9 + 9 + 10 products). Descriptors
// differ in their 14-bit address field only, so each is one add on the low word of the stage's base descriptor.
template <int Q>
__device__ __forceinline__ void issue_stage(uint32_t tm, uint32_t da_lo, uint32_t db_lo, uint32_t aplane16,
                                            uint32_t idesc, bool first_stage) {
  constexpr uint32_t OWNER[NPL] = {2, 0, 1, 2, 2, 1, 0};
  constexpr uint64_t HI = ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
#pragma unroll
    for (int j = 0; j < NPL - i; ++j) {  // the first product into accumulator j of an item is (0, j)
      if (OWNER[i + j] == (uint32_t)Q)
        umma_i8(tm + (uint32_t)((i + j) * NBR), HI | (uint64_t)(da_lo + (uint32_t)i * aplane16),
                HI | (uint64_t)(db_lo + (uint32_t)j * (S_PLANE >> 4)), idesc, (!first_stage || i > 0) ? 1u : 0u);
    }
  }
}

struct Smem {
  unsigned char *G, *S, *V;
  double *redA;          // [2][2][NF][3] producer sums: buffer, producer group (one or two), frequency
  double *part;          // [4][NF][3]    epilogue partial b-sums per warp
  double *nval;          // [NF][2]       (s|r), (c|r) from the w row
  uint64_t *g_full, *g_empty, *v_full, *v_empty, *s_full, *s_empty, *acc_full, *acc_empty, *sums_full, *sums_empty;
  uint32_t* tmem;
  __device__ Smem(unsigned char* raw, const Args& ar) {
    G = raw;
    S = G + (size_t)ar.gst * ar.gslot;
    V = S + SST * S_STAGE;
    redA = reinterpret_cast<double*>(V + VST * V_STAGE);
    part = redA + 2 * 2 * NF * 3;
    nval = part + 4 * NF * 3;
    g_full = reinterpret_cast<uint64_t*>(nval + NF * 2);
    g_empty = g_full + 8;
    v_full = g_empty + 8;
    v_empty = v_full + VST;
    s_full = v_empty + VST;
    s_empty = s_full + SST;
    acc_full = s_empty + SST;
    acc_empty = acc_full + 1;
    sums_full = acc_empty + 1;
    sums_empty = sums_full + 2;
    tmem = reinterpret_cast<uint32_t*>(sums_empty + 2);
  }
};
constexpr size_t SMEM_FIXED = (size_t)SST * S_STAGE + VST * V_STAGE + (2 * 2 * NF * 3 + 4 * NF * 3 + NF * 2) * 8 +
                              (8 + 8 + 2 * VST + 2 * SST + 2 + 4) * 8 + 16;

template <bool NMFP, int NPW>
__global__ void __launch_bounds__(Roles<NPW>::THREADS, 1) fp_sweep_i8_kernel(const Args ar) {
  using R = Roles<NPW>;
  constexpr int NG = R::NG;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem sm(smem_raw, ar);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < ar.gst; ++s) { mbar_init(&sm.g_full[s], 1); mbar_init(&sm.g_empty[s], NISSUE); }
    for (int s = 0; s < VST; ++s) { mbar_init(&sm.v_full[s], 1); mbar_init(&sm.v_empty[s], 8); }
    for (int s = 0; s < SST; ++s) { mbar_init(&sm.s_full[s], 8); mbar_init(&sm.s_empty[s], NISSUE); }
    mbar_init(sm.acc_full, NISSUE);
    mbar_init(sm.acc_empty, 4);
    for (int b = 0; b < 2; ++b) { mbar_init(&sm.sums_full[b], NPW); mbar_init(&sm.sums_empty[b], 1); }
    fence_barrier_init();
  }
  if (wid == 1) {  // the MMA warp owns the tensor-memory allocation (all 512 columns: one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(sm.tmem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *sm.tmem;

  if (wid < 4) {
    // ================= control warpgroup: TMA (warp 0) and MMA issue (warp 1) =================
    reg_set_dec<R::REGS_CTRL>();
    if (wid == 0 && lane == 0) {
      // ring positions and phase parities are carried incrementally (a division by the runtime ring depth per stage
      // costs ~25 dependent instructions on a thread that competes with 16 producer warps for issue slots)
      uint32_t k = 0, sv = 0, vpar = 0, sg = 0, gpar = 0;
      unsigned char* gdst = sm.G;
      for (int item = blockIdx.x; item < ar.nwork; item += gridDim.x) {
        const PulsarMeta pm = ar.meta[ar.pidx[item / ar.ntile]];
        const size_t stage_bytes = (size_t)V_STAGE + (size_t)NPL * pm.i8_rows * KT;
        const int ng = (pm.i8_rows + 127) >> 7;
        for (int grp = 0; grp < ng; ++grp) {  // one pass over the TOAs per group of 128 operand rows
          const int rows_g = min(128, pm.i8_rows - 128 * grp);
          const uint32_t gbytes = (uint32_t)(NPL * rows_g * KT);
          const unsigned char* src = ar.planes + pm.i8_off;
          const size_t goff = (size_t)V_STAGE + (size_t)grp * (NPL * 128 * KT);
          for (int c = 0; c < pm.i8_nst; ++c, ++k) {
            if (k >= VST) wait_wd<2000>(&sm.v_empty[sv], vpar ^ 1u, 1, k);
            mbar_expect_tx(&sm.v_full[sv], V_STAGE);
            tma_load_1d(sm.V + sv * V_STAGE, src, V_STAGE, &sm.v_full[sv]);
            if (k >= (uint32_t)ar.gst) wait_wd<2000>(&sm.g_empty[sg], gpar ^ 1u, 2, k);
            if (k >= (uint32_t)ar.gst) FFP_TRACE(0, k - (uint32_t)ar.gst);   // MMAs of stage k - gst have completed (seen by the TMA thread)
            mbar_expect_tx(&sm.g_full[sg], gbytes);
            tma_load_1d(gdst, src + goff, gbytes, &sm.g_full[sg]);
            src += stage_bytes;
            if (++sv == VST) { sv = 0; vpar ^= 1u; }
            if (++sg == (uint32_t)ar.gst) { sg = 0; gpar ^= 1u; gdst = sm.G; } else gdst += ar.gslot;
          }
        }
      }
    } else if (wid >= 1) {
      // Three MMA-issuing warps, one per remaining control warp (= one per SM sub-partition). A single issuing thread
      // was measured to be the critical path (it never waited: 2900 cycles per stage for 28 MMAs against 1344 of
      // tensor time): next to producer warps it gets an issue slot only every few cycles. Each issuer owns a set of
      // accumulators, so no ordering between issuers is needed: every accumulator is written by one thread only, in
      // program order. The whole warp runs the loop with warp-uniform values (broadcast by shuffle, so that the
      // compiler keeps descriptors in uniform registers instead of serialising lanes around every MMA) and one
      // elected lane issues.
      const bool leader = elect_one();
      // instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, N = 64, M = 128
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NBR >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t tmu = __shfl_sync(0xffffffffu, tm, 0);
      const uint32_t g0 = smem_u32(sm.G), s0 = smem_u32(sm.S);
      uint32_t k = 0, ps = 0;  // global stage counter; pass counter (one pass = one row group of one item)
      uint32_t sg = 0, gpar = 0, ss = 0, spar = 0, ga = g0, sa = s0;  // ring positions, parities, slot addresses
      for (int item = blockIdx.x; item < ar.nwork; item += gridDim.x) {
        const int pi = ar.pidx[item / ar.ntile];
        const int rows = __shfl_sync(0xffffffffu, ar.meta[pi].i8_rows, 0);
        const int nst = __shfl_sync(0xffffffffu, ar.meta[pi].i8_nst, 0);
        for (int grp = 0; 128 * grp < rows; ++grp, ++ps) {
          const uint32_t aplane16 = (uint32_t)(min(128, rows - 128 * grp) * KT) >> 4;
          if (ps > 0) wait_wd<0>(sm.acc_empty, (ps - 1) & 1u, 3, ps);  // the epilogue has drained the accumulators
          for (int c = 0; c < nst; ++c, ++k) {
            wait_wd<0>(&sm.g_full[sg], gpar, 4, k);
            wait_wd<0>(&sm.s_full[ss], spar, 5, k);
            tc_fence_after();
            if (wid == 1 && leader) FFP_TRACE(1, k);   // issuer 0 sees stage k complete
            const uint32_t da_lo = ((ga >> 4) & 0x3fffu) | (1u << 16);
            const uint32_t db_lo = ((sa >> 4) & 0x3fffu) | (1u << 16);
            if (leader) {
              if (wid == 1) issue_stage<0>(tmu, da_lo, db_lo, aplane16, idesc, c == 0);
              else if (wid == 2) issue_stage<1>(tmu, da_lo, db_lo, aplane16, idesc, c == 0);
              else issue_stage<2>(tmu, da_lo, db_lo, aplane16, idesc, c == 0);
              umma_commit(&sm.g_empty[sg]);  // each issuer's commit arrives when ITS MMAs above have read their operands
              umma_commit(&sm.s_empty[ss]);
              if (wid == 1) FFP_TRACE(2, k);           // issuer 0 has issued its MMAs of stage k
            }
            __syncwarp();
            if (++sg == (uint32_t)ar.gst) { sg = 0; gpar ^= 1u; ga = g0; } else ga += (uint32_t)ar.gslot;
            if (++ss == SST) { ss = 0; spar ^= 1u; sa = s0; } else sa += S_STAGE;
          }
          if (leader) umma_commit(sm.acc_full);
          __syncwarp();
        }
      }
    }
  } else if (wid < 8) {
    // ================= epilogue warpgroup: one TMEM lane quarter per warp =================
    reg_set_inc<R::REGS_EPI>();
    const int ew = wid - 4;
    const int lrow = 32 * ew + lane;   // TMEM lane = operand row inside the row group
    uint32_t it = 0, ps = 0;  // items; passes (one per row group of an item)
    for (int item = blockIdx.x; item < ar.nwork; item += gridDim.x, ++it) {
      const int gp = item / ar.ntile, ft = item - gp * ar.ntile;
      const int p = ar.pidx[gp];
      const PulsarMeta pm = ar.meta[p];
      const int64_t f0 = (int64_t)ft * NF;
      // bases wider than 127 columns take one pass per group of 128 operand rows; the b-sums of the groups add up in
      // `part`, the w row sits in the last group
      for (int grp = 0; 128 * grp < pm.i8_rows; ++grp, ++ps) {
      const int row = 128 * grp + lrow;
      const double rs = ar.rowscale[(size_t)p * RS + row];
      const bool has_rows = 128 * grp + 32 * ew < pm.i8_rows;  // warp-uniform
      wait_wd<100000>(sm.acc_full, ps & 1u, 6, ps);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < NF; c0 += 8) {
        double v[24];
        if (has_rows) {
          const uint32_t ta = tm + ((uint32_t)(32 * ew) << 16) + (uint32_t)c0;
          double ysv[8], ycv[8];
          {
            uint32_t acc[NPL][8];
#pragma unroll
            for (int g = 0; g < NPL; ++g) tmem_ld8(ta + (uint32_t)(g * NBR), acc[g]);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              double y = (double)(int)acc[NPL - 1][q];  // smallest weight first
#pragma unroll
              for (int g = NPL - 2; g >= 0; --g) y = fma(y, 0.00390625, (double)(int)acc[g][q]);
              ysv[q] = y * rs;
            }
          }
          {
            uint32_t acc[NPL][8];
#pragma unroll
            for (int g = 0; g < NPL; ++g) tmem_ld8(ta + (uint32_t)(g * NBR + NF), acc[g]);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              double y = (double)(int)acc[NPL - 1][q];
#pragma unroll
              for (int g = NPL - 2; g >= 0; --g) y = fma(y, 0.00390625, (double)(int)acc[g][q]);
              ycv[q] = y * rs;
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const double ys = ysv[q], yc = ycv[q];
            if (row == pm.m) {  // the w row: (s|r) = s.w, (c|r) = c.w  
              sm.nval[2 * (c0 + q)] = ys;
              sm.nval[2 * (c0 + q) + 1] = yc;
            }
            const bool basis = row < pm.mfix;  // rows of the draw-independent block enter the b-sums (plain Fp: all)
            if (NMFP && row >= pm.mfix && row < pm.m && f0 + c0 + q < ar.F) {
              // rows of the per-draw block leave as z' in the layout nmfp_stageB_kernel streams: k-block (row / 4),
              // column block (4 frequencies), then 16 * {sin, cos} + 4 * (frequency % 4) + row % 4
              const int jr = row - pm.mfix + (ar.mvpad - pm.mvar), fi = c0 + q;
              double* z = ar.Z + ((size_t)p * ar.ntile + ft) * ((size_t)ar.mvpad * 64) +
                          (size_t)(((jr >> 2) * 8 + (fi >> 2)) * 32 + 4 * (fi & 3) + (jr & 3));
              z[0] = ys;
              z[16] = yc;
            }
            v[3 * q] = basis ? ys * ys : 0.0;
            v[3 * q + 1] = basis ? ys * yc : 0.0;
            v[3 * q + 2] = basis ? yc * yc : 0.0;
          }
          // transpose-reduce over the 32 rows of this warp: 24 -> 12 -> 6 -> 3 values per lane, then lanes 4q hold
          // the three sums of frequency c0 + q
#pragma unroll
          for (int i = 0; i < 12; ++i) {
            const bool up = lane & 16;
            const double keep = up ? v[i + 12] : v[i], send = up ? v[i] : v[i + 12];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const bool up = lane & 8;
            const double keep = up ? v[i + 6] : v[i], send = up ? v[i] : v[i + 6];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const bool up = lane & 4;
            const double keep = up ? v[i + 3] : v[i], send = up ? v[i] : v[i + 3];
            double t = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            v[i] = t;
          }
        } else {
          v[0] = v[1] = v[2] = 0.0;
        }
        if ((lane & 3) == 0) {
          const int q = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
          double* o = sm.part + ((size_t)ew * NF + c0 + q) * 3;
          if (grp == 0) { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
          else { o[0] += v[0]; o[1] += v[1]; o[2] += v[2]; }   // (this warp's own slots: no other writer)
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.acc_empty);  // tensor memory may be overwritten by the next pass
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // partial sums and the w-row values of the 4 warps are visible
      const uint32_t buf = it & 1u;
      if (ew == 0) {
        wait_wd<2000>(&sm.sums_full[buf], (it >> 1) & 1u, 7, it);
        const int f = lane;
        const int64_t fidx = f0 + f;
        double b[3] = {0, 0, 0}, a[3];
#pragma unroll
        for (int w2 = 0; w2 < 4; ++w2)
#pragma unroll
          for (int k2 = 0; k2 < 3; ++k2) b[k2] += sm.part[((size_t)w2 * NF + f) * 3 + k2];
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2) {
          a[k2] = sm.redA[((buf * 2 + 0) * NF + f) * 3 + k2];
          if (NG == 2) a[k2] += sm.redA[((buf * 2 + 1) * NF + f) * 3 + k2];
        }
        a[2] = pm.ninv_sum - a[0];  // c N^-1 c = sum 1/N - s N^-1 s (s^2 + c^2 = 1 to the last bit of the sincos values)
        const double N0 = sm.nval[2 * f], N1 = sm.nval[2 * f + 1];
        // M = [[ss, sc],[sc, cc]], N = [n0, n1]; LU with partial pivoting
        double m00 = a[0] - b[0], m01 = a[1] - b[1], m10 = m01, m11 = a[2] - b[2];
        double n0 = N0, n1 = N1;
        if (fabs(m10) > fabs(m00)) {  // row swap; the unknowns keep their order
          double t0 = m00; m00 = m10; m10 = t0;
          t0 = m01; m01 = m11; m11 = t0;
          t0 = n0; n0 = n1; n1 = t0;
        }
        const double lq = m10 / m00;
        const double u = m11 - lq * m01;
        const double x1 = (n1 - lq * n0) / u;
        const double x0 = (n0 - m01 * x1) / m00;
        double val = 0.5 * (N0 * x0 + N1 * x1);
        if (NMFP) {
          if (fidx < ar.F) {  // the draw-independent pieces (fixed block removed) for stage B
            double* o = ar.A + ((size_t)p * ar.ntile + ft) * 160 + f;
            o[0] = a[0] - b[0]; o[32] = a[1] - b[1]; o[64] = a[2] - b[2]; o[96] = N0; o[128] = N1;
          }
        } else if (fidx < ar.F) {
          if (!(ar.freqs[fidx] > 0.0)) val = __longlong_as_double(0x7ff8000000000000LL);  // f <= 0: NaN like f**(1/3)
          if (ar.terms) ar.terms[(size_t)p * ar.F + fidx] = val;
          if (ar.inner) {
            double* o = ar.inner + ((size_t)p * ar.F + fidx) * 5;
            o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; o[3] = N0; o[4] = N1;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.sums_empty[buf]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // part / nval are rewritten by the next item
    }
  } else {
    // ================= producers: sin/cos digit planes + the three quadratic sums =================
    if (NPW == 16) reg_set_dec<R::REGS_PROD>();
    else reg_set_inc<R::REGS_PROD>();
    const int pw = wid - 8;
    const uint32_t grp = (uint32_t)(pw >> 3);          // serves the stages with (global stage index % NG) == grp
    // lane = 4 * kg + fl: the 8 lanes of a quarter-warp read two distinct (t, 1/N) entries 64 bytes apart (no bank
    // conflict on the 16-byte loads), and a warp's 4-byte stores cover 4 rows x 32 bytes = all 32 banks
    const int kg = lane >> 2, fl = lane & 3;
    const int f = 4 * (pw & 7) + fl;                   // frequency inside the tile
    const int soff = swz32(f, 4 * kg);                 // word of TOAs 4kg..4kg+3 in row f (sin); cos row: + 1024
    // The schedule is built around one hardware fact: fp64 arithmetic and tcgen05 MMAs share a resource on the SM
    // (tools/probes/umma_fp64_overlap_probe.cu: back-to-back MMAs starve DFMAs; tools/i8_trace.py: next to each other
    // the fp64 part of a stage runs at a quarter of the fp64 rate and the MMAs at ~70 %), and a warp stalled on the fp64
    // pipe cannot issue its integer work either. Conversions, integer work and stores are free underneath the MMAs. So
    // a group announces stage k (s_full) not when its planes are stored but after the fp64 part of its NEXT stage: the
    // MMAs of stage k then start next to the integer part of stage k + NG (digits, byte transpose, stores) instead of
    // next to an fp64 part. Measured alternatives (announce at once, integer part interleaved into the next fp64 part,
    // strictly exclusive phases, batches and coarse phases of several stages): DESIGN.md section 4b.
    auto store_planes = [&](const double (&sv4)[4], const double (&cv4)[4], unsigned char* sb) {
      uint32_t slo[4], shi[4], clo[4], chi[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint2 ds = digits7(sv4[e]), dc = digits7(cv4[e]);
        slo[e] = ds.x; shi[e] = ds.y; clo[e] = dc.x; chi[e] = dc.y;
      }
      // word of plane p (most significant first) = byte (6 - p) of the four values
      {
        const uint32_t a = __byte_perm(slo[0], slo[1], 0x5140), b = __byte_perm(slo[0], slo[1], 0x7362);
        const uint32_t c2 = __byte_perm(slo[2], slo[3], 0x5140), d = __byte_perm(slo[2], slo[3], 0x7362);
        const uint32_t e2 = __byte_perm(shi[0], shi[1], 0x5140), f2 = __byte_perm(shi[0], shi[1], 0x7362);
        const uint32_t g2 = __byte_perm(shi[2], shi[3], 0x5140), h2 = __byte_perm(shi[2], shi[3], 0x7362);
        *reinterpret_cast<uint32_t*>(sb + 6 * S_PLANE) = __byte_perm(a, c2, 0x5410);   // byte 0
        *reinterpret_cast<uint32_t*>(sb + 5 * S_PLANE) = __byte_perm(a, c2, 0x7632);   // byte 1
        *reinterpret_cast<uint32_t*>(sb + 4 * S_PLANE) = __byte_perm(b, d, 0x5410);    // byte 2
        *reinterpret_cast<uint32_t*>(sb + 3 * S_PLANE) = __byte_perm(b, d, 0x7632);    // byte 3
        *reinterpret_cast<uint32_t*>(sb + 2 * S_PLANE) = __byte_perm(e2, g2, 0x5410);  // byte 4
        *reinterpret_cast<uint32_t*>(sb + 1 * S_PLANE) = __byte_perm(e2, g2, 0x7632);  // byte 5
        *reinterpret_cast<uint32_t*>(sb + 0 * S_PLANE) = __byte_perm(f2, h2, 0x5410);  // byte 6
      }
      {
        unsigned char* cb = sb + NF * KT;  // cos rows 32..63
        const uint32_t a = __byte_perm(clo[0], clo[1], 0x5140), b = __byte_perm(clo[0], clo[1], 0x7362);
        const uint32_t c2 = __byte_perm(clo[2], clo[3], 0x5140), d = __byte_perm(clo[2], clo[3], 0x7362);
        const uint32_t e2 = __byte_perm(chi[0], chi[1], 0x5140), f2 = __byte_perm(chi[0], chi[1], 0x7362);
        const uint32_t g2 = __byte_perm(chi[2], chi[3], 0x5140), h2 = __byte_perm(chi[2], chi[3], 0x7362);
        *reinterpret_cast<uint32_t*>(cb + 6 * S_PLANE) = __byte_perm(a, c2, 0x5410);
        *reinterpret_cast<uint32_t*>(cb + 5 * S_PLANE) = __byte_perm(a, c2, 0x7632);
        *reinterpret_cast<uint32_t*>(cb + 4 * S_PLANE) = __byte_perm(b, d, 0x5410);
        *reinterpret_cast<uint32_t*>(cb + 3 * S_PLANE) = __byte_perm(b, d, 0x7632);
        *reinterpret_cast<uint32_t*>(cb + 2 * S_PLANE) = __byte_perm(e2, g2, 0x5410);
        *reinterpret_cast<uint32_t*>(cb + 1 * S_PLANE) = __byte_perm(e2, g2, 0x7632);
        *reinterpret_cast<uint32_t*>(cb + 0 * S_PLANE) = __byte_perm(f2, h2, 0x5410);
      }
    };
    uint32_t kbase = 0, it = 0;
    int pend = -1;  // S slot whose planes are stored but not announced yet
    for (int item = blockIdx.x; item < ar.nwork; item += gridDim.x, ++it) {
      const int gp = item / ar.ntile, ft = item - gp * ar.ntile;
      const PulsarMeta pm = ar.meta[ar.pidx[gp]];
      const int64_t f0 = (int64_t)ft * NF;
      const double fq = ar.freqs[f0 + f < ar.F ? f0 + f : f0];  // a short last tile repeats its first frequency
      const double omega = __dmul_rn(6.283185307179586, fq);     // (2*pi)*f, rounded once (fastfp.py:78)
      const bool fast = __all_sync(0xffffffffu, fabs(omega) * pm.tabs_max <= 0.999 * FFP_SINCOS_MAX);
      double s3[2] = {0.0, 0.0};  // s N^-1 s, s N^-1 c
      const int nst = pm.i8_nst;
      // one pass over the TOAs per group of 128 operand rows (bases wider than 127 columns): the planes are produced
      // again for every pass, the quadratic sums only in the first
      for (int rg = 0; 128 * rg < pm.i8_rows; ++rg, kbase += (uint32_t)nst) {
      const int c0 = NG == 1 ? 0 : (int)((kbase ^ grp) & 1u);
      for (int c = c0; c < nst; c += NG) {
        const uint32_t k = kbase + (uint32_t)c;
        const uint32_t sv = k % VST, ss = k % SST;
        wait_wd<2000>(&sm.v_full[sv], (k / VST) & 1u, 8, k);
        const double2* vv = reinterpret_cast<const double2*>(sm.V + sv * V_STAGE) + 4 * kg;
        if (pw == 0 && lane == 0) FFP_TRACE(3, k);   // fp64 part of stage k starts (inputs present)
        // ---- fp64 part: four (TOA, frequency) pairs per thread
        double ph[4], ninv[4], sv4[4], cv4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double2 tn = vv[e];                    // (t, 1/N)
          ph[e] = __dmul_rn(omega, tn.x);              // ((2*pi)*f)*t, rounded once more
          ninv[e] = tn.y;
        }
        if (fast) {
          sincos_cw_n<4>(ph, sv4, cv4);  // in lockstep: phases inside the Cody-Waite range (checked once per item)
        } else {
          // cold: some phase of this item may exceed the Cody-Waite range (or is NaN/Inf): library sincos
#pragma unroll
          for (int e = 0; e < 4; ++e) {  // unrolled: a dynamic index would put the arrays in local memory
            double s1, c1;
            sincos(ph[e], &s1, &c1);
            sv4[e] = s1; cv4[e] = c1;
          }
        }
        if (rg == 0) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const double sn = sv4[e] * ninv[e];   // c N^-1 c follows from sum 1/N - s N^-1 s (epilogue)
            s3[0] = fma(sn, sv4[e], s3[0]);
            s3[1] = fma(sn, cv4[e], s3[1]);
          }
        }
        // ---- the group's previous stage is complete in shared memory: announce it now (see above)
        __syncwarp();
        if (lane == 0) {
          if (pw == 0) FFP_TRACE(4, k);              // fp64 part of stage k done
          mbar_arrive(&sm.v_empty[sv]);
          if (pend >= 0) mbar_arrive(&sm.s_full[pend]);
        }
        // ---- integer part: digits, 4 x 7 byte transpose, stores
        if (k >= SST) wait_wd<2000>(&sm.s_empty[ss], ((k / SST) - 1) & 1u, 9, k);
        store_planes(sv4, cv4, sm.S + ss * S_STAGE + soff);
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's (async proxy) operand reads
        if (pw == 0 && lane == 0) FFP_TRACE(5, k);   // planes of stage k stored
        pend = (int)ss;
      }
      }
      // the two sums of frequency f: over the 8 lanes that share it (lane bits 2..4), then published per group
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        double t = s3[q];
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        t += __shfl_xor_sync(0xffffffffu, t, 8);
        t += __shfl_xor_sync(0xffffffffu, t, 16);
        s3[q] = t;
      }
      const uint32_t buf = it & 1u;
      if (it >= 2) wait_wd<2000>(&sm.sums_empty[buf], ((it >> 1) - 1) & 1u, 10, it);
      if (kg == 0) {
        double* o = sm.redA + ((buf * 2 + grp) * NF + f) * 3;
        o[0] = s3[0]; o[1] = s3[1];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.sums_full[buf]);
    }
    __syncwarp();
    if (lane == 0 && pend >= 0) mbar_arrive(&sm.s_full[pend]);  // the group's last stage of this CTA
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "n"(512));
}

// ---- measurement helper: the tensor path's own ceilings (fastfp_fp64_peak kinds 17 and 18) -----------------------
// Back-to-back kind::i8 MMAs from one thread per CTA, one CTA per SM, operands in the sweep kernel's SWIZZLE_32B planes:
// N = 256 gives the INT8 tensor peak of the chip, N = 64 with the sweep's 28-product stage gives the rate this
// formulation can reach at most (an M=128, N=64 MMA is bound by its 6 KB of shared-memory operand reads).
template <int N>
__global__ void __launch_bounds__(64, 1) i8_peak_kernel(int iters, int* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  constexpr int A_PLANE = 128 * KT, B_PLANE = N * KT;
  unsigned char* sA = smem;
  unsigned char* sB = smem + NPL * A_PLANE;
  for (int i = threadIdx.x; i < NPL * (A_PLANE + B_PLANE); i += blockDim.x) smem[i] = (unsigned char)((i * 7 + 3) & 3);
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < NPL; ++i)
#pragma unroll
        for (int j = 0; j < NPL - i; ++j)
          umma_i8(tm + (uint32_t)(((i + j) * N) % 512), umma_desc(smem_u32(sA + i * A_PLANE)),
                  umma_desc(smem_u32(sB + j * B_PLANE)), idesc, (it > 0 || i > 0) ? 1u : 0u);
    }
    umma_commit(&bar);
    wait_wd<0>(&bar, 0u, 11, 0u);
    if (out) out[blockIdx.x] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "n"(512));
}

}  // namespace i8

int run_i8_peak(int kind, int iters, double* tops, double* ms_out) {
  using namespace i8;
  int dev = 0, sms = 0;
  FFP_CUDA(cudaGetDevice(&dev));
  FFP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int N = kind == 17 ? 256 : 64;
  const size_t smem = (size_t)NPL * (128 + N) * KT + 1024;
  if (kind == 17) FFP_CUDA(cudaFuncSetAttribute(i8_peak_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else FFP_CUDA(cudaFuncSetAttribute(i8_peak_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  FFP_CUDA(cudaEventCreate(&e0));
  FFP_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    FFP_CUDA(cudaEventRecord(e0));
    if (kind == 17) i8_peak_kernel<256><<<sms, 64, smem>>>(iters, nullptr);
    else i8_peak_kernel<64><<<sms, 64, smem>>>(iters, nullptr);
    FFP_CUDA(cudaEventRecord(e1));
    FFP_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    FFP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  g_launches += 4;
  FFP_CUDA(cudaGetLastError());
  *tops = 2.0 * (double)sms * iters * 28.0 * 128.0 * N * 32.0 / (best * 1e-3) / 1e12;
  *ms_out = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

// ---- host side ------------------------------------------------------------------------------------------------
// A pack can take the tensor path when every pulsar fits the tile: diagonal N, m + 1 <= 128 rows (the basis rows and
// the w row), n <= 16384 TOAs (int32 accumulators: 7 products of |digit|^2 <= 2^14 per TOA stay below 2^31
// up to 18 724 TOAs).
// which pulsars the tensor sweep can take: diagonal N (pack-wide), up to RS operand rows, n <= 16384 (exactness of the
// int32 accumulators); the others of a pack stay on the fp64 kernel in the same sweep
static bool i8_takes(const fastfp_pack* pk, const PulsarMeta& pm) {
  return !pk->ecorr && pm.m + 1 <= i8::RS && pm.n <= 16384;
}
bool i8_eligible(const fastfp_pack* pk) {
  for (const PulsarMeta& pm : pk->meta)
    if (i8_takes(pk, pm)) return true;
  return false;
}

// lay out and build the digit planes from the fp64 packets (after launch_fp_precompute); sets pk->i8_ok
int build_i8_planes(fastfp_pack* pk, cudaStream_t st) {
  pk->i8_ok = false;
  pk->i8_count = 0;
  if (!i8_eligible(pk)) return 0;
  const int P = pk->P;
  int64_t off = 0;
  int rows_max = 0, nst_max = 0;
  for (PulsarMeta& pm : pk->meta) {
    pm.i8_rows = pm.i8_nst = 0;
    pm.i8_off = off;
    if (!i8_takes(pk, pm)) continue;
    // rows (basis + the w row; more than 128 of them form row groups, one pass over the TOAs each) padded to 32:
    // every plane then starts on a 1024-byte boundary of the (1024-aligned) ring, the alignment
    // the operand descriptors of all swizzle modes accept; rows beyond the padding are never loaded (the MMA reads
    // 128 rows per plane, the tail comes from the next plane and lands in output lanes nobody reads)
    pm.i8_rows = (pm.m + 1 + 31) / 32 * 32;
    pm.i8_nst = (pm.n + i8::KT - 1) / i8::KT;
    off += (int64_t)pm.i8_nst * (i8::V_STAGE + i8::NPL * pm.i8_rows * i8::KT);
    rows_max = pm.i8_rows > rows_max ? pm.i8_rows : rows_max;
    nst_max = pm.i8_nst > nst_max ? pm.i8_nst : nst_max;
  }
  FFP_CUDA(cudaMemcpyAsync(pk->d_meta, pk->meta.data(), sizeof(PulsarMeta) * P, cudaMemcpyHostToDevice, st));
  // the last plane of the last stage is read 128 rows deep by the MMA only from shared memory; the global buffer
  // needs no slack, but keep the allocation 16-byte granular for the bulk copies
  FFP_CUDA(cudaMalloc(&pk->d_i8, (size_t)off + 16));
  FFP_CUDA(cudaMalloc(&pk->d_i8_scale, (size_t)P * i8::RS * sizeof(double)));
  FFP_CUDA(cudaMemsetAsync(pk->d_i8_scale, 0, (size_t)P * i8::RS * sizeof(double), st));
  int *d_exp = nullptr, *d_bad = nullptr;
  FFP_CUDA(cudaMalloc(&d_exp, (size_t)P * i8::RS * sizeof(int)));
  FFP_CUDA(cudaMalloc(&d_bad, (size_t)P * sizeof(int)));
  FFP_CUDA(cudaMemsetAsync(d_bad, 0, (size_t)P * sizeof(int), st));
  i8::i8_rowscale_kernel<<<dim3(rows_max, P), 256, 0, st>>>(pk->d_packets, pk->d_meta, pk->d_i8_scale, d_exp, d_bad);
  i8::i8_planes_kernel<<<dim3(nst_max, P), 256, 0, st>>>(pk->d_packets, pk->d_meta, d_exp, pk->d_i8);
  g_launches += 2;
  cudaError_t e = cudaGetLastError();
  std::vector<int> bad(P, 0);
  if (e == cudaSuccess) e = cudaMemcpyAsync(bad.data(), d_bad, (size_t)P * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_exp);
  cudaFree(d_bad);
  if (e != cudaSuccess) return cuda_fail(e, "build_i8_planes");
  // pulsars with a non-finite G or w (singular Sigma, NaN data) cannot be carried by integer planes: fp64 kernel
  std::vector<int> take, rest_flag(P, 1);
  for (int p = 0; p < P; ++p) {
    const bool ok = pk->meta[p].i8_nst > 0 && bad[p] == 0 && (p >= (int)pk->info.size() || pk->info[p] == 0);
    if (ok) { take.push_back(p); rest_flag[p] = 0; }
  }
  pk->i8_rows_max = rows_max;
  pk->i8_bytes = off;
  if (take.empty()) {
    cudaFree(pk->d_i8); cudaFree(pk->d_i8_scale);
    pk->d_i8 = nullptr; pk->d_i8_scale = nullptr;
    return 0;
  }
  cudaFree(pk->d_pidx_all);
  pk->d_pidx_all = nullptr;
  FFP_CUDA(cudaMalloc(&pk->d_pidx_all, sizeof(int) * take.size()));
  FFP_CUDA(cudaMemcpy(pk->d_pidx_all, take.data(), sizeof(int) * take.size(), cudaMemcpyHostToDevice));
  pk->i8_count = (int)take.size();
  // the complement, per kernel family of the fp64 sweep
  for (Group& g : pk->groups) {
    std::vector<int> all((size_t)g.count), rest;
    FFP_CUDA(cudaMemcpy(all.data(), g.d_pidx, sizeof(int) * g.count, cudaMemcpyDeviceToHost));
    for (int p : all)
      if (rest_flag[p]) rest.push_back(p);
    cudaFree(g.d_pidx_rest);
    g.d_pidx_rest = nullptr;
    g.count_rest = (int)rest.size();
    if (g.count_rest) {
      FFP_CUDA(cudaMalloc(&g.d_pidx_rest, sizeof(int) * rest.size()));
      FFP_CUDA(cudaMemcpy(g.d_pidx_rest, rest.data(), sizeof(int) * rest.size(), cudaMemcpyHostToDevice));
    }
  }
  pk->bytes += off + (int64_t)P * i8::RS * 8;
  pk->i8_ok = true;
  return 0;
}

int launch_fp_sweep_i8(const fastfp_pack* pk, const double* d_freqs, int64_t F, double* d_terms, cudaStream_t st,
                       double* d_inner, const NmfpOut* nm) {
  using namespace i8;
  Args a{};
  a.planes = pk->d_i8;
  a.rowscale = pk->d_i8_scale;
  a.meta = pk->d_meta;
  a.pidx = pk->d_pidx_all;
  a.freqs = d_freqs;
  a.F = F;
  a.terms = d_terms;
  a.inner = d_inner;
  a.Z = nm ? nm->Z : nullptr;
  a.A = nm ? nm->A : nullptr;
  a.mvpad = nm ? nm->mvmax : 0;
  const int64_t ntile = (F + NF - 1) / NF, nwork = ntile * pk->i8_count;  // the pulsars this kernel takes
  if (nwork > 0x7fffffffLL) { set_error("frequency batch too large for one launch"); return -1; }
  a.ntile = (int)ntile;
  a.nwork = (int)nwork;
  a.gslot = NPL * (pk->i8_rows_max < 128 ? pk->i8_rows_max : 128) * KT;  // one row group
  const size_t budget = 220 * 1024 - SMEM_FIXED;
  int gst = (int)(budget / a.gslot);
  gst = gst > 8 ? 8 : gst;
  // the last plane of the last slot is read 128 rows deep: the S ring behind the G ring absorbs the overrun
  if (gst < 2) { set_error("internal: no room for the G ring"); return FASTFP_ERR_UNSUPPORTED; }
  a.gst = gst;
  const size_t smem = (size_t)gst * a.gslot + SMEM_FIXED;
  static bool attr_done[64] = {};
  if (!attr_done[pk->device & 63]) {
#define FFP_I8_ATTR(NPW_)                                                                                             \
  FFP_CUDA(cudaFuncSetAttribute(fp_sweep_i8_kernel<false, NPW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
  FFP_CUDA(cudaFuncSetAttribute(fp_sweep_i8_kernel<true, NPW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FFP_I8_ATTR(16) FFP_I8_ATTR(8)
#undef FFP_I8_ATTR
    attr_done[pk->device & 63] = true;
  }
  const unsigned grid = (unsigned)(nwork < pk->num_sms ? nwork : pk->num_sms);
  // producer warps per CTA: 16 (two groups alternating stages; default, C2 18.0 ms) or 8 (FASTFP_B200_I8_NPW=8: 19.2 ms)
#ifdef FFP_I8_TRACE
  long long* d_trace = nullptr;
  const char* trace_path = getenv("FASTFP_B200_I8_TRACE");
  if (trace_path) {
    FFP_CUDA(cudaMalloc(&d_trace, sizeof(long long) * TRACE_EV * TRACE_K));
    FFP_CUDA(cudaMemsetAsync(d_trace, 0, sizeof(long long) * TRACE_EV * TRACE_K, st));
  }
  a.trace = d_trace;
#endif
  static const int npw = getenv("FASTFP_B200_I8_NPW") ? atoi(getenv("FASTFP_B200_I8_NPW")) : 16;
  if (npw == 16) {
    if (nm) fp_sweep_i8_kernel<true, 16><<<grid, Roles<16>::THREADS, smem, st>>>(a);
    else fp_sweep_i8_kernel<false, 16><<<grid, Roles<16>::THREADS, smem, st>>>(a);
  } else {
    if (nm) fp_sweep_i8_kernel<true, 8><<<grid, Roles<8>::THREADS, smem, st>>>(a);
    else fp_sweep_i8_kernel<false, 8><<<grid, Roles<8>::THREADS, smem, st>>>(a);
  }
  g_launches += 1;
  FFP_CUDA(cudaGetLastError());
#ifdef FFP_I8_TRACE
  if (d_trace) {  // diagnostic build only: synchronous dump of CTA 0's event clocks
    std::vector<long long> h((size_t)TRACE_EV * TRACE_K);
    FFP_CUDA(cudaStreamSynchronize(st));
    FFP_CUDA(cudaMemcpy(h.data(), d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    FFP_CUDA(cudaFree(d_trace));
    if (FILE* fh = fopen(trace_path, "w")) {
      for (int k = 0; k < TRACE_K; ++k) {
        for (int e = 0; e < TRACE_EV; ++e) fprintf(fh, "%lld ", h[(size_t)e * TRACE_K + k]);
        fprintf(fh, "\n");
      }
      fclose(fh);
    }
  }
#endif
  return 0;
}

}  // namespace ffp
