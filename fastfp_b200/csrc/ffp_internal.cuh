// Internal declarations shared by the CUDA translation units of libfastfp_b200.so.
// sm_100a only (built with -gencode arch=compute_100a,code=sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

namespace ffp {

// ---- tiling constants of the sweep kernel -------------------------
constexpr int NWC = 8;         // consumer (MMA) warps per sweep CTA: two per SM sub-partition
constexpr int NWP = 16;        // producer (sincos) warps per sweep CTA: four per SM sub-partition
constexpr int NTC = NWC * 32, NTP = NWP * 32;
constexpr int NTHREADS = NTC + NTP;  // 768 threads per sweep CTA, one CTA per SM
constexpr int CTAS_PER_SM = 1;
constexpr int CONSUMER_REGS = 120, PRODUCER_REGS = 56;  // setmaxnreg split of the 768 x 80 pool
constexpr int VST = 16;        // TOA-vector ring depth (t | 1/N | w; TMA -> producer)
constexpr int FLUSH_TOAS = 512;  // level-1 accumulation block, in TOAs
constexpr int MAX_M = 640;     // widest basis the sweep kernel handles (8 warp rows x 10 blocks of 8 rows)

// Sweep configuration. The contraction runs on the fp64 MMA path (mma.sync.m8n8k4.f64): a warp
// owns NMBW row blocks (8 basis rows each) x NNB column blocks (8 columns = 4 frequencies x
// {sin, cos}); WMW consumer warps split the rows, NWC/WMW split the frequencies of the tile.
//   CI  TOAs per staged chunk (KB = CI/4 k-blocks)
//   NWC / NWP  consumer (MMA) / producer (sincos) warps of the CTA. The MMA issue rate of one warp is
//   limited, so narrow accumulator tiles want three consumer warps per SM sub-partition (12 + 8);
//   the default is two per sub-partition and four producer warps per sub-partition (8 + 16).
template <int NMBW_, int NNB_, int WMW_, int CI_, int NWC_ = ffp::NWC, int NWP_ = ffp::NWP>
struct SweepCfg {
  static constexpr int NMBW = NMBW_, NNB = NNB_, WMW = WMW_, CI = CI_;
  static constexpr int NWC = NWC_, NWP = NWP_, NTC = NWC_ * 32, NTP = NWP_ * 32, NTHREADS = NTC + NTP;
  // setmaxnreg split of the launch-time register pool (NTHREADS x the launch register count)
  static constexpr int CREGS = NWC_ == 8 ? CONSUMER_REGS : 112, PREGS = NWC_ == 8 ? PRODUCER_REGS : 64;
  static_assert(NTHREADS <= 1024 && NWC % WMW_ == 0, "warp layout");
  static_assert(NTC * CREGS + NTP * PREGS <= NTHREADS * ((65536 / NTHREADS) / 8 * 8), "register pool");
  static constexpr int WNW = NWC / WMW;        // consumer warps along frequency
  static constexpr int KF = WNW * NNB * 4;     // frequencies per CTA
  static constexpr int NBT = KF / 4;           // column blocks per CTA tile
  static constexpr int NMB = NMBW * WMW;       // row blocks
  static constexpr int MP = 8 * NMB;           // padded basis width (rows of G)
  static constexpr int KB = CI / 4;            // k-blocks (4 TOAs) per chunk
  static constexpr int ST = CI * 2 * KF;       // doubles of one sin/cos tile: [KB][NBT][32]
  static constexpr int VEC = 4 * CI;           // doubles of the vector part: (t, 1/N, w, 0) per TOA
  static constexpr int GT = CI * MP;           // doubles of the G part: [KB][NMB][32] fragments
  static constexpr int PK = VEC + GT;          // doubles per packet
  // basis phase: one warp store covers 8 frequencies x 4 TOAs; a thread keeps XW frequencies
  static constexpr int NX = KF / 8;                       // groups of 8 frequencies
  static constexpr int XW = NX >= NWP ? NX / NWP : 1;      // frequency groups per producer warp
  // producer warps sharing one group split its k-blocks; with few groups and few k-blocks the
  // surplus warps stay idle (they still take part in the barrier protocol)
  static constexpr int KSPLIT = NX >= NWP ? 1 : (NWP / NX < KB ? NWP / NX : KB);
  static constexpr int KBW = KB / KSPLIT;                 // k-blocks per active producer warp
  static constexpr int NACTIVE = NX >= NWP ? NWP : NX * KSPLIT;  // producer warps with work
  static constexpr int NACC = 2 * NMBW * NNB;  // accumulators per thread
  static constexpr int NACCX = NACC + 3 * NNB;  // + epoch sums of the block-N variant
  static constexpr int SLAB = NACCX * NTC + 5 * XW * NTP;  // doubles of level-2 scratch per CTA
  static constexpr int FLUSH = FLUSH_TOAS / CI;  // chunks per level-1 block
  // ring depths: sin/cos tiles (producer -> consumer) and G tiles (TMA -> consumer), as deep as
  // the shared-memory budget allows
  static constexpr int SST = ST * 8 * 8 <= 140 * 1024 ? 8 : 4;
  static constexpr int GBUDGET = 214 * 1024 - SST * ST * 8 - VST * VEC * 8;
  static constexpr int GST = GBUDGET / (GT * 8) >= 8 ? 8 : (GBUDGET / (GT * 8) >= 2 ? GBUDGET / (GT * 8) : 2);
  static constexpr int RED = KF * (3 * WMW + 5 * KSPLIT);  // doubles of epilogue reduction scratch
  static constexpr size_t SMEM =
      (size_t)(SST * ST + GST * GT + VST * VEC + KF + RED + 2 * (SST + GST + VST)) * 8 + 128;
  static_assert(KB % KSPLIT == 0 && NACTIVE <= NWP, "basis-phase mapping");
  static_assert(GST >= 3, "the consumers prefetch across chunk boundaries: three G stages at least");
};

// offset of G element (TOA il within its chunk, basis row j) inside a packet's G part:
// fragment order [k-block][row block][lane], lane = (j%8)*4 + il%4 -- the A-operand layout of
// mma.m8n8k4 (row = lane>>2, k = lane&3), so a warp's fragment load is 256 contiguous bytes.
__host__ __device__ inline int g_frag_index(int il, int j, int nmb) {
  return (((il >> 2) * nmb + (j >> 3)) << 5) + (((j & 7) << 2) | (il & 3));
}

// Per-pulsar descriptor, device-visible.
struct PulsarMeta {
  int64_t pk_off;  // offset (in doubles) of the pulsar's first packet inside `packets`
  int64_t L_off;   // offset of its m x m factor inside `Lbuf`
  int64_t raw_off; // offset of its TOA-length vectors inside the staging arrays
  int64_t T_off;   // offset of its raw T inside the staging array
  int32_t n, m, nch, mpad;
  int32_t mfix, mvar;  // nmfp: leading draw-independent columns / trailing per-draw columns
  int32_t var_off;     // nmfp: offset of this pulsar's varying block in a phiinv_var row
  int32_t ci;          // TOAs per packet for this pulsar's kernel configuration
  double tabs_max;     // max |TOA| (inf if any TOA is not finite): decides the sincos path per tile
  int64_t dm_off;      // block-N packs: offset of this pulsar's per-chunk slot masks
  int64_t i8_off;      // tensor path: byte offset of this pulsar's digit-plane stages
  int32_t i8_rows;     // tensor path: rows stored per plane (basis rows + the w row, padded to 8)
  int32_t i8_nst;      // tensor path: stages of 32 TOAs
  double ninv_sum;     // tensor path: sum_i 1/N_i (c N^-1 c = sum 1/N - s N^-1 s, so the producers form two sums, not three)
};

struct KernelCfg {  // run-time mirror of SweepCfg's parameters
  int nmbw, nnb, wmw, ci;
  int nwc = NWC;  // consumer warps (8, or 12 for the three-row-group configurations)
  int kf() const { return (nwc / wmw) * nnb * 4; }
  int mp() const { return 8 * nmbw * wmw; }
  bool operator<(const KernelCfg& o) const {
    if (nmbw != o.nmbw) return nmbw < o.nmbw;
    if (nnb != o.nnb) return nnb < o.nnb;
    if (wmw != o.wmw) return wmw < o.wmw;
    if (nwc != o.nwc) return nwc < o.nwc;
    return ci < o.ci;
  }
};

struct Group {  // pulsars that share one kernel instantiation
  KernelCfg cfg;
  int count;
  int* d_pidx;  // device array of pulsar indices
  int count_rest = 0;          // ... those of them the tensor sweep does not take (build_i8_planes), when it takes some
  int* d_pidx_rest = nullptr;
};

}  // namespace ffp

// The opaque handle of include/fastfp_b200.h.
struct fastfp_pack {
  int device = 0;
  int P = 0;
  int num_sms = 0;
  bool nmfp = false;
  bool ecorr = false;            // block-diagonal N (kernel ECORR): 8 epoch-slot rows in the G tiles
  unsigned char* d_done_mask = nullptr;
  std::vector<ffp::PulsarMeta> meta;
  std::vector<ffp::Group> groups;
  ffp::PulsarMeta* d_meta = nullptr;
  double* d_packets = nullptr;  // [sum_p nch_p][PK_p]
  double* d_L = nullptr;        // Cholesky factors (Fp) / fixed-block factors (nmfp)
  int* d_info = nullptr;        // per-pulsar factorisation status
  std::vector<int> info;        // host copy, read back when the pack is built (fastfp_pack_factor_info)
  double* d_slab = nullptr;     // level-2 accumulation scratch, one slab per resident CTA
  unsigned int* d_counter = nullptr;  // persistent-CTA work counter
  // INT8 tensor-core path (fp_sweep_i8.cu): digit planes of G, per-row scales; chosen per pack
  unsigned char* d_i8 = nullptr;
  double* d_i8_scale = nullptr;   // [P][128]
  int* d_pidx_all = nullptr;      // pulsars on the tensor sweep (all of them, or those that fit: n <= 16384, finite planes)
  int i8_count = 0;               // their number; the others stay on the fp64 kernel in the same sweep
  bool i8_ok = false;             // the planes exist (every pulsar fits the tile, all values finite)
  int i8_rows_max = 0;
  int64_t i8_bytes = 0;
  int path = 0;                   // FASTFP_PATH_AUTO / _FP64 / _I8 (fastfp_pack_set_path)
  // AUTO resolves to the tensor path wherever the pack can take it (it passed the GPU parity suite on hardware and is
  // the faster kernel: profiles/README.md); FASTFP_PATH_FP64 forces the DMMA kernel
  static constexpr bool kAutoPrefersI8 = true;
  bool use_i8() const { return i8_ok && (path == 2 || (path == 0 && kAutoPrefersI8)); }
  bool i8_all() const { return i8_count == P; }
  int64_t bytes = 0;
  int64_t mvar_total = 0;
  int mvar_max = 0;
  int mvpad = 0;           // nmfp: per-draw block width padded to the stage-B tile (32, 64 or 96)
  // nmfp only
  double* d_S0 = nullptr;  // [P][mvmax][mvmax] Schur complement of the fixed block (no phiinv)
  double* d_zr = nullptr;  // [P][mvmax]  z'_r
  // scratch reused across sweeps (grown on demand)
  mutable double* d_terms = nullptr;
  mutable int64_t terms_cap = 0;
  mutable double* d_freqs = nullptr;
  mutable int64_t freqs_cap = 0;
  mutable double* d_out = nullptr;
  mutable int64_t out_cap = 0;
  mutable double* d_scratch = nullptr;  // nmfp: stage-A tiles of a frequency batch
  mutable int64_t scratch_cap = 0;
  mutable double* d_lf = nullptr;       // nmfp: L^-1 fragments of a draw batch
  mutable int64_t lf_cap = 0;
  mutable double* d_inner = nullptr;   // Fe-statistic: inner products of a frequency batch + antenna patterns
  mutable int64_t inner_cap = 0;
  // staging of the per-draw power-law parameters (fastfp_powerlaw_phiinv): device + pinned host copy,
  // the event marks the last H2D copy out of h_pl; pl_tab is the frequency table already on the device
  mutable double* d_pl = nullptr;
  mutable double* h_pl = nullptr;
  mutable int64_t pl_cap = 0;
  mutable cudaEvent_t pl_event = nullptr;
  mutable std::vector<double> pl_tab;
  // optional per-stage timing of nmfp sweeps (fastfp_nmfp_stage_timing): stage A, factor, stage B
  mutable bool time_stages = false;
  mutable double stage_ms[3] = {0.0, 0.0, 0.0};
};

namespace ffp {

extern std::atomic<int64_t> g_launches;
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define FFP_CUDA(call)                                         \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return ffp::cuda_fail(e__, #call); \
  } while (0)

// grow-on-demand device scratch (contents are not preserved)
template <typename T>
inline int ensure(T** buf, int64_t* cap, int64_t need) {
  if (*cap >= need) return 0;
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *cap = 0;
  FFP_CUDA(cudaMalloc(buf, (size_t)need * sizeof(T)));
  *cap = need;
  return 0;
}

// ---- kernel launchers (defined in the .cu files) ---------------------------------------
// precompute.cu
struct BlockNDev {         // device copies of the block-N (kernel ECORR) side arrays, or all null
  const double* res_w;     // (N^-1 r) * Nvec per TOA
  const int* slot_idx;     // epoch slot (0..7) of a TOA inside its chunk, -1 if none
  const double* slot_val;  // sqrt(beta_e) / Nvec_i
};
int launch_fp_precompute(fastfp_pack* pk, const double* d_toas, const double* d_res,
                         const double* d_Nvec, const double* d_T, cudaStream_t st,
                         double* d_ur_keep = nullptr,  // [P][MAX_M], receives G r
                         const BlockNDev* bn = nullptr);
// fp_sweep*.cu
struct NmfpOut {      // stage-A outputs of the nmfp path (null for plain Fp)
  double* Z;          // [P][ceil(F/32)][mvpad/4][8][32]  z'_s, z'_c tiles in MMA B-fragment order
  double* A;          // [P][ceil(F/32)][5][32]           a_ss, a_sc, a_cc, a_sr, a_cr
  int mvmax;          // padded width of the per-draw block (multiple of 8)
};
int launch_fp_sweep(const fastfp_pack* pk, const double* d_freqs, int64_t F, double* d_terms,
                    cudaStream_t st, const NmfpOut* nm = nullptr, double* d_inner = nullptr, bool rest_only = false);
// the sweep on the path(s) the pack is set to: the tensor kernel for the pulsars it takes, the fp64 kernel for the rest
int launch_sweep(const fastfp_pack* pk, const double* d_freqs, int64_t F, double* d_terms, cudaStream_t st,
                 const NmfpOut* nm = nullptr, double* d_inner = nullptr);
int launch_reduce_terms(const double* d_terms, int P, int64_t F, double* d_out, cudaStream_t st);
// fp_sweep_i8.cu
bool i8_eligible(const fastfp_pack* pk);
int build_i8_planes(fastfp_pack* pk, cudaStream_t st);
int run_i8_peak(int kind, int iters, double* tops, double* ms);
int launch_fp_sweep_i8(const fastfp_pack* pk, const double* d_freqs, int64_t F, double* d_terms, cudaStream_t st,
                       double* d_inner = nullptr, const NmfpOut* nm = nullptr);
// fe.cu
int launch_fe_combine(const double* d_inner, int P, int64_t F, const double* d_fplus, const double* d_fcross, int64_t S,
                      double* d_out, int64_t out_ld, cudaStream_t st);
bool sweep_config(int m, KernelCfg* cfg);
int sweep_max_slab_doubles();
// xcy.cu
int launch_tnt(int64_t n, int64_t m, const double* dN, const double* dT, const double* d_phiinv, double* d_part,
               int nsplit, double* d_out, cudaStream_t st);
int launch_xcy(int64_t n, int64_t m, const double* dN, const double* dT, const double* dS,
               const double* dx, const double* dy, const double* dx0, double* d_work, double* d_out,
               cudaStream_t st);  // dx0: raw x for the x^T N^-1 y term (null: dx)
// microbench.cu
int run_fp64_peak(int kind, int iters, double* tflops, double* ms);

// ---- small PTX helpers: mbarrier + 1-D TMA bulk copy ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.b32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  while (!mbar_try_wait(bar, parity)) __nanosleep(64);  // back off: polling shares the MIO queue with LDS
}
// for the warps on the critical path (the MMA consumers): try_wait already suspends the warp for a
// hardware-chosen interval and wakes it when the phase completes, so no extra sleep that could overshoot
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace ffp
