// extern "C" entry points of libfastfp_b200.so (declared in include/fastfp_b200.h).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "../../include/fastfp_b200.h"
#include "ffp_internal.cuh"

namespace ffp {

std::atomic<int64_t> g_launches{0};
static thread_local std::string t_err;

void set_error(const std::string& msg) { t_err = msg; }
int cuda_fail(cudaError_t e, const char* what) {
  t_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return FASTFP_ERR_CUDA;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};


// nmfp.cu
int nmfp_pack_finish(fastfp_pack* pk, const double* d_toas, const double* d_res,
                     const double* d_Nvec, const double* d_T, const double* d_TNT,
                     const double* d_phiinv_fix, cudaStream_t st, const BlockNDev* bn = nullptr);
int nmfp_sweep_impl(const fastfp_pack* pk, const double* d_freqs, int64_t F,
                    const double* d_phiinv_var, int64_t D, double* d_out, cudaStream_t st);
int nmfp_stage_a_impl(const fastfp_pack* pk, const double* d_freqs, int64_t F, double* dZ, double* dA, cudaStream_t st);
int nmfp_stage_b_only(const fastfp_pack* pk, const double* d_freqs, int64_t F, const double* dZ, const double* dA,
                      int nt_blk, const double* d_phiinv_var, int64_t D, double* d_out, cudaStream_t st);
int powerlaw_phiinv_impl(const fastfp_pack* pk, const double* const* Ffreqs, const double* log10_A,
                         const double* gamma, int64_t D, const double* curn_Ffreqs, int64_t ncurn,
                         const double* curn_log10_A, const double* curn_gamma, double* out,
                         cudaStream_t st);

// Common part of the two pack constructors: validate, lay out, upload the raw arrays.
struct Staging {
  double *d_toas = nullptr, *d_res = nullptr, *d_Nvec = nullptr, *d_T = nullptr;
  ~Staging() {
    cudaFree(d_toas); cudaFree(d_res); cudaFree(d_Nvec); cudaFree(d_T);
  }
};

static int pack_layout(fastfp_pack* pk, int P, const int64_t* n, const int64_t* m,
                       const int64_t* m_fix, const double* const* toas, bool blockn = false) {
  pk->P = P;
  pk->meta.resize(P);
  int64_t pk_off = 0, L_off = 0, raw_off = 0, T_off = 0, dm_off = 0;
  int var_off = 0;
  pk->ecorr = blockn;
  std::map<KernelCfg, std::vector<int>> groups;
  for (int p = 0; p < P; ++p) {
    if (n[p] < 1 || m[p] < 1 || n[p] > 0x7fffff00LL) {
      set_error("pulsar " + std::to_string(p) + ": n and m must be positive");
      return FASTFP_ERR_INVALID;
    }
    KernelCfg kc{};
    // block-diagonal N: one more block of 8 rows (the epoch slots) after the basis rows
    const int m_rows = blockn ? ((int)m[p] + 7) / 8 * 8 + 8 : (int)m[p];
    if (m[p] > MAX_M || !sweep_config(m_rows, &kc)) {
      set_error("pulsar " + std::to_string(p) + ": basis width m=" + std::to_string(m[p]) +
                " exceeds the supported maximum " + std::to_string(MAX_M));
      return FASTFP_ERR_UNSUPPORTED;
    }
    PulsarMeta& pm = pk->meta[p];
    pm.n = (int)n[p];
    pm.m = (int)m[p];
    pm.ci = kc.ci;
    pm.nch = (int)((n[p] + kc.ci - 1) / kc.ci);
    pm.mpad = kc.mp();
    pm.pk_off = pk_off;
    pm.L_off = L_off;
    pm.raw_off = raw_off;
    pm.T_off = T_off;
    pm.mfix = m_fix ? (int)m_fix[p] : pm.m;
    pm.mvar = pm.m - pm.mfix;
    if (pm.mfix < 0 || pm.mvar < 0) {
      set_error("pulsar " + std::to_string(p) + ": m_fix out of range");
      return FASTFP_ERR_INVALID;
    }
    pm.var_off = var_off;
    pm.dm_off = dm_off;
    if (blockn && n[p] % kc.ci != 0) {
      set_error("block-N pack: the TOA count must be a multiple of the chunk size (fastfp_sweep_chunk_toas)");
      return FASTFP_ERR_INVALID;
    }
    dm_off += (n[p] + kc.ci - 1) / kc.ci;
    pm.tabs_max = 0.0;
    if (!toas[p]) { set_error("null toas"); return FASTFP_ERR_INVALID; }
    for (int64_t i = 0; i < n[p]; ++i) {
      const double a = std::fabs(toas[p][i]);
      if (!(a <= 1.7e308)) { pm.tabs_max = INFINITY; break; }
      if (a > pm.tabs_max) pm.tabs_max = a;
    }
    var_off += pm.mvar;
    pk->mvar_max = std::max(pk->mvar_max, pm.mvar);
    pk_off += (int64_t)pm.nch * pm.ci * (4 + pm.mpad);
    L_off += (int64_t)pm.m * pm.m;
    raw_off += pm.n;
    T_off += (int64_t)pm.n * pm.m;
    groups[kc].push_back(p);
  }
  pk->mvar_total = var_off;
  FFP_CUDA(cudaMalloc(&pk->d_meta, sizeof(PulsarMeta) * P));
  FFP_CUDA(cudaMemcpy(pk->d_meta, pk->meta.data(), sizeof(PulsarMeta) * P, cudaMemcpyHostToDevice));
  FFP_CUDA(cudaMalloc(&pk->d_packets, (size_t)pk_off * 8));
  FFP_CUDA(cudaMalloc(&pk->d_L, (size_t)L_off * 8));
  FFP_CUDA(cudaMalloc(&pk->d_info, sizeof(int) * P));
  FFP_CUDA(cudaMemset(pk->d_info, 0, sizeof(int) * P));
  FFP_CUDA(cudaDeviceGetAttribute(&pk->num_sms, cudaDevAttrMultiProcessorCount, pk->device));
  const size_t slab_bytes = (size_t)CTAS_PER_SM * pk->num_sms * sweep_max_slab_doubles() * 8;
  FFP_CUDA(cudaMalloc(&pk->d_slab, slab_bytes));
  FFP_CUDA(cudaMalloc(&pk->d_counter, sizeof(unsigned int)));
  pk->bytes = pk_off * 8 + L_off * 8 + (int64_t)sizeof(PulsarMeta) * P + (int64_t)slab_bytes;
  for (auto& kv : groups) {
    Group g;
    g.cfg = kv.first;
    g.count = (int)kv.second.size();
    FFP_CUDA(cudaMalloc(&g.d_pidx, sizeof(int) * g.count));
    FFP_CUDA(cudaMemcpy(g.d_pidx, kv.second.data(), sizeof(int) * g.count, cudaMemcpyHostToDevice));
    pk->groups.push_back(g);
  }
  return 0;
}

static int upload_ragged(double** dst, const double* const* src, const fastfp_pack* pk, int which,
                         cudaStream_t st) {
  // which: 0 = length n_p, 1 = n_p*m_p (T), 2 = m_p*m_p
  int64_t total = 0;
  for (auto& pm : pk->meta)
    total += which == 0 ? pm.n : which == 1 ? (int64_t)pm.n * pm.m : (int64_t)pm.m * pm.m;
  if (*dst == nullptr) FFP_CUDA(cudaMalloc(dst, (size_t)total * 8));
  int64_t off = 0;
  for (int p = 0; p < pk->P; ++p) {
    const PulsarMeta& pm = pk->meta[p];
    const int64_t cnt = which == 0 ? pm.n : which == 1 ? (int64_t)pm.n * pm.m : (int64_t)pm.m * pm.m;
    if (!src[p]) { set_error("null per-pulsar array"); return FASTFP_ERR_INVALID; }
    FFP_CUDA(cudaMemcpyAsync(*dst + off, src[p], (size_t)cnt * 8, cudaMemcpyHostToDevice, st));
    off += cnt;
  }
  return 0;
}

// sum_i 1/N_i per pulsar in extended precision (the tensor sweep derives c N^-1 c from s N^-1 s with it)
static void set_ninv_sums(fastfp_pack* pk, const double* const* Nvecs) {
  for (int p = 0; p < pk->P; ++p) {
    long double acc = 0.0L;
    for (int i = 0; i < pk->meta[p].n; ++i) acc += 1.0L / (long double)Nvecs[p][i];
    pk->meta[p].ninv_sum = (double)acc;
  }
}

static void pack_free(fastfp_pack* pk) {
  if (!pk) return;
  DeviceGuard g(pk->device);
  for (auto& gr : pk->groups) { cudaFree(gr.d_pidx); cudaFree(gr.d_pidx_rest); }
  cudaFree(pk->d_meta); cudaFree(pk->d_packets); cudaFree(pk->d_L); cudaFree(pk->d_info);
  cudaFree(pk->d_S0); cudaFree(pk->d_zr); cudaFree(pk->d_slab); cudaFree(pk->d_counter); cudaFree(pk->d_done_mask);
  cudaFree(pk->d_terms); cudaFree(pk->d_freqs); cudaFree(pk->d_out); cudaFree(pk->d_scratch); cudaFree(pk->d_lf);
  cudaFree(pk->d_pl); cudaFreeHost(pk->h_pl);
  cudaFree(pk->d_i8); cudaFree(pk->d_i8_scale); cudaFree(pk->d_pidx_all); cudaFree(pk->d_inner);
  if (pk->pl_event) cudaEventDestroy(pk->pl_event);
  delete pk;
}

}  // namespace ffp

using namespace ffp;

extern "C" {

const char* fastfp_last_error(void) { return t_err.c_str(); }
int fastfp_version(void) { return 100; }
int fastfp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
int64_t fastfp_kernel_launches(void) { return g_launches.load(); }

int fastfp_pack_create(int device, int P, const int64_t* n, const int64_t* m,
                       const double* const* toas, const double* const* residuals,
                       const double* const* Nvecs, const double* const* Ts,
                       const double* const* sigmas, void* stream, fastfp_pack_t** out) {
  if (!out || P < 1 || !n || !m || !toas || !residuals || !Nvecs || !Ts || !sigmas) {
    set_error("fastfp_pack_create: null argument or P < 1");
    return FASTFP_ERR_INVALID;
  }
  *out = nullptr;
  DeviceGuard g(device);
  if (!g.ok) { set_error("cannot select CUDA device " + std::to_string(device)); return FASTFP_ERR_CUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  fastfp_pack* pk = new fastfp_pack();
  pk->device = device;
  int rc = pack_layout(pk, P, n, m, nullptr, toas);
  Staging sg;
  if (!rc) rc = upload_ragged(&sg.d_toas, toas, pk, 0, st);
  if (!rc) rc = upload_ragged(&sg.d_res, residuals, pk, 0, st);
  if (!rc) rc = upload_ragged(&sg.d_Nvec, Nvecs, pk, 0, st);
  if (!rc) rc = upload_ragged(&sg.d_T, Ts, pk, 1, st);
  if (!rc) rc = upload_ragged(&pk->d_L, sigmas, pk, 2, st);
  if (!rc) rc = launch_fp_precompute(pk, sg.d_toas, sg.d_res, sg.d_Nvec, sg.d_T, st);
  if (!rc) {
    set_ninv_sums(pk, Nvecs);
    rc = build_i8_planes(pk, st);  // digit planes for the tensor path when every pulsar fits its tile
  }
  if (rc) { pack_free(pk); return rc; }
  *out = pk;
  return FASTFP_OK;
}

int fastfp_pack_set_path(fastfp_pack_t* pk, int path) {
  if (!pk || path < FASTFP_PATH_AUTO || path > FASTFP_PATH_I8) {
    set_error("fastfp_pack_set_path: invalid argument");
    return FASTFP_ERR_INVALID;
  }
  if (path == FASTFP_PATH_I8 && !(pk->i8_ok && pk->i8_all())) {
    set_error("fastfp_pack_set_path: not every pulsar of this pack has INT8 digit planes (block-diagonal N, m > 639, "
              "n > 16384 or non-finite data); FASTFP_PATH_AUTO sweeps those on the fp64 kernel");
    return FASTFP_ERR_UNSUPPORTED;
  }
  pk->path = path;
  return FASTFP_OK;
}

int fastfp_pack_path(const fastfp_pack_t* pk) {
  if (!pk) return FASTFP_ERR_INVALID;
  return pk->use_i8() ? (pk->i8_all() ? FASTFP_PATH_I8 : FASTFP_PATH_MIXED) : FASTFP_PATH_FP64;
}

int fastfp_nmfp_pack_create(int device, int P, const int64_t* n, const int64_t* m,
                            const double* const* toas, const double* const* residuals,
                            const double* const* Nvecs, const double* const* Ts,
                            const double* const* TNTs, const int64_t* m_fix,
                            const double* const* phiinv_fix, void* stream, fastfp_pack_t** out) {
  if (!out || P < 1 || !n || !m || !toas || !residuals || !Nvecs || !Ts || !TNTs || !m_fix ||
      !phiinv_fix) {
    set_error("fastfp_nmfp_pack_create: null argument or P < 1");
    return FASTFP_ERR_INVALID;
  }
  *out = nullptr;
  DeviceGuard g(device);
  if (!g.ok) { set_error("cannot select CUDA device " + std::to_string(device)); return FASTFP_ERR_CUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  fastfp_pack* pk = new fastfp_pack();
  pk->device = device;
  pk->nmfp = true;
  int rc = pack_layout(pk, P, n, m, m_fix, toas);
  Staging sg;
  double *d_TNT = nullptr, *d_pf = nullptr;
  if (!rc) rc = upload_ragged(&sg.d_toas, toas, pk, 0, st);
  if (!rc) rc = upload_ragged(&sg.d_res, residuals, pk, 0, st);
  if (!rc) rc = upload_ragged(&sg.d_Nvec, Nvecs, pk, 0, st);
  if (!rc) rc = upload_ragged(&sg.d_T, Ts, pk, 1, st);
  if (!rc) rc = upload_ragged(&d_TNT, TNTs, pk, 2, st);
  if (!rc) {
    // fixed phiinv: (P, MAX_M) padded
    std::vector<double> pf((size_t)P * MAX_M, 0.0);
    for (int p = 0; p < P; ++p)
      for (int j = 0; j < pk->meta[p].mfix; ++j) pf[(size_t)p * MAX_M + j] = phiinv_fix[p][j];
    cudaError_t e = cudaMalloc(&d_pf, pf.size() * 8);
    if (e == cudaSuccess) e = cudaMemcpy(d_pf, pf.data(), pf.size() * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = cuda_fail(e, "upload phiinv_fix");
  }
  if (!rc) set_ninv_sums(pk, Nvecs);
  if (!rc) rc = nmfp_pack_finish(pk, sg.d_toas, sg.d_res, sg.d_Nvec, sg.d_T, d_TNT, d_pf, st);
  cudaFree(d_TNT);
  cudaFree(d_pf);
  if (rc) { pack_free(pk); return rc; }
  *out = pk;
  return FASTFP_OK;
}

int fastfp_sweep_chunk_toas(int64_t m, int blockn) {
  KernelCfg kc{};
  const int64_t m_rows = blockn ? (m + 7) / 8 * 8 + 8 : m;
  if (m < 1 || m > MAX_M || !sweep_config((int)m_rows, &kc)) return 0;
  return kc.ci;
}

int fastfp_pack_create_blockn(int device, int P, const int64_t* n, const int64_t* m,
                              const double* const* toas, const double* const* residuals,
                              const double* const* residuals_w, const double* const* Nvecs,
                              const double* const* Ts, const double* const* mats,
                              const int32_t* const* slot_idx, const double* const* slot_val,
                              const unsigned char* const* done_mask, const int64_t* m_fix,
                              const double* const* phiinv_fix, void* stream, fastfp_pack_t** out) {
  if (!out || P < 1 || !n || !m || !toas || !residuals || !residuals_w || !Nvecs || !Ts || !mats ||
      !slot_idx || !slot_val || !done_mask || (m_fix && !phiinv_fix)) {
    set_error("fastfp_pack_create_blockn: null argument or P < 1");
    return FASTFP_ERR_INVALID;
  }
  *out = nullptr;
  DeviceGuard g(device);
  if (!g.ok) { set_error("cannot select CUDA device " + std::to_string(device)); return FASTFP_ERR_CUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  fastfp_pack* pk = new fastfp_pack();
  pk->device = device;
  pk->nmfp = m_fix != nullptr;
  int rc = pack_layout(pk, P, n, m, m_fix, toas, true);
  Staging sg;
  double *d_resw = nullptr, *d_sval = nullptr, *d_mat = nullptr, *d_pf = nullptr;
  int* d_sidx = nullptr;
  if (!rc) rc = upload_ragged(&sg.d_toas, toas, pk, 0, st);
  if (!rc) rc = upload_ragged(&sg.d_res, residuals, pk, 0, st);
  if (!rc) rc = upload_ragged(&d_resw, residuals_w, pk, 0, st);
  if (!rc) rc = upload_ragged(&sg.d_Nvec, Nvecs, pk, 0, st);
  if (!rc) rc = upload_ragged(&sg.d_T, Ts, pk, 1, st);
  if (!rc) rc = upload_ragged(&d_sval, slot_val, pk, 0, st);
  if (!rc) {
    int64_t ntot = 0, nchtot = 0;
    for (auto& pm : pk->meta) { ntot += pm.n; nchtot += pm.nch; }
    cudaError_t e = cudaMalloc(&d_sidx, (size_t)ntot * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&pk->d_done_mask, (size_t)nchtot);
    for (int p = 0; p < P && e == cudaSuccess; ++p) {
      const PulsarMeta& pm = pk->meta[p];
      if (!slot_idx[p] || !done_mask[p]) { rc = FASTFP_ERR_INVALID; set_error("null slot array"); break; }
      e = cudaMemcpyAsync(d_sidx + pm.raw_off, slot_idx[p], (size_t)pm.n * sizeof(int), cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(pk->d_done_mask + pm.dm_off, done_mask[p], (size_t)pm.nch, cudaMemcpyHostToDevice, st);
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "block-N side arrays");
  }
  BlockNDev bn{d_resw, d_sidx, d_sval};
  if (!rc && !pk->nmfp) {
    rc = upload_ragged(&pk->d_L, mats, pk, 2, st);  // sigmas
    if (!rc) rc = launch_fp_precompute(pk, sg.d_toas, sg.d_res, sg.d_Nvec, sg.d_T, st, nullptr, &bn);
  } else if (!rc) {
    rc = upload_ragged(&d_mat, mats, pk, 2, st);  // TNTs
    if (!rc) {
      std::vector<double> pf((size_t)P * MAX_M, 0.0);
      for (int p = 0; p < P; ++p)
        for (int j = 0; j < pk->meta[p].mfix; ++j) pf[(size_t)p * MAX_M + j] = phiinv_fix[p][j];
      cudaError_t e = cudaMalloc(&d_pf, pf.size() * 8);
      if (e == cudaSuccess) e = cudaMemcpy(d_pf, pf.data(), pf.size() * 8, cudaMemcpyHostToDevice);
      if (e != cudaSuccess) rc = cuda_fail(e, "upload phiinv_fix");
    }
    if (!rc) rc = nmfp_pack_finish(pk, sg.d_toas, sg.d_res, sg.d_Nvec, sg.d_T, d_mat, d_pf, st, &bn);
  }
  cudaStreamSynchronize(st);
  cudaFree(d_resw); cudaFree(d_sval); cudaFree(d_sidx); cudaFree(d_mat); cudaFree(d_pf);
  if (rc) { pack_free(pk); return rc; }
  *out = pk;
  return FASTFP_OK;
}

void fastfp_pack_destroy(fastfp_pack_t* pack) { pack_free(pack); }
int64_t fastfp_pack_bytes(const fastfp_pack_t* pack) { return pack ? pack->bytes : 0; }
int fastfp_pack_num_pulsars(const fastfp_pack_t* pack) { return pack ? pack->P : 0; }
int64_t fastfp_pack_mvar_total(const fastfp_pack_t* pack) { return pack ? pack->mvar_total : 0; }
int fastfp_pack_factor_info(const fastfp_pack_t* pack, int32_t* info) {
  if (!pack) { set_error("fastfp_pack_factor_info: null pack"); return FASTFP_ERR_INVALID; }
  int bad = 0;
  for (int p = 0; p < pack->P; ++p) {
    const int v = p < (int)pack->info.size() ? pack->info[p] : 0;
    if (info) info[p] = v;
    bad += v != 0;
  }
  return bad;
}

// Frequencies are processed in batches so the (P, F_batch) term buffer stays bounded.
static const int64_t kTermBudgetDoubles = 1LL << 27;  // 1 GiB

// per-pulsar terms of one frequency batch on the path the pack is set to
static int sweep_terms(const fastfp_pack* pk, const double* d_freqs, int64_t F, double* d_terms, cudaStream_t st,
                       double* d_inner = nullptr) {
  return launch_sweep(pk, d_freqs, F, d_terms, st, nullptr, d_inner);
}

static int fp_run(const fastfp_pack* pk, const double* freqs, int64_t F, double* out, int flags,
                  void* stream, bool want_terms) {
  if (!pk || (F > 0 && (!freqs || !out)) || F < 0) {
    set_error("fastfp_fp_sweep: null argument or negative F");
    return FASTFP_ERR_INVALID;
  }
  if (pk->nmfp) { set_error("this pack was built for nmfp; use fastfp_nmfp_sweep"); return FASTFP_ERR_INVALID; }
  if (F == 0) return FASTFP_OK;
  DeviceGuard g(pk->device);
  cudaStream_t st = (cudaStream_t)stream;
  const bool fdev = flags & FASTFP_FREQS_ON_DEVICE, odev = flags & FASTFP_OUT_ON_DEVICE;
  const double* d_freqs = freqs;
  if (!fdev) {
    if (int rc = ensure(&pk->d_freqs, &pk->freqs_cap, F)) return rc;
    FFP_CUDA(cudaMemcpyAsync(pk->d_freqs, freqs, (size_t)F * 8, cudaMemcpyHostToDevice, st));
    d_freqs = pk->d_freqs;
  }
  const int P = pk->P;
  if (want_terms) {
    double* d_terms = out;
    if (!odev) {
      if (int rc = ensure(&pk->d_terms, &pk->terms_cap, (int64_t)P * F)) return rc;
      d_terms = pk->d_terms;
    }
    if (int rc = sweep_terms(pk, d_freqs, F, d_terms, st)) return rc;
    if (!odev) {
      FFP_CUDA(cudaMemcpyAsync(out, d_terms, (size_t)P * F * 8, cudaMemcpyDeviceToHost, st));
      FFP_CUDA(cudaStreamSynchronize(st));
    }
    return FASTFP_OK;
  }
  double* d_out = out;
  if (!odev) {
    if (int rc = ensure(&pk->d_out, &pk->out_cap, F)) return rc;
    d_out = pk->d_out;
  }
  const int64_t FB = std::max<int64_t>(1024, std::min<int64_t>(F, kTermBudgetDoubles / P));
  if (int rc = ensure(&pk->d_terms, &pk->terms_cap, (int64_t)P * std::min(FB, F))) return rc;
  for (int64_t lo = 0; lo < F; lo += FB) {
    const int64_t fb = std::min(FB, F - lo);
    if (int rc = sweep_terms(pk, d_freqs + lo, fb, pk->d_terms, st)) return rc;
    if (int rc = launch_reduce_terms(pk->d_terms, P, fb, d_out + lo, st)) return rc;
  }
  if (!odev) {
    FFP_CUDA(cudaMemcpyAsync(out, d_out, (size_t)F * 8, cudaMemcpyDeviceToHost, st));
    FFP_CUDA(cudaStreamSynchronize(st));
  }
  return FASTFP_OK;
}

int fastfp_fp_sweep(const fastfp_pack_t* pack, const double* freqs, int64_t F, double* out,
                    int flags, void* stream) {
  return fp_run(pack, freqs, F, out, flags, stream, false);
}
int fastfp_fp_terms(const fastfp_pack_t* pack, const double* freqs, int64_t F, double* terms,
                    int flags, void* stream) {
  return fp_run(pack, freqs, F, terms, flags, stream, true);
}

// Fe-statistic sky scan: one sweep for the inner products of every (pulsar, frequency), then the combine kernel
int fastfp_fe_sweep(const fastfp_pack_t* pk, const double* freqs, int64_t F, const double* fplus, const double* fcross,
                    int64_t S, double* out, int flags, void* stream) {
  if (!pk || F < 0 || S < 0 || ((F > 0 && S > 0) && (!freqs || !fplus || !fcross || !out))) {
    set_error("fastfp_fe_sweep: null argument or negative size");
    return FASTFP_ERR_INVALID;
  }
  if (pk->nmfp) { set_error("fastfp_fe_sweep needs a plain-Fp pack (fastfp_pack_create)"); return FASTFP_ERR_INVALID; }
  if (F == 0 || S == 0) return FASTFP_OK;
  DeviceGuard g(pk->device);
  cudaStream_t st = (cudaStream_t)stream;
  const bool fdev = flags & FASTFP_FREQS_ON_DEVICE, odev = flags & FASTFP_OUT_ON_DEVICE;
  const int P = pk->P;
  const double* d_freqs = freqs;
  if (!fdev) {
    if (int rc = ensure(&pk->d_freqs, &pk->freqs_cap, F)) return rc;
    FFP_CUDA(cudaMemcpyAsync(pk->d_freqs, freqs, (size_t)F * 8, cudaMemcpyHostToDevice, st));
    d_freqs = pk->d_freqs;
  }
  double* d_out = out;
  if (!odev) {
    if (int rc = ensure(&pk->d_out, &pk->out_cap, S * F)) return rc;
    d_out = pk->d_out;
  }
  const int64_t FB = std::max<int64_t>(1024, std::min<int64_t>(F, kTermBudgetDoubles / (5 * (int64_t)P)));
  // scratch: the inner products of one frequency batch, then the antenna patterns of the S sky positions
  if (int rc = ensure(&pk->d_inner, &pk->inner_cap, 5 * (int64_t)P * std::min(FB, F) + 2 * S * P)) return rc;
  double* d_fp = pk->d_inner + 5 * (int64_t)P * std::min(FB, F);
  double* d_fx = d_fp + S * P;
  FFP_CUDA(cudaMemcpyAsync(d_fp, fplus, (size_t)S * P * 8, cudaMemcpyHostToDevice, st));
  FFP_CUDA(cudaMemcpyAsync(d_fx, fcross, (size_t)S * P * 8, cudaMemcpyHostToDevice, st));
  for (int64_t lo = 0; lo < F; lo += FB) {
    const int64_t fb = std::min(FB, F - lo);
    if (int rc = sweep_terms(pk, d_freqs + lo, fb, nullptr, st, pk->d_inner)) return rc;
    if (int rc = launch_fe_combine(pk->d_inner, P, fb, d_fp, d_fx, S, d_out + lo, F, st)) return rc;
  }
  if (!odev) {
    FFP_CUDA(cudaMemcpyAsync(out, d_out, (size_t)S * F * 8, cudaMemcpyDeviceToHost, st));
    FFP_CUDA(cudaStreamSynchronize(st));
  } else {
    FFP_CUDA(cudaStreamSynchronize(st));  // fplus / fcross were read from caller-owned host memory
  }
  return FASTFP_OK;
}

int fastfp_nmfp_sweep(const fastfp_pack_t* pk, const double* freqs, int64_t F,
                      const double* phiinv_var, int64_t D, double* out, int flags, void* stream) {
  if (!pk || F < 0 || D < 0 || ((F > 0 && D > 0) && (!freqs || !phiinv_var || !out))) {
    set_error("fastfp_nmfp_sweep: null argument or negative size");
    return FASTFP_ERR_INVALID;
  }
  if (!pk->nmfp) { set_error("this pack was built for plain Fp; use fastfp_fp_sweep"); return FASTFP_ERR_INVALID; }
  if (F == 0 || D == 0) return FASTFP_OK;
  DeviceGuard g(pk->device);
  cudaStream_t st = (cudaStream_t)stream;
  const bool fdev = flags & FASTFP_FREQS_ON_DEVICE, odev = flags & FASTFP_OUT_ON_DEVICE,
             pdev = flags & FASTFP_PARAMS_ON_DEVICE;
  const double* d_freqs = freqs;
  if (!fdev) {
    if (int rc = ensure(&pk->d_freqs, &pk->freqs_cap, F)) return rc;
    FFP_CUDA(cudaMemcpyAsync(pk->d_freqs, freqs, (size_t)F * 8, cudaMemcpyHostToDevice, st));
    d_freqs = pk->d_freqs;
  }
  const double* d_phi = phiinv_var;
  double* d_phi_tmp = nullptr;
  if (!pdev) {
    FFP_CUDA(cudaMalloc(&d_phi_tmp, (size_t)D * pk->mvar_total * 8));
    FFP_CUDA(cudaMemcpyAsync(d_phi_tmp, phiinv_var, (size_t)D * pk->mvar_total * 8,
                             cudaMemcpyHostToDevice, st));
    d_phi = d_phi_tmp;
  }
  double* d_out = out;
  if (!odev) {
    if (int rc = ensure(&pk->d_out, &pk->out_cap, D * F)) { cudaFree(d_phi_tmp); return rc; }
    d_out = pk->d_out;
  }
  int rc = nmfp_sweep_impl(pk, d_freqs, F, d_phi, D, d_out, st);
  if (!rc && !odev) {
    cudaError_t e = cudaMemcpyAsync(out, d_out, (size_t)D * F * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "copy nmfp result to host");
  }
  if (d_phi_tmp) { cudaStreamSynchronize(st); cudaFree(d_phi_tmp); }
  return rc;
}

int fastfp_nmfp_tile_sizes(const fastfp_pack_t* pk, int64_t* z_per_tile, int64_t* a_per_tile) {
  if (!pk || !pk->nmfp || !z_per_tile || !a_per_tile) { set_error("fastfp_nmfp_tile_sizes: not an nmfp pack"); return FASTFP_ERR_INVALID; }
  *z_per_tile = (int64_t)pk->P * pk->mvpad * 64;
  *a_per_tile = (int64_t)pk->P * 160;
  return FASTFP_OK;
}

int fastfp_nmfp_stage_a(const fastfp_pack_t* pk, const double* freqs_dev, int64_t F, double* z_dev, double* a_dev,
                        void* stream) {
  if (!pk || !pk->nmfp || F <= 0 || !freqs_dev || !z_dev || !a_dev) {
    set_error("fastfp_nmfp_stage_a: not an nmfp pack, null argument or F <= 0");
    return FASTFP_ERR_INVALID;
  }
  DeviceGuard g(pk->device);
  return nmfp_stage_a_impl(pk, freqs_dev, F, z_dev, a_dev, (cudaStream_t)stream);
}

int fastfp_nmfp_stage_b(const fastfp_pack_t* pk, const double* freqs_dev, int64_t F, const double* z_dev,
                        const double* a_dev, int64_t tiles_per_block, const double* phiinv_var_dev, int64_t D,
                        double* out_dev, void* stream) {
  if (!pk || !pk->nmfp || F <= 0 || D < 0 || tiles_per_block <= 0 || !freqs_dev || !z_dev || !a_dev ||
      (D > 0 && (!phiinv_var_dev || !out_dev))) {
    set_error("fastfp_nmfp_stage_b: not an nmfp pack, null argument or bad size");
    return FASTFP_ERR_INVALID;
  }
  if (D == 0) return FASTFP_OK;
  DeviceGuard g(pk->device);
  return nmfp_stage_b_only(pk, freqs_dev, F, z_dev, a_dev, (int)tiles_per_block, phiinv_var_dev, D, out_dev,
                           (cudaStream_t)stream);
}

int fastfp_nmfp_stage_timing(fastfp_pack_t* pk, int enable) {
  if (!pk || !pk->nmfp) { set_error("fastfp_nmfp_stage_timing: not an nmfp pack"); return FASTFP_ERR_INVALID; }
  pk->time_stages = enable != 0;
  return FASTFP_OK;
}

int fastfp_nmfp_stage_ms(const fastfp_pack_t* pk, double* ms3) {
  if (!pk || !pk->nmfp || !ms3) { set_error("fastfp_nmfp_stage_ms: not an nmfp pack"); return FASTFP_ERR_INVALID; }
  for (int i = 0; i < 3; ++i) ms3[i] = pk->stage_ms[i];
  return FASTFP_OK;
}

int fastfp_powerlaw_phiinv(const fastfp_pack_t* pk, const double* const* Ffreqs,
                           const double* log10_A, const double* gamma, int64_t D,
                           const double* curn_Ffreqs, int64_t ncurn, const double* curn_log10_A,
                           const double* curn_gamma, double* phiinv_var_dev, void* stream) {
  if (!pk || !pk->nmfp || !Ffreqs || !log10_A || !gamma || D < 0 || !phiinv_var_dev ||
      (ncurn > 0 && (!curn_Ffreqs || !curn_log10_A || !curn_gamma))) {
    set_error("fastfp_powerlaw_phiinv: invalid argument");
    return FASTFP_ERR_INVALID;
  }
  if (D == 0) return FASTFP_OK;
  DeviceGuard g(pk->device);
  return powerlaw_phiinv_impl(pk, Ffreqs, log10_A, gamma, D, curn_Ffreqs, ncurn, curn_log10_A,
                              curn_gamma, phiinv_var_dev, (cudaStream_t)stream);
}

static int xcy_run(int device, int64_t n, int64_t m, const double* Nvec, const double* T, const double* sigma,
                   const double* x, const double* y, const double* x0, double* out, void* stream) {
  if (n < 1 || m < 1 || !Nvec || !T || !sigma || !x || !y || !out) {
    set_error("fastfp_xcy: null argument or non-positive size");
    return FASTFP_ERR_INVALID;
  }
  DeviceGuard g(device);
  if (!g.ok) { set_error("cannot select CUDA device " + std::to_string(device)); return FASTFP_ERR_CUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tot = (size_t)(4 * n + n * m + m * m) + (size_t)(m * m + 3 * m + 2);
  double* d = nullptr;
  FFP_CUDA(cudaMalloc(&d, tot * 8));
  double *dN = d, *dx = dN + n, *dy = dx + n, *dx0 = dy + n, *dT = dx0 + n, *dS = dT + n * m, *dW = dS + m * m;
  double* dO = dW + (m * m + 3 * m);
  cudaError_t e = cudaMemcpyAsync(dN, Nvec, n * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dx, x, n * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dy, y, n * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && x0) e = cudaMemcpyAsync(dx0, x0, n * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dT, T, n * m * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dS, sigma, m * m * 8, cudaMemcpyHostToDevice, st);
  int rc = 0;
  if (e != cudaSuccess) rc = cuda_fail(e, "fastfp_xcy upload");
  if (!rc) rc = launch_xcy(n, m, dN, dT, dS, dx, dy, x0 ? dx0 : nullptr, dW, dO, st);
  if (!rc) {
    e = cudaMemcpyAsync(out, dO, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "fastfp_xcy download");
  }
  cudaFree(d);
  return rc;
}

int fastfp_xcy(int device, int64_t n, int64_t m, const double* Nvec, const double* T,
               const double* sigma, const double* x, const double* y, double* out, void* stream) {
  return xcy_run(device, n, m, Nvec, T, sigma, x, y, nullptr, out, stream);
}

int fastfp_xcy_blockn(int device, int64_t n, int64_t m, const double* Nvec, const double* T,
                      const double* sigma, const double* x, const double* xw, const double* yw,
                      double* out, void* stream) {
  if (!xw || !yw) { set_error("fastfp_xcy_blockn: null argument"); return FASTFP_ERR_INVALID; }
  return xcy_run(device, n, m, Nvec, T, sigma, xw, yw, x, out, stream);
}

int fastfp_tnt(int device, int64_t n, int64_t m, const double* Nvec, const double* T, const double* phiinv,
               double* out, void* stream) {
  if (n < 1 || m < 1 || !Nvec || !T || !out) {
    set_error("fastfp_tnt: null argument or non-positive size");
    return FASTFP_ERR_INVALID;
  }
  DeviceGuard g(device);
  if (!g.ok) { set_error("cannot select CUDA device " + std::to_string(device)); return FASTFP_ERR_CUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  const int nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(64, n / 256));
  const size_t tot = (size_t)(n + n * m + m + m * m) + (size_t)nsplit * m * m;
  double* d = nullptr;
  FFP_CUDA(cudaMalloc(&d, tot * 8));
  double *dN = d, *dT = dN + n, *dP = dT + n * m, *dO = dP + m, *dW = dO + m * m;
  cudaError_t e = cudaMemcpyAsync(dN, Nvec, n * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dT, T, n * m * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && phiinv) e = cudaMemcpyAsync(dP, phiinv, m * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dW, 0, (size_t)nsplit * m * m * 8, st);
  int rc = 0;
  if (e != cudaSuccess) rc = cuda_fail(e, "fastfp_tnt upload");
  if (!rc) rc = launch_tnt(n, m, dN, dT, phiinv ? dP : nullptr, dW, nsplit, dO, st);
  if (!rc) {
    e = cudaMemcpyAsync(out, dO, m * m * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "fastfp_tnt download");
  }
  cudaFree(d);
  return rc;
}

int fastfp_fp64_peak(int device, int kind, int iters, double* tflops, double* ms) {
  if (!tflops || !ms || iters < 1 || kind < 0 || kind > 18) {
    set_error("fastfp_fp64_peak: invalid argument");
    return FASTFP_ERR_INVALID;
  }
  DeviceGuard g(device);
  if (!g.ok) { set_error("cannot select CUDA device"); return FASTFP_ERR_CUDA; }
  if (kind >= 17) return run_i8_peak(kind, iters, tflops, ms);
  return run_fp64_peak(kind, iters, tflops, ms);
}

}  // extern "C"
