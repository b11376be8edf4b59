// Host side of the sweep: configuration choice, launches, and the ordered pulsar sum.
#include <cstdlib>

#include "fp_sweep_kernel.cuh"

namespace ffp {

// Kernel configuration for a basis of width m. A consumer warp owns NMBW
// blocks of 8 basis rows x NNB blocks of 4 frequencies; wider bases put more consumer warps along
// the row direction and fewer frequencies in a tile.
bool sweep_config(int m, KernelCfg* c) {
  if (m < 1 || m > MAX_M) return false;
  if (m <= 40) { *c = {(m + 7) / 8, 4, 1, 16}; return true; }    // 128 frequencies per CTA
  // (a 12-consumer/8-producer split with 3 x 4 block warp tiles -- SweepCfg<3, 4, 3, 32, 12, 8> -- was
  // measured for m = 72: the MMA warps alone get 5% faster, the 8 producer warps fall behind, net -6%)
  if (m <= 80) { *c = {(m + 7) / 8, 2, 1, 32}; return true; }    // 64 frequencies per CTA
  if (m <= 160) { *c = {(m + 15) / 16, 2, 2, 16}; return true; }  // 32 frequencies per CTA
  if (m <= 320) { *c = {(m + 31) / 32, 2, 4, 16}; return true; }  // 16 frequencies per CTA
  // all eight consumer warps along the rows, chunks of 8 TOAs (the G tile of a chunk is 8 x 640 doubles = 40 KB): 8
  // frequencies per CTA. Real pulsars with many DMX columns land here; the tile is reused across 8 frequencies only, so
  // this family runs at a lower fraction of the pipe than the narrow ones.
  *c = {(m + 63) / 64, 2, 8, 8};                                  // m <= 640
  return true;
}

int sweep_max_slab_doubles() { return (40 + 12) * 12 * 32 + 10 * NTP; }  // NACC <= 80, XW <= 2, at most 12 consumer warps

// out[f] = sum over pulsars in pulsar order, starting from 0 (fastfp.py:71,90).
__global__ void reduce_terms_kernel(const double* __restrict__ terms, int P, int64_t F,
                                    double* __restrict__ out) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  double acc = 0.0;
  for (int p = 0; p < P; ++p) acc += terms[(size_t)p * F + f];
  out[f] = acc;
}

int launch_fp_sweep(const fastfp_pack* pk, const double* d_freqs, int64_t F, double* d_terms,
                    cudaStream_t st, const NmfpOut* nm, double* d_inner, bool rest_only) {
#ifdef FFP_DEBUG_SWITCHES  // profiling builds only; the shipped library is compiled without it
  static const int dbg = getenv("FASTFP_DBG") ? atoi(getenv("FASTFP_DBG")) : 0;
#endif
  SweepArgs a{};
  a.packets = pk->d_packets;
  a.meta = pk->d_meta;
  a.freqs = d_freqs;
  a.F = F;
  a.terms = d_terms;
  a.inner = d_inner;
  a.slab = pk->d_slab;
  a.counter = pk->d_counter;
  a.Z = nm ? nm->Z : nullptr;
  a.A = nm ? nm->A : nullptr;
  a.mvmax = nm ? nm->mvmax : 0;
  a.done_mask = pk->d_done_mask;
#ifdef FFP_DEBUG_SWITCHES
  a.dbg = dbg;
#endif
  for (const Group& g0 : pk->groups) {
    Group g = g0;
    if (rest_only) {  // only the pulsars the tensor sweep left out
      if (g0.count_rest == 0) continue;
      g.count = g0.count_rest;
      g.d_pidx = g0.d_pidx_rest;
    }
    int rc;
    if (g.cfg.wmw == 8) rc = dispatch_sweep_xwide(pk, g, a, nm != nullptr, st);
    else if (g.cfg.wmw == 4) rc = dispatch_sweep_wide(pk, g, a, nm != nullptr, st);
    else if (g.cfg.wmw == 1 && g.cfg.nnb == 4) rc = dispatch_sweep_w1(pk, g, a, nm != nullptr, st);
    else if (g.cfg.wmw == 1) rc = dispatch_sweep_w2(pk, g, a, nm != nullptr, st);
    else rc = dispatch_sweep_w4(pk, g, a, nm != nullptr, st);
    if (rc) return rc;
  }
  return 0;
}

int launch_sweep(const fastfp_pack* pk, const double* d_freqs, int64_t F, double* d_terms, cudaStream_t st,
                 const NmfpOut* nm, double* d_inner) {
  if (!pk->use_i8()) return launch_fp_sweep(pk, d_freqs, F, d_terms, st, nm, d_inner);
  if (int rc = launch_fp_sweep_i8(pk, d_freqs, F, d_terms, st, d_inner, nm)) return rc;
  return pk->i8_all() ? 0 : launch_fp_sweep(pk, d_freqs, F, d_terms, st, nm, d_inner, true);
}

int launch_reduce_terms(const double* d_terms, int P, int64_t F, double* d_out, cudaStream_t st) {
  reduce_terms_kernel<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(d_terms, P, F, d_out);
  g_launches += 1;
  FFP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ffp
