// Host side of the sweep: configuration choice, launches, and the ordered pulsar sum.
#include <cstdlib>

#include "fp_sweep_kernel.cuh"

namespace ffp {

// Kernel configuration for a basis of width m. Wider bases put more warps
// along the row direction and fewer frequencies in a tile; the shared-memory budget keeps two
// CTAs resident per SM in every configuration.
bool sweep_config(int m, KernelCfg* c) {
  if (m < 1 || m > MAX_M) return false;
  if (m <= 40) { *c = {(m + 3) / 4, 4, 1, 16}; return true; }
  if (m <= 80) { *c = {(m + 7) / 8, 4, 2, 32}; return true; }
  if (m <= 160) {
    const int tm = (m + 15) / 16;
    *c = {tm, 4, 4, tm <= 8 ? 32 : 16};
    return true;
  }
  *c = {(m + 15) / 16, 2, 4, 16};
  return true;
}

int sweep_max_slab_doubles() { return (10 * 8 + 5) * NT; }  // NACC <= 80 in every configuration

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
                    cudaStream_t st, const NmfpOut* nm, long long* trace) {
  static const int dbg = getenv("FASTFP_DBG") ? atoi(getenv("FASTFP_DBG")) : 0;  // profiling only
  SweepArgs a{};
  a.packets = pk->d_packets;
  a.meta = pk->d_meta;
  a.freqs = d_freqs;
  a.F = F;
  a.terms = d_terms;
  a.slab = pk->d_slab;
  a.counter = pk->d_counter;
  a.Z = nm ? nm->Z : nullptr;
  a.A = nm ? nm->A : nullptr;
  a.mvmax = nm ? nm->mvmax : 0;
  a.trace = trace;
  a.dbg = dbg;
  for (const Group& g : pk->groups) {
    int rc;
    if (g.cfg.tq == 2) rc = dispatch_sweep_wide(pk, g, a, nm != nullptr, st);
    else if (g.cfg.wmw == 1) rc = dispatch_sweep_w1(pk, g, a, nm != nullptr, st);
    else if (g.cfg.wmw == 2) rc = dispatch_sweep_w2(pk, g, a, nm != nullptr, st);
    else rc = dispatch_sweep_w4(pk, g, a, nm != nullptr, st);
    if (rc) return rc;
  }
  return 0;
}

int launch_reduce_terms(const double* d_terms, int P, int64_t F, double* d_out, cudaStream_t st) {
  reduce_terms_kernel<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(d_terms, P, F, d_out);
  g_launches += 1;
  FFP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ffp
