// The frequency-sweep kernel (v2): a persistent CTA of 4 warps takes (pulsar, frequency-tile)
// work items from an atomic counter; two CTAs are resident per SM.
//
// Replaces the body of FastFp.calculate_Fp under jax.vmap (reference fastfp/fastfp.py:69-92,
// examples/run_fp.py:63) -- and, with NMFP = true, the draw-independent part of
// NMFP.calculate_nmfp (fastfp/nmfp.py:96-119) -- for a whole frequency tile at once:
//
//   for each chunk of CI TOAs (two TMA bulk copies: the vectors t | 1/N | w, and the tile G[CI][MP]):
//     basis phase : each thread builds sin/cos of ((2*pi)*f)*t (fastfp.py:78-79 phase order, one
//                   rounding per multiply) for its (frequency, TOA) pairs, stores the pairs into
//                   the shared S tile and accumulates s N^-1 s, s N^-1 c, c N^-1 c, s.w, c.w
//     contraction : Y[MP][2*KF] += G_chunk^T . S_chunk as a register-tiled fp64 outer-product loop
//                   (TM x 2*TQ accumulators per thread)
//   epilogue      : b = Y_s.Y_s, Y_s.Y_c, Y_c.Y_c; M = [[sNs-b_ss, sNc-b_sc],[.., cNc-b_cc]],
//                   N = [s.w, c.w]; general 2x2 solve with partial pivoting (what
//                   jnp.linalg.solve does at fastfp.py:90); term = 0.5 * N . M^-1 N.
//
// The f^(-1/3) prefactor of fastfp.py:78-79 scales N by a and M by a^2 and cancels exactly in
// N^T M^-1 N; it is not applied (f <= 0 still yields NaN as in the reference).
//
// Summation is blocked: level-1 register sums over FLUSH_TOAS TOAs are folded into level-2
// totals that live in an L2-resident scratch slab (one per resident CTA), so the rounding error
// of the n-long sums stays at the level of a BLAS/XLA dot, and the register file holds only one
// set of accumulators.
#pragma once
#include "ffp_internal.cuh"
#include "ffp_sincos.cuh"

namespace ffp {

struct SweepArgs {
  const double* packets;
  const PulsarMeta* meta;
  const int* pidx;
  int ntile_f;
  int nwork;
  const double* freqs;
  int64_t F;
  double* terms;          // [P][F] (plain Fp)
  double* slab;           // level-2 scratch, SLAB doubles per CTA
  unsigned int* counter;  // work counter (zeroed before the launch)
  double* Z;              // nmfp: [P][ceil(F/32)][mvmax][64]
  double* A;              // nmfp: [P][ceil(F/32)][5][32]
  int mvmax;
  long long* trace;
  int dbg;
};

template <class C, bool NMFP>
__global__ void __launch_bounds__(NT, 2) fp_sweep_kernel(const SweepArgs ar) {
  constexpr int TM = C::TM, TQ = C::TQ, CI = C::CI;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Sbuf = reinterpret_cast<double*>(smem_raw);   // [2][CI][SROW]
  double* Gring = Sbuf + 2 * CI * C::SROW;              // [GST][GT]
  double* Vring = Gring + GST * C::GT;                  // [VST][VEC]
  uint64_t* gbar = reinterpret_cast<uint64_t*>(Vring + VST * C::VEC);  // [GST]
  uint64_t* vbar = gbar + GST;                                         // [VST]
  __shared__ int s_work;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // basis-phase mapping: thread <-> (frequency fl, TOA group ig)
  const int fl = tid % C::KF, ig = tid / C::KF;
  // contraction mapping: warp (wm, wn), lane (lm, ln)
  const int wm = wid / C::WNW, wn = wid - wm * C::WNW;
  const int lm = lane >> 3, ln = lane & 7;
  const int aoff = wm * 4 * TM + lm;            // rows aoff + 4*r
  const int boff = 2 * (wn * 8 * TQ + ln);      // (s,c) pair of frequency wn*8*TQ + ln + 8*q at +16*q
  double* const sl = ar.slab + (size_t)blockIdx.x * C::SLAB + tid;
  const bool basis_first = (wid & 1) == 0;
  const bool tr = ar.trace != nullptr && blockIdx.x == 0 && lane == 0;

  if (tid == 0) {
    for (int s = 0; s < GST; ++s) mbar_init(&gbar[s], 1);
    for (int s = 0; s < VST; ++s) mbar_init(&vbar[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  uint32_t g = 0;  // chunks consumed by this CTA so far: ring stage / mbarrier parity bookkeeping
  int nitem = 0;
  for (;;) {
    if (tid == 0) s_work = (int)atomicAdd(ar.counter, 1u);
    __syncthreads();
    const int w = s_work;
    if (w >= ar.nwork) break;
    const int gp = w / ar.ntile_f, ft = w - gp * ar.ntile_f;
    const int p = ar.pidx[gp];
    const PulsarMeta pm = ar.meta[p];
    const double* gpk = ar.packets + pm.pk_off;
    const int nch = pm.nch;
    const int64_t fidx = (int64_t)ft * C::KF + fl;
    const double fval = fidx < ar.F ? ar.freqs[fidx] : 1.0;
    const double omega = __dmul_rn(6.283185307179586, fval);  // (2*pi)*f, rounded once

    auto issue_vec = [&](int c) {
      const uint32_t k = g + (uint32_t)c;
      uint64_t* b = &vbar[k % VST];
      mbar_expect_tx(b, C::VEC * 8);
      tma_load_1d(Vring + (k % VST) * C::VEC, gpk + (size_t)c * C::PK, C::VEC * 8, b);
    };
    auto issue_G = [&](int c) {
      const uint32_t k = g + (uint32_t)c;
      uint64_t* b = &gbar[k % GST];
      mbar_expect_tx(b, C::GT * 8);
      tma_load_1d(Gring + (k % GST) * C::GT, gpk + (size_t)c * C::PK + C::VEC, C::GT * 8, b);
    };
    if (tid == 0) {
      issue_vec(0);
      if (nch > 1) issue_vec(1);
      issue_G(0);
    }

    double acc[TM][2 * TQ];
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int q = 0; q < 2 * TQ; ++q) acc[r][q] = 0.0;
    double s2[5] = {0, 0, 0, 0, 0};
    bool flushed = false;

    auto build_basis = [&](int c) {
      const uint32_t k = g + (uint32_t)c;
      mbar_wait(&vbar[k % VST], (k / VST) & 1u);
      const double* pk = Vring + (k % VST) * C::VEC;
      double* sb = Sbuf + (c & 1) * CI * C::SROW;
      double l0 = 0, l1 = 0, l2 = 0, l3 = 0, l4 = 0;
      bool big = false;
#pragma unroll 8
      for (int kk = 0; kk < C::IPT; ++kk) {  // straight-line: the evaluations interleave
        const int i = ig * C::IPT + kk;
        const double ph = __dmul_rn(omega, pk[i]);  // ((2*pi)*f)*t, rounded once more
        const bool ok = fabs(ph) <= FFP_SINCOS_MAX;
        big |= !ok;
        double s, cs;
        sincos_cw(ok ? ph : 0.0, &s, &cs);
        const double ni = ok ? pk[CI + i] : 0.0, wv = ok ? pk[2 * CI + i] : 0.0;
        *reinterpret_cast<double2*>(sb + i * C::SROW + 2 * fl) = make_double2(s, cs);
        const double sn = s * ni, cn = cs * ni;
        l0 = fma(sn, s, l0);
        l1 = fma(sn, cs, l1);
        l2 = fma(cn, cs, l2);
        l3 = fma(s, wv, l3);
        l4 = fma(cs, wv, l4);
      }
      if (big) {  // cold: phases beyond the Cody-Waite range (or NaN/Inf) take the library path
#pragma unroll 1
        for (int kk = 0; kk < C::IPT; ++kk) {
          const int i = ig * C::IPT + kk;
          const double ph = __dmul_rn(omega, pk[i]);
          if (fabs(ph) <= FFP_SINCOS_MAX) continue;
          double s, cs;
          sincos(ph, &s, &cs);
          const double ni = pk[CI + i], wv = pk[2 * CI + i];
          *reinterpret_cast<double2*>(sb + i * C::SROW + 2 * fl) = make_double2(s, cs);
          const double sn = s * ni, cn = cs * ni;
          l0 = fma(sn, s, l0);
          l1 = fma(sn, cs, l1);
          l2 = fma(cn, cs, l2);
          l3 = fma(s, wv, l3);
          l4 = fma(cs, wv, l4);
        }
      }
      s2[0] += l0; s2[1] += l1; s2[2] += l2; s2[3] += l3; s2[4] += l4;
    };

    auto contract = [&](int c) {
      const uint32_t k = g + (uint32_t)c;
      mbar_wait(&gbar[k % GST], (k / GST) & 1u);
      const double* gt = Gring + (k % GST) * C::GT + aoff;
      const double* sb = Sbuf + (c & 1) * CI * C::SROW + boff;
#pragma unroll 2
      for (int i = 0; i < CI; ++i) {
        double a[TM], b[2 * TQ];
#pragma unroll
        for (int r = 0; r < TM; ++r) a[r] = gt[i * C::MP + 4 * r];
#pragma unroll
        for (int q = 0; q < TQ; ++q) {  // 8-byte loads: 1 shared-memory wavefront each
          b[2 * q] = sb[i * C::SROW + 16 * q];
          b[2 * q + 1] = sb[i * C::SROW + 16 * q + 1];
        }
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
          for (int q = 0; q < 2 * TQ; ++q) acc[r][q] = fma(a[r], b[q], acc[r][q]);
      }
    };

    // fold the level-1 sums into the level-2 totals of this CTA's scratch slab
    auto flush = [&]() {
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int q = 0; q < 2 * TQ; ++q) {
          double* a = sl + (size_t)(r * 2 * TQ + q) * NT;
          double v = acc[r][q];
          if (flushed) v += __ldcg(a);
          __stcg(a, v);
          acc[r][q] = 0.0;
        }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        double* a = sl + (size_t)(C::NACC + k) * NT;
        double v = s2[k];
        if (flushed) v += __ldcg(a);
        __stcg(a, v);
        s2[k] = 0.0;
      }
      flushed = true;
    };

    build_basis(0);
    __syncthreads();
    for (int c = 0; c < nch; ++c) {
      if (tr && nitem == 0 && c < 64) ar.trace[(c * 8 + wid) * 4 + 0] = clock64();
      if (tid == 0) {
        if (c + 1 < nch) issue_G(c + 1);
        if (c + 2 < nch) issue_vec(c + 2);
      }
      if (basis_first) {
        if (c + 1 < nch && !(ar.dbg & 1)) build_basis(c + 1);
        if (tr && nitem == 0 && c < 64) ar.trace[(c * 8 + wid) * 4 + 1] = clock64();
        if (!(ar.dbg & 2)) contract(c);
      } else {
        if (!(ar.dbg & 2)) contract(c);
        if (tr && nitem == 0 && c < 64) ar.trace[(c * 8 + wid) * 4 + 1] = clock64();
        if (c + 1 < nch && !(ar.dbg & 1)) build_basis(c + 1);
      }
      if ((c + 1) % C::FLUSH == 0 && c + 1 < nch && !(ar.dbg & 4)) flush();
      if (tr && nitem == 0 && c < 64) ar.trace[(c * 8 + wid) * 4 + 2] = clock64();
      __syncthreads();
      if (tr && nitem == 0 && c < 64) ar.trace[(c * 8 + wid) * 4 + 3] = clock64();
    }
    g += (uint32_t)nch;
    ++nitem;

    // ---- epilogue ------------------------------------------------------------------------
    if (flushed) {
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int q = 0; q < 2 * TQ; ++q) acc[r][q] += __ldcg(sl + (size_t)(r * 2 * TQ + q) * NT);
#pragma unroll
      for (int k = 0; k < 5; ++k) s2[k] += __ldcg(sl + (size_t)(C::NACC + k) * NT);
    }
    const int mfix = NMFP ? pm.mfix : pm.mpad;  // rows below mfix enter the b-sums
    double bs[TQ][3];
#pragma unroll
    for (int q = 0; q < TQ; ++q) {
      double pss = 0, psc = 0, pcc = 0;
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        const int j = aoff + 4 * r;
        const double ys = acc[r][2 * q], yc = acc[r][2 * q + 1];
        if (!NMFP || j < mfix) {
          pss = fma(ys, ys, pss);
          psc = fma(ys, yc, psc);
          pcc = fma(yc, yc, pcc);
        } else if (j < pm.m) {
          // nmfp: rows of the per-draw block go out as z' (canonical 32-frequency tiles)
          const int64_t f = (int64_t)ft * C::KF + wn * 8 * TQ + ln + 8 * q;
          if (f < ar.F) {
            const int64_t nt32 = (ar.F + 31) >> 5;
            double* z = ar.Z + (((size_t)p * nt32 + (f >> 5)) * ar.mvmax + (j - mfix)) * 64 + 2 * (f & 31);
            *reinterpret_cast<double2*>(z) = make_double2(ys, yc);
          }
        }
      }
      bs[q][0] = pss; bs[q][1] = psc; bs[q][2] = pcc;
    }
#pragma unroll
    for (int q = 0; q < TQ; ++q)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = bs[q][k];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        bs[q][k] = v;
      }
    double* redB = Sbuf;                            // [WMW][KF][3]
    double* redA = Sbuf + C::WMW * C::KF * 3;       // [IG][KF][5]
    if (lm == 0) {
#pragma unroll
      for (int q = 0; q < TQ; ++q) {
        const int f2 = wn * 8 * TQ + ln + 8 * q;
#pragma unroll
        for (int k = 0; k < 3; ++k) redB[(wm * C::KF + f2) * 3 + k] = bs[q][k];
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) redA[(ig * C::KF + fl) * 5 + k] = s2[k];
    __syncthreads();
    if (tid < C::KF && fidx < ar.F) {
      double b[3] = {0, 0, 0}, a[5] = {0, 0, 0, 0, 0};
      for (int w2 = 0; w2 < C::WMW; ++w2)
#pragma unroll
        for (int k = 0; k < 3; ++k) b[k] += redB[(w2 * C::KF + tid) * 3 + k];
      for (int g2 = 0; g2 < C::IG; ++g2)
#pragma unroll
        for (int k = 0; k < 5; ++k) a[k] += redA[(g2 * C::KF + tid) * 5 + k];
      if (NMFP) {
        // draw-independent pieces: a_ss, a_sc, a_cc (fixed block removed), a_sr, a_cr
        const int64_t nt32 = (ar.F + 31) >> 5;
        double* o = ar.A + ((size_t)p * nt32 + (fidx >> 5)) * 160 + (fidx & 31);
        o[0] = a[0] - b[0];
        o[32] = a[1] - b[1];
        o[64] = a[2] - b[2];
        o[96] = a[3];
        o[128] = a[4];
      } else {
        // M = [[ss, sc],[sc, cc]], N = [n0, n1]; LU with partial pivoting
        double m00 = a[0] - b[0], m01 = a[1] - b[1], m10 = m01, m11 = a[2] - b[2];
        double n0 = a[3], n1 = a[4];
        if (fabs(m10) > fabs(m00)) {  // row swap; the unknowns keep their order
          double t0 = m00; m00 = m10; m10 = t0;
          t0 = m01; m01 = m11; m11 = t0;
          t0 = n0; n0 = n1; n1 = t0;
        }
        const double l = m10 / m00;
        const double u = m11 - l * m01;
        const double x1 = (n1 - l * n0) / u;
        const double x0 = (n0 - m01 * x1) / m00;
        double val = 0.5 * (a[3] * x0 + a[4] * x1);
        if (!(fval > 0.0)) val = __longlong_as_double(0x7ff8000000000000LL);
        ar.terms[(size_t)p * ar.F + fidx] = val;
      }
    }
    __syncthreads();  // reduction scratch (Sbuf) and s_work are reused by the next work item
  }
}

// ---- launch helpers -------------------------------------------------------------------------
template <class C, bool NMFP>
int launch_sweep_cfg(const fastfp_pack* pk, const Group& g, const SweepArgs& base, cudaStream_t st) {
  static bool attr_done[64] = {};
  if (!attr_done[pk->device & 63]) {
    FFP_CUDA(cudaFuncSetAttribute(fp_sweep_kernel<C, NMFP>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr_done[pk->device & 63] = true;
  }
  SweepArgs a = base;
  a.pidx = g.d_pidx;
  const int64_t ntile = (a.F + C::KF - 1) / C::KF;
  const int64_t nwork = ntile * g.count;
  if (nwork > 0x7fffffffLL) { set_error("frequency batch too large for one launch"); return -1; }
  a.ntile_f = (int)ntile;
  a.nwork = (int)nwork;
  const int64_t resident = 2LL * pk->num_sms;
  const unsigned grid = (unsigned)(nwork < resident ? nwork : resident);
  FFP_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), st));
  fp_sweep_kernel<C, NMFP><<<grid, NT, C::SMEM, st>>>(a);
  g_launches += 1;
  FFP_CUDA(cudaGetLastError());
  return 0;
}

// one translation unit per WMW family instantiates these (compile time)
int dispatch_sweep_w1(const fastfp_pack*, const Group&, const SweepArgs&, bool nmfp, cudaStream_t);
int dispatch_sweep_w2(const fastfp_pack*, const Group&, const SweepArgs&, bool nmfp, cudaStream_t);
int dispatch_sweep_w4(const fastfp_pack*, const Group&, const SweepArgs&, bool nmfp, cudaStream_t);
int dispatch_sweep_wide(const fastfp_pack*, const Group&, const SweepArgs&, bool nmfp, cudaStream_t);

#define FFP_SWEEP_CASE(TMv, TQv, WMWv, CIv)                                                        \
  if (g.cfg.tm == TMv && g.cfg.tq == TQv && g.cfg.wmw == WMWv && g.cfg.ci == CIv)                   \
    return nmfp ? launch_sweep_cfg<SweepCfg<TMv, TQv, WMWv, CIv>, true>(pk, g, a, st)              \
                : launch_sweep_cfg<SweepCfg<TMv, TQv, WMWv, CIv>, false>(pk, g, a, st);

}  // namespace ffp
