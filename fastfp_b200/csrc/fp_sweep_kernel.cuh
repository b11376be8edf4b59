// The frequency-sweep kernel (fp64 DMMA formulation): a persistent, warp-specialised CTA -- SweepCfg::NWC
// consumer (MMA) warps + SweepCfg::NWP producer (sincos) warps, 8 + 16 = 768 threads by default, one CTA
// per SM -- takes (pulsar, frequency-tile) work items from an atomic counter.
//
// Replaces the body of FastFp.calculate_Fp under jax.vmap (reference fastfp/fastfp.py:69-92,
// examples/run_fp.py:63) -- and, with NMFP = true, the draw-independent part of
// NMFP.calculate_nmfp (fastfp/nmfp.py:96-119) -- for a whole frequency tile at once:
//
//   per chunk of CI TOAs
//     TMA        : two bulk copies per chunk -- the TOA vectors (t, 1/N, w) (to the producers) and
//                  the G tile in MMA-fragment order (to the consumers) -- through mbarrier rings
//     producers  : build sin/cos of ((2*pi)*f)*t (fastfp.py:78-79 phase order, one rounding per
//                  multiply) for their (frequency, TOA) pairs, store them into the shared S-tile
//                  ring in MMA-fragment order, and accumulate s N^-1 s, s N^-1 c, c N^-1 c, s.w, c.w
//     consumers  : Y[MP][2*KF] += G_chunk^T . S_chunk on the fp64 MMA path (mma.sync.m8n8k4.f64;
//                  measured to share the DFMA pipe and its peak, but with one 8-byte operand load
//                  per 128 FMAs instead of one per ~4)
//   epilogue     : b = Y_s.Y_s, Y_s.Y_c, Y_c.Y_c; M = [[sNs-b_ss, sNc-b_sc],[.., cNc-b_cc]],
//                  N = [s.w, c.w]; general 2x2 solve with partial pivoting (what jnp.linalg.solve
//                  does at fastfp.py:90); term = 0.5 * N . M^-1 N.
//
// Producers and consumers are decoupled by full/empty mbarriers, so the dependent sincos chains of
// the producers interleave with the consumers' MMAs on the shared fp64 pipe at run time. With the default
// split each SM sub-partition hosts two consumer warps and four producer warps (the producers' DFMAs queue
// behind the MMAs on the shared pipe, so it takes that many to keep the S ring full); the register file is
// split with setmaxnreg (SweepCfg::CREGS / PREGS = 120 / 56 registers per thread out of the 768 x 80 pool).
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
  double* inner;          // optional [P][F][5]: (s|s), (s|c), (c|c), (s|r), (c|r) (Fe-statistic); null = terms only
  double* slab;           // level-2 scratch, SLAB doubles per CTA
  unsigned int* counter;  // work counter (zeroed before the launch)
  double* Z;              // nmfp: [P][ceil(F/32)][mvmax/4][8][32] (B-fragment order, mvmax = padded)
  double* A;              // nmfp: [P][ceil(F/32)][5][32]
  int mvmax;
  const unsigned char* done_mask;  // block-N packs: per chunk, which of the 8 epoch slots end there
#ifdef FFP_DEBUG_SWITCHES
  int dbg;  // profiling builds only (tools/dbg_split.sh): bit 0 producers' math off, 1 MMAs off, 2 level-2 flush off
#endif
};
// The shipped library has no run-time switch that could skip work: the bits are a compile-time zero.
#ifdef FFP_DEBUG_SWITCHES
#define FFP_DBG(ar, bit) ((ar).dbg & (bit))
#else
#define FFP_DBG(ar, bit) 0
#endif

// D(8x8) += A(8x4) . B(4x8), fp64. Lane l holds A[l>>2][l&3], B[l&3][l>>2], D[l>>2][2*(l&3)+{0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}
template <int R>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R));
}
template <int R>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R));
}
// Shared-memory carve-up, identical for both roles.
template <class C>
struct SweepSmem {
  double* Sring;   // [C::SST][KB][NBT][32]   sin/cos tiles, B-fragment order
  double* Gring;   // [GST][KB][NMB][32]   G tiles, A-fragment order
  double* Vring;   // [VST][VEC]           t | 1/N | w
  double* fq;      // [KF]                 trial frequencies of the current tile
  double* red;     // [RED]                epilogue reduction scratch
  uint64_t *s_full, *s_empty, *g_full, *g_empty, *v_full, *v_empty;
  __device__ explicit SweepSmem(unsigned char* raw) {
    Sring = reinterpret_cast<double*>(raw);
    Gring = Sring + C::SST * C::ST;
    Vring = Gring + C::GST * C::GT;
    fq = Vring + VST * C::VEC;
    red = fq + C::KF;
    s_full = reinterpret_cast<uint64_t*>(red + C::RED);
    s_empty = s_full + C::SST;
    g_full = s_empty + C::SST;
    g_empty = g_full + C::GST;
    v_full = g_empty + C::GST;
    v_empty = v_full + VST;
  }
};

struct WorkItem {
  int p, ft, nch;
  const double* gpk;
};

// ---- producer role: sin/cos tiles + the five weighted sums -----------------------------------
template <class C, bool NMFP, bool ECORR>
__device__ __forceinline__ void producer_loop(const SweepArgs& ar, SweepSmem<C>& sm, volatile int* s_work,
                                              const int pw, const int lane) {
  constexpr int CI = C::CI, XW = C::XW;
  const int tidp = pw * 32 + lane;
  const int bk = lane & 3, bf8 = lane >> 2;
  const int bx0 = C::NX >= C::NWP ? pw * XW : pw % C::NX;         // first group of 8 frequencies
  const int bkb0 = C::NX >= C::NWP ? 0 : (pw / C::NX) * C::KBW;   // first k-block
  const int bsplit = C::NX >= C::NWP ? 0 : pw / C::NX;
  // element (frequency group x, k-block kb) -> S offset (kb*NBT + 2*x + (bf8>>2))*32 + 8*(bf8&3) + 2*bk:
  // the (sin, cos) pair of a (TOA, frequency) is adjacent, so it goes out as one 16-byte store
  const int sofs = (bf8 >> 2) * 32 + 8 * (bf8 & 3) + 2 * bk;
  double* const sl = ar.slab + (size_t)blockIdx.x * C::SLAB + (size_t)C::NACCX * C::NTC + tidp;
  uint32_t g = 0;
  for (;;) {
    __syncthreads();  // B1: work item published
    const int w = *s_work;
    if (w >= ar.nwork) break;
    const int gp = w / ar.ntile_f;
    const PulsarMeta pm = ar.meta[ar.pidx[gp]];
    const double* gpk = ar.packets + pm.pk_off;
    const int nch = pm.nch;
    auto issue_vec = [&](int c) {
      const uint32_t k = g + (uint32_t)c;
      uint64_t* b = &sm.v_full[k % VST];
      mbar_expect_tx(b, C::VEC * 8);
      tma_load_1d(sm.Vring + (k % VST) * C::VEC, gpk + (size_t)c * C::PK, C::VEC * 8, b);
    };
    if (pw == 0 && lane == 0)
      for (int c = 0; c < VST - 1 && c < nch; ++c) issue_vec(c);
    __syncthreads();  // B2: frequencies of the tile are in shared memory
    double omega[XW];
#pragma unroll
    for (int xx = 0; xx < XW; ++xx)
      omega[xx] = __dmul_rn(6.283185307179586, sm.fq[8 * (bx0 + xx) + bf8]);  // (2*pi)*f, rounded once
    bool fast = true;  // warp-uniform in practice; any lane out of range sends its warp down the cold path
#pragma unroll
    for (int xx = 0; xx < XW; ++xx) fast = fast && (fabs(omega[xx]) * pm.tabs_max <= 0.999 * FFP_SINCOS_MAX);
    fast = __all_sync(0xffffffffu, fast);
    double s2[XW][5];
#pragma unroll
    for (int xx = 0; xx < XW; ++xx)
#pragma unroll
      for (int q = 0; q < 5; ++q) s2[xx][q] = 0.0;
    bool flushed = false;

    for (int c = 0; c < nch; ++c) {
      const uint32_t k = g + (uint32_t)c;
      if (pw == 0 && lane == 0 && c + VST - 2 < nch && c >= 1) {
        if (k >= 2) mbar_wait(&sm.v_empty[(k - 2) % VST], ((k - 2) / VST) & 1u);
        issue_vec(c + VST - 2);
      }
      mbar_wait(&sm.v_full[k % VST], (k / VST) & 1u);
      if (k >= C::SST) mbar_wait(&sm.s_empty[k % C::SST], ((k / C::SST) - 1) & 1u);
      const double* pk = sm.Vring + (k % VST) * C::VEC;
      double* sb = sm.Sring + (k % C::SST) * C::ST + sofs;
      if (!FFP_DBG(ar, 1) && pw < C::NACTIVE) {
        if (fast) {
          // straight-line: every phase of this tile is inside the Cody-Waite range (checked once per
          // work item against the pulsar's largest |TOA|), so there is no per-element test
#pragma unroll
          for (int kk = 0; kk < C::KBW; ++kk) {
            const int kb = bkb0 + kk;
            const int i = 4 * kb + bk;
            const double2 tn = *reinterpret_cast<const double2*>(pk + 4 * i);  // (t, 1/N)
            const double wv = pk[4 * i + 2];
#pragma unroll
            for (int xx = 0; xx < XW; ++xx) {
              const double ph = __dmul_rn(omega[xx], tn.x);  // ((2*pi)*f)*t, rounded once more
              double s, cs;
              sincos_cw(ph, &s, &cs);
              *reinterpret_cast<double2*>(sb + (kb * C::NBT + 2 * (bx0 + xx)) * 32) = make_double2(s, cs);
              const double sn = s * tn.y, cn = cs * tn.y;
              s2[xx][0] = fma(sn, s, s2[xx][0]);
              s2[xx][1] = fma(sn, cs, s2[xx][1]);
              s2[xx][2] = fma(cn, cs, s2[xx][2]);
              s2[xx][3] = fma(s, wv, s2[xx][3]);
              s2[xx][4] = fma(cs, wv, s2[xx][4]);
            }
          }
        } else {
          // cold: some phase of this tile may exceed the Cody-Waite range (or is NaN/Inf): library sincos
#pragma unroll 1
          for (int e = 0; e < C::KBW * XW; ++e) {
            const int kb = bkb0 + e / XW, xx = e % XW;
            const int i = 4 * kb + bk;
            const double om = __dmul_rn(6.283185307179586, sm.fq[8 * (bx0 + xx) + bf8]);
            const double ph = __dmul_rn(om, pk[4 * i]);
            double s, cs;
            sincos(ph, &s, &cs);
            const double ni = pk[4 * i + 1], wv = pk[4 * i + 2];
            *reinterpret_cast<double2*>(sb + (kb * C::NBT + 2 * (bx0 + xx)) * 32) = make_double2(s, cs);
            const double sn = s * ni, cn = cs * ni;
#pragma unroll
            for (int x2 = 0; x2 < XW; ++x2)
              if (x2 == xx) {
                s2[x2][0] = fma(sn, s, s2[x2][0]);
                s2[x2][1] = fma(sn, cs, s2[x2][1]);
                s2[x2][2] = fma(cn, cs, s2[x2][2]);
                s2[x2][3] = fma(s, wv, s2[x2][3]);
                s2[x2][4] = fma(cs, wv, s2[x2][4]);
              }
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sm.s_full[k % C::SST]);
        mbar_arrive(&sm.v_empty[k % VST]);
      }
      if ((c + 1) % C::FLUSH == 0 && c + 1 < nch && !FFP_DBG(ar, 4)) {
#pragma unroll
        for (int xx = 0; xx < XW; ++xx)
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            double* a = sl + (size_t)(xx * 5 + q) * C::NTP;
            if (flushed) atomicAdd(a, s2[xx][q]);  // result unused -> RED.ADD.F64, no round trip
            else __stcg(a, s2[xx][q]);
            s2[xx][q] = 0.0;
          }
        flushed = true;
      }
    }
    g += (uint32_t)nch;
    // scalar sums: level-2 totals, then across the 4 lanes that share a frequency
    if (flushed) __threadfence();
#pragma unroll
    for (int xx = 0; xx < XW; ++xx)
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        double v = s2[xx][q];
        if (flushed) v += __ldcg(sl + (size_t)(xx * 5 + q) * C::NTP);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        s2[xx][q] = v;
      }
    double* redA = sm.red + C::WMW * C::KF * 3;  // [KSPLIT][KF][5]
    if (bk == 0 && pw < C::NACTIVE) {
#pragma unroll
      for (int xx = 0; xx < XW; ++xx)
#pragma unroll
        for (int q = 0; q < 5; ++q)
          redA[(bsplit * C::KF + 8 * (bx0 + xx) + bf8) * 5 + q] = s2[xx][q];
    }
    __syncthreads();  // B3: reductions published
    __syncthreads();  // B4: tile finished
  }
}

// ---- consumer role: the contraction and the epilogue -----------------------------------------
template <class C, bool NMFP, bool ECORR>
__device__ __forceinline__ void consumer_loop(const SweepArgs& ar, SweepSmem<C>& sm, volatile int* s_work,
                                              const int cw, const int lane) {
  constexpr int NMBW = C::NMBW, NNB = C::NNB;
  const int tid = cw * 32 + lane;
  const int wm = cw / C::WNW, wn = cw - wm * C::WNW;
  // B fragment: lane holds S[k = lane&3][n = lane>>2], n = 2*(freq%4) + {sin, cos}; stored at 8*(n>>1) + 2*k + (n&1)
  const int bperm = 8 * (lane >> 3) + 2 * (lane & 3) + ((lane >> 2) & 1);
  double* const sl = ar.slab + (size_t)blockIdx.x * C::SLAB + tid;
  uint32_t g = 0;
  for (;;) {
    if (tid == 0) *s_work = (int)atomicAdd(ar.counter, 1u);
    __syncthreads();  // B1
    const int w = *s_work;
    if (w >= ar.nwork) break;
    const int gp = w / ar.ntile_f, ft = w - gp * ar.ntile_f;
    const int p = ar.pidx[gp];
    const PulsarMeta pm = ar.meta[p];
    const double* gpk = ar.packets + pm.pk_off;
    const int nch = pm.nch;
    const int64_t f0 = (int64_t)ft * C::KF;
    auto issue_G = [&](int c) {
      const uint32_t k = g + (uint32_t)c;
      uint64_t* b = &sm.g_full[k % C::GST];
      mbar_expect_tx(b, C::GT * 8);
      tma_load_1d(sm.Gring + (k % C::GST) * C::GT, gpk + (size_t)c * C::PK + C::VEC, C::GT * 8, b);
    };
    if (tid == 0)
      for (int c = 0; c < C::GST - 1 && c < nch; ++c) issue_G(c);  // chunks 0 .. GST-2 in flight
    // a short last tile repeats its first frequency, so which sincos path a warp takes (and with it the last bits of
    // a bin) never depends on how many bins were passed along with it
    if (tid < C::KF) sm.fq[tid] = ar.freqs[f0 + tid < ar.F ? f0 + tid : f0];
    __syncthreads();  // B2

    double acc[NMBW][NNB][2];
#pragma unroll
    for (int r = 0; r < NMBW; ++r)
#pragma unroll
      for (int q = 0; q < NNB; ++q) acc[r][q][0] = acc[r][q][1] = 0.0;
    bool flushed = false;
    // block-diagonal N (kernel ECORR): the last row block of the last warp row holds 8 epoch slots;
    // Y there is sqrt(beta_e) * sum_{i in e} x_i / N_i, folded into these sums when the epoch ends
    double es[NNB][3];
#pragma unroll
    for (int q = 0; q < NNB; ++q) es[q][0] = es[q][1] = es[q][2] = 0.0;
    const bool slot_warp = ECORR && wm == C::WMW - 1;
    const unsigned char* dmask = ECORR ? ar.done_mask + pm.dm_off : nullptr;

    // fragments of the next k-block -- also across chunk boundaries -- are fetched while the MMAs of
    // the current one run
    double a0[NMBW], b0[NNB];
    {
      const uint32_t k = g;
      mbar_wait_spin(&sm.g_full[k % C::GST], (k / C::GST) & 1u);
      mbar_wait_spin(&sm.s_full[k % C::SST], (k / C::SST) & 1u);
      const double* gt = sm.Gring + (k % C::GST) * C::GT + (wm * NMBW) * 32 + lane;
      const double* sb = sm.Sring + (k % C::SST) * C::ST + (wn * NNB) * 32 + bperm;
#pragma unroll
      for (int r = 0; r < NMBW; ++r) a0[r] = gt[r * 32];
#pragma unroll
      for (int q = 0; q < NNB; ++q) b0[q] = sb[q * 32];
    }
    for (int c = 0; c < nch; ++c) {
      const uint32_t k = g + (uint32_t)c;
      // the issuer refills the stage freed two chunks ago, so its wait practically never blocks
      if (tid == 0 && c + C::GST - 2 < nch && c >= 1) {
        if (k >= 2) mbar_wait(&sm.g_empty[(k - 2) % C::GST], ((k - 2) / C::GST) & 1u);
        issue_G(c + C::GST - 2);
      }
      const double* gt = sm.Gring + (k % C::GST) * C::GT + (wm * NMBW) * 32 + lane;
      const double* sb = sm.Sring + (k % C::SST) * C::ST + (wn * NNB) * 32 + bperm;
#pragma unroll
      for (int kb = 0; kb < C::KB; ++kb) {
        double a1[NMBW], b1[NNB];
        if (kb + 1 < C::KB) {
#pragma unroll
          for (int r = 0; r < NMBW; ++r) a1[r] = gt[((kb + 1) * C::NMB + r) * 32];
#pragma unroll
          for (int q = 0; q < NNB; ++q) b1[q] = sb[((kb + 1) * C::NBT + q) * 32];
        } else if (c + 1 < nch) {
          const uint32_t k1 = k + 1;
          mbar_wait_spin(&sm.g_full[k1 % C::GST], (k1 / C::GST) & 1u);
          mbar_wait_spin(&sm.s_full[k1 % C::SST], (k1 / C::SST) & 1u);
          const double* gt1 = sm.Gring + (k1 % C::GST) * C::GT + (wm * NMBW) * 32 + lane;
          const double* sb1 = sm.Sring + (k1 % C::SST) * C::ST + (wn * NNB) * 32 + bperm;
#pragma unroll
          for (int r = 0; r < NMBW; ++r) a1[r] = gt1[r * 32];
#pragma unroll
          for (int q = 0; q < NNB; ++q) b1[q] = sb1[q * 32];
        } else {
#pragma unroll
          for (int r = 0; r < NMBW; ++r) a1[r] = 0.0;
#pragma unroll
          for (int q = 0; q < NNB; ++q) b1[q] = 0.0;
        }
        if (!FFP_DBG(ar, 2)) {
#pragma unroll
          for (int r = 0; r < NMBW; ++r)
#pragma unroll
            for (int q = 0; q < NNB; ++q) dmma_m8n8k4(acc[r][q][0], acc[r][q][1], a0[r], b0[q]);
        }
#pragma unroll
        for (int r = 0; r < NMBW; ++r) a0[r] = a1[r];
#pragma unroll
        for (int q = 0; q < NNB; ++q) b0[q] = b1[q];
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sm.g_empty[k % C::GST]);
        mbar_arrive(&sm.s_empty[k % C::SST]);
      }
      if (ECORR && slot_warp) {
        // epochs that end in this chunk: e_xy += (sqrt(beta) A_x)(sqrt(beta) A_y); the slot restarts
        if ((dmask[c] >> (lane >> 2)) & 1) {
#pragma unroll
          for (int q = 0; q < NNB; ++q) {
            const double ys = acc[NMBW - 1][q][0], yc = acc[NMBW - 1][q][1];
            es[q][0] = fma(ys, ys, es[q][0]);
            es[q][1] = fma(ys, yc, es[q][1]);
            es[q][2] = fma(yc, yc, es[q][2]);
            acc[NMBW - 1][q][0] = acc[NMBW - 1][q][1] = 0.0;
          }
        }
      }
      if ((c + 1) % C::FLUSH == 0 && c + 1 < nch && !FFP_DBG(ar, 4)) {
        // fold the level-1 sums into the level-2 totals of this CTA's scratch slab
#pragma unroll
        for (int r = 0; r < NMBW; ++r)
#pragma unroll
          for (int q = 0; q < NNB; ++q)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (ECORR && r == NMBW - 1 && slot_warp) continue;  // open epoch sums stay in registers
              // the slot is private to this thread: first block stores, later blocks add with a
              // fire-and-forget reduction (RED.ADD.F64) -- a load/add/store chain would expose one L2
              // round trip per accumulator
              double* a = sl + (size_t)((r * NNB + q) * 2 + e) * C::NTC;
              if (flushed) atomicAdd(a, acc[r][q][e]);
              else __stcg(a, acc[r][q][e]);
              acc[r][q][e] = 0.0;
            }
        if (ECORR && slot_warp) {
#pragma unroll
          for (int q = 0; q < NNB; ++q)
#pragma unroll
            for (int e = 0; e < 3; ++e) {
              double* a = sl + (size_t)(C::NACC + q * 3 + e) * C::NTC;
              if (flushed) atomicAdd(a, es[q][e]);
              else __stcg(a, es[q][e]);
              es[q][e] = 0.0;
            }
        }
        flushed = true;
      }
    }
    g += (uint32_t)nch;

    // ---- epilogue ------------------------------------------------------------------------
    if (flushed) {
      __threadfence();  // the reductions above are complete before the read-back
#pragma unroll
      for (int r = 0; r < NMBW; ++r)
#pragma unroll
        for (int q = 0; q < NNB; ++q)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (!(ECORR && r == NMBW - 1 && slot_warp)) acc[r][q][e] += __ldcg(sl + (size_t)((r * NNB + q) * 2 + e) * C::NTC);
      if (ECORR && slot_warp) {
#pragma unroll
        for (int q = 0; q < NNB; ++q)
#pragma unroll
          for (int e = 0; e < 3; ++e) es[q][e] += __ldcg(sl + (size_t)(C::NACC + q * 3 + e) * C::NTC);
      }
    }
    // this thread holds Y[row][freq] for rows 8*(wm*NMBW + r) + (lane>>2) and the tile frequencies
    // 4*(wn*NNB + q) + (lane&3): [..][0] is the sin column, [..][1] the cos column
    const int mfix = pm.mfix;  // rows below mfix enter the b-sums (plain Fp: mfix = m)
    double* redB = sm.red;  // [WMW][KF][3]
#pragma unroll
    for (int q = 0; q < NNB; ++q) {
      double pss = 0, psc = 0, pcc = 0;
#pragma unroll
      for (int r = 0; r < NMBW; ++r) {
        const int j = 8 * (wm * NMBW + r) + (lane >> 2);
        const double ys = acc[r][q][0], yc = acc[r][q][1];
        if (j < mfix) {
          pss = fma(ys, ys, pss);
          psc = fma(ys, yc, psc);
          pcc = fma(yc, yc, pcc);
        } else if (NMFP && j < pm.m) {
          // nmfp: rows of the per-draw block go out as z' (canonical 32-frequency tiles)
          const int64_t f = f0 + 4 * (wn * NNB + q) + (lane & 3);
          if (f < ar.F) {
            // 32-frequency tile, MMA B-fragment order: k-block (row/4), column block (4 freqs), then
            // position 16*sc + 4*(freq%4) + row%4 -- the layout stage B loads without conflicts
            const int64_t nt32 = (ar.F + 31) >> 5;
            const int jr = j - mfix + (ar.mvmax - pm.mvar), fi = (int)(f & 31);  // rows follow the top padding
            double* z = ar.Z + ((size_t)p * nt32 + (f >> 5)) * ((size_t)ar.mvmax * 64) +
                        (size_t)(((jr >> 2) * 8 + (fi >> 2)) * 32 + 4 * (fi & 3) + (jr & 3));
            z[0] = ys;
            z[16] = yc;
          }
        }
      }
      if (ECORR && slot_warp) {  // the block-N correction enters exactly like the Woodbury b-sums
        pss += es[q][0];
        psc += es[q][1];
        pcc += es[q][2];
      }
      double v3[3] = {pss, psc, pcc};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = v3[k];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if ((lane >> 2) == 0) redB[(wm * C::KF + 4 * (wn * NNB + q) + (lane & 3)) * 3 + k] = v;
      }
    }
    __syncthreads();  // B3: reductions (consumer b-sums, producer scalar sums) published
    const int64_t fidx = f0 + tid;
    if (tid < C::KF && fidx < ar.F) {
      const double* redA = sm.red + C::WMW * C::KF * 3;  // [KSPLIT][KF][5]
      double b[3] = {0, 0, 0}, a[5] = {0, 0, 0, 0, 0};
      for (int w2 = 0; w2 < C::WMW; ++w2)
#pragma unroll
        for (int k = 0; k < 3; ++k) b[k] += redB[(w2 * C::KF + tid) * 3 + k];
      for (int g2 = 0; g2 < C::KSPLIT; ++g2)
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
        const double lq = m10 / m00;
        const double u = m11 - lq * m01;
        const double x1 = (n1 - lq * n0) / u;
        const double x0 = (n0 - m01 * x1) / m00;
        double val = 0.5 * (a[3] * x0 + a[4] * x1);
        if (!(sm.fq[tid] > 0.0)) val = __longlong_as_double(0x7ff8000000000000LL);
        if (ar.terms) ar.terms[(size_t)p * ar.F + fidx] = val;
        if (ar.inner) {  // the inner products themselves (no f^(-1/3) prefactor: it cancels in every statistic)
          double* o = ar.inner + ((size_t)p * ar.F + fidx) * 5;
          o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; o[3] = a[3]; o[4] = a[4];
        }
      }
    }
    __syncthreads();  // B4: fq, red and s_work are reused by the next work item
  }
}

template <class C, bool NMFP, bool ECORR>
__global__ void __launch_bounds__(C::NTHREADS, CTAS_PER_SM) fp_sweep_kernel(const SweepArgs ar) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SweepSmem<C> sm(smem_raw);
  __shared__ int s_work;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < C::SST; ++s) { mbar_init(&sm.s_full[s], C::NWP); mbar_init(&sm.s_empty[s], C::NWC); }
    for (int s = 0; s < C::GST; ++s) { mbar_init(&sm.g_full[s], 1); mbar_init(&sm.g_empty[s], C::NWC); }
    for (int s = 0; s < VST; ++s) { mbar_init(&sm.v_full[s], 1); mbar_init(&sm.v_empty[s], C::NWP); }
    fence_barrier_init();
  }
  __syncthreads();
  if (wid < C::NWC) {
    reg_alloc<C::CREGS>();
    consumer_loop<C, NMFP, ECORR>(ar, sm, &s_work, wid, lane);
  } else {
    reg_dealloc<C::PREGS>();
    producer_loop<C, NMFP, ECORR>(ar, sm, &s_work, wid - C::NWC, lane);
  }
}

// ---- launch helpers -------------------------------------------------------------------------
template <class C, bool NMFP, bool ECORR>
int launch_sweep_cfg(const fastfp_pack* pk, const Group& g, const SweepArgs& base, cudaStream_t st) {
  static bool attr_done[64] = {};
  if (!attr_done[pk->device & 63]) {
    FFP_CUDA(cudaFuncSetAttribute(fp_sweep_kernel<C, NMFP, ECORR>,
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
  const int64_t resident = (int64_t)CTAS_PER_SM * pk->num_sms;
  const unsigned grid = (unsigned)(nwork < resident ? nwork : resident);
  FFP_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), st));
  fp_sweep_kernel<C, NMFP, ECORR><<<grid, C::NTHREADS, C::SMEM, st>>>(a);
  g_launches += 1;
  FFP_CUDA(cudaGetLastError());
  return 0;
}

// one translation unit per configuration family instantiates these (compile time)
int dispatch_sweep_w1(const fastfp_pack*, const Group&, const SweepArgs&, bool nmfp, cudaStream_t);
int dispatch_sweep_w2(const fastfp_pack*, const Group&, const SweepArgs&, bool nmfp, cudaStream_t);
int dispatch_sweep_w4(const fastfp_pack*, const Group&, const SweepArgs&, bool nmfp, cudaStream_t);
int dispatch_sweep_wide(const fastfp_pack*, const Group&, const SweepArgs&, bool nmfp, cudaStream_t);
int dispatch_sweep_xwide(const fastfp_pack*, const Group&, const SweepArgs&, bool nmfp, cudaStream_t);

#define FFP_SWEEP_CASE(NMBWv, NNBv, WMWv, CIv) FFP_SWEEP_CASE_W(NMBWv, NNBv, WMWv, CIv, 8, 16)
#define FFP_SWEEP_CASE_W(NMBWv, NNBv, WMWv, CIv, NWCv, NWPv)                                       \
  if (g.cfg.nmbw == NMBWv && g.cfg.nnb == NNBv && g.cfg.wmw == WMWv && g.cfg.ci == CIv &&          \
      g.cfg.nwc == NWCv) {                                                                         \
    using Cfg_ = SweepCfg<NMBWv, NNBv, WMWv, CIv, NWCv, NWPv>;                                      \
    if (pk->ecorr) return nmfp ? launch_sweep_cfg<Cfg_, true, true>(pk, g, a, st)                   \
                               : launch_sweep_cfg<Cfg_, false, true>(pk, g, a, st);                 \
    return nmfp ? launch_sweep_cfg<Cfg_, true, false>(pk, g, a, st)                                 \
                : launch_sweep_cfg<Cfg_, false, false>(pk, g, a, st);                               \
  }

}  // namespace ffp
