// Branch-free double-precision sincos for the Earth-term basis.
//
// The sweep kernel needs sin and cos of phi = ((2*pi)*f)*t (reference fastfp/fastfp.py:78-79)
// for ~1e11 (frequency, TOA) pairs per sweep; |phi| is a few 1e4 rad at most (t ~ 4.6e9 s,
// f <= 1e-6 Hz). CUDA's sincos() keeps a Payne-Hanek slow path behind a call, which turns every
// evaluation into its own control-flow region and stops the compiler from interleaving the
// independent evaluations a thread owns. This version is straight-line code:
//   k = rint(phi * 2/pi); r = phi - k*pi/2 by a three-constant Cody-Waite reduction with FMAs
//   (exact products, so the only error is the final rounding of r: <= 2^-53*|r| + 2e-33*|k|);
//   sin/cos kernels on |r| <= pi/4 with the fdlibm minimax coefficients; quadrant fix-up by
//   selects. Valid for |phi| <= FFP_SINCOS_MAX (k fits comfortably in 2^17); callers route larger
//   arguments to the library function. Measured max error vs long double: < 1 ulp-ish
//   (tests/test_sincos_host.py pins it).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define FFP_HD __host__ __device__ __forceinline__
#else
#define FFP_HD inline
#endif

#define FFP_SINCOS_MAX 1.0e5

namespace ffp {

FFP_HD double ffp_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}

FFP_HD void sincos_cw(double x, double* sp, double* cp) {
  // k = nearest integer to x*(2/pi), via the 1.5*2^52 magic constant (round-to-nearest-even)
  const double magic = 6755399441055744.0;
  const double kd = ffp_fma(x, 0.6366197723675814, magic);
#if defined(__CUDA_ARCH__)
  const int q = __double2loint(kd);
#else
  int64_t bits;
  std::memcpy(&bits, &kd, 8);
  const int q = (int)(uint32_t)bits;
#endif
  const double k = (double)q;  // == kd - magic exactly (|k| < 2^31); a conversion instead of an fp64-pipe add
  double r = ffp_fma(-k, 1.5707963267948966, x);
  r = ffp_fma(-k, 6.123233995736766e-17, r);
  r = ffp_fma(-k, -1.4973849048591698e-33, r);
  const double z = r * r;
  // sin(r) = r + r*z*(S1 + z*(S2 + ... z*S6))
  double ps = ffp_fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = ffp_fma(z, ps, 2.75573137070700676789e-06);
  ps = ffp_fma(z, ps, -1.98412698298579493134e-04);
  ps = ffp_fma(z, ps, 8.33333333332248946124e-03);
  ps = ffp_fma(z, ps, -1.66666666666666324348e-01);
  const double s = ffp_fma(r * z, ps, r);
  // cos(r) = 1 - z/2 + z*z*(C1 + z*(C2 + ... z*C6))
  double pc = ffp_fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = ffp_fma(z, pc, -2.75573143513906633035e-07);
  pc = ffp_fma(z, pc, 2.48015872894767294178e-05);
  pc = ffp_fma(z, pc, -1.38888888888741095749e-03);
  pc = ffp_fma(z, pc, 4.16666666666666019037e-02);
  const double c = ffp_fma(z, ffp_fma(z, pc, -0.5), 1.0);
  // quadrant: k mod 4 = 0:(s,c) 1:(c,-s) 2:(-s,-c) 3:(-c,s); the sign flips are integer XORs on
  // the high words (an fp64 negate + select would cost four more instructions per value)
  const bool swap = (q & 1) != 0;
  const double so = swap ? c : s;
  const double co = swap ? s : c;
  const uint32_t fs = ((uint32_t)q & 2u) << 30, fc = ((uint32_t)(q + 1) & 2u) << 30;
#if defined(__CUDA_ARCH__)
  *sp = __hiloint2double(__double2hiint(so) ^ (int)fs, __double2loint(so));
  *cp = __hiloint2double(__double2hiint(co) ^ (int)fc, __double2loint(co));
#else
  uint64_t bs, bc;
  std::memcpy(&bs, &so, 8);
  std::memcpy(&bc, &co, 8);
  bs ^= (uint64_t)fs << 32;
  bc ^= (uint64_t)fc << 32;
  std::memcpy(sp, &bs, 8);
  std::memcpy(cp, &bc, 8);
#endif
}

// NV evaluations of sincos_cw in lockstep: the same operations per element in the same order (bit-identical results),
// written step by step over the NV arguments so that the independent chains are adjacent in the instruction stream
// (the polynomial constants are materialised once per step, and the scheduler has NV-way parallelism to hide the
// fp64 latency) -- the producers of the tensor-core sweep evaluate four (TOA, frequency) pairs per thread and stage.
template <int NV>
FFP_HD void sincos_cw_n(const double (&x)[NV], double (&sp)[NV], double (&cp)[NV]) {
  const double magic = 6755399441055744.0;
  double kd[NV], k[NV], r[NV], z[NV], ps[NV], pc[NV];
  int q[NV];
#pragma unroll
  for (int e = 0; e < NV; ++e) kd[e] = ffp_fma(x[e], 0.6366197723675814, magic);
#pragma unroll
  for (int e = 0; e < NV; ++e) {
#if defined(__CUDA_ARCH__)
    q[e] = __double2loint(kd[e]);
#else
    int64_t bits;
    std::memcpy(&bits, &kd[e], 8);
    q[e] = (int)(uint32_t)bits;
#endif
    k[e] = (double)q[e];  // == kd[e] - magic exactly; the conversion pipe is not shared with the tensor core
  }
#pragma unroll
  for (int e = 0; e < NV; ++e) r[e] = ffp_fma(-k[e], 1.5707963267948966, x[e]);
#pragma unroll
  for (int e = 0; e < NV; ++e) r[e] = ffp_fma(-k[e], 6.123233995736766e-17, r[e]);
#pragma unroll
  for (int e = 0; e < NV; ++e) r[e] = ffp_fma(-k[e], -1.4973849048591698e-33, r[e]);
#pragma unroll
  for (int e = 0; e < NV; ++e) z[e] = r[e] * r[e];
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    ps[e] = ffp_fma(z[e], 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    pc[e] = ffp_fma(z[e], -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  }
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    ps[e] = ffp_fma(z[e], ps[e], 2.75573137070700676789e-06);
    pc[e] = ffp_fma(z[e], pc[e], -2.75573143513906633035e-07);
  }
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    ps[e] = ffp_fma(z[e], ps[e], -1.98412698298579493134e-04);
    pc[e] = ffp_fma(z[e], pc[e], 2.48015872894767294178e-05);
  }
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    ps[e] = ffp_fma(z[e], ps[e], 8.33333333332248946124e-03);
    pc[e] = ffp_fma(z[e], pc[e], -1.38888888888741095749e-03);
  }
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    ps[e] = ffp_fma(z[e], ps[e], -1.66666666666666324348e-01);
    pc[e] = ffp_fma(z[e], pc[e], 4.16666666666666019037e-02);
  }
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    const double s = ffp_fma(r[e] * z[e], ps[e], r[e]);
    const double c = ffp_fma(z[e], ffp_fma(z[e], pc[e], -0.5), 1.0);
    const bool swap = (q[e] & 1) != 0;
    const double so = swap ? c : s;
    const double co = swap ? s : c;
    const uint32_t fs = ((uint32_t)q[e] & 2u) << 30, fc = ((uint32_t)(q[e] + 1) & 2u) << 30;
#if defined(__CUDA_ARCH__)
    sp[e] = __hiloint2double(__double2hiint(so) ^ (int)fs, __double2loint(so));
    cp[e] = __hiloint2double(__double2hiint(co) ^ (int)fc, __double2loint(co));
#else
    uint64_t bs, bc;
    std::memcpy(&bs, &so, 8);
    std::memcpy(&bc, &co, 8);
    bs ^= (uint64_t)fs << 32;
    bc ^= (uint64_t)fc << 32;
    std::memcpy(&sp[e], &bs, 8);
    std::memcpy(&cp[e], &bc, 8);
#endif
  }
}

}  // namespace ffp
