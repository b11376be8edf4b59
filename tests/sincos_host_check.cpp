// Host check of ffp::sincos_cw against long double sinl/cosl (build: g++ -O2 -mfma).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "../fastfp_b200/csrc/ffp_sincos.cuh"

int main(int argc, char** argv) {
  const long N = argc > 1 ? atol(argv[1]) : 2000000;
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> U(-1.0e5, 1.0e5), V(-50.0, 50.0);
  double max_abs = 0, max_ulp = 0, worst_x = 0;
  for (long i = 0; i < N; ++i) {
    double x = (i & 1) ? U(rng) : V(rng);
    if (i < 64) x = (i - 32) * 0.7853981633974483;  // multiples of pi/4, the reduction's edge
    double s, c;
    ffp::sincos_cw(x, &s, &c);
    const long double ts = sinl((long double)x), tc = cosl((long double)x);
    const double es = fabs((double)(s - ts)), ec = fabs((double)(c - tc));
    const double us = es / (fabs((double)ts) > 0 ? ldexp(1.0, ilogb((double)ts) - 52) : 1e-300);
    const double uc = ec / (fabs((double)tc) > 0 ? ldexp(1.0, ilogb((double)tc) - 52) : 1e-300);
    if (es > max_abs) { max_abs = es; worst_x = x; }
    if (ec > max_abs) { max_abs = ec; worst_x = x; }
    if (fabs((double)ts) > 1e-3 && us > max_ulp) max_ulp = us;
    if (fabs((double)tc) > 1e-3 && uc > max_ulp) max_ulp = uc;
  }
  printf("max_abs_err %.3e max_ulp_err %.3f worst_x %.17g\n", max_abs, max_ulp, worst_x);
  // the lockstep form used by the tensor-core sweep's producers must agree with the scalar form bit for bit
  long mismatch = 0;
  for (long i = 0; i < N / 4; ++i) {
    double x[4], s4[4], c4[4];
    for (int e = 0; e < 4; ++e) x[e] = (i & 1) ? U(rng) : V(rng);
    ffp::sincos_cw_n<4>(x, s4, c4);
    for (int e = 0; e < 4; ++e) {
      double s, c;
      ffp::sincos_cw(x[e], &s, &c);
      if (memcmp(&s, &s4[e], 8) || memcmp(&c, &c4[e], 8)) ++mismatch;
    }
  }
  printf("lockstep_mismatch %ld\n", mismatch);
  return 0;
}
