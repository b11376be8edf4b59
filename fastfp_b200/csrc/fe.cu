// Fe-statistic (Ellis, Siemens & Creighton 2012, the coherent Earth-term statistic; the reference lists it as a
// to-do: README.md:23) from the same per-(pulsar, frequency) inner products the Fp sweep forms
// (fastfp/fastfp.py:81-88): with the antenna patterns F+_p, Fx_p of a sky position the four templates of pulsar p are
//   A_p = [F+ s, F+ c, Fx s, Fx c],   s, c = sin, cos(((2 pi) f) t)          (f^(-1/3) cancels as in Fp)
//   N = sum_p [F+ N_p ; Fx N_p]                       N_p = [(s|r), (c|r)]
//   M = sum_p [[F+^2 M_p, F+ Fx M_p], [F+ Fx M_p, Fx^2 M_p]]      M_p = [[(s|s), (s|c)], [(s|c), (c|c)]]
//   Fe = 1/2 N^T M^-1 N
// so a sky scan costs one sweep (fp_sweep*_kernel with the `inner` output) plus this combine kernel: one thread per
// (sky position, frequency), pulsars summed in pulsar order, general 4x4 solve with partial pivoting (np.linalg.solve).
#include "../../include/fastfp_b200.h"
#include "ffp_internal.cuh"

namespace ffp {

__global__ void fe_combine_kernel(const double* __restrict__ inner, int P, int64_t F, const double* __restrict__ fplus,
                                  const double* __restrict__ fcross, int64_t S, double* __restrict__ out, int64_t out_ld) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t k = blockIdx.y;
  if (f >= F || k >= S) return;
  double N[4] = {0, 0, 0, 0};
  double M[4][4] = {};
  for (int p = 0; p < P; ++p) {
    const double* q = inner + ((size_t)p * F + f) * 5;
    const double ss = q[0], sc = q[1], cc = q[2], sr = q[3], cr = q[4];
    const double fp = fplus[(size_t)k * P + p], fx = fcross[(size_t)k * P + p];
    N[0] += fp * sr; N[1] += fp * cr; N[2] += fx * sr; N[3] += fx * cr;
    const double pp = fp * fp, px = fp * fx, xx = fx * fx;
    M[0][0] += pp * ss; M[0][1] += pp * sc; M[1][1] += pp * cc;
    M[0][2] += px * ss; M[0][3] += px * sc; M[1][2] += px * sc; M[1][3] += px * cc;
    M[2][2] += xx * ss; M[2][3] += xx * sc; M[3][3] += xx * cc;
  }
  M[1][0] = M[0][1]; M[2][0] = M[0][2]; M[3][0] = M[0][3]; M[2][1] = M[1][2]; M[3][1] = M[1][3]; M[3][2] = M[2][3];
  // x = M^-1 N: Gaussian elimination with partial pivoting (fully unrolled: everything stays in registers)
  double b[4] = {N[0], N[1], N[2], N[3]};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    double big = fabs(M[c][c]);
#pragma unroll
    for (int r = c + 1; r < 4; ++r) {
      const double v = fabs(M[r][c]);
      if (v > big) { big = v; piv = r; }
    }
#pragma unroll
    for (int r = c + 1; r < 4; ++r) {
      if (r == piv) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double t = M[c][j]; M[c][j] = M[r][j]; M[r][j] = t; }
        const double t = b[c]; b[c] = b[r]; b[r] = t;
      }
    }
#pragma unroll
    for (int r = c + 1; r < 4; ++r) {
      const double l = M[r][c] / M[c][c];
#pragma unroll
      for (int j = c + 1; j < 4; ++j) M[r][j] -= l * M[c][j];
      b[r] -= l * b[c];
    }
  }
  double x[4];
#pragma unroll
  for (int r = 3; r >= 0; --r) {
    double acc = b[r];
#pragma unroll
    for (int j = r + 1; j < 4; ++j) acc -= M[r][j] * x[j];
    x[r] = acc / M[r][r];
  }
  out[(size_t)k * out_ld + f] = 0.5 * (N[0] * x[0] + N[1] * x[1] + N[2] * x[2] + N[3] * x[3]);
}

int launch_fe_combine(const double* d_inner, int P, int64_t F, const double* d_fplus, const double* d_fcross, int64_t S,
                      double* d_out, int64_t out_ld, cudaStream_t st) {
  if (S > 65535) { set_error("at most 65535 sky positions per call"); return FASTFP_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)((F + 127) / 128), (unsigned)S);
  fe_combine_kernel<<<grid, 128, 0, st>>>(d_inner, P, F, d_fplus, d_fcross, S, d_out, out_ld);
  g_launches += 1;
  FFP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ffp
