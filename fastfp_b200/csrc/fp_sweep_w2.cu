// sweep instantiations: 40 < m <= 80 (two warp rows, 64 frequencies per CTA) -- the m = 72 case
#include "fp_sweep_kernel.cuh"
namespace ffp {
int dispatch_sweep_w2(const fastfp_pack* pk, const Group& g, const SweepArgs& a, bool nmfp, cudaStream_t st) {
  FFP_SWEEP_CASE(6, 4, 2, 32) FFP_SWEEP_CASE(7, 4, 2, 32) FFP_SWEEP_CASE(8, 4, 2, 32)
  FFP_SWEEP_CASE(9, 4, 2, 32) FFP_SWEEP_CASE(10, 4, 2, 32)
  set_error("no sweep kernel for this configuration (w2)");
  return -3;
}
}  // namespace ffp
