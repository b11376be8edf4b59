// placeholder (nmfp path under construction)
#include "../../include/fastfp_b200.h"
#include "ffp_internal.cuh"
namespace ffp {
int nmfp_pack_finish(fastfp_pack*, const double*, const double*, const double*, const double*,
                     const double*, const double*, cudaStream_t) { set_error("nmfp not built yet"); return FASTFP_ERR_UNSUPPORTED; }
int nmfp_sweep_impl(const fastfp_pack*, const double*, int64_t, const double*, int64_t, double*, cudaStream_t) { set_error("nmfp not built yet"); return FASTFP_ERR_UNSUPPORTED; }
int powerlaw_phiinv_impl(const fastfp_pack*, const double* const*, const double*, const double*, int64_t, const double*, int64_t, const double*, const double*, double*, cudaStream_t) { set_error("nmfp not built yet"); return FASTFP_ERR_UNSUPPORTED; }
}
