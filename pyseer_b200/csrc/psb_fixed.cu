// placeholder, replaced below
#include "psb_internal.cuh"
extern "C" int psb_fixed_setup(psb_ctx *c, int32_t N, int32_t q, const double *Z, const double *y, int32_t cont, double a, double b) { psb_set_error("not built yet"); return PSB_ERR_UNSUPPORTED; }
extern "C" int psb_fit_null(psb_ctx *ctx, int32_t n_samples, int32_t q, const double *Z, const double *y, int32_t continuous, int32_t firth, double *out_params, double *out_bse, double *out_llf, uint32_t *out_status) { return PSB_ERR_UNSUPPORTED; }
extern "C" int psb_run_fixed(psb_ctx *c, const psb_params *p) { return PSB_ERR_UNSUPPORTED; }
