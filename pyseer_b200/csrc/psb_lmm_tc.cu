// placeholder, replaced below
#include "psb_internal.cuh"
int psb_lmm_tc_setup(psb_ctx *c) { psb_set_error("int8 tensor path not built yet"); return PSB_ERR_UNSUPPORTED; }
int psb_lmm_tc_run(psb_ctx *c, int n) { return PSB_ERR_UNSUPPORTED; }
void psb_lmm_tc_free(psb_ctx *c) {}
