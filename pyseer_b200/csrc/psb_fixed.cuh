// psb_fixed.cuh -- argument block shared by the fixed-effects kernels (psb_fixed.cu: register-
// resident solver templates for designs up to FX_MAXP columns; psb_fixed_gen.cu: generic
// shared-memory solver for wider designs).
#pragma once
#include "psb_internal.cuh"

#define FX_MAXP 12          // widest design of the register-resident solver templates
#define FX_GEN_MAXP 64      // widest design of the generic shared-memory solver

struct FxArgs {
    const uint32_t *bits;
    const double *Z;          // [q][Npad]
    const uint32_t *y1;       // phenotype == 1 bits
    const uint32_t *valid;    // sample < N bits
    int Wrow, Wn, N, Npad;
    int q;                    // columns of Z
    int has_x;                // 0: null model (no variant column)
    double start0;            // log(mean(y) / (1 - mean(y)))
    int use_warm;             // start Newton from the null-model parameters (exact fallback below)
    double warm[FX_MAXP];     // null-model parameters (Z order)
    // first Newton step of the warm attempt from masked sums (linear tensor tile): per variant
    // sums[v * sums_ld + a] = sum over carriers of w0 z_a (a < q), [q] = sum of y - pi0;
    // HzzInv = (Z' W0 Z)^-1 at the null fit (q x q).  Null when not available.
    const double *sums, *HzzInv;
    int sums_ld;
    double null_llf, null_firth, lrt_pvalue;
    // outputs (indexed by variant id)
    double *pvalue, *beta, *bse, *intercept, *betas;
    uint32_t *flags;
    int *counters;            // [2]: lrt-filtered, [3]: firth list length
    int32_t *firth_list;
    // null-fit outputs (has_x == 0): params[q], bse[q], llf, status
    double *null_out;
};


// modes of the generic kernel
#define FXG_LOGIT 0     // variant fit: Logit Newton + LRT, failures pushed to the Firth list
#define FXG_FIRTH 1     // Firth regression over the Firth list
#define FXG_LINEAGE 2   // model.fit_lineage_effect: response = the variant, argmax Wald
#define FXG_NULL 3      // null fit (no variant column); a.null_out; a.has_x == 0
#define FXG_NULL_FIRTH 4

int psb_fixed_gen_launch(psb_ctx *c, const FxArgs &a, int mode, int n, int lineage_mode,
                         int n_lin, int32_t *lineage_out);
