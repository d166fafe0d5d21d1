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


// extra arguments of the fast Logit path (psb_fixed_fast.cu)
struct FxFast {
    const double *Zi;     // [Wn][Q-1][32] covariate columns 1..Q-1 interleaved per 32-sample word (fp64)
    const float *Zf;      // the same in fp32
    const double *W0;     // [Wn * 32] null-model weights pi0 (1 - pi0), 0 on padding samples
    const uint32_t *vbits; // [Wn padded to whole chunks] sample-valid bits (0 on padding words)
    const double *Hzz0;   // packed lower triangle of Z'W0Z (Q columns; unit diagonal on padding columns)
    double zmax[FX_MAXP]; // max |z_c| per column
    int32_t *slow_list;   // variants handed to the reference-faithful kernel (length in counters[6])
};
int psb_fixed_fast_setup(psb_ctx *c, const double *Z, const double *warm);
void psb_fixed_fast_free(psb_ctx *c);
int psb_fixed_fast_launch(psb_ctx *c, const FxArgs &a, int n);

// Singular information matrices in Firth regression.  model.fit_firth calls np.linalg.pinv on
// -hessian (model.py:450) and firth_likelihood takes log(det(-hessian)) (model.py:410): with an
// exactly collinear design (the reference's own test has variant == covariate,
// tests/model_test.py:316-338) neither raises -- pinv drops the null directions (singular values
// <= 1e-15 sigma_max), det is 0 and the penalised likelihood becomes -log(0) = +inf, which no
// comparison of the step-halving loop ever finds "worse".  This routine restates both for a
// symmetric matrix given as its packed lower triangle: cyclic Jacobi eigendecomposition in the
// caller's scratch (A, Q: P x P doubles each, any address space; p_active = columns in use), V = sum_{|l_i| > cut} q_i q_i' / l_i
// (packed, when V != nullptr) and the return value log det = sum log l_i, -inf when an eigenvalue
// falls under the pinv cut-off, NaN when one is negative beyond it.  A rare path (the Cholesky
// factorisation of the caller failed or met a pivot below 1e-13 of the diagonal): one thread.
static __device__ __noinline__ double psb_sym_pinv_logdet(const double *Hp, double *Vp, int P, int p_active, double *A,
                                                    double *Q) {
    for (int i = 0; i < P; ++i)
        for (int j = 0; j < P; ++j) {
            A[i * P + j] = Hp[i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i];
            Q[i * P + j] = i == j ? 1.0 : 0.0;
        }
    bool finite = true;
    for (int e = 0; e < P * P; ++e) finite = finite && isfinite(A[e]);
    if (!finite) {
        if (Vp) for (int e = 0; e < P * (P + 1) / 2; ++e) Vp[e] = NAN;
        return NAN;
    }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < P; ++i) {
            dia = fma(A[i * P + i], A[i * P + i], dia);
            for (int j = 0; j < i; ++j) off = fma(A[i * P + j], A[i * P + j], off);
        }
        if (off <= 1e-36 * dia || off == 0.0) break;
        for (int p = 0; p < P - 1; ++p)
            for (int q = p + 1; q < P; ++q) {
                const double apq = A[p * P + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * P + q] - A[p * P + p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                const double c = rsqrt(fma(t, t, 1.0)), sn = t * c;
                for (int k = 0; k < P; ++k) {            // columns p, q
                    const double akp = A[k * P + p], akq = A[k * P + q];
                    A[k * P + p] = c * akp - sn * akq;
                    A[k * P + q] = sn * akp + c * akq;
                }
                for (int k = 0; k < P; ++k) {            // rows p, q
                    const double apk = A[p * P + k], aqk = A[q * P + k];
                    A[p * P + k] = c * apk - sn * aqk;
                    A[q * P + k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < P; ++k) {
                    const double qkp = Q[k * P + p], qkq = Q[k * P + q];
                    Q[k * P + p] = c * qkp - sn * qkq;
                    Q[k * P + q] = sn * qkp + c * qkq;
                }
            }
    }
    // (columns >= p_active are the solver's padding: unit diagonal, decoupled, never rotated)
    double lmax = 0.0;
    for (int i = 0; i < p_active; ++i) lmax = fmax(lmax, fabs(A[i * P + i]));
    const double cut = 1e-15 * lmax;                     // numpy.linalg.pinv default rcond
    double logdet = 0.0;
    for (int i = 0; i < p_active; ++i) {
        const double l = A[i * P + i];
        if (fabs(l) <= cut) logdet = isnan(logdet) ? logdet : -INFINITY;
        else if (l < 0.0) logdet = NAN;
        else logdet += log(l);
    }
    if (Vp)
        for (int i = 0; i < P; ++i)
            for (int j = 0; j <= i; ++j) {
                double acc = 0.0;
                for (int k = 0; k < P; ++k) {
                    const double l = A[k * P + k];
                    if (fabs(l) > cut || k >= p_active) acc = fma(Q[i * P + k] / l, Q[j * P + k], acc);
                }
                Vp[i * (i + 1) / 2 + j] = acc;
            }
    return logdet;
}

// modes of the generic kernel
#define FXG_LOGIT 0     // variant fit: Logit Newton + LRT, failures pushed to the Firth list
#define FXG_FIRTH 1     // Firth regression over the Firth list
#define FXG_LINEAGE 2   // model.fit_lineage_effect: response = the variant, argmax Wald
#define FXG_NULL 3      // null fit (no variant column); a.null_out; a.has_x == 0
#define FXG_NULL_FIRTH 4

int psb_fixed_gen_launch(psb_ctx *c, const FxArgs &a, int mode, int n, int lineage_mode,
                         int n_lin, int32_t *lineage_out);
