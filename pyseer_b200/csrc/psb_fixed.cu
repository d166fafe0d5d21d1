// psb_fixed.cu -- fixed-effects (SEER) model: batched OLS t-test, batched Logit Newton with
// likelihood-ratio test, Firth-penalised fallback, and the null-model fits.
//
// Reference being replaced, per variant:
//   model.fixed_effects_regression  model.py:202-394
//   model.fit_firth / firth_likelihood  model.py:397-504
//   model.fit_null                  model.py:73-148
//   statsmodels Logit.fit(method='newton') / OLS.fit()  (third party; behaviour restated in
//   oracle/fixed_oracle.py, which pins it to the goldens of tests/model_test.py)
//
// Design [1, k, m, c] (model.py:274-297) is held as Z = [1, m, c] (q columns, shared by all
// variants, column-contiguous in HBM and L1-resident) plus the variant's bit row: the k
// column never exists in memory.  Internally the parameter order is (Z_0..Z_{q-1}, k).
//
// One warp fits one variant.  Lane l owns samples 32 t + l; every lane accumulates the full
// X'WX (lower triangle) and the score in registers, a butterfly reduction leaves the sums in
// all lanes, and each lane runs the small Cholesky solve redundantly (no divergence, no
// shared memory).  Everything is fp64: the Newton / Firth iterations follow the reference's
// stopping rules literally (35 steps, |dbeta| <= 1e-8; lagged 1e-4 test and step halving for
// Firth), so the iterates -- not just the limits -- agree with the reference.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "psb_internal.cuh"
#include "psb_math.cuh"

#include "psb_fixed.cuh"
#include "psb_fixed_dev.cuh"

// Row of the design for sample i (internal order: Z_0 = 1, Z_1.., then k, then zero padding).
template <int PP>
__device__ __forceinline__ void fx_row(const FxArgs &a, int i, uint32_t xbit, double (&z)[PP]) {
    z[0] = 1.0;
#pragma unroll
    for (int c = 1; c < PP; ++c) {
        double v = 0.0;
        if (c < a.q) v = __ldg(a.Z + (size_t)c * a.Npad + i);
        else if (c == a.q && a.has_x) v = (double)xbit;
        z[c] = v;
    }
}

// beta'z with four interleaved accumulation chains (shorter dependency chain than one FMA
// chain of length PP; the summation order is irrelevant at the 1e-6 tolerance).
template <int PP, class BetaT>
__device__ __forceinline__ double fx_dot(const BetaT &beta, const double (&z)[PP]) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int c = 0; c < PP; c += 4) {
        s0 = fma(beta[c], z[c], s0);
        if (c + 1 < PP) s1 = fma(beta[c + 1], z[c + 1], s1);
        if (c + 2 < PP) s2 = fma(beta[c + 2], z[c + 2], s2);
        if (c + 3 < PP) s3 = fma(beta[c + 3], z[c + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}

// One pass over the samples at parameters beta: X'WX (packed), score X'(y - pi),
// max |y - pi| and the log-likelihood.  All lanes return the full sums.
// (BetaT: a register array or a pointer to the warp's shared-memory copy of the parameters)
template <int PP, class BetaT>
__device__ __forceinline__ void fx_eval(const FxArgs &a, const uint32_t *xrow, int lane,
                                        const BetaT &beta, double (&H)[Tri<PP>::SIZE],
                                        double (&g)[PP], double &maxdev, double &llf,
                                        const bool WITH_LLF, const uint32_t *yrow = nullptr) {
    if (!yrow) yrow = a.y1;
#pragma unroll
    for (int e = 0; e < Tri<PP>::SIZE; ++e) H[e] = 0.0;
#pragma unroll
    for (int c = 0; c < PP; ++c) g[c] = 0.0;
    maxdev = 0.0;
    llf = 0.0;
    for (int w = 0; w < a.Wn; ++w) {
        const uint32_t vw = __ldg(a.valid + w);
        if (!((vw >> lane) & 1u)) continue;
        const uint32_t xw = a.has_x ? __ldg(xrow + w) : 0u;
        const uint32_t yw = __ldg(yrow + w);
        const int i = w * 32 + lane;
        double z[PP];
        fx_row<PP>(a, i, (xw >> lane) & 1u, z);
        const double eta = fx_dot<PP, BetaT>(beta, z);
        const double y = (double)((yw >> lane) & 1u);
        // statsmodels' own formulas (Logit.cdf = 1 / (1 + exp(-x)), hessian weight L (1 - L)):
        // in saturated fits the weight rounds to exactly 0 for eta > ~36.7 and the reference
        // then meets an exactly singular information matrix -- reproduced, not "improved"
        const double ex = exp(-eta);
        const double pi = 1.0 / (1.0 + ex);
        const double wgt = pi * (1.0 - pi);
        const double r = y - pi;
        maxdev = fmax(maxdev, fabs(r));
        if (WITH_LLF) {
            // log cdf((2y-1) eta) = min(s, 0) - log1p(exp(-|s|)),  s = (2y-1) eta
            const double s = y > 0.5 ? eta : -eta;
            llf += fmin(s, 0.0) - log1p(eta >= 0.0 ? ex : 1.0 / ex);
        }
#pragma unroll
        for (int c = 0; c < PP; ++c) {
            g[c] = fma(r, z[c], g[c]);
            const double wz = wgt * z[c];
#pragma unroll
            for (int d = 0; d < PP; ++d)
                if (d <= c) H[Tri<PP>::at(c, d)] = fma(wz, z[d], H[Tri<PP>::at(c, d)]);
        }
    }
#pragma unroll
    for (int e = 0; e < Tri<PP>::SIZE; ++e) H[e] = warp_sum(H[e]);
#pragma unroll
    for (int c = 0; c < PP; ++c) g[c] = warp_sum(g[c]);
    maxdev = warp_max(maxdev);
    if (WITH_LLF) llf = warp_sum(llf);
    // padding columns (and the absent k column of a null fit): unit diagonal, zero score
    const int p = a.q + (a.has_x ? 1 : 0);
#pragma unroll
    for (int c = 0; c < PP; ++c)
        if (c >= p) H[Tri<PP>::at(c, c)] = 1.0;
}

// log-likelihood only
template <int PP, class BetaT>
__device__ __forceinline__ double fx_loglike(const FxArgs &a, const uint32_t *xrow, int lane,
                                             const BetaT &beta) {
    double llf = 0.0;
    for (int w = 0; w < a.Wn; ++w) {
        const uint32_t vw = __ldg(a.valid + w);
        if (!((vw >> lane) & 1u)) continue;
        const uint32_t xw = a.has_x ? __ldg(xrow + w) : 0u;
        const uint32_t yw = __ldg(a.y1 + w);
        double z[PP];
        fx_row<PP>(a, w * 32 + lane, (xw >> lane) & 1u, z);
        const double eta = fx_dot<PP, BetaT>(beta, z);
        const double s = ((yw >> lane) & 1u) ? eta : -eta;
        llf += fmin(s, 0.0) - log1p(exp(-fabs(s)));
    }
    return warp_sum(llf);
}

__device__ __forceinline__ void fx_write_failed(const FxArgs &a, int v, uint32_t f) {
    a.flags[v] = f;
}

// ---------------------------------------------------------------------------------------
// Logit Newton (statsmodels Logit.fit(method='newton'), call site model.py:328-330)
// ---------------------------------------------------------------------------------------
template <int PP, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_fixed_logit(FxArgs a, const int32_t *__restrict__ idx, int n_tested) {
    // the parameter vector lives in shared memory (one copy per warp, read by broadcast in the
    // sample loop): 2 PP registers less pressure on the X'WX accumulators
    __shared__ double s_beta[4][PP];
    double *beta = s_beta[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    const int p = a.q + (a.has_x ? 1 : 0);
    for (int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_tested; t += warps_total) {
        const int v = a.has_x ? idx[t] : 0;
        uint32_t f = a.has_x ? a.flags[v] : 0u;
        if (f & PSB_F_MISSING_DATA) continue;                       // model.py:371-377
        if (f & PSB_F_BAD_CHISQ) {                                  // model.py:326: straight to Firth
            if (lane == 0) a.firth_list[atomicAdd(&a.counters[3], 1)] = v;
            continue;
        }
        const uint32_t *xrow = a.bits + (size_t)v * a.Wrow;
        double H[Tri<PP>::SIZE], g[PP];
        double maxdev, llf = NAN, maxstep = INFINITY;
        uint32_t fail = 0;
        int it = 0;
        bool have_llf = false;
        int n_eval = 0;
        double fast_bse = -1.0;
        const double inv_n = 1.0 / (double)a.N;
        // The log-likelihood is concave, so a converged Newton run ends at the same (unique)
        // maximiser whatever the start.  Attempt 0 starts from the null-model parameters and
        // saves about half of the iterations; unless it converges cleanly within 12 steps the
        // fit is redone from the reference's start vector with the reference's 35-step rule
        // (attempt 1), so separation / non-convergence are flagged exactly as statsmodels does.
        for (int attempt = (a.use_warm && a.has_x) ? 0 : 1; attempt < 2; ++attempt) {
            __syncwarp();
            if (lane < PP) beta[lane] = (attempt == 0) ? (lane < a.q ? a.warm[lane] : 0.0)
                                                       : (lane == 0 ? a.start0 : 0.0);
            __syncwarp();
            if (attempt == 0 && a.sums) {
                // At the null parameters the Z-block of X'WX and the Z-part of the score are the
                // same for every variant (Hzz, 0); the k-border is a set of masked sums that the
                // linear tensor tile already delivered.  First Newton step in closed form:
                //   schur = hxx - hx' Hzz^-1 hx,  d_k = g_k / schur,  d_z = -Hzz^-1 hx d_k
                const double *sv = a.sums + (size_t)v * a.sums_ld;
                double hx[PP], tz[PP];
#pragma unroll
                for (int c = 0; c < PP; ++c) hx[c] = (c < a.q) ? sv[c] : 0.0;
                const double gx = sv[a.q];
                double quad = 0.0;
#pragma unroll
                for (int c = 0; c < PP; ++c) {
                    double acc = 0.0;
                    if (c < a.q) {
#pragma unroll
                        for (int d = 0; d < PP; ++d)
                            if (d < a.q) acc = fma(__ldg(a.HzzInv + c * a.q + d), hx[d], acc);
                    }
                    tz[c] = acc;
                    quad = fma(acc, hx[c], quad);
                }
                const double dk = gx / (hx[0] - quad);        // hxx = sum of w0 over carriers = hx[0]
                if (isfinite(dk)) {
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < PP; ++c) {
                        if (lane == 0 && c < a.q) beta[c] -= tz[c] * dk;
                        if (lane == 0 && c == a.q) beta[c] = dk;
                    }
                    __syncwarp();
                }
            }
            const int maxit = attempt == 0 ? 12 : 35;
            fail = 0;
            it = 0;
            maxstep = INFINITY;
            bool converged = false;
            for (;;) {
                const bool want_llf = maxstep <= 1e-3;       // very likely the last evaluation
                fx_eval<PP>(a, xrow, lane, beta, H, g, maxdev, llf, want_llf);
                ++n_eval;
                have_llf = want_llf;
                if (it > 0 && maxdev <= 1e-8) { fail = PSB_F_PERFECT_SEP; break; }   // _check_perfect_pred
                // reference rule |dbeta| <= 1e-8; the warm attempt stops one quadratic step
                // earlier (|dbeta| <= 1e-6 leaves an error of order 1e-12)
                if (it > 0 && !(maxstep > (attempt == 0 ? 1e-6 : 1e-8))) { converged = true; break; }
                if (it >= maxit) break;                                              // maxiter
                // (X'WX/n - 1e-10 I) step = score/n: statsmodels' ridge sign, see fx_ldl
#pragma unroll
                for (int e = 0; e < Tri<PP>::SIZE; ++e) H[e] *= inv_n;
#pragma unroll
                for (int c = 0; c < PP; ++c) {
                    if (c < p) H[Tri<PP>::at(c, c)] -= 1e-10;
                    g[c] *= inv_n;
                }
                if (!fx_ldl<PP>(H)) { fail = PSB_F_MATRIX_INV; break; }
                double gs[PP];                       // score / n, kept for the llf correction
#pragma unroll
                for (int c = 0; c < PP; ++c) gs[c] = g[c];
                fx_ldl_solve<PP>(H, g);
                maxstep = 0.0;
                double gd = 0.0;
                __syncwarp();
#pragma unroll
                for (int c = 0; c < PP; ++c) {
                    if (lane == 0) beta[c] += g[c];
                    maxstep = fmax(maxstep, fabs(g[c]));
                    gd = fma(gs[c], g[c], gd);
                }
                __syncwarp();
                if (isnan(maxstep)) { fail = PSB_F_MATRIX_INV; break; }
                ++it;
                if (attempt == 0 && a.has_x && want_llf && maxstep <= 1e-7) {
                    // The step just taken is below 1e-7: the fit has converged, and everything
                    // the result needs follows from THIS evaluation without another pass --
                    //   llf(beta + d) = llf(beta) + g'd/2 + O(d^3)   (since (X'WX) d = g),
                    //   bse from the factored matrix (X'WX differs by O(d) relative).
                    double e[PP];
#pragma unroll
                    for (int c = 0; c < PP; ++c) e[c] = (c == a.q) ? 1.0 : 0.0;
                    fx_ldl_solve<PP>(H, e);
                    double var_x = 0.0;
#pragma unroll
                    for (int c = 0; c < PP; ++c)
                        if (c == a.q) var_x = e[c] * inv_n;
                    if (var_x > 0.0 && isfinite(var_x)) {
                        fast_bse = sqrt(var_x);
                        llf += 0.5 * (double)a.N * gd;
                        converged = true;
                        break;
                    }
                }
            }
            if (attempt == 0 && converged && !fail) break;     // accept the warm-started fit
            fast_bse = -1.0;
        }
        if (lane == 0 && a.has_x) atomicAdd(&a.counters[4], n_eval);     // measured, for the roofline
        double bse_x = NAN;
        double bse_all[PP];
        if (!fail && fast_bse > 0.0) {
            bse_x = fast_bse;
        } else if (!fail) {
            // bse = sqrt(diag(inv(X'WX))) at the final parameters (H holds X'WX there)
            // numpy.linalg.inv: 'Singular matrix'.  Variant fits keep the plain test (quasi-separated fits
            // have legitimately tiny pivots, and the reference's LU carries on through them); the null
            // fit, where the answer decides between Newton's result and the Powell fallback
            // (model.py:132-137), recognises numerically dependent columns (fx_chol_firth).
            if (!(a.has_x ? fx_chol<PP>(H) : fx_chol_firth<PP>(H))) {
                fail = PSB_F_MATRIX_INV;
            } else if (a.has_x) {
                double e[PP];
#pragma unroll
                for (int c = 0; c < PP; ++c) e[c] = (c == a.q) ? 1.0 : 0.0;
                fx_chol_solve<PP>(H, e);
#pragma unroll
                for (int c = 0; c < PP; ++c)
                    if (c == a.q) bse_x = sqrt(e[c]);
            } else {
#pragma unroll
                for (int c0 = 0; c0 < PP; ++c0) {
                    double e[PP];
#pragma unroll
                    for (int c = 0; c < PP; ++c) e[c] = (c == c0) ? 1.0 : 0.0;
                    fx_chol_solve<PP>(H, e);
                    bse_all[c0] = sqrt(e[c0]);
                }
            }
        }
        if (!a.has_x) {
            // null fit: params, bse, llf, status
            if (fail) llf = NAN;
            else if (!have_llf) llf = fx_loglike<PP>(a, xrow, lane, beta);
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < PP; ++c)
                    if (c < a.q) {
                        a.null_out[c] = beta[c];
                        a.null_out[a.q + c] = fail ? NAN : bse_all[c];
                    }
                a.null_out[2 * a.q] = llf;
                a.null_out[2 * a.q + 1] = (double)fail;
                a.null_out[2 * a.q + 2] = (double)it;
            }
            continue;
        }
        if (!fail && bse_x > 3.0) fail = PSB_F_HIGH_BSE;                // model.py:332-334
        if (fail) {
            if (lane == 0) {
                a.flags[v] = f | fail;
                a.firth_list[atomicAdd(&a.counters[3], 1)] = v;
            }
            continue;
        }
        if (!have_llf) llf = fx_loglike<PP>(a, xrow, lane, beta);
        if (lane == 0) fx_publish<PP>(a, v, f, beta, bse_x, llf, a.null_llf);
    }
}

// ---------------------------------------------------------------------------------------
// Firth regression (model.fit_firth, model.py:414-504)
// ---------------------------------------------------------------------------------------
template <int PP>
__global__ void __launch_bounds__(128)
k_fixed_firth(FxArgs a, int n_list) {
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    const int p = a.q + (a.has_x ? 1 : 0);
    for (int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_list; t += warps_total) {
        const int v = a.has_x ? a.firth_list[t] : 0;
        uint32_t f = a.has_x ? (a.flags[v] | PSB_F_FIRTH_USED) : 0u;
        const uint32_t *xrow = a.bits + (size_t)v * a.Wrow;
        double beta[PP], prev[PP];
#pragma unroll
        for (int c = 0; c < PP; ++c) beta[c] = prev[c] = 0.0;
        beta[0] = a.start0;
        double H[Tri<PP>::SIZE], V[Tri<PP>::SIZE], g[PP];
        double maxdev, llf_cur;
        bool ok = true, converged = false;
        // state at the current iterate: H = X'WX, llf, FL = -(llf + 0.5 log det H)
        fx_eval<PP>(a, xrow, lane, beta, H, g, maxdev, llf_cur, true);
        double hxx_cur = 0.0;
#pragma unroll
        for (int c = 0; c < PP; ++c)
            if (c == a.q) hxx_cur = H[Tri<PP>::at(c, c)];
        double fl_cur, fitll = NAN, hxx_fit = NAN;
        double last_step_norm = INFINITY;      // || betas[i] - betas[i-1] ||
        int n_iter = 0, n_halve = 0;
        for (int i = 0; i < 1000 && ok; ++i) {
            ++n_iter;
            double logdet;
            if (fx_chol_firth<PP>(H)) {
                logdet = fx_chol_logdet<PP>(H);
                fx_inverse_from_chol<PP>(H, V);        // V = pinv(-hessian), model.py:450
            } else {
                // singular information matrix: np.linalg.pinv / det carry on (model.py:450, :410);
                // the factorisation ran in place, so X'WX is evaluated again first
                fx_eval<PP>(a, xrow, lane, beta, H, g, maxdev, llf_cur, true);
                if (lane == 0 && a.has_x) atomicAdd(&a.counters[8], 1);
                logdet = fx_singular<PP, true>(H, V, p);
                if (isnan(V[0])) { ok = false; break; }
            }
            fl_cur = -(llf_cur + 0.5 * logdet);
            // U = X'(y - pi + h (1/2 - pi)),  h_i = w_i x_i' V x_i   (model.py:455-466)
            double U[PP];
#pragma unroll
            for (int c = 0; c < PP; ++c) U[c] = 0.0;
            for (int w = 0; w < a.Wn; ++w) {
                const uint32_t vw = __ldg(a.valid + w);
                if (!((vw >> lane) & 1u)) continue;
                const uint32_t xw = a.has_x ? __ldg(xrow + w) : 0u;
                const uint32_t yw = __ldg(a.y1 + w);
                double z[PP];
                fx_row<PP>(a, w * 32 + lane, (xw >> lane) & 1u, z);
                const double eta = fx_dot<PP>(beta, z);
                const double pi = 1.0 / (1.0 + exp(-eta));
                const double wgt = pi * (1.0 - pi);
                double quad = 0.0;
#pragma unroll
                for (int c = 0; c < PP; ++c) {
                    double s = 0.0;
#pragma unroll
                    for (int d = 0; d < PP; ++d)
                        s = fma(c >= d ? V[Tri<PP>::at(c, d)] : V[Tri<PP>::at(d, c)], z[d], s);
                    quad = fma(s, z[c], quad);
                }
                const double h = wgt * quad;
                const double y = (double)((yw >> lane) & 1u);
                const double r = y - pi + h * (0.5 - pi);
#pragma unroll
                for (int c = 0; c < PP; ++c) U[c] = fma(r, z[c], U[c]);
            }
            double cand[PP];
#pragma unroll
            for (int c = 0; c < PP; ++c) U[c] = warp_sum(U[c]);
#pragma unroll
            for (int c = 0; c < PP; ++c) {
                double s = 0.0;
#pragma unroll
                for (int d = 0; d < PP; ++d)
                    s = fma(c >= d ? V[Tri<PP>::at(c, d)] : V[Tri<PP>::at(d, c)], (d < p) ? U[d] : 0.0, s);
                cand[c] = beta[c] + ((c < p) ? s : 0.0);
            }
            // step halving while the penalised likelihood gets worse (model.py:470-476)
            double llf_new, fl_new, hxx_new = 0.0;
            int j = 0;
            for (;;) {
                fx_eval<PP>(a, xrow, lane, cand, H, g, maxdev, llf_new, true);
#pragma unroll
                for (int c = 0; c < PP; ++c)
                    if (c == a.q) hxx_new = H[Tri<PP>::at(c, c)];
                // log det via a scratch Cholesky (V is free to be reused as scratch)
#pragma unroll
                for (int e = 0; e < Tri<PP>::SIZE; ++e) V[e] = H[e];
                double ld;
                if (fx_chol_firth<PP>(V)) ld = fx_chol_logdet<PP>(V);
                else {
                    if (lane == 0 && a.has_x) atomicAdd(&a.counters[8], 1);
                    ld = fx_singular<PP, false>(H, V, p);
                }
                fl_new = -(llf_new + 0.5 * ld);
                // model.py:470: `while firth_likelihood(new) > firth_likelihood(old)`.  Near convergence the
                // two values agree to the last few bits and the comparison is decided by rounding noise;
                // a step that halves down to one ulp above the old vector (0.5 ulp rounds back up when the
                // old mantissa is odd) can then stay "worse" for all 1000 halvings -- the reference has
                // this coin flip too (about one Firth fit in 2000 here), with its own noise.  A difference
                // within 16 ulp of the likelihood is treated as "not worse": the fit then follows the path
                // the reference takes whenever its own rounding is not the unlucky one.
                if (!(fl_new > fl_cur + 2e-15 * fabs(fl_cur))) break;
#pragma unroll
                for (int c = 0; c < PP; ++c) cand[c] = beta[c] + 0.5 * (cand[c] - beta[c]);
                if (++j > 1000) { ok = false; break; }
            }
            n_halve += j;
            if (!ok) break;
            // betas.append(new_beta)
            double nrm = 0.0;
#pragma unroll
            for (int c = 0; c < PP; ++c) {
                double d = beta[c] - prev[c];
                nrm = fma(d, d, nrm);
            }
            const double prev_step = sqrt(nrm);            // || betas[i] - betas[i-1] ||
#pragma unroll
            for (int c = 0; c < PP; ++c) {
                prev[c] = beta[c];
                beta[c] = cand[c];
            }
            llf_cur = llf_new;
            hxx_cur = hxx_new;
            fitll = -fl_new;
            hxx_fit = hxx_new;
            last_step_norm = prev_step;
            if (i > 0 && prev_step < 1e-4) { converged = true; break; }    // model.py:480-483
        }
        (void)hxx_cur;
        if (lane == 0 && a.has_x) {                        // diagnostics: longest fit of the launch
            const int old = atomicMax(&a.counters[9], n_iter);
            if (n_iter > old) a.counters[11] = v;
            const int oldh = atomicMax(&a.counters[10], n_halve);
            if (n_halve > oldh) a.counters[12] = v;
        }
        if (ok && !converged) ok = false;                  // model.py:485-486 (limit reached)
        (void)last_step_norm;
        if (!a.has_x) {
            if (lane == 0) {
                a.null_out[2 * a.q] = ok ? fitll : NAN;
                a.null_out[2 * a.q + 1] = ok ? 0.0 : (double)PSB_F_FIRTH_FAIL;
#pragma unroll
                for (int c = 0; c < PP; ++c)
                    if (c < a.q) {
                        a.null_out[c] = beta[c];
                        a.null_out[a.q + c] = NAN;
                    }
            }
            continue;
        }
        if (lane == 0) {
            if (!ok) {
                a.flags[v] = f | PSB_F_FIRTH_FAIL | PSB_F_FILTER;       // model.py:356-362
                atomicAdd(&a.counters[2], 1);
            } else {
                fx_publish<PP>(a, v, f, beta, sqrt(hxx_fit), fitll, a.null_firth);   // bse: model.py:491
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Lineage effects (model.fit_lineage_effect, model.py:151-199): Logit  k ~ [1, lineages, c]
// with the VARIANT as the response, statsmodels' default zero start; result = index of the
// lineage column with the largest Wald statistic |beta| / bse (np.argmax semantics: the first
// NaN wins), or -1 (None) when the fit fails.  mode 0: every tested variant that produced a
// fit (model.py:379-380); mode 1: tested variants that passed the lrt filter (lmm.py:208-211).
// ---------------------------------------------------------------------------------------
template <int PP>
__global__ void __launch_bounds__(128)
k_fixed_lineage(FxArgs a, const int32_t *__restrict__ idx, int n_tested, int mode, int n_lin,
                const int32_t *__restrict__ nmissing, int32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    for (int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_tested; t += warps_total) {
        const int v = idx[t];
        const uint32_t f = a.flags[v];
        const uint32_t skip = mode == 0 ? (PSB_F_PREFILTER | PSB_F_FIRTH_FAIL | PSB_F_MISSING_DATA)
                                        : (PSB_F_PREFILTER | PSB_F_FILTER);
        if ((f & skip) || nmissing[v] > 0) {
            if (lane == 0) out[v] = -1;
            continue;
        }
        const uint32_t *krow = a.bits + (size_t)v * a.Wrow;
        double beta[PP];
#pragma unroll
        for (int c = 0; c < PP; ++c) beta[c] = 0.0;
        double H[Tri<PP>::SIZE], g[PP];
        double maxdev, llf, maxstep = INFINITY;
        bool fail = false;
        int it = 0;
        const double inv_n = 1.0 / (double)a.N;
        for (;;) {
            fx_eval<PP>(a, krow, lane, beta, H, g, maxdev, llf, false, krow);
            if (it > 0 && maxdev <= 1e-8) { fail = true; break; }
            if (it > 0 && !(maxstep > 1e-8)) break;
            if (it >= 35) break;
#pragma unroll
            for (int e = 0; e < Tri<PP>::SIZE; ++e) H[e] *= inv_n;
#pragma unroll
            for (int c = 0; c < PP; ++c) {
                if (c < a.q) H[Tri<PP>::at(c, c)] -= 1e-10;
                g[c] *= inv_n;
            }
            if (!fx_ldl<PP>(H)) { fail = true; break; }
            fx_ldl_solve<PP>(H, g);
            maxstep = 0.0;
#pragma unroll
            for (int c = 0; c < PP; ++c) {
                beta[c] += g[c];
                maxstep = fmax(maxstep, fabs(g[c]));
            }
            if (isnan(maxstep)) { fail = true; break; }
            ++it;
        }
        int best = -1;
        if (!fail && fx_chol<PP>(H)) {
            double bestval = -INFINITY;
            bool seen_nan = false;
#pragma unroll
            for (int c0 = 1; c0 < PP; ++c0) {
                if (c0 <= n_lin) {
                    double e[PP];
#pragma unroll
                    for (int c = 0; c < PP; ++c) e[c] = (c == c0) ? 1.0 : 0.0;
                    fx_chol_solve<PP>(H, e);
                    const double wald = fabs(beta[c0]) / sqrt(e[c0]);
                    if (isnan(wald)) {
                        if (!seen_nan) { best = c0 - 1; seen_nan = true; }
                    } else if (!seen_nan && (best < 0 || wald > bestval)) {
                        best = c0 - 1;
                        bestval = wald;
                    }
                }
            }
        }
        if (lane == 0) out[v] = best;
    }
}

// ---------------------------------------------------------------------------------------
// OLS (statsmodels OLS.fit, call site model.py:300-312) in closed form from the masked sums
//   u = Z'x, xy = x'y (k_bitsums), G = (Z'Z)^-1, Zty, yQy precomputed:
//   xQx = x'x - u'Gu, xQy = xy - u'G Zty, beta_k = xQy / xQx, gamma = G (Zty - u beta_k),
//   RSS = yQy - beta_k xQy, df = N - (q + 1), bse_k = sqrt(RSS / df / xQx).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fixed_ols(int n_tested, const int32_t *__restrict__ idx, const double *__restrict__ sums, int C,
            int col_y, int col_z1, int q, int N, const int32_t *__restrict__ carriers,
            const double *__restrict__ consts /* G[q*q], Zty[q], yQy */, double lrt_pvalue,
            double *__restrict__ pvalue, double *__restrict__ beta_out, double *__restrict__ bse_out,
            double *__restrict__ intercept, double *__restrict__ betas, uint32_t *__restrict__ flags,
            int *__restrict__ counters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tested) return;
    const int v = idx[t];
    uint32_t f = flags[v];
    if (f & PSB_F_MISSING_DATA) return;
    const double *G = consts, *Zty = consts + q * q;
    const double yQy = consts[q * q + q];
    const double *s = sums + (size_t)v * C;
    double u[FX_GEN_MAXP];
    u[0] = (double)carriers[v];
    for (int c = 1; c < q; ++c) u[c] = s[col_z1 + c - 1];
    const double xx = u[0], xy = s[col_y];
    double Gu[FX_GEN_MAXP];
    double uGu = 0.0, uGz = 0.0;
    for (int c = 0; c < q; ++c) {
        double acc = 0.0;
        for (int d = 0; d < q; ++d) acc = fma(G[c * q + d], u[d], acc);
        Gu[c] = acc;
        uGu = fma(acc, u[c], uGu);
        uGz = fma(acc, Zty[c], uGz);
    }
    const double xQx = xx - uGu, xQy = xy - uGz;
    const double bk = xQy / xQx;
    const double rss = yQy - bk * xQy;
    const double df = (double)(N - (q + 1));
    const double bse = sqrt(rss / df / xQx);
    const double tv = bk / bse;
    double p = psb_t2_sf(tv * tv, df);
    if (!(xQx > 1e-9 * fmax(xx, 1.0))) p = NAN;        // k in the span of Z: rank-deficient design
    if (p > lrt_pvalue || !isfinite(p) || !isfinite(bk)) {
        f |= PSB_F_LRT_FAILED | PSB_F_FILTER;
        atomicAdd(&counters[2], 1);
    }
    pvalue[v] = p;
    beta_out[v] = bk;
    bse_out[v] = bse;
    // gamma = G Zty - beta_k G u
    for (int c = 0; c < q; ++c) {
        double acc = 0.0;
        for (int d = 0; d < q; ++d) acc = fma(G[c * q + d], Zty[d], acc);
        double gam = acc - bk * Gu[c];
        if (c == 0) intercept[v] = gam;
        else betas[(size_t)v * (q - 1) + (c - 1)] = gam;
    }
    flags[v] = f;
}

// missing-data-error (model.py:371-377): NaN genotypes reach the design matrix
__global__ void k_fixed_mark_missing(int n_tested, const int32_t *__restrict__ idx,
                                     const int32_t *__restrict__ nmissing, uint32_t *__restrict__ flags,
                                     int *__restrict__ counters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tested) return;
    int v = idx[t];
    if (nmissing[v] > 0) {
        flags[v] |= PSB_F_MISSING_DATA | PSB_F_FILTER;
        atomicAdd(&counters[2], 1);
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
int psb_upload_pheno(psb_ctx *c, const double *y);
void psb_fill_welch_cols(const double *y, int N, int Npad, double *cols, int col_w0, uint64_t *mask_lo,
                         uint64_t *mask_hi);

static int fx_run_null(psb_ctx *c, int mode, std::vector<double> &h);

static bool host_chol_inverse(std::vector<double> &A, int q) {
    // A (q x q, symmetric PD, row-major) -> A^-1 by Cholesky
    std::vector<double> L(A);
    for (int j = 0; j < q; ++j) {
        double d = L[j * q + j];
        for (int k = 0; k < j; ++k) d -= L[j * q + k] * L[j * q + k];
        if (!(d > 0.0)) return false;
        L[j * q + j] = sqrt(d);
        for (int i = j + 1; i < q; ++i) {
            double s = L[i * q + j];
            for (int k = 0; k < j; ++k) s -= L[i * q + k] * L[j * q + k];
            L[i * q + j] = s / L[j * q + j];
        }
    }
    for (int c = 0; c < q; ++c) {
        std::vector<double> e(q, 0.0);
        e[c] = 1.0;
        for (int i = 0; i < q; ++i) {
            double s = e[i];
            for (int k = 0; k < i; ++k) s -= L[i * q + k] * e[k];
            e[i] = s / L[i * q + i];
        }
        for (int i = q - 1; i >= 0; --i) {
            double s = e[i];
            for (int k = i + 1; k < q; ++k) s -= L[k * q + i] * e[k];
            e[i] = s / L[i * q + i];
        }
        for (int i = 0; i < q; ++i) A[i * q + c] = e[i];
    }
    return true;
}

static int fixed_common_setup(psb_ctx *c, int32_t N, int32_t q, const double *Z, const double *y) {
    PSB_REQUIRE(c && Z && y, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(N > 1 && q >= 1, PSB_ERR_ARG, "bad shape N=%d q=%d", N, q);
    PSB_REQUIRE(q + 1 <= FX_GEN_MAXP, PSB_ERR_UNSUPPORTED,
                "design width %d exceeds the device solver's limit of %d columns", q + 1, FX_GEN_MAXP);
    for (int i = 0; i < N; ++i)
        PSB_REQUIRE(Z[(size_t)i * q] == 1.0, PSB_ERR_ARG, "column 0 of Z must be the intercept (ones)");
    PSB_CUDA(cudaSetDevice(c->device));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    psb_free_model(c);
    c->N = N;
    c->Wn = (N + 31) / 32;
    c->Npad = c->Wn * 32;
    c->q = q;
    std::vector<double> Zc((size_t)q * c->Npad, 0.0);
    for (int i = 0; i < N; ++i)
        for (int k = 0; k < q; ++k) Zc[(size_t)k * c->Npad + i] = Z[(size_t)i * q + k];
    PSB_CUDA(cudaMalloc(&c->d_Z, Zc.size() * sizeof(double)));
    PSB_CUDA(cudaMemcpy(c->d_Z, Zc.data(), Zc.size() * sizeof(double), cudaMemcpyHostToDevice));
    return psb_upload_pheno(c, y);
}

extern "C" int psb_fixed_setup(psb_ctx *c, int32_t N, int32_t q, const double *Z, const double *y,
                               int32_t continuous, double null_llf, double null_firth) {
    int rc = fixed_common_setup(c, N, q, Z, y);
    if (rc) return rc;
    c->continuous = continuous ? 1 : 0;
    c->null_llf = null_llf;
    c->null_firth = null_firth;
    double mean = 0.0;
    for (int i = 0; i < N; ++i) mean += y[i];
    mean /= N;
    c->y_mean = mean;
    // masked-sum columns for k_bitsums: y | Z_1..Z_{q-1} | Welch(4)
    c->C = 1 + (q - 1) + 4;
    c->col_b = 0;
    c->col_q0 = 1;
    c->col_w0 = q;
    c->colmask_lo = c->colmask_hi = 0;
    std::vector<double> cols((size_t)c->C * c->Npad, 0.0);
    for (int i = 0; i < N; ++i) {
        cols[i] = y[i];
        for (int k = 1; k < q; ++k) cols[(size_t)k * c->Npad + i] = Z[(size_t)i * q + k];
    }
    psb_fill_welch_cols(y, N, c->Npad, cols.data(), c->col_w0, &c->colmask_lo, &c->colmask_hi);
    rc = psb_upload_welch_T(c, &cols[(size_t)c->col_w0 * c->Npad], &cols[(size_t)(c->col_w0 + 1) * c->Npad]);
    if (rc) return rc;
    PSB_CUDA(cudaMalloc(&c->d_cols, cols.size() * sizeof(double)));
    PSB_CUDA(cudaMemcpy(c->d_cols, cols.data(), cols.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (continuous && q * 2 <= 32) {
        // y'x and Z'x through one linear tensor pass instead of q masked fp64 column sums
        rc = psb_tc_linear_setup(c, cols.data(), q, c->Npad);
        if (rc) return rc;
    }
    if (continuous) {
        // G = (Z'Z)^-1, Zty, yQy
        std::vector<double> G((size_t)q * q, 0.0), Zty(q, 0.0);
        double yty = 0.0;
        for (int i = 0; i < N; ++i) {
            const double *zi = Z + (size_t)i * q;
            for (int a = 0; a < q; ++a) {
                Zty[a] += zi[a] * y[i];
                for (int b = 0; b < q; ++b) G[a * q + b] += zi[a] * zi[b];
            }
            yty += y[i] * y[i];
        }
        PSB_REQUIRE(host_chol_inverse(G, q), PSB_ERR_ARG,
                    "covariate design [1, m, c] is rank deficient; drop collinear columns");
        double zGz = 0.0;
        for (int a = 0; a < q; ++a)
            for (int b = 0; b < q; ++b) zGz += Zty[a] * G[a * q + b] * Zty[b];
        std::vector<double> consts(G);
        consts.insert(consts.end(), Zty.begin(), Zty.end());
        consts.push_back(yty - zGz);
        PSB_CUDA(cudaMalloc(&c->d_fixed_const, consts.size() * sizeof(double)));
        PSB_CUDA(cudaMemcpy(c->d_fixed_const, consts.data(), consts.size() * sizeof(double),
                            cudaMemcpyHostToDevice));
    }
    c->model = PSB_MODEL_FIXED;
    c->h_warm.clear();
    if (!continuous) {
        // null-model parameters: warm start for every variant's Newton run
        std::vector<double> h;
        rc = fx_run_null(c, 0, h);
        if (rc) return rc;
        if (h[2 * q + 1] == 0.0) c->h_warm.assign(h.begin(), h.begin() + q);
        c->logit_first_step = false;
        if ((int)c->h_warm.size() == q && q + 1 <= FX_MAXP && (q + 1) * 2 <= 32) {
            // operands of the closed-form first Newton step: Hzz^-1 and the masked-sum columns
            // [w0 z_0 .. w0 z_{q-1}, y - pi0] for the linear tensor tile
            std::vector<double> Hzz((size_t)q * q, 0.0), lc((size_t)(q + 1) * c->Npad, 0.0);
            for (int i = 0; i < N; ++i) {
                const double *zi = Z + (size_t)i * q;
                double eta = 0.0;
                for (int k = 0; k < q; ++k) eta += c->h_warm[k] * zi[k];
                const double pi = 1.0 / (1.0 + exp(-eta));
                const double w0 = pi * (1.0 - pi);
                for (int a2 = 0; a2 < q; ++a2) {
                    lc[(size_t)a2 * c->Npad + i] = w0 * zi[a2];
                    for (int b2 = 0; b2 < q; ++b2) Hzz[a2 * q + b2] += w0 * zi[a2] * zi[b2];
                }
                lc[(size_t)q * c->Npad + i] = y[i] - pi;
            }
            if (host_chol_inverse(Hzz, q)) {
                rc = psb_tc_linear_setup(c, lc.data(), q + 1, c->Npad);
                if (rc) return rc;
                PSB_CUDA(cudaMalloc(&c->d_fixed_const, Hzz.size() * sizeof(double)));
                PSB_CUDA(cudaMemcpy(c->d_fixed_const, Hzz.data(), Hzz.size() * sizeof(double),
                                    cudaMemcpyHostToDevice));
                c->logit_first_step = true;
                if (!getenv("PSB_LOGIT_FAST") || atoi(getenv("PSB_LOGIT_FAST")) != 0) {
                    rc = psb_fixed_fast_setup(c, Z, c->h_warm.data());
                    if (rc) return rc;
                }
            }
        }
    }
    PSB_UPLOAD_FENCE();
    return PSB_OK;
}

static FxArgs fx_args(psb_ctx *c, const psb_params *prm, int has_x) {
    FxArgs a;
    a.bits = c->d_bits;
    a.Z = c->d_Z;
    a.y1 = c->d_y1bits;
    a.valid = c->d_valid;
    a.Wrow = c->Wrow;
    a.Wn = c->Wn;
    a.N = c->N;
    a.Npad = c->Npad;
    a.q = c->q;
    a.has_x = has_x;
    a.start0 = log(c->y_mean / (1.0 - c->y_mean));
    static const bool warm_ok = !(getenv("PSB_LOGIT_WARM") && atoi(getenv("PSB_LOGIT_WARM")) == 0);   // debugging
    a.use_warm = (warm_ok && has_x && (int)c->h_warm.size() == c->q && c->q + 1 <= FX_MAXP) ? 1 : 0;
    for (int k = 0; k < FX_MAXP; ++k) a.warm[k] = (a.use_warm && k < c->q) ? c->h_warm[k] : 0.0;
    a.null_llf = c->null_llf;
    a.null_firth = c->null_firth;
    a.lrt_pvalue = prm ? prm->lrt_pvalue : 1.0;
    a.pvalue = c->d_pvalue;
    a.beta = c->d_beta;
    a.bse = c->d_bse;
    a.intercept = c->d_extra;
    a.betas = c->d_betas;
    a.flags = c->d_flags;
    a.counters = c->d_counters;
    a.firth_list = c->d_idx2;
    a.null_out = nullptr;
    a.sums = nullptr;
    a.HzzInv = nullptr;
    a.sums_ld = 0;
    return a;
}

template <int PP>
static void launch_logit(psb_ctx *c, const FxArgs &a, int n, int grid, const int32_t *list) {
    static const int minb = getenv("PSB_LOGIT_MINB") ? atoi(getenv("PSB_LOGIT_MINB")) : 2;
    if (PP == 12 && minb == 3) k_fixed_logit<PP, 3><<<grid, 128, 0, c->stream>>>(a, list, n);
    else if (PP == 12 && minb == 4) k_fixed_logit<PP, 4><<<grid, 128, 0, c->stream>>>(a, list, n);
    else k_fixed_logit<PP, 2><<<grid, 128, 0, c->stream>>>(a, list, n);
}
template <int PP>
static void launch_firth(psb_ctx *c, const FxArgs &a, int n, int grid) {
    k_fixed_firth<PP><<<grid, 128, 0, c->stream>>>(a, n);
}

static int fx_dispatch(psb_ctx *c, const FxArgs &a, int n, bool firth, const int32_t *list = nullptr) {
    if (n <= 0) return PSB_OK;
    if (!list) list = c->d_idx;
    const int p = a.q + (a.has_x ? 1 : 0);
    const int warps = 4;
    int grid = std::min(psb_div_up(n, warps), c->sm_count * 8);
    if (p <= 4) firth ? launch_firth<4>(c, a, n, grid) : launch_logit<4>(c, a, n, grid, list);
    else if (p <= 8) firth ? launch_firth<8>(c, a, n, grid) : launch_logit<8>(c, a, n, grid, list);
    else if (p <= 12) firth ? launch_firth<12>(c, a, n, grid) : launch_logit<12>(c, a, n, grid, list);
    else   // wider designs: generic shared-memory solver (psb_fixed_gen.cu)
        return psb_fixed_gen_launch(c, a, firth ? (a.has_x ? FXG_FIRTH : FXG_NULL_FIRTH)
                                                : (a.has_x ? FXG_LOGIT : FXG_NULL), n, 0, 0, nullptr);
    c->launches++;
    PSB_CUDA(cudaGetLastError());
    return PSB_OK;
}

// Null-model fit (no variant column) with the device solver on the context's current design.
// mode bit 0: Firth; bit 1: start from zeros.  h = params[q], bse[q], llf, status, iterations.
static int fx_run_null(psb_ctx *c, int mode, std::vector<double> &h) {
    const int q = c->q;
    int rc = psb_ensure_capacity(c, 1, q > 1 ? q - 1 : 1);
    if (rc) return rc;
    double *d_out = nullptr;
    PSB_CUDA(cudaMalloc(&d_out, (2 * q + 3) * sizeof(double)));
    PSB_UPLOAD_FENCE();          // design / phenotype uploads precede the first kernel
    const uint32_t *save_bits = c->d_bits;
    const int save_wrow = c->Wrow;
    c->d_bits = nullptr;
    c->Wrow = 0;
    FxArgs a = fx_args(c, nullptr, 0);
    c->d_bits = save_bits;
    c->Wrow = save_wrow;
    a.null_out = d_out;
    if (mode & 2) a.start0 = 0.0;      // statsmodels default start (model.py:188)
    rc = fx_dispatch(c, a, 1, (mode & 1) != 0);
    if (rc) {
        cudaFree(d_out);
        return rc;
    }
    h.assign(2 * q + 3, 0.0);
    PSB_CUDA(cudaMemcpyAsync(h.data(), d_out, (2 * q + 3) * sizeof(double), cudaMemcpyDeviceToHost,
                             c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_out);
    return PSB_OK;
}

extern "C" int psb_run_fixed(psb_ctx *c, const psb_params *prm) {
    PSB_NVTX("psb_run_fixed");
    PSB_REQUIRE(c && prm, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(c->model == PSB_MODEL_FIXED, PSB_ERR_STATE, "psb_run_fixed without psb_fixed_setup");
    PSB_REQUIRE((prm->continuous != 0) == (c->continuous != 0), PSB_ERR_ARG,
                "params.continuous differs from psb_fixed_setup");
    PSB_CUDA(cudaSetDevice(c->device));
    int rc = psb_run_begin(c);
    if (rc) return rc;
    PSB_REQUIRE(c->d_bits || c->S == 0, PSB_ERR_STATE, "psb_run_fixed without psb_submit");
    rc = psb_ensure_capacity(c, c->S, c->q > 1 ? c->q - 1 : 1);
    if (rc) return rc;
    rc = psb_table_flip(c);
    if (rc) return rc;
    PSB_CUDA(cudaEventRecord(c->ev_run0, c->stream));
    if (c->S > 0 && c->q > 1)
        PSB_CUDA(cudaMemsetAsync(c->d_betas, 0xFF, (size_t)c->S * (c->q - 1) * sizeof(double), c->stream));
    // stats: binary needs only the popcount table; continuous needs the Welch sums here and
    // y'x, Z'x for the regression -- from the linear tensor pass when it is set up
    const bool tc_linear = c->continuous && c->tmap_Lq && !c->d_miss && psb_bitstats_fits(c);
    if (!c->d_miss && psb_bitstats_fits(c) && (!c->continuous || tc_linear))
        rc = psb_launch_bitstats(c, c->continuous);
    else
        rc = psb_launch_bitsums(c);
    if (rc) return rc;
    rc = psb_launch_prefilter(c, prm, /*lmm_rule=*/0, /*defer_welch=*/0);
    if (rc) return rc;
    int h_cnt[8] = {0};
    PSB_CUDA(cudaMemcpyAsync(h_cnt, c->d_counters, sizeof(h_cnt), cudaMemcpyDeviceToHost, c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    const int n_tested = h_cnt[0];
    if (n_tested > 0 && c->d_miss) {
        k_fixed_mark_missing<<<psb_div_up(n_tested, 256), 256, 0, c->stream>>>(
            n_tested, c->d_idx, c->d_missing, c->d_flags, c->d_counters);
        c->launches++;
        PSB_CUDA(cudaGetLastError());
    }
    PSB_CUDA(cudaEventRecord(c->ev_k0, c->stream));
    if (n_tested > 0) {
        if (c->continuous) {
            if (tc_linear) {
                rc = psb_tc_run(c, n_tested, c->d_sums, c->C);
                if (rc) return rc;
            }
            k_fixed_ols<<<psb_div_up(n_tested, 256), 256, 0, c->stream>>>(
                n_tested, c->d_idx, c->d_sums, c->C, c->col_b, c->col_q0, c->q, c->N, c->d_carriers,
                c->d_fixed_const, prm->lrt_pvalue, c->d_pvalue, c->d_beta, c->d_bse, c->d_extra,
                c->d_betas, c->d_flags, c->d_counters);
            c->launches++;
            PSB_CUDA(cudaGetLastError());
        } else {
            FxArgs a = fx_args(c, prm, 1);
            if (c->logit_first_step && a.use_warm && !c->d_miss) {
                rc = psb_tc_run(c, n_tested, c->d_sums, c->C);
                if (rc) return rc;
                a.sums = c->d_sums;
                a.sums_ld = c->C;
                a.HzzInv = c->d_fixed_const;
            }
            if (c->fx_Q > 0 && a.sums) {
                // fast path over every tested variant; what it hands back (no clean convergence, a
                // fit far from the null model) goes through the reference-faithful kernel from the
                // reference's start vector
                rc = psb_fixed_fast_launch(c, a, n_tested);
                if (rc) return rc;
                PSB_CUDA(cudaMemcpyAsync(h_cnt, c->d_counters, sizeof(h_cnt), cudaMemcpyDeviceToHost,
                                         c->stream));
                PSB_CUDA(cudaStreamSynchronize(c->stream));
                c->fixed_slow = h_cnt[6];
                FxArgs as = a;
                as.use_warm = 0;
                as.sums = nullptr;
                rc = fx_dispatch(c, as, h_cnt[6], false, c->d_idx3);
            } else {
                c->fixed_slow = 0;
                rc = fx_dispatch(c, a, n_tested, false);
            }
            if (rc) return rc;
            PSB_CUDA(cudaMemcpyAsync(h_cnt, c->d_counters, sizeof(h_cnt), cudaMemcpyDeviceToHost,
                                     c->stream));
            PSB_CUDA(cudaStreamSynchronize(c->stream));
            rc = fx_dispatch(c, a, h_cnt[3], true);
            if (rc) return rc;
        }
    }
    PSB_CUDA(cudaEventRecord(c->ev_k1, c->stream));
    c->have_k_ev = true;
    PSB_CUDA(cudaEventRecord(c->ev_run1, c->stream));
    c->have_run_ev = true;
    c->ran = true;
    return psb_run_end(c);
}

template <int PP>
static void launch_lineage(psb_ctx *c, const FxArgs &a, int n, int grid, int mode) {
    k_fixed_lineage<PP><<<grid, 128, 0, c->stream>>>(a, c->d_idx, n, mode, c->n_lin, c->d_missing,
                                                     c->d_lineage);
}

// Lineage design [1, lineage columns, covariates] for model.fit_lineage_effect; call after
// the model set-up (it shares the context's sample count and phenotype-order bit layout).
extern "C" int psb_lineage_setup(psb_ctx *c, int32_t N, int32_t q, const double *Zlin,
                                 int32_t n_lineage) {
    PSB_REQUIRE(c && Zlin, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(c->model != PSB_MODEL_NONE && N == c->N, PSB_ERR_STATE,
                "psb_lineage_setup needs a model set up with the same n_samples");
    PSB_REQUIRE(q >= 2 && q <= FX_GEN_MAXP && n_lineage >= 1 && n_lineage <= q - 1, PSB_ERR_UNSUPPORTED,
                "lineage design of %d columns (%d lineages) is outside the solver's range (<= %d)",
                q, n_lineage, FX_GEN_MAXP);
    PSB_CUDA(cudaSetDevice(c->device));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    const int Npad = ((N + 31) / 32) * 32;
    std::vector<double> Zc((size_t)q * Npad, 0.0);
    for (int i = 0; i < N; ++i) {
        PSB_REQUIRE(Zlin[(size_t)i * q] == 1.0, PSB_ERR_ARG, "column 0 of the lineage design must be ones");
        for (int k = 0; k < q; ++k) Zc[(size_t)k * Npad + i] = Zlin[(size_t)i * q + k];
    }
    if (c->d_Zlin) cudaFree(c->d_Zlin);
    c->d_Zlin = nullptr;
    PSB_CUDA(cudaMalloc(&c->d_Zlin, Zc.size() * sizeof(double)));
    PSB_CUDA(cudaMemcpy(c->d_Zlin, Zc.data(), Zc.size() * sizeof(double), cudaMemcpyHostToDevice));
    c->q_lin = q;
    c->n_lin = n_lineage;
    PSB_UPLOAD_FENCE();
    return PSB_OK;
}

extern "C" int psb_run_lineage(psb_ctx *c, int32_t mode) {
    PSB_NVTX("psb_run_lineage");
    PSB_REQUIRE(c, PSB_ERR_ARG, "ctx is NULL");
    PSB_REQUIRE(c->d_Zlin && c->ran, PSB_ERR_STATE, "psb_run_lineage needs psb_lineage_setup and a finished psb_run_*");
    PSB_CUDA(cudaSetDevice(c->device));
    int h_cnt[8] = {0};
    PSB_CUDA(cudaMemcpyAsync(h_cnt, c->d_counters, sizeof(h_cnt), cudaMemcpyDeviceToHost, c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    const int n = h_cnt[0];
    if (c->S > 0) PSB_CUDA(cudaMemsetAsync(c->d_lineage, 0xFF, (size_t)c->S * sizeof(int32_t), c->stream));
    if (n <= 0) return PSB_OK;
    FxArgs a = fx_args(c, nullptr, 0);
    a.Z = c->d_Zlin;
    a.q = c->q_lin;
    a.Npad = ((c->N + 31) / 32) * 32;
    a.start0 = 0.0;
    a.use_warm = 0;
    const int grid = std::min(psb_div_up(n, 4), c->sm_count * 8);
    const int p = c->q_lin;
    if (p <= 4) launch_lineage<4>(c, a, n, grid, mode);
    else if (p <= 8) launch_lineage<8>(c, a, n, grid, mode);
    else if (p <= 12) launch_lineage<12>(c, a, n, grid, mode);
    else {
        int rc = psb_fixed_gen_launch(c, a, FXG_LINEAGE, n, mode, c->n_lin, c->d_lineage);
        return rc ? rc : psb_run_end(c);
    }
    c->launches++;
    PSB_CUDA(cudaGetLastError());
    return psb_run_end(c);
}

extern "C" int psb_fetch_lineage(psb_ctx *c, int32_t *out) {
    PSB_REQUIRE(c && out, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(c->d_lineage && c->ran, PSB_ERR_STATE, "no lineage results");
    PSB_CUDA(cudaSetDevice(c->device));
    if (c->S > 0)
        PSB_CUDA(cudaMemcpyAsync(out, c->d_lineage, (size_t)c->S * sizeof(int32_t), cudaMemcpyDefault,
                                 c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    return PSB_OK;
}

// model.fit_null (model.py:73-148) with the device solver (binary) or host normal equations
// (continuous: a single q x q system).
extern "C" int psb_fit_null(psb_ctx *c, int32_t N, int32_t q, const double *Z, const double *y,
                            int32_t continuous, int32_t firth, double *out_params, double *out_bse,
                            double *out_llf, uint32_t *out_status) {
    PSB_REQUIRE(out_params && out_bse && out_llf && out_status, PSB_ERR_ARG, "NULL output");
    *out_status = 0;
    if (continuous) {
        PSB_REQUIRE(c && Z && y && N > q && q >= 1, PSB_ERR_ARG, "bad arguments");
        std::vector<double> G((size_t)q * q, 0.0), Zty(q, 0.0);
        double yty = 0.0;
        for (int i = 0; i < N; ++i) {
            const double *zi = Z + (size_t)i * q;
            for (int a = 0; a < q; ++a) {
                Zty[a] += zi[a] * y[i];
                for (int b = 0; b < q; ++b) G[a * q + b] += zi[a] * zi[b];
            }
            yty += y[i] * y[i];
        }
        if (!host_chol_inverse(G, q)) {
            *out_status = PSB_F_MATRIX_INV;
            return PSB_OK;
        }
        double ssr = yty;
        for (int a = 0; a < q; ++a) {
            double s = 0.0;
            for (int b = 0; b < q; ++b) s += G[a * q + b] * Zty[b];
            out_params[a] = s;
        }
        for (int a = 0; a < q; ++a) ssr -= out_params[a] * Zty[a];
        const double df = N - q;
        for (int a = 0; a < q; ++a) out_bse[a] = sqrt(G[a * q + a] * ssr / df);
        *out_llf = -0.5 * N * (log(2.0 * M_PI) + log(ssr / N) + 1.0);
        return PSB_OK;
    }
    // binary: borrow the context's model slot
    int rc = fixed_common_setup(c, N, q, Z, y);
    if (rc) return rc;
    double mean = 0.0;
    for (int i = 0; i < N; ++i) mean += y[i];
    c->y_mean = mean / N;
    c->continuous = 0;
    c->model = PSB_MODEL_FIXED;
    std::vector<double> h;
    rc = fx_run_null(c, firth, h);
    if (rc) return rc;
    for (int a2 = 0; a2 < q; ++a2) {
        out_params[a2] = h[a2];
        out_bse[a2] = h[q + a2];
    }
    *out_llf = h[2 * q];
    *out_status = (uint32_t)h[2 * q + 1];
    psb_free_model(c);
    return PSB_OK;
}
