// psb_fixed_gen.cu -- generic-width logistic solver (designs of 13..64 columns).
//
// Same algorithms and stopping rules as the register-resident templates of psb_fixed.cu
// (statsmodels Logit Newton, model.py:328-330; model.fit_firth, model.py:414-504;
// model.fit_lineage_effect, model.py:151-199; model.fit_null, model.py:73-148) for designs
// that do not fit a thread's registers -- many covariates, or lineage clusters (one column
// per cluster).  One warp per fit; the information matrix, its Cholesky factor, the
// parameter vectors and a 32-sample tile of the design live in shared memory:
//   * each lane evaluates eta, pi, w, y - pi for its own sample of a 32-sample chunk and
//     writes the sample's design row into the tile (sample-major, odd pitch);
//   * the rank-32 update of X'WX is then spread over the lanes by matrix entry;
//   * Cholesky (right-looking) and the triangular solves are warp-parallel per column.
// Always starts from the reference's start vector (no warm start), so iterates match the
// reference one to one.
#include <math.h>

#include <algorithm>

#include "psb_fixed.cuh"
#include "psb_math.cuh"

#define FXG_WARPS 4

struct GenWs {
    double *H;        // packed lower triangle, P (P + 1) / 2
    double *V;        // second triangle (Firth: inverse / scratch factor)
    double *zt;       // 32 x pitch design tile (row = sample of the chunk)
    double *wz;       // 32 x pitch, w_s * z_s (or scratch)
    double *beta, *g, *cand, *prev, *U, *tmp;   // P each
    int P, pitch, tri;
};

__device__ __forceinline__ int tri_at(int a, int b) { return a * (a + 1) / 2 + b; }   // b <= a

__device__ __forceinline__ double gw_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double gw_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One pass over the samples at parameters `par`: H = X'WX, g = X'(y - pi), max |y - pi|, llf.
// p = active columns (q, or q + 1 with the variant); yrow = bit row of the response.
__device__ void gen_eval(const FxArgs &a, const GenWs &ws, const uint32_t *xrow, const uint32_t *yrow,
                         int lane, const double *par, int p, double &maxdev, double &llf) {
    const int P = ws.P, pitch = ws.pitch;
    for (int e = lane; e < ws.tri; e += 32) ws.H[e] = 0.0;
    for (int c = lane; c < P; c += 32) ws.g[c] = 0.0;
    __syncwarp();
    maxdev = 0.0;
    llf = 0.0;
    for (int w = 0; w < a.Wn; ++w) {
        const uint32_t vw = __ldg(a.valid + w);
        const bool on = (vw >> lane) & 1u;
        const uint32_t xw = (a.has_x && xrow) ? __ldg(xrow + w) : 0u;
        const uint32_t yw = __ldg(yrow + w);
        const int i = w * 32 + lane;
        double *zr = ws.zt + lane * pitch;
        double eta = 0.0;
        for (int c = 0; c < P; ++c) {
            double z = 0.0;
            if (on) {
                if (c == 0) z = 1.0;
                else if (c < a.q) z = __ldg(a.Z + (size_t)c * a.Npad + i);
                else if (c == a.q && a.has_x) z = (double)((xw >> lane) & 1u);
            }
            zr[c] = z;
            eta = fma(par[c], z, eta);
        }
        double wgt = 0.0, r = 0.0;
        if (on) {
            const double y = (double)((yw >> lane) & 1u);
            // statsmodels' formulas, saturating exactly as the reference does (psb_fixed.cu)
            const double ex = exp(-eta);
            const double pi = 1.0 / (1.0 + ex);
            wgt = pi * (1.0 - pi);
            r = y - pi;
            maxdev = fmax(maxdev, fabs(r));
            const double s = y > 0.5 ? eta : -eta;
            llf += fmin(s, 0.0) - log1p(eta >= 0.0 ? ex : 1.0 / ex);
        }
        double *wr = ws.wz + lane * pitch;
        for (int c = 0; c < P; ++c) wr[c] = wgt * zr[c];
        ws.tmp[lane] = r;           // tmp holds >= 32 doubles (see carve-up)
        __syncwarp();
        // rank-32 update, lanes over matrix entries
        for (int e = lane; e < ws.tri; e += 32) {
            // (row, col) of packed index e
            int row = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while (tri_at(row + 1, 0) <= e) ++row;
            while (tri_at(row, 0) > e) --row;
            const int col = e - tri_at(row, 0);
            double acc = ws.H[e];
            for (int s = 0; s < 32; ++s) acc = fma(ws.wz[s * pitch + row], ws.zt[s * pitch + col], acc);
            ws.H[e] = acc;
        }
        for (int c = lane; c < P; c += 32) {
            double acc = ws.g[c];
            for (int s = 0; s < 32; ++s) acc = fma(ws.tmp[s], ws.zt[s * pitch + c], acc);
            ws.g[c] = acc;
        }
        __syncwarp();
    }
    maxdev = gw_max(maxdev);
    llf = gw_sum(llf);
    for (int c = p + lane; c < P; c += 32) ws.H[tri_at(c, c)] = 1.0;    // padding columns
    __syncwarp();
}

// In-place right-looking Cholesky of the packed matrix (diagonal stored as 1 / L_jj); returns
// false when a pivot is not positive / finite.
__device__ bool gen_chol(double *A, int P, int lane, const double *orig_diag = nullptr, double rel_floor = 0.0) {
    bool ok = true;
    for (int j = 0; j < P; ++j) {
        const double d = A[tri_at(j, j)];
        if (!(d > (orig_diag ? rel_floor * orig_diag[j] : 0.0)) || !isfinite(d)) ok = false;
        const double inv = rsqrt(d);
        __syncwarp();
        if (lane == 0) A[tri_at(j, j)] = inv;
        for (int i = j + 1 + lane; i < P; i += 32) A[tri_at(i, j)] *= inv;
        __syncwarp();
        // trailing update: entries (i, k), j < k <= i < P
        const int m = P - 1 - j;
        const int cnt = m * (m + 1) / 2;
        for (int e = lane; e < cnt; e += 32) {
            int r = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while (tri_at(r + 1, 0) <= e) ++r;
            while (tri_at(r, 0) > e) --r;
            const int i = j + 1 + r, k = j + 1 + (e - tri_at(r, 0));
            A[tri_at(i, k)] = fma(-A[tri_at(i, j)], A[tri_at(k, j)], A[tri_at(i, k)]);
        }
        __syncwarp();
    }
    return ok;
}

// L D L' without pivoting (unit L below the diagonal, 1 / d_j on it) for the Newton step
// matrix X'WX/n - 1e-10 I, which may be indefinite in separated data (see psb_fixed.cu fx_ldl).
__device__ bool gen_ldl(double *A, int P, int lane) {
    bool ok = true;
    for (int j = 0; j < P; ++j) {
        const double d = A[tri_at(j, j)];
        if (d == 0.0 || !isfinite(d)) ok = false;
        const double inv = 1.0 / d;
        __syncwarp();
        if (lane == 0) A[tri_at(j, j)] = inv;
        const int m = P - 1 - j;
        const int cnt = m * (m + 1) / 2;
        for (int e = lane; e < cnt; e += 32) {
            int r = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while (tri_at(r + 1, 0) <= e) ++r;
            while (tri_at(r, 0) > e) --r;
            const int i = j + 1 + r, k = j + 1 + (e - tri_at(r, 0));
            A[tri_at(i, k)] = fma(-A[tri_at(i, j)] * inv, A[tri_at(k, j)], A[tri_at(i, k)]);
        }
        __syncwarp();
        for (int i = j + 1 + lane; i < P; i += 32) A[tri_at(i, j)] *= inv;
        __syncwarp();
    }
    return ok;
}

__device__ void gen_ldl_solve(const double *L, double *b, int P, int lane) {
    for (int i = 0; i < P; ++i) {
        __syncwarp();
        const double bi = b[i];
        for (int k = i + 1 + lane; k < P; k += 32) b[k] = fma(-L[tri_at(k, i)], bi, b[k]);
    }
    __syncwarp();
    for (int c = lane; c < P; c += 32) b[c] *= L[tri_at(c, c)];
    for (int i = P - 1; i >= 0; --i) {
        __syncwarp();
        const double bi = b[i];
        for (int k = lane; k < i; k += 32) b[k] = fma(-L[tri_at(i, k)], bi, b[k]);
    }
    __syncwarp();
}

__device__ double gen_logdet(const double *L, int P) {
    double s = 0.0;
    for (int c = 0; c < P; ++c) s += log(L[tri_at(c, c)]);
    return -2.0 * s;
}

// b := (L L')^-1 b, warp-parallel column substitutions
__device__ void gen_solve(const double *L, double *b, int P, int lane) {
    for (int i = 0; i < P; ++i) {
        __syncwarp();
        const double bi = b[i] * L[tri_at(i, i)];
        __syncwarp();
        if (lane == 0) b[i] = bi;
        for (int k = i + 1 + lane; k < P; k += 32) b[k] = fma(-L[tri_at(k, i)], bi, b[k]);
    }
    for (int i = P - 1; i >= 0; --i) {
        __syncwarp();
        const double bi = b[i] * L[tri_at(i, i)];
        __syncwarp();
        if (lane == 0) b[i] = bi;
        for (int k = lane; k < i; k += 32) b[k] = fma(-L[tri_at(i, k)], bi, b[k]);
    }
    __syncwarp();
}

__device__ void gen_publish(const FxArgs &a, int v, uint32_t f, const double *beta, double bse,
                            double fit_llf, double null_llf) {
    const double lrstat = -2.0 * (null_llf - fit_llf);
    double p = 1.0;
    if (lrstat > 0.0) p = psb_chi2_sf1(lrstat);
    // (a NaN statistic compares false and leaves p = 1, as in the reference)
    const double kbeta = beta[a.q];
    if (p > a.lrt_pvalue || !isfinite(p) || !isfinite(kbeta)) {
        f |= PSB_F_LRT_FAILED | PSB_F_FILTER;
        atomicAdd(&a.counters[2], 1);
    }
    a.pvalue[v] = p;
    a.beta[v] = kbeta;
    a.bse[v] = bse;
    a.intercept[v] = beta[0];
    for (int c = 1; c < a.q; ++c) a.betas[(size_t)v * (a.q - 1) + (c - 1)] = beta[c];
    a.flags[v] = f;
}

// Numerically singular matrices (fx_chol_firth, psb_fixed_dev.cuh): a pivot below rel_floor of its
// column's own diagonal entry.  `diag`: P doubles of scratch for the original diagonal.
__device__ bool gen_chol_firth(double *A, int P, int lane, double *diag, double rel_floor = 1e-13) {
    __syncwarp();
    for (int j = lane; j < P; j += 32) diag[j] = A[tri_at(j, j)];
    __syncwarp();
    return gen_chol(A, P, lane, diag, rel_floor);
}

// Singular information matrix (psb_sym_pinv_logdet, psb_fixed.cuh): lane 0 works in the design-tile
// area of the warp's shared memory (A) and the warp's global scratch (Q).
__device__ double gen_singular(const GenWs &ws, double *Vout, int p, double *scratchQ, int lane) {
    double ld = 0.0;
    __syncwarp();
    if (lane == 0) ld = psb_sym_pinv_logdet(ws.H, Vout, ws.P, p, ws.zt, scratchQ);
    __syncwarp();
    return __shfl_sync(0xffffffffu, ld, 0);
}

__global__ void __launch_bounds__(FXG_WARPS * 32)
k_fixed_generic(FxArgs a, const int32_t *__restrict__ idx, int n_items, int mode, int P, int lineage_mode,
                int n_lin, const int32_t *__restrict__ nmissing, int32_t *__restrict__ lineage_out,
                double *__restrict__ scratch) {
    extern __shared__ __align__(16) double gsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    GenWs ws;
    ws.P = P;
    ws.pitch = P | 1;
    ws.tri = P * (P + 1) / 2;
    const int vecs = 5 * P + 32 + P;                // beta g cand prev U + tmp(>= 32, >= P)
    const int per_warp = 2 * ws.tri + 2 * 32 * ws.pitch + vecs;
    double *base = gsm + (size_t)warp * per_warp;
    ws.H = base;
    ws.V = ws.H + ws.tri;
    ws.zt = ws.V + ws.tri;
    ws.wz = ws.zt + 32 * ws.pitch;
    ws.beta = ws.wz + 32 * ws.pitch;
    ws.g = ws.beta + P;
    ws.cand = ws.g + P;
    ws.prev = ws.cand + P;
    ws.U = ws.prev + P;
    ws.tmp = ws.U + P;
    const int p = a.q + (a.has_x ? 1 : 0);
    const int warps_total = gridDim.x * FXG_WARPS;
    const double inv_n = 1.0 / (double)a.N;
    double *scratchQ = scratch ? scratch + (size_t)(blockIdx.x * FXG_WARPS + warp) * P * P : nullptr;

    for (int t = blockIdx.x * FXG_WARPS + warp; t < n_items; t += warps_total) {
        int v = 0;
        if (mode == FXG_LOGIT || mode == FXG_LINEAGE) v = idx[t];
        else if (mode == FXG_FIRTH) v = a.firth_list[t];
        uint32_t f = (mode == FXG_NULL || mode == FXG_NULL_FIRTH) ? 0u : a.flags[v];
        const uint32_t *xrow = a.bits ? a.bits + (size_t)v * a.Wrow : nullptr;
        const uint32_t *yrow = mode == FXG_LINEAGE ? xrow : a.y1;
        if (mode == FXG_LOGIT) {
            if (f & PSB_F_MISSING_DATA) continue;
            if (f & PSB_F_BAD_CHISQ) {
                if (lane == 0) a.firth_list[atomicAdd(&a.counters[3], 1)] = v;
                continue;
            }
        }
        if (mode == FXG_LINEAGE) {
            const uint32_t skip = lineage_mode == 0 ? (PSB_F_PREFILTER | PSB_F_FIRTH_FAIL | PSB_F_MISSING_DATA)
                                                    : (PSB_F_PREFILTER | PSB_F_FILTER);
            if ((f & skip) || nmissing[v] > 0) {
                if (lane == 0) lineage_out[v] = -1;
                continue;
            }
        }
        __syncwarp();
        for (int c = lane; c < P; c += 32) {
            ws.beta[c] = (c == 0) ? a.start0 : 0.0;
            ws.prev[c] = 0.0;
        }
        __syncwarp();
        double maxdev, llf = NAN;

        if (mode == FXG_FIRTH || mode == FXG_NULL_FIRTH) {
            // ---------------- Firth (model.py:414-504) ----------------
            f |= PSB_F_FIRTH_USED;
            bool ok = true, converged = false;
            double llf_cur, fl_cur, fitll = NAN, hxx_fit = NAN;
            gen_eval(a, ws, xrow, yrow, lane, ws.beta, p, maxdev, llf_cur);
            double hxx_new = 0.0;
            for (int i = 0; i < 1000 && ok; ++i) {
                if (gen_chol_firth(ws.H, P, lane, ws.g)) {
                    fl_cur = -(llf_cur + 0.5 * gen_logdet(ws.H, P));
                    // V = (L L')^-1, column by column (only the lower triangle is kept)
                    for (int c0 = 0; c0 < P; ++c0) {
                        for (int c = lane; c < P; c += 32) ws.tmp[c] = (c == c0) ? 1.0 : 0.0;
                        gen_solve(ws.H, ws.tmp, P, lane);
                        for (int c = c0 + lane; c < P; c += 32) ws.V[tri_at(c, c0)] = ws.tmp[c];
                        __syncwarp();
                    }
                } else {
                    // singular information matrix: pinv / det as the reference (model.py:450, :410);
                    // the factorisation ran in place, so X'WX is evaluated again first
                    gen_eval(a, ws, xrow, yrow, lane, ws.beta, p, maxdev, llf_cur);
                    const double ld0 = gen_singular(ws, ws.V, p, scratchQ, lane);
                    if (isnan(ws.V[0])) { ok = false; break; }
                    fl_cur = -(llf_cur + 0.5 * ld0);
                }
                // U = X'(y - pi + h (1/2 - pi)),  h_i = w_i x_i' V x_i
                for (int c = lane; c < P; c += 32) ws.U[c] = 0.0;
                __syncwarp();
                for (int w = 0; w < a.Wn; ++w) {
                    const uint32_t vw = __ldg(a.valid + w);
                    const bool on = (vw >> lane) & 1u;
                    const uint32_t xw = (a.has_x && xrow) ? __ldg(xrow + w) : 0u;
                    const uint32_t yw = __ldg(yrow + w);
                    const int smp = w * 32 + lane;
                    double *zr = ws.zt + lane * ws.pitch;
                    double eta = 0.0;
                    for (int c = 0; c < P; ++c) {
                        double z = 0.0;
                        if (on) {
                            if (c == 0) z = 1.0;
                            else if (c < a.q) z = __ldg(a.Z + (size_t)c * a.Npad + smp);
                            else if (c == a.q && a.has_x) z = (double)((xw >> lane) & 1u);
                        }
                        zr[c] = z;
                        eta = fma(ws.beta[c], z, eta);
                    }
                    double r = 0.0;
                    if (on) {
                        const double pi = 1.0 / (1.0 + exp(-eta));
                        const double wgt = pi * (1.0 - pi);
                        double quad = 0.0;
                        for (int c = 0; c < P; ++c) {
                            double sacc = 0.0;
                            for (int d = 0; d < P; ++d)
                                sacc = fma(c >= d ? ws.V[tri_at(c, d)] : ws.V[tri_at(d, c)], zr[d], sacc);
                            quad = fma(sacc, zr[c], quad);
                        }
                        const double y = (double)((yw >> lane) & 1u);
                        r = y - pi + wgt * quad * (0.5 - pi);
                    }
                    ws.tmp[lane] = r;
                    __syncwarp();
                    for (int c = lane; c < P; c += 32) {
                        double acc = ws.U[c];
                        for (int s = 0; s < 32; ++s) acc = fma(ws.tmp[s], ws.zt[s * ws.pitch + c], acc);
                        ws.U[c] = acc;
                    }
                    __syncwarp();
                }
                // cand = beta + V U
                for (int c = lane; c < P; c += 32) {
                    double sacc = 0.0;
                    for (int d = 0; d < p; ++d)
                        sacc = fma(c >= d ? ws.V[tri_at(c, d)] : ws.V[tri_at(d, c)], ws.U[d], sacc);
                    ws.cand[c] = ws.beta[c] + (c < p ? sacc : 0.0);
                }
                __syncwarp();
                double llf_new, fl_new;
                int j = 0;
                for (;;) {
                    gen_eval(a, ws, xrow, yrow, lane, ws.cand, p, maxdev, llf_new);
                    hxx_new = a.has_x ? ws.H[tri_at(a.q, a.q)] : 0.0;
                    for (int e = lane; e < ws.tri; e += 32) ws.V[e] = ws.H[e];
                    __syncwarp();
                    double ld;
                    if (gen_chol_firth(ws.V, P, lane, ws.g)) ld = gen_logdet(ws.V, P);
                    else ld = gen_singular(ws, nullptr, p, scratchQ, lane);
                    fl_new = -(llf_new + 0.5 * ld);
                    if (!(fl_new > fl_cur + 2e-15 * fabs(fl_cur))) break;   // see k_fixed_firth: noise-level differences are "not worse"
                    __syncwarp();
                    for (int c = lane; c < P; c += 32) ws.cand[c] = ws.beta[c] + 0.5 * (ws.cand[c] - ws.beta[c]);
                    __syncwarp();
                    if (++j > 1000) { ok = false; break; }
                }
                if (!ok) break;
                double nrm = 0.0;
                for (int c = 0; c < P; ++c) {
                    const double d = ws.beta[c] - ws.prev[c];
                    nrm = fma(d, d, nrm);
                }
                const double prev_step = sqrt(nrm);
                __syncwarp();
                for (int c = lane; c < P; c += 32) {
                    ws.prev[c] = ws.beta[c];
                    ws.beta[c] = ws.cand[c];
                }
                __syncwarp();
                llf_cur = llf_new;
                fitll = -fl_new;
                hxx_fit = hxx_new;
                if (i > 0 && prev_step < 1e-4) { converged = true; break; }
            }
            if (ok && !converged) ok = false;
            if (mode == FXG_NULL_FIRTH) {
                if (lane == 0) {
                    a.null_out[2 * a.q] = ok ? fitll : NAN;
                    a.null_out[2 * a.q + 1] = ok ? 0.0 : (double)PSB_F_FIRTH_FAIL;
                    for (int c = 0; c < a.q; ++c) {
                        a.null_out[c] = ws.beta[c];
                        a.null_out[a.q + c] = NAN;
                    }
                }
                continue;
            }
            if (lane == 0) {
                if (!ok) {
                    a.flags[v] = f | PSB_F_FIRTH_FAIL | PSB_F_FILTER;
                    atomicAdd(&a.counters[2], 1);
                } else {
                    gen_publish(a, v, f, ws.beta, sqrt(hxx_fit), fitll, a.null_firth);
                }
            }
            continue;
        }

        // ---------------- Logit Newton (statsmodels, 35 steps, |dbeta| <= 1e-8) ----------------
        uint32_t fail = 0;
        int it = 0;
        double maxstep = INFINITY;
        for (;;) {
            gen_eval(a, ws, xrow, yrow, lane, ws.beta, p, maxdev, llf);
            if (it > 0 && maxdev <= 1e-8) { fail = PSB_F_PERFECT_SEP; break; }
            if (it > 0 && !(maxstep > 1e-8)) break;
            if (it >= 35) break;
            for (int e = lane; e < ws.tri; e += 32) ws.V[e] = ws.H[e] * inv_n;
            for (int c = lane; c < P; c += 32) ws.tmp[c] = ws.g[c] * inv_n;
            __syncwarp();
            for (int c = lane; c < p; c += 32) ws.V[tri_at(c, c)] -= 1e-10;   // statsmodels' ridge sign
            __syncwarp();
            if (!gen_ldl(ws.V, P, lane)) { fail = PSB_F_MATRIX_INV; break; }
            gen_ldl_solve(ws.V, ws.tmp, P, lane);
            maxstep = 0.0;
            for (int c = 0; c < P; ++c) maxstep = fmax(maxstep, fabs(ws.tmp[c]));
            __syncwarp();
            for (int c = lane; c < P; c += 32) ws.beta[c] += ws.tmp[c];
            __syncwarp();
            if (isnan(maxstep)) { fail = PSB_F_MATRIX_INV; break; }
            ++it;
        }
        // factor X'WX at the final parameters for the standard errors
        bool have_factor = false;
        if (!fail) {
            // (pivot floor for the null fit only: see k_fixed_logit)
            have_factor = mode == FXG_NULL ? gen_chol_firth(ws.H, P, lane, ws.g, 1e-13) : gen_chol(ws.H, P, lane);
            if (!have_factor) fail = PSB_F_MATRIX_INV;
        }
        if (mode == FXG_LINEAGE) {
            int best = -1;
            if (!fail) {
                double bestval = -INFINITY;
                bool seen_nan = false;
                for (int c0 = 1; c0 <= n_lin; ++c0) {
                    for (int c = lane; c < P; c += 32) ws.tmp[c] = (c == c0) ? 1.0 : 0.0;
                    gen_solve(ws.H, ws.tmp, P, lane);
                    const double wald = fabs(ws.beta[c0]) / sqrt(ws.tmp[c0]);
                    __syncwarp();
                    if (isnan(wald)) {
                        if (!seen_nan) { best = c0 - 1; seen_nan = true; }
                    } else if (!seen_nan && (best < 0 || wald > bestval)) {
                        best = c0 - 1;
                        bestval = wald;
                    }
                }
            }
            if (lane == 0) lineage_out[v] = best;
            continue;
        }
        if (mode == FXG_NULL) {
            for (int c0 = 0; c0 < a.q; ++c0) {
                double bse = NAN;
                if (!fail) {
                    for (int c = lane; c < P; c += 32) ws.tmp[c] = (c == c0) ? 1.0 : 0.0;
                    gen_solve(ws.H, ws.tmp, P, lane);
                    bse = sqrt(ws.tmp[c0]);
                    __syncwarp();
                }
                if (lane == 0) {
                    a.null_out[c0] = ws.beta[c0];
                    a.null_out[a.q + c0] = bse;
                }
            }
            if (lane == 0) {
                a.null_out[2 * a.q] = fail ? NAN : llf;
                a.null_out[2 * a.q + 1] = (double)fail;
                a.null_out[2 * a.q + 2] = (double)it;
            }
            continue;
        }
        double bse_x = NAN;
        if (!fail) {
            for (int c = lane; c < P; c += 32) ws.tmp[c] = (c == a.q) ? 1.0 : 0.0;
            gen_solve(ws.H, ws.tmp, P, lane);
            bse_x = sqrt(ws.tmp[a.q]);
            __syncwarp();
            if (bse_x > 3.0) fail = PSB_F_HIGH_BSE;
        }
        if (fail) {
            if (lane == 0) {
                a.flags[v] = f | fail;
                a.firth_list[atomicAdd(&a.counters[3], 1)] = v;
            }
            continue;
        }
        if (lane == 0) gen_publish(a, v, f, ws.beta, bse_x, llf, a.null_llf);
    }
}

int psb_fixed_gen_launch(psb_ctx *c, const FxArgs &a, int mode, int n, int lineage_mode, int n_lin,
                         int32_t *lineage_out) {
    if (n <= 0) return PSB_OK;
    const int p = a.q + (a.has_x ? 1 : 0);
    PSB_REQUIRE(p <= FX_GEN_MAXP, PSB_ERR_UNSUPPORTED,
                "design width %d exceeds the device solver's limit of %d columns", p, FX_GEN_MAXP);
    const int P = p;
    const int pitch = P | 1, tri = P * (P + 1) / 2;
    const size_t per_warp = (size_t)2 * tri + (size_t)2 * 32 * pitch + 6 * P + 32;
    const size_t smem = per_warp * FXG_WARPS * sizeof(double);
    PSB_CUDA(cudaFuncSetAttribute(k_fixed_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::min(psb_div_up(n, FXG_WARPS), c->sm_count * 4);
    double *scratch = nullptr;
    if (mode == FXG_FIRTH || mode == FXG_NULL_FIRTH) {
        // per-warp P x P scratch of the singular-matrix path (gen_singular)
        const size_t need = (size_t)grid * FXG_WARPS * P * P * sizeof(double);
        if (need > c->gen_scratch_cap) {
            PSB_CUDA(cudaStreamSynchronize(c->stream));
            if (c->d_gen_scratch) cudaFree(c->d_gen_scratch);
            c->d_gen_scratch = nullptr;
            c->gen_scratch_cap = 0;
            PSB_CUDA(cudaMalloc(&c->d_gen_scratch, need));
            c->gen_scratch_cap = need;
        }
        scratch = c->d_gen_scratch;
    }
    k_fixed_generic<<<grid, FXG_WARPS * 32, smem, c->stream>>>(a, c->d_idx, n, mode, P, lineage_mode,
                                                              n_lin, c->d_missing, lineage_out, scratch);
    c->launches++;
    PSB_CUDA(cudaGetLastError());
    return PSB_OK;
}
