// psb_fixed_dev.cuh -- device helpers shared by the register-resident fixed-effects kernels
// (psb_fixed.cu: reference-faithful Logit Newton, Firth, lineage; psb_fixed_fast.cu: the fast Logit
// path): packed symmetric matrices in registers, Cholesky / L D L' factorisations and solves, warp
// reductions, publication of a fitted variant.
#pragma once
#include <math.h>

#include "psb_fixed.cuh"
#include "psb_math.cuh"

template <int PP>
struct Tri {
    static constexpr int SIZE = PP * (PP + 1) / 2;
    __host__ __device__ static constexpr int at(int a, int b) { return a * (a + 1) / 2 + b; }   // b <= a
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// In-place lower Cholesky of the packed symmetric matrix; returns false when a pivot is not
// positive (matrix not PD) or not finite.
// (all loops run over the full constant range with constant-foldable guards, so that the
// unroller turns every index into a literal and the arrays stay in registers)
//
// Right-looking (outer-product) form: after column j is scaled, the trailing submatrix
// update is (PP-j)^2/2 independent FMAs, so the dependency chain per column is just
// sqrt -> reciprocal -> multiply -> FMA.  The diagonal is stored as 1 / L_jj.
template <int PP>
__device__ __forceinline__ bool fx_chol(double (&A)[Tri<PP>::SIZE]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < PP; ++j) {
        const double d = A[Tri<PP>::at(j, j)];
        if (!(d > 0.0) || !isfinite(d)) ok = false;
        const double inv = rsqrt(d);
        A[Tri<PP>::at(j, j)] = inv;
#pragma unroll
        for (int i = 0; i < PP; ++i)
            if (i > j) A[Tri<PP>::at(i, j)] *= inv;
#pragma unroll
        for (int i = 0; i < PP; ++i)
#pragma unroll
            for (int k = 0; k < PP; ++k)
                if (i > j && k > j && k <= i)
                    A[Tri<PP>::at(i, k)] = fma(-A[Tri<PP>::at(i, j)], A[Tri<PP>::at(k, j)], A[Tri<PP>::at(i, k)]);
    }
    return ok;
}

// b := (L L')^-1 b   (column-oriented substitutions; diagonal of L holds reciprocals)
template <int PP>
__device__ __forceinline__ void fx_chol_solve(const double (&L)[Tri<PP>::SIZE], double (&b)[PP]) {
#pragma unroll
    for (int i = 0; i < PP; ++i) {
        b[i] *= L[Tri<PP>::at(i, i)];
#pragma unroll
        for (int k = 0; k < PP; ++k)
            if (k > i) b[k] = fma(-L[Tri<PP>::at(k, i)], b[i], b[k]);
    }
#pragma unroll
    for (int ii = 0; ii < PP; ++ii) {
        const int i = PP - 1 - ii;
        b[i] *= L[Tri<PP>::at(i, i)];
#pragma unroll
        for (int k = 0; k < PP; ++k)
            if (k < i) b[k] = fma(-L[Tri<PP>::at(i, k)], b[i], b[k]);
    }
}

// log det of the factored matrix: -2 sum log(1 / L_jj)
template <int PP>
__device__ __forceinline__ double fx_chol_logdet(const double (&L)[Tri<PP>::SIZE]) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < PP; ++c) s += log(L[Tri<PP>::at(c, c)]);
    return -2.0 * s;
}

// Cholesky that also recognises NUMERICALLY singular matrices: a pivot is the squared norm of what
// is left of column j after the earlier columns have been projected out, so a pivot below
// rel_floor times the column's own diagonal entry means "linearly dependent up to rounding"
// (an exactly duplicated column leaves +-a few ulp of its diagonal behind).  Measured against the
// column itself -- not against the largest diagonal entry -- because quasi-separated fits have
// legitimately tiny diagonal entries (weights of order e^-35 after statsmodels' 35 Newton steps)
// that the reference's LU happily inverts.  Used where the reference's own behaviour on a singular
// matrix is reproduced: Firth regression (pinv / det, psb_sym_pinv_logdet) and the null fit (Powell).
template <int PP>
__device__ __forceinline__ bool fx_chol_firth(double (&A)[Tri<PP>::SIZE], double rel_floor = 1e-13) {
    double orig[PP];
#pragma unroll
    for (int j = 0; j < PP; ++j) orig[j] = A[Tri<PP>::at(j, j)];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < PP; ++j) {
        const double d = A[Tri<PP>::at(j, j)];
        if (!(d > rel_floor * orig[j]) || !isfinite(d)) ok = false;
        const double inv = rsqrt(d);
        A[Tri<PP>::at(j, j)] = inv;
#pragma unroll
        for (int i = 0; i < PP; ++i)
            if (i > j) A[Tri<PP>::at(i, j)] *= inv;
#pragma unroll
        for (int i = 0; i < PP; ++i)
#pragma unroll
            for (int k = 0; k < PP; ++k)
                if (i > j && k > j && k <= i)
                    A[Tri<PP>::at(i, k)] = fma(-A[Tri<PP>::at(i, j)], A[Tri<PP>::at(k, j)], A[Tri<PP>::at(i, k)]);
    }
    return ok;
}

// Singular H: pinv (when WANT_V) and log det by eigendecomposition, through local-memory copies so
// that the caller's arrays stay in registers.
template <int PP, bool WANT_V>
__device__ __forceinline__ double fx_singular(const double (&H)[Tri<PP>::SIZE], double (&V)[Tri<PP>::SIZE],
                                              int p_active) {
    double Hl[Tri<PP>::SIZE], Vl[Tri<PP>::SIZE], A[PP * PP], Q[PP * PP];
#pragma unroll
    for (int e = 0; e < Tri<PP>::SIZE; ++e) Hl[e] = H[e];
    const double ld = psb_sym_pinv_logdet(Hl, WANT_V ? Vl : nullptr, PP, p_active, A, Q);
    if (WANT_V) {
#pragma unroll
        for (int e = 0; e < Tri<PP>::SIZE; ++e) V[e] = Vl[e];
    }
    return ld;
}

// statsmodels' Newton step matrix X'WX/n - 1e-10 I (the ridge lands on the NEGATIVE definite
// hessian, base/model.py:fit + base/optimizer.py:_fit_newton) can turn indefinite in separated
// data, where the reference's LU solve simply carries on.  L D L' without pivoting solves the
// same system for any matrix with non-zero leading minors; a zero / non-finite pivot is the
// analogue of numpy's "Singular matrix".  Unit lower L below the diagonal, 1 / d_j on it.
template <int PP>
__device__ __forceinline__ bool fx_ldl(double (&A)[Tri<PP>::SIZE]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < PP; ++j) {
        const double d = A[Tri<PP>::at(j, j)];
        if (d == 0.0 || !isfinite(d)) ok = false;
        const double inv = 1.0 / d;
        A[Tri<PP>::at(j, j)] = inv;
#pragma unroll
        for (int i = 0; i < PP; ++i)
#pragma unroll
            for (int k = 0; k < PP; ++k)
                if (i > j && k > j && k <= i)
                    A[Tri<PP>::at(i, k)] = fma(-A[Tri<PP>::at(i, j)] * inv, A[Tri<PP>::at(k, j)], A[Tri<PP>::at(i, k)]);
#pragma unroll
        for (int i = 0; i < PP; ++i)
            if (i > j) A[Tri<PP>::at(i, j)] *= inv;
    }
    return ok;
}

// b := (L D L')^-1 b
template <int PP>
__device__ __forceinline__ void fx_ldl_solve(const double (&L)[Tri<PP>::SIZE], double (&b)[PP]) {
#pragma unroll
    for (int i = 0; i < PP; ++i) {
#pragma unroll
        for (int k = 0; k < PP; ++k)
            if (k > i) b[k] = fma(-L[Tri<PP>::at(k, i)], b[i], b[k]);
    }
#pragma unroll
    for (int i = 0; i < PP; ++i) b[i] *= L[Tri<PP>::at(i, i)];
#pragma unroll
    for (int ii = 0; ii < PP; ++ii) {
        const int i = PP - 1 - ii;
#pragma unroll
        for (int k = 0; k < PP; ++k)
            if (k < i) b[k] = fma(-L[Tri<PP>::at(i, k)], b[i], b[k]);
    }
}

// Straightforward, register-friendly inverse: solve for each unit vector (PP solves).  Used by
// the Firth path only; V is returned packed (lower triangle).
template <int PP>
__device__ __forceinline__ void fx_inverse_from_chol(const double (&L)[Tri<PP>::SIZE],
                                                     double (&V)[Tri<PP>::SIZE]) {
#pragma unroll
    for (int c = 0; c < PP; ++c) {
        double e[PP];
#pragma unroll
        for (int i = 0; i < PP; ++i) e[i] = (i == c) ? 1.0 : 0.0;
        fx_chol_solve<PP>(L, e);
#pragma unroll
        for (int i = 0; i < PP; ++i)
            if (i >= c) V[Tri<PP>::at(i, c)] = e[i];
    }
}

// Publishes a fitted variant: LRT against the matching null, lrt filter, result columns.
template <int PP, class BetaT>
__device__ __forceinline__ void fx_publish(const FxArgs &a, int v, uint32_t f, const BetaT &beta,
                                           double bse, double fit_llf, double null_llf) {
    const double lrstat = -2.0 * (null_llf - fit_llf);          // model.py:336, :366
    double p = 1.0;
    if (lrstat > 0.0) p = psb_chi2_sf1(lrstat);      // NaN compares false: p stays 1, as in the reference
    double kbeta = 0.0;
#pragma unroll
    for (int c = 0; c < PP; ++c)
        if (c == a.q) kbeta = beta[c];
    if (p > a.lrt_pvalue || !isfinite(p) || !isfinite(kbeta)) {   // model.py:384
        f |= PSB_F_LRT_FAILED | PSB_F_FILTER;
        atomicAdd(&a.counters[2], 1);
    }
    a.pvalue[v] = p;
    a.beta[v] = kbeta;
    a.bse[v] = bse;
    a.intercept[v] = beta[0];
#pragma unroll
    for (int c = 1; c < PP; ++c)
        if (c < a.q) a.betas[(size_t)v * (a.q - 1) + (c - 1)] = beta[c];
    a.flags[v] = f;
}

