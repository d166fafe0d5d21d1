// psb_math.cuh -- fp64 special functions shared by the kernels and the host test hooks.
//
// Replaces the scipy calls on the per-variant path:
//   stats.chi2.sf(x, 1)            model.py:68 (via chi2_contingency), :339, :369
//   stats.f.sf(x, 1, dfd)          lmm.py:251-253
//   2 * stats.t.sf(|t|, df)        model.py:53-55 (ttest_ind), OLS pvalues model.py:312
// All three reduce to erfc or to the regularised incomplete beta I_x(a, 1/2).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define PSB_HD __host__ __device__ __forceinline__
#else
#define PSB_HD inline
#endif

PSB_HD double psb_t2_sf(double x, double dfd);

// Welch t-test p-value, scipy.stats.ttest_ind(p[k==1], p[k==0], equal_var=False) (model.py:53-55),
// from the sums over carriers (s1 = sum y, s1q = sum y^2, n1 samples) and non-carriers (s0, s0q, n0)
PSB_HD double psb_welch_prep(double s1, double s1q, double s0, double s0q, double n1, double n0) {
    const double nan = NAN;
    if (n1 < 1.0 || n0 < 1.0) return nan;
    double m1 = s1 / n1, m0 = s0 / n0;
    double v1 = (s1q - s1 * m1) / (n1 - 1.0);   // nan when n1 == 1
    double v0 = (s0q - s0 * m0) / (n0 - 1.0);
    if (n1 < 2.0) v1 = nan;
    if (n0 < 2.0) v0 = nan;
    double vn1 = v1 / n1, vn0 = v0 / n0;
    double df = (vn1 + vn0) * (vn1 + vn0) / (vn1 * vn1 / (n1 - 1.0) + vn0 * vn0 / (n0 - 1.0));
    if (isnan(df)) df = 1.0;
    double t = (m1 - m0) / sqrt(vn1 + vn0);
    return psb_t2_sf(t * t, df);
}

// chi2.sf(x, 1) = erfc(sqrt(x / 2))
PSB_HD double psb_chi2_sf1(double x) {
    if (isnan(x)) return x;
    if (x <= 0.0) return 1.0;
    return erfc(sqrt(0.5 * x));
}

// ln Gamma(a + 1/2) - ln Gamma(a): asymptotic series for large a (absolute error
// < 1e-16 for a >= 16), direct lgamma below.
PSB_HD double psb_lgamma_half_diff(double a) {
    if (a >= 16.0) {
        double r = 1.0 / a, r2 = r * r;
        // 1/2 ln a - 1/(8a) + 1/(192 a^3) - 1/(640 a^5) + 17/(14336 a^7) - 31/(18432 a^9)
        double s = r * (-0.125 + r2 * (1.0 / 192.0 + r2 * (-1.0 / 640.0 +
                   r2 * (17.0 / 14336.0 + r2 * (-31.0 / 18432.0)))));
        return 0.5 * log(a) + s;
    }
    return lgamma(a + 0.5) - lgamma(a);
}

// Continued fraction for the incomplete beta function (modified Lentz).
PSB_HD double psb_betacf(double a, double b, double x) {
    const double FPMIN = 1e-300, EPS = 1e-16;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < FPMIN) d = FPMIN;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 20000; ++m) {
        double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) <= EPS) break;
    }
    return h;
}

// Survival function of t^2 with df degrees of freedom evaluated at t2 >= 0:
//   P(T^2 > t2) = I_{df/(df+t2)}(df/2, 1/2)  ( = f.sf(t2, 1, df) = 2 t.sf(sqrt(t2), df) )
PSB_HD double psb_t2_sf(double t2, double df) {
    if (isnan(t2) || isnan(df)) return NAN;
    if (!(df > 0.0)) return NAN;
    if (t2 <= 0.0) return 1.0;
    if (isinf(t2)) return 0.0;
    double a = 0.5 * df, b = 0.5;
    double x = df / (df + t2);          // in (0,1)
    double xc = t2 / (df + t2);         // 1 - x without cancellation
    double lnx = (xc < 0.5) ? log1p(-xc) : log(x);
    double ln1mx = (x < 0.5) ? log1p(-x) : log(xc);
    // ln B(a, 1/2) = lgamma(a) + lgamma(1/2) - lgamma(a + 1/2)
    double lnB = 0.5723649429247000870717135 - psb_lgamma_half_diff(a);
    double bt = exp(a * lnx + b * ln1mx - lnB);
    if (x < (a + 1.0) / (a + b + 2.0))
        return bt * psb_betacf(a, b, x) / a;
    return 1.0 - bt * psb_betacf(b, a, xc) / b;
}
