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

// Continued fraction for the incomplete beta function, evaluated by the forward recurrence of its
// numerators and denominators (Wallis), renormalised by the last denominator after every double
// step: two reciprocals per step -- one shared by both partial numerators, one for the
// renormalisation -- where the modified Lentz scheme divides six times.  This is the per-variant
// cost of the Welch pre-filter and of F.sf in the LMM epilogue (fp64 divisions are ~25 instructions).
PSB_HD double psb_betacf(double a, double b, double x) {
    const double TINY = 1e-290, EPS = 2.3e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double am = 1.0, bm = 1.0, az = 1.0, bz = 1.0 - qab * x / qap;
    for (int m = 1; m <= 20000; ++m) {
        const double em = (double)m, tem = em + em;
        const double u = qam + tem, v = a + tem, w = qap + tem;
        const double rden = x / (u * v * w);
        const double d1 = em * (b - em) * w * rden;
        const double ap = fma(d1, am, az), bp = fma(d1, bm, bz);
        const double d2 = -(a + em) * (qab + em) * u * rden;
        const double app = fma(d2, az, ap);
        double bpp = fma(d2, bz, bp);
        if (fabs(bpp) < TINY) bpp = TINY;
        const double r = 1.0 / bpp;
        const double aold = az;
        am = ap * r;
        bm = bp * r;
        az = app * r;
        bz = 1.0;
        if (fabs(az - aold) <= EPS * fabs(az)) break;
    }
    return az;
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
    // Tail: I_x(a, 1/2) by its continued fraction, which is fast and well conditioned for
    // x < (a + 1) / (a + b + 2) once t2 is past ~12 (<= 20 steps).
    // Bulk: p = 1 - I_{1-x}(1/2, a) with the hypergeometric series
    //     I_xc(1/2, a) = 2 bt sum_n T_n,  T_0 = 1,  T_{n+1} = T_n xc (a + 1/2 + n) / (n + 3/2)
    // -- positive terms, one reciprocal and four multiply-adds each, <= ~50 of them for t2 <= 12 --
    // instead of the textbook switch between the two continued fractions at t2 ~ 3, which puts the
    // slow side of BOTH (50 steps of six divisions at t2 = 3.1, df = 5000) where most null variants
    // are; a warp runs as long as its slowest lane.  Past the textbook switch point the
    // complementary continued fraction is also ill conditioned (its value grows to ~150 at t2 = 12),
    // the series is not.  Small df (a < 16) keep the textbook rule: the series converges like xc^n.
    const bool direct = x < (a + 1.0) / (a + b + 2.0);
    if (direct && (t2 > 12.0 || a < 16.0)) return bt * psb_betacf(a, b, x) / a;
    if (direct || a >= 16.0) {
        double term = 1.0, sum = 1.0;
        const double am = a - 1.0;
        for (int n = 0; n < 4000; ++n) {
            // (a + 1/2 + n) / (n + 3/2) = 1 + (a - 1) / (n + 3/2)
            const double r = 1.0 / ((double)n + 1.5);
            term *= xc * fma(am, r, 1.0);
            sum += term;
            if (term <= 1e-17 * sum) break;
        }
        return 1.0 - 2.0 * bt * sum;
    }
    return 1.0 - bt * psb_betacf(b, a, xc) / b;
}
