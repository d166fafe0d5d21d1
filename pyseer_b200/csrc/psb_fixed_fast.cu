// psb_fixed_fast.cu -- the fast path of the batched Logit fit (model.fixed_effects_regression,
// model.py:315-352; statsmodels Logit.fit(method='newton')), for designs of up to 12 columns.
//
// Same estimator as k_fixed_logit (psb_fixed.cu), which stays the reference-faithful path: a
// converged Newton run ends at the unique maximiser of a concave likelihood whatever the start and
// whatever (positive definite) matrix scaled its steps, so this kernel is free to choose both, as
// long as the score it drives to zero is exact.  A variant that does not converge cleanly here --
// separation, a singular matrix, a fit far from the null model -- is handed to k_fixed_logit, which
// repeats statsmodels' iteration literally and decides the flags.
//
// What makes it fast (one warp per variant, lane = sample of a 32-sample word, as before):
//   * the Z block of X'WX is accumulated as the DIFFERENCE to its value at the null fit,
//         Hzz = Z'W0Z + sum_i (w_i - w0_i) z_i z_i',
//     in FP32 (66 FFMA per sample at q = 11 instead of 66 DFMA; half the registers): Z'W0Z is
//     exact (once per run, fp64), and the difference is small next to it for a fit near the null
//     model -- every variant but the few with a strong effect -- so its FP32 rounding (1e-7
//     relative to the difference) stays below 1e-8 of Hzz.  Fits that move further than
//     |d eta| <= 1 from the null model are sent to the exact kernel.
//   * the variant column is binary: its border of X'WX and its score are masked sums (no multiplies
//     by x, hxx = hzx_0), the intercept column is never loaded or multiplied;
//   * the score, the border and eta stay FP64;
//   * covariate columns are stored interleaved per 32-sample word, [word][column][lane], in fp64 and
//     in fp32: every load of the sample loop is one coalesced line at a constant offset;
//   * design width is a template parameter and padding columns are physical zero columns: no
//     per-column predicates in the sample loop.
// The iteration itself is the warm-started one of k_fixed_logit: null-model parameters, first step
// in closed form from the masked sums of the linear tensor tile, full evaluations until a step falls
// below 1e-7, results from that same evaluation (llf by the quadratic correction, bse from the
// factored matrix).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "psb_internal.cuh"
#include "psb_fixed_dev.cuh"
#include "psb_tc_ptx.cuh"

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#define FF_CH 8          // 32-sample words per staged chunk
#define FF_WARPS 4
#define FF_NBUF 3        // staged tiles: chunk ch + 2 is fetched while chunk ch is consumed

template <int Q>
struct FastAcc {
    double g[Q];        // Z'(y - pi)            (g[0]: intercept)
    double gx;          // x'(y - pi)
    float dhz[Q];       // sum over carriers of (w - w0) z_c   (dhz[0]: of w - w0); the null-model part
                        // sum over carriers of w0 z_c is exact (masked sums of the linear tensor tile)
    float dH[Q * (Q + 1) / 2];   // sum (w - w0) z_c z_d, d <= c
    double lprod;       // running product of the samples' likelihoods, exponent kept apart in lexp:
    int lexp;           //   llf = log(lprod) + lexp ln 2  (one log per lane and pass instead of one per sample)
    double maxdev;
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int c = 0; c < Q; ++c) {
            g[c] = 0.0;
            dhz[c] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < Q * (Q + 1) / 2; ++e) dH[e] = 0.f;
        gx = maxdev = 0.0;
        lprod = 1.0;
        lexp = 0;
    }
};

// exp(x) to ~1e-14 relative: x = (32 k + j) ln2/32 + r, |r| <= ln2/64, exp = 2^k * 2^(j/32) * e^r with
// a 32-entry table (shared memory) and a degree-5 polynomial (Estrin form: dependency depth 3); 10 FP64
// operations against ~35 of the correctly rounded library exp.  The power of two is clamped to
// 2^+-1000 with two integer operations (exp(-eta) only feeds 1 / (1 + .), which saturates long before).
__constant__ double c_ff[16] = {46.16624130844682903551758979206, /* [0] 32 / ln 2 */
                                0.021660849392498290195,          /* [1] ln 2 / 32 */
                                1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, /* [2..5] */
                                6755399441055744.0,               /* [6] 1.5 * 2^52 */
                                0.69314718055994530942,           /* [7] ln 2 */
                                2.0 / 3.0, 2.0 / 5.0, 2.0 / 7.0, 2.0 / 9.0, 2.0 / 11.0, 2.0 / 13.0, 2.0 / 15.0, /* [8..14] */
                                2.0 / 17.0};
__device__ __forceinline__ double ff_exp(double x, const double *__restrict__ tab) {
    const double t = fma(x, c_ff[0], c_ff[6]);                  // low word of t = rint(x * 32 / ln2)
    const int n = __double2loint(t);
    const double kf = t - c_ff[6];
    const double r = fma(-kf, c_ff[1], x);
    const double r2 = r * r;
    const double pa = fma(r, c_ff[2], c_ff[3]);                 // 1/24 + r/120
    const double pb = fma(r, c_ff[4], c_ff[5]);                 // 1/2 + r/6
    const double pc = 1.0 + r;
    const double p = fma(fma(pa, r2, pb), r2, pc);
    const double s = tab[n & 31] * p;
    const int k = max(min(n >> 5, 1000), -1000);
    return __hiloint2double(__double2hiint(s) + (k << 20), __double2loint(s));
}
// 1 / d for d >= 2^-1000: hardware seed (MUFU.RCP64H, ~2^-23) and one Newton step (2^-46)
__device__ __forceinline__ double ff_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    return fma(y, e, y);
}
// the same with two Newton steps (full double precision): the log-likelihood evaluations, where the
// one-step error -- always of one sign -- would add up over the N samples
__device__ __forceinline__ double ff_rcp2(double d) {
    double y = ff_rcp(d);
    const double e = fma(-d, y, 1.0);
    return fma(y, e, y);
}
// log(p) for p in (0, 1] to ~1e-15 absolute: p = m 2^k, m in [0.75, 1.5), log m = 2 atanh(s),
// s = (m - 1) / (m + 1), |s| <= 0.2, odd series to s^17
__device__ __forceinline__ double ff_log(double p) {
    int hi = __double2hiint(p);
    int k = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;                        // m in [1, 2)
    if (hi >= 0x3ff80000) { hi -= 0x00100000; ++k; }            // m in [0.75, 1.5)
    const double m = __hiloint2double(hi, __double2loint(p));
    const double sgm = (m - 1.0) * ff_rcp2(m + 1.0);
    const double s2 = sgm * sgm;
    double q = fma(s2, c_ff[15], c_ff[14]);
    q = fma(q, s2, c_ff[13]);
    q = fma(q, s2, c_ff[12]);
    q = fma(q, s2, c_ff[11]);
    q = fma(q, s2, c_ff[10]);
    q = fma(q, s2, c_ff[9]);
    q = fma(q, s2, c_ff[8]);
    q = fma(q, s2, 2.0);
    return fma((double)k, c_ff[7], q * sgm);
}

// Two 32-sample words at once, in straight-line code: the two dependency chains (eta -> exp ->
// reciprocal -> weights -> accumulations) interleave in the instruction stream, which is what keeps
// the pipes busy at two warps per scheduler.  zp / fp / wp point at the first word's covariates in the
// staged tile (the second word follows at +ZW / +32).
template <int Q>
__device__ __forceinline__ void fast_pair(const double *__restrict__ zp, const float *__restrict__ fp,
                                          const double *__restrict__ wp, const uint32_t xbits,
                                          const uint32_t ybits, const uint32_t okbits,
                                          const double (&b)[Q + 1], const double *__restrict__ etab,
                                          FastAcc<Q> &acc) {
    constexpr int ZW = (Q - 1) * 32;
    double z[2][Q];
    float zf[2][Q];
    double w0[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int c = 1; c < Q; ++c) {
            z[u][c] = zp[u * ZW + (c - 1) * 32];
            zf[u][c] = fp[u * ZW + (c - 1) * 32];
        }
        w0[u] = wp[u * 32];
    }
    double ex[2], pi[2], w[2], r[2], eta[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        // eta = beta_0 + sum_c beta_c z_c + x beta_x  (four chains)
        const bool xb = (xbits >> u) & 1u;
        double e0 = b[0], e1 = xb ? b[Q] : 0.0, e2 = 0.0, e3 = 0.0;
#pragma unroll
        for (int c = 1; c < Q; c += 4) {
            e0 = fma(b[c], z[u][c], e0);
            if (c + 1 < Q) e1 = fma(b[c + 1], z[u][c + 1], e1);
            if (c + 2 < Q) e2 = fma(b[c + 2], z[u][c + 2], e2);
            if (c + 3 < Q) e3 = fma(b[c + 3], z[u][c + 3], e3);
        }
        eta[u] = (e0 + e1) + (e2 + e3);
    }
    // pi = 1 / (1 + exp(-eta)), w = pi (1 - pi) as statsmodels writes them, with exp and the
    // reciprocal evaluated to 1e-13 relative (ff_exp, ff_rcp): the score only needs pi to ~1e-9 for
    // coefficients good to 1e-10 (random errors average out over N samples); the exact kernel keeps
    // the correctly rounded library functions where saturation decides flags
#pragma unroll
    for (int u = 0; u < 2; ++u) ex[u] = ff_exp(-eta[u], etab);
#pragma unroll
    for (int u = 0; u < 2; ++u) pi[u] = ff_rcp2(1.0 + ex[u]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const bool yb = (ybits >> u) & 1u, ok = (okbits >> u) & 1u;
        w[u] = ok ? pi[u] * (1.0 - pi[u]) : 0.0;
        r[u] = ok ? (yb ? 1.0 : 0.0) - pi[u] : 0.0;
        acc.maxdev = fmax(acc.maxdev, fabs(r[u]));
        // likelihood of the sample: cdf((2y-1) eta) = pi (y = 1) or 1 - pi = exp(-eta) pi (y = 0),
        // multiplied into the running product (relative rounding 1e-16 per factor: the log of the
        // product is good to 1e-14 absolute, better than a sum of N rounded logs)
        acc.lprod *= ok ? (yb ? pi[u] : ex[u] * pi[u]) : 1.0;
    }
    {
        // renormalise the product: its binary exponent moves to lexp (two factors >= 2^-1022 each per
        // step cannot underflow a product kept in [1, 2))
        const int hi = __double2hiint(acc.lprod);
        acc.lexp += (hi >> 20) - 1023;
        acc.lprod = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(acc.lprod));
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const bool xb = (xbits >> u) & 1u;
        acc.g[0] += r[u];
        acc.gx += xb ? r[u] : 0.0;
#pragma unroll
        for (int c = 1; c < Q; ++c) acc.g[c] = fma(r[u], z[u][c], acc.g[c]);
    }
    // FP32: differences to the null-model values of the Z block and of the variant's border
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const bool xb = (xbits >> u) & 1u;
        const float dw = (float)(w[u] - w0[u]);
        const float dwx = xb ? dw : 0.f;
        acc.dH[0] += dw;
        acc.dhz[0] += dwx;
#pragma unroll
        for (int c = 1; c < Q; ++c) {
            const float wz = dw * zf[u][c];
            acc.dhz[c] = fmaf(dwx, zf[u][c], acc.dhz[c]);
            acc.dH[c * (c + 1) / 2] += wz;
#pragma unroll
            for (int d = 1; d < Q; ++d)
                if (d <= c) acc.dH[c * (c + 1) / 2 + d] = fmaf(wz, zf[u][d], acc.dH[c * (c + 1) / 2 + d]);
        }
    }
}

template <int Q>
__device__ __forceinline__ void fast_chunk(const double *__restrict__ zt, const float *__restrict__ ft,
                                           const double *__restrict__ wt, const uint32_t *__restrict__ xrow,
                                           const uint32_t *__restrict__ y1, const uint32_t *__restrict__ vbits,
                                           int t0, int wlast, int lane, const double (&b)[Q + 1],
                                           const double *__restrict__ etab, FastAcc<Q> &acc) {
    constexpr int ZW = (Q - 1) * 32;
#pragma unroll 1
    for (int k = 0; k < FF_CH; k += 2) {
        // words beyond the last one are zero padded in the staged arrays and masked by vbits; the
        // variant row and the phenotype row are read at a clamped index
        const int ta = min(t0 + k, wlast), tb = min(t0 + k + 1, wlast);
        const uint32_t xa = __ldg(xrow + ta), xb = __ldg(xrow + tb);
        const uint32_t ya = __ldg(y1 + ta), yb = __ldg(y1 + tb);
        const uint32_t va = __ldg(vbits + t0 + k), vb = __ldg(vbits + t0 + k + 1);
        const uint32_t xbits = ((xa >> lane) & 1u) | (((xb >> lane) & 1u) << 1);
        const uint32_t ybits = ((ya >> lane) & 1u) | (((yb >> lane) & 1u) << 1);
        const uint32_t okbits = ((va >> lane) & 1u) | (((vb >> lane) & 1u) << 1);
        fast_pair<Q>(zt + k * ZW, ft + k * ZW, wt + k * 32, xbits, ybits, okbits, b, etab, acc);
    }
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// one bulk copy global -> shared, completion counted in bytes on an mbarrier (UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
#define FF_BULK 1        // 1: a tile is three bulk copies issued by one thread; 0: 16-byte cp.async by all

// Execution model.  A CTA of FF_WARPS warps sweeps the samples in lockstep PASSES: every pass is one
// evaluation (score, X'WX, llf) of the variant each warp currently owns.  The covariate columns of a
// chunk of FF_CH words (fp64 + fp32 + null weights: 32 KB at q = 11) are staged once per CTA into a
// ring of three shared-memory tiles with cp.async (one CTA barrier per chunk), so that the sample loop reads them at
// shared-memory latency and the L2 -> SM traffic is a quarter of what per-warp loads would cost.
// (A ring of bulk copies completing on mbarriers, which lets the warps drift apart by a chunk or two
// instead of meeting at a CTA barrier per chunk, was measured SLOWER -- 9.8 against 11.2 M k-mers/s:
// the warps spinning on mbarrier.try_wait take issue slots from the one still computing, while a
// warp parked at bar.sync costs nothing.)
// Between passes every warp solves its Newton step and either keeps its variant for another pass,
// publishes it, or hands it to the exact kernel; free warps take the next variant from a global
// work counter.
template <int Q, int MINB>
__global__ void __launch_bounds__(FF_WARPS * 32, MINB)
k_fixed_logit_fast(FxArgs a, FxFast ff, const int32_t *__restrict__ idx, int n_tested, int *__restrict__ work) {
    constexpr int P = Q + 1;                 // columns: Z_0 .. Z_{Q-1} (zero padded beyond a.q), x
    constexpr int ZW = (Q - 1) * 32;         // covariate values per word
    extern __shared__ __align__(16) unsigned char ff_smem[];
    double *sZ = reinterpret_cast<double *>(ff_smem);                         // [FF_NBUF][FF_CH][Q-1][32]
    double *sW0 = sZ + FF_NBUF * FF_CH * ZW;                                  // [FF_NBUF][FF_CH][32]
    float *sZf = reinterpret_cast<float *>(sW0 + FF_NBUF * FF_CH * 32);      // [FF_NBUF][FF_CH][Q-1][32]
    double *s_H0 = reinterpret_cast<double *>(sZf + FF_NBUF * FF_CH * ZW);    // packed Tri<Q>
    double *s_beta = s_H0 + Q * (Q + 1) / 2;                                  // [FF_WARPS][P]
    double *s_etab = s_beta + FF_WARPS * P;                                   // 2^(j/32), j = 0..31
    for (int e = threadIdx.x; e < Q * (Q + 1) / 2; e += blockDim.x) s_H0[e] = ff.Hzz0[e];
    if (threadIdx.x < 32) s_etab[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / 32.0));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *beta = s_beta + warp * P;
    const double inv_n = 1.0 / (double)a.N;
    const int nch = (a.Wn + FF_CH - 1) / FF_CH;

    int v = -1;                 // variant owned by this warp (-1: none)
    uint32_t f = 0;
    bool exhausted = false;
    double maxstep = INFINITY;
    int it = 0, n_eval = 0;
    const uint32_t *xrow = a.bits;
    const double *sv = a.sums;     // this variant's masked sums at the null fit: [c] = sum_carriers w0 z_c
    __syncthreads();

#if FF_BULK
    // Staging by the bulk-copy engine: one thread posts the three pieces of a tile (20 + 10 + 2 KB at
    // q = 11) on the tile's mbarrier; the 1920 16-byte cp.async instructions a tile cost before were
    // ~8 % of the CTA's issue slots and a third of its LSU traffic.
    uint64_t *s_mb = reinterpret_cast<uint64_t *>(s_etab + 32);
    if (threadIdx.x == 0) {
        for (int i = 0; i < FF_NBUF; ++i) mbar_init(smem_u32(&s_mb[i]), 1);
        fence_barrier_init();
    }
    __syncthreads();
    uint32_t mb_phase = 0;             // bit b: parity of tile b's next completion
    auto stage = [&](int ch, int buf) {
        if (threadIdx.x == 0) {
            const uint32_t bar = smem_u32(&s_mb[buf]);
            mbar_arrive_expect_tx(bar, (uint32_t)(FF_CH * ZW * 12 + FF_CH * 32 * 8));
            bulk_g2s(smem_u32(sZ + (size_t)buf * FF_CH * ZW), ff.Zi + (size_t)ch * FF_CH * ZW, FF_CH * ZW * 8, bar);
            bulk_g2s(smem_u32(sZf + (size_t)buf * FF_CH * ZW), ff.Zf + (size_t)ch * FF_CH * ZW, FF_CH * ZW * 4, bar);
            bulk_g2s(smem_u32(sW0 + (size_t)buf * FF_CH * 32), ff.W0 + (size_t)ch * FF_CH * 32, FF_CH * 32 * 8, bar);
        }
    };
#else
    auto stage = [&](int ch, int buf) {
        // chunk ch of the three arrays (each padded to whole chunks) -> buffer buf
        const char *gz = reinterpret_cast<const char *>(ff.Zi + (size_t)ch * FF_CH * ZW);
        const char *gf = reinterpret_cast<const char *>(ff.Zf + (size_t)ch * FF_CH * ZW);
        const char *gw = reinterpret_cast<const char *>(ff.W0 + (size_t)ch * FF_CH * 32);
        char *dz = reinterpret_cast<char *>(sZ + (size_t)buf * FF_CH * ZW);
        char *df = reinterpret_cast<char *>(sZf + (size_t)buf * FF_CH * ZW);
        char *dw = reinterpret_cast<char *>(sW0 + (size_t)buf * FF_CH * 32);
        for (int e = threadIdx.x; e < FF_CH * ZW * 8 / 16; e += blockDim.x) cp_async16(dz + e * 16, gz + e * 16);
        for (int e = threadIdx.x; e < FF_CH * ZW * 4 / 16; e += blockDim.x) cp_async16(df + e * 16, gf + e * 16);
        for (int e = threadIdx.x; e < FF_CH * 32 * 8 / 16; e += blockDim.x) cp_async16(dw + e * 16, gw + e * 16);
        cp_async_commit();
    };

#endif

    for (;;) {
        // ---- pass boundary: free warps take the next variant ---------------------------------
        while (v < 0 && !exhausted) {
            int t = 0;
            if (lane == 0) t = atomicAdd(work, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= n_tested) { exhausted = true; break; }
            const int cand = idx[t];
            f = a.flags[cand];
            if (f & PSB_F_MISSING_DATA) continue;                       // model.py:371-377
            if (f & PSB_F_BAD_CHISQ) {                                  // model.py:326: straight to Firth
                if (lane == 0) a.firth_list[atomicAdd(&a.counters[3], 1)] = cand;
                continue;
            }
            v = cand;
            xrow = a.bits + (size_t)v * a.Wrow;
            __syncwarp();
            if (lane < P) beta[lane] = lane < a.q ? a.warm[lane] : 0.0;
            __syncwarp();
            // first Newton step from the null parameters in closed form (see k_fixed_logit):
            //   schur = hxx - hx' Hzz^-1 hx,  d_x = g_x / schur,  d_z = -Hzz^-1 hx d_x
            sv = a.sums + (size_t)v * a.sums_ld;
            double hx[Q], tz[Q];
#pragma unroll
            for (int c = 0; c < Q; ++c) hx[c] = (c < a.q) ? sv[c] : 0.0;
            const double gx0 = sv[a.q];
            double quad = 0.0;
#pragma unroll
            for (int c = 0; c < Q; ++c) {
                double s = 0.0;
                if (c < a.q) {
#pragma unroll
                    for (int d = 0; d < Q; ++d)
                        if (d < a.q) s = fma(__ldg(a.HzzInv + c * a.q + d), hx[d], s);
                }
                tz[c] = s;
                quad = fma(s, hx[c], quad);
            }
            const double dk = gx0 / (hx[0] - quad);
            if (isfinite(dk)) {
                __syncwarp();
#pragma unroll
                for (int c = 0; c < Q; ++c)
                    if (lane == 0 && c < a.q) beta[c] -= tz[c] * dk;
                if (lane == 0) beta[Q] = dk;
                __syncwarp();
            }
            maxstep = INFINITY;
            it = 0;
            n_eval = 0;
        }
        if (!__syncthreads_or(v >= 0)) break;          // no warp of the CTA has work left

        // ---- one pass over the samples -----------------------------------------------------------
        const bool active = v >= 0;
        FastAcc<Q> acc;
        acc.clear();
        double breg[Q + 1];                  // this pass's parameters, in registers for the whole sweep
#pragma unroll
        for (int c = 0; c <= Q; ++c) breg[c] = beta[c];
        // Three tiles, ONE barrier per chunk: the tile refilled during chunk ch (for chunk ch + 2) was
        // read during chunk ch - 1, which every warp left before it passed this chunk's barrier.
        stage(0, 0);
        if (nch > 1) stage(1, 1);
        for (int ch = 0; ch < nch; ++ch) {
            const int buf = ch % FF_NBUF;
#if FF_BULK
            mbar_wait(smem_u32(&s_mb[buf]), (mb_phase >> buf) & 1u);
            mb_phase ^= 1u << buf;
            __syncthreads();
            if (ch + 2 < nch) stage(ch + 2, (ch + 2) % FF_NBUF);
#else
            if (ch + 1 < nch) cp_async_wait<1>();
            else cp_async_wait<0>();
            __syncthreads();
            if (ch + 2 < nch) stage(ch + 2, (ch + 2) % FF_NBUF);
            else cp_async_commit();        // keep one group per iteration so that wait<1> counts right
#endif
            if (active) {
                const double *zt = sZ + (size_t)buf * FF_CH * ZW + lane;
                const float *ft = sZf + (size_t)buf * FF_CH * ZW + lane;
                const double *wt = sW0 + (size_t)buf * FF_CH * 32 + lane;
                fast_chunk<Q>(zt, ft, wt, xrow, a.y1, ff.vbits, ch * FF_CH, a.Wn - 1, lane, breg, s_etab, acc);
            }
        }
        __syncthreads();                   // the tiles are free before the next pass refills them
        if (!active) continue;

        // ---- this warp's Newton step -----------------------------------------------------------
#pragma unroll
        for (int c = 0; c < Q; ++c) {
            acc.g[c] = warp_sum(acc.g[c]);
            acc.dhz[c] = warp_sum_f(acc.dhz[c]);
        }
        acc.gx = warp_sum(acc.gx);
#pragma unroll
        for (int e = 0; e < Q * (Q + 1) / 2; ++e) acc.dH[e] = warp_sum_f(acc.dH[e]);
        acc.maxdev = warp_max(acc.maxdev);
        // log-likelihood at the evaluation point: one log per lane
        const double pass_llf = warp_sum(fma((double)acc.lexp, 0.69314718055994530942, log(acc.lprod)));
        ++n_eval;
        bool slow = false, done = false;
        double bse = NAN, llf = NAN;
        if (acc.maxdev <= 1e-8) {
            slow = true;                                               // _check_perfect_pred territory
        } else {
            // (X'WX/n - 1e-10 I) step = score/n   (statsmodels' ridge sign, see fx_ldl)
            double H[Tri<P>::SIZE], g[P];
#pragma unroll
            for (int c = 0; c < Q; ++c) {
#pragma unroll
                for (int d = 0; d < Q; ++d)
                    if (d <= c)
                        H[Tri<P>::at(c, d)] = (s_H0[c * (c + 1) / 2 + d] + (double)acc.dH[c * (c + 1) / 2 + d]) * inv_n;
                // border: exact null-model part (masked sums of the tensor tile) + FP32 difference
                H[Tri<P>::at(Q, c)] = ((c < a.q ? sv[c] : 0.0) + (double)acc.dhz[c]) * inv_n;
                g[c] = acc.g[c] * inv_n;
            }
            H[Tri<P>::at(Q, Q)] = (sv[0] + (double)acc.dhz[0]) * inv_n;
            g[Q] = acc.gx * inv_n;
#pragma unroll
            for (int c = 0; c < P; ++c)
                if (c < a.q || c == Q) H[Tri<P>::at(c, c)] -= 1e-10;
            if (!fx_ldl<P>(H)) {
                slow = true;
            } else {
                double gs[P];
#pragma unroll
                for (int c = 0; c < P; ++c) gs[c] = g[c];
                fx_ldl_solve<P>(H, g);
                maxstep = 0.0;
                double gd = 0.0;
                __syncwarp();
#pragma unroll
                for (int c = 0; c < P; ++c) {
                    if (lane == 0) beta[c] += g[c];
                    maxstep = fmax(maxstep, fabs(g[c]));
                    gd = fma(gs[c], g[c], gd);
                }
                __syncwarp();
                ++it;
                if (!(maxstep < 1e3)) {
                    slow = true;                                       // NaN or running away
                } else if (maxstep <= 1e-7) {
                    // converged: llf(beta + d) = llf(beta) + g'd/2 + O(d^3), bse from the factored matrix
                    double e[P];
#pragma unroll
                    for (int c = 0; c < P; ++c) e[c] = (c == Q) ? 1.0 : 0.0;
                    fx_ldl_solve<P>(H, e);
                    const double var_x = e[Q] * inv_n;
                    if (var_x > 0.0 && isfinite(var_x)) {
                        bse = sqrt(var_x);
                        llf = pass_llf + 0.5 * (double)a.N * gd;
                        done = true;
                    } else {
                        slow = true;
                    }
                } else if (it >= 12) {
                    slow = true;
                }
            }
        }
        if (!done && !slow) continue;                   // another pass for the same variant
        if (lane == 0) atomicAdd(&a.counters[4], n_eval);
        if (done) {
            // the FP32 difference is trusted for fits within |d eta| <= 1 of the null model
            double disp = fabs(beta[Q]);
#pragma unroll
            for (int c = 0; c < Q; ++c)
                if (c < a.q) disp = fma(fabs(beta[c] - a.warm[c]), ff.zmax[c], disp);
            if (!(disp <= 1.0)) { done = false; slow = true; }
        }
        if (lane == 0) {
            if (!done) {
                ff.slow_list[atomicAdd(&a.counters[6], 1)] = v;
            } else if (bse > 3.0) {                                  // model.py:332-334
                a.flags[v] = f | PSB_F_HIGH_BSE;
                a.firth_list[atomicAdd(&a.counters[3], 1)] = v;
            } else {
                const double lrstat = -2.0 * (a.null_llf - llf);        // model.py:336-339
                double p = 1.0;
                if (lrstat > 0.0) p = psb_chi2_sf1(lrstat);
                const double kbeta = beta[Q];
                uint32_t fo = f;
                if (p > a.lrt_pvalue || !isfinite(p) || !isfinite(kbeta)) {  // model.py:384
                    fo |= PSB_F_LRT_FAILED | PSB_F_FILTER;
                    atomicAdd(&a.counters[2], 1);
                }
                a.pvalue[v] = p;
                a.beta[v] = kbeta;
                a.bse[v] = bse;
                a.intercept[v] = beta[0];
                for (int c = 1; c < a.q; ++c) a.betas[(size_t)v * (a.q - 1) + (c - 1)] = beta[c];
                a.flags[v] = fo;
            }
        }
        __syncwarp();
        v = -1;
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int fast_width(int q) { return q + 1 <= 4 ? 3 : (q + 1 <= 8 ? 7 : 11); }

// Once per run (psb_fixed_setup): interleaved covariate columns in fp64 and fp32, null-model
// weights, the exact Z block of X'WX at the null fit (packed lower triangle, unit diagonal on the
// padding columns) and max |z_c| per column.
int psb_fixed_fast_setup(psb_ctx *c, const double *Z, const double *warm) {
    const int N = c->N, q = c->q;
    if (q + 1 > FX_MAXP) return PSB_OK;
    const int Q = fast_width(q);
    const int Wn = (c->Wn + FF_CH - 1) / FF_CH * FF_CH;          // whole chunks (zero padded)
    std::vector<double> Zi((size_t)Wn * (Q - 1) * 32, 0.0), W0((size_t)Wn * 32, 0.0), H0((size_t)Q * (Q + 1) / 2, 0.0);
    std::vector<float> Zf(Zi.size(), 0.f);
    std::vector<uint32_t> vb(Wn, 0u);                      // sample < N, zero on the padding words
    for (int i = 0; i < N; ++i) vb[i >> 5] |= 1u << (i & 31);
    std::vector<double> zmax(FX_MAXP, 0.0);
    for (int i = 0; i < N; ++i) {
        const double *zi = Z + (size_t)i * q;
        double eta = 0.0;
        for (int k = 0; k < q; ++k) eta += warm[k] * zi[k];
        const double pi = 1.0 / (1.0 + exp(-eta));
        const double w0 = pi * (1.0 - pi);
        W0[i] = w0;
        const int t = i >> 5, lane = i & 31;
        for (int k = 0; k < q; ++k) {
            zmax[k] = std::max(zmax[k], fabs(zi[k]));
            if (k >= 1) {
                Zi[((size_t)t * (Q - 1) + (k - 1)) * 32 + lane] = zi[k];
                Zf[((size_t)t * (Q - 1) + (k - 1)) * 32 + lane] = (float)zi[k];
            }
            // the kernel adds fp32(w - w0) * fp32(z_c) * fp32(z_d): the base must be the sum of
            // w0 * fp32(z_c) * fp32(z_d)?  No: the base is the exact Z'W0Z; the fp32 rounding of z
            // only touches the (small) difference.
            for (int d = 0; d <= k; ++d) H0[(size_t)k * (k + 1) / 2 + d] += w0 * zi[k] * zi[d];
        }
    }
    for (int k = q; k < Q; ++k) H0[(size_t)k * (k + 1) / 2 + k] = 1.0;
    auto up = [&](void **dst, const void *src, size_t bytes) -> int {
        PSB_CUDA(cudaMalloc(dst, bytes));
        PSB_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return PSB_OK;
    };
    int rc;
    if ((rc = up((void **)&c->d_fx_Zi, Zi.data(), Zi.size() * sizeof(double)))) return rc;
    if ((rc = up((void **)&c->d_fx_Zf, Zf.data(), Zf.size() * sizeof(float)))) return rc;
    if ((rc = up((void **)&c->d_fx_W0, W0.data(), W0.size() * sizeof(double)))) return rc;
    if ((rc = up((void **)&c->d_fx_H0, H0.data(), H0.size() * sizeof(double)))) return rc;
    if ((rc = up((void **)&c->d_fx_vb, vb.data(), vb.size() * sizeof(uint32_t)))) return rc;
    c->fx_zmax.assign(zmax.begin(), zmax.end());
    c->fx_Q = Q;
    return PSB_OK;
}

void psb_fixed_fast_free(psb_ctx *c) {
    if (c->d_fx_Zi) cudaFree(c->d_fx_Zi);
    if (c->d_fx_Zf) cudaFree(c->d_fx_Zf);
    if (c->d_fx_W0) cudaFree(c->d_fx_W0);
    if (c->d_fx_H0) cudaFree(c->d_fx_H0);
    if (c->d_fx_vb) cudaFree(c->d_fx_vb);
    c->d_fx_vb = nullptr;
    c->d_fx_Zi = nullptr;
    c->d_fx_Zf = nullptr;
    c->d_fx_W0 = c->d_fx_H0 = nullptr;
    c->fx_Q = 0;
}

template <int Q>
static size_t fast_smem() {
    return (size_t)FF_NBUF * FF_CH * (Q - 1) * 32 * (8 + 4) + (size_t)FF_NBUF * FF_CH * 32 * 8 +
           ((size_t)Q * (Q + 1) / 2 + FF_WARPS * (Q + 1) + 32) * 8 + FF_NBUF * 8;
}

template <int Q, int MINB>
static int launch_fast_b(psb_ctx *c, const FxArgs &a, const FxFast &ff, int n, int *work) {
    const size_t smem = fast_smem<Q>();
    PSB_CUDA(cudaFuncSetAttribute(k_fixed_logit_fast<Q, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::min(psb_div_up(n, FF_WARPS), c->sm_count * MINB);
    k_fixed_logit_fast<Q, MINB><<<grid, FF_WARPS * 32, smem, c->stream>>>(a, ff, c->d_idx, n, work);
    return PSB_OK;
}

template <int Q>
static int launch_fast(psb_ctx *c, const FxArgs &a, const FxFast &ff, int n, int *work) {
    static const int minb = getenv("PSB_LOGIT_MINB") ? atoi(getenv("PSB_LOGIT_MINB")) : 2;
    if (minb == 3) return launch_fast_b<Q, 3>(c, a, ff, n, work);
    if (minb == 1) return launch_fast_b<Q, 1>(c, a, ff, n, work);
    return launch_fast_b<Q, 2>(c, a, ff, n, work);
}

// Launches the fast kernel over the tested variants; the variants it hands back are in
// c->d_idx3[0 .. counters[6]).
int psb_fixed_fast_launch(psb_ctx *c, const FxArgs &a, int n) {
    if (n <= 0) return PSB_OK;
    FxFast ff;
    ff.Zi = c->d_fx_Zi;
    ff.Zf = c->d_fx_Zf;
    ff.W0 = c->d_fx_W0;
    ff.Hzz0 = c->d_fx_H0;
    ff.vbits = c->d_fx_vb;
    ff.slow_list = c->d_idx3;
    for (int k = 0; k < FX_MAXP; ++k) ff.zmax[k] = c->fx_zmax[k];
    int *work = c->d_counters + 7;          // counters[7]: the work counter of this launch (zeroed by k_prefilter's memset)
    int rc;
    if (c->fx_Q == 3) rc = launch_fast<3>(c, a, ff, n, work);
    else if (c->fx_Q == 7) rc = launch_fast<7>(c, a, ff, n, work);
    else rc = launch_fast<11>(c, a, ff, n, work);
    if (rc) return rc;
    c->launches++;
    PSB_CUDA(cudaGetLastError());
    return PSB_OK;
}
