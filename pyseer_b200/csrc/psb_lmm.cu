// psb_lmm.cu -- linear mixed model path: setup of the rotated operands, the FP64
// contraction kernel, the per-variant epilogue and psb_run_lmm.
//
// Reference being replaced (per block of variants):
//   lmm.fit_lmm            lmm.py:125-226   (filters, notes, result tuples)
//   lmm.fit_lmm_block      lmm.py:228-260   (F-test, bse, frac_h2)
//   LMM.nLLeval / rotate / nLLcore / computeAKA / computeAKB
//                          fastlmm/lmm_cov.py:597-684, 165-194, 686-838, 885-916
//
// Algebra.  With P = I - X X^+, Sd = h2 S + (1 - h2),
// L = P U diag(Sd^-1/2)  (N x J, J = N - D),  v = P U diag(1/Sd) U' P y,  YKY = y' M y:
//   a = snpsKsnps = || L' x ||^2      b = snpsKY = x' v
//   beta = b / a,  var_beta = (YKY - b beta) / (J - 1) / a,  frac = b beta / YKY,
//   p = F.sf(beta^2 / var_beta; 1, N - (D + 1)).
// x is a 0/1 column, so L' x is a sum of the rows of L selected by the carrier bits: the
// FP64 kernel below does predicated adds only (no multiplies); the tensor-core kernel
// (psb_lmm_tc.cu) does the same contraction exactly in int8 slices.
#include <math.h>

#include <algorithm>

#include "psb_internal.cuh"
#include "psb_math.cuh"

// ------------------------------------------------------------------------------------
// FP64 contraction: a[v] = sum_j ( sum_i x[v,i] L[i,j] )^2
// CTA = 128 variants x all components; thread micro-tile 8 variants x 4 components.
// ------------------------------------------------------------------------------------
#define QF_BM 128
#define QF_BN 64
#define QF_BK 32
#define QF_THREADS 256

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__global__ void __launch_bounds__(QF_THREADS, 1)
k_lmm_quadform_fp64(const uint32_t *__restrict__ bits, int Wrow, const int32_t *__restrict__ idx,
                    const int *__restrict__ n_tested_dev, const double *__restrict__ L, int Lrows,
                    int Jpad, double *__restrict__ a_out) {
    const int n_tested = *n_tested_dev;
    __shared__ __align__(16) double Ls[2][QF_BK][QF_BN];
    __shared__ uint32_t Xs[2][QF_BM];
    __shared__ int32_t rows[QF_BM];

    const int tid = threadIdx.x;
    const int vg = tid >> 4;   // 0..15 -> variants vg*8 .. vg*8+7
    const int jg = tid & 15;   // 0..15 -> components jg*4 .. jg*4+3
    const int nkb = Lrows / QF_BK;
    const int njt = Jpad / QF_BN;

    for (int tile = blockIdx.x; tile * QF_BM < n_tested; tile += gridDim.x) {
        __syncthreads();
        if (tid < QF_BM) {
            int t = tile * QF_BM + tid;
            rows[tid] = t < n_tested ? idx[t] : -1;
        }
        __syncthreads();
        double part[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) part[v] = 0.0;

        const int total = njt * nkb;
        // prologue: stage 0
        auto issue = [&](int it, int buf) {
            int jt = it / nkb, kb = it - jt * nkb;
            const double *src = L + (size_t)(kb * QF_BK) * Jpad + jt * QF_BN;
            // 32 rows x 64 doubles = 1024 16-byte chunks; 4 per thread
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int e = tid + r * QF_THREADS;
                int row = e >> 5, c2 = e & 31;
                cp_async16(&Ls[buf][row][c2 * 2], src + (size_t)row * Jpad + c2 * 2);
            }
            if (tid < QF_BM) {
                int rr = rows[tid];
                if (rr >= 0) cp_async4(&Xs[buf][tid], bits + (size_t)rr * Wrow + kb);
                else Xs[buf][tid] = 0u;
            }
            cp_async_commit();
        };
        issue(0, 0);
        double acc[8][4];
        for (int it = 0; it < total; ++it) {
            int buf = it & 1;
            int kb = it % nkb;
            if (kb == 0) {
#pragma unroll
                for (int v = 0; v < 8; ++v)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[v][j] = 0.0;
            }
            if (it + 1 < total) {
                issue(it + 1, buf ^ 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            uint32_t w[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) w[v] = Xs[buf][vg * 8 + v];
#pragma unroll 8
            for (int k = 0; k < QF_BK; ++k) {
                const double2 *lp = reinterpret_cast<const double2 *>(&Ls[buf][k][jg * 4]);
                double2 l01 = lp[0], l23 = lp[1];
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    if ((w[v] >> k) & 1u) {
                        acc[v][0] += l01.x;
                        acc[v][1] += l01.y;
                        acc[v][2] += l23.x;
                        acc[v][3] += l23.y;
                    }
                }
            }
            if (kb == nkb - 1) {
#pragma unroll
                for (int v = 0; v < 8; ++v)
#pragma unroll
                    for (int j = 0; j < 4; ++j) part[v] = fma(acc[v][j], acc[v][j], part[v]);
            }
            __syncthreads();
        }
        // reduce over the 16 component groups (lanes differing in the low 4 bits)
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            double p = part[v];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
            if (jg == 0) {
                int rr = rows[vg * 8 + v];
                if (rr >= 0) a_out[rr] = p;
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// Epilogue: nLLcore tail (lmm_cov.py:799-815) + fit_lmm_block (lmm.py:247-258) +
// the lrt filter of fit_lmm (lmm.py:201-224).  One thread per tested variant.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_lmm_epilogue(const int *__restrict__ n_tested_dev, const int32_t *__restrict__ idx,
               const double *__restrict__ a_in,
               const double *__restrict__ b_in, const double *__restrict__ pp_in,
               const double *__restrict__ sums, int C, int col_b, int col_q0, int nq, int N,
               const int32_t *__restrict__ carriers, const int32_t *__restrict__ nmissing,
               double YKY, double dof1, double lrt_pvalue, double *__restrict__ pvalue,
               double *__restrict__ beta_out, double *__restrict__ bse_out,
               double *__restrict__ frac_out, uint32_t *__restrict__ flags,
               int *__restrict__ counters, int defer_welch, int col_w0, double T1, double T2,
               double filter_pvalue, double *__restrict__ prep_out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *n_tested_dev) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    int v = idx[t];
    uint32_t f = flags[v];
    if (defer_welch) {
        // pre_filtering (model.py:53-55) applied after the fact: the sums over carriers came out
        // of the tensor pass, the sums over non-carriers follow from the totals.  lmm.py:172 rule.
        const double *w = sums + (size_t)v * C + col_w0;
        const double n1 = (double)(carriers[v] - nmissing[v]), n0 = (double)(N - carriers[v]);
        const double prep = psb_welch_prep(w[0], w[1], T1 - w[0], T2 - w[1], n1, n0);
        prep_out[v] = prep;
        if (prep >= filter_pvalue || !isfinite(prep)) {
            flags[v] = (f & ~PSB_F_TESTED) | PSB_F_PREFILTER_FAILED | PSB_F_PREFILTER;
            atomicAdd(&counters[5], 1);
            return;
        }
    }
    double p, beta, var_beta, frac;
    if (nmissing[v] > 0) {
        // NaN genotypes propagate through rotate()/nLLcore: every statistic is NaN
        p = beta = var_beta = frac = nan;
    } else {
        const double *s = sums + (size_t)v * C;
        double a = a_in[v];
        // b = x'v and ||Q'x||^2: from the tensor pass when it carries them, else from the
        // masked column sums
        double b, pp = 0.0;
        if (b_in) {
            b = b_in[v];
            pp = pp_in[v];
        } else {
            b = s[col_b];
            for (int d = 0; d < nq; ++d) pp = fma(s[col_q0 + d], s[col_q0 + d], pp);
        }
        // rotate(): columns with std(P x) <= 1e-10 are zeroed (lmm_cov.py:179-181)
        double c = (double)carriers[v];
        double ss = c - pp;   // || P x ||^2
        if (ss <= fmax(1e-20 * (double)N, 1e-12 * c)) {
            a = 0.0;
            b = 0.0;
        }
        beta = b / a;
        if (isnan(beta) && b == 0.0) beta = 0.0;       // lmm_cov.py:803-805
        double veb = b * beta;
        double r2 = YKY - veb;
        var_beta = r2 / dof1 / a;                      // lmm_cov.py:813
        frac = veb / YKY;                              // lmm_cov.py:814
        double chi2 = beta * beta / var_beta;          // lmm.py:248
        p = psb_t2_sf(chi2, dof1);                     // lmm.py:251-253
    }
    if (p >= lrt_pvalue || !isfinite(p)) {             // lmm.py:201-207
        f |= PSB_F_LRT_FAILED | PSB_F_FILTER;
        pvalue[v] = p;
        atomicAdd(&counters[2], 1);
    } else {
        pvalue[v] = p;
        beta_out[v] = beta;
        bse_out[v] = sqrt(var_beta);
        frac_out[v] = sqrt(frac);
    }
    flags[v] = f;
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
static void pack_pheno_bits(const double *y, int N, int Wn, std::vector<uint32_t> &y1,
                            std::vector<uint32_t> &y0, std::vector<uint32_t> &valid) {
    y1.assign(Wn, 0u);
    y0.assign(Wn, 0u);
    valid.assign(Wn, 0u);
    for (int i = 0; i < N; ++i) {
        valid[i >> 5] |= 1u << (i & 31);
        if (y[i] == 1.0) y1[i >> 5] |= 1u << (i & 31);
        if (y[i] == 0.0) y0[i >> 5] |= 1u << (i & 31);
    }
}

int psb_upload_pheno(psb_ctx *c, const double *y) {
    std::vector<uint32_t> y1, y0, valid;
    pack_pheno_bits(y, c->N, c->Wn, y1, y0, valid);
    c->n_y1 = c->n_y0 = 0;
    for (int i = 0; i < c->N; ++i) {
        c->n_y1 += (y[i] == 1.0);
        c->n_y0 += (y[i] == 0.0);
    }
    PSB_CUDA(cudaMalloc(&c->d_y1bits, c->Wn * sizeof(uint32_t)));
    PSB_CUDA(cudaMalloc(&c->d_y0bits, c->Wn * sizeof(uint32_t)));
    PSB_CUDA(cudaMalloc(&c->d_valid, c->Wn * sizeof(uint32_t)));
    PSB_CUDA(cudaMemcpy(c->d_y1bits, y1.data(), c->Wn * sizeof(uint32_t), cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(c->d_y0bits, y0.data(), c->Wn * sizeof(uint32_t), cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(c->d_valid, valid.data(), c->Wn * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return PSB_OK;
}

// Welch columns: centred y and its square under the k==1 and k==0 masks.
void psb_fill_welch_cols(const double *y, int N, int Npad, double *cols, int col_w0,
                         uint64_t *mask_lo, uint64_t *mask_hi) {
    double mean = 0.0;
    for (int i = 0; i < N; ++i) mean += y[i];
    mean /= (double)(N > 0 ? N : 1);
    int kinds[4] = {PSB_MASK_K1, PSB_MASK_K1, PSB_MASK_K0, PSB_MASK_K0};
    for (int k = 0; k < 4; ++k) {
        int c = col_w0 + k;
        double *col = cols + (size_t)c * Npad;
        for (int i = 0; i < N; ++i) {
            double yc = y[i] - mean;
            col[i] = (k & 1) ? yc * yc : yc;
        }
        if (c < 32) *mask_lo |= (uint64_t)kinds[k] << (2 * c);
        else *mask_hi |= (uint64_t)kinds[k] << (2 * (c - 32));
    }
}

// Orthonormal basis of the column space of X (N x D row-major) by twice-applied modified
// Gram-Schmidt; dependent columns are dropped (pinv semantics of Linreg, lmm_cov.py:861-872).
int psb_orthobasis(const double *X, int N, int D, std::vector<double> &Q) {
    Q.clear();
    int r = 0;
    std::vector<double> col(N);
    for (int d = 0; d < D; ++d) {
        double n0 = 0.0;
        for (int i = 0; i < N; ++i) {
            col[i] = X[(size_t)i * D + d];
            n0 += col[i] * col[i];
        }
        n0 = sqrt(n0);
        if (n0 == 0.0) continue;
        for (int pass = 0; pass < 2; ++pass)
            for (int e = 0; e < r; ++e) {
                const double *qe = &Q[(size_t)e * N];
                double dot = 0.0;
                for (int i = 0; i < N; ++i) dot += qe[i] * col[i];
                for (int i = 0; i < N; ++i) col[i] -= dot * qe[i];
            }
        double n1 = 0.0;
        for (int i = 0; i < N; ++i) n1 += col[i] * col[i];
        n1 = sqrt(n1);
        if (n1 <= 1e-10 * n0) continue;
        Q.resize((size_t)(r + 1) * N);
        for (int i = 0; i < N; ++i) Q[(size_t)r * N + i] = col[i] / n1;
        ++r;
    }
    return r;
}

extern "C" int psb_lmm_setup(psb_ctx *c, int32_t N, int32_t D, const double *X, const double *y,
                             const double *U, const double *S, double h2, int32_t precision) {
    PSB_REQUIRE(c && X && y && U && S, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(N > 1 && D >= 1 && D < N, PSB_ERR_ARG, "bad shape N=%d D=%d", N, D);
    PSB_REQUIRE(precision == 0 || (precision >= 3 && precision <= 7) || precision == 46, PSB_ERR_ARG,
                "precision must be 0 (fp64), 3..7 int8 slices or 46 (two passes: 4 slices, 6 for the tail), got %d",
                precision);
    // lmm_cov.py:667-670: nLLeval returns no 'beta' for h2 outside [0,1) -> KeyError in
    // fit_lmm_block (tests/lmm_test.py:416-417)
    PSB_REQUIRE(h2 >= 0.0 && h2 < 1.0, PSB_ERR_H2, "h2 = %g outside [0, 1)", h2);
    PSB_CUDA(cudaSetDevice(c->device));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    psb_free_model(c);

    const int J = N - D;
    c->N = N;
    c->Wn = (N + 31) / 32;
    c->D = D;
    c->J = J;
    c->h2 = h2;
    c->precision = precision;

    std::vector<double> Q;
    const int r = psb_orthobasis(X, N, D, Q);

    // P y, with rotate()'s constant-column rule
    std::vector<double> Py(y, y + N);
    for (int e = 0; e < r; ++e) {
        const double *qe = &Q[(size_t)e * N];
        double dot = 0.0;
        for (int i = 0; i < N; ++i) dot += qe[i] * y[i];
        for (int i = 0; i < N; ++i) Py[i] -= dot * qe[i];
    }
    {
        double m = 0.0, s2 = 0.0;
        for (int i = 0; i < N; ++i) m += Py[i];
        m /= N;
        for (int i = 0; i < N; ++i) s2 += (Py[i] - m) * (Py[i] - m);
        if (sqrt(s2 / N) <= 1e-10) std::fill(Py.begin(), Py.end(), 0.0);
    }
    // P U = U - Q (Q' U)
    std::vector<double> PU(U, U + (size_t)N * J);
    std::vector<double> QtU((size_t)(r > 0 ? r : 1) * J, 0.0);
    for (int e = 0; e < r; ++e) {
        const double *qe = &Q[(size_t)e * N];
        double *out = &QtU[(size_t)e * J];
        for (int i = 0; i < N; ++i) {
            const double qi = qe[i];
            const double *ui = U + (size_t)i * J;
            for (int j = 0; j < J; ++j) out[j] += qi * ui[j];
        }
    }
    for (int i = 0; i < N; ++i) {
        double *pi = &PU[(size_t)i * J];
        for (int e = 0; e < r; ++e) {
            const double qi = Q[(size_t)e * N + i];
            const double *qu = &QtU[(size_t)e * J];
            for (int j = 0; j < J; ++j) pi[j] -= qi * qu[j];
        }
    }
    std::vector<double> Sd(J), UY(J, 0.0);
    for (int j = 0; j < J; ++j) Sd[j] = h2 * S[j] + (1.0 - h2);     // lmm_cov.py:665
    for (int i = 0; i < N; ++i) {
        const double yi = Py[i];
        const double *pi = &PU[(size_t)i * J];
        for (int j = 0; j < J; ++j) UY[j] += pi[j] * yi;
    }
    double YKY = 0.0;
    for (int j = 0; j < J; ++j) YKY += UY[j] / Sd[j] * UY[j];       // computeAKA
    c->YKY = YKY;

    int rc0 = PSB_OK;
    // column matrix for k_bitsums: v | Q_0..Q_{r-1} | Welch(4)
    c->Npad = c->Wn * 32;
    c->C = 1 + r + 4;
    c->col_b = 0;
    c->col_q0 = 1;
    c->col_w0 = 1 + r;
    c->colmask_lo = c->colmask_hi = 0;
    std::vector<double> cols((size_t)c->C * c->Npad, 0.0);
    {
        std::vector<double> wy(J);
        for (int j = 0; j < J; ++j) wy[j] = UY[j] / Sd[j];
        for (int i = 0; i < N; ++i) {
            const double *pi = &PU[(size_t)i * J];
            double acc = 0.0;
            for (int j = 0; j < J; ++j) acc += pi[j] * wy[j];
            cols[i] = acc;
        }
    }
    for (int e = 0; e < r; ++e)
        std::copy(&Q[(size_t)e * N], &Q[(size_t)e * N] + N, &cols[(size_t)(1 + e) * c->Npad]);
    psb_fill_welch_cols(y, N, c->Npad, cols.data(), c->col_w0, &c->colmask_lo, &c->colmask_hi);
    rc0 = psb_upload_welch_T(c, &cols[(size_t)c->col_w0 * c->Npad], &cols[(size_t)(c->col_w0 + 1) * c->Npad]);
    if (rc0) return rc0;
    PSB_CUDA(cudaMalloc(&c->d_cols, cols.size() * sizeof(double)));
    PSB_CUDA(cudaMemcpy(c->d_cols, cols.data(), cols.size() * sizeof(double), cudaMemcpyHostToDevice));
    int rc = psb_upload_pheno(c, y);
    if (rc) return rc;

    // L = P U Sd^-1/2, zero padded to [Lrows][Jpad]
    c->Lrows = ((N + QF_BK - 1) / QF_BK) * QF_BK;
    c->Jpad = ((J + QF_BN - 1) / QF_BN) * QF_BN;
    {
        std::vector<double> isd(J);
        for (int j = 0; j < J; ++j) isd[j] = 1.0 / sqrt(Sd[j]);
        for (int i = 0; i < N; ++i) {
            double *pi = &PU[(size_t)i * J];
            for (int j = 0; j < J; ++j) pi[j] *= isd[j];
        }
    }
    PSB_CUDA(cudaMalloc(&c->d_L, (size_t)c->Lrows * c->Jpad * sizeof(double)));
    PSB_CUDA(cudaMemset(c->d_L, 0, (size_t)c->Lrows * c->Jpad * sizeof(double)));
    PSB_CUDA(cudaMemcpy2D(c->d_L, (size_t)c->Jpad * sizeof(double), PU.data(),
                          (size_t)J * sizeof(double), (size_t)J * sizeof(double), N,
                          cudaMemcpyHostToDevice));
    PSB_UPLOAD_FENCE();
    c->model = PSB_MODEL_LMM;
    if (precision > 0) {
        // precision 46: two operand images -- 6 slices for the variants whose F statistic exceeds
        // refine_F (built first and stashed), 4 slices for everyone (the working image)
        const int passes = precision == 46 ? 2 : 1;
        for (int pass = 0; pass < passes; ++pass) {
            if (precision == 46) c->precision = pass == 0 ? 6 : 4;
            rc = psb_lmm_tc_setup(c, cols.data(), cols.data() + c->Npad, r, c->Npad,
                                  &cols[(size_t)c->col_w0 * c->Npad], &cols[(size_t)(c->col_w0 + 1) * c->Npad]);
            if (rc) {
                psb_free_model(c);
                return rc;
            }
            if (passes == 2 && pass == 0) {
                if (c->tc_special == 0) break;       // no x'v from the tensor pass: single pass at 6 slices
                psb_lmm_tc_stash(c);
            }
        }
        if (getenv("PSB_REFINE_F")) c->refine_F = atof(getenv("PSB_REFINE_F"));
    }
    return PSB_OK;
}

// Two-pass mode: which tested variants go through the contraction again at the refinement precision.
// F = beta^2 / var(beta) = (J - 1) b^2 / (a YKY - b^2)  (lmm.py:248, lmm_cov.py:799-815).  The relative
// error of the p-value is about F / 2 times that of a (1e-8 at 4 slices), so everything above
// refine_F is redone (and everything that is not finite).
__global__ void __launch_bounds__(256)
k_lmm_select_refine(const int *__restrict__ n_tested_dev, const int32_t *__restrict__ idx,
                    const double *__restrict__ a_in, const double *__restrict__ b_in, double YKY, double dof1,
                    double thr, int32_t *__restrict__ list, int *__restrict__ n_list) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *n_tested_dev) return;
    const int v = idx[t];
    const double a = a_in[v], b = b_in[v];
    const double F = dof1 * b * b / (a * YKY - b * b);
    if (!(F <= thr)) list[atomicAdd(n_list, 1)] = v;
}

extern "C" int psb_run_lmm(psb_ctx *c, const psb_params *prm) {
    PSB_NVTX("psb_run_lmm");
    PSB_REQUIRE(c && prm, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(c->model == PSB_MODEL_LMM, PSB_ERR_STATE, "psb_run_lmm without psb_lmm_setup");
    PSB_CUDA(cudaSetDevice(c->device));
    int rc = psb_run_begin(c);
    if (rc) return rc;
    PSB_REQUIRE(c->d_bits || c->S == 0, PSB_ERR_STATE, "psb_run_lmm without psb_submit");
    rc = psb_ensure_capacity(c, c->S, 0);
    if (rc) return rc;
    rc = psb_table_flip(c);
    if (rc) return rc;
    PSB_CUDA(cudaEventRecord(c->ev_run0, c->stream));
    // With the tensor path carrying x'v and Q'x, the stats pass only needs popcounts and
    // (continuous phenotype) the two Welch sums; otherwise all masked column sums.
    const bool tc_sums = c->precision > 0 && c->tc_special > 0;
    // Continuous phenotype with a pre-filter that does not gate (pyseer's default --filter-pvalue
    // 1): the Welch sums ride the tensor pass (two more column pairs in its special tile) and the
    // test moves into the epilogue, so the stats pass is popcounts only (HBM speed).  PSB_WELCH_TC=0
    // keeps the CUDA-core sums.
    static const bool welch_tc_ok = !(getenv("PSB_WELCH_TC") && atoi(getenv("PSB_WELCH_TC")) == 0);
    const bool defer = welch_tc_ok && tc_sums && c->tc_welch && !c->d_miss && prm->continuous &&
                       prm->filter_pvalue >= 1.0 && !(prm->options & PSB_OPT_NO_PREFILTER);
    c->tc_welch_run = defer;
    if (tc_sums && !c->d_miss && (defer || psb_bitstats_fits(c)))
        rc = psb_launch_bitstats(c, defer ? 0 : prm->continuous);
    else
        rc = psb_launch_bitsums(c);
    if (rc) return rc;
    rc = psb_launch_prefilter(c, prm, /*lmm_rule=*/1, defer ? 1 : 0);
    if (rc) return rc;
    // The number of tested variants stays on the device (counters[0]): the whole run is queued
    // without a host round trip, so the next psb_submit copy overlaps these kernels.
    const int upper = (int)c->S;
    PSB_CUDA(cudaEventRecord(c->ev_k0, c->stream));
    if (upper > 0) {
        if (c->precision == 0) {
            int tiles = psb_div_up(upper, QF_BM);
            int grid = std::min(tiles, c->sm_count);
            k_lmm_quadform_fp64<<<grid, QF_THREADS, 0, c->stream>>>(
                c->d_bits, c->Wrow, c->d_idx, c->d_counters, c->d_L, c->Lrows, c->Jpad, c->d_a);
            c->launches++;
            PSB_CUDA(cudaGetLastError());
        } else {
            rc = psb_lmm_tc_run(c, upper);
            if (rc) return rc;
            if (c->tc_alt && tc_sums) {
                // second pass at the refinement precision over the (few) variants in the far tail;
                // the list is in no particular order, every variant writes its own slots
                k_lmm_select_refine<<<psb_div_up(upper, 256), 256, 0, c->stream>>>(
                    c->d_counters, c->d_idx, c->d_a, c->d_b, c->YKY, (double)(c->J - 1), c->refine_F,
                    c->d_idx3, c->d_counters + 6);
                c->launches++;
                PSB_CUDA(cudaGetLastError());
                psb_lmm_tc_swap(c);
                rc = psb_lmm_tc_run_list(c, upper, c->d_idx3, c->d_counters + 6);
                psb_lmm_tc_swap(c);
                if (rc) return rc;
            }
        }
    }
    PSB_CUDA(cudaEventRecord(c->ev_k1, c->stream));
    c->have_k_ev = true;
    if (upper > 0) {
        k_lmm_epilogue<<<psb_div_up(upper, 256), 256, 0, c->stream>>>(
            c->d_counters, c->d_idx, c->d_a, tc_sums ? c->d_b : nullptr, tc_sums ? c->d_pp : nullptr,
            c->d_sums, c->C, c->col_b, c->col_q0, c->col_w0 - c->col_q0,
            c->N, c->d_carriers, c->d_missing, c->YKY, (double)(c->J - 1), prm->lrt_pvalue,
            c->d_pvalue, c->d_beta, c->d_bse, c->d_extra, c->d_flags, c->d_counters,
            defer ? 1 : 0, c->col_w0, c->welch_T1, c->welch_T2, prm->filter_pvalue, c->d_prep);
        c->launches++;
        PSB_CUDA(cudaGetLastError());
    }
    PSB_CUDA(cudaEventRecord(c->ev_run1, c->stream));
    c->have_run_ev = true;
    c->ran = true;
    return psb_run_end(c);
}
