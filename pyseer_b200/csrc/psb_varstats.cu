// psb_varstats.cu -- per-variant popcounts, 2x2 table and masked column sums.
//
// Replaces the tail of input.read_variant (input.py:439-452: kstrains count, af, missing),
// the table building of model.pre_filtering (model.py:57-61) and every O(N) masked
// reduction the models need (x'v for the LMM score, Z'x / y'x for OLS, Welch sums).
//
// One warp per variant.  The packed row is read once, coalesced (lane l owns word l of
// each 32-word chunk); for the fp64 column sums the chunk's words are broadcast by
// shuffle and lane l takes bit l of word t, so the column read col[32 t + l] is a
// coalesced 256-byte line that stays in L1/L2 (the columns are shared by all variants).
#include <algorithm>
#include <vector>

#include "psb_internal.cuh"

#define BITSUMS_MAXC 40

template <int HAS_MISS>
__global__ void __launch_bounds__(256)
k_bitsums(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ miss, int64_t S,
          int Wrow, int Wn, const uint32_t *__restrict__ y1, const uint32_t *__restrict__ y0,
          const uint32_t *__restrict__ valid, const double *__restrict__ cols,
          uint64_t mask_lo, uint64_t mask_hi, int C, int Npad, int32_t *__restrict__ carriers,
          int32_t *__restrict__ nmissing, int32_t *__restrict__ tab, double *__restrict__ sums) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (; v < S; v += warps_total) {
        const uint32_t *row = bits + v * Wrow;
        const uint32_t *mrow = HAS_MISS ? miss + v * Wrow : nullptr;
        int c_all = 0, c_miss = 0, n11 = 0, n10 = 0, n01 = 0, n00 = 0;
        double acc[BITSUMS_MAXC];
#pragma unroll
        for (int c = 0; c < BITSUMS_MAXC; ++c) acc[c] = 0.0;
        for (int w0 = 0; w0 < Wn; w0 += 32) {
            int w = w0 + lane;
            uint32_t x = 0, m = 0, vb = 0, a1 = 0, a0 = 0;
            if (w < Wn) {
                x = __ldg(row + w);
                if (HAS_MISS) m = __ldg(mrow + w);
                vb = __ldg(valid + w);
                a1 = __ldg(y1 + w);
                a0 = __ldg(y0 + w);
            }
            m &= vb;
            x &= vb & ~m;          // a NaN genotype is not a carrier bit
            uint32_t k1 = x & ~m, k0 = ~x & ~m & vb;
            c_all += __popc(x | m);
            c_miss += __popc(m);
            n11 += __popc(a1 & k1);
            n10 += __popc(a1 & k0);
            n01 += __popc(a0 & k1);
            n00 += __popc(a0 & k0);
            if (C > 0) {
                int tmax = min(32, Wn - w0);
                for (int t = 0; t < tmax; ++t) {
                    uint32_t xt = __shfl_sync(0xffffffffu, x, t);
                    uint32_t k1t = __shfl_sync(0xffffffffu, k1, t);
                    uint32_t k0t = __shfl_sync(0xffffffffu, k0, t);
                    bool bx = (xt >> lane) & 1u, b1 = (k1t >> lane) & 1u, b0 = (k0t >> lane) & 1u;
                    const double *cp = cols + (size_t)(w0 + t) * 32 + lane;
#pragma unroll
                    for (int c = 0; c < BITSUMS_MAXC; ++c) {
                        if (c < C) {
                            int mk = c < 32 ? (int)((mask_lo >> (2 * c)) & 3u) : (int)((mask_hi >> (2 * (c - 32))) & 3u);
                            bool on = mk == PSB_MASK_RAW ? bx : (mk == PSB_MASK_K1 ? b1 : b0);
                            double val = __ldg(cp + (size_t)c * Npad);
                            if (on) acc[c] += val;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c_all += __shfl_xor_sync(0xffffffffu, c_all, o);
            c_miss += __shfl_xor_sync(0xffffffffu, c_miss, o);
            n11 += __shfl_xor_sync(0xffffffffu, n11, o);
            n10 += __shfl_xor_sync(0xffffffffu, n10, o);
            n01 += __shfl_xor_sync(0xffffffffu, n01, o);
            n00 += __shfl_xor_sync(0xffffffffu, n00, o);
        }
#pragma unroll
        for (int c = 0; c < BITSUMS_MAXC; ++c) {
            if (c < C) {
                double a = acc[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) sums[v * C + c] = a;
            }
        }
        if (lane == 0) {
            carriers[v] = c_all;
            nmissing[v] = c_miss;
            tab[v * 4 + 0] = n11;
            tab[v * 4 + 1] = n10;
            tab[v * 4 + 2] = n01;
            tab[v * 4 + 3] = n00;
        }
    }
}

int psb_launch_bitsums(psb_ctx *c) {
    uint64_t mask_lo = c->colmask_lo, mask_hi = c->colmask_hi;
    PSB_REQUIRE(c->C <= BITSUMS_MAXC, PSB_ERR_UNSUPPORTED,
                "too many covariate columns for the device path (%d > %d)", c->C, BITSUMS_MAXC);
    if (c->S == 0) return PSB_OK;
    int warps_per_block = 8;
    int64_t blocks = (c->S + warps_per_block - 1) / warps_per_block;
    int64_t maxb = (int64_t)c->sm_count * 16;
    if (blocks > maxb) blocks = maxb;
    if (c->d_miss)
        k_bitsums<1><<<(int)blocks, 256, 0, c->stream>>>(
            c->d_bits, c->d_miss, c->S, c->Wrow, c->Wn, c->d_y1bits, c->d_y0bits, c->d_valid,
            c->d_cols, mask_lo, mask_hi, c->C, c->Npad, c->d_carriers, c->d_missing, c->d_tab, c->d_sums);
    else
        k_bitsums<0><<<(int)blocks, 256, 0, c->stream>>>(
            c->d_bits, nullptr, c->S, c->Wrow, c->Wn, c->d_y1bits, c->d_y0bits, c->d_valid,
            c->d_cols, mask_lo, mask_hi, c->C, c->Npad, c->d_carriers, c->d_missing, c->d_tab, c->d_sums);
    c->launches++;
    PSB_CUDA(cudaGetLastError());
    return PSB_OK;
}

// ---------------------------------------------------------------------------------
// AF filter + pre_filtering (input.py:608 / :693, model.py:31-70) + compaction of the
// variants that go on to a model fit.  One thread per variant.
// ---------------------------------------------------------------------------------
#include "psb_math.cuh"

__global__ void __launch_bounds__(256)
k_prefilter(int64_t S, int N, const int32_t *__restrict__ carriers,
            const int32_t *__restrict__ nmissing, const int32_t *__restrict__ tab,
            const double *__restrict__ sums, int C, int col_w0, psb_params prm, int lmm_rule,
            int defer_welch, double *__restrict__ af_out, double *__restrict__ prep_out,
            double *__restrict__ pvalue, double *__restrict__ beta, double *__restrict__ bse,
            double *__restrict__ extra, uint32_t *__restrict__ flags, int32_t *__restrict__ idx,
            int *__restrict__ counters) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= S) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    int carr = carriers[v], nm = nmissing[v];
    double af = (double)carr / (double)N;           // input.py:446
    double missing = (double)nm / (double)N;        // input.py:452
    uint32_t f = 0;
    double prep = nan;
    if (prm.options & PSB_OPT_NO_PREFILTER) {
        f = PSB_F_TESTED;
        int pos = atomicAdd(&counters[0], 1);
        idx[pos] = (int32_t)v;
    } else if (!(prm.min_af <= af && af <= prm.max_af) || missing > prm.max_missing) {
        f = PSB_F_AF_FILTER | PSB_F_PREFILTER;
    } else {
        if (prm.continuous && defer_welch) {
            // the Welch sums come out of the tensor pass (special tile) and the test is applied in
            // k_lmm_epilogue: every variant that passes the AF filter goes on
            prep = nan;
        } else if (prm.continuous) {
            // Welch t-test, scipy.stats.ttest_ind(p[k==1], p[k==0], equal_var=False)
            const double *s = sums + v * C + col_w0;
            prep = psb_welch_prep(s[0], s[1], s[2], s[3], (double)(carr - nm), (double)(N - carr));
        } else {
            int n11 = tab[v * 4 + 0], n10 = tab[v * 4 + 1], n01 = tab[v * 4 + 2], n00 = tab[v * 4 + 3];
            int le1 = (n11 <= 1) + (n10 <= 1) + (n01 <= 1) + (n00 <= 1);
            int le5 = (n11 <= 5) + (n10 <= 5) + (n01 <= 5) + (n00 <= 5);
            if (le1 > 0 || le5 > 1) f |= PSB_F_BAD_CHISQ;      // model.py:65-66
            double r1 = (double)(n11 + n10), r0 = (double)(n01 + n00);
            double c1 = (double)(n11 + n01), c0 = (double)(n10 + n00);
            double n = r1 + r0;
            double e11 = r1 * c1 / n, e10 = r1 * c0 / n, e01 = r0 * c1 / n, e00 = r0 * c0 / n;
            if (n > 0.0 && (e11 == 0.0 || e10 == 0.0 || e01 == 0.0 || e00 == 0.0)) {
                prep = nan;   // scipy raises ValueError here; reported as a failed pre-filter
            } else {
                double d11 = n11 - e11, d10 = n10 - e10, d01 = n01 - e01, d00 = n00 - e00;
                double chi2 = d11 * d11 / e11 + d10 * d10 / e10 + d01 * d01 / e01 + d00 * d00 / e00;
                prep = psb_chi2_sf1(chi2);
            }
        }
        bool fail = lmm_rule ? (prep >= prm.filter_pvalue) : (prep > prm.filter_pvalue);
        const bool deferred = prm.continuous && defer_welch;
        if (!deferred && (fail || !isfinite(prep))) {
            f |= PSB_F_PREFILTER_FAILED | PSB_F_PREFILTER;
        } else {
            f |= PSB_F_TESTED;
            int pos = atomicAdd(&counters[0], 1);
            idx[pos] = (int32_t)v;
        }
    }
    if (f & PSB_F_PREFILTER) atomicAdd(&counters[1], 1);
    af_out[v] = af;
    prep_out[v] = prep;
    pvalue[v] = nan;
    beta[v] = nan;
    bse[v] = nan;
    extra[v] = nan;
    flags[v] = f;
}

int psb_launch_prefilter(psb_ctx *c, const psb_params *prm, int lmm_rule, int defer_welch) {
    PSB_CUDA(cudaMemsetAsync(c->d_counters, 0, PSB_N_COUNTERS * sizeof(int), c->stream));
    if (c->S == 0) return PSB_OK;
    int blocks = psb_div_up(c->S, 256);
    k_prefilter<<<blocks, 256, 0, c->stream>>>(c->S, c->N, c->d_carriers, c->d_missing, c->d_tab,
                                               c->d_sums, c->C, c->col_w0, *prm, lmm_rule, defer_welch,
                                               c->d_af,
                                               c->d_prep, c->d_pvalue, c->d_beta, c->d_bse,
                                               c->d_extra, c->d_flags, c->d_idx, c->d_counters);
    c->launches++;
    PSB_CUDA(cudaGetLastError());
    return PSB_OK;
}

// ---------------------------------------------------------------------------------
// Fast stats pass (no missing genotypes): carriers, 2x2 table and -- continuous phenotype
// only -- the Welch sums over carriers.  Used when the tensor pass carries x'v and Q'x, so
// the only fp64 columns left are (yc, yc^2).  The two columns live in shared memory,
// transposed to [bit][word] so that lane l (word l of a 32-word chunk) reads consecutive
// 16-byte cells; a warp walks 4 variants at a time so every cell read serves 4 rows.
// Non-carrier sums follow from the totals (T - sum over carriers).
// ---------------------------------------------------------------------------------
#define BST_VPW 8

template <int NC>
__global__ void __launch_bounds__(256)
k_bitstats(const uint32_t *__restrict__ bits, int64_t S, int Wrow, int Wn,
           const uint32_t *__restrict__ y1, const uint32_t *__restrict__ y0,
           const uint32_t *__restrict__ valid, const double2 *__restrict__ colsT, double T1, double T2,
           int n_y1, int n_y0, int C, int col_w0, int32_t *__restrict__ carriers,
           int32_t *__restrict__ nmissing, int32_t *__restrict__ tab, double *__restrict__ sums) {
    extern __shared__ __align__(16) unsigned char bst_smem[];
    double2 *sT = reinterpret_cast<double2 *>(bst_smem);
    if (NC) {
        for (int e = threadIdx.x; e < 32 * Wn; e += blockDim.x) sT[e] = colsT[e];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * BST_VPW;
    for (; base < S; base += warps_total * BST_VPW) {
        int c_all[BST_VPW], n11[BST_VPW], n01[BST_VPW];
        double s1[BST_VPW], s2[BST_VPW];
#pragma unroll
        for (int u = 0; u < BST_VPW; ++u) {
            c_all[u] = n11[u] = n01[u] = 0;
            s1[u] = s2[u] = 0.0;
        }
        for (int w = lane; w < Wn; w += 32) {
            const uint32_t vb = __ldg(valid + w), a1 = __ldg(y1 + w), a0 = __ldg(y0 + w);
            uint32_t x[BST_VPW];
#pragma unroll
            for (int u = 0; u < BST_VPW; ++u) {
                x[u] = (base + u < S) ? (__ldg(bits + (base + u) * Wrow + w) & vb) : 0u;
                c_all[u] += __popc(x[u]);
                n11[u] += __popc(x[u] & a1);
                n01[u] += __popc(x[u] & a0);
            }
            if (NC) {
                const double2 *cp = sT + w;
#pragma unroll 8
                for (int b = 0; b < 32; ++b) {
                    const double2 yv = cp[b * Wn];
#pragma unroll
                    for (int u = 0; u < BST_VPW; ++u) {
                        if ((x[u] >> b) & 1u) {
                            s1[u] += yv.x;
                            s2[u] += yv.y;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < BST_VPW; ++u) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c_all[u] += __shfl_xor_sync(0xffffffffu, c_all[u], o);
                n11[u] += __shfl_xor_sync(0xffffffffu, n11[u], o);
                n01[u] += __shfl_xor_sync(0xffffffffu, n01[u], o);
                if (NC) {
                    s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], o);
                    s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], o);
                }
            }
        }
        if (lane < BST_VPW && base + lane < S) {
            // lane u publishes variant base + u (static indexing keeps the arrays in registers)
            int ca = 0, a11 = 0, a01 = 0;
            double r1 = 0.0, r2 = 0.0;
#pragma unroll
            for (int u = 0; u < BST_VPW; ++u)
                if (lane == u) {
                    ca = c_all[u]; a11 = n11[u]; a01 = n01[u]; r1 = s1[u]; r2 = s2[u];
                }
            const int64_t v = base + lane;
            carriers[v] = ca;
            nmissing[v] = 0;
            tab[v * 4 + 0] = a11;
            tab[v * 4 + 1] = n_y1 - a11;
            tab[v * 4 + 2] = a01;
            tab[v * 4 + 3] = n_y0 - a01;
            if (NC) {
                double *s = sums + v * C + col_w0;
                s[0] = r1;
                s[1] = r2;
                s[2] = T1 - r1;
                s[3] = T2 - r2;
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// Popcount-only stats pass at HBM speed (no missing genotypes, no fp64 columns): carriers and
// the 2x2 table.  G lanes share a row and read it in 16-byte cells (G = 32 / 16 / 8 by row
// width, so short rows still fill the warp: 32 / G rows side by side); BSV_U row groups are in
// flight per warp, i.e. every lane has up to 16 independent 16-byte loads outstanding.  The three
// counts of a row travel as one packed 64-bit word (21 bits each) through a transposing butterfly:
// each exchange halves the number of live values, 2.25 shuffles per row instead of 15.
// The phenotype masks sit in shared memory, zero padded to the row width.
// ---------------------------------------------------------------------------------
#define BSV_U 8

template <int G>
__global__ void __launch_bounds__(256)
k_bitstats_v(const uint32_t *__restrict__ bits, int64_t S, int Wrow, int Wn,
             const uint32_t *__restrict__ y1, const uint32_t *__restrict__ y0,
             const uint32_t *__restrict__ valid, int n_y1, int n_y0,
             int32_t *__restrict__ carriers, int32_t *__restrict__ nmissing, int32_t *__restrict__ tab) {
    extern __shared__ __align__(16) unsigned char bsv_smem[];
    uint4 *sV = reinterpret_cast<uint4 *>(bsv_smem);          // [Wq] valid, then y1, then y0
    const int Wq = Wrow >> 2;
    {
        uint32_t *w = reinterpret_cast<uint32_t *>(bsv_smem);
        for (int e = threadIdx.x; e < Wrow; e += blockDim.x) {
            const bool in = e < Wn;
            w[e] = in ? valid[e] : 0u;
            w[Wrow + e] = in ? y1[e] : 0u;
            w[2 * Wrow + e] = in ? y0[e] : 0u;
        }
        __syncthreads();
    }
    const uint4 *s1 = sV + Wq, *s0 = sV + 2 * Wq;
    constexpr int RPG = 32 / G;                  // rows side by side in a warp
    constexpr int RPP = RPG * BSV_U;             // rows per warp pass
    const int lane = threadIdx.x & 31, sg = lane / G, ql = lane % G;
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPP;
    const uint4 *rows = reinterpret_cast<const uint4 *>(bits);
    for (; base < S; base += warps_total * RPP) {
        int c_all[BSV_U], n11[BSV_U], n01[BSV_U];
#pragma unroll
        for (int u = 0; u < BSV_U; ++u) c_all[u] = n11[u] = n01[u] = 0;
        for (int q = ql; q < Wq; q += G) {
            uint4 x[BSV_U];
#pragma unroll
            for (int u = 0; u < BSV_U; ++u) {
                const int64_t r = base + u * RPG + sg;
                x[u] = r < S ? __ldcs(rows + r * Wq + q) : make_uint4(0u, 0u, 0u, 0u);
            }
            const uint4 vb = sV[q], a1 = s1[q], a0 = s0[q];
#pragma unroll
            for (int u = 0; u < BSV_U; ++u) {
                const uint32_t xa = x[u].x & vb.x, xb = x[u].y & vb.y, xc = x[u].z & vb.z, xd = x[u].w & vb.w;
                c_all[u] += __popc(xa) + __popc(xb) + __popc(xc) + __popc(xd);
                n11[u] += __popc(xa & a1.x) + __popc(xb & a1.y) + __popc(xc & a1.z) + __popc(xd & a1.w);
                n01[u] += __popc(xa & a0.x) + __popc(xb & a0.y) + __popc(xc & a0.z) + __popc(xd & a0.w);
            }
        }
        unsigned long long v[BSV_U];
#pragma unroll
        for (int u = 0; u < BSV_U; ++u)
            v[u] = (unsigned long long)c_all[u] | ((unsigned long long)n11[u] << 21) |
                   ((unsigned long long)n01[u] << 42);
        // transposing butterfly inside the G lanes of a row group: offsets G/2 .. 1; while more than
        // one value is live a lane keeps one half and hands the other half over
        int usel = 0;
#pragma unroll
        for (int o = G / 2, n = BSV_U; o > 0; o >>= 1) {
            if (n > 1) {
                const int half = n / 2;
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int i = 0; i < half; ++i) {
                    const unsigned long long send = up ? v[i] : v[i + half];
                    const unsigned long long keep = up ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
                usel = usel * 2 + (up ? 1 : 0);
                n = half;
            } else {
                v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
            }
        }
        // G >= BSV_U, so one value is left: row group `usel`, complete in the lanes whose low
        // log2(G / BSV_U) bits are zero
        const int64_t r = base + usel * RPG + sg;
        if ((ql & (G / BSV_U - 1)) == 0 && r < S) {
            const int ca = (int)(v[0] & 0x1fffffu), a11 = (int)((v[0] >> 21) & 0x1fffffu),
                      a01 = (int)((v[0] >> 42) & 0x1fffffu);
            carriers[r] = ca;
            nmissing[r] = 0;
            *reinterpret_cast<int4 *>(tab + r * 4) = make_int4(a11, n_y1 - a11, a01, n_y0 - a01);
        }
    }
}

// ---------------------------------------------------------------------------------
// The same pass as a streaming kernel: tiles of whole rows come in through the bulk-copy engine
// (cp.async.bulk on an mbarrier, two tiles per CTA: one lands while the other is counted); four
// threads share a row of the tile and walk it in 16-byte cells of shared memory, so the loads in
// flight do not depend on how many warps are resident and the only cross-lane traffic is two
// shuffle steps per row.  Thread (row r, part p) visits cells (4 i + p + rot_r) mod Wq with rot_r
// chosen so that the 32 lanes of a warp (8 rows x 4 parts) spread evenly over the eight 16-byte bank
// groups for any row width.  MODE 0: no 0/1 phenotype classes (continuous phenotype: carriers only,
// one popcount per word); 1: every valid sample is a case or a control (carriers = n11 + n01, two
// popcounts); 2: general (three).
// ---------------------------------------------------------------------------------
#define BSS_THREADS 512

__device__ __forceinline__ uint32_t bss_smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void bss_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(BSS_THREADS)
k_bitstats_stream(const uint32_t *__restrict__ bits, int64_t S, int Wrow, int Wn, int tile_rows,
                  const uint32_t *__restrict__ y1, const uint32_t *__restrict__ y0,
                  const uint32_t *__restrict__ valid, int n_y1, int n_y0,
                  int32_t *__restrict__ carriers, int32_t *__restrict__ nmissing, int32_t *__restrict__ tab) {
    extern __shared__ __align__(128) unsigned char bss_smem[];
    const int Wq = Wrow >> 2;
    const size_t tile_cells = (size_t)tile_rows * Wq;
    uint4 *sTile = reinterpret_cast<uint4 *>(bss_smem);                          // two tiles
    uint4 *sV = sTile + 2 * tile_cells;                                          // valid | y1 | y0
    uint64_t *mb = reinterpret_cast<uint64_t *>(sV + 3 * (size_t)Wq);
    {
        uint32_t *w = reinterpret_cast<uint32_t *>(sV);
        for (int e = threadIdx.x; e < Wrow; e += blockDim.x) {
            const bool in = e < Wn;
            w[e] = in ? valid[e] : 0u;
            w[Wrow + e] = in ? y1[e] : 0u;
            w[2 * Wrow + e] = in ? y0[e] : 0u;
        }
        if (threadIdx.x == 0) {
            for (int b = 0; b < 2; ++b)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bss_smem_u32(&mb[b])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    const uint4 *s1 = sV + Wq, *s0 = sV + 2 * Wq;
    const int64_t n_tiles = (S + tile_rows - 1) / tile_rows;
    auto post = [&](int64_t tile, int buf) {          // one thread: bulk copy of a tile
        const int64_t r0 = tile * tile_rows;
        const uint32_t bytes = (uint32_t)(min((int64_t)tile_rows, S - r0) * Wrow * 4);
        const uint32_t bar = bss_smem_u32(&mb[buf]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         bss_smem_u32(sTile + (size_t)buf * tile_cells)),
                     "l"(bits + (size_t)r0 * Wrow), "r"(bytes), "r"(bar)
                     : "memory");
    };
    int64_t tile = blockIdx.x;
    if (threadIdx.x == 0) {
        if (tile < n_tiles) post(tile, 0);
        if (tile + gridDim.x < n_tiles) post(tile + gridDim.x, 1);
    }
    const int part = threadIdx.x & 3;
    uint32_t phase = 0;
    for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        bss_wait(bss_smem_u32(&mb[buf]), (phase >> buf) & 1u);
        phase ^= 1u << buf;
        const int64_t r0 = tile * tile_rows;
        const int rows = (int)min((int64_t)tile_rows, S - r0);
        const uint4 *tbase = sTile + (size_t)buf * tile_cells;
        for (int rb = 0; rb < rows; rb += BSS_THREADS / 4) {
            const int r = rb + (threadIdx.x >> 2);
            const bool live = r < rows;
            const uint4 *row = tbase + (size_t)(live ? r : 0) * Wq;
            int c_all = 0, n11 = 0, n01 = 0;
            if (live) {
                // (r Wq + rot) mod 8 alternates between 0 and 4 from row to row
                int c = (part + ((4 * (r & 1) - r * Wq) & 7)) % Wq;
                for (int k = part; k < Wq; k += 4) {
                    const uint4 x = row[c];
                    const uint4 vb = sV[c];
                    const uint32_t xa = x.x & vb.x, xb = x.y & vb.y, xc = x.z & vb.z, xd = x.w & vb.w;
                    if (MODE != 1) c_all += __popc(xa) + __popc(xb) + __popc(xc) + __popc(xd);
                    if (MODE != 0) {
                        const uint4 a1 = s1[c], a0 = s0[c];
                        n11 += __popc(xa & a1.x) + __popc(xb & a1.y) + __popc(xc & a1.z) + __popc(xd & a1.w);
                        n01 += __popc(xa & a0.x) + __popc(xb & a0.y) + __popc(xc & a0.z) + __popc(xd & a0.w);
                    }
                    c += 4;
                    while (c >= Wq) c -= Wq;
                }
            }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                if (MODE != 1) c_all += __shfl_xor_sync(0xffffffffu, c_all, o);
                if (MODE != 0) {
                    n11 += __shfl_xor_sync(0xffffffffu, n11, o);
                    n01 += __shfl_xor_sync(0xffffffffu, n01, o);
                }
            }
            if (MODE == 1) c_all = n11 + n01;
            if (live && part == 0) {
                const int64_t v = r0 + r;
                carriers[v] = c_all;
                nmissing[v] = 0;
                *reinterpret_cast<int4 *>(tab + v * 4) = make_int4(n11, n_y1 - n11, n01, n_y0 - n01);
            }
        }
        __syncthreads();                              // every thread has left the tile: refill it
        if (threadIdx.x == 0 && tile + 2 * (int64_t)gridDim.x < n_tiles) post(tile + 2 * (int64_t)gridDim.x, buf);
    }
}

int psb_upload_welch_T(psb_ctx *c, const double *yc, const double *yc2) {
    const int Wn = c->Wn, N = c->N;
    std::vector<double> t((size_t)32 * Wn * 2, 0.0);
    double T1 = 0.0, T2 = 0.0;
    for (int i = 0; i < N; ++i) {
        int w = i >> 5, b = i & 31;
        t[((size_t)b * Wn + w) * 2 + 0] = yc[i];
        t[((size_t)b * Wn + w) * 2 + 1] = yc2[i];
        T1 += yc[i];
        T2 += yc2[i];
    }
    c->welch_T1 = T1;
    c->welch_T2 = T2;
    PSB_CUDA(cudaMalloc(&c->d_wcolsT, t.size() * sizeof(double)));
    PSB_CUDA(cudaMemcpy(c->d_wcolsT, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
    return PSB_OK;
}

static size_t bitstats_smem(const psb_ctx *c) { return (size_t)32 * c->Wn * sizeof(double2); }

bool psb_bitstats_fits(psb_ctx *c) {
    return c->d_wcolsT != nullptr && bitstats_smem(c) <= 200 * 1024;
}

int psb_launch_bitstats(psb_ctx *c, int continuous) {
    if (c->S == 0) return PSB_OK;
    const size_t smem = continuous ? bitstats_smem(c) : 0;
    int per_sm = continuous ? (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / (smem + 1024))) : 8;
    int64_t blocks = (c->S + 8 * BST_VPW - 1) / (8 * BST_VPW);
    int64_t maxb = (int64_t)c->sm_count * per_sm;
    if (blocks > maxb) blocks = maxb;
    const double2 *ct = reinterpret_cast<const double2 *>(c->d_wcolsT);
    if (continuous) {
        PSB_CUDA(cudaFuncSetAttribute(k_bitstats<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_bitstats<2><<<(int)blocks, 256, smem, c->stream>>>(
            c->d_bits, c->S, c->Wrow, c->Wn, c->d_y1bits, c->d_y0bits, c->d_valid, ct, c->welch_T1,
            c->welch_T2, c->n_y1, c->n_y0, c->C, c->col_w0, c->d_carriers, c->d_missing, c->d_tab, c->d_sums);
    } else if ((size_t)c->Wrow * 4 <= 16 * 1024 && !(getenv("PSB_BITSTATS_STREAM") && atoi(getenv("PSB_BITSTATS_STREAM")) == 0)) {
        // popcounts only, streamed through shared memory by the bulk-copy engine
        const size_t row_bytes = (size_t)c->Wrow * 4;
        int tile_rows = (int)std::min<size_t>(512, (80 * 1024) / row_bytes);
        tile_rows = std::max(tile_rows, 1);
        const size_t sm = 2 * (size_t)tile_rows * row_bytes + (size_t)c->Wrow * 12 + 16;
        const int64_t n_tiles = (c->S + tile_rows - 1) / tile_rows;
        const int nb = (int)std::min<int64_t>(n_tiles, c->sm_count);
        const int mode = (c->n_y1 == 0 && c->n_y0 == 0) ? 0 : (c->n_y1 + c->n_y0 == c->N ? 1 : 2);
#define BSS_LAUNCH(M)                                                                                          \
        PSB_CUDA(cudaFuncSetAttribute(k_bitstats_stream<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        k_bitstats_stream<M><<<nb, BSS_THREADS, sm, c->stream>>>(c->d_bits, c->S, c->Wrow, c->Wn, tile_rows,    \
                                                                c->d_y1bits, c->d_y0bits, c->d_valid, c->n_y1, \
                                                                c->n_y0, c->d_carriers, c->d_missing, c->d_tab)
        if (mode == 0) { BSS_LAUNCH(0); }
        else if (mode == 1) { BSS_LAUNCH(1); }
        else { BSS_LAUNCH(2); }
#undef BSS_LAUNCH
    } else if (c->N < (1 << 21) && (size_t)c->Wrow * 12 <= 48 * 1024 &&
               !(getenv("PSB_BITSTATS_V") && atoi(getenv("PSB_BITSTATS_V")) == 0)) {
        // popcounts only: 16-byte loads, G lanes per row
        const int Wq = c->Wrow >> 2;
        const size_t sm = (size_t)c->Wrow * 12;
        const int G = Wq <= 8 ? 8 : Wq <= 16 ? 16 : 32;
        const int64_t rpp = (int64_t)(32 / G) * BSV_U * 8;             // rows per block pass
        int64_t nb = std::min<int64_t>((c->S + rpp - 1) / rpp, (int64_t)c->sm_count * 4);
#define BSV_LAUNCH(GG)                                                                                  \
        k_bitstats_v<GG><<<(int)nb, 256, sm, c->stream>>>(c->d_bits, c->S, c->Wrow, c->Wn, c->d_y1bits, \
                                                          c->d_y0bits, c->d_valid, c->n_y1, c->n_y0,     \
                                                          c->d_carriers, c->d_missing, c->d_tab)
        if (G == 8) BSV_LAUNCH(8);
        else if (G == 16) BSV_LAUNCH(16);
        else BSV_LAUNCH(32);
#undef BSV_LAUNCH
    } else {
        k_bitstats<0><<<(int)blocks, 256, 0, c->stream>>>(
            c->d_bits, c->S, c->Wrow, c->Wn, c->d_y1bits, c->d_y0bits, c->d_valid, ct, c->welch_T1,
            c->welch_T2, c->n_y1, c->n_y0, c->C, c->col_w0, c->d_carriers, c->d_missing, c->d_tab, c->d_sums);
    }
    c->launches++;
    PSB_CUDA(cudaGetLastError());
    return PSB_OK;
}
