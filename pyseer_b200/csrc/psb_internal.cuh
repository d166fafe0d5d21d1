// psb_internal.cuh -- context layout and helpers shared by the translation units of
// libpyseer_b200.so.  Not part of the public ABI (see include/pyseer_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/pyseer_b200.h"

#define PSB_N_COUNTERS 16  // device counters per result set (see psb_varstats.cu: psb_launch_prefilter)
#define PSB_MODEL_NONE 0
#define PSB_MODEL_LMM 1
#define PSB_MODEL_FIXED 2

// mask kinds for k_bitsums columns
#define PSB_MASK_RAW 0   // x (NaN genotypes count as absent)
#define PSB_MASK_K1 1    // x & ~missing            (k == 1)
#define PSB_MASK_K0 2    // ~x & ~missing & valid   (k == 0)

void psb_set_error(const char *fmt, ...);

// NVTX ranges around the phases of the C ABI (visible in Nsight Systems / ncu --nvtx; header-only
// NVTX 3, no link dependency, a no-op without an attached tool)
#include <nvtx3/nvToolsExt.h>
struct psb_nvtx_range {
    explicit psb_nvtx_range(const char *name) { nvtxRangePushA(name); }
    ~psb_nvtx_range() { nvtxRangePop(); }
};
#define PSB_NVTX(name) psb_nvtx_range psb_nvtx_range_##__LINE__(name)

#define PSB_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e_ = (call);                                                    \
        if (e_ != cudaSuccess) {                                                    \
            psb_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__,            \
                          cudaGetErrorName(e_), cudaGetErrorString(e_));            \
            return PSB_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

// Setup uploads use plain cudaMemcpy / cudaMemset on the legacy default stream.  The library's
// streams are non-blocking (no implicit ordering with it) and a pageable copy may return before
// its DMA has landed, so every setup path fences once before a kernel can read the state.
#define PSB_UPLOAD_FENCE() PSB_CUDA(cudaDeviceSynchronize())


#define PSB_REQUIRE(cond, code, ...)                                                \
    do {                                                                            \
        if (!(cond)) {                                                              \
            psb_set_error(__VA_ARGS__);                                             \
            return (code);                                                          \
        }                                                                           \
    } while (0)

struct psb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_run0 = nullptr, ev_run1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    cudaEvent_t ev_user[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool have_run_ev = false, have_k_ev = false;
    int64_t launches = 0;

    int model = PSB_MODEL_NONE;
    int N = 0;    // samples
    int Wn = 0;   // ceil(N / 32)

    // phenotype helpers (both models)
    uint32_t *d_y1bits = nullptr, *d_y0bits = nullptr, *d_valid = nullptr;  // Wn words
    // column matrix for k_bitsums: [C][Npad] fp64
    double *d_cols = nullptr;
    uint64_t colmask_lo = 0, colmask_hi = 0;   // 2 bits per column: PSB_MASK_*
    int C = 0, Npad = 0;
    int col_b = -1, col_q0 = -1, col_w0 = -1;   // LMM: b column, first Q column; Welch cols
    double y_mean = 0.0;
    // fast stats pass (psb_varstats.cu k_bitstats): Welch columns transposed [32][Wn] x
    // (yc, yc^2), their totals, and the phenotype class counts
    double *d_wcolsT = nullptr;
    double welch_T1 = 0.0, welch_T2 = 0.0;
    int n_y1 = 0, n_y0 = 0;

    // ---- LMM ----
    int D = 0, J = 0, Jpad = 0, Kpad = 0;
    int precision = 0;
    double h2 = 0.0, YKY = 0.0;
    double *d_L = nullptr;        // [Npad32][Jpad64] fp64, L = P U Sd^-1/2
    int Lrows = 0;
    // tensor-core operands (psb_lmm_tc.cu)
    int8_t *d_Lq = nullptr;       // [jtiles][slices][32 comps][Kpad] int8 (K-major)
    double *d_scale2 = nullptr;   // [Jpad32] (s_j 2^-B)^2
    int n_slices = 0, jtiles = 0;
    bool tc_tri = false;          // regular tiles hold the triangular operand M'' (psb_lmm_tc.cu)
    uint8_t *d_shift = nullptr;   // [Jq] per-column left shift of the integer epilogue (triangular form)
    bool tc_int_epi = false;      // triangular tiles are recombined and summed in int64
    int tc_special = 0;           // hi/lo column pairs carried by the special tile (0 = none)
    bool tc_welch = false;        // the special tile also carries the Welch columns (yc, yc^2) after them
    bool tc_welch_run = false;    // ... and this run takes the Welch sums from there (psb_run_lmm)
    void *tmap_Lq = nullptr;      // host copy of the CUtensorMap (128 B): box = 128 samples x all sliced rows
    void *tmap_Lq_half = nullptr; // same with half of the sliced rows per box (two-SM mode)
    double *d_tc_part = nullptr;  // partial quadratic forms of a split launch (psb_lmm_tc.cu)
    void *tc_alt = nullptr;       // two-pass mode: the refinement precision's operand image (psb_lmm_tc.cu: TcState)
    double refine_F = 30.0;       // ... applied to variants whose F statistic exceeds this

    // ---- fixed effects ----
    int q = 0;
    int continuous = 0;
    double null_llf = 0.0, null_firth = 0.0;
    double *d_Z = nullptr;        // [q][Npad] (column c contiguous)
    double *d_yv = nullptr;       // [Npad]
    double *d_fixed_const = nullptr;  // OLS precomputed: see psb_fixed.cu
    std::vector<double> h_ZtZinv; // q x q
    std::vector<double> h_Zty;    // q
    double *d_Zlin = nullptr;     // lineage design [q_lin][Npad] (model.fit_lineage_effect)
    int q_lin = 0, n_lin = 0;
    bool logit_first_step = false; // closed-form first Newton step operands are set up
    // fast Logit path (psb_fixed_fast.cu): interleaved covariates (fp64, fp32), null weights, Z'W0Z
    double *d_fx_Zi = nullptr, *d_fx_W0 = nullptr, *d_fx_H0 = nullptr;
    float *d_fx_Zf = nullptr;
    uint32_t *d_fx_vb = nullptr;
    std::vector<double> fx_zmax;
    int fx_Q = 0;                  // 0: not set up
    int fixed_slow = 0;            // variants of the last run that took the exact kernel after the fast one
    std::vector<double> h_warm;   // null-model Logit parameters (warm start), q

    // ---- variants ----
    const uint32_t *d_bits = nullptr, *d_miss = nullptr;
    uint32_t *own_bits = nullptr, *own_miss = nullptr;       // psb_synth_device rows
    size_t own_bits_cap = 0, own_miss_cap = 0;   // bytes
    // psb_submit staging: two slots filled on a copy stream, so that the copy of batch i+1
    // overlaps the kernels of batch i (input.py's k-mer streaming as a pinned-host -> device
    // pipeline).  ev_copy[s]: rows of slot s have landed; ev_used[s]: last run that read slot s.
    cudaStream_t copy_stream = nullptr;
    uint32_t *stage_bits[2] = {nullptr, nullptr}, *stage_miss[2] = {nullptr, nullptr};
    size_t stage_bits_cap[2] = {0, 0}, stage_miss_cap[2] = {0, 0};
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_used[2] = {nullptr, nullptr};
    bool used_valid[2] = {false, false};
    // rows handed over by psb_submit* / psb_synth_device but not yet adopted by a run: a submit
    // does not disturb the table of the previous run, which can still be fetched
    const uint32_t *sub_bits = nullptr, *sub_miss = nullptr;
    int64_t sub_S = 0;
    int sub_Wrow = 0, sub_slot = -1;
    bool sub_valid = false;
    int stage_slot = 1;          // slot of the most recent psb_submit
    int bits_slot = -1;          // staging slot d_bits points into (-1: not a staging slot)
    bool copy_pending = false;   // the compute stream has not yet waited for ev_copy[bits_slot]
    int64_t S = 0;
    int Wrow = 0;

    // ---- result table + workspace (capacity in variants) ----
    int64_t cap = 0;
    int32_t *d_carriers = nullptr, *d_missing = nullptr;
    double *d_af = nullptr, *d_prep = nullptr, *d_pvalue = nullptr, *d_beta = nullptr,
           *d_bse = nullptr, *d_extra = nullptr, *d_betas = nullptr;
    uint32_t *d_flags = nullptr;
    int betas_cols = 0;
    int32_t *d_tab = nullptr;     // [cap][4] 2x2 table n11 n10 n01 n00
    double *d_sums = nullptr;     // [cap][C]
    size_t sums_cap = 0;
    int32_t *d_idx = nullptr;     // compacted tested variant ids (cap + 256)
    int32_t *d_idx2 = nullptr;    // second list (Firth candidates)
    int32_t *d_idx3 = nullptr;    // third list (variants the fast Logit kernel hands to the exact one)
    double *d_gen_scratch = nullptr;   // generic solver: per-warp scratch of the singular-matrix path
    size_t gen_scratch_cap = 0;
    int *d_counters = nullptr;    // [8] device counters
    int32_t *d_lineage = nullptr; // [cap] index of the strongest lineage, -1 = None
    double *d_a = nullptr;        // [cap] quadratic forms
    double *d_b = nullptr;        // [cap] x'v from the tensor path
    double *d_pp = nullptr;       // [cap] ||Q'x||^2 from the tensor path
    int64_t counts[4] = {0, 0, 0, 0};
    bool ran = false;
    // Second set of the columns psb_fetch returns (and of the counters): psb_run_lmm / psb_run_fixed
    // alternate between the two sets, so that the table of run i can travel to the host on its own
    // stream (psb_fetch_begin) while run i + 1 is computed.  d_* above always name the CURRENT set.
    int32_t *alt_carriers = nullptr, *alt_missing = nullptr;
    double *alt_af = nullptr, *alt_prep = nullptr, *alt_pvalue = nullptr, *alt_beta = nullptr,
           *alt_bse = nullptr, *alt_extra = nullptr, *alt_betas = nullptr;
    uint32_t *alt_flags = nullptr;
    int *alt_counters = nullptr;
    int tab_cur = 0;                       // which set d_* name
    cudaStream_t fetch_stream = nullptr;
    cudaEvent_t ev_fetch[2] = {nullptr, nullptr};   // psb_fetch_begin of set s has landed on the host
    bool fetch_valid[2] = {false, false};
    int fetch_set = -1;                    // set of the last psb_fetch_begin
    int64_t fetch_S[2] = {0, 0};
    int *h_counters[2] = {nullptr, nullptr};        // pinned: counters of the fetched run
    void *kin = nullptr;          // psb_kinship.cu state
    void *d_dig = nullptr;        // psb_patterns.cu: MD5 digests of the batch's rows
    int64_t dig_cap = 0;
    void *text = nullptr;         // psb_text.cu state (device k-mer text parser)

    // ---- burden regions (psb_burden.cu) ----
    uint32_t *bur_vbits = nullptr, *bur_vmiss = nullptr;   // member record rows (host submits)
    uint32_t *bur_out = nullptr, *bur_outmiss = nullptr;   // region rows
    int64_t *bur_offs = nullptr;
    int32_t *bur_members = nullptr;
    size_t bur_vbits_cap = 0, bur_vmiss_cap = 0, bur_out_cap = 0, bur_outmiss_cap = 0,
           bur_offs_cap = 0, bur_members_cap = 0;
    cudaEvent_t ev_bur_copy = nullptr, ev_bur_done = nullptr;
    bool bur_done_valid = false;
};

int psb_ensure_capacity(psb_ctx *ctx, int64_t S, int betas_cols);
int psb_run_begin(psb_ctx *ctx);
int psb_run_end(psb_ctx *ctx);
int psb_table_flip(psb_ctx *ctx);
int psb_free_model(psb_ctx *ctx);
void psb_kinship_release(psb_ctx *ctx);
void psb_burden_release(psb_ctx *ctx);
void psb_text_release(psb_ctx *ctx);
void psb_patterns_release(psb_ctx *ctx);

// psb_varstats.cu
int psb_launch_bitsums(psb_ctx *ctx);
int psb_launch_prefilter(psb_ctx *ctx, const psb_params *prm, int lmm_rule, int defer_welch);
int psb_upload_welch_T(psb_ctx *ctx, const double *yc, const double *yc2);
bool psb_bitstats_fits(psb_ctx *ctx);
int psb_launch_bitstats(psb_ctx *ctx, int continuous);

// psb_lmm_tc.cu
int psb_lmm_tc_setup(psb_ctx *ctx, const double *h_v, const double *h_Q, int r, int ldq,
                     const double *h_w1, const double *h_w2);
int psb_lmm_tc_run(psb_ctx *ctx, int n_tested);
int psb_lmm_tc_run_list(psb_ctx *ctx, int upper, const int32_t *list, const int *count_dev);
void psb_lmm_tc_swap(psb_ctx *ctx);
int psb_lmm_tc_stash(psb_ctx *ctx);
int psb_tc_linear_setup(psb_ctx *ctx, const double *cols, int ncols, int ld);
int psb_tc_run(psb_ctx *ctx, int n_tested, double *lin_out, int lin_ld);
void psb_lmm_tc_free(psb_ctx *ctx);

static inline int psb_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
