// psb_api.cu -- context lifecycle, staging, result table and measurement entry points
// of the C ABI declared in include/pyseer_b200.h.
#include <stdarg.h>
#include <string.h>

#include "psb_internal.cuh"
#include <stdlib.h>
#include <utility>
#include "psb_math.cuh"
#include "psb_fixed.cuh"

static thread_local char g_err[1024] = "";

void psb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

int psb_abi_version(void) { return PSB_ABI_VERSION; }
const char *psb_last_error(void) { return g_err; }

int psb_device_count(int *count) {
    PSB_REQUIRE(count, PSB_ERR_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return PSB_OK;
}

int psb_create(int device_id, psb_ctx **out) {
    PSB_REQUIRE(out, PSB_ERR_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    PSB_CUDA(cudaGetDeviceCount(&n));
    PSB_REQUIRE(device_id >= 0 && device_id < n, PSB_ERR_ARG,
                "device %d not available (%d visible); libpyseer_b200 has no CPU fallback",
                device_id, n);
    PSB_CUDA(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    PSB_CUDA(cudaGetDeviceProperties(&prop, device_id));
    PSB_REQUIRE(prop.major >= 10, PSB_ERR_UNSUPPORTED,
                "device %d is sm_%d%d; this library is built for sm_100a only", device_id,
                prop.major, prop.minor);
    psb_ctx *c = new psb_ctx();
    c->device = device_id;
    c->sm_count = prop.multiProcessorCount;
    PSB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PSB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        PSB_CUDA(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
        PSB_CUDA(cudaEventCreateWithFlags(&c->ev_used[i], cudaEventDisableTiming));
    }
    PSB_CUDA(cudaEventCreate(&c->ev_run0));
    PSB_CUDA(cudaEventCreate(&c->ev_run1));
    PSB_CUDA(cudaEventCreate(&c->ev_k0));
    PSB_CUDA(cudaEventCreate(&c->ev_k1));
    for (int i = 0; i < 8; ++i) PSB_CUDA(cudaEventCreate(&c->ev_user[i]));
    PSB_CUDA(cudaMalloc(&c->d_counters, PSB_N_COUNTERS * sizeof(int)));
    PSB_CUDA(cudaMalloc(&c->alt_counters, PSB_N_COUNTERS * sizeof(int)));
    PSB_CUDA(cudaMemset(c->d_counters, 0, PSB_N_COUNTERS * sizeof(int)));
    PSB_CUDA(cudaMemset(c->alt_counters, 0, PSB_N_COUNTERS * sizeof(int)));
    PSB_CUDA(cudaStreamCreateWithFlags(&c->fetch_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        PSB_CUDA(cudaEventCreateWithFlags(&c->ev_fetch[i], cudaEventDisableTiming));
        PSB_CUDA(cudaHostAlloc((void **)&c->h_counters[i], PSB_N_COUNTERS * sizeof(int), cudaHostAllocDefault));
    }
    *out = c;
    return PSB_OK;
}

}  // extern "C"

static void free_dev(void *p) {
    if (p) cudaFree(p);
}

int psb_free_model(psb_ctx *c) {
    free_dev(c->d_y1bits); c->d_y1bits = nullptr;
    free_dev(c->d_y0bits); c->d_y0bits = nullptr;
    free_dev(c->d_valid); c->d_valid = nullptr;
    free_dev(c->d_cols); c->d_cols = nullptr;
    free_dev(c->d_wcolsT); c->d_wcolsT = nullptr;
    free_dev(c->d_L); c->d_L = nullptr;
    psb_lmm_tc_free(c);
    free_dev(c->d_Z); c->d_Z = nullptr;
    free_dev(c->d_Zlin); c->d_Zlin = nullptr; c->q_lin = c->n_lin = 0;
    free_dev(c->d_yv); c->d_yv = nullptr;
    free_dev(c->d_fixed_const); c->d_fixed_const = nullptr;
    free_dev(c->d_sums); c->d_sums = nullptr; c->sums_cap = 0;
    c->h_warm.clear();
    c->logit_first_step = false;
    psb_fixed_fast_free(c);
    c->model = PSB_MODEL_NONE;
    c->ran = false;
    return PSB_OK;
}

static void free_tables(psb_ctx *c);
extern "C" {
}
static void free_tables(psb_ctx *c) {
    free_dev(c->d_carriers); free_dev(c->d_missing); free_dev(c->d_af); free_dev(c->d_prep);
    free_dev(c->d_pvalue); free_dev(c->d_beta); free_dev(c->d_bse); free_dev(c->d_extra);
    free_dev(c->d_betas); free_dev(c->d_flags); free_dev(c->d_tab); free_dev(c->d_idx);
    free_dev(c->d_idx2); free_dev(c->d_idx3); free_dev(c->d_a); free_dev(c->d_b); free_dev(c->d_pp); free_dev(c->d_lineage);
    free_dev(c->alt_carriers); free_dev(c->alt_missing); free_dev(c->alt_af); free_dev(c->alt_prep);
    free_dev(c->alt_pvalue); free_dev(c->alt_beta); free_dev(c->alt_bse); free_dev(c->alt_extra);
    free_dev(c->alt_betas); free_dev(c->alt_flags);
    c->alt_carriers = c->alt_missing = nullptr;
    c->alt_af = c->alt_prep = c->alt_pvalue = c->alt_beta = c->alt_bse = c->alt_extra = c->alt_betas = nullptr;
    c->alt_flags = nullptr;
    c->fetch_valid[0] = c->fetch_valid[1] = false;
    c->d_carriers = c->d_missing = nullptr;
    c->d_af = c->d_prep = c->d_pvalue = c->d_beta = c->d_bse = c->d_extra = c->d_betas = nullptr;
    c->d_flags = nullptr; c->d_tab = nullptr; c->d_idx = c->d_idx2 = c->d_idx3 = nullptr; c->d_a = c->d_b = c->d_pp = nullptr; c->d_lineage = nullptr;
    c->cap = 0; c->betas_cols = 0;
}

extern "C" {
int psb_destroy(psb_ctx *c) {
    if (!c) return PSB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamSynchronize(c->fetch_stream);
    psb_kinship_release(c);
    psb_burden_release(c);
    psb_text_release(c);
    psb_patterns_release(c);
    psb_free_model(c);
    free_tables(c);
    for (int i = 0; i < 2; ++i) {
        free_dev(c->stage_bits[i]); free_dev(c->stage_miss[i]);
        cudaEventDestroy(c->ev_copy[i]); cudaEventDestroy(c->ev_used[i]);
    }
    cudaStreamDestroy(c->copy_stream);
    free_dev(c->own_bits); free_dev(c->own_miss); free_dev(c->d_counters); free_dev(c->d_gen_scratch);
    free_dev(c->alt_counters);
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(c->ev_fetch[i]);
        if (c->h_counters[i]) cudaFreeHost(c->h_counters[i]);
    }
    cudaStreamDestroy(c->fetch_stream);
    cudaEventDestroy(c->ev_run0); cudaEventDestroy(c->ev_run1);
    cudaEventDestroy(c->ev_k0); cudaEventDestroy(c->ev_k1);
    for (int i = 0; i < 8; ++i) cudaEventDestroy(c->ev_user[i]);
    cudaStreamDestroy(c->stream);
    delete c;
    return PSB_OK;
}

int psb_sync(psb_ctx *c) {
    PSB_REQUIRE(c, PSB_ERR_ARG, "ctx is NULL");
    PSB_CUDA(cudaSetDevice(c->device));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    return PSB_OK;
}

}  // extern "C"

// Called by psb_run_*: order the compute stream after the pending psb_submit copy ...
int psb_run_begin(psb_ctx *c) {
    if (c->sub_valid) {
        // adopt the submitted rows as the current batch
        c->d_bits = c->sub_bits;
        c->d_miss = c->sub_miss;
        c->S = c->sub_S;
        c->Wrow = c->sub_Wrow;
        c->bits_slot = c->sub_slot;
        c->copy_pending = c->sub_slot >= 0;
        c->sub_valid = false;
        c->ran = false;
    }
    if (c->copy_pending && c->bits_slot >= 0) {
        PSB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy[c->bits_slot], 0));
        c->copy_pending = false;
    }
    return PSB_OK;
}
// ... and mark the staging slot as read once the run's kernels are queued.
int psb_run_end(psb_ctx *c) {
    if (c->bits_slot >= 0) {
        PSB_CUDA(cudaEventRecord(c->ev_used[c->bits_slot], c->stream));
        c->used_valid[c->bits_slot] = true;
    }
    return PSB_OK;
}

// psb_run_lmm / psb_run_fixed write the OTHER set of result columns than the run before them; a
// psb_fetch_begin that is still draining that set (two runs back) is waited for on the device.
int psb_table_flip(psb_ctx *c) {
    std::swap(c->d_carriers, c->alt_carriers);
    std::swap(c->d_missing, c->alt_missing);
    std::swap(c->d_af, c->alt_af);
    std::swap(c->d_prep, c->alt_prep);
    std::swap(c->d_pvalue, c->alt_pvalue);
    std::swap(c->d_beta, c->alt_beta);
    std::swap(c->d_bse, c->alt_bse);
    std::swap(c->d_extra, c->alt_extra);
    std::swap(c->d_betas, c->alt_betas);
    std::swap(c->d_flags, c->alt_flags);
    std::swap(c->d_counters, c->alt_counters);
    c->tab_cur ^= 1;
    if (c->fetch_valid[c->tab_cur]) {
        PSB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_fetch[c->tab_cur], 0));
        c->fetch_valid[c->tab_cur] = false;
    }
    return PSB_OK;
}

int psb_ensure_capacity(psb_ctx *c, int64_t S, int betas_cols) {
    if (S > c->cap || betas_cols > c->betas_cols) {
        PSB_CUDA(cudaStreamSynchronize(c->stream));
        PSB_CUDA(cudaStreamSynchronize(c->fetch_stream));
        free_tables(c);
        int64_t cap = S < 1024 ? 1024 : S;
        PSB_CUDA(cudaMalloc(&c->d_carriers, cap * sizeof(int32_t)));
        PSB_CUDA(cudaMalloc(&c->d_missing, cap * sizeof(int32_t)));
        PSB_CUDA(cudaMalloc(&c->d_af, cap * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->d_prep, cap * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->d_pvalue, cap * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->d_beta, cap * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->d_bse, cap * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->d_extra, cap * sizeof(double)));
        if (betas_cols > 0)
            PSB_CUDA(cudaMalloc(&c->d_betas, cap * betas_cols * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->d_flags, cap * sizeof(uint32_t)));
        PSB_CUDA(cudaMalloc(&c->alt_carriers, cap * sizeof(int32_t)));
        PSB_CUDA(cudaMalloc(&c->alt_missing, cap * sizeof(int32_t)));
        PSB_CUDA(cudaMalloc(&c->alt_af, cap * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->alt_prep, cap * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->alt_pvalue, cap * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->alt_beta, cap * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->alt_bse, cap * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->alt_extra, cap * sizeof(double)));
        if (betas_cols > 0)
            PSB_CUDA(cudaMalloc(&c->alt_betas, cap * betas_cols * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->alt_flags, cap * sizeof(uint32_t)));
        PSB_CUDA(cudaMalloc(&c->d_tab, cap * 4 * sizeof(int32_t)));
        PSB_CUDA(cudaMalloc(&c->d_idx, (cap + 256) * sizeof(int32_t)));
        PSB_CUDA(cudaMalloc(&c->d_idx2, (cap + 256) * sizeof(int32_t)));
        PSB_CUDA(cudaMalloc(&c->d_idx3, (cap + 256) * sizeof(int32_t)));
        PSB_CUDA(cudaMalloc(&c->d_lineage, cap * sizeof(int32_t)));
        PSB_CUDA(cudaMalloc(&c->d_a, cap * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->d_b, cap * sizeof(double)));
        PSB_CUDA(cudaMalloc(&c->d_pp, cap * sizeof(double)));
        c->cap = cap;
        c->betas_cols = betas_cols;
    }
    size_t need = (size_t)c->cap * (size_t)(c->C > 0 ? c->C : 1) * sizeof(double);
    if (need > c->sums_cap) {
        PSB_CUDA(cudaStreamSynchronize(c->stream));
        free_dev(c->d_sums);
        c->d_sums = nullptr;
        PSB_CUDA(cudaMalloc(&c->d_sums, need));
        c->sums_cap = need;
    }
    return PSB_OK;
}

extern "C" {

static int check_rows(psb_ctx *c, int64_t n_variants, int32_t words_per_row) {
    PSB_REQUIRE(c, PSB_ERR_ARG, "ctx is NULL");
    PSB_REQUIRE(c->model != PSB_MODEL_NONE, PSB_ERR_STATE,
                "psb_submit before psb_lmm_setup / psb_fixed_setup");
    PSB_REQUIRE(n_variants >= 0 && n_variants < (1ll << 31) - 512, PSB_ERR_ARG,
                "n_variants %lld out of range", (long long)n_variants);
    PSB_REQUIRE(words_per_row >= c->Wn && words_per_row % 4 == 0, PSB_ERR_ARG,
                "words_per_row %d must be a multiple of 4 and >= ceil(N/32) = %d",
                words_per_row, c->Wn);
    return PSB_OK;
}

static int stage_reserve(psb_ctx *c, uint32_t **buf, size_t *cap, size_t bytes) {
    if (bytes <= *cap) return PSB_OK;
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->copy_stream));
    free_dev(*buf);
    *buf = nullptr;
    *cap = 0;
    PSB_CUDA(cudaMalloc(buf, bytes));
    *cap = bytes;
    return PSB_OK;
}

int psb_submit(psb_ctx *c, const uint32_t *bits, const uint32_t *missing, int64_t n_variants,
               int32_t words_per_row) {
    PSB_NVTX("psb_submit");
    int rc = check_rows(c, n_variants, words_per_row);
    if (rc) return rc;
    PSB_REQUIRE(bits || n_variants == 0, PSB_ERR_ARG, "bits is NULL");
    PSB_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)n_variants * words_per_row * sizeof(uint32_t);
    const int slot = c->stage_slot ^ 1;          // alternate: the other slot may still be in use
    rc = stage_reserve(c, &c->stage_bits[slot], &c->stage_bits_cap[slot], bytes);
    if (rc) return rc;
    if (missing) {
        rc = stage_reserve(c, &c->stage_miss[slot], &c->stage_miss_cap[slot], bytes);
        if (rc) return rc;
    }
    // the kernels of the run that last read this slot must have finished before it is refilled
    if (c->used_valid[slot]) PSB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_used[slot], 0));
    if (bytes) {
        PSB_CUDA(cudaMemcpyAsync(c->stage_bits[slot], bits, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        if (missing)
            PSB_CUDA(cudaMemcpyAsync(c->stage_miss[slot], missing, bytes, cudaMemcpyHostToDevice,
                                     c->copy_stream));
    }
    PSB_CUDA(cudaEventRecord(c->ev_copy[slot], c->copy_stream));
    c->stage_slot = slot;
    c->sub_bits = c->stage_bits[slot];
    c->sub_miss = missing ? c->stage_miss[slot] : nullptr;
    c->sub_S = n_variants;
    c->sub_Wrow = words_per_row;
    c->sub_slot = slot;
    c->sub_valid = true;
    return PSB_OK;
}

// Asynchronous psb_fetch: queues the copies of the current table (the last psb_run_*) on the
// library's fetch stream, behind the run's last kernel, and returns.  The next psb_run_* writes the
// other set of columns, so it may be queued right away -- the device never idles while a table
// travels.  psb_fetch_wait blocks until the copies of the last psb_fetch_begin have landed and
// returns that run's counters (as psb_counts).  `out` must stay valid until then.
int psb_fetch_begin(psb_ctx *c, const psb_results *out) {
    PSB_NVTX("psb_fetch_begin");
    PSB_REQUIRE(c && out, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(c->ran && c->have_run_ev, PSB_ERR_STATE, "psb_fetch_begin before psb_run_*");
    PSB_CUDA(cudaSetDevice(c->device));
    const int64_t S = c->S;
    cudaStream_t st = c->fetch_stream;
    PSB_CUDA(cudaStreamWaitEvent(st, c->ev_run1, 0));
#define CP(field, src, type)                                                                 \
    if (out->field && S > 0)                                                                 \
        PSB_CUDA(cudaMemcpyAsync(out->field, src, S * sizeof(type), cudaMemcpyDefault, st));
    CP(carriers, c->d_carriers, int32_t)
    CP(missing, c->d_missing, int32_t)
    CP(af, c->d_af, double)
    CP(prep, c->d_prep, double)
    CP(pvalue, c->d_pvalue, double)
    CP(beta, c->d_beta, double)
    CP(bse, c->d_bse, double)
    CP(extra, c->d_extra, double)
    CP(flags, c->d_flags, uint32_t)
#undef CP
    if (out->betas && S > 0 && c->model == PSB_MODEL_FIXED && c->q > 1)
        PSB_CUDA(cudaMemcpyAsync(out->betas, c->d_betas, S * (c->q - 1) * sizeof(double),
                                 cudaMemcpyDefault, st));
    PSB_CUDA(cudaMemcpyAsync(c->h_counters[c->tab_cur], c->d_counters, PSB_N_COUNTERS * sizeof(int), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaEventRecord(c->ev_fetch[c->tab_cur], st));
    c->fetch_valid[c->tab_cur] = true;
    c->fetch_set = c->tab_cur;
    c->fetch_S[c->tab_cur] = S;
    return PSB_OK;
}

int psb_fetch_wait(psb_ctx *c, int64_t counts_out[4]) {
    PSB_REQUIRE(c, PSB_ERR_ARG, "ctx is NULL");
    PSB_REQUIRE(c->fetch_set >= 0, PSB_ERR_STATE, "psb_fetch_wait before psb_fetch_begin");
    PSB_CUDA(cudaSetDevice(c->device));
    const int s = c->fetch_set;
    PSB_CUDA(cudaEventSynchronize(c->ev_fetch[s]));
    if (counts_out) {
        const int *h = c->h_counters[s];
        counts_out[0] = c->fetch_S[s];
        counts_out[1] = h[1] + h[5];
        counts_out[2] = h[0] - h[5];
        counts_out[3] = h[0] - h[5] - h[2];
    }
    return PSB_OK;
}

int psb_submit_device(psb_ctx *c, const void *d_bits, const void *d_missing, int64_t n_variants,
                      int32_t words_per_row) {
    int rc = check_rows(c, n_variants, words_per_row);
    if (rc) return rc;
    PSB_REQUIRE(d_bits || n_variants == 0, PSB_ERR_ARG, "d_bits is NULL");
    c->sub_bits = (const uint32_t *)d_bits;
    c->sub_miss = (const uint32_t *)d_missing;
    c->sub_S = n_variants;
    c->sub_Wrow = words_per_row;
    c->sub_slot = -1;
    c->sub_valid = true;
    return PSB_OK;
}

int psb_fetch(psb_ctx *c, const psb_results *out) {
    PSB_NVTX("psb_fetch");
    PSB_REQUIRE(c && out, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(c->ran, PSB_ERR_STATE, "psb_fetch before psb_run_*");
    PSB_CUDA(cudaSetDevice(c->device));
    int64_t S = c->S;
    cudaStream_t st = c->stream;
#define CP(field, src, type)                                                                 \
    if (out->field && S > 0)                                                                 \
        PSB_CUDA(cudaMemcpyAsync(out->field, src, S * sizeof(type), cudaMemcpyDefault, st));
    CP(carriers, c->d_carriers, int32_t)
    CP(missing, c->d_missing, int32_t)
    CP(af, c->d_af, double)
    CP(prep, c->d_prep, double)
    CP(pvalue, c->d_pvalue, double)
    CP(beta, c->d_beta, double)
    CP(bse, c->d_bse, double)
    CP(extra, c->d_extra, double)
    CP(flags, c->d_flags, uint32_t)
#undef CP
    if (out->betas && S > 0 && c->model == PSB_MODEL_FIXED && c->q > 1)
        PSB_CUDA(cudaMemcpyAsync(out->betas, c->d_betas, S * (c->q - 1) * sizeof(double),
                                 cudaMemcpyDefault, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    return PSB_OK;
}

int psb_results_device(psb_ctx *c, psb_results *o) {
    PSB_REQUIRE(c && o, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(c->ran, PSB_ERR_STATE, "psb_results_device before psb_run_*");
    o->carriers = c->d_carriers; o->missing = c->d_missing; o->af = c->d_af; o->prep = c->d_prep;
    o->pvalue = c->d_pvalue; o->beta = c->d_beta; o->bse = c->d_bse; o->extra = c->d_extra;
    o->betas = c->d_betas; o->flags = c->d_flags;
    return PSB_OK;
}

int psb_counts(psb_ctx *c, int64_t out[4]) {
    PSB_REQUIRE(c && out, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(c->ran, PSB_ERR_STATE, "psb_counts before psb_run_*");
    PSB_CUDA(cudaSetDevice(c->device));
    int h[PSB_N_COUNTERS];
    PSB_CUDA(cudaMemcpyAsync(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    // h[5]: variants the LMM epilogue found to fail the (deferred) Welch pre-filter after they had
    // been compacted as tested
    out[0] = c->S;              // loaded
    out[1] = h[1] + h[5];       // pre-filtered (af + prefilter)
    out[2] = h[0] - h[5];       // tested
    out[3] = h[0] - h[5] - h[2];  // tested and passed the lrt filter
    return PSB_OK;
}

int psb_last_stats(psb_ctx *c, int64_t out[4]) {
    PSB_REQUIRE(c && out, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(c->ran, PSB_ERR_STATE, "psb_last_stats before psb_run_*");
    PSB_CUDA(cudaSetDevice(c->device));
    int h[PSB_N_COUNTERS];
    PSB_CUDA(cudaMemcpyAsync(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    out[0] = h[4];      // Newton evaluations (passes over the samples) of the Logit kernel
    out[1] = h[3];      // variants handed to the Firth kernel
    out[2] = h[2];      // variants that failed the lrt filter
    if (getenv("PSB_DEBUG_STATS"))
        fprintf(stderr, "psb stats: firth max iterations %d (variant %d), max halvings of a fit %d (variant %d)\n", h[9], h[11], h[10], h[12]);
    out[3] = h[8];      // penalised-likelihood evaluations of the Firth kernel that met a singular matrix
    return PSB_OK;
}

int psb_last_ms(psb_ctx *c, int32_t which, float *ms) {
    PSB_REQUIRE(c && ms, PSB_ERR_ARG, "NULL argument");
    PSB_CUDA(cudaSetDevice(c->device));
    *ms = 0.f;
    if (which == 0) {
        PSB_REQUIRE(c->have_run_ev, PSB_ERR_STATE, "no run recorded");
        PSB_CUDA(cudaEventSynchronize(c->ev_run1));
        PSB_CUDA(cudaEventElapsedTime(ms, c->ev_run0, c->ev_run1));
    } else {
        PSB_REQUIRE(c->have_k_ev, PSB_ERR_STATE, "no kernel timing recorded");
        PSB_CUDA(cudaEventSynchronize(c->ev_k1));
        PSB_CUDA(cudaEventElapsedTime(ms, c->ev_k0, c->ev_k1));
    }
    return PSB_OK;
}

int psb_event_record(psb_ctx *c, int32_t slot) {
    PSB_REQUIRE(c && slot >= 0 && slot < 8, PSB_ERR_ARG, "bad event slot");
    PSB_CUDA(cudaSetDevice(c->device));
    PSB_CUDA(cudaEventRecord(c->ev_user[slot], c->stream));
    return PSB_OK;
}

int psb_event_elapsed(psb_ctx *c, int32_t a, int32_t b, float *ms) {
    PSB_REQUIRE(c && ms && a >= 0 && a < 8 && b >= 0 && b < 8, PSB_ERR_ARG, "bad event slot");
    PSB_CUDA(cudaSetDevice(c->device));
    PSB_CUDA(cudaEventSynchronize(c->ev_user[b]));
    PSB_CUDA(cudaEventElapsedTime(ms, c->ev_user[a], c->ev_user[b]));
    return PSB_OK;
}

int psb_host_alloc(size_t bytes, void **out) {
    PSB_REQUIRE(out, PSB_ERR_ARG, "out is NULL");
    *out = nullptr;
    PSB_CUDA(cudaHostAlloc(out, bytes > 0 ? bytes : 1, cudaHostAllocPortable));
    return PSB_OK;
}

int psb_host_free(void *ptr) {
    if (ptr) PSB_CUDA(cudaFreeHost(ptr));
    return PSB_OK;
}

int psb_download_bits(psb_ctx *c, uint32_t *out_bits) {
    PSB_REQUIRE(c && out_bits, PSB_ERR_ARG, "NULL argument");
    PSB_CUDA(cudaSetDevice(c->device));
    int rc = psb_run_begin(c);
    if (rc) return rc;
    PSB_REQUIRE(c->d_bits || c->S == 0, PSB_ERR_STATE, "no rows submitted");
    size_t bytes = (size_t)c->S * c->Wrow * sizeof(uint32_t);
    if (bytes)
        PSB_CUDA(cudaMemcpyAsync(out_bits, c->d_bits, bytes, cudaMemcpyDeviceToHost, c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    return PSB_OK;
}

int psb_download_rows(psb_ctx *c, uint32_t *out_bits, uint32_t *out_missing, int32_t *has_missing) {
    PSB_REQUIRE(c && out_bits, PSB_ERR_ARG, "NULL argument");
    PSB_CUDA(cudaSetDevice(c->device));
    int rc = psb_run_begin(c);
    if (rc) return rc;
    PSB_REQUIRE(c->d_bits || c->S == 0, PSB_ERR_STATE, "no rows submitted");
    size_t bytes = (size_t)c->S * c->Wrow * sizeof(uint32_t);
    if (bytes) {
        PSB_CUDA(cudaMemcpyAsync(out_bits, c->d_bits, bytes, cudaMemcpyDeviceToHost, c->stream));
        if (out_missing && c->d_miss)
            PSB_CUDA(cudaMemcpyAsync(out_missing, c->d_miss, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    if (has_missing) *has_missing = c->d_miss ? 1 : 0;
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    return PSB_OK;
}

int psb_launch_count(psb_ctx *c, int64_t *n) {
    PSB_REQUIRE(c && n, PSB_ERR_ARG, "NULL argument");
    *n = c->launches;
    return PSB_OK;
}

double psb_host_chi2_sf1(double x) { return psb_chi2_sf1(x); }
double psb_host_f_sf_1(double x, double dfd) { return psb_t2_sf(x, dfd); }
double psb_host_t2_sf(double t, double df) { return psb_t2_sf(t * t, df); }

}  // extern "C"
