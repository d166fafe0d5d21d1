// psb_synth.cu -- seeded synthetic k-mer presence/absence rows (bench + tests).
//
// x_{s,i} ~ Bernoulli(af_s), af_s ~ U(af_lo, af_hi); counter-based so the row
// of variant id s does not depend on batch boundaries or GPU count.  Variants whose id is
// a multiple of `planted_every` are correlated with the phenotype sign so the far p-value
// tail is exercised.  The host and device generators are the same function.
#include "psb_internal.cuh"

__host__ __device__ __forceinline__ uint64_t psb_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct psb_synth_row {
    uint64_t key;
    double af;
    double delta;   // 0 unless planted
    double sep;     // 0 unless "separated": carried by positive-sign samples only, with this probability
};

__host__ __device__ __forceinline__ psb_synth_row psb_synth_rowinfo(uint64_t seed, int64_t vid,
                                                                    double af_lo, double af_hi,
                                                                    int planted_every, int separated_every) {
    psb_synth_row r;
    r.key = psb_mix64(seed ^ psb_mix64((uint64_t)vid));
    double u = (double)(r.key >> 11) * (1.0 / 9007199254740992.0);
    r.af = af_lo + (af_hi - af_lo) * u;
    r.delta = 0.0;
    r.sep = 0.0;
    if (planted_every > 0 && vid % planted_every == 0) {
        r.af = 0.5;
        r.delta = 0.30 * (double)((r.key >> 3) & 0xFFFFull) * (1.0 / 65536.0);
    } else if (separated_every > 0 && vid % separated_every == separated_every / 2) {
        // rare variant found in positive-sign samples only (a k-mer private to a clade of cases):
        // an empty cell of the 2x2 table -> 'bad-chisq' -> Firth regression (model.py:326, 355)
        r.sep = 0.02 + 0.04 * (double)((r.key >> 19) & 0xFFFFull) * (1.0 / 65536.0);
    }
    return r;
}

// word w of the row: samples 32w .. 32w+31
__host__ __device__ __forceinline__ uint32_t psb_synth_word(const psb_synth_row &r, int w, int N,
                                                            const int8_t *y_sign) {
    uint32_t out = 0;
    for (int h = 0; h < 16; ++h) {
        uint64_t z = psb_mix64(r.key + 0x632BE59BD9B4E019ull * (uint64_t)(w * 16 + h + 1));
        for (int e = 0; e < 2; ++e) {
            int i = w * 32 + h * 2 + e;
            if (i >= N) break;
            uint32_t u32 = e ? (uint32_t)(z >> 32) : (uint32_t)z;
            double p = r.af;
            if (r.delta != 0.0 && y_sign) p += r.delta * (double)y_sign[i];
            if (r.sep != 0.0 && y_sign) p = y_sign[i] > 0 ? r.sep : 0.0;
            double thr = p * 4294967296.0;
            uint32_t t = thr >= 4294967295.0 ? 0xFFFFFFFFu : (thr <= 0.0 ? 0u : (uint32_t)thr);
            if (u32 < t) out |= 1u << (h * 2 + e);
        }
    }
    return out;
}

__global__ void k_synth(uint32_t *__restrict__ bits, int64_t n_variants, int Wrow, int N,
                        uint64_t seed, int64_t first, double af_lo, double af_hi,
                        int planted_every, int separated_every, const int8_t *__restrict__ y_sign) {
    int64_t total = n_variants * Wrow;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = e / Wrow;
        int w = (int)(e - s * Wrow);
        uint32_t word = 0;
        if (w * 32 < N) {
            psb_synth_row r = psb_synth_rowinfo(seed, first + s, af_lo, af_hi, planted_every, separated_every);
            word = psb_synth_word(r, w, N, y_sign);
        }
        bits[e] = word;
    }
}

extern "C" int psb_synth_device(psb_ctx *c, uint64_t seed, int64_t first_variant, int64_t n_variants,
                                int32_t n_samples, double af_lo, double af_hi,
                                int32_t planted_every, int32_t separated_every, const int8_t *y_sign) {
    PSB_REQUIRE(c, PSB_ERR_ARG, "ctx is NULL");
    PSB_REQUIRE(c->model != PSB_MODEL_NONE && n_samples == c->N, PSB_ERR_STATE,
                "psb_synth_device needs a model set up with the same n_samples");
    PSB_REQUIRE((planted_every <= 0 && separated_every <= 0) || y_sign, PSB_ERR_ARG, "planted variants need y_sign");
    PSB_CUDA(cudaSetDevice(c->device));
    int Wrow = ((c->Wn + 3) / 4) * 4;
    size_t bytes = (size_t)n_variants * Wrow * sizeof(uint32_t);
    if (bytes > c->own_bits_cap) {
        PSB_CUDA(cudaStreamSynchronize(c->stream));
        if (c->own_bits) cudaFree(c->own_bits);
        c->own_bits = nullptr;
        c->own_bits_cap = 0;
        PSB_CUDA(cudaMalloc(&c->own_bits, bytes));
        c->own_bits_cap = bytes;
    }
    int8_t *d_sign = nullptr;
    if (y_sign) {
        PSB_CUDA(cudaMalloc(&d_sign, n_samples));
        PSB_CUDA(cudaMemcpyAsync(d_sign, y_sign, n_samples, cudaMemcpyHostToDevice, c->stream));
    }
    if (n_variants > 0) {
        k_synth<<<c->sm_count * 8, 256, 0, c->stream>>>(c->own_bits, n_variants, Wrow, n_samples,
                                                       seed, first_variant, af_lo, af_hi,
                                                       planted_every, separated_every, d_sign);
        c->launches++;
        PSB_CUDA(cudaGetLastError());
    }
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    if (d_sign) cudaFree(d_sign);
    c->sub_bits = c->own_bits;
    c->sub_miss = nullptr;
    c->sub_S = n_variants;
    c->sub_Wrow = Wrow;
    c->sub_slot = -1;
    c->sub_valid = true;
    return PSB_OK;
}

extern "C" int psb_synth_host(uint64_t seed, int64_t first_variant, int64_t n_variants,
                              int32_t n_samples, double af_lo, double af_hi,
                              int32_t planted_every, int32_t separated_every, const int8_t *y_sign,
                              uint32_t *out_bits, int32_t words_per_row) {
    PSB_REQUIRE(out_bits, PSB_ERR_ARG, "out_bits is NULL");
    PSB_REQUIRE(words_per_row * 32 >= n_samples, PSB_ERR_ARG, "words_per_row too small");
    PSB_REQUIRE((planted_every <= 0 && separated_every <= 0) || y_sign, PSB_ERR_ARG, "planted variants need y_sign");
    for (int64_t s = 0; s < n_variants; ++s) {
        psb_synth_row r = psb_synth_rowinfo(seed, first_variant + s, af_lo, af_hi, planted_every, separated_every);
        for (int w = 0; w < words_per_row; ++w)
            out_bits[s * words_per_row + w] = (w * 32 < n_samples) ? psb_synth_word(r, w, n_samples, y_sign) : 0u;
    }
    return PSB_OK;
}
