// psb_text.cu -- k-mer text straight to the device: the tokenising half of input.read_variant's
// k-mer branch (pyseer/input.py:377-388 with the common tail :438-452) as a kernel.
//
// The reference splits every line `name | s1:c1 s2:c2 ...` in Python, looks each sample up in a
// dictionary and builds a length-N vector.  The native host reader (psb_io.cu) does the same on
// --cpu threads and tops out near 1 GB/s of text.  Here the host only cuts the decompressed text
// into lines (psb_reader_next_text); the text itself goes to the GPU through the staging copy
// stream and one warp per line turns it into the packed row psb_submit would have received:
//
//   * the sample tokens start after the first '|' (the host reader's rule; whitespace = ' ', '\t');
//   * a lane owns one byte of a 32-byte window; a lane whose byte opens a token hashes the sample
//     name up to ':' (FNV-1a, 32 bit), probes an open-addressing table of the phenotyped samples
//     (built once by psb_text_setup), confirms the match byte by byte against the name pool and
//     sets bit (i % 32) of word (i / 32) of the row in shared memory;
//   * the finished row is written once, coalesced, into the staging slot.
//
// Unknown samples are ignored, a sample listed twice sets the same bit (the reference's dictionary
// semantics).  This is HBM/PCIe-bound byte work: 19 KB of text per variant at N = 5000 against the
// 640 B packed row, so the text copy (not the kernel) sets the rate.
#include <string.h>

#include <algorithm>
#include <vector>

#include "psb_internal.cuh"

struct psb_text_state {
    uint64_t *d_table = nullptr;      // (hash32 << 32) | (sample index + 1); 0 = empty
    uint32_t mask = 0;
    char *d_pool = nullptr;           // sample names back to back
    int32_t *d_pool_off = nullptr;    // [N + 1]
    char *d_text = nullptr;           // text of the batch being parsed
    size_t text_cap = 0;
    int64_t *d_lstart = nullptr;      // [lines_cap]
    int32_t *d_llen = nullptr;
    size_t lines_cap = 0;
    int32_t *d_info[2] = {nullptr, nullptr};   // per staging slot: bit 1 = no observation, bit 2 = no '|'
    size_t info_cap[2] = {0, 0};
};

__host__ __device__ __forceinline__ uint32_t psb_fnv1a_step(uint32_t h, unsigned char c) {
    return (h ^ (uint32_t)c) * 16777619u;
}

#define TXT_WARPS 8

__global__ void __launch_bounds__(TXT_WARPS * 32)
k_text_kmers(const char *__restrict__ text, const int64_t *__restrict__ lstart,
             const int32_t *__restrict__ llen, int64_t n_lines, const uint64_t *__restrict__ table,
             uint32_t mask, const char *__restrict__ pool, const int32_t *__restrict__ pool_off,
             uint32_t *__restrict__ rows, int Wrow, int32_t *__restrict__ info) {
    extern __shared__ uint32_t txt_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *row = txt_smem + (size_t)warp * Wrow;
    for (int w = lane; w < Wrow; w += 32) row[w] = 0u;
    __syncwarp();
    const int64_t warps_total = (int64_t)gridDim.x * TXT_WARPS;
    for (int64_t line = (int64_t)blockIdx.x * TXT_WARPS + warp; line < n_lines; line += warps_total) {
        const unsigned char *L = reinterpret_cast<const unsigned char *>(text) + lstart[line];
        const int len = llen[line];
        // first '|' of the line
        int bar = len;
        for (int base = 0; base < len; base += 32) {
            const int i = base + lane;
            const unsigned m = __ballot_sync(0xffffffffu, i < len && L[i] == '|');
            if (m) {
                bar = base + __ffs(m) - 1;
                break;
            }
        }
        int flags = 0;
        if (bar == len) {
            flags = 4;                                   // malformed: reported by the host wrapper
        } else {
            bool prev_ws = true;                         // the byte before the first sample byte is '|'
            for (int base = bar + 1; base < len; base += 32) {
                const int i = base + lane;
                const unsigned char c = i < len ? L[i] : (unsigned char)' ';
                const bool ws = c == ' ' || c == '\t';
                const unsigned wsm = __ballot_sync(0xffffffffu, ws);
                const bool before_ws = lane == 0 ? prev_ws : ((wsm >> (lane - 1)) & 1u) != 0;
                prev_ws = (wsm >> 31) & 1u;
                if (!ws && before_ws) {
                    // a token opens here: the sample name runs to ':' or to the end of the token
                    uint32_t h = 2166136261u;
                    int j = i;
                    while (j < len) {
                        const unsigned char ch = L[j];
                        if (ch == ':' || ch == ' ' || ch == '\t') break;
                        h = psb_fnv1a_step(h, ch);
                        ++j;
                    }
                    const int nl = j - i;
                    uint32_t slot = h & mask;
                    for (;;) {
                        const uint64_t e = __ldg(table + slot);
                        if (e == 0ull) break;
                        if ((uint32_t)(e >> 32) == h) {
                            const int s = (int)(uint32_t)e - 1;
                            const int o = __ldg(pool_off + s);
                            if (__ldg(pool_off + s + 1) - o == nl) {
                                bool same = true;
                                for (int t = 0; t < nl; ++t)
                                    if ((unsigned char)__ldg(pool + o + t) != L[i + t]) {
                                        same = false;
                                        break;
                                    }
                                if (same) {
                                    atomicOr(row + (s >> 5), 1u << (s & 31));
                                    break;
                                }
                            }
                        }
                        slot = (slot + 1) & mask;
                    }
                }
            }
        }
        __syncwarp();
        uint32_t any = 0u;
        uint32_t *out = rows + (size_t)line * Wrow;
        for (int w = lane; w < Wrow; w += 32) {
            const uint32_t v = row[w];
            any |= v;
            out[w] = v;
            row[w] = 0u;
        }
        any = __reduce_or_sync(0xffffffffu, any);
        if (lane == 0) info[line] = flags | (any ? 0 : 2);
        __syncwarp();
    }
}

static void text_free(psb_text_state *t) {
    if (!t) return;
    cudaFree(t->d_table);
    cudaFree(t->d_pool);
    cudaFree(t->d_pool_off);
    cudaFree(t->d_text);
    cudaFree(t->d_lstart);
    cudaFree(t->d_llen);
    cudaFree(t->d_info[0]);
    cudaFree(t->d_info[1]);
    delete t;
}

void psb_text_release(psb_ctx *c) {
    text_free(static_cast<psb_text_state *>(c->text));
    c->text = nullptr;
}

template <typename T>
static int text_reserve(psb_ctx *c, T **buf, size_t *cap, size_t n) {
    if (n <= *cap) return PSB_OK;
    // the buffers are only touched on the copy stream, whose earlier work must drain first
    PSB_CUDA(cudaStreamSynchronize(c->copy_stream));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    const size_t want = n + n / 4 + 64;
    PSB_CUDA(cudaMalloc(buf, want * sizeof(T)));
    *cap = want;
    return PSB_OK;
}

extern "C" {

// Builds the device lookup table of the phenotyped samples (phenotype order = bit order of the
// rows, as input.py:450 builds k).  Call once after psb_lmm_setup / psb_fixed_setup.
int psb_text_setup(psb_ctx *c, const char *const *sample_names, int32_t n_samples) {
    PSB_REQUIRE(c && sample_names, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(c->model != PSB_MODEL_NONE, PSB_ERR_STATE, "psb_text_setup before psb_lmm_setup / psb_fixed_setup");
    PSB_REQUIRE(n_samples == c->N, PSB_ERR_ARG, "n_samples %d does not match the model's %d", n_samples, c->N);
    PSB_CUDA(cudaSetDevice(c->device));
    psb_text_release(c);
    psb_text_state *t = new psb_text_state();
    c->text = t;
    uint32_t cap = 64;
    while (cap < 4u * (uint32_t)n_samples) cap <<= 1;
    std::vector<uint64_t> tab(cap, 0ull);
    std::vector<char> pool;
    std::vector<int32_t> off(n_samples + 1, 0);
    for (int s = 0; s < n_samples; ++s) {
        const char *nm = sample_names[s];
        PSB_REQUIRE(nm, PSB_ERR_ARG, "sample name %d is NULL", s);
        const size_t nl = strlen(nm);
        uint32_t h = 2166136261u;
        for (size_t k = 0; k < nl; ++k) h = psb_fnv1a_step(h, (unsigned char)nm[k]);
        off[s] = (int32_t)pool.size();
        pool.insert(pool.end(), nm, nm + nl);
        // a name listed twice keeps its FIRST index, like the map of the host reader (emplace)
        uint32_t slot = h & (cap - 1);
        bool dup = false;
        while (tab[slot] != 0ull) {
            if ((uint32_t)(tab[slot] >> 32) == h) {
                const int o = (int)(uint32_t)tab[slot] - 1;
                if ((size_t)(off[o + 1] - off[o]) == nl && memcmp(pool.data() + off[o], nm, nl) == 0) {
                    dup = true;
                    break;
                }
            }
            slot = (slot + 1) & (cap - 1);
        }
        off[s + 1] = (int32_t)pool.size();
        if (!dup) tab[slot] = ((uint64_t)h << 32) | (uint64_t)(uint32_t)(s + 1);
    }
    if (pool.empty()) pool.push_back('\0');
    t->mask = cap - 1;
    PSB_CUDA(cudaMalloc(&t->d_table, cap * sizeof(uint64_t)));
    PSB_CUDA(cudaMalloc(&t->d_pool, pool.size()));
    PSB_CUDA(cudaMalloc(&t->d_pool_off, off.size() * sizeof(int32_t)));
    PSB_CUDA(cudaMemcpy(t->d_table, tab.data(), cap * sizeof(uint64_t), cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(t->d_pool, pool.data(), pool.size(), cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(t->d_pool_off, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    PSB_UPLOAD_FENCE();
    return PSB_OK;
}

// psb_submit for k-mer TEXT: `text` holds n_lines lines, line v = text[line_start[v] .. + line_len[v])
// without its newline and trailing blanks (what psb_reader_next_text returns).  The text is copied
// on the staging copy stream (page-locked `text` makes that a true asynchronous DMA; the caller
// keeps it alive until the batch has been fetched) and parsed there into the rows of the next
// staging slot -- exactly as if psb_submit had been called with the rows the host reader builds.
int psb_submit_text(psb_ctx *c, const char *text, int64_t text_bytes, const int64_t *line_start,
                    const int32_t *line_len, int64_t n_lines) {
    PSB_NVTX("psb_submit_text");
    PSB_REQUIRE(c, PSB_ERR_ARG, "ctx is NULL");
    PSB_REQUIRE(c->text, PSB_ERR_STATE, "psb_submit_text before psb_text_setup");
    PSB_REQUIRE(n_lines >= 0 && n_lines < (1ll << 31) - 512, PSB_ERR_ARG, "n_lines %lld out of range",
                (long long)n_lines);
    PSB_REQUIRE(n_lines == 0 || (text && line_start && line_len && text_bytes > 0), PSB_ERR_ARG, "NULL argument");
    PSB_CUDA(cudaSetDevice(c->device));
    psb_text_state *t = static_cast<psb_text_state *>(c->text);
    const int Wrow = 4 * ((c->N + 127) / 128);
    const size_t smem = (size_t)TXT_WARPS * Wrow * sizeof(uint32_t);
    PSB_REQUIRE(smem <= 200 * 1024, PSB_ERR_UNSUPPORTED, "too many samples (%d) for the device text parser", c->N);
    const size_t bytes = (size_t)n_lines * Wrow * sizeof(uint32_t);
    const int slot = c->stage_slot ^ 1;
    if (bytes > c->stage_bits_cap[slot]) {
        PSB_CUDA(cudaStreamSynchronize(c->stream));
        PSB_CUDA(cudaStreamSynchronize(c->copy_stream));
        cudaFree(c->stage_bits[slot]);
        c->stage_bits[slot] = nullptr;
        c->stage_bits_cap[slot] = 0;
        PSB_CUDA(cudaMalloc(&c->stage_bits[slot], bytes));
        c->stage_bits_cap[slot] = bytes;
    }
    int rc = text_reserve(c, &t->d_text, &t->text_cap, (size_t)text_bytes);
    if (rc) return rc;
    size_t lc = t->lines_cap;
    rc = text_reserve(c, &t->d_lstart, &lc, (size_t)n_lines);
    if (rc) return rc;
    lc = t->lines_cap;
    rc = text_reserve(c, &t->d_llen, &lc, (size_t)n_lines);
    if (rc) return rc;
    t->lines_cap = lc;
    rc = text_reserve(c, &t->d_info[slot], &t->info_cap[slot], (size_t)n_lines);
    if (rc) return rc;
    if (c->used_valid[slot]) PSB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_used[slot], 0));
    if (n_lines > 0) {
        PSB_CUDA(cudaMemcpyAsync(t->d_text, text, (size_t)text_bytes, cudaMemcpyHostToDevice, c->copy_stream));
        PSB_CUDA(cudaMemcpyAsync(t->d_lstart, line_start, (size_t)n_lines * sizeof(int64_t),
                                 cudaMemcpyHostToDevice, c->copy_stream));
        PSB_CUDA(cudaMemcpyAsync(t->d_llen, line_len, (size_t)n_lines * sizeof(int32_t), cudaMemcpyHostToDevice,
                                 c->copy_stream));
        if (smem > 48 * 1024)
            PSB_CUDA(cudaFuncSetAttribute(k_text_kmers, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t blocks = std::min<int64_t>((n_lines + TXT_WARPS - 1) / TXT_WARPS, (int64_t)c->sm_count * 8);
        k_text_kmers<<<(int)blocks, TXT_WARPS * 32, smem, c->copy_stream>>>(
            t->d_text, t->d_lstart, t->d_llen, n_lines, t->d_table, t->mask, t->d_pool, t->d_pool_off,
            c->stage_bits[slot], Wrow, t->d_info[slot]);
        c->launches++;
        PSB_CUDA(cudaGetLastError());
    }
    PSB_CUDA(cudaEventRecord(c->ev_copy[slot], c->copy_stream));
    c->stage_slot = slot;
    c->sub_bits = c->stage_bits[slot];
    c->sub_miss = nullptr;
    c->sub_S = n_lines;
    c->sub_Wrow = Wrow;
    c->sub_slot = slot;
    c->sub_valid = true;
    return PSB_OK;
}

// Per-line flags of the batch the last psb_run_* worked on, when it came from psb_submit_text:
// bit 1 (2) = no observation in the selected samples (input.py:447-448), bit 2 (4) = line without a
// '|' separator (the host reader's PSB_ERR_ARG).  Synchronises with the device.
int psb_text_info(psb_ctx *c, int32_t *info, int64_t n) {
    PSB_REQUIRE(c && info, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(c->text, PSB_ERR_STATE, "psb_text_info before psb_text_setup");
    psb_text_state *t = static_cast<psb_text_state *>(c->text);
    PSB_REQUIRE(c->bits_slot >= 0 && n <= c->S && (size_t)n <= t->info_cap[c->bits_slot], PSB_ERR_STATE,
                "the current batch was not submitted as text");
    PSB_CUDA(cudaSetDevice(c->device));
    PSB_CUDA(cudaMemcpyAsync(info, t->d_info[c->bits_slot], (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost,
                             c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    return PSB_OK;
}

}  // extern "C"
