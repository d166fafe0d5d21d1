// psb_io.cu -- native variant-file reader: k-mer text (`name | s1:1 s2:1 ...`) and Rtab
// rows straight into the packed bit rows psb_submit takes.
//
// Replaces the per-line Python of input.read_variant (pyseer/input.py:301-454; k-mer branch
// :377-388, Rtab branch :412-436, common tail :438-452) for the two text formats: sample
// names are resolved through a hash map built once from the phenotype order, presence goes
// directly to bit (i % 32) of word (i / 32), NaN genotypes ('.' or '' in Rtab) to the
// `missing` rows.  gzip input is read through zlib (gzread handles plain files as well).
// Host-only code; lives in the same shared library as the kernels.
#include <zlib.h>

#include <algorithm>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "psb_internal.cuh"

struct psb_reader {
    gzFile fh = nullptr;
    int var_type = 0;                       // 0 = k-mers, 1 = Rtab
    int n_samples = 0;
    std::vector<std::string> names;         // owns the keys of `index`
    std::unordered_map<std::string_view, int> index;
    std::vector<int> rtab_col;              // Rtab column -> sample index or -1
    std::vector<char> buf;                  // read buffer
    size_t pos = 0, len = 0;
    bool eof = false;
    std::string line;
    std::vector<std::string> lines;         // lines of the batch being parsed
    int n_threads = 1;
};

static bool reader_fill(psb_reader *r) {
    if (r->eof) return false;
    int got = gzread(r->fh, r->buf.data(), (unsigned)r->buf.size());
    if (got <= 0) {
        r->eof = true;
        r->pos = r->len = 0;
        return false;
    }
    r->pos = 0;
    r->len = (size_t)got;
    return true;
}

// next line (without the newline) into r->line; false at end of file
static bool reader_getline(psb_reader *r) {
    r->line.clear();
    for (;;) {
        if (r->pos == r->len && !reader_fill(r)) return !r->line.empty();
        const char *p = r->buf.data() + r->pos;
        const char *nl = (const char *)memchr(p, '\n', r->len - r->pos);
        if (nl) {
            r->line.append(p, nl - p);
            r->pos += (size_t)(nl - p) + 1;
            return true;
        }
        r->line.append(p, r->len - r->pos);
        r->pos = r->len;
    }
}

extern "C" int psb_reader_open(const char *path, int32_t var_type, const char *const *sample_names,
                               int32_t n_samples, psb_reader **out) {
    PSB_REQUIRE(path && sample_names && out, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(var_type == 0 || var_type == 1, PSB_ERR_ARG, "var_type must be 0 (k-mers) or 1 (Rtab)");
    PSB_REQUIRE(n_samples > 0, PSB_ERR_ARG, "no samples");
    *out = nullptr;
    gzFile fh = gzopen(path, "rb");
    PSB_REQUIRE(fh, PSB_ERR_ARG, "cannot open %s", path);
    gzbuffer(fh, 1 << 20);
    psb_reader *r = new psb_reader();
    r->fh = fh;
    r->var_type = var_type;
    r->n_samples = n_samples;
    r->buf.resize(4 << 20);
    r->names.reserve(n_samples);
    for (int i = 0; i < n_samples; ++i) r->names.emplace_back(sample_names[i]);
    r->index.reserve((size_t)n_samples * 2);
    for (int i = 0; i < n_samples; ++i) r->index.emplace(std::string_view(r->names[i]), i);
    if (var_type == 1) {
        // header: first field is the row label, the rest are sample names in column order
        if (!reader_getline(r)) {
            gzclose(fh);
            delete r;
            psb_set_error("%s: empty Rtab file", path);
            return PSB_ERR_ARG;
        }
        const std::string &h = r->line;
        size_t i = 0, n = h.size();
        int field = 0;
        while (i < n) {
            while (i < n && (h[i] == ' ' || h[i] == '\t' || h[i] == '\r')) ++i;
            size_t j = i;
            while (j < n && h[j] != ' ' && h[j] != '\t' && h[j] != '\r') ++j;
            if (j > i) {
                if (field > 0) {
                    auto it = r->index.find(std::string_view(h.data() + i, j - i));
                    r->rtab_col.push_back(it == r->index.end() ? -1 : it->second);
                }
                ++field;
            }
            i = j;
        }
    }
    *out = r;
    return PSB_OK;
}

extern "C" int psb_reader_close(psb_reader *r) {
    if (!r) return PSB_OK;
    if (r->fh) gzclose(r->fh);
    delete r;
    return PSB_OK;
}

// One parsed line -> one packed row.  Returns PSB_OK or an error code with *err set to a static
// message (worker threads do not touch the thread-local error text).
struct psb_line_out {
    int flags = 0;
    int rc = PSB_OK;
    const char *err = nullptr;
};

static void parse_line(const psb_reader *r, const char *L, size_t len, uint32_t *row, uint32_t *mrow,
                       int words_per_row, psb_line_out *out) {
    memset(row, 0, (size_t)words_per_row * 4);
    if (mrow) memset(mrow, 0, (size_t)words_per_row * 4);
    int flags = 0;
    bool seen = false;
#define LINE_REQUIRE(cond, code, msg)  \
    do {                               \
        if (!(cond)) {                 \
            out->rc = (code);          \
            out->err = (msg);          \
            return;                    \
        }                              \
    } while (0)
    if (r->var_type == 0) {
        // samples after the first '|'
        const char *bar = (const char *)memchr(L, '|', len);
        LINE_REQUIRE(bar, PSB_ERR_ARG, "k-mer line without '|' separator");
        size_t k = (size_t)(bar - L) + 1;
        while (k < len) {
            while (k < len && (L[k] == ' ' || L[k] == '\t')) ++k;
            size_t e = k;
            while (e < len && L[e] != ' ' && L[e] != '\t') ++e;
            if (e > k) {
                size_t c = k;
                while (c < e && L[c] != ':') ++c;
                auto it = r->index.find(std::string_view(L + k, c - k));
                if (it != r->index.end()) {
                    row[it->second >> 5] |= 1u << (it->second & 31);
                    seen = true;
                }
            }
            k = e;
        }
    } else {
        size_t i = 0;
        while (i < len && L[i] != '\t') ++i;
        LINE_REQUIRE(i < len, PSB_ERR_ARG, "No sample data found; is this a Rtab file?");
        size_t col = 0, k = i + 1;
        const size_t ncol = r->rtab_col.size();
        for (;;) {
            size_t e = k;
            while (e < len && L[e] != '\t') ++e;
            LINE_REQUIRE(col < ncol, PSB_ERR_ARG, "Unexpected mismatch between header and data row");
            const size_t fl = e - k;
            const int s = r->rtab_col[col];
            if (fl == 1 && L[k] == '1') {
                if (s >= 0) { row[s >> 5] |= 1u << (s & 31); seen = true; }
            } else if (fl == 0 || (fl == 1 && L[k] == '.')) {
                if (s >= 0) {
                    flags |= 1;
                    seen = true;
                    if (mrow) mrow[s >> 5] |= 1u << (s & 31);
                }
            } else {
                LINE_REQUIRE(fl == 1 && L[k] == '0', PSB_ERR_ARG, "Rtab file not binary");
            }
            ++col;
            if (e >= len) break;
            k = e + 1;
        }
        LINE_REQUIRE(col == ncol, PSB_ERR_ARG, "Unexpected mismatch between header and data row");
        LINE_REQUIRE(!(flags & 1) || mrow, PSB_ERR_ARG,
                     "row has missing genotypes but no missing buffer was given");
    }
#undef LINE_REQUIRE
    if (!seen) flags |= 2;
    out->flags = flags;
}

// Parser threads for psb_reader_next (pyseer's --cpu, which this package uses for parsing only):
// the lines of a batch are read and split serially (zlib is a serial stream), then parsed into
// their rows in parallel -- lines are independent and every row is written by one thread.
extern "C" int psb_reader_set_threads(psb_reader *r, int32_t n_threads) {
    PSB_REQUIRE(r, PSB_ERR_ARG, "reader is NULL");
    r->n_threads = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
    return PSB_OK;
}

// Reads up to max_variants rows.  bits / missing: max_variants x words_per_row (zeroed here);
// names: concatenated NUL-terminated variant names (names_cap bytes); name_off[v] = offset of
// name v; info[v]: bit 0 = row has missing genotypes, bit 1 = no observation in the selected
// samples (the reference writes "No observations of ..." to stderr, input.py:447-448).
// *n_read = rows produced (0 at end of file).  Returns PSB_ERR_NOMEM when names_cap is too
// small for a single name, PSB_ERR_ARG on a malformed row.
extern "C" int psb_reader_next(psb_reader *r, int64_t max_variants, uint32_t *bits, uint32_t *missing,
                               int32_t words_per_row, char *names, int64_t names_cap,
                               int64_t *name_off, int32_t *info, int64_t *n_read, int32_t *any_missing) {
    PSB_REQUIRE(r && bits && names && name_off && info && n_read, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(words_per_row * 32 >= r->n_samples, PSB_ERR_ARG, "words_per_row too small");
    *n_read = 0;
    if (any_missing) *any_missing = 0;
    // ---- serial: lines of the batch and their names -----------------------------------
    std::vector<std::string> &lines = r->lines;
    int64_t n = 0, used = 0;
    while (n < max_variants) {
        // stop early when the name buffer may not hold another name (keep 64 KiB free)
        if (names_cap - used < (1 << 16) && n > 0) break;
        if (!reader_getline(r)) break;
        std::string &L = r->line;
        size_t len = L.size();
        while (len > 0 && (L[len - 1] == '\r' || L[len - 1] == ' ' || L[len - 1] == '\t')) --len;
        if (len == 0) continue;
        L.resize(len);
        size_t i = 0, j = 0;
        if (r->var_type == 0) {      // name = first whitespace-delimited token
            while (i < len && (L[i] == ' ' || L[i] == '\t')) ++i;
            j = i;
            while (j < len && L[j] != ' ' && L[j] != '\t') ++j;
        } else {                     // name = first tab-delimited field
            while (j < len && L[j] != '\t') ++j;
        }
        const size_t name_len = j - i;
        PSB_REQUIRE((int64_t)name_len + 1 <= names_cap - used, PSB_ERR_NOMEM, "name buffer too small");
        memcpy(names + used, L.data() + i, name_len);
        names[used + name_len] = '\0';
        name_off[n] = used;
        used += (int64_t)name_len + 1;
        if ((size_t)n >= lines.size()) lines.emplace_back();
        lines[n].swap(L);
        ++n;
    }
    // ---- parallel: one row per line ------------------------------------------------------
    std::vector<psb_line_out> outs((size_t)n);
    auto work = [&](int64_t lo, int64_t hi) {
        for (int64_t v = lo; v < hi; ++v)
            parse_line(r, lines[v].data(), lines[v].size(), bits + v * words_per_row,
                       missing ? missing + v * words_per_row : nullptr, words_per_row, &outs[v]);
    };
    const int T = (int)std::min<int64_t>(r->n_threads, std::max<int64_t>(1, n / 16));
    if (T <= 1) {
        work(0, n);
    } else {
        std::vector<std::thread> pool;
        for (int t = 1; t < T; ++t) pool.emplace_back(work, n * t / T, n * (t + 1) / T);
        work(0, n / T);
        for (auto &th : pool) th.join();
    }
    for (int64_t v = 0; v < n; ++v) {
        if (outs[v].rc != PSB_OK) {
            psb_set_error("%s", outs[v].err);
            return outs[v].rc;
        }
        info[v] = outs[v].flags;
        if ((outs[v].flags & 1) && any_missing) *any_missing = 1;
    }
    *n_read = n;
    return PSB_OK;
}
