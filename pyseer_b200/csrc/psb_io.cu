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

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <string_view>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <vector>

#include "psb_internal.cuh"
#include "psb_pgz.h"

struct psb_reader {
    gzFile fh = nullptr;
    psb_pgz *pgz = nullptr;                 // plain gzip of some size: inflated on the parser threads (psb_pgz.cu)
    bool pgz_err = false;
    FILE *raw = nullptr;                    // plain text, or BGZF input inflated block-parallel (bgzf_fill)
    bool bgzf = false;
    std::vector<unsigned char> bgzf_in;
    bool bgzf_err = false;
    int var_type = 0;                       // 0 = k-mers, 1 = Rtab, 2 = VCF
    int n_samples = 0;
    std::vector<std::string> names;         // owns the keys of `index`
    std::unordered_map<std::string_view, int> index;
    std::vector<int> rtab_col;              // Rtab column -> sample index or -1
    std::vector<char> buf;                  // read buffer
    size_t pos = 0, len = 0;
    bool eof = false;
    std::string line;
    std::string pending;                    // a line whose name did not fit the caller's buffer:
    bool have_pending = false;              //   first line of the next psb_reader_next
    bool drained = false;                   // reader_getline has returned false
    std::vector<std::string> lines;         // lines of the batch being parsed
    int n_threads = 1;
    // VCF: name under construction, and contig / position / REF length of the records of the last batch
    std::string vcf_name;
    std::vector<std::string> vcf_contig;
    std::vector<int64_t> vcf_pos;
    std::vector<int32_t> vcf_reflen;
    // text mode (psb_reader_next_text): bytes read past the end of the last batch, file offset of the
    // plain-text reads, mean line length seen so far
    std::vector<char> carry;
    int64_t file_off = 0;
    double text_bytes_seen = 0.0, text_lines_seen = 0.0;
    bool text_eof = false;
};

// ---- BGZF (bgzip) input: independent deflate blocks of <= 64 KiB, inflated on the parser threads ----
// A bgzip file is a series of gzip members whose extra field carries the member's size ('B','C'
// subfield, SAM specification 4.1).  Members do not depend on each other, so a group of them is
// inflated in parallel -- plain gzip is one serial deflate stream (18.8 k variants/s at N = 5000
// however many threads parse); bgzip'ed k-mer files decompress at (threads) x that rate.
static bool bgzf_probe(FILE *f) {
    unsigned char h[18];
    const size_t got = fread(h, 1, sizeof(h), f);
    rewind(f);
    return got == sizeof(h) && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[10] == 6 && h[11] == 0 &&
           h[12] == 'B' && h[13] == 'C' && h[14] == 2 && h[15] == 0;
}

struct bgzf_block {
    size_t in_off, in_len;      // deflate payload inside r->bgzf_in
    size_t out_off, out_len;    // where its text lands in r->buf
};

static int bgzf_inflate_one(const unsigned char *in, size_t in_len, unsigned char *out, size_t out_len) {
    // the library's own inflate (psb_pgz.cu; ~1.5x zlib on k-mer text), zlib when it declines
    static const bool own = !(getenv("PSB_PGZ") && atoi(getenv("PSB_PGZ")) == 0);
    if (own && psb_pgz_inflate_exact(in, in_len, out, out_len) == 0) return 0;
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return -1;
    zs.next_in = const_cast<unsigned char *>(in);
    zs.avail_in = (uInt)in_len;
    zs.next_out = out;
    zs.avail_out = (uInt)out_len;
    const int rc = inflate(&zs, Z_FINISH);
    const bool ok = rc == Z_STREAM_END && zs.total_out == out_len;
    inflateEnd(&zs);
    return ok ? 0 : -1;
}

// next group of blocks -> r->buf; false at end of file (or on a corrupt block: r->bgzf_err set)
static bool bgzf_fill(psb_reader *r) {
    const size_t group_bytes = 8u << 20;                 // compressed bytes per group
    std::vector<unsigned char> &in = r->bgzf_in;
    in.clear();
    std::vector<bgzf_block> blocks;
    size_t out_total = 0;
    while (in.size() < group_bytes) {
        unsigned char h[18];
        const size_t got = fread(h, 1, sizeof(h), r->raw);
        if (got == 0) break;
        if (got != sizeof(h) || h[0] != 0x1f || h[1] != 0x8b || !(h[3] & 4) || h[12] != 'B' || h[13] != 'C') {
            r->bgzf_err = true;
            return false;
        }
        const size_t bsize = (size_t)h[16] + ((size_t)h[17] << 8) + 1;       // whole member
        const size_t xlen = (size_t)h[10] + ((size_t)h[11] << 8);
        const size_t head = 12 + xlen;
        if (bsize < head + 8) { r->bgzf_err = true; return false; }
        const size_t rest = bsize - sizeof(h);
        const size_t at = in.size();
        in.resize(at + rest);
        if (fread(in.data() + at, 1, rest, r->raw) != rest) { r->bgzf_err = true; return false; }
        const size_t payload_off = at + (head - sizeof(h));
        const size_t payload_len = bsize - head - 8;
        const unsigned char *tr = in.data() + at + rest - 8;
        const size_t isize = (size_t)tr[4] | ((size_t)tr[5] << 8) | ((size_t)tr[6] << 16) | ((size_t)tr[7] << 24);
        if (isize == 0) continue;                                            // the empty end-of-file member
        blocks.push_back({payload_off, payload_len, out_total, isize});
        out_total += isize;
    }
    if (blocks.empty()) return false;
    if (r->buf.size() < out_total) r->buf.resize(out_total);
    std::atomic<int> next(0), bad(0);
    auto work = [&]() {
        for (;;) {
            const int b = next.fetch_add(1);
            if (b >= (int)blocks.size()) return;
            const bgzf_block &k = blocks[b];
            if (bgzf_inflate_one(in.data() + k.in_off, k.in_len, (unsigned char *)r->buf.data() + k.out_off, k.out_len))
                bad.store(1);
        }
    };
    const int T = std::max(1, std::min<int>(r->n_threads, (int)blocks.size()));
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    if (bad.load()) { r->bgzf_err = true; return false; }
    r->pos = 0;
    r->len = out_total;
    return true;
}

static bool reader_fill(psb_reader *r) {
    if (r->eof) return false;
    if (r->raw && r->bgzf) {
        if (!bgzf_fill(r)) {
            r->eof = true;
            r->pos = r->len = 0;
            return false;
        }
        return true;
    }
    if (r->raw) {                       // plain text: straight from the file
        const size_t got = fread(r->buf.data(), 1, r->buf.size(), r->raw);
        if (got == 0) {
            r->eof = true;
            r->pos = r->len = 0;
            return false;
        }
        r->pos = 0;
        r->len = got;
        return true;
    }
    int64_t got;
    if (r->pgz) {
        got = psb_pgz_read(r->pgz, r->buf.data(), (int64_t)r->buf.size());
        if (got < 0) r->pgz_err = true;
    } else {
        got = gzread(r->fh, r->buf.data(), (unsigned)r->buf.size());
    }
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
    PSB_REQUIRE(var_type >= 0 && var_type <= 2, PSB_ERR_ARG, "var_type must be 0 (k-mers), 1 (Rtab) or 2 (VCF)");
    PSB_REQUIRE(n_samples > 0, PSB_ERR_ARG, "no samples");
    *out = nullptr;
    FILE *raw = fopen(path, "rb");
    PSB_REQUIRE(raw, PSB_ERR_ARG, "cannot open %s", path);
    const bool bgzf = bgzf_probe(raw) && !(getenv("PSB_BGZF") && atoi(getenv("PSB_BGZF")) == 0);
    unsigned char magic[2] = {0, 0};
    const bool gz = fread(magic, 1, 2, raw) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    rewind(raw);
    gzFile fh = nullptr;
    psb_pgz *pgz = nullptr;
    if (gz && !bgzf) {
        // files of some size go through the parallel inflater (PSB_PGZ=0: always zlib; PSB_PGZ_MIN /
        // PSB_PGZ_CHUNK: size threshold and compressed bytes per work item, for the tests)
        fseek(raw, 0, SEEK_END);
        const long fsize = ftell(raw);
        fclose(raw);
        raw = nullptr;
        const long pgz_min = getenv("PSB_PGZ_MIN") ? atol(getenv("PSB_PGZ_MIN")) : (4l << 20);
        if (!(getenv("PSB_PGZ") && atoi(getenv("PSB_PGZ")) == 0) && fsize >= pgz_min)
            pgz = psb_pgz_open(path, 1, getenv("PSB_PGZ_CHUNK") ? (size_t)atol(getenv("PSB_PGZ_CHUNK")) : 0);
        if (!pgz) {
            fh = gzopen(path, "rb");
            PSB_REQUIRE(fh, PSB_ERR_ARG, "cannot open %s", path);
            gzbuffer(fh, 1 << 20);
        }
    }
    psb_reader *r = new psb_reader();
    r->fh = fh;
    r->pgz = pgz;
    r->raw = raw;
    r->bgzf = bgzf;
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
            if (fh) gzclose(fh);
            if (pgz) psb_pgz_close(pgz);
            if (raw) fclose(raw);
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
    if (var_type == 2) {
        // skip the meta lines; the #CHROM line names the sample columns (9 fixed fields first)
        bool found = false;
        while (reader_getline(r)) {
            const std::string &h = r->line;
            if (h.rfind("#CHROM", 0) != 0) continue;
            size_t i = 0, n = h.size();
            while (n > 0 && (h[n - 1] == '\r' || h[n - 1] == '\n')) --n;
            int field = 0;
            while (i <= n) {
                size_t j = i;
                while (j < n && h[j] != '\t') ++j;
                if (field >= 9) {
                    auto it = r->index.find(std::string_view(h.data() + i, j - i));
                    r->rtab_col.push_back(it == r->index.end() ? -1 : it->second);
                }
                ++field;
                if (j >= n) break;
                i = j + 1;
            }
            found = true;
            break;
        }
        if (!found) {
            if (fh) gzclose(fh);
            if (pgz) psb_pgz_close(pgz);
            if (raw) fclose(raw);
            delete r;
            psb_set_error("%s: no #CHROM header line found; is this a VCF file?", path);
            return PSB_ERR_ARG;
        }
    }
    *out = r;
    return PSB_OK;
}

extern "C" int psb_reader_close(psb_reader *r) {
    if (!r) return PSB_OK;
    if (r->fh) gzclose(r->fh);
    if (r->pgz) psb_pgz_close(r->pgz);
    if (r->raw) fclose(r->raw);
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
    } else if (r->var_type == 2) {
        // VCF record, input.read_vcf_var (input.py:457-502), dominant encoding.  Fields: CHROM POS ID
        // REF ALT QUAL FILTER INFO FORMAT samples...  flags bit 2: more than one ALT allele, bit 3:
        // FILTER neither empty nor PASS -- both skipped by the reference (the row stays empty).
        const char *fld[9];
        size_t fl[9];
        size_t k = 0;
        for (int c = 0; c < 9; ++c) {
            size_t e = k;
            while (e < len && L[e] != '\t') ++e;
            fld[c] = L + k;
            fl[c] = e - k;
            LINE_REQUIRE(e < len || c == 8, PSB_ERR_ARG, "VCF record with fewer than 9 fields");
            k = e + 1;
        }
        // ALT: '.' = none; more than one allele -> skipped
        if (memchr(fld[4], ',', fl[4])) flags |= 4;
        // FILTER: '.' or empty passes; otherwise one of the ';'-separated names must be PASS
        if (!(fl[6] == 0 || (fl[6] == 1 && fld[6][0] == '.'))) {
            bool pass = false;
            size_t a = 0;
            while (a <= fl[6]) {
                size_t b = a;
                while (b < fl[6] && fld[6][b] != ';') ++b;
                if (b - a == 4 && memcmp(fld[6] + a, "PASS", 4) == 0) pass = true;
                a = b + 1;
            }
            if (!pass && !(flags & 4)) flags |= 8;
        }
        if (!(flags & 12)) {
            // position of GT among the ':'-separated FORMAT keys (-1: every genotype is missing)
            int gi = -1, key = 0;
            for (size_t a = 0; a <= fl[8];) {
                size_t b = a;
                while (b < fl[8] && fld[8][b] != ':') ++b;
                if (b - a == 2 && fld[8][a] == 'G' && fld[8][a + 1] == 'T') { gi = key; break; }
                ++key;
                a = b + 1;
            }
            size_t col = 0;
            const size_t ncol = r->rtab_col.size();
            while (k <= len && col < ncol) {
                size_t e = k;
                while (e < len && L[e] != '\t') ++e;
                const int s = r->rtab_col[col];
                if (s >= 0) {
                    int st = 0;                       // 0 absent, 1 carrier, 2 missing
                    if (gi < 0) {
                        st = 2;
                    } else {
                        // the gi-th ':' field of the cell
                        size_t a = k;
                        for (int q = 0; q < gi && a < e; ++q) {
                            while (a < e && L[a] != ':') ++a;
                            if (a < e) ++a;
                        }
                        size_t b = a;
                        while (b < e && L[b] != ':') ++b;
                        // haplotypes separated by '/' or '|'
                        size_t h0 = a;
                        for (;;) {
                            size_t h1 = h0;
                            while (h1 < b && L[h1] != '/' && L[h1] != '|') ++h1;
                            const size_t hl = h1 - h0;
                            // an empty token (cell shorter than FORMAT, empty cell, trailing '/'):
                            // pysam reports None, which read_vcf_var treats like '.' (input.py:486-488)
                            if (hl == 0 || (hl == 1 && L[h0] == '.')) {
                                if (st == 0) st = 2;
                            } else if (!(hl == 1 && L[h0] == '0')) {
                                st = 1;
                                break;
                            } else if (st == 2) {
                                st = 0;
                            }
                            if (h1 >= b) break;
                            h0 = h1 + 1;
                        }
                    }
                    if (st == 1) { row[s >> 5] |= 1u << (s & 31); seen = true; }
                    if (st == 2) {
                        flags |= 1;
                        seen = true;
                        if (mrow) mrow[s >> 5] |= 1u << (s & 31);
                    }
                }
                ++col;
                k = e + 1;
            }
            LINE_REQUIRE(!(flags & 1) || mrow, PSB_ERR_ARG,
                         "record has missing genotypes but no missing buffer was given");
        } else {
            seen = true;          // no "No observations" message for skipped records
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
    if (r->pgz) psb_pgz_set_threads(r->pgz, r->n_threads);
    return PSB_OK;
}

// Reads up to max_variants rows.  bits / missing: max_variants x words_per_row (zeroed here);
// names: concatenated NUL-terminated variant names (names_cap bytes); name_off[v] = offset of
// name v; info[v]: bit 0 = row has missing genotypes, bit 1 = no observation in the selected
// samples (the reference writes "No observations of ..." to stderr, input.py:447-448).
// *n_read = rows produced (0 at end of file).  A batch also ends when the next name does not fit
// what is left of `names`: that line is kept and opens the next call (psb_reader_at_eof tells the
// two apart).  Returns PSB_ERR_NOMEM (line kept as well) when names_cap is too
// small for a single name, PSB_ERR_ARG on a malformed row.
extern "C" int psb_reader_next(psb_reader *r, int64_t max_variants, uint32_t *bits, uint32_t *missing,
                               int32_t words_per_row, char *names, int64_t names_cap,
                               int64_t *name_off, int32_t *info, int64_t *n_read, int32_t *any_missing) {
    PSB_REQUIRE(r && bits && names && name_off && info && n_read, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(words_per_row * 32 >= r->n_samples, PSB_ERR_ARG, "words_per_row too small");
    *n_read = 0;
    if (any_missing) *any_missing = 0;
    // ---- serial: lines of the batch and their names -----------------------------------
    r->vcf_contig.clear();
    r->vcf_pos.clear();
    r->vcf_reflen.clear();
    std::vector<std::string> &lines = r->lines;
    int64_t n = 0, used = 0;
    while (n < max_variants) {
        if (r->have_pending) {
            r->line.swap(r->pending);
            r->have_pending = false;
        } else if (!reader_getline(r)) {
            PSB_REQUIRE(!r->bgzf_err, PSB_ERR_ARG, "corrupt BGZF block in the variant file");
            PSB_REQUIRE(!r->pgz_err, PSB_ERR_ARG, "variant file: %s", psb_pgz_error(r->pgz));
            r->drained = true;
            break;
        }
        std::string &L = r->line;
        size_t len = L.size();
        while (len > 0 && (L[len - 1] == '\r' || L[len - 1] == ' ' || L[len - 1] == '\t')) --len;
        if (len == 0) continue;
        L.resize(len);
        size_t i = 0, j = 0;
        size_t vt0 = 0, ve0 = 0, vt1 = 0, vreflen = 0;
        if (r->var_type == 0) {      // name = first whitespace-delimited token
            while (i < len && (L[i] == ' ' || L[i] == '\t')) ++i;
            j = i;
            while (j < len && L[j] != ' ' && L[j] != '\t') ++j;
        } else if (r->var_type == 2) {
            // '_'.join([contig, pos, ref] + alts)  (input.py:471-472); also record contig / pos / ref
            // length for the burden-region lookup (psb_reader_vcf_info)
            if (L[0] == '#') continue;
            size_t t[5] = {0, 0, 0, 0, 0}, e[5] = {0, 0, 0, 0, 0}, k = 0;
            bool ok = true;
            for (int c = 0; c < 5; ++c) {
                size_t q = k;
                while (q < len && L[q] != '\t') ++q;
                t[c] = k;
                e[c] = q;
                if (q >= len) { ok = false; break; }
                k = q + 1;
            }
            PSB_REQUIRE(ok, PSB_ERR_ARG, "VCF record with fewer than 9 fields");
            r->vcf_name.assign(L.data() + t[0], e[0] - t[0]);
            r->vcf_name.push_back('_');
            r->vcf_name.append(L.data() + t[1], e[1] - t[1]);
            r->vcf_name.push_back('_');
            r->vcf_name.append(L.data() + t[3], e[3] - t[3]);
            if (!(e[4] - t[4] == 1 && L[t[4]] == '.')) {
                r->vcf_name.push_back('_');
                for (size_t q = t[4]; q < e[4]; ++q) r->vcf_name.push_back(L[q] == ',' ? '_' : L[q]);
            }
            vt0 = t[0]; ve0 = e[0]; vt1 = t[1]; vreflen = e[3] - t[3];
        } else {                     // name = first tab-delimited field
            while (j < len && L[j] != '\t') ++j;
        }
        const char *name_ptr = r->var_type == 2 ? r->vcf_name.data() : L.data() + i;
        const size_t name_len = r->var_type == 2 ? r->vcf_name.size() : j - i;
        if ((int64_t)name_len + 1 > names_cap - used) {
            // the name does not fit what is left of the caller's buffer: the line opens the next
            // call (nothing is lost; with n == 0 the caller has to come back with a larger buffer)
            r->pending.swap(L);
            r->have_pending = true;
            PSB_REQUIRE(n > 0, PSB_ERR_NOMEM, "name buffer too small for a name of %zu bytes", name_len);
            break;
        }
        if (r->var_type == 2) {
            r->vcf_contig.emplace_back(L.data() + vt0, ve0 - vt0);
            r->vcf_pos.push_back(strtoll(L.c_str() + vt1, nullptr, 10));
            r->vcf_reflen.push_back((int32_t)vreflen);
        }
        memcpy(names + used, name_ptr, name_len);
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

// ---------------------------------------------------------------------------------------
// Text mode: the decompressed k-mer text itself, cut into lines, for the device parser
// (psb_submit_text, psb_text.cu).  The host no longer tokenises: it reads (plain: pread on --cpu
// threads; BGZF: block-parallel inflate; gzip: zlib) STRAIGHT into the caller's page-locked buffer,
// finds the newlines (on the same threads) and copies the variant names out.  Rules of the row
// reader kept: trailing '\r', ' ', '\t' are trimmed, empty lines skipped, the name is the first
// blank-delimited token.
// ---------------------------------------------------------------------------------------
static int64_t text_read_more(psb_reader *r, char *dst, int64_t want) {
    if (r->text_eof || want <= 0) return 0;
    int64_t got = 0;
    if (r->raw && !r->bgzf) {
        const int fd = fileno(r->raw);
        const int T = (int)std::max<int64_t>(1, std::min<int64_t>(r->n_threads, want >> 22));
        std::vector<int64_t> part(T, 0);
        auto work = [&](int t) {
            const int64_t lo = want * t / T, hi = want * (t + 1) / T;
            int64_t done = 0;
            while (lo + done < hi) {
                const ssize_t k = pread(fd, dst + lo + done, (size_t)(hi - lo - done), (off_t)(r->file_off + lo + done));
                if (k <= 0) break;
                done += k;
            }
            part[t] = done;
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto &th : pool) th.join();
        // a short slice can only be the one that met the end of the file
        for (int t = 0; t < T; ++t) {
            got += part[t];
            if (part[t] < want * (t + 1) / T - want * t / T) break;
        }
        r->file_off += got;
        if (got < want) r->text_eof = true;
        return got;
    }
    if (r->raw && r->bgzf) {
        // whole blocks whose text fits `want`, inflated in parallel into their final place
        std::vector<unsigned char> &in = r->bgzf_in;
        in.clear();
        std::vector<bgzf_block> blocks;
        int64_t out_total = 0;
        for (;;) {
            unsigned char h[18];
            const long at_file = ftell(r->raw);
            const size_t n = fread(h, 1, sizeof(h), r->raw);
            if (n == 0) { r->text_eof = true; break; }
            if (n != sizeof(h) || h[0] != 0x1f || h[1] != 0x8b || !(h[3] & 4) || h[12] != 'B' || h[13] != 'C') {
                r->bgzf_err = true;
                return -1;
            }
            const size_t bsize = (size_t)h[16] + ((size_t)h[17] << 8) + 1;
            const size_t xlen = (size_t)h[10] + ((size_t)h[11] << 8);
            const size_t head = 12 + xlen;
            if (bsize < head + 8) { r->bgzf_err = true; return -1; }
            const size_t rest = bsize - sizeof(h);
            const size_t at = in.size();
            in.resize(at + rest);
            if (fread(in.data() + at, 1, rest, r->raw) != rest) { r->bgzf_err = true; return -1; }
            const unsigned char *tr = in.data() + at + rest - 8;
            const size_t isize = (size_t)tr[4] | ((size_t)tr[5] << 8) | ((size_t)tr[6] << 16) | ((size_t)tr[7] << 24);
            if (isize == 0) { in.resize(at); continue; }
            if (out_total + (int64_t)isize > want) {
                // does not fit any more: this block opens the next call
                in.resize(at);
                fseek(r->raw, at_file, SEEK_SET);
                break;
            }
            blocks.push_back({at + (head - sizeof(h)), bsize - head - 8, (size_t)out_total, isize});
            out_total += (int64_t)isize;
        }
        if (blocks.empty()) return 0;
        std::atomic<int> next(0), bad(0);
        auto work = [&]() {
            for (;;) {
                const int b = next.fetch_add(1);
                if (b >= (int)blocks.size()) return;
                const bgzf_block &k = blocks[b];
                if (bgzf_inflate_one(in.data() + k.in_off, k.in_len, (unsigned char *)dst + k.out_off, k.out_len))
                    bad.store(1);
            }
        };
        const int T = std::max(1, std::min<int>(r->n_threads, (int)blocks.size()));
        std::vector<std::thread> pool;
        for (int t = 1; t < T; ++t) pool.emplace_back(work);
        work();
        for (auto &th : pool) th.join();
        if (bad.load()) { r->bgzf_err = true; return -1; }
        return out_total;
    }
    if (r->pgz) {
        got = psb_pgz_read(r->pgz, dst, want);
        if (got < 0) { r->pgz_err = true; return -1; }
        if (got < want) r->text_eof = true;
        return got;
    }
    while (got < want) {
        const int k = gzread(r->fh, dst + got, (unsigned)std::min<int64_t>(want - got, 1 << 30));
        if (k <= 0) { r->text_eof = true; break; }
        got += k;
    }
    return got;
}

// positions of the '\n' bytes of dst[lo, hi), on up to n_threads threads
static void text_find_newlines(const char *dst, int64_t lo, int64_t hi, int n_threads, std::vector<int64_t> &nl) {
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, (hi - lo) >> 22));
    std::vector<std::vector<int64_t>> part(T);
    auto work = [&](int t) {
        const int64_t a = lo + (hi - lo) * t / T, b = lo + (hi - lo) * (t + 1) / T;
        const char *p = dst + a, *e = dst + b;
        while (p < e) {
            const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
            if (!q) break;
            part[t].push_back(q - dst);
            p = q + 1;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    for (int t = 0; t < T; ++t) nl.insert(nl.end(), part[t].begin(), part[t].end());
}

// Fills dst (dst_cap bytes, ideally page-locked) with the text of up to max_lines k-mer lines.
// line_start[v] / line_len[v]: line v inside dst, trimmed; names / name_off as psb_reader_next.
// *n_read lines use the first *n_bytes bytes of dst.  Fewer than max_lines lines come back at the
// end of the file (psb_reader_at_eof) or when dst / names cannot take more: then *n_read is a
// multiple of line_multiple (the caller's block size) so that block boundaries do not move, and
// the rest opens the next call.  PSB_ERR_NOMEM: not even line_multiple lines fit.
extern "C" int psb_reader_next_text(psb_reader *r, int64_t max_lines, int64_t line_multiple, char *dst,
                                    int64_t dst_cap, int64_t *line_start, int32_t *line_len, char *names,
                                    int64_t names_cap, int64_t *name_off, int64_t *n_read, int64_t *n_bytes) {
    PSB_REQUIRE(r && dst && line_start && line_len && names && name_off && n_read && n_bytes, PSB_ERR_ARG,
                "NULL argument");
    PSB_REQUIRE(r->var_type == 0, PSB_ERR_ARG, "text mode reads k-mer files only");
    PSB_REQUIRE(!r->have_pending && r->pos == r->len, PSB_ERR_STATE, "text mode cannot follow psb_reader_next");
    PSB_REQUIRE(max_lines > 0 && dst_cap > 0, PSB_ERR_ARG, "empty buffer");
    if (line_multiple < 1) line_multiple = 1;
    *n_read = 0;
    *n_bytes = 0;
    int64_t have = (int64_t)r->carry.size();
    PSB_REQUIRE(have <= dst_cap, PSB_ERR_NOMEM, "text buffer of %lld bytes too small", (long long)dst_cap);
    if (have) memcpy(dst, r->carry.data(), (size_t)have);
    r->carry.clear();
    int64_t n = 0, used = 0, scanned = 0, line_at = 0, consumed = 0;
    bool full = false, bad_line = false;
    std::vector<int64_t> nl;
    auto take_line = [&](int64_t a, int64_t b) -> bool {      // dst[a, b) without its newline
        int64_t e = b;
        while (e > a && (dst[e - 1] == '\r' || dst[e - 1] == ' ' || dst[e - 1] == '\t')) --e;
        if (e == a) return true;                              // empty line: skipped
        int64_t i = a;
        while (i < e && (dst[i] == ' ' || dst[i] == '\t')) ++i;
        int64_t j = i;
        while (j < e && dst[j] != ' ' && dst[j] != '\t') ++j;
        if (j - i + 1 > names_cap - used || e - a > 0x7fffffffll) return false;
        // the row reader's check (parse_line): the sample list starts after the first '|' (found within
        // the first bytes of a well-formed line)
        if (!memchr(dst + a, '|', (size_t)(e - a))) bad_line = true;
        memcpy(names + used, dst + i, (size_t)(j - i));
        names[used + (j - i)] = '\0';
        name_off[n] = used;
        used += j - i + 1;
        line_start[n] = a;
        line_len[n] = (int32_t)(e - a);
        ++n;
        return true;
    };
    for (;;) {
        // lines complete in dst[scanned, have)
        nl.clear();
        text_find_newlines(dst, scanned, have, r->n_threads, nl);
        scanned = have;
        for (size_t k = 0; k < nl.size() && n < max_lines && !full; ++k) {
            if (!take_line(line_at, nl[k])) { full = true; break; }
            line_at = nl[k] + 1;
            consumed = line_at;
        }
        if (n >= max_lines || full) break;
        if (r->text_eof) {
            // the last line of the file may come without a newline
            if (line_at < have) {
                if (take_line(line_at, have)) {
                    line_at = have;
                    consumed = have;
                } else {
                    full = true;
                }
            }
            break;
        }
        // more text: what the missing lines should take at the mean line length seen so far
        const double mean = r->text_lines_seen > 0 ? r->text_bytes_seen / r->text_lines_seen
                                                   : (n > 0 ? (double)consumed / (double)n : 0.0);
        int64_t want = mean > 0 ? (int64_t)((double)(max_lines - n) * mean * 1.01) - (have - line_at) + (256 << 10)
                                : (int64_t)8 << 20;
        if (want < (1 << 20)) want = 1 << 20;
        if (want > dst_cap - have) want = dst_cap - have;
        if (want <= 0 || (r->bgzf && want < (1 << 16))) { full = true; break; }
        const int64_t got = text_read_more(r, dst + have, want);
        PSB_REQUIRE(got >= 0 || !r->pgz_err, PSB_ERR_ARG, "variant file: %s", psb_pgz_error(r->pgz));
        PSB_REQUIRE(got >= 0, PSB_ERR_ARG, "corrupt BGZF block in the variant file");
        if (got == 0 && !r->text_eof) { full = true; break; }     // BGZF: the next block does not fit
        have += got;
    }
    PSB_REQUIRE(!bad_line, PSB_ERR_ARG, "k-mer line without '|' separator");
    if (full && n < max_lines) {
        // cut short by a buffer: keep whole blocks of the caller's block size
        const int64_t keep = n / line_multiple * line_multiple;
        if (keep <= 0) {
            // nothing is lost: the text read so far opens the next call (with larger buffers)
            r->carry.assign(dst, dst + have);
            psb_set_error("text / name buffers too small for %lld lines", (long long)line_multiple);
            return PSB_ERR_NOMEM;
        }
        if (keep < n) {
            n = keep;
            consumed = line_start[keep];      // the first dropped line opens the next call
        }
    }
    r->carry.assign(dst + consumed, dst + have);
    r->text_bytes_seen += (double)consumed;
    r->text_lines_seen += (double)n;
    r->drained = r->text_eof && r->carry.empty();
    *n_read = n;
    *n_bytes = consumed;
    return PSB_OK;
}

// *at_eof != 0 once the file has been read to its end and no line is held back: tells a short batch
// that ended at the end of the file from one that ended because the names buffer was full.
extern "C" int psb_reader_at_eof(psb_reader *r, int32_t *at_eof) {
    PSB_REQUIRE(r && at_eof, PSB_ERR_ARG, "NULL argument");
    *at_eof = (r->drained && !r->have_pending) ? 1 : 0;
    return PSB_OK;
}

// ---------------------------------------------------------------------------------------
// TSV lines of a whole result table: utils.format_output (pyseer/utils.py:39-105) applied to the
// Seer / LMM tuples the result loop of main() builds (__main__.py:547-568, 783-803), without the
// per-variant Python objects.  Host-only code.
// ---------------------------------------------------------------------------------------
// '%.2E' % Decimal(x) if isfinite(x) else ''.  Six of these per line made snprintf the cost of the
// formatter; the fast path scales |x| into [100, 1000) with a table of powers of ten (relative error of
// the scaled value < 4e-16, i.e. < 4e-13 absolute) and rounds it -- unless its fraction lies within 1e-6
// of one half, where the exact binary value decides (ties to even) and snprintf is asked.
static const double *fmt_p10() {
    static double tab[301];
    static bool ready = false;
    if (!ready) {
        for (int k = 0; k <= 300; ++k) {
            char lit[16];
            snprintf(lit, sizeof(lit), "1e%d", k);
            tab[k] = strtod(lit, nullptr);              // correctly rounded
        }
        ready = true;
    }
    return tab;
}
static const double *const k_p10 = fmt_p10();

static inline char *fmt_num(char *p, double x) {
    if (!isfinite(x)) return p;
    const double a = fabs(x);
    if (a >= 1e-290 && a <= 1e290) {
        int e = (int)floor(log10(a));
        double s = 0.0;
        for (int tries = 0; tries < 3; ++tries) {
            s = e >= 2 ? a / k_p10[e - 2] : a * k_p10[2 - e];
            if (s < 100.0) --e;
            else if (s >= 1000.0) ++e;
            else break;
        }
        if (s >= 100.0 && s < 1000.0) {
            const double fl = floor(s), fr = s - fl;
            if (fabs(fr - 0.5) > 1e-6) {
                int n = (int)fl + (fr > 0.5 ? 1 : 0);
                if (n == 1000) {
                    n = 100;
                    ++e;
                }
                if (signbit(x)) *p++ = '-';
                *p++ = (char)('0' + n / 100);
                *p++ = '.';
                *p++ = (char)('0' + (n / 10) % 10);
                *p++ = (char)('0' + n % 10);
                *p++ = 'E';
                int ae = e;
                if (ae < 0) {
                    *p++ = '-';
                    ae = -ae;
                } else {
                    *p++ = '+';
                }
                if (ae >= 100) {
                    *p++ = (char)('0' + ae / 100);
                    ae %= 100;
                }
                *p++ = (char)('0' + ae / 10);
                *p++ = (char)('0' + ae % 10);
                return p;
            }
        }
    }
    p += snprintf(p, 32, "%.2E", x);
    return p;
}

static const struct { uint32_t bit; const char *text; } k_notes[] = {
    {PSB_F_AF_FILTER, "af-filter"},
    {PSB_F_PREFILTER_FAILED, "pre-filtering-failed"},
    {PSB_F_BAD_CHISQ, "bad-chisq"},
    {PSB_F_HIGH_BSE, "high-bse"},
    {PSB_F_PERFECT_SEP, "perfectly-separable-data"},
    {PSB_F_MATRIX_INV, "matrix-inversion-error"},
    {PSB_F_FIRTH_FAIL, "firth-fail"},
    {PSB_F_MISSING_DATA, "missing-data-error"},
    {PSB_F_LRT_FAILED, "lrt-filtering-failed"},
};

// Formats rows [b0, b1) (b0 a multiple of block_size) into `dst`; counts[0..2] += pre-filtered,
// tested, printed.
struct fmt_lineage {                 // --lineage column: name of lineage[v], "NA" for a negative index
    const int32_t *idx = nullptr;    // nullptr: no such column
    const char *names = nullptr;
    const int64_t *off = nullptr;
    int32_t n = 0;
};

static void format_range(int model, int64_t b0, int64_t b1, const char *names, const int64_t *name_off,
                         const psb_results *cols, int n_betas, int block_size, int print_filtered,
                         const fmt_lineage &lin, std::string &dst, int64_t counts[3]) {
    char num[40];
    for (int64_t c0 = b0; c0 < b1; c0 += block_size) {
        const int64_t c1 = std::min<int64_t>(b1, c0 + block_size);
        for (int pass = 0; pass < (model == 1 ? 2 : 1); ++pass) {
            for (int64_t v = c0; v < c1; ++v) {
                const uint32_t f = cols->flags[v];
                const bool pre = (f & PSB_F_PREFILTER) != 0;
                if (model == 1 && pre != (pass == 0)) continue;
                if (pre) {
                    counts[0]++;
                    if (!print_filtered) continue;
                } else {
                    counts[1]++;
                    if ((f & PSB_F_FILTER) && !print_filtered) continue;
                }
                auto put = [&](double x) { dst.append(num, (size_t)(fmt_num(num, x) - num)); };
                dst.append(names + name_off[v]);
                dst.push_back('\t');
                put(cols->af[v]);
                dst.push_back('\t');
                put(cols->prep[v]);
                dst.push_back('\t');
                // fields the tuple leaves at NaN stay empty
                const bool fitted = !pre && !(f & (PSB_F_FIRTH_FAIL | PSB_F_MISSING_DATA)) &&
                                    !(model == 1 && (f & PSB_F_FILTER));
                const bool has_p = !pre && !(model == 0 && (f & (PSB_F_FIRTH_FAIL | PSB_F_MISSING_DATA)));
                if (has_p) put(cols->pvalue[v]);
                dst.push_back('\t');
                if (fitted) put(cols->beta[v]);
                dst.push_back('\t');
                if (fitted) put(cols->bse[v]);
                dst.push_back('\t');
                if (fitted) put(cols->extra[v]);
                if (model == 0 && fitted && n_betas > 0) {
                    for (int c = 0; c < n_betas; ++c) {
                        dst.push_back('\t');
                        put(cols->betas[v * n_betas + c]);
                    }
                }
                if (lin.idx) {                                   // utils.py:93-97
                    dst.push_back('\t');
                    const int32_t l = lin.idx[v];
                    if (l >= 0 && l < lin.n) dst.append(lin.names + lin.off[l]);
                    else dst.append("NA");
                }
                dst.push_back('\t');
                bool first = true;
                for (const auto &nt : k_notes)
                    if (f & nt.bit) {
                        if (!first) dst.push_back(',');
                        dst.append(nt.text);
                        first = false;
                    }
                dst.push_back('\n');
                counts[2]++;
            }
        }
    }
}

// model: 0 = fixed effects (Seer: ... bse, intercept, betas[n_betas]), 1 = LMM (... bse, variant_h2).
// names: NUL-terminated variant names back to back, name_off[v] their offsets.  cols: HOST pointers
// of the result table (psb_fetch).  Rows are visited in blocks of block_size; inside a block the LMM
// model emits the pre-filtered variants first (lmm.py:158-226), fixed effects keep input order.
// counts[0..2] += pre-filtered, tested, printed (__main__.py:549-565, 793-817).  n_threads > 1 formats
// ranges of whole blocks in parallel and concatenates them in order.  Returns PSB_ERR_NOMEM when
// out_cap is too small (nothing usable in `out` then).
static int format_rows_impl(int32_t model, int64_t n, const char *names, const int64_t *name_off,
                            const psb_results *cols, int32_t n_betas, int32_t block_size,
                            int32_t print_filtered, int32_t n_threads, const fmt_lineage &lin, char *out,
                            int64_t out_cap, int64_t *out_len, int64_t counts[3]) {
    PSB_REQUIRE(names && name_off && cols && out && out_len && counts, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(cols->af && cols->prep && cols->pvalue && cols->beta && cols->bse && cols->extra &&
                    cols->flags && (n_betas == 0 || cols->betas),
                PSB_ERR_ARG, "result columns missing");
    PSB_REQUIRE(model == 0 || model == 1, PSB_ERR_ARG, "model must be 0 (seer) or 1 (lmm)");
    if (block_size < 1) block_size = 1;
    const int64_t n_blocks = (n + block_size - 1) / block_size;
    int T = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
    T = (int)std::min<int64_t>(T, std::max<int64_t>(1, n / 2048));
    T = (int)std::min<int64_t>(T, std::max<int64_t>(1, n_blocks));
    std::vector<std::string> parts((size_t)T);
    std::vector<int64_t> cnt((size_t)T * 3, 0);
    auto work = [&](int t) {
        const int64_t b0 = (n_blocks * t / T) * block_size;
        const int64_t b1 = std::min<int64_t>(n, (n_blocks * (t + 1) / T) * block_size);
        parts[t].reserve((size_t)((b1 - b0) * (64 + 10 * (7 + n_betas) + (lin.idx ? 16 : 0))));
        format_range(model, b0, b1, names, name_off, cols, n_betas, block_size, print_filtered, lin, parts[t],
                     &cnt[(size_t)t * 3]);
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto &th : pool) th.join();
    }
    int64_t total = 0;
    for (const auto &s2 : parts) total += (int64_t)s2.size();
    PSB_REQUIRE(total <= out_cap, PSB_ERR_NOMEM, "output buffer too small (%lld bytes needed)", (long long)total);
    char *p = out;
    for (int t = 0; t < T; ++t) {
        memcpy(p, parts[t].data(), parts[t].size());
        p += parts[t].size();
        for (int k = 0; k < 3; ++k) counts[k] += cnt[(size_t)t * 3 + k];
    }
    *out_len = total;
    return PSB_OK;
}

extern "C" int psb_format_rows(int32_t model, int64_t n, const char *names, const int64_t *name_off,
                               const psb_results *cols, int32_t n_betas, int32_t block_size,
                               int32_t print_filtered, int32_t n_threads, char *out, int64_t out_cap,
                               int64_t *out_len, int64_t counts[3]) {
    return format_rows_impl(model, n, names, name_off, cols, n_betas, block_size, print_filtered, n_threads,
                            fmt_lineage(), out, out_cap, out_len, counts);
}

// The same with the lineage column of --lineage runs (utils.py:93-97) between the coefficients and the
// notes: lineage[v] indexes the n_lineages NUL-terminated names in lineage_names (lineage_off their
// offsets); a negative index prints NA.
extern "C" int psb_format_rows_lineage(int32_t model, int64_t n, const char *names, const int64_t *name_off,
                                       const psb_results *cols, int32_t n_betas, int32_t block_size,
                                       int32_t print_filtered, int32_t n_threads, const int32_t *lineage,
                                       const char *lineage_names, const int64_t *lineage_off,
                                       int32_t n_lineages, char *out, int64_t out_cap, int64_t *out_len,
                                       int64_t counts[3]) {
    PSB_REQUIRE(lineage && lineage_names && lineage_off && n_lineages >= 0, PSB_ERR_ARG, "NULL lineage argument");
    fmt_lineage lin;
    lin.idx = lineage;
    lin.names = lineage_names;
    lin.off = lineage_off;
    lin.n = n_lineages;
    return format_rows_impl(model, n, names, name_off, cols, n_betas, block_size, print_filtered, n_threads, lin,
                            out, out_cap, out_len, counts);
}

// The similarity tool's output (pyseer/similarity.py:118-120: DataFrame(K, index, columns).to_csv(sep='\t')):
// header line of sample names after an empty index label, then one line per sample, entries printed as
// pandas prints the float64 counts ('1234.0').  pandas needs ~12 s for N = 5000; this is a memory-bound
// loop.  PSB_ERR_UNSUPPORTED when an entry is not a non-negative integer below 1e15 (the caller falls back
// to pandas; K = G G' never holds such values).  names: NUL-terminated, name_off their offsets.
extern "C" int psb_format_matrix(const double *K, int32_t n, const char *names, const int64_t *name_off,
                                 int32_t n_threads, char *out, int64_t out_cap, int64_t *out_len) {
    PSB_REQUIRE(K && names && name_off && out && out_len && n > 0, PSB_ERR_ARG, "NULL argument");
    int T = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
    T = std::min(T, std::max(1, n / 64));
    std::vector<std::string> parts((size_t)T);
    std::atomic<int> bad(0);
    auto work = [&](int t) {
        const int r0 = (int)((int64_t)n * t / T), r1 = (int)((int64_t)n * (t + 1) / T);
        std::string &dst = parts[t];
        dst.reserve((size_t)(r1 - r0) * ((size_t)n * 8 + 64));
        char num[32];
        for (int i = r0; i < r1 && !bad.load(std::memory_order_relaxed); ++i) {
            dst.append(names + name_off[i]);
            const double *row = K + (size_t)i * n;
            for (int j = 0; j < n; ++j) {
                const double x = row[j];
                if (!(x >= 0.0 && x < 1e15) || x != floor(x)) { bad.store(1); break; }
                unsigned long long v = (unsigned long long)x;
                char *e = num + sizeof(num), *p2 = e;
                do {
                    *--p2 = (char)('0' + v % 10);
                    v /= 10;
                } while (v);
                dst.push_back('\t');
                dst.append(p2, (size_t)(e - p2));
                dst.append(".0", 2);
            }
            dst.push_back('\n');
        }
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto &th : pool) th.join();
    }
    if (bad.load()) {
        psb_set_error("matrix entries are not small non-negative integers");
        return PSB_ERR_UNSUPPORTED;
    }
    std::string head;
    for (int j = 0; j < n; ++j) {
        head.push_back('\t');
        head.append(names + name_off[j]);
    }
    head.push_back('\n');
    int64_t total = (int64_t)head.size();
    for (const auto &s2 : parts) total += (int64_t)s2.size();
    PSB_REQUIRE(total <= out_cap, PSB_ERR_NOMEM, "output buffer too small (%lld bytes needed)", (long long)total);
    char *p = out;
    memcpy(p, head.data(), head.size());
    p += head.size();
    for (const auto &s2 : parts) {
        memcpy(p, s2.data(), s2.size());
        p += s2.size();
    }
    *out_len = total;
    return PSB_OK;
}

// VCF only: contig (NUL-terminated, back to back in `contigs` with contig_off[v] offsets), 1-based
// position and REF length of the records returned by the last psb_reader_next -- what the burden
// branch needs to decide which records a region fetches (input.py:395-407).
extern "C" int psb_reader_vcf_info(psb_reader *r, int64_t n, char *contigs, int64_t contigs_cap,
                                   int64_t *contig_off, int64_t *pos, int32_t *ref_len) {
    PSB_REQUIRE(r && contigs && contig_off && pos && ref_len, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(r->var_type == 2, PSB_ERR_STATE, "not a VCF reader");
    PSB_REQUIRE(n == (int64_t)r->vcf_contig.size(), PSB_ERR_ARG, "n = %lld, last batch had %zu records",
                (long long)n, r->vcf_contig.size());
    int64_t used = 0;
    for (int64_t v = 0; v < n; ++v) {
        const std::string &c = r->vcf_contig[v];
        PSB_REQUIRE(used + (int64_t)c.size() + 1 <= contigs_cap, PSB_ERR_NOMEM, "contig buffer too small");
        memcpy(contigs + used, c.data(), c.size());
        contigs[used + c.size()] = '\0';
        contig_off[v] = used;
        used += (int64_t)c.size() + 1;
        pos[v] = r->vcf_pos[v];
        ref_len[v] = r->vcf_reflen[v];
    }
    return PSB_OK;
}

// ---------------------------------------------------------------------------------------
// Pattern hashes for --output-patterns: input.hash_pattern (pyseer/input.py:710-723) =
// b2a_base64(md5(k.view(uint8))) of the presence/absence vector k the reference builds (input.py:450):
// int64 0/1 per sample, or float64 0.0/1.0/NaN when the variant has missing genotypes.  MD5 after
// RFC 1321, written out here (host-only code).
// ---------------------------------------------------------------------------------------
namespace {
struct Md5 {
    uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
    uint64_t total = 0;
    unsigned char buf[64];
    size_t fill = 0;

    static inline uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }

    void block(const unsigned char *p) {
        static const uint32_t K[64] = {
            0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
            0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
            0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
            0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
            0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
            0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
            0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
            0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
        static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                                  5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
        uint32_t w[16];
        for (int i = 0; i < 16; ++i)
            w[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
                   ((uint32_t)p[4 * i + 3] << 24);
        uint32_t A = a, B = b, C = c, D = d;
        for (int i = 0; i < 64; ++i) {
            uint32_t f;
            int g;
            if (i < 16) { f = (B & C) | (~B & D); g = i; }
            else if (i < 32) { f = (D & B) | (~D & C); g = (5 * i + 1) & 15; }
            else if (i < 48) { f = B ^ C ^ D; g = (3 * i + 5) & 15; }
            else { f = C ^ (B | ~D); g = (7 * i) & 15; }
            const uint32_t t = D;
            D = C;
            C = B;
            B = B + rol(A + f + K[i] + w[g], S[i]);
            A = t;
        }
        a += A; b += B; c += C; d += D;
    }

    void update(const unsigned char *p, size_t n) {
        total += n;
        while (n > 0) {
            const size_t take = std::min(n, (size_t)64 - fill);
            memcpy(buf + fill, p, take);
            fill += take;
            p += take;
            n -= take;
            if (fill == 64) { block(buf); fill = 0; }
        }
    }

    void finish(unsigned char out[16]) {
        const uint64_t bits = total * 8;
        const unsigned char pad = 0x80, zero = 0;
        update(&pad, 1);
        while (fill != 56) update(&zero, 1);
        unsigned char len[8];
        for (int i = 0; i < 8; ++i) len[i] = (unsigned char)(bits >> (8 * i));
        update(len, 8);
        const uint32_t st[4] = {a, b, c, d};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) out[4 * i + j] = (unsigned char)(st[i] >> (8 * j));
    }
};
}  // namespace

// out: 25 bytes per hashed row (24 base64 characters + '\n', what binascii.b2a_base64 returns).
// flags (nullable): rows with PSB_F_PREFILTER set are skipped -- the result loop hashes the tested
// variants only (__main__.py:559-560).  *n_out = rows hashed.
extern "C" int psb_hash_patterns(const uint32_t *bits, const uint32_t *missing, int64_t n_variants,
                                 int32_t words_per_row, int32_t n_samples, const uint32_t *flags, char *out,
                                 int64_t *n_out) {
    PSB_REQUIRE(bits && out && n_out, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(words_per_row * 32 >= n_samples && n_samples > 0, PSB_ERR_ARG, "bad row shape");
    static const char b64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    std::vector<unsigned char> k((size_t)n_samples * 8);
    int64_t m = 0;
    for (int64_t v = 0; v < n_variants; ++v) {
        if (flags && (flags[v] & PSB_F_PREFILTER)) continue;
        const uint32_t *row = bits + v * words_per_row;
        const uint32_t *mrow = missing ? missing + v * words_per_row : nullptr;
        bool any_missing = false;
        if (mrow)
            for (int w = 0; w < (n_samples + 31) / 32 && !any_missing; ++w) any_missing = mrow[w] != 0;
        memset(k.data(), 0, k.size());
        for (int i = 0; i < n_samples; ++i) {
            const bool on = (row[i >> 5] >> (i & 31)) & 1u;
            const bool ms = any_missing && ((mrow[i >> 5] >> (i & 31)) & 1u);
            unsigned char *c = &k[(size_t)i * 8];
            if (!any_missing) {
                c[0] = on ? 1 : 0;                                   // int64, little endian
            } else if (ms) {
                c[6] = 0xf8; c[7] = 0x7f;                            // float64 NaN (numpy.nan)
            } else if (on) {
                c[6] = 0xf0; c[7] = 0x3f;                            // float64 1.0
            }
        }
        Md5 h;
        h.update(k.data(), k.size());
        unsigned char dg[18] = {0};
        h.finish(dg);
        char *o = out + m * 25;
        for (int i = 0, j = 0; i < 18; i += 3, j += 4) {
            const uint32_t t = ((uint32_t)dg[i] << 16) | ((uint32_t)dg[i + 1] << 8) | dg[i + 2];
            o[j] = b64[(t >> 18) & 63];
            o[j + 1] = b64[(t >> 12) & 63];
            o[j + 2] = b64[(t >> 6) & 63];
            o[j + 3] = b64[t & 63];
        }
        o[22] = '=';                                                 // 16 bytes -> 22 characters + '=='
        o[23] = '=';
        o[24] = '\n';
        ++m;
    }
    *n_out = m;
    return PSB_OK;
}
