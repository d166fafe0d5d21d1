// psb_pgz.cu -- parallel inflate of one plain gzip stream (host code; no CUDA in this file).
//
// Why: pyseer's variant files are gzip'ed text (input.open_variant_file, pyseer/input.py:268-298).  A
// gzip file is ONE deflate stream: every match may point up to 32 KiB back, so zlib can only walk it
// from the start -- ~0.4 GB/s of text, 21 k variants/s at N = 5000 however many threads parse behind
// it.  bgzip files (independent 64 KiB members) already inflate block-parallel (psb_io.cu); this
// file does the same for plain gzip with the two-stage scheme of pugz / rapidgzip:
//
//   1. the compressed bytes of a batch are cut into chunks; every chunk but the first SEARCHES
//      for the start of a deflate block at or after its nominal offset (a non-final dynamic-Huffman
//      header whose code-length code and both Huffman codes are valid and whose first block decodes);
//   2. all chunks are decoded in parallel by the decoder below, which writes 16-bit symbols: a
//      literal byte, or -- where a match reaches back before the chunk's first byte, into the 32 KiB
//      the chunk cannot know -- a MARKER naming the position in that unknown window;
//   3. a chunk stops at a block boundary that is exactly the start another chunk found.  The first
//      chunk of a batch starts at a position known to be a true boundary, so by induction every chunk
//      whose start was reached exactly by its predecessor holds a true piece of the stream; a start
//      that no predecessor lands on was a false positive and its chunk is thrown away (the predecessor
//      simply decodes on to the next start).  Nothing is trusted that the serial decoding does not
//      confirm;
//   4. windows are chained through the kept chunks (32 KiB each, serial, microseconds), then all
//      chunks replace their markers and are checksummed in parallel; the member's CRC-32 and length
//      are verified against the gzip trailer as zlib would.
//
// Stored and fixed-Huffman blocks are decoded but never searched for (gzip emits dynamic blocks for
// text; a region without them is decoded serially by the preceding chunk).  Several gzip members in
// one file: a member that ends inside a batch is followed through (the chain walk decodes the first
// blocks of the next member up to the next start a chunk found; window and CRC start afresh).
#include "psb_pgz.h"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int LP = 11;                              // primary bits, literal / length table
constexpr int DP = 8;                               // primary bits, distance table
constexpr int LT_CAP = (1 << LP) + 288 * 16;        // + one subtable of 2^(15-11) per long code at most
constexpr int DT_CAP = (1 << DP) + 32 * 128;
constexpr uint32_t SUB = 0x80000000u;
constexpr size_t WIN = 32768;

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                               67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
                                9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t PRE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// LSB-first bit reader over [base, end).  pos() = bit offset of the next unread bit.
struct BitIn {
    const uint8_t *base = nullptr, *p = nullptr, *end = nullptr;
    uint64_t buf = 0;
    int cnt = 0;                                    // valid bits in buf (negative: read past the end)
    void init(const uint8_t *b, const uint8_t *e, uint64_t bit) {
        base = b;
        end = e;
        p = b + (bit >> 3);
        if (p > e) p = e;
        buf = 0;
        cnt = 0;
        refill();
        drop((int)(bit & 7));
    }
    inline void refill() {
        if (p + 8 <= end) {
            uint64_t w;
            memcpy(&w, p, 8);
            buf |= w << cnt;                        // bits above cnt are true stream bits, not yet counted
            const int n = (63 - cnt) >> 3;
            p += n;
            cnt += n * 8;
        } else {
            while (cnt <= 56 && p < end) {
                buf |= (uint64_t)*p++ << cnt;
                cnt += 8;
            }
        }
    }
    inline void drop(int n) {
        buf >>= n;
        cnt -= n;
    }
    inline uint32_t bits(int n) {
        const uint32_t v = (uint32_t)(buf & ((1ull << n) - 1));
        drop(n);
        return v;
    }
    uint64_t pos() const { return (uint64_t)(p - base) * 8 - (uint64_t)cnt; }
};

// Canonical Huffman decoding table: entry = symbol << 8 | code length; SUB | offset << 4 | bits points
// at a subtable for the codes longer than `pbits`; 0 = no code.  Returns 0 for a complete code, 1 for
// a single code of length 1 (the one incomplete set zlib accepts), 2 for no code at all, -1 otherwise.
int build_table(const uint8_t *lens, int n, int pbits, uint32_t *tab, int cap) {
    int count[16] = {0};
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    const int psize = 1 << pbits;
    memset(tab, 0, (size_t)psize * 4);
    if (count[0] == n) return 2;
    int left = 1, maxlen = 0;
    for (int l = 1; l <= 15; ++l) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return -1;                    // over-subscribed
        if (count[l]) maxlen = l;
    }
    int status = 0;
    if (left > 0) {
        if (maxlen != 1) return -1;                 // incomplete
        status = 1;
    }
    uint32_t next[16], code = 0;
    count[0] = 0;
    for (int l = 1; l <= 15; ++l) {
        code = (code + (uint32_t)count[l - 1]) << 1;
        next[l] = code;
    }
    int used = psize;
    const int sb = maxlen > pbits ? maxlen - pbits : 0;
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t c = next[l]++;
        uint32_t r = 0;
        for (int b = 0; b < l; ++b) r |= ((c >> b) & 1u) << (l - 1 - b);
        const uint32_t e = ((uint32_t)s << 8) | (uint32_t)l;
        if (l <= pbits) {
            for (uint32_t i = r; i < (uint32_t)psize; i += 1u << l) tab[i] = e;
        } else {
            const uint32_t pre = r & (uint32_t)(psize - 1);
            uint32_t pe = tab[pre];
            if (!(pe & SUB)) {
                if (pe || used + (1 << sb) > cap) return -1;
                memset(tab + used, 0, (size_t)4 << sb);
                pe = SUB | ((uint32_t)used << 4) | (uint32_t)sb;
                tab[pre] = pe;
                used += 1 << sb;
            }
            uint32_t *sub = tab + ((pe >> 4) & 0x7ffffffu);
            const uint32_t hi = r >> pbits;
            const int hl = l - pbits;
            for (uint32_t i = hi; i < (1u << sb); i += 1u << hl) sub[i] = e;
        }
    }
    return status;
}

struct Tables {
    uint32_t lt[LT_CAP];
    uint32_t dt[DT_CAP];
};

void fixed_tables(Tables &t) {
    uint8_t l[288], d[32];
    for (int i = 0; i < 144; ++i) l[i] = 8;
    for (int i = 144; i < 256; ++i) l[i] = 9;
    for (int i = 256; i < 280; ++i) l[i] = 7;
    for (int i = 280; i < 288; ++i) l[i] = 8;
    for (int i = 0; i < 32; ++i) d[i] = 5;
    build_table(l, 288, LP, t.lt, LT_CAP);
    build_table(d, 32, DP, t.dt, DT_CAP);
}

// Header of a dynamic block (after the 3 block bits) -> tables.  strict: what a block-start
// candidate must satisfy on top of what zlib accepts (complete literal / length code).
bool read_dynamic(BitIn &in, Tables &t, bool strict) {
    in.refill();
    const int hlit = (int)in.bits(5) + 257, hdist = (int)in.bits(5) + 1, hclen = (int)in.bits(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    uint8_t pl[19] = {0};
    in.refill();
    for (int i = 0; i < hclen; ++i) {
        if (i == 12) in.refill();
        pl[PRE_ORDER[i]] = (uint8_t)in.bits(3);
    }
    if (in.cnt < 0) return false;
    // the code-length code: complete, at most 7 bits -> one flat table
    int kraft = 0;
    for (int i = 0; i < 19; ++i)
        if (pl[i]) kraft += 128 >> pl[i];
    if (kraft != 128) return false;
    uint32_t pt[128];
    if (build_table(pl, 19, 7, pt, 128) != 0) return false;
    uint8_t lens[286 + 30 + 140];
    const int n = hlit + hdist;
    int i = 0;
    while (i < n) {
        in.refill();
        if (in.cnt < 0) return false;
        const uint32_t e = pt[in.buf & 127];
        if (!e) return false;
        in.drop((int)(e & 0xff));
        const int s = (int)(e >> 8);
        if (s < 16) {
            lens[i++] = (uint8_t)s;
        } else if (s == 16) {
            if (i == 0) return false;
            int rep = 3 + (int)in.bits(2);
            const uint8_t v = lens[i - 1];
            if (i + rep > n) return false;
            while (rep--) lens[i++] = v;
        } else {
            int rep = s == 17 ? 3 + (int)in.bits(3) : 11 + (int)in.bits(7);
            if (i + rep > n) return false;
            while (rep--) lens[i++] = 0;
        }
    }
    if (in.cnt < 0) return false;
    if (lens[256] == 0) return false;               // no end-of-block code
    const int ls = build_table(lens, hlit, LP, t.lt, LT_CAP);
    if (ls < 0 || ls == 2 || (strict && ls != 0)) return false;
    const int ds = build_table(lens + hlit, hdist, DP, t.dt, DT_CAP);
    if (ds < 0) return false;
    return true;
}

struct Chunk {
    uint16_t *sym = nullptr;        // [WIN markers][n symbols]
    size_t cap = 0, n = 0;
    uint64_t start = 0, end = 0;    // bit positions; end: the block boundary the decoding stopped at
    int status = -2;                // 0: stopped at a boundary, 1: end of the member (end = byte after the deflate data, in bits), -1: error, -2: no start found
    const char *why = "";
    size_t out_off = 0;
    uint32_t crc = 0;
    bool bad_marker = false;
    bool member_start = false;       // first chunk of a gzip member (nothing exists before it)
    Chunk() = default;
    Chunk(const Chunk &) = delete;
    Chunk &operator=(const Chunk &) = delete;
    Chunk(Chunk &&o) noexcept { *this = std::move(o); }
    Chunk &operator=(Chunk &&o) noexcept {
        if (this != &o) {
            free(sym);
            sym = o.sym; cap = o.cap; n = o.n; start = o.start; end = o.end; status = o.status; why = o.why;
            out_off = o.out_off; crc = o.crc; bad_marker = o.bad_marker; member_start = o.member_start;
            o.sym = nullptr; o.cap = 0;
        }
        return *this;
    }
    ~Chunk() { free(sym); }
    size_t max_syms = (size_t)1 << 28;      // 512 MB of symbols: beyond that the serial inflater takes over
    bool reserve(size_t want) {
        if (want <= cap) return true;
        if (want > max_syms + WIN + 8192) return false;
        size_t nc = std::max(want, cap + cap / 2 + 4096);
        uint16_t *q = (uint16_t *)realloc(sym, nc * sizeof(uint16_t));
        if (!q) return false;
        sym = q;
        cap = nc;
        return true;
    }
};

// Decodes blocks from bit `start` until a block boundary that is one of stops[] (ascending), or lies
// at / beyond `limit`, or the member ends.  window_known: number of bytes before `start` that exist at
// all (member start: 0) -- matches reaching further back are errors, not markers.
constexpr uint64_t NONE = ~(uint64_t)0, PENDING = ~(uint64_t)0 - 1;

// The starts the chunks of a batch found, published while the batch is being decoded: a chunk that
// reaches the range of chunk j at a block boundary compares with live[j] (waiting for it if the
// search of chunk j has not finished -- it started at the same time and takes a millisecond).
struct LiveStarts {
    const std::atomic<uint64_t> *start;     // [n]
    const uint64_t *nominal;                // [n + 1]
    int k, n;                               // this chunk, chunks in the batch
};

void decode_range(const uint8_t *base, const uint8_t *end, uint64_t start, const uint64_t *stops, int n_stops,
                  uint64_t limit, uint64_t known_before, Chunk &c, const Tables &fixed, size_t max_blocks = ~(size_t)0,
                  const LiveStarts *live = nullptr) {
    c.start = start;
    c.status = -1;
    c.n = 0;
    // room for ~6 bytes of text per compressed byte to begin with (grown as needed)
    const size_t guess = max_blocks == 1 ? ((size_t)1 << 18)
                                         : (size_t)std::min<uint64_t>(((limit > start ? limit - start : 0) >> 3) * 6 + 65536,
                                                                      (uint64_t)1 << 26);
    if (!c.reserve(WIN + guess + 1024)) { c.why = "out of memory"; return; }
    for (size_t i = 0; i < WIN; ++i) c.sym[i] = (uint16_t)(0x8000u | i);
    uint16_t *out = c.sym;
    size_t o = WIN;
    BitIn in;
    in.init(base, end, start);
    Tables *dyn = new Tables;
    int si = 0, lj = live ? live->k : 0;
    size_t blocks = 0;
    for (;;) {
        const uint64_t pos = in.pos();
        if (in.cnt < 0) { c.why = "unexpected end of the deflate stream"; break; }
        if (blocks > 0) {
            bool at_stop = false;
            if (live) {
                while (lj + 1 < live->n && pos >= live->nominal[lj + 1]) ++lj;     // the chunk whose range holds pos
                if (lj > live->k) {
                    uint64_t s;
                    while ((s = live->start[lj].load(std::memory_order_acquire)) == PENDING) std::this_thread::yield();
                    at_stop = s == pos;
                }
            } else {
                while (si < n_stops && stops[si] < pos) ++si;
                at_stop = si < n_stops && stops[si] == pos;
            }
            if (at_stop || pos >= limit || blocks >= max_blocks) {
                c.status = 0;
                c.end = pos;
                break;
            }
        }
        in.refill();
        const uint32_t bfinal = in.bits(1), btype = in.bits(2);
        const Tables *t = nullptr;
        if (btype == 0) {
            in.drop(in.cnt & 7);                    // to the byte boundary
            in.refill();
            const uint32_t len = in.bits(16), nlen = in.bits(16);
            if (in.cnt < 0 || (len ^ 0xffffu) != nlen) { c.why = "invalid stored block"; break; }
            const uint64_t q = in.pos() >> 3;
            if (base + q + len > end) { c.why = "unexpected end of the deflate stream"; break; }
            if (o + len + 1024 > c.cap) {
                if (!c.reserve(o + len + 1024)) { c.why = "out of memory"; break; }
                out = c.sym;
            }
            const uint8_t *s = base + q;
            for (uint32_t i = 0; i < len; ++i) out[o + i] = s[i];
            o += len;
            in.init(base, end, (q + len) * 8);
        } else if (btype == 3) {
            c.why = "invalid block type";
            break;
        } else {
            if (btype == 1) {
                t = &fixed;
            } else {
                if (!read_dynamic(in, *dyn, false)) { c.why = "invalid code lengths"; break; }
                t = dyn;
            }
            const uint32_t *lt = t->lt, *dt = t->dt;
            const uint64_t far_limit = known_before;       // bytes that exist before `start`
            bool ok = false;
            for (;;) {
                in.refill();
                if (in.cnt < 0) { c.why = "unexpected end of the deflate stream"; break; }
                uint32_t e = lt[in.buf & ((1u << LP) - 1)];
                if (e & SUB) e = lt[((e >> 4) & 0x7ffffffu) + ((in.buf >> LP) & ((1u << (e & 15)) - 1))];
                if (!e) { c.why = "invalid literal/length code"; break; }
                in.drop((int)(e & 0xff));
                uint32_t s = e >> 8;
                if (s < 256) {
                    out[o++] = (uint16_t)s;
                    // a second and third literal from the bits already in the buffer (>= 41 left)
                    e = lt[in.buf & ((1u << LP) - 1)];
                    if (!(e & SUB) && e && (e >> 8) < 256) {
                        in.drop((int)(e & 0xff));
                        out[o++] = (uint16_t)(e >> 8);
                        e = lt[in.buf & ((1u << LP) - 1)];
                        if (!(e & SUB) && e && (e >> 8) < 256) {
                            in.drop((int)(e & 0xff));
                            out[o++] = (uint16_t)(e >> 8);
                        }
                    }
                    if (o + 300 > c.cap) {
                        if (!c.reserve(o + (o >> 1) + 4096)) { c.why = "out of memory"; break; }
                        out = c.sym;
                    }
                    continue;
                }
                if (s == 256) { ok = true; break; }
                s -= 257;
                if (s >= 29) { c.why = "invalid length code"; break; }
                const uint32_t len = LEN_BASE[s] + in.bits(LEN_EXTRA[s]);
                uint32_t d = dt[in.buf & ((1u << DP) - 1)];
                if (d & SUB) d = dt[((d >> 4) & 0x7ffffffu) + ((in.buf >> DP) & ((1u << (d & 15)) - 1))];
                if (!d) { c.why = "invalid distance code"; break; }
                in.drop((int)(d & 0xff));
                d >>= 8;
                if (d >= 30) { c.why = "invalid distance code"; break; }
                const uint32_t dist = DIST_BASE[d] + in.bits(DIST_EXTRA[d]);
                if (dist > o - WIN && (uint64_t)(dist - (o - WIN)) > far_limit) { c.why = "invalid distance too far back"; break; }
                uint16_t *dp = out + o;
                const uint16_t *sp = dp - dist;
                o += len;
                if (dist >= 8) {
                    uint16_t *de = dp + len;
                    do {
                        memcpy(dp, sp, 16);
                        dp += 8;
                        sp += 8;
                    } while (dp < de);
                } else {
                    for (uint32_t i = 0; i < len; ++i) dp[i] = sp[i];
                }
                if (o + 300 > c.cap) {
                    if (!c.reserve(o + (o >> 1) + 4096)) { c.why = "out of memory"; break; }
                    out = c.sym;
                }
            }
            if (!ok) break;
        }
        ++blocks;
        if (bfinal) {
            if (in.cnt < 0) { c.why = "unexpected end of the deflate stream"; break; }
            in.drop(in.cnt & 7);
            c.status = 1;
            c.end = in.pos();
            break;
        }
    }
    delete dyn;
    c.n = o - WIN;
}

// A whole raw deflate stream whose decompressed size is known (a BGZF block: <= 64 KiB, no window before
// it) straight to bytes -- the decoder above without markers.  false: not a valid stream of that size.
bool inflate_exact(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, const Tables &fixed, Tables &dyn) {
    BitIn b;
    b.init(in, in + in_len, 0);
    size_t o = 0;
    for (;;) {
        b.refill();
        if (b.cnt < 3) return false;
        const uint32_t bfinal = b.bits(1), btype = b.bits(2);
        if (btype == 0) {
            b.drop(b.cnt & 7);
            b.refill();
            const uint32_t len = b.bits(16), nlen = b.bits(16);
            if (b.cnt < 0 || (len ^ 0xffffu) != nlen) return false;
            const uint64_t q = b.pos() >> 3;
            if (q + len > in_len || o + len > out_len) return false;
            memcpy(out + o, in + q, len);
            o += len;
            b.init(in, in + in_len, (q + len) * 8);
        } else if (btype == 3) {
            return false;
        } else {
            const Tables *t = &fixed;
            if (btype == 2) {
                if (!read_dynamic(b, dyn, false)) return false;
                t = &dyn;
            }
            const uint32_t *lt = t->lt, *dt = t->dt;
            for (;;) {
                b.refill();
                if (b.cnt < 0) return false;
                uint32_t e = lt[b.buf & ((1u << LP) - 1)];
                if (e & SUB) e = lt[((e >> 4) & 0x7ffffffu) + ((b.buf >> LP) & ((1u << (e & 15)) - 1))];
                if (!e) return false;
                b.drop((int)(e & 0xff));
                uint32_t s = e >> 8;
                if (s < 256) {
                    if (o >= out_len) return false;
                    out[o++] = (uint8_t)s;
                    e = lt[b.buf & ((1u << LP) - 1)];
                    if (!(e & SUB) && e && (e >> 8) < 256 && o < out_len) {
                        b.drop((int)(e & 0xff));
                        out[o++] = (uint8_t)(e >> 8);
                    }
                    continue;
                }
                if (s == 256) break;
                s -= 257;
                if (s >= 29) return false;
                const uint32_t len = LEN_BASE[s] + b.bits(LEN_EXTRA[s]);
                uint32_t d = dt[b.buf & ((1u << DP) - 1)];
                if (d & SUB) d = dt[((d >> 4) & 0x7ffffffu) + ((b.buf >> DP) & ((1u << (d & 15)) - 1))];
                if (!d) return false;
                b.drop((int)(d & 0xff));
                d >>= 8;
                if (d >= 30) return false;
                const uint32_t dist = DIST_BASE[d] + b.bits(DIST_EXTRA[d]);
                if (dist > o || o + len > out_len) return false;
                uint8_t *dp = out + o;
                const uint8_t *sp = dp - dist;
                if (dist >= 16 && o + len + 16 <= out_len) {         // 16 bytes at a time, may run over by < 16
                    uint8_t *de = dp + len;
                    do {
                        memcpy(dp, sp, 16);
                        dp += 16;
                        sp += 16;
                    } while (dp < de);
                } else {
                    for (uint32_t i = 0; i < len; ++i) dp[i] = sp[i];
                }
                o += len;
            }
        }
        if (bfinal) break;
    }
    return b.cnt >= 0 && o == out_len;
}

inline uint64_t peek64(const uint8_t *base, const uint8_t *end, uint64_t bit) {
    const uint8_t *p = base + (bit >> 3);
    uint64_t w = 0;
    if (p + 8 <= end) memcpy(&w, p, 8);
    else if (p < end) memcpy(&w, p, (size_t)(end - p));
    return w >> (bit & 7);
}

// First bit position in [from, to) where a non-final dynamic block plausibly starts and whose first
// block decodes; ~0 when there is none.
uint64_t find_block(const uint8_t *base, const uint8_t *end, uint64_t from, uint64_t to, const Tables &fixed) {
    Tables *t = new Tables;
    Chunk trial;
    uint64_t found = ~(uint64_t)0;
    for (uint64_t bit = from; bit < to; ++bit) {
        const uint64_t v = peek64(base, end, bit);
        if ((v & 7) != 4) continue;                             // BFINAL = 0, BTYPE = 2
        if (((v >> 3) & 31) > 29 || ((v >> 8) & 31) > 29) continue;
        // code-length code complete?  (3 bits each, HCLEN + 4 of them from bit 17; 56 bits are in v)
        const int hclen = (int)((v >> 13) & 15) + 4;
        int kraft = 0;
        uint64_t w = v >> 17;
        const int first = hclen < 13 ? hclen : 13;              // 17 + 39 = 56 bits
        for (int i = 0; i < first; ++i) {
            const int l = (int)(w & 7);
            w >>= 3;
            if (l) kraft += 128 >> l;
        }
        if (hclen > 13) {
            uint64_t w2 = peek64(base, end, bit + 56);
            for (int i = 13; i < hclen; ++i) {
                const int l = (int)(w2 & 7);
                w2 >>= 3;
                if (l) kraft += 128 >> l;
            }
        }
        if (kraft != 128) continue;
        BitIn in;
        in.init(base, end, bit + 3);
        if (!read_dynamic(in, *t, true)) continue;
        decode_range(base, end, bit, nullptr, 0, ~(uint64_t)0, WIN, trial, fixed, 1);
        if (trial.status != 0 || trial.n == 0) continue;
        found = bit;
        break;
    }
    delete t;
    return found;
}

// Worker threads kept for the life of a reader: a batch is three short parallel regions, and creating
// 3 x 15 threads per 16 MB of input cost a sixth of the run.  run() hands items 0 .. n-1 to up to
// `threads` threads (the caller is one of them) and returns when all are done.
class WorkPool {
  public:
    ~WorkPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto &t : th_) t.join();
    }
    void run(int threads, int n, const std::function<void(int)> &f) {
        if (threads > n) threads = n;
        if (threads <= 1) {
            for (int i = 0; i < n; ++i) f(i);
            return;
        }
        while ((int)th_.size() < threads - 1) {
            const int id = (int)th_.size();
            th_.emplace_back([this, id]() { loop(id); });
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &f;
            n_items_ = n;
            next_.store(0);
            want_ = threads - 1;
            pending_ = want_;
            ++gen_;
        }
        cv_work_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return pending_ == 0; });
    }

  private:
    void work() {
        for (;;) {
            const int i = next_.fetch_add(1);
            if (i >= n_items_) return;
            (*fn_)(i);
        }
    }
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return stop_ || (gen_ != seen && id < want_); });
                if (stop_) return;
                seen = gen_;
            }
            work();
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) cv_done_.notify_one();
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int)> *fn_ = nullptr;
    int n_items_ = 0, want_ = 0, pending_ = 0;
    std::atomic<int> next_{0};
    uint64_t gen_ = 0;
    bool stop_ = false;
};

}  // namespace

struct psb_pgz {
    int fd = -1;
    const uint8_t *map = nullptr;
    size_t size = 0;
    int n_threads = 1;
    size_t chunk_bytes = 1 << 20;
    uint64_t pos = 0;                   // bit position of the next block (a true boundary)
    bool in_member = false, eof = false, failed = false;
    uint8_t window[WIN];
    size_t window_len = 0;              // bytes of the member before `pos` (capped at WIN)
    uint32_t crc = 0;
    uint64_t member_out = 0;
    uint8_t *out = nullptr;             // decompressed bytes of the last batch
    size_t out_cap = 0, out_len = 0, out_pos = 0;
    std::string err;
    Tables fixed;
    int64_t n_chunks = 0, n_wasted = 0;
    std::vector<Chunk> pool;            // symbol buffers, kept from batch to batch
    std::vector<Chunk> bridge;          // ... and of the chunks that open a further member inside a batch
    WorkPool workers;
    size_t max_syms = (size_t)1 << 28;
    // serial continuation (zlib from the last confirmed block boundary): taken when a chunk the chain
    // needs could not be decoded here -- a stream this decoder mishandles, a chunk that inflates
    // beyond the symbol budget -- so that only streams zlib rejects as well are reported corrupt
    bool serial = false, zs_open = false;
    z_stream zs;
    int64_t n_serial = 0;               // bytes that came through the serial continuation
    double t_phase[4] = {0, 0, 0, 0};   // seconds in: block search, decoding, windows, markers + CRC
};

static bool pgz_fail(psb_pgz *z, const char *msg) {
    z->failed = true;
    z->err = msg;
    return false;
}

// gzip member header at byte `at` -> *first = byte of the first deflate block; false when there is no
// (complete) header there
static bool pgz_header_end(const psb_pgz *z, size_t at, size_t *first) {
    const uint8_t *m = z->map;
    const size_t n = z->size;
    if (at + 18 > n || m[at] != 0x1f || m[at + 1] != 0x8b || m[at + 2] != 8) return false;
    const int flg = m[at + 3];
    size_t p = at + 10;
    if (flg & 4) {
        if (p + 2 > n) return false;
        p += 2 + ((size_t)m[p] | ((size_t)m[p + 1] << 8));
    }
    for (int k = 0; k < 2; ++k)
        if (flg & (k == 0 ? 8 : 16)) {
            while (p < n && m[p]) ++p;
            ++p;
        }
    if (flg & 2) p += 2;
    if (p >= n) return false;
    *first = p;
    return true;
}

// the member after the trailer that starts at byte `tr`: *first = byte of its first deflate block
static bool pgz_next_member(const psb_pgz *z, size_t tr, size_t *first) {
    size_t nx = tr + 8;
    while (nx < z->size && z->map[nx] == 0) ++nx;               // zero padding after a member (gzip ignores it)
    return nx < z->size && pgz_header_end(z, nx, first);
}

// gzip member header at byte `at` -> bit position of the first deflate block; false when there is
// no (complete) header there
static bool pgz_header(psb_pgz *z, size_t at) {
    const uint8_t *m = z->map;
    const size_t n = z->size;
    if (at + 18 > n || m[at] != 0x1f || m[at + 1] != 0x8b || m[at + 2] != 8) return false;
    const int flg = m[at + 3];
    size_t p = at + 10;
    if (flg & 4) {
        if (p + 2 > n) return false;
        p += 2 + ((size_t)m[p] | ((size_t)m[p + 1] << 8));
    }
    for (int k = 0; k < 2; ++k)
        if (flg & (k == 0 ? 8 : 16)) {
            while (p < n && m[p]) ++p;
            ++p;
        }
    if (flg & 2) p += 2;
    if (p >= n) return false;
    z->pos = (uint64_t)p * 8;
    z->in_member = true;
    z->window_len = 0;
    z->crc = (uint32_t)crc32(0L, Z_NULL, 0);
    z->member_out = 0;
    return true;
}

// One raw deflate stream of known decompressed size -> bytes (thread safe; 0 = ok).
int psb_pgz_inflate_exact(const unsigned char *in, size_t in_len, unsigned char *out, size_t out_len) {
    static const Tables *fixed = []() {
        Tables *t = new Tables;
        fixed_tables(*t);
        return t;
    }();
    Tables dyn;                                     // 27 KB on the stack: no allocator traffic per 64 KiB block
    return inflate_exact(in, in_len, out, out_len, *fixed, dyn) ? 0 : -1;
}

psb_pgz *psb_pgz_open(const char *path, int n_threads, size_t chunk_bytes) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return nullptr;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 18) { close(fd); return nullptr; }
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { close(fd); return nullptr; }
    madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
    psb_pgz *z = new psb_pgz();
    z->fd = fd;
    z->map = (const uint8_t *)m;
    z->size = (size_t)st.st_size;
    z->n_threads = n_threads < 1 ? 1 : n_threads;
    if (chunk_bytes) z->chunk_bytes = chunk_bytes < 4096 ? 4096 : chunk_bytes;
    if (getenv("PSB_PGZ_MAX_SYMS")) z->max_syms = (size_t)atol(getenv("PSB_PGZ_MAX_SYMS"));   // test hook
    fixed_tables(z->fixed);
    if (!pgz_header(z, 0)) {
        psb_pgz_close(z);
        return nullptr;
    }
    return z;
}

void psb_pgz_set_threads(psb_pgz *z, int n_threads) {
    if (z) z->n_threads = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
}

const char *psb_pgz_error(const psb_pgz *z) { return z ? z->err.c_str() : "no reader"; }

void psb_pgz_stats(const psb_pgz *z, int64_t out[2]) {
    out[0] = z ? z->n_chunks : 0;
    out[1] = z ? z->n_wasted : 0;
}

void psb_pgz_close(psb_pgz *z) {
    if (!z) return;
    if (z->zs_open) inflateEnd(&z->zs);
    if (z->map) munmap((void *)z->map, z->size);
    if (z->fd >= 0) close(z->fd);
    free(z->out);
    delete z;
}

// Decodes the next batch into z->out.  false: error (z->err) or end of file (z->eof).
// Whole chunks that fit `direct` (direct_cap bytes) are written there and counted in *direct_used;
// the rest of the batch goes to z->out.
static bool pgz_batch(psb_pgz *z, uint8_t *direct, size_t direct_cap, size_t *direct_used) {
    z->out_len = z->out_pos = 0;
    *direct_used = 0;
    if (z->failed || z->eof) return false;
    const uint8_t *base = z->map, *end = z->map + z->size;
    const int T = std::min(z->n_threads, 32);        // work items per batch (symbol buffers: ~12-20 MB each)
    const uint64_t total_bits = (uint64_t)z->size * 8;
    // nominal chunk boundaries: byte aligned, chunk_bytes apart, from the current position
    const uint64_t p0 = z->pos;
    std::vector<uint64_t> nominal(T + 1);
    nominal[0] = p0;
    for (int k = 1; k <= T; ++k) nominal[k] = std::min(total_bits, ((p0 >> 3) + (uint64_t)k * z->chunk_bytes) * 8);
    int n = T;
    while (n > 1 && nominal[n - 1] >= total_bits) --n;          // chunks that would start at the end of the file
    const uint64_t limit = nominal[n];
    if ((int)z->pool.size() < n) z->pool.resize(n);
    std::vector<Chunk> &ch = z->pool;
    for (int k = 0; k < n; ++k) {
        ch[k].status = -2;
        ch[k].n = 0;
        ch[k].max_syms = z->max_syms;
    }
    const auto t0 = std::chrono::steady_clock::now();
    // ---- 1 + 2. every chunk: find its block start (all but the first), publish it, decode ----
    std::vector<std::atomic<uint64_t>> live(n);
    live[0].store(p0);
    for (int k = 1; k < n; ++k) live[k].store(PENDING);
    z->workers.run(T, n, [&](int k) {
        if (k > 0) live[k].store(find_block(base, end, nominal[k], nominal[k + 1], z->fixed), std::memory_order_release);
        const uint64_t st = live[k].load();
        if (st == NONE) return;
        const LiveStarts ls = {live.data(), nominal.data(), k, n};
        decode_range(base, end, st, nullptr, 0, limit, k == 0 ? (uint64_t)z->window_len : (uint64_t)WIN, ch[k],
                     z->fixed, ~(size_t)0, &ls);
    });
    const auto t1 = t0;
    std::vector<uint64_t> stops;
    std::vector<int> stop_owner;
    for (int k = 1; k < n; ++k)
        if (live[k].load() != NONE) {
            stops.push_back(live[k].load());
            stop_owner.push_back(k);
        }
    const auto t2 = std::chrono::steady_clock::now();
    // ---- 3. the chain of chunks the serial decoding confirms ----
    // A member that ends inside the batch does not end the batch: the first block of the next member is
    // decoded here (a "bridge", serial, at most up to the next start a chunk found) and the chain goes on
    // through the chunks already decoded -- concatenated gzip files keep all their threads busy.
    constexpr int MAX_BRIDGES = 8;
    if (z->bridge.size() < (size_t)MAX_BRIDGES) z->bridge.resize(MAX_BRIDGES);
    for (int k = 0; k < n; ++k) ch[k].member_start = false;
    std::vector<Chunk *> used;
    Chunk *cur = &ch[0];
    int n_bridges = 0;
    bool member_end = false;
    for (;;) {
        Chunk &c = *cur;
        if (c.status < 0) {
            // what was confirmed so far is delivered; zlib continues from the boundary this chunk started at
            z->serial = true;
            z->err = c.why;
            break;
        }
        used.push_back(cur);
        if (c.status == 1) {
            size_t first = 0;
            const size_t tr = (size_t)(c.end >> 3);
            if (tr + 8 <= z->size && n_bridges < MAX_BRIDGES && pgz_next_member(z, tr, &first) &&
                (uint64_t)first * 8 < limit) {
                Chunk &b = z->bridge[n_bridges];
                b.max_syms = z->max_syms;
                const uint64_t at = (uint64_t)first * 8;
                const size_t s0 = std::upper_bound(stops.begin(), stops.end(), at) - stops.begin();
                decode_range(base, end, at, stops.data() + s0, (int)(stops.size() - s0), limit, 0, b, z->fixed);
                b.member_start = true;
                if (b.status >= 0) {
                    ++n_bridges;
                    cur = &b;
                    continue;
                }
                // the new member does not decode here: it opens the next batch (and goes to zlib there)
            }
            member_end = true;
            break;
        }
        const auto it = std::lower_bound(stops.begin(), stops.end(), c.end);
        if (it != stops.end() && *it == c.end) {
            cur = &ch[stop_owner[it - stops.begin()]];
            continue;
        }
        break;                                                  // stopped at / beyond the batch limit
    }
    z->n_chunks += n;
    z->n_wasted += n + n_bridges - (int64_t)used.size();
    if (getenv("PSB_PGZ_DEBUG")) {
        fprintf(stderr, "batch at %llu: n %d bridges %d used %zu member_end %d serial %d; starts:", (unsigned long long)p0, n,
                n_bridges, used.size(), (int)member_end, (int)z->serial);
        for (int k = 0; k < n; ++k)
            fprintf(stderr, " %lld(%d,%lld)", live[k].load() == NONE ? -1ll : (long long)(live[k].load() - nominal[k]), ch[k].status,
                    ch[k].status >= 0 ? (long long)(ch[k].end - nominal[k]) : -1ll);
        fprintf(stderr, "\n");
    }
    if (used.empty()) return true;                              // nothing confirmed: the serial reader starts at z->pos
    // ---- 4. windows (serial, 32 KiB per chunk), then markers -> bytes and CRC in parallel ----
    // destinations: the leading chunks that fit the caller's buffer go straight there
    size_t total = 0, n_direct = 0, direct_bytes = 0;
    for (size_t u = 0; u < used.size(); ++u) {
        Chunk &c = *used[u];
        if (n_direct == u && direct && direct_bytes + c.n <= direct_cap) {
            c.out_off = direct_bytes;
            direct_bytes += c.n;
            ++n_direct;
        } else {
            c.out_off = total;
            total += c.n;
        }
    }
    if (total > z->out_cap) {
        free(z->out);
        z->out_cap = total + (total >> 3) + 4096;
        z->out = (uint8_t *)malloc(z->out_cap);
        if (!z->out) { z->out_cap = 0; return pgz_fail(z, "out of memory"); }
    }
    std::vector<std::vector<uint8_t>> wins(used.size());        // window in front of each used chunk
    std::vector<size_t> win_len(used.size());
    {
        std::vector<uint8_t> w(z->window, z->window + WIN);     // right aligned: w[WIN-1] = last byte
        size_t wl = z->window_len;
        for (size_t u = 0; u < used.size(); ++u) {
            const Chunk &c = *used[u];
            if (c.member_start) {                               // a new member: nothing before its first byte
                std::fill(w.begin(), w.end(), (uint8_t)0);
                wl = 0;
            }
            wins[u] = w;
            win_len[u] = wl;
            // next window = last WIN bytes of (w ++ resolved chunk)
            std::vector<uint8_t> nw(WIN);
            const size_t take = std::min(c.n, WIN);
            if (take < WIN) memcpy(nw.data(), w.data() + take, WIN - take);
            const uint16_t *s = c.sym + WIN + c.n - take;
            for (size_t i = 0; i < take; ++i) {
                const uint16_t v = s[i];
                nw[WIN - take + i] = v < 0x8000u ? (uint8_t)v : w[v & 0x7fffu];
            }
            w.swap(nw);
            wl = std::min(WIN, wl + c.n);
        }
        memcpy(z->window, w.data(), WIN);
        z->window_len = wl;
    }
    const auto t3 = std::chrono::steady_clock::now();
    z->workers.run(T, (int)used.size(), [&](int u) {
        Chunk &c = *used[u];
        const uint8_t *w = wins[u].data();
        const size_t lowest = WIN - win_len[u];         // markers below this index name bytes that do not exist
        const uint16_t *s = c.sym + WIN;
        uint8_t *o = ((size_t)u < n_direct ? direct : z->out) + c.out_off;
        bool bad = false;
        if (lowest > 0) {                               // only the first 32 KiB of a member can hold such markers
            for (size_t i = 0; i < c.n; ++i) bad |= s[i] >= 0x8000u && (size_t)(s[i] & 0x7fffu) < lowest;
        }
        // symbol -> byte through one flat table (identity for literals, the window for markers);
        // runs of 16 literals are packed without it
        std::vector<uint8_t> lut(65536);
        for (int b = 0; b < 256; ++b) lut[b] = (uint8_t)b;
        memcpy(lut.data() + 0x8000, w, WIN);
        const uint8_t *L = lut.data();
        size_t i = 0;
#if defined(__SSE2__)
        for (; i + 16 <= c.n; i += 16) {
            const __m128i a = _mm_loadu_si128((const __m128i *)(s + i));
            const __m128i b = _mm_loadu_si128((const __m128i *)(s + i + 8));
            if (_mm_movemask_epi8(_mm_or_si128(a, b)) & 0xAAAA) {
                for (int j = 0; j < 16; ++j) o[i + j] = L[s[i + j]];
            } else {
                _mm_storeu_si128((__m128i *)(o + i), _mm_packus_epi16(a, b));
            }
        }
#endif
        for (; i < c.n; ++i) o[i] = L[s[i]];
        c.bad_marker = bad;
        size_t done = 0;
        uint32_t crc = (uint32_t)crc32(0L, Z_NULL, 0);
        while (done < c.n) {
            const size_t step = std::min<size_t>(c.n - done, (size_t)1 << 30);
            crc = (uint32_t)crc32(crc, o + done, (uInt)step);
            done += step;
        }
        c.crc = crc;
    });
    const auto t4 = std::chrono::steady_clock::now();
    z->t_phase[0] += std::chrono::duration<double>(t1 - t0).count();
    z->t_phase[1] += std::chrono::duration<double>(t2 - t1).count();
    z->t_phase[2] += std::chrono::duration<double>(t3 - t2).count();
    z->t_phase[3] += std::chrono::duration<double>(t4 - t3).count();
    auto trailer_ok = [&](const Chunk &c) -> bool {            // CRC-32 and length of the member that ends with c
        const size_t tr = (size_t)(c.end >> 3);
        if (tr + 8 > z->size) return pgz_fail(z, "corrupt gzip stream: truncated trailer");
        const uint8_t *t = z->map + tr;
        const uint32_t want_crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        const uint32_t want_len = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
        if (want_crc != z->crc) return pgz_fail(z, "corrupt gzip stream: CRC-32 mismatch");
        if (want_len != (uint32_t)(z->member_out & 0xffffffffu)) return pgz_fail(z, "corrupt gzip stream: length mismatch");
        return true;
    };
    for (size_t u = 0; u < used.size(); ++u) {
        const Chunk &c = *used[u];
        if (c.bad_marker) return pgz_fail(z, "corrupt gzip stream: invalid distance too far back");
        if (c.member_start) {
            z->crc = (uint32_t)crc32(0L, Z_NULL, 0);
            z->member_out = 0;
        }
        z->crc = (uint32_t)crc32_combine(z->crc, c.crc, (z_off_t)c.n);
        z->member_out += c.n;
        if (c.status == 1 && u + 1 < used.size() && !trailer_ok(c)) return false;   // a member that ended inside the batch
    }
    z->out_len = total;
    *direct_used = direct_bytes;
    const Chunk &last = *used.back();
    if (z->serial) {
        z->pos = last.end;                                      // a confirmed boundary; window and CRC are current
        return true;
    }
    if (member_end) {
        if (!trailer_ok(last)) return false;
        const size_t tr = (size_t)(last.end >> 3);
        z->in_member = false;
        size_t nx = tr + 8;
        while (nx < z->size && z->map[nx] == 0) ++nx;           // zero padding after a member (gzip ignores it)
        // another member, or the end of the data (trailing garbage is ignored, as gzip and zlib do)
        if (nx >= z->size || !pgz_header(z, nx)) z->eof = true;
    } else {
        z->pos = last.end;
        if (z->pos >= total_bits) return pgz_fail(z, "corrupt gzip stream: unexpected end of file");
    }
    return true;
}

// zlib from z->pos (a block boundary of the current member, window known) to the end of the file
static int64_t pgz_serial_read(psb_pgz *z, char *dst, int64_t want) {
    int64_t got = 0;
    while (got < want) {
        if (!z->zs_open) {
            if (z->eof) break;
            memset(&z->zs, 0, sizeof(z->zs));
            if (inflateInit2(&z->zs, -15) != Z_OK) { pgz_fail(z, "inflateInit2 failed"); return -1; }
            size_t byte = (size_t)(z->pos >> 3);
            const int bit = (int)(z->pos & 7);
            if (bit) {
                inflatePrime(&z->zs, 8 - bit, z->map[byte] >> bit);
                ++byte;
            }
            if (z->window_len)
                inflateSetDictionary(&z->zs, z->window + (WIN - z->window_len), (uInt)z->window_len);
            z->zs.next_in = const_cast<Bytef *>(z->map + byte);
            z->zs.avail_in = 0;
            z->zs_open = true;
        }
        if (z->zs.avail_in == 0) {
            const size_t left = (size_t)((z->map + z->size) - z->zs.next_in);
            z->zs.avail_in = (uInt)std::min<size_t>(left, (size_t)1 << 30);
        }
        const size_t room = (size_t)std::min<int64_t>(want - got, (int64_t)1 << 30);
        z->zs.next_out = (Bytef *)dst + got;
        z->zs.avail_out = (uInt)room;
        const int rc = inflate(&z->zs, Z_NO_FLUSH);
        const size_t made = room - z->zs.avail_out;
        z->crc = (uint32_t)crc32(z->crc, (const Bytef *)dst + got, (uInt)made);
        z->member_out += made;
        z->n_serial += (int64_t)made;
        got += (int64_t)made;
        if (rc == Z_STREAM_END) {
            const uint8_t *t = z->zs.next_in;
            inflateEnd(&z->zs);
            z->zs_open = false;
            if (t + 8 > z->map + z->size) { pgz_fail(z, "corrupt gzip stream: truncated trailer"); return -1; }
            const uint32_t want_crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
            const uint32_t want_len = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
            if (want_crc != z->crc) { pgz_fail(z, "corrupt gzip stream: CRC-32 mismatch"); return -1; }
            if (want_len != (uint32_t)(z->member_out & 0xffffffffu)) { pgz_fail(z, "corrupt gzip stream: length mismatch"); return -1; }
            size_t nx = (size_t)(t + 8 - z->map);
            while (nx < z->size && z->map[nx] == 0) ++nx;
            if (nx >= z->size || !pgz_header(z, nx)) z->eof = true;
        } else if (rc != Z_OK && !(rc == Z_BUF_ERROR && made > 0)) {
            inflateEnd(&z->zs);
            z->zs_open = false;
            z->err = std::string("corrupt gzip stream: ") + (z->zs.msg ? z->zs.msg : (rc == Z_BUF_ERROR ? "unexpected end of file" : "inflate failed")) +
                     (z->err.empty() ? std::string() : " (parallel decoder: " + z->err + ")");
            z->failed = true;
            return -1;
        }
    }
    return got;
}

int64_t psb_pgz_read(psb_pgz *z, char *dst, int64_t want) {
    if (!z || want < 0) return -1;
    int64_t got = 0;
    while (got < want) {
        if (z->out_pos == z->out_len) {
            if (z->failed) return -1;
            if (z->serial) {
                const int64_t k = pgz_serial_read(z, dst + got, want - got);
                if (k < 0) return -1;
                got += k;
                break;
            }
            if (z->eof) break;
            size_t direct = 0;
            const bool ok = pgz_batch(z, (uint8_t *)dst + got, (size_t)(want - got), &direct);
            if (z->failed) return -1;
            got += (int64_t)direct;
            if (!ok) break;
            continue;
        }
        const size_t n = std::min<size_t>((size_t)(want - got), z->out_len - z->out_pos);
        memcpy(dst + got, z->out + z->out_pos, n);
        z->out_pos += n;
        got += (int64_t)n;
    }
    return got;
}

// Test hook: inflates the whole file through the parallel reader; CRC-32 and length of the text.
extern "C" int psb_pgz_selftest(const char *path, int32_t n_threads, int64_t chunk_bytes, uint32_t *crc_out,
                                int64_t *len_out, int64_t stats_out[2]) {
    psb_pgz *z = psb_pgz_open(path, n_threads, (size_t)chunk_bytes);
    if (!z) return -1;
    // reads of an odd size of about 1 MiB: chunks that fit land in the caller's buffer, larger ones come
    // through the reader's own buffer (PSB_PGZ_SELFTEST_BUF: another size, for timing)
    std::vector<char> buf(getenv("PSB_PGZ_SELFTEST_BUF") ? (size_t)atol(getenv("PSB_PGZ_SELFTEST_BUF")) : ((size_t)1 << 20) + 4099);
    uint32_t crc = (uint32_t)crc32(0L, Z_NULL, 0);
    int64_t total = 0;
    int rc = 0;
    for (;;) {
        const int64_t k = psb_pgz_read(z, buf.data(), (int64_t)buf.size());
        if (k < 0) { rc = -2; break; }
        if (k == 0) break;
        if (crc_out) crc = (uint32_t)crc32(crc, (const Bytef *)buf.data(), (uInt)k);
        total += k;
    }
    if (crc_out) *crc_out = crc;
    if (len_out) *len_out = total;
    if (stats_out) psb_pgz_stats(z, stats_out);
    if (getenv("PSB_PGZ_TIMES") && z->n_serial) fprintf(stderr, "psb_pgz: %lld bytes through the serial continuation\n", (long long)z->n_serial);
    if (getenv("PSB_PGZ_TIMES"))
        fprintf(stderr, "psb_pgz phases: search %.3f decode %.3f windows %.3f resolve %.3f s\n", z->t_phase[0], z->t_phase[1],
                z->t_phase[2], z->t_phase[3]);
    if (rc) fprintf(stderr, "psb_pgz: %s\n", psb_pgz_error(z));
    psb_pgz_close(z);
    return rc;
}
