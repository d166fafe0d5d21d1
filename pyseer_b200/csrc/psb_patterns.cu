// psb_patterns.cu -- input.hash_pattern (pyseer/input.py:710-723) for every row of a batch, on the device.
//
// The reference writes, per tested variant, base64(md5(k)) of the vector k it fits -- int64 0/1 over the
// samples in phenotype order, float64 with NaN where a genotype is missing (input.py:450) -- to
// --output-patterns (__main__.py:559-560), 8 N bytes hashed per variant: 40 KB at N = 5000, 16 k
// variants/s on one host core (psb_hash_patterns), far below what the readers deliver.  Here one thread
// hashes one row straight from its packed bits: a 64-byte MD5 block is 8 samples, its sixteen words are
// (bit, 0) pairs (int64) or (0, high word of 0.0 / 1.0 / NaN) pairs (float64), so the message is never
// materialised.  The 16-byte digests go back to the host, which base64-encodes those of the tested rows.
#include "psb_internal.cuh"

#define MD5_F(x, y, z) (((x) & (y)) | (~(x) & (z)))
#define MD5_G(x, y, z) (((x) & (z)) | ((y) & ~(z)))
#define MD5_H(x, y, z) ((x) ^ (y) ^ (z))
#define MD5_I(x, y, z) ((y) ^ ((x) | ~(z)))
#define MD5_STEP(f, a, b, c, d, x, t, s) \
    a += f(b, c, d) + (x) + (t);         \
    a = __funnelshift_l(a, a, s) + b;

__device__ __forceinline__ void md5_block(uint32_t (&st)[4], const uint32_t (&w)[16]) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
    MD5_STEP(MD5_F, a, b, c, d, w[0], 0xd76aa478u, 7)
    MD5_STEP(MD5_F, d, a, b, c, w[1], 0xe8c7b756u, 12)
    MD5_STEP(MD5_F, c, d, a, b, w[2], 0x242070dbu, 17)
    MD5_STEP(MD5_F, b, c, d, a, w[3], 0xc1bdceeeu, 22)
    MD5_STEP(MD5_F, a, b, c, d, w[4], 0xf57c0fafu, 7)
    MD5_STEP(MD5_F, d, a, b, c, w[5], 0x4787c62au, 12)
    MD5_STEP(MD5_F, c, d, a, b, w[6], 0xa8304613u, 17)
    MD5_STEP(MD5_F, b, c, d, a, w[7], 0xfd469501u, 22)
    MD5_STEP(MD5_F, a, b, c, d, w[8], 0x698098d8u, 7)
    MD5_STEP(MD5_F, d, a, b, c, w[9], 0x8b44f7afu, 12)
    MD5_STEP(MD5_F, c, d, a, b, w[10], 0xffff5bb1u, 17)
    MD5_STEP(MD5_F, b, c, d, a, w[11], 0x895cd7beu, 22)
    MD5_STEP(MD5_F, a, b, c, d, w[12], 0x6b901122u, 7)
    MD5_STEP(MD5_F, d, a, b, c, w[13], 0xfd987193u, 12)
    MD5_STEP(MD5_F, c, d, a, b, w[14], 0xa679438eu, 17)
    MD5_STEP(MD5_F, b, c, d, a, w[15], 0x49b40821u, 22)
    MD5_STEP(MD5_G, a, b, c, d, w[1], 0xf61e2562u, 5)
    MD5_STEP(MD5_G, d, a, b, c, w[6], 0xc040b340u, 9)
    MD5_STEP(MD5_G, c, d, a, b, w[11], 0x265e5a51u, 14)
    MD5_STEP(MD5_G, b, c, d, a, w[0], 0xe9b6c7aau, 20)
    MD5_STEP(MD5_G, a, b, c, d, w[5], 0xd62f105du, 5)
    MD5_STEP(MD5_G, d, a, b, c, w[10], 0x02441453u, 9)
    MD5_STEP(MD5_G, c, d, a, b, w[15], 0xd8a1e681u, 14)
    MD5_STEP(MD5_G, b, c, d, a, w[4], 0xe7d3fbc8u, 20)
    MD5_STEP(MD5_G, a, b, c, d, w[9], 0x21e1cde6u, 5)
    MD5_STEP(MD5_G, d, a, b, c, w[14], 0xc33707d6u, 9)
    MD5_STEP(MD5_G, c, d, a, b, w[3], 0xf4d50d87u, 14)
    MD5_STEP(MD5_G, b, c, d, a, w[8], 0x455a14edu, 20)
    MD5_STEP(MD5_G, a, b, c, d, w[13], 0xa9e3e905u, 5)
    MD5_STEP(MD5_G, d, a, b, c, w[2], 0xfcefa3f8u, 9)
    MD5_STEP(MD5_G, c, d, a, b, w[7], 0x676f02d9u, 14)
    MD5_STEP(MD5_G, b, c, d, a, w[12], 0x8d2a4c8au, 20)
    MD5_STEP(MD5_H, a, b, c, d, w[5], 0xfffa3942u, 4)
    MD5_STEP(MD5_H, d, a, b, c, w[8], 0x8771f681u, 11)
    MD5_STEP(MD5_H, c, d, a, b, w[11], 0x6d9d6122u, 16)
    MD5_STEP(MD5_H, b, c, d, a, w[14], 0xfde5380cu, 23)
    MD5_STEP(MD5_H, a, b, c, d, w[1], 0xa4beea44u, 4)
    MD5_STEP(MD5_H, d, a, b, c, w[4], 0x4bdecfa9u, 11)
    MD5_STEP(MD5_H, c, d, a, b, w[7], 0xf6bb4b60u, 16)
    MD5_STEP(MD5_H, b, c, d, a, w[10], 0xbebfbc70u, 23)
    MD5_STEP(MD5_H, a, b, c, d, w[13], 0x289b7ec6u, 4)
    MD5_STEP(MD5_H, d, a, b, c, w[0], 0xeaa127fau, 11)
    MD5_STEP(MD5_H, c, d, a, b, w[3], 0xd4ef3085u, 16)
    MD5_STEP(MD5_H, b, c, d, a, w[6], 0x04881d05u, 23)
    MD5_STEP(MD5_H, a, b, c, d, w[9], 0xd9d4d039u, 4)
    MD5_STEP(MD5_H, d, a, b, c, w[12], 0xe6db99e5u, 11)
    MD5_STEP(MD5_H, c, d, a, b, w[15], 0x1fa27cf8u, 16)
    MD5_STEP(MD5_H, b, c, d, a, w[2], 0xc4ac5665u, 23)
    MD5_STEP(MD5_I, a, b, c, d, w[0], 0xf4292244u, 6)
    MD5_STEP(MD5_I, d, a, b, c, w[7], 0x432aff97u, 10)
    MD5_STEP(MD5_I, c, d, a, b, w[14], 0xab9423a7u, 15)
    MD5_STEP(MD5_I, b, c, d, a, w[5], 0xfc93a039u, 21)
    MD5_STEP(MD5_I, a, b, c, d, w[12], 0x655b59c3u, 6)
    MD5_STEP(MD5_I, d, a, b, c, w[3], 0x8f0ccc92u, 10)
    MD5_STEP(MD5_I, c, d, a, b, w[10], 0xffeff47du, 15)
    MD5_STEP(MD5_I, b, c, d, a, w[1], 0x85845dd1u, 21)
    MD5_STEP(MD5_I, a, b, c, d, w[8], 0x6fa87e4fu, 6)
    MD5_STEP(MD5_I, d, a, b, c, w[15], 0xfe2ce6e0u, 10)
    MD5_STEP(MD5_I, c, d, a, b, w[6], 0xa3014314u, 15)
    MD5_STEP(MD5_I, b, c, d, a, w[13], 0x4e0811a1u, 21)
    MD5_STEP(MD5_I, a, b, c, d, w[4], 0xf7537e82u, 6)
    MD5_STEP(MD5_I, d, a, b, c, w[11], 0xbd3af235u, 10)
    MD5_STEP(MD5_I, c, d, a, b, w[2], 0x2ad7d2bbu, 15)
    MD5_STEP(MD5_I, b, c, d, a, w[9], 0xeb86d391u, 21)
    st[0] += a;
    st[1] += b;
    st[2] += c;
    st[3] += d;
}

// words of the 8 samples whose presence / missing bits are the low 8 bits of `on` / `ms`
__device__ __forceinline__ void pattern_words(uint32_t on, uint32_t ms, bool as_float, int n, uint32_t (&w)[16]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool live = j < n;
        const uint32_t x = (on >> j) & 1u, m = (ms >> j) & 1u;
        if (!as_float) {
            w[2 * j] = live ? x : 0u;                                        // int64, little endian
            w[2 * j + 1] = 0u;
        } else {
            w[2 * j] = 0u;                                                   // float64: 0.0, 1.0 or numpy.nan
            w[2 * j + 1] = live ? (m ? 0x7ff80000u : (x ? 0x3ff00000u : 0u)) : 0u;
        }
    }
}

__global__ void __launch_bounds__(128)
k_pattern_md5(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ miss, int64_t S, int Wrow, int N,
              uint32_t *__restrict__ digests) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= S) return;
    const uint32_t *row = bits + v * Wrow;
    const uint32_t *mrow = miss ? miss + v * Wrow : nullptr;
    const int Wn = (N + 31) >> 5;
    bool as_float = false;                      // a row with a missing genotype is hashed as float64
    if (mrow)
        for (int t = 0; t < Wn; ++t) as_float |= __ldg(mrow + t) != 0u;
    uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    uint32_t w[16];
    const int nfull = N >> 3;
    uint32_t word = 0, mword = 0;
    for (int blk = 0; blk < nfull; ++blk) {
        if ((blk & 3) == 0) {
            word = __ldg(row + (blk >> 2));
            mword = as_float ? __ldg(mrow + (blk >> 2)) : 0u;
        }
        const int sh = (blk & 3) * 8;
        pattern_words(word >> sh, mword >> sh, as_float, 8, w);
        md5_block(st, w);
    }
    // last block(s): the remaining N % 8 samples, the 0x80 byte, zeros, the length in bits
    const int r = N & 7;
    uint32_t on = 0, ms = 0;
    if (r) {
        const int sh = (nfull & 3) * 8;
        on = __ldg(row + (nfull >> 2)) >> sh;
        ms = as_float ? __ldg(mrow + (nfull >> 2)) >> sh : 0u;
    }
    pattern_words(on, ms, as_float, r, w);
    w[2 * r] = 0x80u;                           // byte 8 r
    const unsigned long long nbits = (unsigned long long)N * 64ull;
    if (r == 7) {                               // 56 data bytes + 0x80: the length needs one more block
        md5_block(st, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = 0u;
    }
    w[14] = (uint32_t)nbits;
    w[15] = (uint32_t)(nbits >> 32);
    md5_block(st, w);
    uint4 o = make_uint4(st[0], st[1], st[2], st[3]);
    reinterpret_cast<uint4 *>(digests)[v] = o;
}

void psb_patterns_release(psb_ctx *c) {
    if (c->d_dig) cudaFree(c->d_dig);
    c->d_dig = nullptr;
    c->dig_cap = 0;
}

// MD5 digests (16 bytes each, as hashlib's digest()) of the vectors input.hash_pattern hashes, for
// every row of the batch last submitted / run.
extern "C" int psb_pattern_digests(psb_ctx *c, uint8_t *out) {
    PSB_REQUIRE(c && out, PSB_ERR_ARG, "NULL argument");
    PSB_CUDA(cudaSetDevice(c->device));
    int rc = psb_run_begin(c);
    if (rc) return rc;
    PSB_REQUIRE(c->d_bits || c->S == 0, PSB_ERR_STATE, "no rows submitted");
    if (c->S == 0) return PSB_OK;
    if (c->S > c->dig_cap) {
        PSB_CUDA(cudaStreamSynchronize(c->stream));
        psb_patterns_release(c);
        PSB_CUDA(cudaMalloc(&c->d_dig, (size_t)c->S * 16));
        c->dig_cap = c->S;
    }
    k_pattern_md5<<<(unsigned)((c->S + 127) / 128), 128, 0, c->stream>>>(c->d_bits, c->d_miss, c->S, c->Wrow, c->N,
                                                                        (uint32_t *)c->d_dig);
    c->launches++;
    PSB_CUDA(cudaGetLastError());
    PSB_CUDA(cudaMemcpyAsync(out, c->d_dig, (size_t)c->S * 16, cudaMemcpyDeviceToHost, c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    return PSB_OK;
}
