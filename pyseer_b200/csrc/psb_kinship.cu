// psb_kinship.cu -- sample similarity (kinship) matrix from packed variants.
//
// Replaces the dense product of pyseer/similarity.py:99-116 (`K = G G'` with G the N x V
// presence/absence matrix of the variants that pass the AF / missing filter, as loaded by
// input.load_var_block): K[i][j] = number of kept variants carried by both samples.  With
// 1 bit per genotype this is an AND + POPCOUNT contraction over the variant axis:
//   k_kin_keep       carriers per variant -> keep flag (input.py:693 filter)
//   k_kin_transpose  32 x 32 bit-block transposes: variant-major rows -> sample-major words
//                    XT[chunk][sample] (bit l = variant 32 chunk + l), dropped variants zeroed
//   k_kin_accum      K[I tile][J tile] += sum_chunks popc(XT[c][i] & XT[c][j]), 64 x 64 tiles,
//                    4 x 4 outputs per thread, upper triangle only (mirrored on fetch)
// Exact integer arithmetic (int32 per launch, int64 accumulator).
#include <algorithm>
#include <vector>

#include "psb_internal.cuh"

#define KIN_TILE 64
#define KIN_CK 64          // chunks (of 32 variants) staged per shared-memory pass

__global__ void __launch_bounds__(256)
k_kin_keep(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ miss, int64_t S, int Wrow,
           int Wn, int N, double min_af, double max_af, double max_missing,
           uint8_t *__restrict__ keep) {
    const int lane = threadIdx.x & 31;
    int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (; v < S; v += stride) {
        int c = 0, m = 0;
        for (int w = lane; w < Wn; w += 32) {
            uint32_t x = __ldg(bits + v * Wrow + w);
            uint32_t mm = miss ? __ldg(miss + v * Wrow + w) : 0u;
            c += __popc(x | mm);
            m += __popc(mm);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, o);
            m += __shfl_xor_sync(0xffffffffu, m, o);
        }
        if (lane == 0) {
            const double af = (double)c / (double)N, missing = (double)m / (double)N;
            keep[v] = !(af < min_af || af > max_af || missing > max_missing);
        }
    }
}

// one warp per (chunk, word): 32 variants x 32 samples bit block, transposed with ballots
__global__ void __launch_bounds__(256)
k_kin_transpose(const uint32_t *__restrict__ bits, const uint8_t *__restrict__ keep, int64_t S, int Wrow,
                int Wn, int Npad, int64_t n_chunks, uint32_t *__restrict__ XT) {
    const int lane = threadIdx.x & 31;
    const int64_t total = n_chunks * Wn;
    int64_t job = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (; job < total; job += stride) {
        const int64_t c = job / Wn;
        const int w = (int)(job - c * Wn);
        const int64_t v = c * 32 + lane;
        uint32_t x = 0;
        if (v < S && keep[v]) x = __ldg(bits + v * Wrow + w);
        uint32_t mine = 0;
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            uint32_t t = __ballot_sync(0xffffffffu, (x >> b) & 1u);
            if (lane == b) mine = t;
        }
        XT[c * Npad + w * 32 + lane] = mine;
    }
}

__global__ void __launch_bounds__(256)
k_kin_accum(const uint32_t *__restrict__ XT, int64_t n_chunks, int Npad, int N, int n_tiles,
            long long *__restrict__ K) {
    __shared__ __align__(16) uint32_t As[KIN_CK][KIN_TILE];
    __shared__ __align__(16) uint32_t Bs[KIN_CK][KIN_TILE];
    // upper-triangular tile index -> (ti, tj), tj >= ti
    int t = blockIdx.x, ti = 0;
    while (t >= n_tiles - ti) { t -= n_tiles - ti; ++ti; }
    const int tj = ti + t;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    int acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[r][s] = 0;
    for (int64_t c0 = 0; c0 < n_chunks; c0 += KIN_CK) {
        __syncthreads();
        for (int e = threadIdx.x; e < KIN_CK * KIN_TILE; e += 256) {
            const int cc = e / KIN_TILE, col = e - cc * KIN_TILE;
            const int64_t c = c0 + cc;
            uint32_t a = 0, b = 0;
            if (c < n_chunks) {
                a = XT[c * Npad + ti * KIN_TILE + col];
                b = XT[c * Npad + tj * KIN_TILE + col];
            }
            As[cc][col] = a;
            Bs[cc][col] = b;
        }
        __syncthreads();
#pragma unroll 8
        for (int cc = 0; cc < KIN_CK; ++cc) {
            const uint4 a4 = *reinterpret_cast<const uint4 *>(&As[cc][ty * 4]);
            const uint4 b4 = *reinterpret_cast<const uint4 *>(&Bs[cc][tx * 4]);
            const uint32_t a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int s = 0; s < 4; ++s) acc[r][s] += __popc(a[r] & b[s]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int i = ti * KIN_TILE + ty * 4 + r, j = tj * KIN_TILE + tx * 4 + s;
            if (i < N && j < N) K[(size_t)i * N + j] += acc[r][s];
        }
}

struct psb_kin {
    int N = 0, Wn = 0, Npad = 0;
    long long *d_K = nullptr;
    uint32_t *d_bits = nullptr, *d_miss = nullptr, *d_XT = nullptr;
    uint8_t *d_keep = nullptr;
    size_t cap_bytes = 0;
    int64_t kept = 0, seen = 0;
};

static void kin_free(psb_kin *k) {
    if (!k) return;
    cudaFree(k->d_K); cudaFree(k->d_bits); cudaFree(k->d_miss); cudaFree(k->d_XT); cudaFree(k->d_keep);
    delete k;
}

void psb_kinship_release(psb_ctx *c) {
    kin_free((psb_kin *)c->kin);
    c->kin = nullptr;
}

extern "C" int psb_kinship_begin(psb_ctx *c, int32_t n_samples) {
    PSB_REQUIRE(c && n_samples > 0, PSB_ERR_ARG, "bad argument");
    PSB_CUDA(cudaSetDevice(c->device));
    psb_kinship_release(c);
    psb_kin *k = new psb_kin();
    k->N = n_samples;
    k->Wn = (n_samples + 31) / 32;
    k->Npad = ((n_samples + KIN_TILE - 1) / KIN_TILE) * KIN_TILE;
    c->kin = k;
    PSB_CUDA(cudaMalloc(&k->d_K, (size_t)n_samples * n_samples * sizeof(long long)));
    PSB_CUDA(cudaMemsetAsync(k->d_K, 0, (size_t)n_samples * n_samples * sizeof(long long), c->stream));
    return PSB_OK;
}

extern "C" int psb_kinship_add(psb_ctx *c, const uint32_t *bits, const uint32_t *missing,
                               int64_t n_variants, int32_t words_per_row, double min_af,
                               double max_af, double max_missing) {
    PSB_REQUIRE(c && c->kin, PSB_ERR_STATE, "psb_kinship_add before psb_kinship_begin");
    psb_kin *k = (psb_kin *)c->kin;
    PSB_REQUIRE(bits || n_variants == 0, PSB_ERR_ARG, "bits is NULL");
    PSB_REQUIRE(words_per_row >= k->Wn, PSB_ERR_ARG, "words_per_row too small");
    if (n_variants == 0) return PSB_OK;
    PSB_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)n_variants * words_per_row * 4;
    const int64_t n_chunks = (n_variants + 31) / 32;
    if (bytes > k->cap_bytes || (missing && !k->d_miss)) {
        PSB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(k->d_bits); cudaFree(k->d_miss); cudaFree(k->d_XT); cudaFree(k->d_keep);
        k->d_bits = k->d_miss = k->d_XT = nullptr;
        k->d_keep = nullptr;
        k->cap_bytes = 0;
        PSB_CUDA(cudaMalloc(&k->d_bits, bytes));
        if (missing) PSB_CUDA(cudaMalloc(&k->d_miss, bytes));
        PSB_CUDA(cudaMalloc(&k->d_XT, (size_t)n_chunks * k->Npad * 4));
        PSB_CUDA(cudaMalloc(&k->d_keep, (size_t)n_variants));
        k->cap_bytes = bytes;
    }
    PSB_CUDA(cudaMemcpyAsync(k->d_bits, bits, bytes, cudaMemcpyHostToDevice, c->stream));
    if (missing) PSB_CUDA(cudaMemcpyAsync(k->d_miss, missing, bytes, cudaMemcpyHostToDevice, c->stream));
    PSB_CUDA(cudaMemsetAsync(k->d_XT, 0, (size_t)n_chunks * k->Npad * 4, c->stream));
    const int blocks = (int)std::min<int64_t>((n_variants + 7) / 8, (int64_t)c->sm_count * 16);
    k_kin_keep<<<blocks, 256, 0, c->stream>>>(k->d_bits, missing ? k->d_miss : nullptr, n_variants,
                                             words_per_row, k->Wn, k->N, min_af, max_af, max_missing,
                                             k->d_keep);
    const int tb = (int)std::min<int64_t>((n_chunks * k->Wn + 7) / 8, (int64_t)c->sm_count * 16);
    k_kin_transpose<<<tb, 256, 0, c->stream>>>(k->d_bits, k->d_keep, n_variants, words_per_row, k->Wn,
                                               k->Npad, n_chunks, k->d_XT);
    const int nt = k->Npad / KIN_TILE;
    k_kin_accum<<<nt * (nt + 1) / 2, 256, 0, c->stream>>>(k->d_XT, n_chunks, k->Npad, k->N, nt, k->d_K);
    c->launches += 3;
    PSB_CUDA(cudaGetLastError());
    PSB_CUDA(cudaStreamSynchronize(c->stream));      // the host buffers may be reused
    k->seen += n_variants;
    return PSB_OK;
}

extern "C" int psb_kinship_fetch(psb_ctx *c, double *K_out) {
    PSB_REQUIRE(c && c->kin && K_out, PSB_ERR_STATE, "psb_kinship_fetch before psb_kinship_begin");
    psb_kin *k = (psb_kin *)c->kin;
    PSB_CUDA(cudaSetDevice(c->device));
    const size_t n2 = (size_t)k->N * k->N;
    std::vector<long long> h(n2);
    PSB_CUDA(cudaMemcpyAsync(h.data(), k->d_K, n2 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    const int N = k->N;
    for (int i = 0; i < N; ++i)
        for (int j = i; j < N; ++j) {
            // every (i, j) with j >= i lies in a computed tile (tj >= ti); mirror it
            const long long v = h[(size_t)i * N + j];
            K_out[(size_t)i * N + j] = (double)v;
            K_out[(size_t)j * N + i] = (double)v;
        }
    return PSB_OK;
}
