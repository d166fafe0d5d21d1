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
//
// Tensor path (default; PSB_KIN_TC=0 selects the CUDA-core kernels above, which stay as the
// agreement check of tests/test_kinship_gpu.py):
//   k_kin_expand     kept variants of a chunk -> E[sample][variant] int8 0/1 (sample-major, the
//                    contraction axis contiguous: the K-major operand layout of tcgen05.mma)
//   k_kin_tc         K[I tile][J tile] += E_I E_J' : tcgen05.mma.kind::i8, M = N = 128, K = 32, both
//                    operands TMA-staged (128-byte swizzle) in a ring of shared-memory stages,
//                    int32 accumulators in tensor memory, upper-triangular tiles only, split over
//                    the variant axis when there are fewer tiles than SMs
// 0/1 operands and int32 accumulation: the same exact integers as the popcount form.
#include <algorithm>
#include <vector>

#include "psb_internal.cuh"
#include "psb_tc_ptx.cuh"

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


// ---------------------------------------------------------------------------------------
// tensor path
// ---------------------------------------------------------------------------------------
#define KT_TILE 128          // samples per tile side (M and N of one MMA)
#define KT_KSTAGE 128        // variants per pipeline stage = one 128-byte swizzle row
#define KT_STAGES 6
#define KT_THREADS 192       // warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: epilogue
#define KT_BOX_BYTES (KT_TILE * KT_KSTAGE)
#define KT_VCHUNK 65536      // variants expanded per pass (E = Npad x 64 KiB)

// One block: 128 variants x 1024 samples.  Rows are read coalesced (a warp = 32 consecutive
// words of one variant), written as 16-byte pieces of the samples' 128-byte K segments.
// Byte vg * 16 + m of a segment holds variant m * 8 + vg of the block: the contraction does
// not care about the order of the variants as long as both operands share it (they are the
// same matrix), and this order keeps the shared-memory reads of a warp on distinct banks.
__global__ void __launch_bounds__(256)
k_kin_expand(const uint32_t *__restrict__ bits, const uint8_t *__restrict__ keep, int64_t v_base,
             int64_t S, int Wrow, int Wn, int Npad, int Vpad, uint8_t *__restrict__ E) {
    __shared__ uint32_t sm[128][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t v0 = v_base + (int64_t)blockIdx.x * 128;
    const int w0 = blockIdx.y * 32;
    for (int r = warp; r < 128; r += 8) {
        const int64_t v = v0 + r;
        const int w = w0 + lane;
        uint32_t x = 0;
        if (v < S && w < Wn && keep[v]) x = __ldg(bits + v * Wrow + w);
        sm[r][lane] = x;
    }
    __syncthreads();
    for (int it = threadIdx.x; it < 8192; it += 256) {
        const int vg = it & 7, s = it >> 3;
        const int row = w0 * 32 + s;
        if (row >= Npad) break;
        const int w = s >> 5, b = s & 31;
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t acc = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc |= ((sm[(q * 4 + k) * 8 + vg][w] >> b) & 1u) << (8 * k);
            o[q] = acc;
        }
        *reinterpret_cast<uint4 *>(E + (size_t)row * Vpad + (size_t)blockIdx.x * 128 + vg * 16) =
            make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32
__device__ __forceinline__ void tc_mma_i8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct KtArgs {
    int N, n_tiles, n_stages, nsplit;
    unsigned long long *K;
};

__global__ void __launch_bounds__(KT_THREADS, 1)
k_kin_tc(const __grid_constant__ CUtensorMap tmap, const KtArgs a) {
    // instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, N = 128, M = 128
    constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KT_TILE >> 3) << 17) |
                               ((uint32_t)(KT_TILE >> 4) << 24);
    extern __shared__ __align__(1024) uint8_t kt_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)kt_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                                        // KT_STAGES boxes
    uint8_t *sB = smem + KT_STAGES * KT_BOX_BYTES;             // KT_STAGES boxes
    uint64_t *full = (uint64_t *)(smem + 2 * KT_STAGES * KT_BOX_BYTES);
    uint64_t *empty = full + KT_STAGES;
    uint64_t *accFull = empty + KT_STAGES;
    uint32_t *tmem_slot = (uint32_t *)(accFull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work item: upper-triangular tile (ti, tj >= ti) and its range of K stages
    int t = (int)blockIdx.x / a.nsplit, ti = 0;
    const int g = (int)blockIdx.x - t * a.nsplit;
    while (t >= a.n_tiles - ti) { t -= a.n_tiles - ti; ++ti; }
    const int tj = ti + t;
    const bool diag = ti == tj;                                // one box serves both operands
    const int k_lo = (int)((long long)a.n_stages * g / a.nsplit);
    const int k_hi = (int)((long long)a.n_stages * (g + 1) / a.nsplit);

    if (threadIdx.x == 0) {
        for (int i = 0; i < KT_STAGES; ++i) {
            mbar_init(smem_u32(&full[i]), 1);
            mbar_init(smem_u32(&empty[i]), 1);
        }
        mbar_init(smem_u32(accFull), 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(KT_TILE));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
    const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB);

    if (warp == 0) {
        // ===================== TMA producer =====================
        int st = 0;
        uint32_t ph = 0;
        for (int ks = k_lo; ks < k_hi; ++ks) {
            mbar_wait(empty0 + st * 8, ph ^ 1);
            if (elect_one()) {
                const uint32_t bar = full0 + st * 8;
                mbar_arrive_expect_tx(bar, diag ? KT_BOX_BYTES : 2 * KT_BOX_BYTES);
                tma_load_2d(sA0 + st * KT_BOX_BYTES, &tmap, bar, ks * KT_KSTAGE, ti * KT_TILE);
                if (!diag) tma_load_2d(sB0 + st * KT_BOX_BYTES, &tmap, bar, ks * KT_KSTAGE, tj * KT_TILE);
            }
            __syncwarp();
            if (++st == KT_STAGES) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int st = 0;
        uint32_t ph = 0;
        for (int ks = k_lo; ks < k_hi; ++ks) {
            mbar_wait(full0 + st * 8, ph);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t adesc = make_b_desc(sA0 + st * KT_BOX_BYTES);
                const uint64_t bdesc = diag ? adesc : make_b_desc(sB0 + st * KT_BOX_BYTES);
#pragma unroll
                for (int kk = 0; kk < KT_KSTAGE / 32; ++kk)
                    tc_mma_i8_ss(tmem_base, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), IDESC,
                                 (ks != k_lo || kk != 0) ? 1u : 0u);
                tc_commit(empty0 + st * 8);
                if (ks == k_hi - 1) tc_commit(smem_u32(accFull));
            }
            __syncwarp();
            if (++st == KT_STAGES) { st = 0; ph ^= 1; }
        }
    } else {
        // ===================== epilogue (warps 2..5: TMEM lane quarter warp % 4) =====================
        const int q = warp & 3;
        const int i = ti * KT_TILE + q * 32 + lane;            // sample of this thread's TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        mbar_wait(smem_u32(accFull), 0);
        tc_fence_after();
        for (int c0 = 0; c0 < KT_TILE; c0 += 16) {
            int32_t r[16];
            tc_ld16(lane_addr + c0, r);
            tc_wait_ld();
            if (i < a.N) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int j = tj * KT_TILE + c0 + c;
                    if (j >= i && j < a.N && r[c] != 0)
                        atomicAdd(a.K + (size_t)i * a.N + j, (unsigned long long)r[c]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(KT_TILE));
    }
}

typedef CUresult (*PFN_kinEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                       const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                       const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct psb_kin {
    int N = 0, Wn = 0, Npad = 0, NpadT = 0;
    long long *d_K = nullptr;
    uint32_t *d_bits = nullptr, *d_miss = nullptr, *d_XT = nullptr;
    uint8_t *d_keep = nullptr, *d_E = nullptr;
    size_t cap_bytes = 0, cap_xt = 0, cap_e = 0, cap_keep = 0;
    int64_t kept = 0, seen = 0;
};

static void kin_free(psb_kin *k) {
    if (!k) return;
    cudaFree(k->d_K); cudaFree(k->d_bits); cudaFree(k->d_miss); cudaFree(k->d_XT); cudaFree(k->d_keep);
    cudaFree(k->d_E);
    delete k;
}

// K += E E' over the kept variants of the rows in d_bits, chunk by chunk through the tensor cores
static int kin_accum_tc(psb_ctx *c, psb_kin *k, const uint32_t *d_rows, int64_t n_variants, int words_per_row) {
    static void *fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult qres;
        PSB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        PSB_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, PSB_ERR_CUDA,
                    "cuTensorMapEncodeTiled not available from the driver");
    }
    const int64_t vmax = std::min<int64_t>(n_variants, KT_VCHUNK);
    const size_t need = (size_t)k->NpadT * (size_t)((vmax + KT_KSTAGE - 1) / KT_KSTAGE * KT_KSTAGE);
    if (need > k->cap_e) {
        PSB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(k->d_E);
        k->d_E = nullptr;
        k->cap_e = 0;
        PSB_CUDA(cudaMalloc(&k->d_E, need));
        k->cap_e = need;
    }
    const size_t smem = 1024 + 2 * KT_STAGES * KT_BOX_BYTES + (2 * KT_STAGES + 1) * 8 + 16;
    PSB_CUDA(cudaFuncSetAttribute(k_kin_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nt = k->NpadT / KT_TILE, tiles = nt * (nt + 1) / 2;
    for (int64_t v0 = 0; v0 < n_variants; v0 += KT_VCHUNK) {
        const int64_t vc = std::min<int64_t>(n_variants - v0, KT_VCHUNK);
        const int n_stages = (int)((vc + KT_KSTAGE - 1) / KT_KSTAGE);
        const int Vpad = n_stages * KT_KSTAGE;
        k_kin_expand<<<dim3(n_stages, (k->NpadT + 1023) / 1024), 256, 0, c->stream>>>(
            d_rows, k->d_keep, v0, n_variants, words_per_row, k->Wn, k->NpadT, Vpad, k->d_E);
        CUtensorMap tm;
        cuuint64_t gdim[2] = {(cuuint64_t)Vpad, (cuuint64_t)k->NpadT};
        cuuint64_t gstr[1] = {(cuuint64_t)Vpad};
        cuuint32_t box[2] = {KT_KSTAGE, KT_TILE};
        cuuint32_t estr[2] = {1, 1};
        CUresult cr = ((PFN_kinEncodeTiled)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, k->d_E, gdim, gstr, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PSB_REQUIRE(cr == CUDA_SUCCESS, PSB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
        KtArgs a;
        a.N = k->N;
        a.n_tiles = nt;
        a.n_stages = n_stages;
        // fewer tiles than two waves of SMs: deal the variant axis of a tile to several CTAs
        a.nsplit = std::max(1, std::min(n_stages, (2 * c->sm_count + tiles - 1) / tiles));
        if (tiles >= 2 * c->sm_count) a.nsplit = 1;
        a.K = (unsigned long long *)k->d_K;
        k_kin_tc<<<tiles * a.nsplit, KT_THREADS, smem, c->stream>>>(tm, a);
        c->launches += 2;
        PSB_CUDA(cudaGetLastError());
    }
    return PSB_OK;
}

// K += (kept rows)' (kept rows) for n_variants packed rows already on the device
static int kin_accumulate(psb_ctx *c, psb_kin *k, const uint32_t *d_rows, const uint32_t *d_miss, int64_t n_variants,
                          int words_per_row, double min_af, double max_af, double max_missing) {
    const int64_t n_chunks = (n_variants + 31) / 32;
    // test hook: PSB_KIN_TC=0 keeps the contraction on the CUDA cores (AND + POPCOUNT)
    const bool use_tc = !(getenv("PSB_KIN_TC") && atoi(getenv("PSB_KIN_TC")) == 0);
    if ((size_t)n_variants > k->cap_keep) {
        PSB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(k->d_keep);
        k->d_keep = nullptr;
        k->cap_keep = 0;
        PSB_CUDA(cudaMalloc(&k->d_keep, (size_t)n_variants));
        k->cap_keep = (size_t)n_variants;
    }
    const int blocks = (int)std::min<int64_t>((n_variants + 7) / 8, (int64_t)c->sm_count * 16);
    k_kin_keep<<<blocks, 256, 0, c->stream>>>(d_rows, d_miss, n_variants, words_per_row, k->Wn, k->N, min_af, max_af,
                                             max_missing, k->d_keep);
    c->launches += 1;
    if (use_tc) {
        int rc = kin_accum_tc(c, k, d_rows, n_variants, words_per_row);
        if (rc != PSB_OK) return rc;
    } else {
        const size_t xt_bytes = (size_t)n_chunks * k->Npad * 4;
        if (xt_bytes > k->cap_xt) {
            PSB_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(k->d_XT);
            k->d_XT = nullptr;
            k->cap_xt = 0;
            PSB_CUDA(cudaMalloc(&k->d_XT, xt_bytes));
            k->cap_xt = xt_bytes;
        }
        PSB_CUDA(cudaMemsetAsync(k->d_XT, 0, xt_bytes, c->stream));
        const int tb = (int)std::min<int64_t>((n_chunks * k->Wn + 7) / 8, (int64_t)c->sm_count * 16);
        k_kin_transpose<<<tb, 256, 0, c->stream>>>(d_rows, k->d_keep, n_variants, words_per_row, k->Wn, k->Npad,
                                                   n_chunks, k->d_XT);
        const int nt = k->Npad / KIN_TILE;
        k_kin_accum<<<nt * (nt + 1) / 2, 256, 0, c->stream>>>(k->d_XT, n_chunks, k->Npad, k->N, nt, k->d_K);
        c->launches += 2;
    }
    PSB_CUDA(cudaGetLastError());
    return PSB_OK;
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
    k->NpadT = ((n_samples + KT_TILE - 1) / KT_TILE) * KT_TILE;
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
    if (bytes > k->cap_bytes || (missing && !k->d_miss)) {
        PSB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(k->d_bits); cudaFree(k->d_miss);
        k->d_bits = k->d_miss = nullptr;
        k->cap_bytes = 0;
        PSB_CUDA(cudaMalloc(&k->d_bits, bytes));
        if (missing) PSB_CUDA(cudaMalloc(&k->d_miss, bytes));
        k->cap_bytes = bytes;
    }
    PSB_CUDA(cudaMemcpyAsync(k->d_bits, bits, bytes, cudaMemcpyHostToDevice, c->stream));
    if (missing) PSB_CUDA(cudaMemcpyAsync(k->d_miss, missing, bytes, cudaMemcpyHostToDevice, c->stream));
    int rc = kin_accumulate(c, k, k->d_bits, missing ? k->d_miss : nullptr, n_variants, words_per_row, min_af, max_af,
                            max_missing);
    if (rc != PSB_OK) return rc;
    PSB_CUDA(cudaGetLastError());
    PSB_CUDA(cudaStreamSynchronize(c->stream));      // the host buffers may be reused
    k->seen += n_variants;
    return PSB_OK;
}

// The same for the rows of the batch last submitted to the context (psb_submit / psb_submit_text /
// psb_submit_device): no host rows at all when the k-mer text was tokenised on the device.
extern "C" int psb_kinship_add_submitted(psb_ctx *c, double min_af, double max_af, double max_missing) {
    PSB_REQUIRE(c && c->kin, PSB_ERR_STATE, "psb_kinship_add_submitted before psb_kinship_begin");
    psb_kin *k = (psb_kin *)c->kin;
    PSB_CUDA(cudaSetDevice(c->device));
    int rc = psb_run_begin(c);                       // adopt the submitted rows, order the stream after their copy
    if (rc) return rc;
    PSB_REQUIRE(c->d_bits || c->S == 0, PSB_ERR_STATE, "no rows submitted");
    PSB_REQUIRE(c->N == k->N && c->Wrow >= k->Wn, PSB_ERR_ARG, "the submitted rows are over %d samples, the matrix over %d",
                c->N, k->N);
    if (c->S == 0) return PSB_OK;
    rc = kin_accumulate(c, k, c->d_bits, c->d_miss, c->S, c->Wrow, min_af, max_af, max_missing);
    if (rc != PSB_OK) return rc;
    rc = psb_run_end(c);                             // the staging slot has been read
    if (rc) return rc;
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    k->seen += c->S;
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
