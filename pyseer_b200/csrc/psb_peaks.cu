// psb_peaks.cu -- measured peak rates of the two pipes the hot kernels are bound by, for the
// roofline denominators of bench.py (MEASURED_PEAKS.json only carries an HBM copy rate and a bf16
// GEMM rate; SURVEY 8d asks for the fraction against the dtype actually executed):
//
//   int8 tensor   tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 256, K = 32, A from tensor
//                 memory, B from shared memory (the operand forms k_lmm_quadform_tc uses), issued
//                 back to back on all-ones operands with nothing else in the way: no TMA, no
//                 expansion, no epilogue.  The accumulators are read back once and checked
//                 (every entry = 32 x number of MMAs), so the rate belongs to MMAs that ran.
//   fp64 FMA      8 independent DFMA chains per thread, 8 x 256 threads per SM (k_fixed_logit's pipe).
//
// Both run under whatever clocks the board holds at that moment (bench.py calls this right after the
// timed region, while the power state is the one the hot kernel saw).
#include <math.h>

#include <vector>

#include "psb_internal.cuh"
#include "psb_tc_ptx.cuh"

#define PK_N 256
#define PK_THREADS 128

__global__ void __launch_bounds__(PK_THREADS, 1) k_peak_i8(int iters, int32_t *__restrict__ out) {
    extern __shared__ __align__(1024) uint8_t pk_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)pk_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sB = smem;                                        // PK_N rows x 128 bytes, K-major
    uint64_t *bar = (uint64_t *)(sB + PK_N * 128);
    uint32_t *tmem_slot = (uint32_t *)(bar + 1);
    const int warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < PK_N * 128 / 4; e += PK_THREADS) ((uint32_t *)sB)[e] = 0x01010101u;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bar), 1);
        fence_barrier_init();
    }
    fence_proxy_async();                                       // generic writes -> async proxy (UMMA reads)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    // A: 128 samples of ones per variant (lane) = 32 TMEM columns after the accumulator
    const uint32_t ones[8] = {0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u,
                              0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u};
    for (int c = 0; c < 32; c += 8) tc_st8(lane_addr + PK_N + c, ones);
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PK_N >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
    if (warp == 0) {
        if (elect_one()) {
            const uint64_t bdesc = make_b_desc(smem_u32(sB));
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    tc_mma_i8_ts(tmem_base, tmem_base + PK_N + kk * 8, bdesc + (uint64_t)(kk * 2), IDESC,
                                 (it != 0 || kk != 0) ? 1u : 0u);
            }
            tc_commit(smem_u32(bar));
        }
        __syncwarp();
    }
    mbar_wait(smem_u32(bar), 0);
    tc_fence_after();
    int32_t r[8];
    tc_ld8(lane_addr, r);
    tc_wait_ld();
    out[(size_t)blockIdx.x * PK_THREADS + threadIdx.x] = (r[0] == r[7] && r[0] == r[3]) ? r[0] : -1;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

__global__ void __launch_bounds__(256) k_peak_f64(int iters, double x, double y, double *__restrict__ out) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
        a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" int psb_measure_peaks(psb_ctx *c, double out[4]) {
    PSB_REQUIRE(c && out, PSB_ERR_ARG, "NULL argument");
    PSB_CUDA(cudaSetDevice(c->device));
    out[0] = out[1] = out[2] = out[3] = 0.0;
    cudaEvent_t e0, e1;
    PSB_CUDA(cudaEventCreate(&e0));
    PSB_CUDA(cudaEventCreate(&e1));
    const int grid = c->sm_count;
    // ---- int8 tensor ----
    {
        const int iters = 40000;                       // x 4 MMAs of 128 x 256 x 32
        int32_t *d_out = nullptr;
        PSB_CUDA(cudaMalloc(&d_out, (size_t)grid * PK_THREADS * sizeof(int32_t)));
        const size_t smem = PK_N * 128 + 1024 + 64;
        PSB_CUDA(cudaFuncSetAttribute(k_peak_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {            // first launch warms up
            PSB_CUDA(cudaEventRecord(e0, c->stream));
            k_peak_i8<<<grid, PK_THREADS, smem, c->stream>>>(iters, d_out);
            PSB_CUDA(cudaEventRecord(e1, c->stream));
            PSB_CUDA(cudaEventSynchronize(e1));
            PSB_CUDA(cudaGetLastError());
            c->launches++;
            float ms = 0.f;
            PSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        std::vector<int32_t> h((size_t)grid * PK_THREADS);
        PSB_CUDA(cudaMemcpy(h.data(), d_out, h.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        cudaFree(d_out);
        const int32_t want = 32 * 4 * iters;
        for (size_t i = 0; i < h.size(); ++i)
            PSB_REQUIRE(h[i] == want, PSB_ERR_NUMERIC, "int8 peak probe: accumulator %zu = %d, expected %d", i,
                        h[i], want);
        const double ops = 2.0 * 128.0 * PK_N * 32.0 * 4.0 * iters * grid;
        out[0] = ops / (best * 1e-3) / 1e12;           // TOP/s
        out[2] = best;
    }
    // ---- fp64 FMA ----
    {
        const int iters = 20000, blocks = grid * 8;
        double *d_out = nullptr;
        PSB_CUDA(cudaMalloc(&d_out, (size_t)blocks * 256 * sizeof(double)));
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            PSB_CUDA(cudaEventRecord(e0, c->stream));
            k_peak_f64<<<blocks, 256, 0, c->stream>>>(iters, 0.999999, 1e-6, d_out);
            PSB_CUDA(cudaEventRecord(e1, c->stream));
            PSB_CUDA(cudaEventSynchronize(e1));
            PSB_CUDA(cudaGetLastError());
            c->launches++;
            float ms = 0.f;
            PSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        cudaFree(d_out);
        out[1] = 2.0 * 8.0 * iters * (double)blocks * 256.0 / (best * 1e-3) / 1e12;   // TFLOP/s
        out[3] = best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return PSB_OK;
}
