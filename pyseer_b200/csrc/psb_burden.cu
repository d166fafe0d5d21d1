// psb_burden.cu -- burden regions on the device.
//
// Replaces the region branch of input.read_variant (input.py:395-411): a burden "variant" is
// the union of every VCF record the region(s) of one burden-file line fetch, accumulated in
// one dictionary by read_vcf_var (input.py:457-502, dominant encoding).  In terms of the
// per-record rows (x = carries a non-reference allele, m = genotype missing), a sample of the
// region row is
//      x_region = OR over the member records of x                      (1 is absorbing)
//      m_region = m of the LAST member record, unless x_region          (input.py:489-497: a
//                 called reference haplotype deletes an earlier NaN, a '.' only sets NaN for a
//                 sample that is not yet in the dictionary)
// Member lists come as CSR (region_offsets / members) in fetch order; a record may belong to
// several regions and appear more than once.
//
// One thread owns one 16-byte chunk of one region row and streams the same chunk of every
// member row: adjacent threads read adjacent chunks, so every member row is read once,
// coalesced, and the region row is written once.  HBM-bound byte work:
// (members + 1) * N/8 bytes per region (+ the same again when a missing matrix is given).
#include "psb_internal.cuh"

__global__ void __launch_bounds__(256)
k_burden_or(const uint4 *__restrict__ vbits, const uint4 *__restrict__ vmiss,
            const int64_t *__restrict__ offs, const int32_t *__restrict__ members,
            uint4 *__restrict__ out_bits, uint4 *__restrict__ out_miss, int64_t n_regions,
            int chunks) {
    const int64_t total = n_regions * chunks;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / chunks;
        const int ch = (int)(e - r * chunks);
        const int64_t m0 = offs[r], m1 = offs[r + 1];
        uint4 acc = make_uint4(0u, 0u, 0u, 0u);
        int64_t m = m0;
        // four independent 16-byte loads in flight per thread
        for (; m + 4 <= m1; m += 4) {
            const uint4 a = __ldg(vbits + (size_t)members[m] * chunks + ch);
            const uint4 b = __ldg(vbits + (size_t)members[m + 1] * chunks + ch);
            const uint4 c = __ldg(vbits + (size_t)members[m + 2] * chunks + ch);
            const uint4 d = __ldg(vbits + (size_t)members[m + 3] * chunks + ch);
            acc.x |= a.x | b.x | c.x | d.x;
            acc.y |= a.y | b.y | c.y | d.y;
            acc.z |= a.z | b.z | c.z | d.z;
            acc.w |= a.w | b.w | c.w | d.w;
        }
        for (; m < m1; ++m) {
            const uint4 a = __ldg(vbits + (size_t)members[m] * chunks + ch);
            acc.x |= a.x; acc.y |= a.y; acc.z |= a.z; acc.w |= a.w;
        }
        out_bits[e] = acc;
        if (out_miss) {
            uint4 last = make_uint4(0u, 0u, 0u, 0u);
            if (m1 > m0) last = __ldg(vmiss + (size_t)members[m1 - 1] * chunks + ch);
            last.x &= ~acc.x; last.y &= ~acc.y; last.z &= ~acc.z; last.w &= ~acc.w;
            out_miss[e] = last;
        }
    }
}

static int reserve(psb_ctx *c, void **buf, size_t *cap, size_t bytes) {
    if (bytes <= *cap) return PSB_OK;
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->copy_stream));
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    PSB_CUDA(cudaMalloc(buf, bytes > 0 ? bytes : 16));
    *cap = bytes > 0 ? bytes : 16;
    return PSB_OK;
}

void psb_burden_release(psb_ctx *c) {
    if (c->bur_vbits) cudaFree(c->bur_vbits);
    if (c->bur_vmiss) cudaFree(c->bur_vmiss);
    if (c->bur_out) cudaFree(c->bur_out);
    if (c->bur_outmiss) cudaFree(c->bur_outmiss);
    if (c->bur_offs) cudaFree(c->bur_offs);
    if (c->bur_members) cudaFree(c->bur_members);
    if (c->ev_bur_copy) cudaEventDestroy(c->ev_bur_copy);
    if (c->ev_bur_done) cudaEventDestroy(c->ev_bur_done);
    c->bur_vbits = c->bur_vmiss = c->bur_out = c->bur_outmiss = nullptr;
    c->bur_offs = nullptr;
    c->bur_members = nullptr;
    c->ev_bur_copy = c->ev_bur_done = nullptr;
    c->bur_vbits_cap = c->bur_vmiss_cap = c->bur_out_cap = c->bur_outmiss_cap = 0;
    c->bur_offs_cap = c->bur_members_cap = 0;
}

// Shared tail of the two entry points: member rows are on the device (d_vbits / d_vmiss,
// ready once `ready` has fired, or already ordered on the compute stream when ready is null).
static int burden_reduce(psb_ctx *c, const uint32_t *d_vbits, const uint32_t *d_vmiss,
                         int64_t n_variants, int32_t wpr, const int64_t *region_offsets,
                         const int32_t *members, int64_t n_regions, bool rows_on_copy_stream) {
    PSB_REQUIRE(n_regions >= 0 && n_regions < (1ll << 31) - 512, PSB_ERR_ARG,
                "n_regions %lld out of range", (long long)n_regions);
    PSB_REQUIRE(region_offsets || n_regions == 0, PSB_ERR_ARG, "region_offsets is NULL");
    const int64_t n_members = n_regions > 0 ? region_offsets[n_regions] : 0;
    PSB_REQUIRE(n_regions == 0 || region_offsets[0] == 0, PSB_ERR_ARG, "region_offsets[0] must be 0");
    for (int64_t r = 0; r < n_regions; ++r)
        PSB_REQUIRE(region_offsets[r + 1] >= region_offsets[r], PSB_ERR_ARG,
                    "region_offsets must be non-decreasing (region %lld)", (long long)r);
    PSB_REQUIRE(members || n_members == 0, PSB_ERR_ARG, "members is NULL");
    for (int64_t m = 0; m < n_members; ++m)
        PSB_REQUIRE(members[m] >= 0 && members[m] < n_variants, PSB_ERR_ARG,
                    "member %lld refers to record %d of %lld", (long long)m, members[m],
                    (long long)n_variants);
    if (!c->ev_bur_copy) {
        PSB_CUDA(cudaEventCreateWithFlags(&c->ev_bur_copy, cudaEventDisableTiming));
        PSB_CUDA(cudaEventCreateWithFlags(&c->ev_bur_done, cudaEventDisableTiming));
    }
    const size_t out_bytes = (size_t)n_regions * wpr * sizeof(uint32_t);
    int rc = reserve(c, (void **)&c->bur_out, &c->bur_out_cap, out_bytes);
    if (rc) return rc;
    if (d_vmiss) {
        rc = reserve(c, (void **)&c->bur_outmiss, &c->bur_outmiss_cap, out_bytes);
        if (rc) return rc;
    }
    rc = reserve(c, (void **)&c->bur_offs, &c->bur_offs_cap, (size_t)(n_regions + 1) * sizeof(int64_t));
    if (rc) return rc;
    rc = reserve(c, (void **)&c->bur_members, &c->bur_members_cap, (size_t)n_members * sizeof(int32_t));
    if (rc) return rc;
    // the member lists are small next to the rows: they ride the copy stream behind them
    if (n_regions > 0)
        PSB_CUDA(cudaMemcpyAsync(c->bur_offs, region_offsets, (size_t)(n_regions + 1) * sizeof(int64_t),
                                 cudaMemcpyHostToDevice, c->copy_stream));
    if (n_members > 0)
        PSB_CUDA(cudaMemcpyAsync(c->bur_members, members, (size_t)n_members * sizeof(int32_t),
                                 cudaMemcpyHostToDevice, c->copy_stream));
    (void)rows_on_copy_stream;
    PSB_CUDA(cudaEventRecord(c->ev_bur_copy, c->copy_stream));
    PSB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_bur_copy, 0));
    if (n_regions > 0) {
        const int chunks = wpr / 4;
        const int64_t total = n_regions * chunks;
        int64_t want = (total + 255) / 256;
        const int64_t cap = (int64_t)c->sm_count * 8;     // 8 resident CTAs of 256 threads per SM
        const int grid = (int)(want < cap ? want : cap);
        k_burden_or<<<grid, 256, 0, c->stream>>>(
            (const uint4 *)d_vbits, (const uint4 *)d_vmiss, c->bur_offs, c->bur_members,
            (uint4 *)c->bur_out, d_vmiss ? (uint4 *)c->bur_outmiss : nullptr, n_regions, chunks);
        c->launches++;
        PSB_CUDA(cudaGetLastError());
    }
    PSB_CUDA(cudaEventRecord(c->ev_bur_done, c->stream));
    c->bur_done_valid = true;
    // the member lists were read from pageable or caller memory: make sure the staged copies
    // have left the host buffers before returning (the ABI only borrows them for the call)
    PSB_CUDA(cudaEventSynchronize(c->ev_bur_copy));
    c->sub_bits = c->bur_out;
    c->sub_miss = d_vmiss ? c->bur_outmiss : nullptr;
    c->sub_S = n_regions;
    c->sub_Wrow = wpr;
    c->sub_slot = -1;
    c->sub_valid = true;
    return PSB_OK;
}

static int burden_check(psb_ctx *c, int64_t n_variants, int32_t wpr) {
    PSB_REQUIRE(c, PSB_ERR_ARG, "ctx is NULL");
    PSB_REQUIRE(c->model != PSB_MODEL_NONE, PSB_ERR_STATE,
                "psb_submit_burden before psb_lmm_setup / psb_fixed_setup");
    PSB_REQUIRE(n_variants >= 0 && n_variants < (1ll << 31) - 512, PSB_ERR_ARG,
                "n_variants %lld out of range", (long long)n_variants);
    PSB_REQUIRE(wpr >= c->Wn && wpr % 4 == 0, PSB_ERR_ARG,
                "words_per_row %d must be a multiple of 4 and >= ceil(N/32) = %d", wpr, c->Wn);
    return PSB_OK;
}

extern "C" int psb_submit_burden(psb_ctx *c, const uint32_t *bits, const uint32_t *missing,
                                 int64_t n_variants, int32_t words_per_row,
                                 const int64_t *region_offsets, const int32_t *members,
                                 int64_t n_regions) {
    int rc = burden_check(c, n_variants, words_per_row);
    if (rc) return rc;
    PSB_REQUIRE(bits || n_variants == 0, PSB_ERR_ARG, "bits is NULL");
    PSB_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)n_variants * words_per_row * sizeof(uint32_t);
    rc = reserve(c, (void **)&c->bur_vbits, &c->bur_vbits_cap, bytes);
    if (rc) return rc;
    if (missing) {
        rc = reserve(c, (void **)&c->bur_vmiss, &c->bur_vmiss_cap, bytes);
        if (rc) return rc;
    }
    // the previous reduction must have finished reading the member rows before they are replaced
    if (c->bur_done_valid) PSB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_bur_done, 0));
    if (bytes) {
        PSB_CUDA(cudaMemcpyAsync(c->bur_vbits, bits, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        if (missing)
            PSB_CUDA(cudaMemcpyAsync(c->bur_vmiss, missing, bytes, cudaMemcpyHostToDevice, c->copy_stream));
    }
    return burden_reduce(c, c->bur_vbits, missing ? c->bur_vmiss : nullptr, n_variants,
                         words_per_row, region_offsets, members, n_regions, true);
}

extern "C" int psb_submit_burden_device(psb_ctx *c, const void *d_bits, const void *d_missing,
                                        int64_t n_variants, int32_t words_per_row,
                                        const int64_t *region_offsets, const int32_t *members,
                                        int64_t n_regions) {
    int rc = burden_check(c, n_variants, words_per_row);
    if (rc) return rc;
    PSB_REQUIRE(d_bits || n_variants == 0, PSB_ERR_ARG, "d_bits is NULL");
    PSB_CUDA(cudaSetDevice(c->device));
    if (c->bur_done_valid) PSB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_bur_done, 0));
    return burden_reduce(c, (const uint32_t *)d_bits, (const uint32_t *)d_missing, n_variants,
                         words_per_row, region_offsets, members, n_regions, false);
}

// Device pointer of the rows left submitted by the last psb_submit* / psb_synth_device (bench:
// synthetic member rows are generated on the device, then reduced with psb_submit_burden_device).
extern "C" int psb_submitted_device(psb_ctx *c, const void **d_bits, const void **d_missing,
                                    int64_t *n_variants, int32_t *words_per_row) {
    PSB_REQUIRE(c, PSB_ERR_ARG, "ctx is NULL");
    const bool sub = c->sub_valid;
    PSB_REQUIRE(sub || c->d_bits, PSB_ERR_STATE, "no rows submitted");
    if (d_bits) *d_bits = sub ? c->sub_bits : c->d_bits;
    if (d_missing) *d_missing = sub ? c->sub_miss : c->d_miss;
    if (n_variants) *n_variants = sub ? c->sub_S : c->S;
    if (words_per_row) *words_per_row = sub ? c->sub_Wrow : c->Wrow;
    return PSB_OK;
}
