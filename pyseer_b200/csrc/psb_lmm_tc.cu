// psb_lmm_tc.cu -- the LMM rotation / quadratic form on the 5th-generation tensor cores.
//
// Replaces the dense contraction of fastlmm nLLeval: rotate() `U.T.dot(P snps)`
// (lmm_cov.py:165-194) followed by computeAKA `(Usnps / Sd * Usnps).sum(0)`
// (lmm_cov.py:885-899), i.e. for every tested variant x (a 0/1 column)
//        a = || L' x ||^2 ,      L = P U diag(Sd^-1/2)   (N x J, J = N - D).
//
// Exactness.  x is 0/1, so L'x is a sum of selected rows of L.  Every column j of L is
// scaled by a power of two and rounded to a (8k-2)-bit integer, which is split into k
// balanced base-256 digits d_s in [-128,127] ("slices").  int8 x {0,1} products
// accumulate exactly in int32 (|sum| <= 128 N), and the epilogue recombines
// g_j = sum_s 256^s D_s in fp64 -- exact below 2^53 -- so the only error left is the
// rounding of L to 8k-2 bits relative to its column maximum (k = 5: 2^-38).  This is an
// integer GEMM by construction; it is not a reduced-precision approximation of an fp GEMM.
//
// Kernel structure (one persistent CTA per SM, 128 variants per tile, 14 warps):
//   warp 0      TMA producer: streams the k-sliced int8 operand (B, K-major, 128B swizzle)
//               through an smem ring with cp.async.bulk.tensor + mbarrier complete_tx.
//   warp 1      MMA issuer: one thread issues tcgen05.mma.kind::i8, M=128, N=32k, K=32;
//               A (the variants) comes from TENSOR MEMORY, B from shared memory,
//               accumulators (int32) live in TMEM, double buffered.
//   warps 2-9   expanders (two groups of four warps taking alternate K stages): each thread
//               owns one variant (= one TMEM lane); it turns its packed presence/absence
//               bits (staged once per tile in smem, 1 bit/sample) into 0/1 bytes in
//               registers -- (w >> s) & 0x01010101, two ALU ops per 4 samples, with the
//               matching sample permutation baked into the B operand -- and writes them
//               straight into TMEM with tcgen05.st: the expanded operand never touches
//               shared memory or HBM.
//   warps 10-13 epilogue: tcgen05.ld the int32 accumulators, recombine the slices in fp64,
//               square, scale and accumulate a[v] in a register across all component tiles.
// One extra component tile carries v = M y (score numerator b = x'v, computeAKB
// lmm_cov.py:902-916) and the covariate basis Q (for rotate()'s constant-column rule), each
// as a hi + lo pair of sliced columns, i.e. to 2 x (8k-2) bits: b and ||Q'x||^2 come out of
// the same pass at fp64 accuracy.
// HBM sees N/8 bytes per variant; L2 serves the sliced operand to all CTAs.
#include <cuda.h>

#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "psb_internal.cuh"

#define TC_THREADS 448
#define TC_EXP_WARPS 8         // expander warps (2 groups x 4 lane quarters)
#define TC_TILE_V 128          // variants per CTA tile (UMMA M)
#define TC_JT 32               // components per accumulator tile
#define TC_KBOX 128            // samples per TMA box / swizzle row (128 bytes of int8)
#define TC_KSTAGE 256          // samples per pipeline stage: two boxes, eight K = 32 MMAs.  A stage
                               // must carry enough tensor work (8 x 80 cycles at k = 5) to cover the
                               // per-stage control path of the issuing warp (~370 cycles measured:
                               // barrier poll, elect, descriptor arithmetic, commit) -- with 128-sample
                               // stages that path, not the tensor pipe, set the pace.
#define TC_MAX_BSTAGES 6

#include "psb_tc_ptx.cuh"

struct TcArgs {
    const uint32_t *bits;     // packed rows
    const int32_t *idx;       // tested variant ids
    const double *scale2;     // per component (s_j)^2; plain s_j for the special tile
    double *a_out;            // [variant id]  || L'x ||^2
    double *b_out;            // [variant id]  x'v            (special tile; may be null)
    double *pp_out;           // [variant id]  || Q'x ||^2    (special tile; may be null)
    int n_special;            // hi/lo column pairs in the special tile (0 = no special tile)
    double *lin_out;          // raw pair sums x'u_p -> lin_out[variant id * lin_ld + p] (may be null)
    int lin_ld;
    double *w_out;            // Welch sums over carriers (pairs n_special, n_special + 1 of the special
    int w_ld;                 // tile: yc, yc^2) -> w_out[variant id * w_ld + {0, 1}] (null: not carried)
    int Wrow;                 // words per packed row in global memory
    int n_tested;             // upper bound used for the grid; the kernels read the real count:
    const int *n_tested_dev;  // device counter written by k_prefilter (no host round trip)
    int nks;                  // K stages (Kpad / 128)
    int jtiles;
    int pitch;                // smem bit-row pitch in words (== 4 mod 32)
    int bits_in_smem;         // 1: the tile's packed rows are staged in shared memory
    int tri;                  // 1: regular tiles hold the triangular operand M'' (see below)
    int int_epi;              // 1: triangular tiles are recombined and summed in int64 (below)
    const uint8_t *shift;     // per-column left shift for int_epi
    int stages_per_tile;      // K stages a variant tile goes through (all component tiles)
    int n_bstages;
    int pair;                 // launched as clusters of two CTAs: 1 = each CTA issues its own MMAs, the
                              // operand stream is shared through TMA multicast; 2 = two-SM MMAs
                              // (cta_group::2), each CTA holds half of every B stage
    int split_max;            // > 1: component tiles of a variant tile may be dealt to up to this many units
    double *a_part;           // [split_max][part_ld] partial quadratic forms of a split launch
    int part_ld;
    int skip_box;             // triangular form: skip the all-zero first box of a diagonal stage
    int debug;                // timing experiments only (PSB_TC_DEBUG, results are wrong when set):
                              // 1: one MMA per stage instead of four; 2: expanders store without
                              // expanding; 4: epilogue releases the accumulators without reading them
};

// Sequence of component tiles a variant tile goes through, and the first K stage of each.
// Triangular form (tri): a = x'Mx with M = L L' symmetric and x binary equals
//   sum_j x_j T_j,   T_j = sum_{i >= j} M''_ij x_i,   M''_jj = M_jj, M''_ij = 2 M_ij (i > j),
// so column tile jt only needs the samples i >= 32 jt: K stages below (32 jt) / 128 are skipped
// and the contraction costs N^2/2 instead of N (N - D) multiply-adds per variant.  Long and
// short tiles alternate so that the epilogue of a short one hides behind the next long one.
__device__ __forceinline__ void tc_tile_of(const TcArgs &a, int q, int &jt, int &ks0) {
    const int nreg = a.jtiles - (a.n_special > 0 ? 1 : 0);
    if (q >= nreg || !a.tri) {
        jt = q;
        ks0 = 0;
        return;
    }
    jt = (q & 1) ? (nreg - 1 - (q >> 1)) : (q >> 1);
    ks0 = (jt * TC_JT) / TC_KSTAGE;
}

// how many units share the component tiles of one variant tile (1: no split)
__host__ __device__ __forceinline__ int tc_split_of(int split_max, int n_tile_units, int n_units, int jtiles) {
    int G = 1;
    if (split_max > 1 && n_tile_units > 0 && n_tile_units * 2 <= n_units) {
        G = n_units / n_tile_units;
        if (G > split_max) G = split_max;
        if (G > jtiles / 4) G = jtiles / 4 > 0 ? jtiles / 4 : 1;
    }
    return G;
}

struct TcStageIter {
    int q, ks, jt;
    __device__ __forceinline__ void init(const TcArgs &a) {
        q = 0;
        int k0;
        tc_tile_of(a, 0, jt, k0);
        ks = k0;
    }
    __device__ __forceinline__ bool valid(const TcArgs &a) const { return q < a.jtiles; }
    __device__ __forceinline__ void next(const TcArgs &a) {
        if (++ks >= a.nks) {
            ++q;
            if (q < a.jtiles) {
                int k0;
                tc_tile_of(a, q, jt, k0);
                ks = k0;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------
template <int NSL, bool TWO>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_lmm_quadform_tc(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_half,
                  const TcArgs args) {
    constexpr int UMMA_N = TC_JT * NSL;                 // accumulator columns per buffer
    constexpr int A_COL0 = 2 * UMMA_N;                  // first TMEM column of the A ring
    constexpr int A_COLS = TC_KSTAGE / 4;               // TMEM columns of one A stage (4 samples each)
    constexpr int NA = (512 - A_COL0) / A_COLS;         // A ring stages
    constexpr uint32_t BOX_BYTES = UMMA_N * TC_KBOX;    // one TMA box: UMMA_N rows x 128 samples
    constexpr uint32_t STAGE_BYTES = 2 * BOX_BYTES;
    static_assert(NA >= 1, "too many slices for the TMEM budget");
    // instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, N, M = 128
    constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UMMA_N >> 3) << 17) |
                               ((uint32_t)(TC_TILE_V >> 4) << 24);
    // two-SM MMA: M = 256 over the CTA pair
    constexpr uint32_t IDESC2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UMMA_N >> 3) << 17) |
                                ((uint32_t)((2 * TC_TILE_V) >> 4) << 24);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int ns = args.n_bstages;                 // operand (B) ring depth; may exceed NA
    const int na = ns < NA ? ns : NA;              // A ring depth in use
    // Two-SM mode (args.pair == 2): the pair issues ONE MMA of M = 256 per K step -- 128 variants
    // (TMEM lanes) in each CTA -- whose B operand is split between the two shared memories: each
    // CTA loads and keeps only HALF of the sliced rows of every stage, so the bytes entering
    // each SM (64 B/clk L2 port) and the shared memory per stage halve.  CTA 0 issues the MMAs;
    // both CTAs' TMA bytes and expander arrivals complete on CTA 0's full barriers, its commits
    // are multicast to both CTAs' empty / accFull barriers, both epilogues release on CTA 0's
    // accEmpty barriers.
    constexpr bool two = TWO;
    const uint32_t box_bytes = two ? BOX_BYTES / 2 : BOX_BYTES;       // per CTA
    const uint32_t stage_bytes = 2 * box_bytes;
    uint8_t *sB = smem;                                               // ns stages
    uint32_t *sBits = (uint32_t *)(sB + (size_t)ns * stage_bytes);    // 128 x pitch words
    uint64_t *bars = (uint64_t *)(sBits + (size_t)TC_TILE_V * args.pitch);
    // One barrier pair per K stage in flight, indexed by the B slot t % ns of stage t: full[s]
    // completes when the TMA bytes of the B stage have landed AND the four expander warps have
    // stored the A stage in TMEM; empty[s] is signalled by tcgen05.commit once the stage's MMAs
    // have retired.  The A stage of stage t lives in TMEM slot t % NA.  The B ring may be
    // deeper than the A ring (ns >= NA): the producer refills B slot t % ns after the commit
    // of stage t - ns, the expanders rewrite A slot t % NA after the commit of stage t - NA,
    // which they observe on the same barrier array (empty[(t - NA) % ns]) -- one commit per
    // stage serves both.  The extra depth covers the L2 latency of the TMA loads.
    uint64_t *full = bars;                        // [ns]
    uint64_t *empty = full + 16;                  // [ns]
    uint64_t *accFull = empty + 16;               // [2]
    uint64_t *accEmpty = accFull + 2;             // [2]
    uint32_t *tmem_slot = (uint32_t *)(accEmpty + 2);
    int32_t *sRows = (int32_t *)(tmem_slot + 2);  // [128]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tested = args.n_tested_dev ? *args.n_tested_dev : args.n_tested;
    const int n_tiles = (n_tested + TC_TILE_V - 1) / TC_TILE_V;
    // Pair mode: CTAs 2i and 2i+1 form a cluster.  Both need the same operand stream (only
    // their variants differ), so each loads HALF of every B stage and multicasts it into the
    // shared memory of both: L2 serves every stage once per pair instead of once per CTA.  The
    // two CTAs therefore walk the stage sequence in lockstep -- a B slot is refilled only after
    // the MMAs of BOTH have retired (each commit arrives on the empty barrier of both) -- and
    // run the same number of variant tiles: the even CTA's tile index decides, the odd
    // CTA may run a last tile without variants.
    const int pair = args.pair;
    const uint32_t cta_rank = pair ? cluster_ctarank() : 0u;
    const int odd = pair ? (int)(blockIdx.x & 1u) : 0;     // tile - odd = the even CTA's tile
    // Work items.  A unit is a CTA (or a CTA pair, which walks its two tiles in lockstep); with
    // enough variant tiles an item is a whole tile (pair of tiles).  When there are FEWER tiles than
    // units -- the refinement pass of the two-pass mode (the few variants with F > 30), small batches
    // -- the component tiles of a variant tile are dealt to G units, each adding its part of
    // a = x'M''x with one atomicAdd per variant (a_out is zeroed by the host wrapper), so that a
    // handful of variants costs a G-th of a tile's latency instead of all of it.  The special tile
    // (last of the sequence) belongs to the last group, which alone writes b, ||Q'x||^2 and the sums.
    const int unit0 = pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_units = pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int n_tile_units = pair ? (n_tiles + 1) / 2 : n_tiles;
    const int G = tc_split_of(args.split_max, n_tile_units, n_units, args.jtiles);
    const int n_items = n_tile_units * G;
    // item -> this CTA's tile and its range [q_lo, q_hi) of the component-tile sequence
    auto item_of = [&](int item, int &tile, int &q_lo, int &q_hi) {
        const int tu = item / G, g = item - tu * G;
        tile = pair ? 2 * tu + odd : tu;
        q_lo = args.jtiles * g / G;
        q_hi = args.jtiles * (g + 1) / G;
    };
    // K stages of the component tiles [q_lo, q_hi)
    auto item_stages = [&](int q_lo, int q_hi) -> int {
        if (G == 1) return args.stages_per_tile;
        int n = 0;
        for (int q = q_lo; q < q_hi; ++q) {
            int jt, ks0;
            tc_tile_of(args, q, jt, ks0);
            n += args.nks - ks0;
        }
        return n;
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < ns; ++i) {
            mbar_init(smem_u32(&full[i]), two ? 1 + 4 + 4 : 1 + 4);
            mbar_init(smem_u32(&empty[i]), pair == 1 ? 2 : 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&accFull[i]), 1);
            mbar_init(smem_u32(&accEmpty[i]), two ? 8 : 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (two) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(tmem_slot)),
                         "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(tmem_slot)),
                         "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (pair) cluster_sync_all();      // the peer's barriers exist before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
    const uint32_t sB0 = smem_u32(sB);
    // CTA 0's full / accEmpty barriers as seen from this CTA (two-SM mode)
    const uint32_t lead_full0 = two ? mapa_shared(full0, 0u) : full0;
    const uint32_t lead_accEmpty0 = two ? mapa_shared(smem_u32(accEmpty), 0u) : smem_u32(accEmpty);

    if (warp == 0) {
        // ===================== TMA producer (whole warp loops, one lane issues) ==========
        int st = 0;
        uint32_t ph = 0;
        for (int item = unit0; item < n_items; item += n_units) {
            int tile_u, q_lo, q_hi;
            item_of(item, tile_u, q_lo, q_hi);
            for (int q = q_lo; q < q_hi; ++q) {
                int jt, ks0;
                tc_tile_of(args, q, jt, ks0);
                for (int ks = ks0; ks < args.nks; ++ks) {
                    mbar_wait(empty0 + st * 8, ph ^ 1);
                    if (elect_one()) {
                        const uint32_t bar = full0 + st * 8;
                        if constexpr (two) {
                            // this CTA's half of the rows of both boxes; bytes complete on CTA 0
                            if (cta_rank == 0) mbar_arrive_expect_tx(bar, 2 * stage_bytes);
                            const uint32_t lbar = lead_full0 + st * 8;
                            const int row = jt * UMMA_N + (int)cta_rank * (UMMA_N / 2);
                            tma_load_2d_cg2(sB0 + st * stage_bytes, &tmap_half, lbar, ks * TC_KSTAGE, row);
                            tma_load_2d_cg2(sB0 + st * stage_bytes + box_bytes, &tmap_half, lbar,
                                            ks * TC_KSTAGE + TC_KBOX, row);
                        } else if (pair) {
                            mbar_arrive_expect_tx(bar, STAGE_BYTES);
                            // each CTA of the pair fetches one of the stage's two boxes for both
                            tma_load_2d_mc(sB0 + st * STAGE_BYTES + cta_rank * BOX_BYTES, &tmap, bar,
                                           ks * TC_KSTAGE + (int)cta_rank * TC_KBOX, jt * UMMA_N, (uint16_t)3);
                        } else {
                            mbar_arrive_expect_tx(bar, STAGE_BYTES);
                            tma_load_2d(sB0 + st * STAGE_BYTES, &tmap, bar, ks * TC_KSTAGE, jt * UMMA_N);
                            tma_load_2d(sB0 + st * STAGE_BYTES + BOX_BYTES, &tmap, bar, ks * TC_KSTAGE + TC_KBOX,
                                        jt * UMMA_N);
                        }
                    }
                    __syncwarp();
                    if (++st == ns) { st = 0; ph ^= 1; }
                }
            }
        }
        // Pair mode: the peer's last commits arrive on this CTA's empty barriers; wait for them
        // so that nothing is in flight towards this CTA's shared memory when it exits.
        if (pair) {
            long long total = 0;
            for (int item = unit0; item < n_items; item += n_units) {
                int tile_u, q_lo, q_hi;
                item_of(item, tile_u, q_lo, q_hi);
                total += item_stages(q_lo, q_hi);
            }
            const int tail = total < ns ? (int)total : ns;
            for (int i = 0; i < tail; ++i) {
                mbar_wait(empty0 + st * 8, ph ^ 1);
                if (++st == ns) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp loops, one lane issues) ============
        // (two-SM mode: CTA 0 issues for the pair, CTA 1's warp only owns its TMEM allocation)
        int st = 0, sta = 0, acc = 0;
        uint32_t ph = 0, phacc = 0;
        const int jt_first_special = args.jtiles - (args.n_special > 0 ? 1 : 0);
        if (!(two && cta_rank != 0))
        for (int item = unit0; item < n_items; item += n_units) {
            int tile_u, q_lo, q_hi;
            item_of(item, tile_u, q_lo, q_hi);
            for (int q = q_lo; q < q_hi; ++q) {
                int jt, ks0;
                tc_tile_of(args, q, jt, ks0);
                mbar_wait(smem_u32(&accEmpty[acc]), phacc ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * UMMA_N);
                for (int ks = ks0; ks < args.nks; ++ks) {
                    mbar_wait(full0 + st * 8, ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t bdesc = make_b_desc(sB0 + st * stage_bytes);
                        const uint32_t a_tmem = tmem_base + (uint32_t)(A_COL0 + sta * A_COLS);
                        const int nkk = (args.debug & 1) ? 2 : 8;
                        // triangular form: the first stage of a column tile starts at a multiple of 256
                        // samples; when the tile's first sample lies in its second box, the first box
                        // (four MMAs) only meets zeros of M'' and is skipped
                        const int kk0 = (args.skip_box && ks == ks0 && jt < jt_first_special &&
                                         ((jt * TC_JT) & (TC_KSTAGE - 1)) >= TC_KBOX) ? 4 : 0;
                        if constexpr (two) {
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                if (kk < nkk && kk >= kk0)
                                    tc_mma_i8_ts_2(d_tmem, a_tmem + kk * 8,
                                                   bdesc + (uint64_t)((kk >> 2) * ((BOX_BYTES / 2) >> 4) + (kk & 3) * 2),
                                                   IDESC2, (ks != ks0 || kk != kk0) ? 1u : 0u);
                            tc_commit_2mc(empty0 + st * 8, (uint16_t)3);
                            if (ks == args.nks - 1) tc_commit_2mc(smem_u32(&accFull[acc]), (uint16_t)3);
                        } else {
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                if (kk < nkk && kk >= kk0)
                                    tc_mma_i8_ts(d_tmem, a_tmem + kk * 8,
                                                 bdesc + (uint64_t)((kk >> 2) * (BOX_BYTES >> 4) + (kk & 3) * 2), IDESC,
                                                 (ks != ks0 || kk != kk0) ? 1u : 0u);
                            if (pair) tc_commit_mc(empty0 + st * 8, (uint16_t)3);
                            else tc_commit(empty0 + st * 8);
                            if (ks == args.nks - 1) tc_commit(smem_u32(&accFull[acc]));
                        }
                    }
                    __syncwarp();
                    if (++st == ns) { st = 0; ph ^= 1; }
                    if (++sta == na) sta = 0;
                }
                if (++acc == 2) { acc = 0; phacc ^= 1; }
            }
        }
    } else if (warp < 2 + TC_EXP_WARPS) {
        // ===================== expanders (warps 2..9) =====================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int grp = (warp - 2) >> 2;              // group 0 / 1: even / odd K stages
        const int v = q * 32 + lane;                  // variant (= TMEM lane) within the tile
        const int et = (warp - 2) * 32 + lane;        // 0..255 cooperative-copy index
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int chunks_per_row = 2 * args.nks;      // 16-byte chunks (4 words = 128 samples)
        // The group takes every second stage t of the global sequence.  sb: barrier (B) slot
        // t % ns of its next stage; sa: A slot t % na; (we, wph): slot and phase of
        // empty[(t - na) % ns], the commit that frees the A slot -- not waited for during the
        // first na stages of the kernel (`lead` of them belong to this group).
        int sb = grp % ns, sa = grp % na;
        int lead = (na - grp + 1) / 2;
        int we = grp + 2 * lead - na;
        uint32_t wph = 0;
        uint32_t parity = 0;                          // parity of the global stage counter
        const int nks = args.nks, jtiles = args.jtiles;
        const int nreg_x = (args.tri != 0) ? jtiles - (args.n_special > 0 ? 1 : 0) : 0;
        // first K stage of the q-th component tile of the sequence (tc_tile_of)
        auto first_ks = [&](int q2) -> int {
            if (q2 >= nreg_x) return 0;
            const int jt = (q2 & 1) ? (nreg_x - 1 - (q2 >> 1)) : (q2 >> 1);
            return (jt * TC_JT) / TC_KSTAGE;
        };
        // two stages forward in the sequence; the cursor is valid while cq < q_end (the end of the
        // item's component-tile range)
        int q_end = jtiles;
        auto advance2 = [&](int &cq, int &cks) {
            cks += 2;
            while (cks >= nks) {
                const int over = cks - nks;
                if (++cq >= q_end) return;
                cks = first_ks(cq) + over;
            }
        };
        for (int item = unit0; item < n_items; item += n_units) {
            int tile, q_lo;
            item_of(item, tile, q_lo, q_end);
            // all expanders are done reading the previous tile's bits
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et < TC_TILE_V) {
                int t = tile * TC_TILE_V + et;
                sRows[et] = t < n_tested ? args.idx[t] : -1;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (args.bits_in_smem) {
                const int total = TC_TILE_V * chunks_per_row;
                for (int e = et; e < total; e += 32 * TC_EXP_WARPS) {
                    int r = e / chunks_per_row, ch = e - r * chunks_per_row;
                    int row = sRows[r];
                    uint4 val = make_uint4(0u, 0u, 0u, 0u);
                    if (row >= 0 && ch * 4 < args.Wrow)
                        val = __ldg(reinterpret_cast<const uint4 *>(args.bits + (size_t)row * args.Wrow) + ch);
                    *reinterpret_cast<uint4 *>(sBits + (size_t)r * args.pitch + ch * 4) = val;
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // Large N (the tile's bits would crowd the operand ring out of shared memory): each
            // thread reads its 16 bytes per stage straight from its global row (L2-resident
            // after the first component tile), fetched one of its stages ahead.
            const int grow_id = sRows[v];
            const uint4 *grow = reinterpret_cast<const uint4 *>(args.bits + (size_t)(grow_id < 0 ? 0 : grow_id) * args.Wrow);
            const int gwords4 = args.Wrow >> 2;
            auto gload = [&](int ch) -> uint4 {       // ch: 16-byte chunk of the row
                return (grow_id >= 0 && ch < gwords4) ? __ldg(grow + ch) : make_uint4(0u, 0u, 0u, 0u);
            };
            // This group takes every second stage of the tile's (component tile, K stage)
            // sequence, starting at the first one whose global parity matches the group.  The
            // cursor (cq, cks) steps two stages at a time; leaving a component tile (rare: once
            // per ~20 stages) carries the overshoot into the next one.  Everything the loop
            // needs from `args` sits in registers: the per-stage control flow is a handful of
            // instructions next to the 64 ALU operations of the expansion itself.
            int cq = q_lo, cks = first_ks(q_lo) + (((int)parity != grp) ? 1 : 0) - 2;
            advance2(cq, cks);
            // expand one 128-sample stage of this thread's variant into A slot sa, then signal
            auto emit = [&](const uint4 &w4a, const uint4 &w4b) {
                const uint32_t ws[8] = {w4a.x, w4a.y, w4a.z, w4a.w, w4b.x, w4b.y, w4b.z, w4b.w};
                // register 8 i + b, byte t <- sample 8 t + b of word i (the B operand is stored
                // with the same permutation, see k_tc_quantise).  The expansion does not depend
                // on the A slot being free, so it runs BEFORE the wait: what is left on the
                // critical path between the commit that frees the slot and this stage's arrival
                // is one 32-register tcgen05.st, the expansion of the second half and its store.
                uint32_t r[32];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int b = 0; b < 8; ++b) r[i * 8 + b] = (ws[i] >> b) & 0x01010101u;
                if (lead > 0) {
                    --lead;
                } else {
                    mbar_wait(empty0 + we * 8, wph);
                    we += 2;
                    while (we >= ns) { we -= ns; wph ^= 1u; }
                }
                tc_fence_after();
                const uint32_t base = lane_addr + (uint32_t)(A_COL0 + sa * A_COLS);
                tc_st32(base, r);
                // second half: the store above has read its registers when it issued
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int b = 0; b < 8; ++b) r[i * 8 + b] = (ws[4 + i] >> b) & 0x01010101u;
                tc_st32(base + 32, r);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (two) mbar_arrive_cluster(lead_full0 + sb * 8);
                    else mbar_arrive(full0 + sb * 8);
                }
                sb += 2;
                while (sb >= ns) sb -= ns;
                sa += 2;
                while (sa >= na) sa -= na;
            };
            if (args.bits_in_smem) {
                const uint32_t myrow_s = smem_u32(sBits + (size_t)v * args.pitch);
                while (cq < q_end) {
                    const uint4 w4a = lds128(myrow_s + (uint32_t)cks * 32u);
                    const uint4 w4b = lds128(myrow_s + (uint32_t)cks * 32u + 16u);
                    emit(w4a, w4b);
                    advance2(cq, cks);
                }
            } else {
                // global-row mode: the words of the group's next two stages are in flight
                int pq = cq, pks = cks;
                const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
                uint4 nxa = zero4, nxb = zero4, nx2a = zero4, nx2b = zero4;
                if (pq < q_end) { nxa = gload(2 * pks); nxb = gload(2 * pks + 1); }
                advance2(pq, pks);
                if (pq < q_end) { nx2a = gload(2 * pks); nx2b = gload(2 * pks + 1); }
                while (cq < q_end) {
                    const uint4 w4a = nxa, w4b = nxb;
                    nxa = nx2a;
                    nxb = nx2b;
                    if (pq < q_end) {
                        advance2(pq, pks);
                        if (pq < q_end) { nx2a = gload(2 * pks); nx2b = gload(2 * pks + 1); }
                    }
                    emit(w4a, w4b);
                    advance2(cq, cks);
                }
            }
            parity ^= (uint32_t)(item_stages(q_lo, q_end) & 1);
        }
    } else {
        // ===================== epilogue (warps 10..13) =====================
        constexpr int CH = NSL <= 5 ? 16 : 8;         // accumulator columns per TMEM load
        const int q = warp & 3;
        const int v = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int jt_special = args.n_special > 0 ? args.jtiles - 1 : -1;
        int acc = 0;
        uint32_t phacc = 0;
        for (int item = unit0; item < n_items; item += n_units) {
            int tile, q_lo, q_hi;
            item_of(item, tile, q_lo, q_hi);
            double a = 0.0, bsum = 0.0, pp = 0.0;
            const int t_own = tile * TC_TILE_V + v;
            const int row_own = t_own < n_tested ? args.idx[t_own] : -1;
            const uint32_t *grow = args.bits + (size_t)(row_own < 0 ? 0 : row_own) * args.Wrow;
            for (int q2 = q_lo; q2 < q_hi; ++q2) {
                int jt, ks0_unused;
                tc_tile_of(args, q2, jt, ks0_unused);
                // triangular form: column j of this tile counts only if the variant carries
                // sample j -- the tile's 32 samples are word jt of the variant's packed row
                uint32_t xword = 0u;
                if (args.tri && jt != jt_special && row_own >= 0) xword = __ldg(grow + jt);
                mbar_wait(smem_u32(&accFull[acc]), phacc);
                tc_fence_after();
                const uint32_t col0 = (uint32_t)(acc * UMMA_N);
                if (args.debug & 4) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (two) mbar_arrive_cluster(lead_accEmpty0 + acc * 8);
                        else mbar_arrive(smem_u32(&accEmpty[acc]));
                    }
                    if (++acc == 2) { acc = 0; phacc ^= 1; }
                    continue;
                }
                if (args.int_epi && jt != jt_special) {
                    // Triangular tile, integer epilogue: the columns of a tile share one scale
                    // 2^(e_tile - shmax - B) up to a small per-column left shift, so the slices
                    // are recombined (g = sum_s 256^s D_s) and the carried columns summed in
                    // int64 -- exact, no per-column int->fp64 conversion -- and the tile costs
                    // one conversion and one fused multiply-add.
                    long long t64 = 0;
#pragma unroll
                    for (int c0 = 0; c0 < TC_JT; c0 += CH) {
                        int32_t d[NSL][CH];
#pragma unroll
                        for (int s = 0; s < NSL; ++s) {
                            if constexpr (CH == 16) tc_ld16(lane_addr + col0 + s * TC_JT + c0, d[s]);
                            else tc_ld8(lane_addr + col0 + s * TC_JT + c0, d[s]);
                        }
                        uint32_t shw[CH / 4];
                        if constexpr (CH == 16) {
                            const uint4 s4 = __ldg(reinterpret_cast<const uint4 *>(args.shift + jt * TC_JT + c0));
                            shw[0] = s4.x; shw[1] = s4.y; shw[2] = s4.z; shw[3] = s4.w;
                        } else {
                            const uint2 s2 = __ldg(reinterpret_cast<const uint2 *>(args.shift + jt * TC_JT + c0));
                            shw[0] = s2.x; shw[1] = s2.y;
                        }
                        tc_wait_ld();
#pragma unroll
                        for (int c = 0; c < CH; ++c) {
                            // pairs of slices fit int32 (|D_s| <= 128 N, N <= 16384)
                            long long g = 0;
                            int s = NSL - 1;
                            if (NSL & 1) { g = (long long)d[s][c]; --s; }
#pragma unroll
                            for (; s >= 1; s -= 2)
                                g = g * 65536 + (long long)(d[s - 1][c] + d[s][c] * 256);
                            const uint32_t sh = (shw[c >> 2] >> (8 * (c & 3))) & 0xffu;
                            if ((xword >> (c0 + c)) & 1u) t64 += g << sh;
                        }
                    }
                    a = fma((double)t64, __ldg(args.scale2 + jt * TC_JT), a);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (two) mbar_arrive_cluster(lead_accEmpty0 + acc * 8);
                        else mbar_arrive(smem_u32(&accEmpty[acc]));
                    }
                    if (++acc == 2) { acc = 0; phacc ^= 1; }
                    continue;
                }
                double prev = 0.0;                    // hi part of the current hi/lo pair
#pragma unroll
                for (int c0 = 0; c0 < TC_JT; c0 += CH) {
                    int32_t d[NSL][CH];
#pragma unroll
                    for (int s = 0; s < NSL; ++s) {
                        if constexpr (CH == 16) tc_ld16(lane_addr + col0 + s * TC_JT + c0, d[s]);
                        else tc_ld8(lane_addr + col0 + s * TC_JT + c0, d[s]);
                    }
                    tc_wait_ld();
                    const double *sc = args.scale2 + jt * TC_JT + c0;
                    if (jt != jt_special) {
#pragma unroll
                        for (int c = 0; c < CH; ++c) {
                            double g = (double)d[NSL - 1][c];
#pragma unroll
                            for (int s = NSL - 2; s >= 0; --s) g = fma(g, 256.0, (double)d[s][c]);
                            if (args.tri) {
                                if ((xword >> (c0 + c)) & 1u) a = fma(g, __ldg(sc + c), a);
                            } else {
                                a = fma(g * g, __ldg(sc + c), a);
                            }
                        }
                    } else {
                        // columns come in (hi, lo) pairs: pair 0 = v, pairs 1.. = Q_e
#pragma unroll
                        for (int c = 0; c < CH; ++c) {
                            double g = (double)d[NSL - 1][c];
#pragma unroll
                            for (int s = NSL - 2; s >= 0; --s) g = fma(g, 256.0, (double)d[s][c]);
                            g *= __ldg(sc + c);
                            if ((c & 1) == 0) {
                                prev = g;
                            } else {
                                double val = prev + g;
                                int pair = (c0 + c) >> 1;
                                if (pair == 0) bsum = val;
                                else if (pair <= args.n_special - 1) pp = fma(val, val, pp);
                                else if (args.w_out && row_own >= 0 && pair <= args.n_special + 1)
                                    args.w_out[(size_t)row_own * args.w_ld + (pair - args.n_special)] = val;
                                if (args.lin_out && row_own >= 0 && pair < args.n_special)
                                    args.lin_out[(size_t)row_own * args.lin_ld + pair] = val;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                        if (two) mbar_arrive_cluster(lead_accEmpty0 + acc * 8);
                        else mbar_arrive(smem_u32(&accEmpty[acc]));
                    }
                if (++acc == 2) { acc = 0; phacc ^= 1; }
            }
            if (row_own >= 0) {
                if (args.a_out) {
                    if (G == 1) args.a_out[row_own] = a;
                    else args.a_part[(size_t)(item % G) * args.part_ld + t_own] = a;   // summed by k_tc_split_reduce
                }
                if (jt_special >= 0 && args.b_out && q_hi == args.jtiles) {
                    args.b_out[row_own] = bsum;
                    args.pp_out[row_own] = pp;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (pair) cluster_sync_all();      // nobody leaves while the peer may still signal or read here
    if (warp == 1) {
        tc_fence_after();
        if constexpr (two)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------------------------------
// Quantisation of L into k balanced base-256 digits, K-major:
//   Lq[(jt * k + s) * 32 + c][kpos(i)] = digit s of rint(L[i][jt*32+c] * 2^(8k-2-e_j)),
// e_j = exponent with max_i |L[i][j]| < 2^e_j;  scale2[j] = 2^(2 (e_j - 8k + 2)).
// kpos permutes the samples inside every group of 32 so that the expanders can produce
// register b of a word with (w >> b) & 0x01010101: byte t of that register is sample
// 8 t + b, and it is operand position 4 b + t.
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int tc_kpos(int i) {
    int i32 = i & 31;
    return (i & ~31) | (((i32 & 7) << 2) | (i32 >> 3));
}

// value of the operand at (sample i, column j): L itself, or (tri) the triangular M''
__device__ __forceinline__ double tc_src(const double *__restrict__ A, int ld, int i, int j, int tri) {
    const double x = A[(size_t)i * ld + j];
    if (!tri) return x;
    return i > j ? 2.0 * x : (i == j ? x : 0.0);
}

// shmax >= 0: integer epilogue (triangular form).  The 32 columns of a tile share the scale
// 2^(e_tile - shmax - B), e_tile = the largest column exponent of the tile; column j is
// quantised with e'_j = max(e_j, e_tile - shmax) and enters the int64 tile sum shifted left by
// e'_j - (e_tile - shmax) in [0, shmax]: full 8k-2 bit precision relative to its own maximum
// for columns within 2^shmax of the tile's largest, graceful below.  The tile scale goes to
// scale2[32 jt].  One warp = one tile (blockDim is a multiple of 32).
__global__ void k_tc_colexp(const double *__restrict__ L, int N, int Jpad, int J, int nsl,
                            int *__restrict__ expo, double *__restrict__ scale2, int Jq, int tri,
                            int shmax, uint8_t *__restrict__ shift) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Jq) return;                      // Jq is a multiple of 32: whole warps leave
    double mx = 0.0;
    if (j < J)
        for (int i = 0; i < N; ++i) mx = fmax(mx, fabs(tc_src(L, Jpad, i, j, tri)));
    int e = 0;
    if (mx > 0.0) {
        frexp(mx, &e);          // mx = f * 2^e, f in [0.5, 1)  =>  mx < 2^e
    }
    if (shmax >= 0) {
        const int INT_MIN_E = -100000;
        const int et = __reduce_max_sync(0xffffffffu, mx > 0.0 ? e : INT_MIN_E);
        const int base = et - shmax;          // exponent of the tile's unit
        const int ej = (mx > 0.0 && e > base) ? e : base;
        expo[j] = et == INT_MIN_E ? 0 : ej;
        shift[j] = (uint8_t)(et == INT_MIN_E ? 0 : ej - base);
        if ((threadIdx.x & 31) == 0)
            scale2[j] = et == INT_MIN_E ? 0.0 : ldexp(1.0, base - (8 * nsl - 2));
        return;
    }
    expo[j] = e;
    // regular tiles: squared scale (a += g^2 s^2); triangular tiles: plain scale (a += g s)
    scale2[j] = (mx > 0.0) ? ldexp(1.0, (tri ? 1 : 2) * (e - (8 * nsl - 2))) : 0.0;
}

// M = L L' (N x N, fp64) for the triangular form; 64 x 64 tiles, 4 x 4 outputs per thread
__global__ void __launch_bounds__(256)
k_tc_syrk(const double *__restrict__ L, int Lrows, int Jpad, double *__restrict__ M, int ldm) {
    __shared__ double As[16][64 + 1], Bs[16][64 + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < Jpad; k0 += 16) {
        __syncthreads();
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int r = e >> 4, kk = e & 15;       // row within the tile, k within the chunk
            As[kk][r] = (i0 + r < Lrows) ? L[(size_t)(i0 + r) * Jpad + k0 + kk] : 0.0;
            Bs[kk][r] = (j0 + r < Lrows) ? L[(size_t)(j0 + r) * Jpad + k0 + kk] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { a[r] = As[kk][ty * 4 + r]; b[r] = Bs[kk][tx * 4 + r]; }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = i0 + ty * 4 + r, j = j0 + tx * 4 + c;
            if (i < ldm && j < ldm) M[(size_t)i * ldm + j] = acc[r][c];
        }
}

// Uniform [0, 1) from a counter (splitmix64 finaliser): the dither of the stochastic rounding below.
__device__ __forceinline__ double tc_u01(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// Rounding to the (8k-2)-bit grid is STOCHASTIC (floor(v + u), u uniform and keyed by the entry):
// unbiased, so the rounding errors of the ~N^2/4 entries a quadratic form sums add up like a random
// walk whatever the matrix looks like.  Round-to-nearest is biased on structured operands: with
// h2 = 0 every off-diagonal entry of M'' is the same number (-2/N) and takes the same rounding
// error, which a variant with c carriers collects c^2/2 times -- measured on a clonal kinship at
// N = 5000: p-values off by 1e-4 relative at k = 4 (4e-7 at k = 5) with round-to-nearest, 1e-9 with
// the dither.  PSB_TC_DITHER=0 restores round-to-nearest.
__global__ void k_tc_quantise(const double *__restrict__ L, int N, int Jpad, int J, int nsl,
                              const int *__restrict__ expo, int8_t *__restrict__ Lq, int Kpad,
                              int Jq, int tri, int dither) {
    // thread per (i, j): i fastest so that the int8 stores of a warp are contiguous
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)Jq * Kpad;
    if (e >= total) return;
    int j = (int)(e / Kpad);
    int i = (int)(e - (size_t)j * Kpad);
    long long qv = 0;
    if (i < N && j < J) {
        double x = tc_src(L, Jpad, i, j, tri);
        const double v = ldexp(x, (8 * nsl - 2) - expo[j]);
        qv = dither ? (long long)floor(v + tc_u01((uint64_t)e)) : llrint(v);
    }
    int jt = j / TC_JT, c = j - jt * TC_JT;
    for (int s = 0; s < nsl; ++s) {
        long long d = ((qv + 128) & 255) - 128;       // balanced digit in [-128, 127]
        qv = (qv - d) >> 8;
        Lq[((size_t)(jt * nsl + s) * TC_JT + c) * Kpad + tc_kpos(i)] = (int8_t)d;
    }
}

// ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static size_t tc_smem_bytes(int nsl, int nb, int pitch, int mode) {
    size_t stage = (size_t)TC_JT * nsl * TC_KSTAGE / (mode == 2 ? 2 : 1);   // per CTA
    return 1024 + (size_t)nb * stage + (size_t)TC_TILE_V * pitch * 4 +
           (32 + 4) * 8 + 16 + TC_TILE_V * 4 + 64;
}

template <int NSL, bool TWO>
static int tc_launch2(psb_ctx *c, const TcArgs &args, int grid, size_t smem) {
    PSB_CUDA(cudaFuncSetAttribute(k_lmm_quadform_tc<NSL, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = args.pair ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PSB_CUDA(cudaLaunchKernelEx(&cfg, k_lmm_quadform_tc<NSL, TWO>, *(const CUtensorMap *)c->tmap_Lq,
                                *(const CUtensorMap *)c->tmap_Lq_half, args));
    return PSB_OK;
}
template <int NSL>
static int tc_launch(psb_ctx *c, const TcArgs &args, int grid, size_t smem) {
    return args.pair == 2 ? tc_launch2<NSL, true>(c, args, grid, smem)
                          : tc_launch2<NSL, false>(c, args, grid, smem);
}

// Tensor map over Lq: inner dim = samples (bytes), outer = (jtile, slice, comp) rows.
static int tc_make_tensor_map(psb_ctx *c, int Jall, int nsl) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PSB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    PSB_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, PSB_ERR_CUDA,
                "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)c->Kpad, (cuuint64_t)Jall * nsl};
    cuuint64_t gstr[1] = {(cuuint64_t)c->Kpad};
    cuuint32_t estr[2] = {1, 1};
    // one box = 128 samples (one swizzle row) x the sliced rows of a component tile ([0]: all of
    // them; [1]: half of them, what one CTA of a two-SM pair holds); a pipeline stage is two boxes
    CUtensorMap *tm[2] = {new CUtensorMap, new CUtensorMap};
    for (int h = 0; h < 2; ++h) {
        cuuint32_t box[2] = {TC_KBOX, (cuuint32_t)(TC_JT * nsl) >> h};
        CUresult cr = ((PFN_encodeTiled)fn)(tm[h], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, c->d_Lq, gdim, gstr,
                                           box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            delete tm[0];
            delete tm[1];
            psb_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
            return PSB_ERR_CUDA;
        }
    }
    c->tmap_Lq = tm[0];
    c->tmap_Lq_half = tm[1];
    return PSB_OK;
}

// Special tile (host side, exact): each source column u (v = M y, then the orthonormal
// covariate basis Q_e) becomes a pair of sliced columns hi = quantised u and lo = quantised
// (u - hi), so that g_hi s_hi + g_lo s_lo reproduces x'u to 2 (8k-2) bits.
static void tc_special_column(const double *u, int N, int nsl, int Kpad, int8_t *tile, int comp,
                              double *scale) {
    const int B = 8 * nsl - 2;
    std::vector<double> rem(u, u + N);
    for (int part = 0; part < 2; ++part) {
        double mx = 0.0;
        for (int i = 0; i < N; ++i) mx = std::max(mx, fabs(rem[i]));
        int e = 0;
        if (mx > 0.0) frexp(mx, &e);
        scale[comp + part] = mx > 0.0 ? ldexp(1.0, e - B) : 0.0;
        for (int i = 0; i < N; ++i) {
            long long qv = mx > 0.0 ? llrint(ldexp(rem[i], B - e)) : 0;
            rem[i] -= ldexp((double)qv, e - B);          // exact
            for (int s = 0; s < nsl; ++s) {
                long long d = ((qv + 128) & 255) - 128;
                qv = (qv - d) >> 8;
                tile[((size_t)s * TC_JT + comp + part) * Kpad + tc_kpos(i)] = (int8_t)d;
            }
        }
    }
}

int psb_lmm_tc_setup(psb_ctx *c, const double *h_v, const double *h_Q, int r, int ldq,
                     const double *h_w1, const double *h_w2) {
    const int nsl = c->precision;
    PSB_REQUIRE(nsl >= 3 && nsl <= 7, PSB_ERR_ARG, "int8 slice count must be in 3..7, got %d", nsl);
    const int N = c->N;
    c->n_slices = nsl;
    // triangular form (default): the operand is M'' (N x N, from M = L L'), half the work of
    // the rectangular L form, which stays available (PSB_TC_TRI=0) as a cross-check
    c->tc_tri = !(getenv("PSB_TC_TRI") && atoi(getenv("PSB_TC_TRI")) == 0);
    const int J = c->tc_tri ? N : c->J;
    const double *src = c->d_L;
    int src_ld = c->Jpad;
    double *d_M = nullptr;
    if (c->tc_tri) {
        const int ldm = ((N + 63) / 64) * 64;
        PSB_CUDA(cudaMalloc(&d_M, (size_t)ldm * ldm * sizeof(double)));
        dim3 g(ldm / 64, ldm / 64);
        k_tc_syrk<<<g, 256, 0, c->stream>>>(c->d_L, c->Lrows, c->Jpad, d_M, ldm);
        c->launches++;
        PSB_CUDA(cudaGetLastError());
        src = d_M;
        src_ld = ldm;
    }
    const int jt_reg = (J + TC_JT - 1) / TC_JT;
    // the special tile holds 1 + r hi/lo pairs; with more than 15 covariate columns the
    // masked column sums stay on the CUDA-core path (psb_varstats.cu)
    c->tc_special = (1 + r) * 2 <= TC_JT ? 1 + r : 0;
    // ... and, when two more pairs fit, the Welch columns (yc, yc^2) of pre_filtering (model.py:53-55)
    c->tc_welch = c->tc_special > 0 && h_w1 && h_w2 && (1 + r + 2) * 2 <= TC_JT;
    c->jtiles = jt_reg + (c->tc_special ? 1 : 0);
    const int Jq = jt_reg * TC_JT;
    const int Jall = c->jtiles * TC_JT;
    c->Kpad = ((N + TC_KSTAGE - 1) / TC_KSTAGE) * TC_KSTAGE;
    int *d_expo = nullptr;
    PSB_CUDA(cudaMalloc(&d_expo, Jq * sizeof(int)));
    PSB_CUDA(cudaMalloc(&c->d_scale2, Jall * sizeof(double)));
    size_t lq_bytes = (size_t)Jall * nsl * c->Kpad;
    PSB_CUDA(cudaMalloc(&c->d_Lq, lq_bytes));
    // integer epilogue: 2^B (digits) * N (carriers) * 2^shmax * 32 (columns) must stay below 2^62
    int shmax = -1;
    c->tc_int_epi = false;
    if (c->tc_tri && N <= 16384 && !(getenv("PSB_TC_INT_EPI") && atoi(getenv("PSB_TC_INT_EPI")) == 0)) {
        int lg = 0;
        while ((1 << lg) < N) ++lg;
        shmax = std::min(5, 62 - (8 * nsl - 2) - 5 - lg);
        c->tc_int_epi = shmax >= 0;
    }
    PSB_CUDA(cudaMalloc(&c->d_shift, Jq));
    PSB_CUDA(cudaMemsetAsync(c->d_shift, 0, Jq, c->stream));
    k_tc_colexp<<<psb_div_up(Jq, 128), 128, 0, c->stream>>>(src, N, src_ld, J, nsl, d_expo,
                                                           c->d_scale2, Jq, c->tc_tri ? 1 : 0,
                                                           c->tc_int_epi ? shmax : -1, c->d_shift);
    c->launches++;
    PSB_CUDA(cudaGetLastError());
    size_t total = (size_t)Jq * c->Kpad;
    k_tc_quantise<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(src, N, src_ld, J, nsl,
                                                                         d_expo, c->d_Lq, c->Kpad, Jq,
                                                                         c->tc_tri ? 1 : 0,
                                                                         (getenv("PSB_TC_DITHER") && atoi(getenv("PSB_TC_DITHER")) == 0) ? 0 : 1);
    c->launches++;
    PSB_CUDA(cudaGetLastError());
    if (c->tc_special) {
        size_t tile_bytes = (size_t)nsl * TC_JT * c->Kpad;
        std::vector<int8_t> tile(tile_bytes, 0);
        std::vector<double> sc(TC_JT, 0.0);
        tc_special_column(h_v, N, nsl, c->Kpad, tile.data(), 0, sc.data());
        for (int e = 0; e < r; ++e)
            tc_special_column(h_Q + (size_t)e * ldq, N, nsl, c->Kpad, tile.data(), 2 + 2 * e, sc.data());
        if (c->tc_welch) {
            tc_special_column(h_w1, N, nsl, c->Kpad, tile.data(), 2 * (1 + r), sc.data());
            tc_special_column(h_w2, N, nsl, c->Kpad, tile.data(), 2 * (2 + r), sc.data());
        }
        PSB_CUDA(cudaMemcpyAsync(c->d_Lq + (size_t)jt_reg * tile_bytes, tile.data(), tile_bytes,
                                 cudaMemcpyHostToDevice, c->stream));
        PSB_CUDA(cudaMemcpyAsync(c->d_scale2 + Jq, sc.data(), TC_JT * sizeof(double),
                                 cudaMemcpyHostToDevice, c->stream));
        PSB_CUDA(cudaStreamSynchronize(c->stream));
    }
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_expo);
    if (d_M) cudaFree(d_M);

    return tc_make_tensor_map(c, Jall, nsl);
}

// Linear-only operand: `ncols` columns (each N doubles, column e at cols + e * ld) as hi/lo
// sliced pairs in one component tile.  Used by the fixed-effects OLS path for y'x and Z'x
// (model.py:300-312 needs nothing else from the variant): one tensor pass replaces ncols
// masked fp64 column sums per variant.
int psb_tc_linear_setup(psb_ctx *c, const double *cols, int ncols, int ld) {
    const int nsl = 5;
    PSB_REQUIRE(ncols >= 1 && ncols * 2 <= TC_JT, PSB_ERR_UNSUPPORTED, "too many linear columns");
    psb_lmm_tc_free(c);
    const int N = c->N;
    c->n_slices = nsl;
    c->tc_special = ncols;
    c->tc_welch = false;
    c->tc_tri = false;
    c->tc_int_epi = false;
    c->jtiles = 1;
    c->Kpad = ((N + TC_KSTAGE - 1) / TC_KSTAGE) * TC_KSTAGE;
    const size_t tile_bytes = (size_t)nsl * TC_JT * c->Kpad;
    std::vector<int8_t> tile(tile_bytes, 0);
    std::vector<double> sc(TC_JT, 0.0);
    for (int e = 0; e < ncols; ++e)
        tc_special_column(cols + (size_t)e * ld, N, nsl, c->Kpad, tile.data(), 2 * e, sc.data());
    PSB_CUDA(cudaMalloc(&c->d_Lq, tile_bytes));
    PSB_CUDA(cudaMalloc(&c->d_scale2, TC_JT * sizeof(double)));
    PSB_CUDA(cudaMemcpy(c->d_Lq, tile.data(), tile_bytes, cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(c->d_scale2, sc.data(), TC_JT * sizeof(double), cudaMemcpyHostToDevice));
    PSB_UPLOAD_FENCE();
    return tc_make_tensor_map(c, TC_JT, nsl);
}

int psb_lmm_tc_run(psb_ctx *c, int n_tested) { return psb_tc_run(c, n_tested, nullptr, 0); }

// the same over another list of variants (ids in `list`, their number in *count_dev on the device)
static const int32_t *g_tc_list = nullptr;
static const int *g_tc_count = nullptr;
int psb_lmm_tc_run_list(psb_ctx *c, int upper, const int32_t *list, const int *count_dev) {
    g_tc_list = list;
    g_tc_count = count_dev;
    const int rc = psb_tc_run(c, upper, nullptr, 0);
    g_tc_list = nullptr;
    g_tc_count = nullptr;
    return rc;
}

// lin_out != null: linear-only use (psb_tc_linear_setup): the pair sums go to
// lin_out[variant * lin_ld + pair] and no quadratic form is produced.
// Second half of a split launch: a[row] = sum_g a_part[g][t] in a fixed order (the kernel's G is
// recomputed from the same device-side count; nothing to do when it was 1).
__global__ void k_tc_split_reduce(const int32_t *__restrict__ idx, const int *__restrict__ count_dev, int upper,
                                  int split_max, int n_units, int pair, int jtiles,
                                  const double *__restrict__ part, int part_ld, double *__restrict__ a) {
    const int n = count_dev ? min(*count_dev, upper) : upper;
    const int n_tiles = (n + TC_TILE_V - 1) / TC_TILE_V;
    const int G = tc_split_of(split_max, pair ? (n_tiles + 1) / 2 : n_tiles, n_units, jtiles);
    if (G == 1) return;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int g = 0; g < G; ++g) s += part[(size_t)g * part_ld + t];
        a[idx[t]] = s;
    }
}

int psb_tc_run(psb_ctx *c, int n_tested, double *lin_out, int lin_ld) {
    const int nsl = c->n_slices;
    TcArgs a;
    a.bits = c->d_bits;
    a.idx = g_tc_list ? g_tc_list : c->d_idx;
    a.scale2 = c->d_scale2;
    a.a_out = lin_out ? nullptr : c->d_a;
    a.b_out = lin_out ? nullptr : c->d_b;
    a.pp_out = lin_out ? nullptr : c->d_pp;
    a.lin_out = lin_out;
    a.lin_ld = lin_ld;
    const bool welch = !lin_out && c->tc_welch && c->tc_welch_run;
    a.w_out = welch ? c->d_sums + c->col_w0 : nullptr;
    a.w_ld = c->C;
    a.n_special = c->tc_special;
    a.Wrow = c->Wrow;
    a.n_tested = n_tested;
    a.n_tested_dev = g_tc_count ? g_tc_count : c->d_counters;      // counters[0] = tested variants
    a.nks = c->Kpad / TC_KSTAGE;
    a.jtiles = c->jtiles;
    a.tri = c->tc_tri ? 1 : 0;
    a.int_epi = c->tc_int_epi ? 1 : 0;
    a.shift = c->d_shift;
    {
        const int nreg = c->jtiles - (c->tc_special > 0 ? 1 : 0);
        int stages = (c->tc_special > 0 ? a.nks : 0);
        for (int jt = 0; jt < nreg; ++jt) stages += a.nks - (a.tri ? (jt * TC_JT) / TC_KSTAGE : 0);
        a.stages_per_tile = stages;
    }
    int pitch = a.nks * (TC_KSTAGE / 32);
    while (pitch % 32 != 4) pitch += 4;
    a.pitch = pitch;
    int smem_max = 0;
    PSB_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    // CTA-pair mode (PSB_TC_PAIR): 2 (default) = two-SM MMAs, each CTA holds half of every B
    // stage; 1 = independent MMAs, operand stream shared by TMA multicast; 0 = single CTAs
    a.pair = c->sm_count >= 2 ? 2 : 0;
    if (getenv("PSB_TC_PAIR")) a.pair = std::max(0, std::min(2, atoi(getenv("PSB_TC_PAIR"))));
    if (c->sm_count < 2) a.pair = 0;
    a.debug = getenv("PSB_TC_DEBUG") ? atoi(getenv("PSB_TC_DEBUG")) : 0;
    a.skip_box = (a.tri && !(getenv("PSB_TC_SKIP") && atoi(getenv("PSB_TC_SKIP")) == 0)) ? 1 : 0;
    const int na = (512 - 2 * TC_JT * nsl) / (TC_KSTAGE / 4);   // TMEM A-ring depth (kernel: NA)
    // Operand ring depth: as many stages as shared memory holds, at most TC_MAX_BSTAGES.  The
    // tile's packed rows share shared memory with the ring; they stay there as long as the ring
    // keeps at least the depth of the TMEM A ring (measured at N=5000: rows in shared memory
    // beat global rows with a deeper ring), otherwise the expanders read their rows from
    // global memory (PSB_TC_BITS_SMEM=0/1 forces either mode, PSB_TC_STAGES caps the depth).
    auto depth = [&](int pit) {
        int nb = TC_MAX_BSTAGES;
        while (nb > 2 && tc_smem_bytes(nsl, nb, pit, a.pair) > (size_t)smem_max) --nb;
        return nb;
    };
    const int nb_smem = depth(pitch), nb_glob = depth(4);
    a.bits_in_smem = (tc_smem_bytes(nsl, nb_smem, pitch, a.pair) <= (size_t)smem_max && nb_smem >= na) ? 1 : 0;
    if (getenv("PSB_TC_GLOBAL_BITS")) a.bits_in_smem = 0;      // test hook: force the large-N mode
    if (getenv("PSB_TC_BITS_SMEM") && tc_smem_bytes(nsl, nb_smem, pitch, a.pair) <= (size_t)smem_max)
        a.bits_in_smem = atoi(getenv("PSB_TC_BITS_SMEM")) ? 1 : 0;
    if (!a.bits_in_smem) {
        pitch = 4;                               // no bit tile in shared memory
        a.pitch = pitch;
    }
    int nb = a.bits_in_smem ? nb_smem : nb_glob;
    if (getenv("PSB_TC_STAGES")) nb = std::max(2, std::min(nb, atoi(getenv("PSB_TC_STAGES"))));
    PSB_REQUIRE(tc_smem_bytes(nsl, nb, pitch, a.pair) <= (size_t)smem_max, PSB_ERR_UNSUPPORTED,
                "n_samples = %d needs more shared memory than the tensor path has; use precision 0",
                c->N);
    a.n_bstages = nb;
    const size_t smem = tc_smem_bytes(nsl, nb, pitch, a.pair);
    const int tiles = psb_div_up(n_tested, TC_TILE_V);
    int grid = std::min(tiles, c->sm_count);
    if (a.pair) grid = std::min((tiles + 1) & ~1, c->sm_count & ~1);
    // Few variant tiles (the device-side count decides: the refinement pass, small batches): their
    // component tiles are dealt to several units (see the kernel), which needs the whole grid and a
    // zeroed accumulator.
    a.split_max = (a.a_out && !(getenv("PSB_TC_SPLIT") && atoi(getenv("PSB_TC_SPLIT")) == 0)) ? 16 : 1;
    a.a_part = nullptr;
    a.part_ld = 0;
    if (a.split_max > 1 && n_tested > 0) {
        grid = a.pair ? (c->sm_count & ~1) : c->sm_count;
        a.part_ld = (c->sm_count / 2 + 1) * TC_TILE_V;         // a split launch has at most this many variants
        if (!c->d_tc_part) PSB_CUDA(cudaMalloc(&c->d_tc_part, (size_t)16 * a.part_ld * sizeof(double)));
        a.a_part = c->d_tc_part;
    }
    int rc = PSB_OK;
    switch (nsl) {
        case 3: rc = tc_launch<3>(c, a, grid, smem); break;
        case 4: rc = tc_launch<4>(c, a, grid, smem); break;
        case 5: rc = tc_launch<5>(c, a, grid, smem); break;
        case 6: rc = tc_launch<6>(c, a, grid, smem); break;
        case 7: rc = tc_launch<7>(c, a, grid, smem); break;
        default:
            psb_set_error("unsupported slice count %d", nsl);
            return PSB_ERR_ARG;
    }
    if (rc) return rc;
    c->launches++;
    PSB_CUDA(cudaGetLastError());
    if (a.split_max > 1 && n_tested > 0) {
        k_tc_split_reduce<<<37, 256, 0, c->stream>>>(a.idx, a.n_tested_dev, n_tested, a.split_max, a.pair ? grid / 2 : grid,
                                                    a.pair ? 1 : 0, a.jtiles, a.a_part, a.part_ld, c->d_a);
        PSB_CUDA(cudaGetLastError());
    }
    return PSB_OK;
}

// ---- two operand images in one context: the working precision and the refinement precision of
// the two-pass mode (psb_lmm_setup, precision 46) ---------------------------------------------------
struct TcState {
    int8_t *d_Lq;
    double *d_scale2;
    uint8_t *d_shift;
    void *tmap_Lq, *tmap_Lq_half;
    int n_slices, jtiles, Kpad, tc_special;
    bool tc_tri, tc_int_epi, tc_welch;
};
static void tc_save(const psb_ctx *c, TcState &s) {
    s.d_Lq = c->d_Lq; s.d_scale2 = c->d_scale2; s.d_shift = c->d_shift;
    s.tmap_Lq = c->tmap_Lq; s.tmap_Lq_half = c->tmap_Lq_half;
    s.n_slices = c->n_slices; s.jtiles = c->jtiles; s.Kpad = c->Kpad; s.tc_special = c->tc_special;
    s.tc_tri = c->tc_tri; s.tc_int_epi = c->tc_int_epi; s.tc_welch = c->tc_welch;
}
static void tc_load(psb_ctx *c, const TcState &s) {
    c->d_Lq = s.d_Lq; c->d_scale2 = s.d_scale2; c->d_shift = s.d_shift;
    c->tmap_Lq = s.tmap_Lq; c->tmap_Lq_half = s.tmap_Lq_half;
    c->n_slices = s.n_slices; c->jtiles = s.jtiles; c->Kpad = s.Kpad; c->tc_special = s.tc_special;
    c->tc_tri = s.tc_tri; c->tc_int_epi = s.tc_int_epi; c->tc_welch = s.tc_welch;
}
// swaps the context's current operand image with the alternate one
void psb_lmm_tc_swap(psb_ctx *c) {
    if (!c->tc_alt) return;
    TcState cur;
    tc_save(c, cur);
    tc_load(c, *(TcState *)c->tc_alt);
    *(TcState *)c->tc_alt = cur;
}
// the image just built by psb_lmm_tc_setup becomes the alternate one; the context is left without
// a current image (the caller builds the working one next)
int psb_lmm_tc_stash(psb_ctx *c) {
    TcState *s = new TcState;
    tc_save(c, *s);
    c->tc_alt = s;
    c->d_Lq = nullptr;
    c->d_scale2 = nullptr;
    c->d_shift = nullptr;
    c->tmap_Lq = c->tmap_Lq_half = nullptr;
    return PSB_OK;
}

void psb_lmm_tc_free(psb_ctx *c) {
    if (c->d_tc_part) cudaFree(c->d_tc_part);
    c->d_tc_part = nullptr;
    if (c->tc_alt) {
        TcState *s = (TcState *)c->tc_alt;
        if (s->d_Lq) cudaFree(s->d_Lq);
        if (s->d_scale2) cudaFree(s->d_scale2);
        if (s->d_shift) cudaFree(s->d_shift);
        if (s->tmap_Lq) delete (CUtensorMap *)s->tmap_Lq;
        if (s->tmap_Lq_half) delete (CUtensorMap *)s->tmap_Lq_half;
        delete s;
        c->tc_alt = nullptr;
    }
    if (c->d_Lq) cudaFree(c->d_Lq);
    if (c->d_scale2) cudaFree(c->d_scale2);
    if (c->d_shift) cudaFree(c->d_shift);
    c->d_Lq = nullptr;
    c->d_scale2 = nullptr;
    c->d_shift = nullptr;
    if (c->tmap_Lq) delete (CUtensorMap *)c->tmap_Lq;
    if (c->tmap_Lq_half) delete (CUtensorMap *)c->tmap_Lq_half;
    c->tmap_Lq_half = nullptr;
    c->tmap_Lq = nullptr;

}
