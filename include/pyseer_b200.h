/*
 * pyseer_b200.h -- C ABI of libpyseer_b200.so
 *
 * Drop-in boundary for pyseer's per-variant association loop.  pyseer itself has no
 * FFI; the seam is the two Python call signatures its worker map invokes
 * (pyseer/__main__.py:541-544 -> lmm.fit_lmm, :777-780 -> model.fixed_effects_regression)
 * plus the namedtuples they return (pyseer/classes.py:3-22).  Each entry point below
 * names the reference interface it replaces.  Plain C: pointers and sizes only, no
 * exceptions across the boundary, every function returns an int status
 * (PSB_OK or a negative psb_status; text via psb_last_error()).
 *
 * Ownership: every pointer passed in is caller-owned and only read (or written, for
 * outputs) during the call, except psb_submit_device() whose device buffer must stay
 * valid until the next psb_submit*() / psb_destroy().  The library owns all device
 * memory it allocates.  One psb_ctx drives one GPU on one stream; contexts are
 * independent (one per process under torchrun, or several in one process).
 */
#ifndef PYSEER_B200_H
#define PYSEER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSB_ABI_VERSION 2

typedef enum psb_status {
    PSB_OK = 0,
    PSB_ERR_CUDA = -1,      /* CUDA runtime / driver error                       */
    PSB_ERR_ARG = -2,       /* bad argument (shape mismatch, null pointer ...)   */
    PSB_ERR_STATE = -3,     /* call out of order (run before setup/submit ...)   */
    PSB_ERR_H2 = -4,        /* h2 outside [0,1): lmm_cov.py:667-670 (KeyError)   */
    PSB_ERR_NOMEM = -5,
    PSB_ERR_UNSUPPORTED = -6,
    PSB_ERR_NUMERIC = -7    /* an iterative library routine did not converge      */
} psb_status;

/* Per-variant flag word.  Bits map 1:1 onto the reference's `notes` strings and the
 * `prefilter` / `filter` booleans of classes.Seer / classes.LMM. */
#define PSB_F_AF_FILTER        0x0001u  /* 'af-filter'            model.py:255 lmm.py:163 */
#define PSB_F_PREFILTER_FAILED 0x0002u  /* 'pre-filtering-failed' model.py:266 lmm.py:174 */
#define PSB_F_BAD_CHISQ        0x0004u  /* 'bad-chisq'            model.py:264 lmm.py:172 */
#define PSB_F_HIGH_BSE         0x0008u  /* 'high-bse'             model.py:332            */
#define PSB_F_PERFECT_SEP      0x0010u  /* 'perfectly-separable-data' model.py:345        */
#define PSB_F_MATRIX_INV       0x0020u  /* 'matrix-inversion-error'   model.py:349        */
#define PSB_F_FIRTH_FAIL       0x0040u  /* 'firth-fail'           model.py:357            */
#define PSB_F_MISSING_DATA     0x0080u  /* 'missing-data-error'   model.py:371            */
#define PSB_F_LRT_FAILED       0x0100u  /* 'lrt-filtering-failed' model.py:384 lmm.py:201 */
#define PSB_F_PREFILTER        0x0200u  /* .prefilter == True                             */
#define PSB_F_FILTER           0x0400u  /* .filter == True                                */
#define PSB_F_TESTED           0x0800u  /* counted in `tested` (__main__.py:558/793)      */
#define PSB_F_FIRTH_USED       0x1000u  /* informational: result came from fit_firth      */

typedef struct psb_ctx psb_ctx;

/* Thresholds of one run: pyseer options --min-af --max-af --max-missing
 * --filter-pvalue --lrt-pvalue (__main__.py:180-206) and the `continuous` switch. */
typedef struct psb_params {
    double min_af;
    double max_af;
    double max_missing;
    double filter_pvalue;
    double lrt_pvalue;
    int32_t continuous;     /* phenotype treated as continuous by pre_filtering()   */
    int32_t options;        /* PSB_OPT_* bits                                        */
} psb_params;

/* fit every submitted variant, no AF filter and no pre_filtering(): the semantics of a
 * direct lmm.fit_lmm_block call (lmm.py:228-260), which filters nothing */
#define PSB_OPT_NO_PREFILTER 0x1

/* Result table, structure-of-arrays, one entry per submitted variant in submission
 * order.  Any pointer may be NULL (that column is skipped).  Doubles follow the
 * reference: NaN where the reference leaves the namedtuple field at NaN. */
typedef struct psb_results {
    int32_t  *carriers;   /* len(kstrains): samples with the variant, missing included (input.py:439) */
    int32_t  *missing;    /* samples with NaN genotype                                               */
    double   *af;         /* carriers / N (input.py:446)                                             */
    double   *prep;       /* filter-pvalue (model.pre_filtering)                                     */
    double   *pvalue;     /* lrt-pvalue                                                              */
    double   *beta;       /* kbeta                                                                   */
    double   *bse;        /* beta-std-err                                                            */
    double   *extra;      /* LMM: frac_h2 (variant_h2); fixed effects: intercept                     */
    double   *betas;      /* fixed effects only: n_variants x (q-1) row-major covariate slopes       */
    uint32_t *flags;      /* PSB_F_*                                                                 */
} psb_results;

/* ---- library / context ------------------------------------------------------ */
int psb_abi_version(void);
const char *psb_last_error(void);
int psb_device_count(int *count);
/* replaces multiprocessing.Pool(options.cpu) (__main__.py:517-519): one context per GPU */
int psb_create(int device_id, psb_ctx **out);
int psb_destroy(psb_ctx *ctx);
int psb_sync(psb_ctx *ctx);

/* ---- once-per-run model state ------------------------------------------------ */
/* LMM state after lmm.initialise_lmm (lmm.py:26-122): covariates X (N x D row-major,
 * last column ones, lmm.py:95-99), phenotype y, kernel eigenvectors U (N x (N-D)
 * row-major) and eigenvalues S (lmm_cov.py:88-103) and h2 (lmm_cov.findH2).  Builds
 * the rotated operands used by fit_lmm_block (lmm.py:228-260).
 * precision: 0 = FP64 CUDA-core contraction; k in [3,7] = exact k-slice int8
 * tensor-core contraction (tcgen05); 46 = two passes, 4 slices for every variant and 6 slices again
 * for those whose F statistic exceeds 30 (the far tail, where the relative error of a p-value is
 * F / 2 times that of the quadratic form).  Returns PSB_ERR_H2 when h2 is outside [0,1). */
int psb_lmm_setup(psb_ctx *ctx, int32_t n_samples, int32_t n_cov, const double *X,
                  const double *y, const double *U, const double *S, double h2,
                  int32_t precision);

/* Symmetric eigendecomposition on the device, replacing the host `eigh` of LMM.setSU_fromK
 * (fastlmm/lmm_cov.py:88-103) inside lmm.initialise_lmm (lmm.py:26-122) -- the O(N^3),
 * once-per-run part.  A: n x n symmetric (host); w_out: n eigenvalues in ascending order;
 * V_out: n x n row-major with the eigenvector of w_out[j] in column j (numpy.linalg.eigh layout).
 * Runs cuSOLVER's fp64 syevd (loaded lazily with dlopen; PSB_ERR_UNSUPPORTED when it is not
 * installed, in which case the host side keeps its NumPy eigh).  Needs no model set-up. */
int psb_eigh(psb_ctx *ctx, int32_t n, const double *A, double *w_out, double *V_out);
/* All of LMM.setSU_fromK (fastlmm/lmm_cov.py:88-103) on the device: K (n x n row-major, host; the
 * normalised kinship, before the + I), covariate design X (n x d row-major) and its pseudo-inverse
 * Xp (d x n, Linreg's Xdagger, lmm_cov.py:861-880) -> K_ = regress(regress(K + I)') formed as the
 * reference forms it (two rank-d updates), then syevd.  w_out / V_out as psb_eigh; the caller keeps the
 * pairs d..n-1 and subtracts 1 from the eigenvalues.  1 <= d <= 16. */
int psb_spectral(psb_ctx *ctx, int32_t n, int32_t d, const double *K, const double *X, const double *Xp,
                 double *w_out, double *V_out);
/* The O(N) sums of the null-model likelihood LMM.nLLeval(h2) (fastlmm/lmm_cov.py:597-684) that
 * LMM.findH2's grid + Brent search (lmm_cov.py:427-478, mingrid.py:13-73) evaluates ~30 times:
 * for each of the n_h2 values, yky_out = sum_j uy2[j] / (h2 S[j] + 1 - h2) and logdet_out =
 * sum_j log(h2 S[j] + 1 - h2) over the J = N - D spectrum (S: eigenvalues, uy2: squares of the rotated
 * phenotype U'Py; both host arrays).  Needs no model set-up. */
int psb_lmm_nll_terms(psb_ctx *ctx, int32_t J, const double *S, const double *uy2, int32_t n_h2,
                      const double *h2, double *yky_out, double *logdet_out);

/* Fixed-effects state shared by every fixed_effects_regression call (model.py:202-205):
 * Z = [1, m, c] (N x q row-major, column 0 ones; model.py:274-297 minus the variant
 * column), phenotype y, `continuous`, null log-likelihoods null_res / null_firth
 * (model.fit_null; __main__.py:378-380, 449-450). */
int psb_fixed_setup(psb_ctx *ctx, int32_t n_samples, int32_t q, const double *Z,
                    const double *y, int32_t continuous, double null_llf,
                    double null_firth_llf);

/* Null model fits on the device with the same solver (model.fit_null, model.py:73-148).
 * Design Z (N x q row-major), no variant column.  out_params: q doubles; out_bse: q
 * doubles; out_llf: log-likelihood; firth bit 0 -> Firth-penalised fit (out_llf only), bit 1 ->
 * start from zeros instead of the log-odds intercept (statsmodels' default, used by
 * model.fit_lineage_effect, model.py:181-190).  The context's model slot is reused: call
 * before psb_fixed_setup / psb_lmm_setup.
 * Returns PSB_OK and *out_status = PSB_F_* bits (0 = fitted). */
int psb_fit_null(psb_ctx *ctx, int32_t n_samples, int32_t q, const double *Z,
                 const double *y, int32_t continuous, int32_t firth, double *out_params,
                 double *out_bse, double *out_llf, uint32_t *out_status);

/* ---- variants ---------------------------------------------------------------- */
/* Packed presence/absence rows replacing the tuples of input.iter_variants /
 * load_var_block (input.py:505-707): n_variants rows of words_per_row uint32, bit
 * (i % 32) of word (i / 32) = sample i of the phenotype index order; words_per_row
 * >= ceil(N/32) and a multiple of 4 (16-byte rows); padding bits must be zero.
 * `missing` (same shape, nullable) marks NaN genotypes (Rtab '.', VCF no-call).
 * psb_submit copies host->device asynchronously on the context's COPY stream into one of two
 * staging slots (pass pinned memory): the copy of batch i+1 overlaps the kernels of batch i,
 * and the table of the previous run stays fetchable until the next psb_run_* adopts the new
 * rows.  Pipelined use:  submit(0) run(0) | submit(1) fetch(0) run(1) | submit(2) fetch(1) ...
 * psb_submit_device adopts device-resident rows without a copy. */
int psb_submit(psb_ctx *ctx, const uint32_t *bits, const uint32_t *missing,
               int64_t n_variants, int32_t words_per_row);
int psb_submit_device(psb_ctx *ctx, const void *d_bits, const void *d_missing,
                      int64_t n_variants, int32_t words_per_row);

/* Burden regions (`--vcf --burden`): replaces the region branch of input.read_variant
 * (input.py:395-411; load_burden :250-266).  `bits` / `missing` hold one packed row per VCF
 * RECORD (read_vcf_var applied to an empty dictionary, input.py:457-502); region r is the
 * union of the records members[region_offsets[r] .. region_offsets[r+1]) in fetch order:
 * carrier if any member carries the alternative allele, missing if the LAST member is missing
 * and no member carries (the dictionary semantics of input.py:489-497).  The reduction runs on
 * the device and leaves n_regions rows submitted, as psb_submit would; region_offsets has
 * n_regions + 1 entries starting at 0.  psb_submit_burden_device takes device-resident record
 * rows (they must stay valid until the next psb_run_* has been queued); psb_submitted_device
 * returns the device pointers of the rows currently submitted (e.g. by psb_synth_device). */
int psb_submit_burden(psb_ctx *ctx, const uint32_t *bits, const uint32_t *missing,
                      int64_t n_variants, int32_t words_per_row, const int64_t *region_offsets,
                      const int32_t *members, int64_t n_regions);
int psb_submit_burden_device(psb_ctx *ctx, const void *d_bits, const void *d_missing,
                             int64_t n_variants, int32_t words_per_row,
                             const int64_t *region_offsets, const int32_t *members,
                             int64_t n_regions);
int psb_submitted_device(psb_ctx *ctx, const void **d_bits, const void **d_missing,
                         int64_t *n_variants, int32_t *words_per_row);

/* ---- the hot path ------------------------------------------------------------ */
/* replaces lmm.fit_lmm (lmm.py:125-226) for every submitted variant */
int psb_run_lmm(psb_ctx *ctx, const psb_params *params);
/* replaces model.fixed_effects_regression (model.py:202-394) for every submitted variant */
int psb_run_fixed(psb_ctx *ctx, const psb_params *params);

/* Lineage effects, model.fit_lineage_effect (model.py:151-199; called from model.py:379-380 and
 * lmm.py:209-211): Logit of the VARIANT on Zlin = [1, lineage columns, covariates] (N x q
 * row-major, n_lineage lineage columns after the intercept).  psb_lineage_setup goes after the
 * model set-up; psb_run_lineage after psb_run_* on the same submitted rows (mode 0: every
 * tested variant that produced a fit, the fixed-effects rule; mode 1: tested variants that
 * passed the lrt filter, the LMM rule); psb_fetch_lineage returns, per submitted variant, the
 * index of the lineage with the largest Wald statistic or -1 (None). */
int psb_lineage_setup(psb_ctx *ctx, int32_t n_samples, int32_t q, const double *Zlin,
                      int32_t n_lineage);
int psb_run_lineage(psb_ctx *ctx, int32_t mode);
int psb_fetch_lineage(psb_ctx *ctx, int32_t *out);

/* Copies the result table to caller arrays (synchronises the context stream).  The
 * destination pointers may be host memory (pinned or pageable) or device memory (e.g. a
 * buffer that is then gathered across ranks with NCCL). */
int psb_fetch(psb_ctx *ctx, const psb_results *out);
/* Asynchronous form: psb_fetch_begin queues the copies of the last run's table on the library's fetch
 * stream and returns; the next psb_run_* (which writes a second set of result columns) may be queued
 * at once, so the device does not idle while a table travels to the host -- the result loop of
 * __main__.py:547-568 overlapped with the next block's fits.  psb_fetch_wait blocks until the copies of
 * the last psb_fetch_begin have landed; counts_out (nullable) = that run's psb_counts.  `out` and the
 * buffers it names must stay valid until then. */
int psb_fetch_begin(psb_ctx *ctx, const psb_results *out);
int psb_fetch_wait(psb_ctx *ctx, int64_t counts_out[4]);
/* Device pointers of the result table of the last run (valid until the next run);
 * lets the caller gather tables across ranks with NCCL without a host round trip. */
int psb_results_device(psb_ctx *ctx, psb_results *out_device_ptrs);
/* counters of the last run: [0] loaded, [1] pre-filtered, [2] tested, [3] passed filter
 * (__main__.py:831-834 semantics, printed == tested - filtered unless --print-filtered) */
int psb_counts(psb_ctx *ctx, int64_t out[4]);

/* ---- multi-GPU: the gather of the result table ---------------------------------------- */
/* Replaces multiprocessing.Pool(options.cpu) and the ordered pool.starmap of the worker map
 * (__main__.py:517-519, :541-546, :777-780): variants are independent given the once-per-run state,
 * ranks take contiguous variant ranges, and the one exchange of the path is the gather of the
 * per-variant result table on a root rank, in rank order (= input order).  NCCL is loaded with dlopen
 * (libnccl.so.2, or the file named by $PSB_NCCL_LIB); PSB_ERR_UNSUPPORTED when it is absent.
 * A psb_comm holds the LOCAL members of the communicator: one context per process
 * (psb_comm_init_rank; rank 0 makes the id with psb_comm_unique_id and hands it to the other ranks
 * by any means -- the Python side uses a file) or all contexts of a process (psb_comm_init_all: rank
 * i = ctxs[i]).
 * psb_comm_gather_begin queues, behind the last psb_run_* of every local member: a device-side
 * pack of its table (header + columns, rows_max rows; the same rows_max on every rank, >= the
 * rows of any rank), then ncclSend to `root` / the root's ncclRecv from every rank, on a stream of
 * the communicator -- the next psb_run_* only waits for the pack, so the transfer overlaps its
 * kernels.  psb_comm_gather_wait joins that stream into the contexts' streams and the host.
 * psb_comm_gather_fetch (root's process) copies rank src_rank's table out of the receive buffer to
 * host or device pointers and reports its row count and psb_counts[4].
 * bcast / allreduce (op 0 = sum, 1 = max, fp64) / barrier work on HOST buffers and need
 * one-context communicators: what a launcher needs for the once-per-run state and the timing. */
typedef struct psb_comm psb_comm;
#define PSB_COMM_ID_BYTES 128
int psb_comm_unique_id(uint8_t id[PSB_COMM_ID_BYTES]);
int psb_comm_init_rank(psb_ctx *ctx, int32_t world, int32_t rank, const uint8_t id[PSB_COMM_ID_BYTES],
                       psb_comm **out);
int psb_comm_init_all(psb_ctx *const *ctxs, int32_t n, psb_comm **out);
int psb_comm_destroy(psb_comm *comm);
int psb_comm_info(psb_comm *comm, int32_t *world, int32_t *n_local, int32_t *nccl_version);
int psb_comm_bcast(psb_comm *comm, void *host_buf, size_t bytes, int32_t root);
int psb_comm_allreduce(psb_comm *comm, double *vals, int32_t n, int32_t op);
int psb_comm_barrier(psb_comm *comm);
int psb_comm_gather_begin(psb_comm *comm, int32_t root, int64_t rows_max);
int psb_comm_gather_wait(psb_comm *comm);
int psb_comm_gather_fetch(psb_comm *comm, int32_t src_rank, const psb_results *out, int64_t *n_rows,
                          int64_t counts[4]);
int psb_comm_gather_bytes(psb_comm *comm, int64_t *bytes_per_rank);

/* ---- pinned host staging -------------------------------------------------------- */
/* Page-locked host buffers for psb_submit / psb_fetch (input.py's k-mer streaming becomes
 * a pinned-host -> device staging pipeline).  psb_download_bits copies the currently
 * submitted device rows (e.g. made by psb_synth_device) into a host buffer of
 * n_variants * words_per_row uint32. */
int psb_host_alloc(size_t bytes, void **out);
int psb_host_free(void *ptr);
int psb_download_bits(psb_ctx *ctx, uint32_t *out_bits);
/* same, with the missing-genotype rows (skipped when the batch has none: *has_missing = 0) */
int psb_download_rows(psb_ctx *ctx, uint32_t *out_bits, uint32_t *out_missing,
                      int32_t *has_missing);

/* ---- sample similarity (kinship) matrix ----------------------------------------------- */
/* K = G G' of pyseer/similarity.py:99-116 over packed rows: K[i][j] = number of variants that
 * pass the AF / missing filter (input.py:693) and are carried by both samples.  Independent of
 * the association models: begin(N), add batches of host rows (same layout as psb_submit), fetch
 * the N x N row-major matrix (exact integers as doubles). */
int psb_kinship_begin(psb_ctx *ctx, int32_t n_samples);
int psb_kinship_add(psb_ctx *ctx, const uint32_t *bits, const uint32_t *missing, int64_t n_variants,
                    int32_t words_per_row, double min_af, double max_af, double max_missing);
/* the same for the rows of the batch last submitted to the context (psb_submit, psb_submit_text: k-mer
 * text tokenised on the device, psb_submit_device) -- the variant file then never exists as host rows */
int psb_kinship_add_submitted(psb_ctx *ctx, double min_af, double max_af, double max_missing);
int psb_kinship_fetch(psb_ctx *ctx, double *K_out);
/* The tool's output, DataFrame(K, index=samples, columns=samples).to_csv(sep='\t') of
 * pyseer/similarity.py:118-120, written natively: a header line of the sample names after an empty index
 * label, then a line per sample with the counts as pandas prints float64 ('1234.0').  names: NUL-terminated
 * sample names back to back, name_off their offsets.  PSB_ERR_UNSUPPORTED when an entry is not a
 * non-negative integer below 1e15 (use pandas then); PSB_ERR_NOMEM when out_cap is too small. */
int psb_format_matrix(const double *K, int32_t n, const char *names, const int64_t *name_off, int32_t n_threads,
                      char *out, int64_t out_cap, int64_t *out_len);

/* ---- native variant-file reader ------------------------------------------------- */
/* Replaces the per-line Python of input.read_variant (input.py:301-454) for the k-mer text
 * format (`name | s1:1 s2:1 ...`, var_type 0), Rtab (var_type 1) and VCF (var_type 2, below); plain or
 * gzip files.
 * sample_names: the phenotype index order (bit i of a row = sample_names[i]).
 * psb_reader_next fills up to max_variants packed rows (bits, and missing when non-NULL; both
 * max_variants x words_per_row, zeroed by the call), the NUL-terminated names back to back in
 * `names` with name_off[v] their offsets, and info[v] (bit 0: row has missing genotypes,
 * bit 1: no observation in the selected samples, input.py:447-448).  *n_read == 0 at end of
 * file; *any_missing != 0 when at least one row had missing genotypes.  Only the first *n_read rows
 * of bits / missing are written. */
typedef struct psb_reader psb_reader;
int psb_reader_open(const char *path, int32_t var_type, const char *const *sample_names,
                    int32_t n_samples, psb_reader **out);
int psb_reader_next(psb_reader *reader, int64_t max_variants, uint32_t *bits, uint32_t *missing,
                    int32_t words_per_row, char *names, int64_t names_cap, int64_t *name_off,
                    int32_t *info, int64_t *n_read, int32_t *any_missing);
/* A batch shorter than max_variants ended either at the end of the file (*at_eof != 0) or because the
 * next name did not fit `names` (the line is kept for the next call; PSB_ERR_NOMEM when not even the
 * first name of a call fits: come back with a larger buffer). */
int psb_reader_at_eof(psb_reader *reader, int32_t *at_eof);
/* Compressed input (input.open_variant_file, input.py:268-298, reads it through Python's gzip): bgzip
 * files are inflated block-parallel; PLAIN gzip files of 4 MB and more -- one serial deflate stream --
 * are inflated on the reader's threads too (csrc/psb_pgz.cu: block starts searched per chunk, chunks
 * decoded to 16-bit symbols with markers for the unknown 32 KiB window, only the chunks the serial
 * chain confirms are kept, CRC-32 and length checked against the trailer; PSB_PGZ=0 selects zlib).
 * psb_pgz_selftest inflates a whole gzip file that way on n_threads threads with chunk_bytes
 * compressed bytes per work item (0: default 1 MiB) and reports the CRC-32 (NULL: not computed) and
 * length of the text and stats[0..1] = work items decoded / thrown away; < 0 on a corrupt stream. */
int psb_pgz_selftest(const char *path, int32_t n_threads, int64_t chunk_bytes, uint32_t *crc_out,
                     int64_t *len_out, int64_t stats_out[2]);
/* ---- k-mer text parsed on the device ------------------------------------------------ */
/* input.py's k-mer streaming (pyseer/input.py:505-707 -> read_variant :377-388, :438-452) as a
 * page-locked host -> device staging pipeline: the host cuts the decompressed text into lines,
 * the GPU tokenises them.
 * psb_reader_next_text (var_type 0 readers): fills dst (dst_cap bytes; allocate it with
 * psb_host_alloc) with the text of up to max_lines lines, read straight into it (plain text:
 * pread on the reader's threads; bgzip: block-parallel inflate; gzip: chunk-parallel inflate, above).  line_start[v] /
 * line_len[v] locate line v inside dst without its newline and trailing blanks (empty lines are
 * skipped); names / name_off as psb_reader_next.  The *n_read lines occupy dst[0, *n_bytes).  A short
 * batch ends at the end of the file (psb_reader_at_eof) or where dst / names are full -- then *n_read is
 * a multiple of line_multiple (PSB_ERR_NOMEM if not even that many lines fit).  Do not mix with
 * psb_reader_next on the same reader.
 * psb_text_setup: device lookup table of the sample names (phenotype order), once per context after
 * the model set-up.  psb_submit_text: psb_submit for such a batch -- copies text / line_start /
 * line_len on the staging copy stream and parses every line there into the packed row psb_submit
 * would have been given (samples after the first '|', blank-separated tokens, name up to ':',
 * unknown samples ignored); `text` must stay valid until the batch has been fetched.
 * psb_text_info (after psb_run_*): per line, bit 1 (value 2) = no observation in the selected samples
 * (input.py:447-448), bit 2 (value 4) = no '|' separator on the line (psb_reader_next: PSB_ERR_ARG). */
int psb_reader_next_text(psb_reader *reader, int64_t max_lines, int64_t line_multiple, char *dst,
                         int64_t dst_cap, int64_t *line_start, int32_t *line_len, char *names,
                         int64_t names_cap, int64_t *name_off, int64_t *n_read, int64_t *n_bytes);
int psb_text_setup(psb_ctx *ctx, const char *const *sample_names, int32_t n_samples);
int psb_submit_text(psb_ctx *ctx, const char *text, int64_t text_bytes, const int64_t *line_start,
                    const int32_t *line_len, int64_t n_lines);
int psb_text_info(psb_ctx *ctx, int32_t *info, int64_t n_lines);
/* var_type 2 = VCF text (plain or gzip), input.read_vcf_var (input.py:457-502), dominant encoding:
 * a sample carries the variant when a haplotype of its GT is a non-reference allele, '.' haplotypes
 * mark it missing unless a called one follows; names are CHROM_POS_REF[_ALT]; info bit 2 (value 4):
 * record with more than one ALT allele, bit 3 (value 8): FILTER neither empty nor PASS -- both are
 * skipped by the reference and come back as empty rows.  psb_reader_vcf_info returns contig, 1-based
 * position and REF length of the records of the last psb_reader_next (burden-region lookup,
 * input.py:395-407). */
int psb_reader_vcf_info(psb_reader *reader, int64_t n_records, char *contigs, int64_t contigs_cap,
                        int64_t *contig_off, int64_t *pos, int32_t *ref_len);
/* parser threads for psb_reader_next (pyseer's --cpu): lines are read serially, parsed in parallel */
int psb_reader_set_threads(psb_reader *reader, int32_t n_threads);
int psb_reader_close(psb_reader *reader);

/* ---- result formatting --------------------------------------------------------- */
/* TSV lines of utils.format_output (utils.py:39-105) for a whole fetched result table, in the order
 * and with the counters of the result loop of main() (__main__.py:547-568, 783-803): model 0 =
 * fixed effects (af, filter-pvalue, lrt-pvalue, beta, beta-std-err, intercept, n_betas slopes,
 * notes), 1 = LMM (..., beta-std-err, variant_h2, notes; inside every block of block_size variants
 * the pre-filtered ones come first, lmm.py:158-226).  Numbers as '%.2E', empty when not finite or
 * not set by the model; notes from the flag bits.  names / name_off: NUL-terminated variant names
 * back to back and their offsets; cols: HOST pointers.  counts[0..2] += pre-filtered, tested,
 * printed.  n_threads > 1 formats ranges of whole blocks in parallel.  Host-only; covers runs without
 * --print-samples / --lineage. */
int psb_format_rows(int32_t model, int64_t n_variants, const char *names, const int64_t *name_off,
                    const psb_results *cols, int32_t n_betas, int32_t block_size,
                    int32_t print_filtered, int32_t n_threads, char *out, int64_t out_cap,
                    int64_t *out_len, int64_t counts[3]);
/* The same with the lineage column of --lineage runs (utils.py:93-97) between the coefficients and the
 * notes: lineage[v] >= 0 indexes the n_lineages NUL-terminated names back to back in lineage_names
 * (lineage_off their offsets), a negative index prints NA. */
int psb_format_rows_lineage(int32_t model, int64_t n_variants, const char *names, const int64_t *name_off,
                            const psb_results *cols, int32_t n_betas, int32_t block_size,
                            int32_t print_filtered, int32_t n_threads, const int32_t *lineage,
                            const char *lineage_names, const int64_t *lineage_off, int32_t n_lineages,
                            char *out, int64_t out_cap, int64_t *out_len, int64_t counts[3]);

/* Pattern hashes of --output-patterns: input.hash_pattern (input.py:710-723) of the vector k the
 * reference builds from a variant (input.py:450; int64, or float64 with NaN when genotypes are
 * missing), from host packed rows.  25 bytes per hashed row in `out` (24 base64 characters + '\n');
 * rows whose flags carry PSB_F_PREFILTER are skipped when `flags` is given (__main__.py:559-560). */
int psb_hash_patterns(const uint32_t *bits, const uint32_t *missing, int64_t n_variants,
                      int32_t words_per_row, int32_t n_samples, const uint32_t *flags, char *out,
                      int64_t *n_out);
/* The same hashes computed on the device from the rows of the batch last submitted / run (one thread
 * per row, the 8 N byte message generated from the packed bits on the fly): out = n_variants x 16
 * bytes, the MD5 digest of each row's vector k (int64 0/1, or float64 with NaN when the row has a
 * missing genotype); base64 of a digest + '\n' is the reference's 25-byte entry. */
int psb_pattern_digests(psb_ctx *ctx, uint8_t *out);

/* ---- measurement ------------------------------------------------------------- */
/* Work counters of the last psb_run_fixed: [0] Newton evaluations (passes over the samples)
 * summed over variants, [1] variants handed to the Firth kernel, [2] variants that failed
 * the lrt filter. */
int psb_last_stats(psb_ctx *ctx, int64_t out[4]);
/* CUDA-event timers on the context stream.  which: 0 = whole last psb_run_*,
 * 1 = dominant kernel of the last run (LMM: the rotation/quadratic-form contraction;
 * fixed effects: the regression kernel).  Returns milliseconds. */
int psb_last_ms(psb_ctx *ctx, int32_t which, float *ms);
/* User event slots (0..7) recorded on the context stream, for timing a region that spans
 * several calls with CUDA events; psb_event_elapsed synchronises on slot b. */
int psb_event_record(psb_ctx *ctx, int32_t slot);
int psb_event_elapsed(psb_ctx *ctx, int32_t slot_a, int32_t slot_b, float *ms);
/* number of kernels the library launched on this context since creation */
int psb_launch_count(psb_ctx *ctx, int64_t *n);
/* Measured peak rates of the pipes the hot kernels are bound by, under the clocks the board holds
 * right now (roofline denominators): out[0] = dense int8 tensor TOP/s (tcgen05.mma kind::i8,
 * M128 N256 K32, A from tensor memory, issued back to back on all SMs; accumulators verified),
 * out[1] = fp64 FMA TFLOP/s (8 independent chains per thread), out[2], out[3] = the two probes'
 * durations in ms. */
int psb_measure_peaks(psb_ctx *ctx, double out[4]);

/* ---- synthetic inputs (bench / tests) ----------------------------------------- */
/* Fills a library-owned device buffer with seeded Bernoulli(af_s) rows, af_s ~
 * U(af_lo, af_hi); variant id = first_variant + row (counter-based, so shards do not
 * depend on the GPU count).  planted_every > 0 plants a phenotype-correlated variant at
 * ids divisible by it (needs y_sign: N int8 of +1/-1/0, host pointer); separated_every > 0 makes
 * ids = separated_every / 2 (mod separated_every) rare variants carried by positive-sign samples
 * only (an empty cell of the 2x2 table: 'bad-chisq' -> Firth regression, model.py:326).  The rows are
 * left submitted (as by psb_submit_device).  Same generator on host: psb_synth_host. */
int psb_synth_device(psb_ctx *ctx, uint64_t seed, int64_t first_variant,
                     int64_t n_variants, int32_t n_samples, double af_lo, double af_hi,
                     int32_t planted_every, int32_t separated_every, const int8_t *y_sign);
int psb_synth_host(uint64_t seed, int64_t first_variant, int64_t n_variants,
                   int32_t n_samples, double af_lo, double af_hi, int32_t planted_every,
                   int32_t separated_every, const int8_t *y_sign, uint32_t *out_bits,
                   int32_t words_per_row);

/* ---- host-evaluated special functions (same code the kernels run; CPU tests) ---- */
double psb_host_chi2_sf1(double x);                 /* scipy.stats.chi2.sf(x, 1)        */
double psb_host_f_sf_1(double x, double dfd);       /* scipy.stats.f.sf(x, 1, dfd)      */
double psb_host_t2_sf(double t, double df);         /* 2 * scipy.stats.t.sf(|t|, df)    */

#ifdef __cplusplus
}
#endif
#endif /* PYSEER_B200_H */
