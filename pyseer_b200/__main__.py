"""``python -m pyseer_b200`` -- pyseer's association CLI (pyseer/__main__.py) with the per-variant
loop running on the GPU.

Same options, same TSV on stdout, same counters on stderr for the fixed-effects (SEER) and
``--lmm`` models with ``--kmers`` / ``--pres`` / ``--vcf`` (``--burden``) input.  What the reference does with a
``multiprocessing.Pool`` over variants (``__main__.py:517-593, 762-827``) is done here by
submitting blocks of packed variants to the engine; ``--cpu`` selects the parser threads of
the native k-mer / Rtab reader, ``--bits-cache`` keeps the packed rows for later runs.
Whole-genome models (``--wg``) are not part of this path and are rejected with a message.
"""
import argparse
from binascii import b2a_base64
import operator
import os
import sys

import numpy as np
import pandas as pd

from . import __version__
from . import _lib
from . import classes as var_obj
from .engine import notes_from_flags
from .input import (VariantReader, CachedVariantReader, open_variants, VcfReader, hash_pattern, hash_patterns, load_covariates, load_lineage,
                    load_phenotypes, load_structure)
from .utils import format_output, format_table


def get_options(argv=None):
    ap = argparse.ArgumentParser(prog='pyseer_b200',
                                 description='SEER / LMM association tests on B200 (pyseer CLI)')
    ph = ap.add_argument_group('Phenotype')
    ph.add_argument('--phenotypes', required=True)
    ph.add_argument('--phenotype-column', default=None)
    va = ap.add_argument_group('Variants')
    g = va.add_mutually_exclusive_group(required=True)
    g.add_argument('--kmers', default=None)
    g.add_argument('--vcf', default=None)
    g.add_argument('--pres', default=None)
    va.add_argument('--burden')
    di = ap.add_argument_group('Distances')
    d = di.add_mutually_exclusive_group()
    d.add_argument('--distances')
    d.add_argument('--load-m')
    di.add_argument('--similarity')
    di.add_argument('--load-lmm')
    di.add_argument('--save-m')
    di.add_argument('--save-lmm')
    di.add_argument('--mds', default='classic')
    di.add_argument('--max-dimensions', type=int, default=10)
    di.add_argument('--no-distances', action='store_true', default=False)
    mo = ap.add_argument_group('Association options')
    mo.add_argument('--continuous', action='store_true', default=False)
    mo.add_argument('--lmm', action='store_true', default=False)
    mo.add_argument('--wg', default=None)
    mo.add_argument('--lineage', action='store_true', default=False)
    mo.add_argument('--lineage-clusters', default=None)
    mo.add_argument('--lineage-file', default='lineage_effects.txt')
    fi = ap.add_argument_group('Filtering options')
    fi.add_argument('--min-af', type=float, default=0.01)
    fi.add_argument('--max-af', type=float, default=0.99)
    fi.add_argument('--max-missing', type=float, default=0.05)
    fi.add_argument('--filter-pvalue', type=float, default=1)
    fi.add_argument('--lrt-pvalue', type=float, default=1)
    co = ap.add_argument_group('Covariates')
    co.add_argument('--covariates', default=None)
    co.add_argument('--use-covariates', default=None, nargs='*')
    ot = ap.add_argument_group('Other')
    ot.add_argument('--print-samples', action='store_true', default=False)
    ot.add_argument('--print-filtered', action='store_true', default=False)
    ot.add_argument('--output-patterns', default=False)
    ot.add_argument('--uncompressed', action='store_true', default=False)
    ot.add_argument('--bits-cache', default=None,
                    help='packed binary cache of the --kmers / --pres file: written on the first run, '
                         'read instead of parsing the text on later runs with the same samples')
    ot.add_argument('--cpu', type=int, default=1, help='parser threads of the native k-mer / Rtab reader')
    ot.add_argument('--block_size', type=int, default=3000)
    ot.add_argument('--gpu', type=int, default=0, help='CUDA device index (the first one with --gpus)')
    ot.add_argument('--gpus', type=int, default=1,
                    help='number of GPUs: batches of variants are dealt to the GPUs in input order and the '
                         'result tables gathered over NCCL on the first one (what --cpu is to pyseer)')
    ot.add_argument('--gpu-batch', type=int, default=None,
                    help='variants per GPU submission (rounded to a multiple of --block_size); default '
                         '48000, or 12000 for k-mer text tokenised on the device (its page-locked text '
                         'buffers stay small and more batches are in flight)')
    ot.add_argument('--lmm-precision', type=int, default=None,
                    help='0 = FP64 contraction, 3..7 = exact int8-slice tensor-core contraction, 46 = two '
                         'passes: 4 slices, 6 again for the far tail (default)')
    ot.add_argument('--version', action='version', version='%(prog)s ' + __version__)
    return ap.parse_args(argv)


def _die(msg):
    sys.stderr.write(msg + '\n')
    sys.exit(1)


def main(argv=None):
    o = get_options(argv)
    # option checks of __main__.py:257-306, in the reference's order and with its messages
    if o.lmm and o.wg:
        _die('Choose only one alternative model. Either --lmm, --wg or neither')
    if o.wg:
        _die('Whole-genome models (--wg) are not part of the GPU path; use pyseer for them')
    if o.max_dimensions < 1:
        _die('Minimum number of dimensions after MDS is 1')
    if o.burden and not o.vcf:
        _die('Burden test can only be performed with VCF input')
    if o.lmm and not o.similarity and not o.load_lmm:
        _die('Must provide a similarity matrix or lmm cache for random effects')
    if not o.no_distances:
        if (o.lmm and (o.distances or o.load_m) and not o.lineage) or \
                (not o.lmm and (o.similarity or o.load_lmm)):
            _die('Must use distance matrix with fixed effects, or similarity matrix with random '
                 'effects\nUnless performing a lineage analysis with random effects')
        if o.lmm and not (o.distances or o.load_m) and o.lineage:
            _die('Must also provide a distance matrix to report lineage effects')
        if not o.lmm and not o.distances and not o.load_m:
            _die('Option --no-distances must be used when no distance matrix is provided')
    else:
        if not o.lmm and not o.lineage_clusters and o.lineage:
            _die('Must provide a lineage clusters file when --no-distances and --lineage are used '
                 'together in fixed-effects mode')
        if o.distances or o.load_m:
            _die('Cannot use --no-distances with --distances or --load-m')
        if o.lmm:
            _die('Cannot use --no-distances with --lmm')
    if o.block_size < 1:
        _die('Block size must be at least 1')
    if o.gpus < 1:
        _die('--gpus must be at least 1')

    p = load_phenotypes(o.phenotypes, o.phenotype_column)
    sys.stderr.write('Read ' + str(len(p)) + ' phenotypes\n')
    if not o.continuous:
        if p.values[(p.values != 0) & (p.values != 1)].size > 0:
            o.continuous = True
            sys.stderr.write('Detected continuous phenotype\n')
        else:
            sys.stderr.write('Detected binary phenotype\n')

    if o.covariates is not None:
        cov = load_covariates(o.covariates, o.use_covariates, p)
        if cov is None:
            sys.exit(1)
    else:
        cov = pd.DataFrame([])

    from . import model as fx
    model = None
    lmm = None
    m = np.empty(shape=(0, 0))
    null_fit = firth_null = None
    # fixed effects, or lineage effects from the MDS components, need p ~ m
    if (o.lineage and not o.lineage_clusters) or not o.lmm:
        if not o.no_distances:
            if o.load_m and os.path.isfile(o.load_m):
                m = pd.read_pickle(o.load_m)
                sys.stderr.write('Loaded projection with dimension ' + str(m.shape) + '\n')
            else:
                seed = os.environ.get('PYSEERSEED', None)
                m = load_structure(o.distances, p, o.max_dimensions, o.mds, o.cpu,
                                   int(seed) if seed is not None else None)
                if o.save_m:
                    m.to_pickle(o.save_m + '.pkl')
            if o.max_dimensions > m.shape[1]:
                sys.stderr.write('Population MDS scaling restricted to %d dimensions instead of '
                                 'requested %d\n' % (m.shape[1], o.max_dimensions))
                o.max_dimensions = m.shape[1]
            common = p.index.intersection(m.index)
            sys.stderr.write('Analysing ' + str(len(common)) + ' samples found in both phenotype '
                             'and structure matrix\n')
            p = p.loc[common]
            m = m.loc[p.index].values[:, :o.max_dimensions]
        if cov.shape[1] > 0:
            cov = cov.loc[p.index]
        null_fit = fx.fit_null(p.values, m, cov, o.continuous, device=o.gpu)
        firth_null = fx.fit_null(p.values, m, cov, o.continuous, True, device=o.gpu) \
            if (not o.continuous and not o.lmm) else True
        if null_fit is None or firth_null is None:
            _die('Could not fit null model, exiting')

    # lineage effects (__main__.py:388-446)
    lineage_clusters = None
    lineage_dict = None
    lineage_samples = None
    if o.lineage_clusters:
        lineage_clusters, lineage_dict = load_lineage(o.lineage_clusters, p)
    if o.lineage:
        lineage_samples = p.index
        lineage_wald = {}
        if o.lineage_clusters:
            # the cluster design is not full rank: drop the lineage least associated with
            # the phenotype (single-predictor fits, the clusters being orthogonal)
            for lineage, design in zip(lineage_dict, lineage_clusters.T):
                fit = fx.fit_null(p.values, design.reshape(-1, 1).astype(float), cov, o.continuous,
                                  device=o.gpu)
                if fit is None:
                    _die('Could not fit lineage null model, exiting')
                lineage_wald[lineage] = np.absolute(fit.params[1]) / fit.bse[1]
            min_lineage = min(lineage_wald.items(), key=operator.itemgetter(1))[0]
            min_index = lineage_dict.index(min_lineage)
            lineage_clusters = np.delete(lineage_clusters, min_index, 1)
            del lineage_dict[min_index]
        else:
            lineage_dict = ['MDS' + str(i + 1) for i in range(o.max_dimensions)]
            lineage_clusters = m
            for lineage, slope, se in zip(lineage_dict, null_fit.params[1:], null_fit.bse[1:]):
                lineage_wald[lineage] = np.absolute(slope) / se
        sys.stderr.write('Writing lineage effects to %s\n' % o.lineage_file)
        from scipy.stats import norm
        with open(o.lineage_file, 'w') as lineage_out:
            lineage_out.write('\t'.join(['lineage', 'wald_test', 'p-value']) + '\n')
            for lineage, wald in sorted(lineage_wald.items(), key=operator.itemgetter(1),
                                        reverse=True):
                pval = 2 * (1 - norm.cdf(wald))
                lineage_out.write('\t'.join([lineage, str(wald), str(pval)]) + '\n')

    if not o.lmm:
        model = fx.FixedModel(p.values, m, cov, o.continuous, null_fit.llf,
                              firth_null if not o.continuous else 0.0, device=o.gpu,
                              lineage=(lineage_clusters, cov) if o.lineage else None)
    else:
        from . import lmm as lm
        sys.stderr.write('Setting up LMM\n')
        p, lmm, h2 = lm.initialise_lmm(p, cov, o.similarity, o.load_lmm, o.save_lmm, lineage_samples,
                                       device=o.gpu, precision=o.lmm_precision)
        sys.stderr.write('h^2 = ' + '{0:.2f}'.format(h2) + '\n')

    if o.vcf:
        # burden regions are reduced on the device (psb_submit_burden) through the context the
        # model already owns
        eng = (lmm.engine(h2) if o.lmm else model.engine) if o.burden else None

        def device_union(vbits, vmiss, offsets, members):
            eng.submit_burden(vbits, vmiss, offsets, members)
            return eng.download_rows()

        reader = VcfReader(o.vcf, p, o.burden, reducer=device_union if o.burden else None,
                           threads=o.cpu)
    else:
        reader = open_variants('kmers' if o.kmers else 'Rtab', o.kmers or o.pres, p, o.uncompressed,
                               cache=o.bits_cache, threads=o.cpu)

    header = ['variant', 'af', 'filter-pvalue', 'lrt-pvalue', 'beta', 'beta-std-err']
    if not o.lmm:
        header.append('intercept')
        if not o.no_distances:
            header += ['PC%d' % i for i in range(1, o.max_dimensions + 1)]
        if o.covariates is not None:
            header += [x for x in cov.columns]
    else:
        header.append('variant_h2')
    if o.lineage:
        header.append('lineage')
    if o.print_samples:
        header += ['k-samples', 'nk-samples']
    header.append('notes')
    print('\t'.join(header))

    patterns = open(o.output_patterns, 'wb') if o.output_patterns else None
    out = sys.stdout
    nan = np.nan
    model_name = 'lmm' if o.lmm else 'seer'
    # k-mer text is tokenised on the device (psb_submit_text) unless sample lists are printed; what else
    # needs the packed rows on the host (the packed cache being written, the LMM's per-block lineage fits)
    # gets them back per batch; pattern hashes and fixed-effects lineage fits happen where the rows are
    text_mode = type(reader) is VariantReader and reader.var_type == 'kmers' and not o.print_samples and \
        os.environ.get('PYSEER_B200_TEXT', '1') != '0'
    # --bits-cache being written by this run: from the rows the device parsed, brought back per batch
    cache_writer = getattr(reader, 'cache_writer', None) if text_mode else None
    # the rows come back from the device for the cache, and for the LMM's per-block lineage fits
    rows_back = cache_writer is not None or (text_mode and o.lmm and o.lineage)
    # measured at N = 5000 (profiles/r02_cli_batch_sweep.json): plain text streams at 84 k variants/s in
    # batches of 24000 lines, 550 k/s in batches of 12000, 640 k/s in batches of 6000
    gpu_batch = o.gpu_batch if o.gpu_batch else (12000 if text_mode else 48000)
    gpu_batch = max(1, gpu_batch // o.block_size) * o.block_size
    counters = {'prefilter': 0, 'tested': 0, 'printed': 0}

    def samples_of(batch, j):
        return reader.sample_lists(batch, j) if o.print_samples else ([], [])

    def emit(batch, r):
        """Result loop of main() (__main__.py:547-568, 783-803) for one batch."""
        flags = r.flags
        if cache_writer is not None and batch.text is not None:
            cache_writer.add(batch)
        if batch.text is not None or batch.report_empty:
            # batch parsed on the device, or read from the packed cache: what the host reader reports while
            # it reads (a row without missing genotypes: "no observation" is carriers == 0)
            for i in np.nonzero(r.carriers[:batch.n] == 0)[0]:
                sys.stderr.write('No observations of ' + batch.names[i] + ' in selected samples\n')
        if batch.skipped is not None and batch.skipped.any():
            # records the reference never hands to a model (k is None, input.py:603-611)
            flags[batch.skipped] = _lib.F_AF_FILTER | _lib.F_PREFILTER
        if not o.print_samples and os.environ.get('PYSEER_B200_NATIVE_FORMAT', '1') != '0':
            # no per-variant sample lists asked for: the whole batch goes through the library's formatter
            # (psb_format_rows[_lineage]: same lines, order and counters as the loop below, far faster)
            lin = None
            if o.lineage:
                # the lineage column as the loop below fills it: NA (-1) for pre-filtered variants; LMM: the
                # lineage of the block's LAST variant for every fitted variant of the block (lmm.py:160-162
                # / :209-211, the loop variable `k` is reused), NA for the lrt-filtered ones; fixed effects:
                # the variant's own lineage (model.py:379-380)
                fl = np.asarray(flags[:batch.n])
                lin = np.full(batch.n, -1, dtype=np.int32)
                if o.lmm:
                    fitted = (fl & (_lib.F_PREFILTER | _lib.F_FILTER)) == 0
                    for b0 in range(0, batch.n, o.block_size):
                        b1 = min(b0 + o.block_size, batch.n)
                        bl = fx.fit_lineage_effect(lineage_clusters, cov.values, reader.k_vector(batch, b1 - 1),
                                                   device=o.gpu)
                        if bl is not None and np.isfinite(bl):
                            lin[b0:b1][fitted[b0:b1]] = int(bl)
                else:
                    rl = np.asarray(r.lineage[:batch.n])
                    sel = ((fl & _lib.F_PREFILTER) == 0) & (rl >= 0)
                    lin[sel] = rl[sel]
            text, n_pre, n_tested, n_printed = format_table(r, batch.names, model_name, o.block_size,
                                                            o.print_filtered, threads=o.cpu, lineage=lin,
                                                            lineage_names=lineage_dict if o.lineage else None)
            counters['prefilter'] += n_pre
            counters['tested'] += n_tested
            counters['printed'] += n_printed
            if hasattr(out, 'buffer'):          # a real stream: the bytes go out as they are
                out.flush()
                out.buffer.write(text)
            else:
                out.write(text.decode())
            if patterns is not None:
                # hash_pattern of every tested variant, in input order (__main__.py:559-560)
                if batch.digests is not None:
                    keep = (np.asarray(flags[:batch.n]) & _lib.F_PREFILTER) == 0
                    patterns.write(b''.join(b2a_base64(d.tobytes()) for d in batch.digests[:batch.n][keep]))
                else:
                    patterns.write(hash_patterns(batch.bits, batch.missing, reader.n_samples, flags))
            return
        # the reference emits each block of --block_size variants as: filtered ones first
        # (LMM only, lmm.py:158-226), then the tested ones; fixed effects keep input order
        for b0 in range(0, batch.n, o.block_size):
            idx = range(b0, min(b0 + o.block_size, batch.n))
            block_lineage = None
            if o.lmm:
                order = [j for j in idx if flags[j] & _lib.F_PREFILTER] + \
                        [j for j in idx if not (flags[j] & _lib.F_PREFILTER)]
                if o.lineage:
                    # lmm.py:160-162 / :209-211: every variant of a block is reported with the
                    # lineage of the block's LAST variant (the loop variable `k` is reused)
                    block_lineage = fx.fit_lineage_effect(lineage_clusters, cov.values,
                                                          reader.k_vector(batch, idx[-1]), device=o.gpu)
            else:
                order = idx
            for j in order:
                f = int(flags[j])
                if f & _lib.F_PREFILTER:
                    counters['prefilter'] += 1
                    if not o.print_filtered:
                        continue
                else:
                    counters['tested'] += 1
                    if patterns is not None:
                        patterns.write(b2a_base64(batch.digests[j].tobytes()) if batch.digests is not None
                                       else hash_pattern(reader.k_vector(batch, j)))
                    if (f & _lib.F_FILTER) and not o.print_filtered:
                        continue
                ks, nks = samples_of(batch, j)
                notes = notes_from_flags(f)
                if o.lmm:
                    if f & _lib.F_PREFILTER:
                        item = var_obj.LMM(batch.names[j], None, r.af[j], r.prep[j], nan, nan, nan, nan,
                                           None, ks, nks, notes, True, False)
                    elif f & _lib.F_FILTER:
                        item = var_obj.LMM(batch.names[j], None, r.af[j], r.prep[j], r.pvalue[j], nan,
                                           nan, nan, None, ks, nks, notes, False, True)
                    else:
                        item = var_obj.LMM(batch.names[j], None, r.af[j], r.prep[j], r.pvalue[j],
                                           r.beta[j], r.bse[j], r.extra[j], block_lineage, ks, nks,
                                           notes, False, False)
                else:
                    item = fx.seer_from_row(r, j, batch.names[j], None, r.af[j], ks, nks)
                    if o.lineage and not (f & _lib.F_PREFILTER) and r.lineage[j] >= 0:
                        item = item._replace(max_lineage=int(r.lineage[j]))
                counters['printed'] += 1
                out.write(format_output(item, lineage_dict if o.lineage else None, model_name,
                                        o.print_samples) + '\n')

    # ---- the streaming pipeline (pipeline.py): reader thread -> pinned staging -> GPU(s) -> output
    # thread, all in input order ------------------------------------------------------------------
    import queue
    import threading
    from . import pipeline
    from .pipeline import BatchRunner, Prefetch
    from .engine import Engine
    n_gpus = max(1, o.gpus)
    thresholds = dict(min_af=o.min_af, max_af=o.max_af, max_missing=o.max_missing,
                      filter_pvalue=o.filter_pvalue, lrt_pvalue=o.lrt_pvalue, continuous=o.continuous)
    extra_models = []
    if o.lmm:
        engines = [lmm.engine(h2)]
        S_, U_ = lmm.getSU()
        for g in range(1, n_gpus):
            e = Engine(o.gpu + g)
            e.lmm_setup(lmm.X, lmm.Y[:, 0], U_, S_, h2, engines[0].lmm_precision)
            engines.append(e)
        run_one = lambda e: e.run_lmm(**thresholds)                                   # noqa: E731
        n_betas = 0
    else:
        engines = [model.engine]
        for g in range(1, n_gpus):
            extra_models.append(fx.FixedModel(p.values, m, cov, o.continuous, null_fit.llf,
                                              firth_null if not o.continuous else 0.0, device=o.gpu + g,
                                              lineage=(lineage_clusters, cov) if o.lineage else None))
            engines.append(extra_models[-1].engine)
        run_one = lambda e: e.run_fixed(**thresholds)                                 # noqa: E731
        n_betas = max(model.Z.shape[1] - 1, 0)
    comm = None
    if n_gpus > 1:
        from .comm import Comm
        comm = Comm.local(engines)
    runner = BatchRunner(engines, run_one, n_betas=n_betas,
                         lineage=(False if (o.lineage and not o.lmm) else None), comm=comm,
                         rows_max=gpu_batch, digests=patterns is not None, rows=rows_back)
    pool = None
    name_bytes = sum(len(x) for x in reader.samples) + 3 * reader.n_samples if text_mode else 0
    if text_mode and o.block_size * (name_bytes + 4096) > (2 << 30):
        text_mode = False      # one block of worst-case lines would not fit a 2 GB text buffer: host parser
        cache_writer = None    # ... which writes the cache while it reads
    if text_mode:
        # page-locked text buffers sized for a full batch at an allele frequency of 0.5 (a batch cut
        # short by its buffer keeps whole blocks, see psb_reader_next_text) and never smaller than one
        # block of lines that list every sample
        per_line = name_bytes // 2 + 512
        text_bytes = int(min(max(32 << 20, gpu_batch * per_line), 768 << 20))
        text_bytes = max(text_bytes, o.block_size * (name_bytes + 4096))
        for e in engines:
            e.text_setup(reader.samples)
        pool = pipeline.TextPool(2 * n_gpus + 2, gpu_batch, text_bytes)
        source = reader.text_batches(gpu_batch, block_size=o.block_size, pool=pool)
    elif isinstance(reader, CachedVariantReader):
        pool = pipeline.PinnedPool(5 * n_gpus + 3, gpu_batch, reader.W, reader.var_type == 'Rtab')
        source = reader.batches(gpu_batch, pool=pool, defer_empty=True)
    elif isinstance(reader, VariantReader):
        pool = pipeline.PinnedPool(5 * n_gpus + 3, gpu_batch, reader.W, reader.var_type == 'Rtab')
        source = reader.batches(gpu_batch, pool=pool)
    else:
        source = reader.batches(gpu_batch)
    batches = Prefetch(source, depth=n_gpus + 1)
    outq = queue.Queue(maxsize=2 * n_gpus)
    out_err = []

    def output_loop():
        while True:
            item = outq.get()
            if item is None:
                return
            if out_err:
                continue                    # keep draining so that the producer never blocks
            try:
                emit(*item)
                if pool is not None:
                    pool.put(item[0].token)
            except BaseException as e:      # noqa: BLE001 -- re-raised in the main thread
                out_err.append(e)

    writer = threading.Thread(target=output_loop, daemon=True)
    writer.start()
    import time
    t_stream = time.time()
    streamed = False
    try:
        for batch, r in runner.results(batches):
            if out_err:
                break
            outq.put((batch, r))
        streamed = True
    finally:
        batches.cancel()
        outq.put(None)
        writer.join()
        if cache_writer is not None:
            # complete only when every batch went through; an interrupted cache is discarded
            complete = streamed and not out_err
            cache_writer.close(complete=complete)
            if not complete:
                try:
                    os.unlink(reader.cache_path)
                except OSError:
                    pass
        runner.close()
        if comm is not None:
            comm.close()
        for e in engines[1:]:
            if o.lmm:
                e.close()
        for fm in extra_models:
            fm.close()
        if pool is not None:
            pool.close()
    if out_err:
        raise out_err[0]
    prefilter, tested, printed = counters['prefilter'], counters['tested'], counters['printed']
    if os.environ.get('PYSEER_B200_TIMING'):
        # streaming part only (reader -> GPU -> output), without the once-per-run set-up
        dt = time.time() - t_stream
        sys.stderr.write('pipeline: %d variants in %.3f s = %.0f variants/s\n'
                         % (prefilter + tested, dt, (prefilter + tested) / max(dt, 1e-9)))
    reader.close()
    if patterns is not None:
        patterns.close()
    if model is not None:
        model.close()
    if lmm is not None:
        lmm.close()

    sys.stderr.write('%d loaded variants\n' % (prefilter + tested))
    sys.stderr.write('%d pre-filtered variants\n' % prefilter)
    sys.stderr.write('%d tested variants\n' % tested)
    sys.stderr.write('%d printed variants\n' % printed)
    return 0


if __name__ == '__main__':
    sys.exit(main())
