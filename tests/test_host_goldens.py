"""Host side of the path against vectors produced by the UNMODIFIED reference (oracle/gen_golden_host.py,
run where /root/reference exists): pyseer.utils.format_output, pyseer.input.hash_pattern,
load_phenotypes / load_structure / load_covariates / load_lineage, read_variant over the k-mer, Rtab
and VCF fixtures including the burden branch, and pyseer.cmdscale."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope='module')
def hg():
    with open(os.path.join(GOLDEN, 'host_goldens.json')) as fh:
        return json.load(fh)


@pytest.fixture(scope='module')
def hz():
    with np.load(os.path.join(GOLDEN, 'host_goldens.npz')) as d:
        return {k: d[k] for k in d.files}


def _num(x):
    # non-finite values are stored as their repr ('nan', 'inf')
    return float(x) if isinstance(x, str) and x in ('nan', 'inf', '-inf') else x


def _pheno():
    from pyseer_b200.input import load_phenotypes
    return load_phenotypes(os.path.join(GOLDEN, 'subset.pheno'), None)


def test_format_output(hg):
    """utils.py:39-105, every (model, lineage, print_samples) combination, NaN / inf fields."""
    from pyseer_b200.classes import Seer, LMM
    from pyseer_b200.utils import format_output
    lineages = ['MDS1', 'MDS2', 'BAPS_3']
    for case in hg['format_output']:
        f = {k: (_num(v) if not isinstance(v, list) else v) for k, v in case['fields'].items()}
        betas = np.array([_num(b) for b in f['betas']], dtype=float)
        ml = f['max_lineage']
        if isinstance(ml, float) and np.isfinite(ml):
            ml = int(ml)          # lineage indices are stored as JSON numbers
        seer = Seer(f['kmer'], 'pat', f['af'], f['prep'], f['pvalue'], f['kbeta'], f['bse'], f['intercept'],
                    betas, ml, f['kstrains'], f['nkstrains'], f['notes'], False, False)
        lmm = LMM(f['kmer'], 'pat', f['af'], f['prep'], f['pvalue'], f['kbeta'], f['bse'], f['frac_h2'], ml,
                  f['kstrains'], f['nkstrains'], f['notes'], False, False)
        for key, want in case['lines'].items():
            model, lin, smp = key.split('|')
            item = lmm if model == 'lmm' else seer
            got = format_output(item, lineages if lin == 'lin' else None, model=model,
                                print_samples=(smp == 'samples'))
            assert got == want, (f['kmer'], key)


def test_hash_pattern(hg):
    from pyseer_b200.input import hash_pattern
    for c in hg['hash_pattern']:
        k = np.array([np.nan if x is None else x for x in c['k']], dtype=c['dtype'])
        assert hash_pattern(k).decode() == c['hash']


def test_loaders(hg, hz):
    from pyseer_b200.input import load_structure, load_covariates, load_lineage
    p = _pheno()
    assert [str(x) for x in p.index] == hg['phenotypes']['index']
    assert np.array_equal(p.values.astype(float), np.array(hg['phenotypes']['values']))
    with contextlib.redirect_stderr(io.StringIO()):
        m = load_structure(os.path.join(GOLDEN, 'distances50.tsv'), p, 10, 'classic', 1, None)
        lin, labels = load_lineage(os.path.join(GOLDEN, 'lineage_clusters50.txt'), p)
    m = np.asarray(getattr(m, 'values', m), dtype=float)
    assert m.shape == hz['structure_m'].shape
    assert np.allclose(m, hz['structure_m'], rtol=1e-9, atol=1e-12)
    cov = load_covariates(os.path.join(GOLDEN, 'covariates.txt'), ['2q', '3'], p)
    assert [str(c) for c in cov.columns] == hg['covariates']['columns']
    assert [str(i) for i in cov.index] == hg['covariates']['index']
    assert np.array_equal(np.asarray(cov.values, dtype=float), hz['covariates'])
    assert [str(x) for x in labels] == hg['lineage']['labels']
    assert np.array_equal(np.asarray(lin, dtype=float), hz['lineage'])


def test_cmdscale(hz):
    from pyseer_b200.cmdscale import cmdscale
    for i in range(3):
        Y, e = cmdscale(hz['cmdscale_D%d' % i])
        assert np.allclose(e, hz['cmdscale_e%d' % i], rtol=1e-9, atol=1e-12)
        assert np.allclose(Y, hz['cmdscale_Y%d' % i], rtol=1e-8, atol=1e-10)


def _check_rows(reader, batches, rows, n):
    from pyseer_b200.input import hash_pattern
    j = 0
    for b in batches:
        for i in range(b.n):
            want = rows[j]
            j += 1
            if want['name'] is None:
                # the reference's None sentinel: counted as loaded, ends up AF-filtered
                assert b.names[i] == 'NA' and not b.bits[i].any()
                continue
            assert b.names[i] == want['name']
            k = reader.k_vector(b, i)
            assert hash_pattern(k).decode() == want['hash'], want['name']
            ks, nks = reader.sample_lists(b, i)
            assert len(ks) == want['carriers']
            assert len(ks) / float(n) == want['af']
            miss = 0 if b.missing is None else int(np.unpackbits(b.missing[i].view(np.uint8)).sum())
            assert miss / float(n) == want['missing']
    assert j == len(rows)


def test_kmer_and_rtab_readers_match_read_variant(hg):
    """input.read_variant (input.py:301-454) run by the reference itself on the fixtures."""
    from pyseer_b200.input import VariantReader
    p = _pheno()
    n = len(p.index)
    for var_type, fn, key in (('kmers', 'kmers.gz', 'kmers'), ('Rtab', 'presence_absence.Rtab.gz', 'rtab')):
        rd = VariantReader(var_type, os.path.join(GOLDEN, fn), p)
        err = io.StringIO()
        with contextlib.redirect_stderr(err):
            batches = list(rd.batches(97))
        _check_rows(rd, batches, hg[key]['rows'], n)
        rd.close()
        assert err.getvalue() == hg[key]['stderr']


def test_vcf_reader_matches_read_vcf_var(hg):
    """input.read_vcf_var (input.py:457-502) and the burden branch (input.py:395-411), run by the
    reference on the VCF fixture with a stand-in for pysam's records."""
    from oracle import input_oracle
    from pyseer_b200.input import VcfReader
    p = _pheno()
    n = len(p.index)
    rd = VcfReader(os.path.join(GOLDEN, 'variants50.vcf.gz'), p)
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        batches = list(rd.batches(300))
    _check_rows(rd, batches, hg['vcf']['rows'], n)
    assert err.getvalue() == hg['vcf']['stderr']
    for key, fn in (('burden', 'burden_regions.txt'), ('burden_multiple', 'burden_regions_multiple.txt')):
        rd = VcfReader(os.path.join(GOLDEN, 'variants50.vcf.gz'), p, os.path.join(GOLDEN, fn),
                       reducer=input_oracle.burden_union)
        with contextlib.redirect_stderr(io.StringIO()):
            batches = list(rd.batches(100))
        _check_rows(rd, batches, hg[key]['rows'], n)


def test_native_pattern_hashes(hg):
    """psb_hash_patterns against the reference's hash_pattern on the k-mer, Rtab and VCF fixtures
    (int64 vectors and float64 vectors with NaN), and the flags filter."""
    from pyseer_b200 import _lib
    from pyseer_b200.input import VariantReader, VcfReader, hash_patterns, hash_pattern
    p = _pheno()
    n = len(p.index)
    for kind, fn, key in (('kmers', 'kmers.gz', 'kmers'), ('Rtab', 'presence_absence.Rtab.gz', 'rtab'),
                          ('vcf', 'variants50.vcf.gz', 'vcf')):
        rd = VcfReader(os.path.join(GOLDEN, fn), p) if kind == 'vcf' else \
            VariantReader(kind, os.path.join(GOLDEN, fn), p)
        with contextlib.redirect_stderr(io.StringIO()):
            batches = list(rd.batches(400))
        rows = hg[key]['rows']
        j = 0
        n_missing_rows = 0
        for b in batches:
            blob = hash_patterns(b.bits, b.missing, n)
            assert len(blob) == 25 * b.n
            for i in range(b.n):
                want = rows[j]
                j += 1
                got = blob[25 * i:25 * i + 25]
                assert got == hash_pattern(rd.k_vector(b, i))
                if want['name'] is not None:
                    assert got.decode() == want['hash'], want['name']
                    n_missing_rows += want['missing'] > 0
            # flags: pre-filtered rows are left out
            flags = np.zeros(b.n, dtype=np.uint32)
            flags[::3] = _lib.F_PREFILTER
            sub = hash_patterns(b.bits, b.missing, n, flags)
            keep = [i for i in range(b.n) if i % 3]
            assert sub == b''.join(blob[25 * i:25 * i + 25] for i in keep)
        rd.close()
        if kind == 'vcf':
            assert n_missing_rows > 0          # the float64 / NaN branch is exercised
