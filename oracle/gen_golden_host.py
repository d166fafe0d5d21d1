#!/usr/bin/env python
"""Golden vectors for the HOST side of the path, produced by the UNMODIFIED reference.
TEST INFRASTRUCTURE.  Run in the build container only (reads /root/reference):

    python oracle/gen_golden_host.py

pysam is absent from the image; pyseer/input.py only needs the name `VariantFile` at import time,
so a stub module is installed before importing it.  Everything below calls the reference's own
functions -- nothing is restated here:

  host_goldens.json
    format_output     pyseer.utils.format_output (utils.py:39-105) on seeded Seer / LMM tuples
    hash_pattern      pyseer.input.hash_pattern (input.py:710-723)
    phenotypes / structure / covariates / lineage
                      load_phenotypes (:24-59), load_structure (:62-137), load_covariates
                      (:195-247), load_lineage (:139-177) on the fixtures under tests/golden/
    kmers / rtab      read_variant (:301-454) over the k-mer and Rtab fixtures: name, af, missing,
                      carriers, pattern hash, "No observations" messages
    vcf / burden      read_vcf_var (:457-502) and the burden branch of read_variant (:395-411) over
                      the VCF fixture, records served by a stand-in for pysam's fetch
  host_goldens.npz    cmdscale (cmdscale.py) on seeded distance matrices, MDS projections
"""
import contextlib
import gzip
import io
import json
import os
import re
import sys
import types
import warnings

import numpy as np
import pandas as pd

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

sys.path.insert(0, REF)
_ps = types.ModuleType('pysam')
_ps.VariantFile = object
sys.modules['pysam'] = _ps
warnings.simplefilter('ignore')
import pyseer.input as ri          # noqa: E402
import pyseer.utils as ru          # noqa: E402
import pyseer.cmdscale as rc       # noqa: E402
from pyseer.classes import Seer, LMM   # noqa: E402

ri.sys.stderr  # the module writes its messages to sys.stderr, captured below


def _f(x):
    return None if x is None else (float(x) if np.isfinite(x) else repr(float(x)))


def format_output_cases():
    rng = np.random.RandomState(11)
    out = []
    lineages = ['MDS1', 'MDS2', 'BAPS_3']
    for i in range(40):
        def val(scale=1.0):
            r = rng.uniform()
            if r < 0.12:
                return float('nan')
            if r < 0.16:
                return float('inf')
            return float(rng.normal() * scale * 10 ** rng.randint(-12, 6))
        nb = int(rng.randint(0, 5))
        ks = ['s%d' % j for j in sorted(rng.choice(20, rng.randint(0, 6), replace=False))]
        nks = ['s%d' % j for j in sorted(rng.choice(20, rng.randint(0, 6), replace=False))]
        notes = sorted(rng.choice(['bad-chisq', 'high-bse', 'af-filter', 'firth-fail'], rng.randint(0, 3),
                                  replace=False).tolist())
        ml = [None, 0, 1, 2, float('nan')][rng.randint(0, 5)]
        fields = dict(kmer='K%d' % i, af=abs(val()), prep=abs(val()), pvalue=abs(val()), kbeta=val(),
                      bse=abs(val()), intercept=val(), frac_h2=abs(val()),
                      betas=[val() for _ in range(nb)], max_lineage=ml, kstrains=ks, nkstrains=nks,
                      notes=notes)
        seer = Seer(fields['kmer'], 'pat', fields['af'], fields['prep'], fields['pvalue'], fields['kbeta'],
                    fields['bse'], fields['intercept'], np.array(fields['betas']), ml, ks, nks,
                    notes, False, False)
        lmm = LMM(fields['kmer'], 'pat', fields['af'], fields['prep'], fields['pvalue'], fields['kbeta'],
                  fields['bse'], fields['frac_h2'], ml, ks, nks, notes, False, False)
        lines = {}
        for model, item in (('seer', seer), ('lmm', lmm), ('enet', seer)):
            for ld in (None, lineages):
                for ps in (False, True):
                    key = '%s|%s|%s' % (model, 'lin' if ld else 'nolin', 'samples' if ps else 'nosamples')
                    lines[key] = ru.format_output(item, ld, model=model, print_samples=ps)
        enc = dict(fields)
        for k in ('af', 'prep', 'pvalue', 'kbeta', 'bse', 'intercept', 'frac_h2'):
            enc[k] = _f(enc[k])
        enc['betas'] = [_f(b) for b in enc['betas']]
        enc['max_lineage'] = _f(ml) if ml is not None else None
        out.append({'fields': enc, 'lines': lines})
    return out


def hash_cases():
    rng = np.random.RandomState(5)
    out = []
    for n in (1, 7, 50, 333):
        k = (rng.uniform(size=n) < 0.4).astype(np.int64)
        out.append({'k': k.tolist(), 'dtype': 'int64', 'hash': ri.hash_pattern(k).decode()})
        kf = k.astype(np.float64)
        kf[rng.randint(0, n)] = np.nan
        out.append({'k': [None if np.isnan(x) else x for x in kf], 'dtype': 'float64',
                    'hash': ri.hash_pattern(kf).decode()})
    return out


class _Call(dict):
    pass


class _Record(object):
    """What read_vcf_var touches on a pysam VariantRecord."""

    def __init__(self, f, names):
        self.contig = f[0]
        self.pos = int(f[1])
        alts = None if f[4] == '.' else tuple(f[4].split(','))
        self.alts = alts
        self.alleles = (f[3],) + (alts or ())
        filt = [] if f[6] in ('.', '') else f[6].split(';')
        self.filter = {k: None for k in filt}
        fmt = f[8].split(':')
        gi = fmt.index('GT') if 'GT' in fmt else -1
        self.samples = {}
        for name, cell in zip(names, f[9:]):
            c = _Call()
            if gi >= 0:
                gt = cell.split(':')[gi].replace('|', '/').split('/')
                c['GT'] = tuple(None if h == '.' else int(h) for h in gt)
            self.samples[name] = c
        self.start = self.pos - 1
        self.stop = self.start + len(f[3])


class _Vcf(object):
    """Stand-in for pysam.VariantFile over an in-memory VCF: iteration and fetch()."""

    def __init__(self, path):
        with gzip.open(path, 'rt') as fh:
            lines = [l.rstrip('\n').split('\t') for l in fh if not l.startswith('##') and l.strip()]
        self.names = lines[0][9:]
        self.records = [_Record(f, self.names) for f in lines[1:]]

    def fetch(self, contig, start, stop):
        return [r for r in self.records if r.contig == contig and r.start < stop and r.stop > start]

    def __iter__(self):
        return iter(self.records)


def variant_stream(infile, p, var_type, burden=False, burden_regions=None, sample_order=None):
    """read_variant until eof, as iter_variants drives it (input.py:553-566)."""
    all_strains = set(p.index)
    rows = []
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        while True:
            eof, k, var_name, kstrains, nkstrains, af, missing = ri.read_variant(
                infile, p, var_type, burden, burden_regions, False, all_strains, sample_order or [])
            if eof:
                break
            if k is None:
                rows.append({'name': None})
                continue
            rows.append({'name': var_name, 'af': float(af), 'missing': float(missing),
                         'carriers': len(kstrains), 'hash': ri.hash_pattern(k).decode(),
                         'k': [None if (isinstance(x, float) and np.isnan(x)) else int(x) for x in k.tolist()]})
    return rows, err.getvalue()


def int_fixtures():
    """run_test.sh:52 ("sample names are all integers"): the reference's fixtures restricted to the 50
    phenotyped samples, as gen_golden.py does for the named ones."""
    import shutil
    t = os.path.join(REF, 'tests')
    shutil.copy(os.path.join(t, 'subset_int.pheno'), os.path.join(GOLDEN, 'subset_int.pheno'))
    shutil.copy(os.path.join(t, 'kmers_int.gz'), os.path.join(GOLDEN, 'kmers_int.gz'))
    for c in ('30.log', '30.err'):
        shutil.copy(os.path.join(t, 'baseline', c), os.path.join(GOLDEN, 'baseline', c))
    ph = pd.read_csv(os.path.join(t, 'subset_int.pheno'), sep='\t', index_col=0)
    keep = [str(x) for x in ph.index]
    d = pd.read_csv(os.path.join(t, 'distances_int.tsv.gz'), sep='\t', index_col=0)
    d.index = d.index.astype(str)
    d.columns = d.columns.astype(str)
    d.loc[keep, keep].to_csv(os.path.join(GOLDEN, 'distances50_int.tsv'), sep='\t')


def main():
    int_fixtures()
    out = {}
    out['format_output'] = format_output_cases()
    out['hash_pattern'] = hash_cases()

    pheno = os.path.join(GOLDEN, 'subset.pheno')
    p = ri.load_phenotypes(pheno, None)
    out['phenotypes'] = {'index': [str(x) for x in p.index], 'values': [float(x) for x in p.values]}

    with contextlib.redirect_stderr(io.StringIO()):
        m = ri.load_structure(os.path.join(GOLDEN, 'distances50.tsv'), p, 10, 'classic', 1, None)
    cov = ri.load_covariates(os.path.join(GOLDEN, 'covariates.txt'), ['2q', '3'], p)
    with contextlib.redirect_stderr(io.StringIO()):
        lin, lin_names = ri.load_lineage(os.path.join(GOLDEN, 'lineage_clusters50.txt'), p)
    out['covariates'] = {'columns': [str(c) for c in cov.columns], 'index': [str(i) for i in cov.index]}
    out['lineage'] = {'labels': [str(x) for x in lin_names]}

    # k-mer and Rtab fixtures through read_variant (files opened as open_variant_file does, :268-299)
    slim = lambda rows: [{k: v for k, v in r.items() if k != 'k'} for r in rows]
    with gzip.open(os.path.join(GOLDEN, 'kmers.gz'), 'r') as fh:
        rows, err = variant_stream(fh, p, 'kmers')
    out['kmers'] = {'rows': slim(rows), 'stderr': err}
    import tempfile
    with tempfile.NamedTemporaryFile('w', suffix='.Rtab', delete=False) as tf:
        with gzip.open(os.path.join(GOLDEN, 'presence_absence.Rtab.gz'), 'rt') as src:
            tf.write(src.read())
    with open(tf.name) as fh:
        sample_order = [str(x) for x in fh.readline().rstrip().split()[1:]]
        rows, err = variant_stream(fh, p, 'Rtab', sample_order=sample_order)
    os.unlink(tf.name)
    out['rtab'] = {'rows': slim(rows), 'stderr': err}

    # VCF records and burden regions through read_vcf_var / the burden branch
    from collections import deque
    vcf = _Vcf(os.path.join(GOLDEN, 'variants50.vcf.gz'))
    rows, err = variant_stream(iter(vcf.records), p, 'vcf')
    out['vcf'] = {'rows': rows, 'stderr': err}
    for tag, fn in (('burden', 'burden_regions.txt'), ('burden_multiple', 'burden_regions_multiple.txt')):
        regions = []
        ri.load_burden(os.path.join(GOLDEN, fn), regions)
        rows, err = variant_stream(vcf, p, 'vcf', True, deque(regions))
        out[tag] = {'rows': rows, 'stderr': err}

    with open(os.path.join(GOLDEN, 'host_goldens.json'), 'w') as fh:
        json.dump(out, fh, indent=0, sort_keys=True)

    rng = np.random.RandomState(2)
    arrs = {'structure_m': np.asarray(m.values if hasattr(m, 'values') else m, dtype=float),
            'covariates': np.asarray(cov.values, dtype=float), 'lineage': np.asarray(lin, dtype=float)}
    for i, n in enumerate((5, 12, 40)):
        x = rng.normal(size=(n, 3 + i))
        D = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1))
        Y, e = rc.cmdscale(D)
        arrs['cmdscale_D%d' % i] = D
        arrs['cmdscale_Y%d' % i] = Y
        arrs['cmdscale_e%d' % i] = e
    np.savez_compressed(os.path.join(GOLDEN, 'host_goldens.npz'), **arrs)
    print('wrote host_goldens.json / .npz:', {k: (len(v) if isinstance(v, list) else list(v)[:4])
                                               for k, v in out.items()})


if __name__ == '__main__':
    main()
