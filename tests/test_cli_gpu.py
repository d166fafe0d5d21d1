"""End-to-end CLI parity against the reference's own baseline logs (tests/baseline/*.log and
*.err of the reference, produced by the real pyseer + statsmodels; copied to
tests/golden/baseline/).  Inputs are the reference's fixtures restricted to the 50 phenotyped
samples (pyseer intersects samples before MDS / kinship normalisation, so results are the
same).  Values are printed at '%.2E': they must agree to one unit of the last digit; PC
coefficients are defined up to the sign of the MDS eigenvectors."""
import contextlib
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

G = lambda f: os.path.join(GOLDEN, f)
FIXED = ['--kmers', G('kmers.gz'), '--phenotypes', G('subset.pheno')]
CASES = {
    # run_test.sh:19
    '1': FIXED + ['--distances', G('distances50.tsv')],
    # run_test.sh:24-25 (pop_struct.pkl is the saved MDS of the same distances)
    '5': FIXED + ['--max-dimensions', '3', '--distances', G('distances50.tsv'),
                  '--phenotype-column', 'continuous'],
    '6': FIXED + ['--max-dimensions', '3', '--continuous', '--distances', G('distances50.tsv')],
    # run_test.sh:28
    '9': FIXED + ['--max-dimensions', '3', '--covariates', G('covariates.txt'), '--use-covariates',
                  '2q', '3', '--distances', G('distances50.tsv')],
    # run_test.sh:33
    '14': ['--pres', G('presence_absence.Rtab.gz'), '--phenotypes', G('subset.pheno'),
           '--distances', G('distances50.tsv'), '--max-dimensions', '3'],
    # run_test.sh:34
    '15': FIXED + ['--distances', G('distances50.tsv'), '--max-dimensions', '3', '--mds', 'classic',
                   '--continuous'],
    # run_test.sh:42
    '20': FIXED + ['--similarity', G('similarity50.tsv'), '--lmm'],
    # run_test.sh:46-47
    '24': ['--pres', G('presence_absence.Rtab.gz'), '--phenotypes', G('subset.pheno'), '--lmm',
           '--similarity', G('similarity50.tsv')],
    '25': FIXED + ['--lmm', '--similarity', G('similarity50.tsv'), '--covariates',
                   G('covariates.txt'), '--use-covariates', '2q', '3'],
    # run_test.sh:31-32, :59 (VCF, burden regions), :45 (LMM with VCF)
    '12': ['--vcf', G('variants50.vcf.gz'), '--phenotypes', G('subset.pheno'), '--distances',
           G('distances50.tsv'), '--max-dimensions', '3'],
    '13': ['--vcf', G('variants50.vcf.gz'), '--burden', G('burden_regions.txt'), '--phenotypes',
           G('subset.pheno'), '--distances', G('distances50.tsv'), '--max-dimensions', '3'],
    '37': ['--vcf', G('variants50.vcf.gz'), '--burden', G('burden_regions_multiple.txt'),
           '--phenotypes', G('subset.pheno'), '--distances', G('distances50.tsv'),
           '--max-dimensions', '3'],
    '23': ['--vcf', G('variants50.vcf.gz'), '--phenotypes', G('subset.pheno'), '--lmm',
           '--similarity', G('similarity50.tsv')],
    # run_test.sh:22 (filters), :49 (pattern output)
    '3': FIXED + ['--filter-pvalue', '1E-5', '--lrt-pvalue', '1E-8', '--distances',
                  G('distances50.tsv')],
    '27': FIXED + ['--lmm', '--similarity', G('similarity50.tsv')],
    # run_test.sh:40-41, :44 (lineage effects: MDS components, user clusters, LMM)
    '18': FIXED + ['--distances', G('distances50.tsv'), '--max-dimensions', '3', '--lineage'],
    '19': FIXED + ['--distances', G('distances50.tsv'), '--max-dimensions', '3', '--lineage',
                   '--lineage-clusters', G('lineage_clusters50.txt')],
    '22': FIXED + ['--lmm', '--similarity', G('similarity50.tsv'), '--lineage', '--distances',
                   G('distances50.tsv')],
    # run_test.sh:50-51
    '28': FIXED + ['--no-distances'],
    '29': FIXED + ['--no-distances', '--use-covariates', '3', '--covariates', G('covariates.txt')],
}


def _table(text):
    lines = [l for l in text.split('\n') if l]
    header = lines[0].split('\t')
    rows = {}
    for l in lines[1:]:
        f = l.split('\t')
        rows[f[0]] = dict(zip(header[1:], f[1:]))
    return header, rows


def _counters(err):
    out = {}
    for l in err.split('\n'):
        for key in ('loaded', 'pre-filtered', 'tested', 'printed'):
            if l.endswith(key + ' variants'):
                out[key] = int(l.split()[0])
    return out


def _same(a, b, abs_only=False):
    if a == b:
        return True
    if a == '' or b == '':
        return False
    x, y = float(a), float(b)
    if abs_only:
        x, y = abs(x), abs(y)
    if abs(x) < 1e-7 and abs(y) < 1e-7:            # zero up to convergence / rounding noise
        return True
    # one unit of the third significant digit
    return abs(x - y) <= 1.5e-2 * max(abs(y), 1e-300)


@pytest.mark.parametrize('case', sorted(CASES, key=int))
def test_baseline(case, tmp_path):
    from pyseer_b200.__main__ import main
    out, err = io.StringIO(), io.StringIO()
    args = list(CASES[case])
    if case == '27':
        args += ['--output-patterns', str(tmp_path / 'patterns.txt')]
    if case in ('18', '19', '22'):
        args += ['--lineage-file', str(tmp_path / 'lineage_effects.txt')]
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
        main(args)
    if case == '27':
        # one base64 MD5 line per tested variant (__main__.py:559-560; scripts/count_patterns.py)
        pats = open(str(tmp_path / 'patterns.txt'), 'rb').read().split(b'\n')[:-1]
        assert len(pats) == _counters(err.getvalue())['tested'] and all(len(x) == 24 for x in pats)
    ref_out = open(os.path.join(GOLDEN, 'baseline', case + '.log')).read()
    ref_err = open(os.path.join(GOLDEN, 'baseline', case + '.err')).read()
    assert _counters(err.getvalue()) == _counters(ref_err)
    h, rows = _table(out.getvalue())
    rh, rrows = _table(ref_out)
    assert h == rh
    assert list(rows) == list(rrows)                   # same variants, same order
    bad = []
    lineage_diff = 0
    for name, ref in rrows.items():
        got = rows[name]
        for col in rh[1:]:
            if col == 'notes':
                ok = set(got[col].split(',')) == set(ref[col].split(','))
            elif col == 'lineage':
                ok = got[col] == ref[col]
                if not ok and case == '19':
                    # 18 cluster columns on 50 samples: nearly every fit is quasi-separated and
                    # the arg-max of the Wald statistics is decided by rounding noise (the
                    # NumPy restatement itself differs from the reference's log on 26 of 188
                    # rows); counted, not required to be identical
                    lineage_diff += 1
                    ok = True
            else:
                ok = _same(got[col], ref[col], abs_only=col.startswith('PC'))
            if not ok:
                bad.append((name[:20], col, got[col], ref[col]))
    assert not bad, bad[:10]
    assert lineage_diff <= 0.25 * len(rrows)


def _write_kmers(path, n_samples, n_kmers, seed=3):
    """Synthetic --kmers text file (plain) + phenotype and similarity files next to it."""
    import benchdata
    from oracle import synth
    X, y, K = benchdata.lmm_problem(n_samples, seed=seed)
    ys = np.where(y > np.median(y), 1, -1).astype(np.int8)
    bits = synth.synth_rows(77, 0, n_kmers, n_samples, 0.0, 1.0, 50, ys, 0)
    x = synth.unpack_rows(bits, n_samples)
    names = np.array(['s%d' % i for i in range(n_samples)])
    with open(path, 'w') as fh:
        for s in range(n_kmers):
            fh.write('K%07d | ' % s + ' '.join(t + ':1' for t in names[x[s] != 0]) + '\n')
    base = os.path.dirname(path)
    with open(os.path.join(base, 'pheno.tsv'), 'w') as fh:
        fh.write('samples\tcontinuous\tbinary\n')
        for i in range(n_samples):
            fh.write('s%d\t%r\t%d\n' % (i, float(y[i]), int(y[i] > np.median(y))))
    import pandas as pd
    pd.DataFrame(K, index=names, columns=names).to_csv(os.path.join(base, 'sim.tsv'), sep='\t')
    pd.DataFrame(np.sqrt(np.maximum(0, np.add.outer(np.diag(K), np.diag(K)) - 2 * K)), index=names,
                 columns=names).to_csv(os.path.join(base, 'dist.tsv'), sep='\t')
    return base


def _cli(args):
    from pyseer_b200.__main__ import main
    out, err = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
        main(args)
    return out.getvalue(), err.getvalue()


@pytest.mark.parametrize('mode', ['lmm', 'fixed'])
def test_pipeline_batches_do_not_change_output(mode, tmp_path):
    """The streaming pipeline (reader thread, pinned staging, submit k+1 while k runs, output thread):
    many small batches in flight print the same bytes as one batch."""
    base = _write_kmers(str(tmp_path / 'kmers.txt'), 200, 5000)
    common = ['--kmers', str(tmp_path / 'kmers.txt'), '--uncompressed', '--phenotypes',
              os.path.join(base, 'pheno.tsv'), '--print-filtered', '--block_size', '100']
    if mode == 'lmm':
        common += ['--lmm', '--similarity', os.path.join(base, 'sim.tsv'), '--phenotype-column', 'continuous']
    else:
        common += ['--distances', os.path.join(base, 'dist.tsv'), '--max-dimensions', '4',
                   '--phenotype-column', 'binary']
    one, e1 = _cli(common + ['--gpu-batch', '100000'])
    many, e2 = _cli(common + ['--gpu-batch', '300', '--cpu', '4'])
    assert one == many and _counters(e1) == _counters(e2) and _counters(e1)['loaded'] == 5000


@pytest.mark.parametrize('mode', ['lmm', 'fixed'])
def test_two_gpus_print_the_same_table(mode, tmp_path):
    """--gpus 2: batches dealt to two GPUs, tables gathered over NCCL on the first one, output in input
    order -- byte-identical to the one-GPU run.  Needs two GPUs (`gpurun --gpus 2`)."""
    from pyseer_b200.engine import device_count
    if device_count() < 2:
        pytest.skip('needs 2 GPUs')
    base = _write_kmers(str(tmp_path / 'kmers.txt'), 200, 5000)
    common = ['--kmers', str(tmp_path / 'kmers.txt'), '--uncompressed', '--phenotypes',
              os.path.join(base, 'pheno.tsv'), '--print-filtered', '--block_size', '100', '--gpu-batch', '700']
    if mode == 'lmm':
        common += ['--lmm', '--similarity', os.path.join(base, 'sim.tsv'), '--phenotype-column', 'continuous']
    else:
        common += ['--distances', os.path.join(base, 'dist.tsv'), '--max-dimensions', '4',
                   '--phenotype-column', 'binary']
    one, e1 = _cli(common)
    two, e2 = _cli(common + ['--gpus', '2'])
    assert one == two and _counters(e1) == _counters(e2)


@pytest.mark.parametrize('case', ['18', '19', '22'])
def test_lineage_runs_fast_path_equals_result_loop(case, tmp_path, monkeypatch):
    """--lineage runs: k-mer text tokenised on the device (rows brought back for the LMM's per-block
    lineage fits) + psb_format_rows_lineage print what the host parser + the row-by-row result loop
    print -- same lines, same lineage labels, same counters, in small blocks and batches."""
    extra = ['--lineage-file', str(tmp_path / 'lineage.txt'), '--block_size', '30', '--gpu-batch', '60',
             '--print-filtered']
    outs = {}
    for tag, env in (('fast', {'PYSEER_B200_TEXT': '1', 'PYSEER_B200_NATIVE_FORMAT': '1'}),
                     ('loop', {'PYSEER_B200_TEXT': '0', 'PYSEER_B200_NATIVE_FORMAT': '0'})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        outs[tag] = _cli(list(CASES[case]) + extra)
    a, b = outs['fast'][0].split('\n'), outs['loop'][0].split('\n')
    assert len(a) == len(b) > 150
    for x, y in zip(a, b):
        fx_, fy = x.split('\t'), y.split('\t')
        assert fx_[:-1] == fy[:-1] and set(fx_[-1].split(',')) == set(fy[-1].split(',')), (x, y)
    assert _counters(outs['fast'][1]) == _counters(outs['loop'][1])
