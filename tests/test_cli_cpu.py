"""The CLI's host side end to end WITHOUT a GPU: readers (native k-mer / Rtab / VCF parsers, burden
regions), batching, the native formatter and pattern hashes, counters -- with the device engine
replaced by test doubles that compute every batch through the oracle (oracle/ is test
infrastructure; the product path has no such route).  Outputs are compared with the reference's own
baseline logs exactly as tests/test_cli_gpu.py does on the B200."""
import contextlib
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_cli_gpu import CASES, _counters, _same, _table      # same cases, same comparison

pytestmark = []          # (test_cli_gpu's module-level gpu mark does not apply here)


class _FakeEngine(object):
    """The engine calls the CLI's pipeline makes (submit / run_* / fetch, pipeline.py), answered by
    the oracle; burden unions through the oracle's packed-row rule instead of psb_submit_burden."""
    lmm_precision = 5

    def __init__(self, fixed_model=None, lmm=None, h2=None):
        self.fixed_model, self.lmm, self.h2 = fixed_model, lmm, h2
        self._sub = self._res = None

    def submit_burden(self, bits, missing, offsets, members):
        from oracle.input_oracle import burden_union
        self._rows = burden_union(bits, missing, offsets, members)

    def download_rows(self):
        bits, miss = self._rows if getattr(self, '_rows', None) is not None else self._sub
        return bits, (miss if miss is not None and miss.any() else None)

    def submit(self, bits, missing=None):
        self._sub = (np.array(bits), None if missing is None else np.array(missing))
        self._info = None

    # the device text parser's contract (psb_submit_text), restated in Python: samples after the first
    # '|', blank-separated tokens, name up to ':', unknown samples ignored
    def text_setup(self, samples):
        self._index = {}
        for i, x in enumerate(samples):
            self._index.setdefault(str(x), i)
        self._n = len(samples)

    def submit_text(self, text, n_bytes, line_start, line_len, n_lines):
        from pyseer_b200.engine import words_per_row
        bits = np.zeros((n_lines, words_per_row(self._n)), dtype=np.uint32)
        info = np.zeros(n_lines, dtype=np.int32)
        for v in range(n_lines):
            line = bytes(text[line_start[v]:line_start[v] + line_len[v]]).decode()
            if '|' not in line:
                info[v] = 4 | 2
                continue
            for tok in line.split('|', 1)[1].replace('\t', ' ').split(' '):
                i = self._index.get(tok.split(':')[0]) if tok else None
                if i is not None:
                    bits[v, i >> 5] |= np.uint32(1 << (i & 31))
            if not bits[v].any():
                info[v] |= 2
        self._sub = (bits, None)
        self._info = info

    def text_info(self, n_lines):
        return self._cur_info[:n_lines]

    def pattern_digests(self):
        # psb_pattern_digests' contract: md5 of the vector k the reference hashes (input.py:710-723)
        import hashlib
        bits, miss = self._sub
        return np.array([np.frombuffer(hashlib.md5(_k_of(bits, miss, j, self._n).tobytes()).digest(), dtype=np.uint8)
                         for j in range(bits.shape[0])]).reshape(-1, 16)

    def run_fixed(self, min_af, max_af, max_missing, filter_pvalue, lrt_pvalue, continuous):
        self._cur_info = self._info
        self._res = _fake_run_fixed_bits(self.fixed_model, self._sub[0], self._sub[1], filter_pvalue,
                                         lrt_pvalue, min_af, max_af, max_missing)

    def run_lmm(self, min_af, max_af, max_missing, filter_pvalue, lrt_pvalue, continuous):
        self._cur_info = self._info
        self._res = _fake_run_lmm_bits(self.lmm, self.h2, self._sub[0], self._sub[1], continuous,
                                       filter_pvalue, lrt_pvalue, min_af, max_af, max_missing)

    def fetch(self):
        return self._res

    def close(self):
        pass


class _FakePool(object):
    """pipeline.PinnedPool without cudaHostAlloc (no GPU in the CPU suite)."""

    def __init__(self, n, rows, W, with_missing):
        self.shape, self.with_missing = (rows, W), with_missing

    def get(self):
        return (np.empty(self.shape, dtype=np.uint32),
                np.empty(self.shape, dtype=np.uint32) if self.with_missing else None, 0)

    def put(self, token):
        pass

    def close(self):
        pass


class _FakeTextPool(object):
    """pipeline.TextPool without cudaHostAlloc."""

    def __init__(self, n, rows, text_bytes):
        self.rows, self.text_bytes = rows, text_bytes

    def get(self):
        return (np.empty(self.text_bytes, dtype=np.uint8), np.empty(self.rows, dtype=np.int64),
                np.empty(self.rows, dtype=np.int32), 0)

    def put(self, token):
        pass

    def close(self):
        pass


class _FakeAsyncFetch(object):
    """pipeline.AsyncFetch for the engine double: the 'copy' is taken when it is begun (the double's runs
    are synchronous), in the place the pipeline begins it -- before the next run is queued."""

    def __init__(self, eng, rows, n_betas):
        self.eng = eng

    def begin(self, n):
        self.r = self.eng.fetch()

    def wait(self):
        return self.r

    def close(self):
        pass


def _patch(monkeypatch):
    from pyseer_b200 import model as fx, lmm as lm, pipeline
    monkeypatch.setattr(fx, 'fit_null', _fake_fit_null)
    monkeypatch.setattr(fx, 'FixedModel', _FakeFixedModel)
    monkeypatch.setattr(lm.KinshipLMM, 'close', lambda self: None)
    monkeypatch.setattr(lm.KinshipLMM, 'engine', lambda self, h2: _FakeEngine(lmm=self, h2=h2))
    monkeypatch.setattr(pipeline, 'PinnedPool', _FakePool)
    monkeypatch.setattr(pipeline, 'TextPool', _FakeTextPool)
    monkeypatch.setattr(pipeline, 'AsyncFetch', _FakeAsyncFetch)


class _FakeFixedModel(object):
    def __init__(self, p, m, cov, continuous, null_res, null_firth, device=0, lineage=None):
        self.p = np.asarray(p, dtype=float).reshape(-1)
        self.m = np.asarray(m)
        self.cov = np.asarray(getattr(cov, 'values', cov))
        self.continuous = bool(continuous)
        self.null_llf = getattr(null_res, 'llf', null_res)
        self.null_firth = null_firth
        self.Z = np.ones((self.p.shape[0], 1 + (self.m.shape[1] if self.m.ndim == 2 and self.m.shape[0] == self.p.shape[0] else 0)
                          + (self.cov.shape[1] if self.cov.ndim == 2 and self.cov.shape[0] == self.p.shape[0] else 0)))
        self.engine = _FakeEngine(fixed_model=self)

    def close(self):
        pass


def _k_of(bits, missing, j, n):
    from pyseer_b200.engine import unpack_rows
    x = unpack_rows(bits[j:j + 1], n)[0].astype(float)
    if missing is not None:
        mm = unpack_rows(missing[j:j + 1], n)[0].astype(bool)
        if mm.any():
            x[mm] = np.nan
            return x
    return x.astype(np.int64)


def _results(n, nb):
    from pyseer_b200.engine import Results
    r = Results()
    for f in ('af', 'prep', 'pvalue', 'beta', 'bse', 'extra'):
        setattr(r, f, np.full(n, np.nan))
    r.betas = np.full((n, nb), np.nan)
    r.flags = np.zeros(n, dtype=np.uint32)
    r.carriers = np.zeros(n, dtype=np.int32)
    r.missing = np.zeros(n, dtype=np.int32)
    r.lineage = None
    return r


def _flags(notes, prefilter, filt):
    from pyseer_b200 import _lib
    bit = {s: b for b, s in _lib.NOTE_BITS}
    f = 0
    for s in notes:
        f |= bit[s]
    if prefilter:
        f |= _lib.F_PREFILTER
    else:
        f |= _lib.F_TESTED
    if filt:
        f |= _lib.F_FILTER
    return f


def _fake_run_fixed_bits(model, bits, missing, filter_pvalue, lrt_pvalue, min_af=-1.0, max_af=2.0,
                         max_missing=2.0, lineage=False):
    """model.fixed_effects_regression per variant through the oracle, iter_variants' AF filter first
    (input.py:608)."""
    from oracle import fixed_oracle as fo
    n = model.p.shape[0]
    nb = model.m.shape[1] if model.m.ndim == 2 and model.m.shape[0] == n else 0
    nb += model.cov.shape[1] if model.cov.ndim == 2 and model.cov.shape[0] == n else 0
    S = bits.shape[0]
    r = _results(S, nb)
    for j in range(S):
        k = _k_of(bits, missing, j, n)
        nan_mask = np.isnan(k) if k.dtype.kind == 'f' else np.zeros(n, dtype=bool)
        carriers = int(np.nansum(k == 1) + nan_mask.sum())          # missing count as carriers (:439)
        af = carriers / float(n)
        miss = nan_mask.sum() / float(n)
        keep = (min_af <= af <= max_af) and not (miss > max_missing)
        o = fo.fixed_effects_regression('v', model.p if keep else None, k.astype(float), model.m, model.cov, af,
                                        'p', False, None, filter_pvalue, lrt_pvalue, model.null_llf,
                                        model.null_firth, [], [], model.continuous)
        r.af[j], r.prep[j], r.pvalue[j], r.beta[j], r.bse[j] = af, o.prep, o.pvalue, o.kbeta, o.bse
        r.extra[j] = o.intercept
        betas = np.atleast_1d(np.asarray(o.betas, dtype=float)) if nb else np.empty(0)
        if nb and betas.shape[0] == nb:
            r.betas[j] = betas
        r.flags[j] = _flags(set(o.notes), o.prefilter, o.filter)
    return r


def _fake_run_lmm_bits(lmm, h2, bits, missing, continuous, filter_pvalue, lrt_pvalue, min_af=-1.0,
                       max_af=2.0, max_missing=2.0):
    """lmm.fit_lmm over the batch through the oracle (load_var_block's AF rule, input.py:693)."""
    from oracle import lmm_oracle as lo
    n = lmm.Y.shape[0]
    S = bits.shape[0]
    olmm = lo.OracleLMM(lmm.X, lmm.Y, None)
    olmm.S, olmm.U = lmm.getSU()
    r = _results(S, 0)
    y = lmm.Y[:, 0]
    nan = float('nan')
    variants, mat = [], np.zeros((n, S))
    for j in range(S):
        k = _k_of(bits, missing, j, n)
        nan_mask = np.isnan(k) if k.dtype.kind == 'f' else np.zeros(n, dtype=bool)
        af = (int(np.nansum(k == 1)) + int(nan_mask.sum())) / float(n)
        miss = nan_mask.sum() / float(n)
        ok = not (af < min_af or af > max_af or miss > max_missing)
        var = lo.LMM(str(j), 'p' if ok else None, af, nan, nan, nan, nan, nan, nan, [], [], set(), True, True)
        variants.append((var, y, k.astype(float) if ok else None))
        if ok:
            mat[:, j] = k.astype(float)
        r.af[j] = af
    out = lo.fit_lmm(olmm, h2, variants, mat, False, [], np.empty((0, 0)), continuous, filter_pvalue,
                     lrt_pvalue)
    for o in out:
        j = int(o.kmer)
        r.prep[j], r.pvalue[j], r.beta[j], r.bse[j], r.extra[j] = o.prep, o.pvalue, o.kbeta, o.bse, o.frac_h2
        r.flags[j] = _flags(o.notes, o.prefilter, o.filter)
    return r


def _fake_fit_null(p, m, cov, continuous, firth=False, device=0):
    from oracle import fixed_oracle as fo
    return fo.fit_null(np.asarray(p, dtype=float), np.asarray(m), np.asarray(getattr(cov, 'values', cov)),
                       continuous, firth)


CPU_CASES = ['1', '3', '5', '6', '9', '12', '13', '14', '15', '20', '23', '24', '25', '27', '28', '29', '37']


@pytest.mark.parametrize('case', CPU_CASES)
def test_baseline_with_oracle_engine(case, tmp_path, monkeypatch):
    from pyseer_b200.__main__ import main
    _patch(monkeypatch)
    out, err = io.StringIO(), io.StringIO()
    args = list(CASES[case]) + ['--cpu', '3']
    if case == '27':
        args += ['--output-patterns', str(tmp_path / 'patterns.txt')]
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err), np.errstate(all='ignore'):
        main(args)
    ref_out = open(os.path.join(GOLDEN, 'baseline', case + '.log')).read()
    ref_err = open(os.path.join(GOLDEN, 'baseline', case + '.err')).read()
    assert _counters(err.getvalue()) == _counters(ref_err)
    if case == '27':
        # one hash per tested variant, equal to the reference's hash_pattern of the same vectors
        import json
        pats = open(str(tmp_path / 'patterns.txt'), 'rb').read().split(b'\n')[:-1]
        assert len(pats) == _counters(err.getvalue())['tested'] and all(len(x) == 24 for x in pats)
        rows = json.load(open(os.path.join(GOLDEN, 'host_goldens.json')))['kmers']['rows']
        want = set(r['hash'].strip() for r in rows)
        assert set(x.decode() for x in pats) <= want
    h, rows = _table(out.getvalue())
    rh, rrows = _table(ref_out)
    assert h == rh
    assert list(rows) == list(rrows)                   # same variants, same order
    bad = []
    for name, ref in rrows.items():
        got = rows[name]
        for col in rh[1:]:
            if col == 'notes':
                ok = set(got[col].split(',')) == set(ref[col].split(','))
            else:
                ok = _same(got[col], ref[col], abs_only=col.startswith('PC'))
            if not ok:
                bad.append((name[:20], col, got[col], ref[col]))
    assert not bad, bad[:10]


def test_bits_cache_and_formatter_paths_agree(tmp_path, monkeypatch):
    """The same run with and without --bits-cache (first run writes it, second reads it), and with the
    row-by-row Python formatter instead of the native one, prints the same bytes."""
    from pyseer_b200.__main__ import main
    _patch(monkeypatch)
    cache = str(tmp_path / 'kmers.bits')

    def run(extra, env=None):
        out, err = io.StringIO(), io.StringIO()
        if env:
            monkeypatch.setenv(*env)
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err), np.errstate(all='ignore'):
            main(list(CASES['3']) + ['--print-filtered'] + extra)
        if env:
            monkeypatch.delenv(env[0])
        return out.getvalue(), _counters(err.getvalue()), err.getvalue()

    base, counts, _ = run([])
    first, c1, e1 = run(['--bits-cache', cache])
    second, c2, e2 = run(['--bits-cache', cache])
    assert 'Reading packed variants from' in e2 and 'Reading packed variants from' not in e1
    assert base == first == second and counts == c1 == c2
    slow, c3, _ = run([], env=('PYSEER_B200_NATIVE_FORMAT', '0'))
    assert c3 == counts
    a, b = base.split('\n'), slow.split('\n')
    assert len(a) == len(b)
    for x, y in zip(a, b):
        fx_, fy = x.split('\t'), y.split('\t')
        assert fx_[:-1] == fy[:-1] and set(fx_[-1].split(',')) == set(fy[-1].split(','))


def _run_cli(args, monkeypatch):
    from pyseer_b200.__main__ import main
    _patch(monkeypatch)
    out, err = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err), np.errstate(all='ignore'):
        main(args)
    return out.getvalue(), err.getvalue()


def _compare(case, out, err):
    ref_out = open(os.path.join(GOLDEN, 'baseline', case + '.log')).read()
    ref_err = open(os.path.join(GOLDEN, 'baseline', case + '.err')).read()
    assert _counters(err) == _counters(ref_err)
    h, rows = _table(out)
    rh, rrows = _table(ref_out)
    assert h == rh and list(rows) == list(rrows)
    bad = []
    for name, ref in rrows.items():
        for col in rh[1:]:
            got = rows[name][col]
            if col == 'notes':
                ok = set(got.split(',')) == set(ref[col].split(','))
            elif col in ('k-samples', 'nk-samples'):
                ok = got == ref[col]
            else:
                ok = _same(got, ref[col], abs_only=col.startswith('PC'))
            if not ok:
                bad.append((name[:20], col, got[:40], ref[col][:40]))
    assert not bad, bad[:10]


def test_more_reference_invocations(tmp_path, monkeypatch):
    """run_test.sh:21, 26-27, 29-30, 42-43, 48: --save-m / --load-m, --print-samples (the row-by-row
    formatter with sample lists), uncompressed k-mers, --cpu 2, --save-lmm / --load-lmm."""
    import gzip
    G = lambda f: os.path.join(GOLDEN, f)
    fixed = ['--kmers', G('kmers.gz'), '--phenotypes', G('subset.pheno')]
    # 1 + 2 + 8: save the projection, load it again
    out, err = _run_cli(fixed + ['--distances', G('distances50.tsv'), '--save-m', str(tmp_path / 'pop')],
                        monkeypatch)
    _compare('1', out, err)
    pkl = str(tmp_path / 'pop.pkl')
    for case in ('2', '8'):
        out, err = _run_cli(fixed + ['--load-m', pkl], monkeypatch)
        _compare(case, out, err)
    # 7: sample lists
    out, err = _run_cli(fixed + ['--max-dimensions', '3', '--print-samples', '--load-m', pkl], monkeypatch)
    _compare('7', out, err)
    # 10: uncompressed text
    txt = str(tmp_path / 'kmers.txt')
    with gzip.open(G('kmers.gz'), 'rb') as src, open(txt, 'wb') as dst:
        dst.write(src.read())
    out, err = _run_cli(['--kmers', txt, '--phenotypes', G('subset.pheno'), '--uncompressed', '--load-m', pkl],
                        monkeypatch)
    _compare('10', out, err)
    # 11: --cpu 2
    out, err = _run_cli(fixed + ['--cpu', '2', '--load-m', pkl], monkeypatch)
    _compare('11', out, err)
    # 20 + 21 + 26: LMM cache written, then loaded
    cache = str(tmp_path / 'lmm.cache')
    out, err = _run_cli(fixed + ['--similarity', G('similarity50.tsv'), '--lmm', '--save-lmm', cache], monkeypatch)
    _compare('20', out, err)
    out, err = _run_cli(fixed + ['--lmm', '--load-lmm', cache + '.npz'], monkeypatch)
    _compare('21', out, err)
    out, err = _run_cli(fixed + ['--lmm', '--load-lmm', cache + '.npz', '--cpu', '2'], monkeypatch)
    _compare('26', out, err)


def test_integer_sample_names(monkeypatch):
    """run_test.sh:52: sample names that are all integers stay strings through every loader."""
    G = lambda f: os.path.join(GOLDEN, f)
    out, err = _run_cli(['--kmers', G('kmers_int.gz'), '--phenotypes', G('subset_int.pheno'), '--distances',
                         G('distances50_int.tsv')], monkeypatch)
    _compare('30', out, err)
