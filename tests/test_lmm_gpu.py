"""GPU parity tests for the LMM path: CUDA (through the C ABI) vs the oracle and the
reference goldens.  Tolerance: 1e-6 relative on beta / bse / frac_h2 / p-values
(BASELINE.json north_star), counts and flags bit-exact."""
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-6
# contraction back-ends under test: 0 = FP64 CUDA cores, k = k exact int8 slices on tcgen05
PRECISIONS = [int(t) for t in os.environ.get('PSB_TEST_PRECISIONS', '0,4,5,6,7,46').split(',')]
PREC2 = [p for p in PRECISIONS if p in (0, 5, 46)] or PRECISIONS[:1]


def _close(a, b, rtol=RTOL):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    big = fin & (np.abs(b) > 1e-290)
    assert np.allclose(a[big], b[big], rtol=rtol, atol=0), np.nanmax(np.abs(a[big] / b[big] - 1))
    small = fin & ~big
    assert np.all(np.abs(a[small]) < 1e-280)


def _setup_subset(lmmfix, with_cov=False, precision=0):
    from pyseer_b200 import lmm as plmm
    names = list(lmmfix['sim_subset_names'])
    idx = [names.index(s) for s in lmmfix['samples']]
    K = lmmfix['sim_subset'][np.ix_(idx, idx)].copy()
    n = K.shape[0]
    K *= float(n) / np.diag(K).sum()
    if with_cov:
        cn = list(lmmfix['cov_names'])
        ci = [cn.index(s) for s in lmmfix['samples']]
        X = np.c_[lmmfix['cov_quantitative'][ci], lmmfix['cov_categorical'][ci].astype(float),
                  np.ones(n)]
    else:
        X = np.ones((n, 1))
    m = plmm.KinshipLMM(X, lmmfix['pheno_binary'].reshape(-1, 1), K, precision=precision)
    res = m.findH2()
    return m, res


@pytest.mark.parametrize('precision', PREC2)
def test_reference_goldens(goldens, lmmfix, utd, precision):
    """tests/lmm_test.py:395-420 and :136-392 replayed on the GPU path."""
    from pyseer_b200 import lmm as plmm
    from pyseer_b200.classes import LMM
    m, res = _setup_subset(lmmfix, precision=precision)
    assert abs(res['nLL'][0] - goldens['lmm_nLL']) < 1e-6 and abs(res['h2']) < 1e-9
    h2 = res['h2']
    k = utd['k'][:50]
    r = plmm.fit_lmm_block(m, h2, k.reshape(-1, 1))
    g = goldens['lmm_fit']
    assert abs(r['beta'][0] - g['kbeta']) < 1e-7
    assert abs(r['bse'][0] - g['bse']) < 1e-7
    assert abs(r['frac_h2'][0] - g['frac_h2']) < 1e-7
    assert abs(r['p_values'][0] - g['pvalue']) < 1e-7
    with pytest.raises(KeyError):
        plmm.fit_lmm_block(m, 1, k.reshape(-1, 1))
    with pytest.raises(AssertionError):
        plmm.fit_lmm_block(m, h2, k.reshape(-1, 1)[:10])

    p = lmmfix['pheno_binary']
    nan = np.nan

    def run(kk, pattern='pattern', cont=False, fp=1, lp=1):
        var = LMM('variant', pattern, 0.2, nan, nan, nan, nan, nan, nan, ['a'], ['b'], set(),
                  True, True)
        return plmm.fit_lmm(m, h2, [(var, p, kk)], kk.reshape(-1, 1).copy(), False, [],
                            np.empty((0, 0)), cont, fp, lp)[0]

    r = run(k)
    for f, v in g.items():
        assert abs(getattr(r, f) - v) < 1e-7, f
    assert r.notes == set() and not r.prefilter and not r.filter and r.max_lineage is None
    r = run(k, pattern=None)
    assert r.notes == {'af-filter'} and r.prefilter and not r.filter and np.isnan(r.prep)
    bad_k = np.array([1.] * 5 + [0.] * 45)
    r = run(bad_k)
    for f, v in goldens['lmm_fit_badchisq'].items():
        assert abs(getattr(r, f) - v) < 1e-7, f
    assert r.notes == {'bad-chisq'}
    r = run(k, fp=0.05)
    assert r.notes == {'pre-filtering-failed'} and r.prefilter and not r.filter
    assert abs(r.prep - g['prep']) < 1e-7 and np.isnan(r.pvalue)
    r = run(k, lp=0.05)
    assert r.notes == {'lrt-filtering-failed'} and r.filter and not r.prefilter
    assert abs(r.pvalue - g['pvalue']) < 1e-7 and np.isnan(r.kbeta)
    r = run(k, cont=True)
    assert abs(r.prep - goldens['lmm_fit_cont_prep']) < 1e-7
    assert abs(r.pvalue - g['pvalue']) < 1e-7
    m.close()
    # covariates (tests/lmm_test.py:80-88)
    m, res = _setup_subset(lmmfix, with_cov=True, precision=precision)
    assert abs(res['nLL'][0] - goldens['lmm_cov_nLL']) < 1e-6
    m.close()


@pytest.mark.parametrize('precision', PREC2)
@pytest.mark.parametrize('tag', ['interior_cont', 'interior_cov', 'interior_binary'])
def test_against_reference_module_vectors(tag, precision):
    """Vectors computed by the reference's unmodified fastlmm.lmm_cov (interior h2)."""
    from pyseer_b200 import lmm as plmm
    d = load_golden('lmm_ref_%s.npz' % tag)
    n = d['y'].shape[0]
    X = np.c_[d['cov'], np.ones(n)] if d['cov'].shape[1] else np.ones((n, 1))
    m = plmm.KinshipLMM(X, d['y'].reshape(-1, 1), d['K'].copy(), precision=precision)
    res = m.findH2()
    # the h2 search returns the best *evaluated* point of a flat minimum: its last digits
    # depend on the host BLAS; the statistics below are compared at the golden h2
    assert abs(res['h2'] - d['h2'][0]) < 1e-5
    assert abs(res['nLL'][0] - d['nLL'][0]) < 1e-7
    r = plmm.fit_lmm_block(m, float(d['h2'][0]), d['snps'].astype(float))
    ok = np.isfinite(d['p_values'])
    _close(r['beta'][ok], d['beta'][ok])
    _close(r['bse'][ok] ** 2, d['variance_beta'][ok])
    _close(r['p_values'][ok], d['p_values'][ok])
    big = ok & (d['frac'] > 1e-12)
    _close(r['frac_h2'][big] ** 2, d['frac'][big])
    # constant columns (2, 3): rotate() zeroes them -> beta 0, p 1 (lmm_cov.py:180-181, 803-805)
    assert r['p_values'][2] == 1.0 and r['p_values'][3] == 1.0
    m.close()


def _synthetic(n, seed):
    rng = np.random.RandomState(seed)
    G = (rng.uniform(size=(n, 2 * n)) < rng.uniform(0.05, 0.95, 2 * n)).astype(float)
    K = G.dot(G.T)
    g = G.dot(rng.normal(size=2 * n))
    y = (g - g.mean()) / g.std() * np.sqrt(0.5) + np.sqrt(0.5) * rng.normal(size=n)
    return K, y


@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('n,nv,binary', [(50, 200, True), (333, 1500, False), (1000, 3000, True)])
def test_oracle_parity_full_path(n, nv, binary, precision):
    """fit_lmm semantics (AF filter, pre-filter, LMM fit, lrt filter) against the oracle on
    seeded synthetic k-mers, including planted low-p variants and edge rows."""
    from oracle import lmm_oracle as lo
    from pyseer_b200 import lmm as plmm, _lib
    from pyseer_b200.engine import synth_host, unpack_rows
    K, y = _synthetic(n, 100 + n)
    if binary:
        y = (y > np.median(y)).astype(float)
    ys = np.where(y > np.median(y), 1, -1).astype(np.int8)
    bits = synth_host(20261017, 0, nv, n, af_lo=0.0, af_hi=1.0, planted_every=50, y_sign=ys)
    bits[5] = 0                       # empty row
    bits[6] = bits[7]                 # duplicate pattern
    x = unpack_rows(bits, n)
    continuous = not binary
    min_af, max_af, fp, lp = 0.02, 0.98, 0.5, 0.3

    olmm, oh2, _ = lo.initialise_lmm(y, None, K.copy())
    Kn = K * (float(n) / np.diag(K).sum())
    m = plmm.KinshipLMM(np.ones((n, 1)), y.reshape(-1, 1), Kn.copy(), precision=precision)
    h2 = m.findH2()['h2']
    assert abs(h2 - oh2) < 1e-5
    h2 = oh2
    r = plmm.run_lmm_bits(m, h2, bits, None, continuous, fp, lp, min_af=min_af, max_af=max_af,
                          max_missing=0.05)

    # oracle, variant by variant through fit_lmm (lmm.py:125-226)
    nan = np.nan
    variants, cols = [], []
    af = x.sum(1) / float(n)
    for s in range(nv):
        ok = (af[s] >= min_af) and (af[s] <= max_af)
        var = lo.LMM('v%d' % s, 'pat' if ok else None, af[s], nan, nan, nan, nan, nan, nan, [], [],
                     set(), True, True)
        variants.append((var, y, x[s].astype(float)))
    mat = x.T.astype(float).copy()
    mat[:, ~((af >= min_af) & (af <= max_af))] = 0
    out = lo.fit_lmm(olmm, oh2, variants, mat, False, [], np.empty((0, 0)), continuous, fp, lp)
    byname = {o.kmer: o for o in out}
    n_pref = n_tested = 0
    for s in range(nv):
        o = byname['v%d' % s]
        f = int(r.flags[s])
        from pyseer_b200.engine import notes_from_flags
        assert notes_from_flags(f) == o.notes, (s, notes_from_flags(f), o.notes)
        assert bool(f & _lib.F_PREFILTER) == o.prefilter and bool(f & _lib.F_FILTER) == o.filter
        assert r.af[s] == o.af
        n_pref += o.prefilter
        n_tested += (not o.prefilter)
    assert r.counts['loaded'] == nv and r.counts['prefiltered'] == n_pref
    assert r.counts['tested'] == n_tested
    for fld, col in (('prep', r.prep), ('pvalue', r.pvalue), ('kbeta', r.beta), ('bse', r.bse),
                     ('frac_h2', r.extra)):
        ref = np.array([getattr(byname['v%d' % s], fld) for s in range(nv)], dtype=float)
        _close(col, ref)
    if n >= 333:
        assert np.nanmin(r.pvalue) < 1e-8      # the planted tail is exercised
    m.close()


@pytest.mark.parametrize('precision', PREC2)
def test_missing_genotypes(precision):
    """NaN genotypes: excluded from the 2x2 table, counted as carriers in af, NaN statistics
    and 'lrt-filtering-failed' (lmm.py:201; input.py:439-452)."""
    from oracle import lmm_oracle as lo
    from pyseer_b200 import lmm as plmm, _lib
    from pyseer_b200.engine import pack_rows
    n = 120
    K, y = _synthetic(n, 5)
    y = (y > np.median(y)).astype(float)
    rng = np.random.RandomState(3)
    k = (rng.uniform(size=(20, n)) < 0.4).astype(float)
    k[3, 10] = np.nan
    k[4, :3] = np.nan
    bits, miss = pack_rows(k)
    Kn = K * (float(n) / np.diag(K).sum())
    m = plmm.KinshipLMM(np.ones((n, 1)), y.reshape(-1, 1), Kn.copy(), precision=precision)
    h2 = m.findH2()['h2']
    r = plmm.run_lmm_bits(m, h2, bits, miss, False, 1, 1, 0.01, 0.99, 0.05)
    for s in (3, 4):
        kk = k[s]
        prep, bad = lo.pre_filtering(y, kk, False)
        assert abs(r.prep[s] / prep - 1) < 1e-9
        assert r.missing[s] == np.isnan(kk).sum()
        assert r.carriers[s] == np.nansum(kk) + np.isnan(kk).sum()
        assert np.isnan(r.pvalue[s]) and (r.flags[s] & _lib.F_LRT_FAILED)
    assert np.isfinite(r.pvalue[[0, 1, 2, 5]]).all()
    m.close()


def test_large_n_mode_matches_default(monkeypatch):
    """The tensor kernel's large-N mode (packed rows read from global memory instead of a
    shared-memory tile, used when N is too large for both the tile and the operand ring) gives
    the same table as the default mode."""
    from pyseer_b200 import lmm as plmm
    from pyseer_b200.engine import synth_host
    n, nv = 700, 2000
    K, y = _synthetic(n, 3)
    Kn = K * (float(n) / np.diag(K).sum())
    bits = synth_host(5, 0, nv, n)
    out = []
    for force in (False, True):
        if force:
            monkeypatch.setenv('PSB_TC_GLOBAL_BITS', '1')
        m = plmm.KinshipLMM(np.ones((n, 1)), y.reshape(-1, 1), Kn.copy(), precision=5)
        h2 = m.findH2()['h2']
        out.append(plmm.run_lmm_bits(m, h2, bits, None, True, 1.0, 1.0, 0.01, 0.99, 0.05))
        m.close()
    for f in ('pvalue', 'beta', 'bse', 'extra'):
        assert np.array_equal(getattr(out[0], f), getattr(out[1], f), equal_nan=True), f


@pytest.mark.parametrize('n,nv', [(300, 700), (1100, 300)])
def test_tensor_kernel_modes_agree(n, nv, monkeypatch):
    """The launch modes of the tensor kernel -- single CTAs, multicast pairs, two-SM MMAs
    (PSB_TC_PAIR = 0 / 1 / 2), rows in shared memory or read from global memory -- run the same
    integer contraction: their results are bit-identical."""
    from pyseer_b200 import lmm as plmm
    from pyseer_b200.engine import synth_host, unpack_rows
    rng = np.random.RandomState(n)
    G = (rng.uniform(size=(n, 2 * n)) < rng.uniform(0.05, 0.95, 2 * n)).astype(float)
    K = G.dot(G.T)
    K *= n / np.diag(K).sum()
    y = G.dot(rng.normal(size=2 * n)) + rng.normal(size=n) * np.sqrt(n)
    snps = unpack_rows(synth_host(5, 0, nv, n), n).T.astype(float)
    m = plmm.KinshipLMM(np.ones((n, 1)), y.reshape(-1, 1), K, precision=5)
    m.findH2()
    ref = None
    for env in ({'PSB_TC_PAIR': '2'}, {'PSB_TC_PAIR': '1'}, {'PSB_TC_PAIR': '0'},
                {'PSB_TC_PAIR': '2', 'PSB_TC_GLOBAL_BITS': '1'}, {'PSB_TC_PAIR': '2', 'PSB_TC_STAGES': '2'}):
        for k in ('PSB_TC_PAIR', 'PSB_TC_GLOBAL_BITS', 'PSB_TC_STAGES'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        r = plmm.fit_lmm_block(m, 0.4, snps)
        if ref is None:
            ref = r
        else:
            for key in ('p_values', 'beta', 'bse', 'frac_h2'):
                assert np.array_equal(r[key], ref[key], equal_nan=True), (env, key)
    m.close()
