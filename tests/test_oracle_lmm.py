"""Pin the LMM oracle to the reference's goldens (tests/lmm_test.py) and to vectors
computed by the reference's unmodified fastlmm.lmm_cov (tests/golden/lmm_ref_*.npz)."""
import numpy as np
import pytest

from oracle import lmm_oracle as lo
from conftest import load_golden


def _setup(lmmfix, with_cov=False):
    names = list(lmmfix['sim_subset_names'])
    idx = [names.index(s) for s in lmmfix['samples']]
    K = lmmfix['sim_subset'][np.ix_(idx, idx)]
    cov = None
    if with_cov:
        cn = list(lmmfix['cov_names'])
        ci = [cn.index(s) for s in lmmfix['samples']]
        # tests/lmm_test.py:82 reads covariates.txt unchanged: both columns numeric
        cov = np.c_[lmmfix['cov_quantitative'][ci],
                    lmmfix['cov_categorical'][ci].astype(float)]
    return lo.initialise_lmm(lmmfix['pheno_binary'], cov, K)


def test_initialise_lmm(goldens, lmmfix):
    lmm, h2, res = _setup(lmmfix)
    assert abs(res['nLL'][0] - goldens['lmm_nLL']) < 1e-6
    assert abs(h2 - goldens['lmm_h2']) < 1e-9
    lmm, h2, res = _setup(lmmfix, True)
    assert abs(res['nLL'][0] - goldens['lmm_cov_nLL']) < 1e-7
    assert abs(h2) < 1e-9


def test_fit_lmm_block(goldens, lmmfix, utd):
    lmm, h2, _ = _setup(lmmfix)
    k = utd['k'][:50]
    r = lo.fit_lmm_block(lmm, h2, k.reshape(-1, 1))
    g = goldens['lmm_fit']
    assert abs(r['beta'][0] - g['kbeta']) < 1e-9
    assert abs(r['bse'][0] - g['bse']) < 1e-9
    assert abs(r['frac_h2'][0] - g['frac_h2']) < 1e-9
    assert abs(r['p_values'][0] - g['pvalue']) < 1e-9
    with pytest.raises(KeyError):
        lo.fit_lmm_block(lmm, 1, k.reshape(-1, 1))
    with pytest.raises(AssertionError):
        lo.fit_lmm_block(lmm, h2, k.reshape(-1, 1)[:10])


def _var(pattern='pattern'):
    nan = np.nan
    return lo.LMM('variant', pattern, 0.2, nan, nan, nan, nan, nan, nan, [], [], set(),
                  True, True)


def test_fit_lmm(goldens, lmmfix, utd):
    lmm, h2, _ = _setup(lmmfix)
    p = lmmfix['pheno_binary']
    k = utd['k'][:50]
    g = goldens['lmm_fit']

    def run(kk, pattern='pattern', cont=False, fp=1, lp=1, lineage=False, lin=None, cov=None):
        return lo.fit_lmm(lmm, h2, [(_var(pattern), p, kk)], kk.reshape(-1, 1).copy(),
                          lineage, lin if lin is not None else [], cov if cov is not None else
                          np.empty((0, 0)), cont, fp, lp)[0]

    r = run(k)
    for f, v in g.items():
        assert abs(getattr(r, f) - v) < 1e-9, f
    assert r.notes == set() and not r.prefilter and not r.filter
    r = run(k, pattern=None)
    assert r.notes == {'af-filter'} and r.prefilter and not r.filter
    bad_k = np.array([1.] * 5 + [0.] * 45)
    r = run(bad_k)
    for f, v in goldens['lmm_fit_badchisq'].items():
        assert abs(getattr(r, f) - v) < 1e-9, f
    assert r.notes == {'bad-chisq'}
    r = run(k, fp=0.05)
    assert r.notes == {'pre-filtering-failed'} and r.prefilter and abs(r.prep - g['prep']) < 1e-9
    r = run(k, lp=0.05)
    assert r.notes == {'lrt-filtering-failed'} and r.filter and abs(r.pvalue - g['pvalue']) < 1e-9
    r = run(k, lineage=True, lin=utd['m'][:50])
    assert r.max_lineage == goldens['lmm_lineage_index']
    r = run(k, cont=True)
    assert abs(r.prep - goldens['lmm_fit_cont_prep']) < 1e-9
    assert abs(r.pvalue - g['pvalue']) < 1e-9


@pytest.mark.parametrize('tag', ['interior_cont', 'interior_cov', 'interior_binary'])
def test_against_reference_module(tag):
    d = load_golden('lmm_ref_%s.npz' % tag)
    cov = d['cov'] if d['cov'].shape[1] else None
    lmm, h2, res = lo.initialise_lmm(d['y'], cov, d['K'])
    assert abs(h2 - d['h2'][0]) < 1e-5
    h2 = float(d['h2'][0])
    assert abs(res['nLL'][0] - d['nLL'][0]) < 1e-7
    assert 0.05 < h2 < 0.95      # interior: exercises the 1/Sd weighting
    r = lo.fit_lmm_block(lmm, h2, d['snps'].astype(float))
    ok = np.isfinite(d['p_values'])
    assert np.array_equal(ok, np.isfinite(r['p_values']))
    assert np.allclose(r['beta'][ok], d['beta'][ok], rtol=1e-8, atol=1e-12)
    assert np.allclose(r['bse'][ok] ** 2, d['variance_beta'][ok], rtol=1e-8)
    assert np.allclose(r['p_values'][ok], d['p_values'][ok], rtol=1e-7)
    assert np.allclose(r['frac_h2'][ok] ** 2, d['frac'][ok], rtol=1e-8, atol=1e-14)
