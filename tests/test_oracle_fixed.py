"""Pin the fixed-effects oracle to the reference's own goldens
(/root/reference/tests/model_test.py; constants in tests/golden/reference_goldens.json)."""
import numpy as np

from oracle import fixed_oracle as fo


def test_prefilter(goldens, utd):
    prep, bad = fo.pre_filtering(utd['p_binary'], utd['k'], False)
    assert abs(prep - goldens['prefilter_binary_p']) < 1e-12 and not bad
    p = np.concatenate((np.ones(50), np.zeros(50)))
    k = np.concatenate((np.ones(45), np.zeros(55)))
    prep, bad = fo.pre_filtering(p, k, False)
    assert abs(prep / goldens['prefilter_binary_bad_p'] - 1) < 1e-10 and bad
    prep, bad = fo.pre_filtering(np.random.RandomState(0).random_sample(100), utd['k'], False)
    assert np.isnan(prep) and bad
    prep, bad = fo.pre_filtering(utd['p_continuous'], utd['k'], True)
    assert abs(prep - goldens['prefilter_cont_p']) < 1e-12 and not bad
    prep, bad = fo.pre_filtering(p, k, True)
    assert abs(prep / goldens['prefilter_cont_binaryp_p'] - 1) < 1e-10 and not bad


def test_fit_null(goldens, utd):
    none = np.empty((0, 0))
    r = fo.fit_null(utd['p_binary'], utd['m'], none, False)
    assert np.abs(r.params - goldens['null_binary_params']).max() < 1e-7
    assert abs(fo.fit_null(utd['p_binary'], utd['m'], none, False, True)
               - goldens['null_binary_firth']) < 1e-9
    r = fo.fit_null(utd['p_binary'], utd['m'], utd['cov'], False)
    assert np.abs(r.params - goldens['null_binary_cov_params']).max() < 1e-7
    assert abs(fo.fit_null(utd['p_binary'], utd['m'], utd['cov'], False, True)
               - goldens['null_binary_cov_firth']) < 1e-9
    p = np.array([1.] * 10 + [0.] * 90)
    assert fo.fit_null(p, p.reshape(-1, 1).copy(), none, False) is None
    r = fo.fit_null(utd['p_continuous'], utd['m'], none, True)
    assert np.abs(r.params - goldens['null_cont_params']).max() < 1e-7
    r = fo.fit_null(utd['p_continuous'], utd['m'], utd['cov'], True)
    assert np.abs(r.params - goldens['null_cont_cov_params']).max() < 1e-7


def test_lineage(goldens, utd):
    none = np.empty((0, 0))
    assert fo.fit_lineage_effect(utd['m'], none, utd['k']) == goldens['lineage_index']
    assert fo.fit_lineage_effect(utd['lin'], none, utd['k']) == goldens['lineage_index']
    assert fo.fit_lineage_effect(utd['m'], utd['cov'], utd['k']) == goldens['lineage_index']
    k = np.array([1.] * 10 + [0.] * 90)
    assert fo.fit_lineage_effect(k.reshape(-1, 1).copy(), none, k) is None


def test_firth(goldens, utd):
    p, m = utd['p_binary'], utd['m']
    assert abs(fo.firth_likelihood(utd['firth_vars'], m, p) - goldens['firth_likelihood']) < 1e-9
    assert fo.firth_likelihood(utd['firth_vars'] + 100, m, p) == np.inf
    start = np.zeros(m.shape[1])
    start[0] = np.log(np.mean(p) / (1 - np.mean(p)))
    intercept, kbeta, beta, bse, fitll = fo.fit_firth(start, m, p)
    assert abs(intercept - goldens['firth_intercept']) < 1e-9
    assert abs(kbeta - goldens['firth_kbeta']) < 1e-9
    assert np.abs(np.array(beta) - goldens['firth_beta']).max() < 1e-7
    assert abs(bse - goldens['firth_bse']) < 1e-9
    assert abs(fitll - goldens['firth_ll']) < 1e-9
    assert fo.fit_firth(start, m, p, step_limit=10, convergence_limit=1e-10) is None


def _check(s, g, notes, prefilter, filt):
    for f in ['prep', 'pvalue', 'kbeta', 'bse', 'intercept']:
        assert abs(getattr(s, f) - g[f]) < 1e-7, f
    assert np.abs(np.asarray(s.betas) - g['betas']).max() < 1e-7
    assert s.notes == notes and s.prefilter == prefilter and s.filter == filt


def test_fixed_effects_binary(goldens, utd):
    p, k, m = utd['p_binary'], utd['k'], utd['m']
    none = np.empty((0, 0))
    args = dict(variant='variant', af=0.2, pattern='test', null_res=-9.9, null_firth=-9.9,
                kstrains=[], nkstrains=[], continuous=False)
    s = fo.fixed_effects_regression(p=p, k=k, m=m, c=none, lineage_effects=False, lin=None,
                                    pret=1, lrtt=1, **args)
    _check(s, goldens['fe_binary'], set(), False, False)
    s = fo.fixed_effects_regression(p=None, k=k, m=m, c=none, lineage_effects=False, lin=None,
                                    pret=1, lrtt=1, **args)
    assert s.notes == {'af-filter'} and s.prefilter and not s.filter and np.isnan(s.prep)
    s = fo.fixed_effects_regression(p=p, k=k, m=m, c=none, lineage_effects=False, lin=None,
                                    pret=0.05, lrtt=1, **args)
    assert s.notes == {'pre-filtering-failed'} and s.prefilter and np.isnan(s.pvalue)
    s = fo.fixed_effects_regression(p=p, k=k, m=m, c=none, lineage_effects=False, lin=None,
                                    pret=1, lrtt=0.05, **args)
    _check(s, goldens['fe_binary'], {'lrt-filtering-failed'}, False, True)
    pb = np.array([1.] * 10 + [0.] * 90)
    s = fo.fixed_effects_regression(p=pb, k=pb.copy(), m=pb.reshape(-1, 1).copy(), c=none,
                                    lineage_effects=False, lin=None, pret=1, lrtt=1, **args)
    assert s.notes == {'bad-chisq'}
    s = fo.fixed_effects_regression(p=p, k=k, m=m, c=utd['cov'], lineage_effects=False,
                                    lin=None, pret=1, lrtt=1, **args)
    _check(s, goldens['fe_binary_cov'], set(), False, False)
    s = fo.fixed_effects_regression(p=p, k=k, m=m, c=none, lineage_effects=True,
                                    lin=utd['lin'], pret=1, lrtt=1, **args)
    _check(s, goldens['fe_binary'], set(), False, False)
    assert s.max_lineage == 2


def test_fixed_effects_continuous(goldens, utd):
    p, k, m = utd['p_continuous'], utd['k'], utd['m']
    none = np.empty((0, 0))
    args = dict(variant='variant', af=0.2, pattern='test', null_res=None, null_firth=-9.9,
                kstrains=[], nkstrains=[], continuous=True, lineage_effects=False, lin=None)
    s = fo.fixed_effects_regression(p=p, k=k, m=m, c=none, pret=1, lrtt=1, **args)
    _check(s, goldens['fe_cont'], set(), False, False)
    s = fo.fixed_effects_regression(p=p, k=k, m=m, c=none, pret=0.05, lrtt=1, **args)
    assert s.notes == {'pre-filtering-failed'} and abs(s.prep - goldens['fe_cont']['prep']) < 1e-9
    s = fo.fixed_effects_regression(p=p, k=k, m=m, c=none, pret=1, lrtt=1e-50, **args)
    _check(s, goldens['fe_cont'], {'lrt-filtering-failed'}, False, True)
    s = fo.fixed_effects_regression(p=p, k=k, m=m, c=utd['cov'], pret=1, lrtt=1, **args)
    _check(s, goldens['fe_cont_cov'], set(), False, False)


def test_null_fit_powell_fallback_oracle():
    """model.py:132-137 restated: Newton's 'Singular matrix' on a duplicated column -> statsmodels'
    Powell optimiser; the likelihood reaches the value of the fit without the duplicate."""
    import numpy as np
    from oracle import fixed_oracle as fo
    rng = np.random.RandomState(0)
    n = 200
    m = rng.uniform(-1, 1, size=(n, 3))
    y = (m[:, 0] + rng.normal(size=n) > 0).astype(float)
    none = np.empty((0, 0))
    a = fo.fit_null(y, m, none, False)
    b = fo.fit_null(y, np.c_[m, m[:, 1]], none, False)
    assert b is not None and abs(a.llf - b.llf) < 1e-4 and np.all(np.isnan(b.bse))
