"""Oracle parity at the BASELINE sizes (BASELINE.json configs[1..4]) on >= 1e5 sampled variants per
config, planted tail included, plus adversarial inputs for the triangular int8 form of the LMM
kernel (clonal kinship, h2 in {0, 0.5, 0.99}, D = 4, allele frequencies at the filter thresholds).

The CUDA path runs through the C ABI on seeded synthetic rows (psb_synth_device); the oracle's
answer for the same variant ids comes from `python -m oracle.cpu_arm sample` (NumPy twin of the
generator + oracle/lmm_oracle.py / oracle/fixed_oracle.py on forked workers, a process of its own
so that no fork happens under a live CUDA context).

Bar (BASELINE.json north_star): carriers, af, notes / filter flags and counters exact; prep,
lrt-pvalue, beta, bse, variant_h2 / intercept / slopes within 1e-6 relative.
PSB_TEST_SAMPLE scales the number of sampled variants (default 102000 per config).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

RTOL = 1e-6
SEED = 20261017
SAMPLE = int(os.environ.get('PSB_TEST_SAMPLE', '102000'))
CHUNK = 3000
# contraction back-end of the LMM cases: the library default unless PSB_TEST_BASELINE_PRECISION names one
PRECISION = int(os.environ.get('PSB_TEST_BASELINE_PRECISION', os.environ.get('PYSEER_B200_LMM_PRECISION', '46')))


def _tasks(n_total, id_space, chunk=CHUNK):
    """Chunks of `chunk` consecutive variant ids at offsets spread over the id space of the config
    (multiples of 1000, so every chunk starts on a planted variant)."""
    k = max(1, n_total // chunk)
    rng = np.random.RandomState(k)
    starts = np.sort(rng.choice(max(1, id_space // 1000 - chunk // 1000), size=k, replace=False)) * 1000
    return np.array([[int(s), chunk] for s in starts], dtype=np.int64)


def _oracle(tmp_path, tag, **state):
    st = tmp_path / ('state_%s.npz' % tag)
    out = tmp_path / ('oracle_%s.npz' % tag)
    np.savez(st, **state)
    subprocess.check_call([sys.executable, '-m', 'oracle.cpu_arm', 'sample', str(st), str(out)],
                          cwd=ROOT)
    os.unlink(st)
    with np.load(out) as d:
        return {k: d[k] for k in d.files}


def _compare(cols, ref, names, what, firth_noise=0):
    """cols: dict of GPU columns; ref: oracle.  Exact: carriers, af, flags.  1e-6: the rest.

    firth_noise: how many Firth fits may differ in 'firth-fail'.  model.fit_firth halves its step
    while firth_likelihood(new) > firth_likelihood(old) (model.py:470-476); one iterate before
    convergence the two values differ by rounding noise (1e-13 relative), and when the noise says
    "worse" the halving can park one ulp away from the old iterate for all 1000 tries -> None ->
    'firth-fail'.  Which variants that happens to is decided by the last bits of LAPACK's det / pinv
    in the reference (the device kernel treats a difference within 16 ulp as "not worse", so it no
    longer loses fits -- or 85 ms per lost fit -- to its own noise; what remains are the reference's
    unlucky rows); the rows are excluded from the value columns and counted."""
    vis = np.uint32(0x07FF)                       # note bits + prefilter + filter
    assert np.array_equal(cols['carriers'], ref['carriers']), what
    assert np.array_equal(cols['af'], ref['res'][:, 0]), what
    diff = (cols['flags'] & vis) ^ (ref['flags'] & vis)
    bad = np.where(diff != 0)[0]
    keep = np.ones(diff.shape[0], dtype=bool)
    if firth_noise:
        went_firth = (ref['flags'] & np.uint32(0x0004 | 0x0008 | 0x0010 | 0x0020)) != 0
        noise = bad[went_firth[bad] & ((diff[bad] & ~np.uint32(0x0040 | 0x0400 | 0x0100)) == 0)]
        assert noise.size <= firth_noise, (what, noise.size, firth_noise)
        keep[noise] = False
        bad = np.setdiff1d(bad, noise)
    assert bad.size == 0, (what, bad[:10], cols['flags'][bad[:10]], ref['flags'][bad[:10]])
    worst = {'firth_noise_rows': int((~keep).sum())}
    for j, name in names:
        # NaN wherever the reference leaves the tuple field unset (filtered variants, failed fits)
        a, b = cols[name][keep], ref['res'][keep, j]
        fin = np.isfinite(b)
        mism = np.where(np.isfinite(a) != fin)[0]
        assert mism.size == 0, (what, name, np.where(keep)[0][mism[:10]], a[mism[:10]], b[mism[:10]])
        big = fin & (np.abs(b) > 1e-290)
        # Signed coefficients pass through zero: among 1e5 variants x 12 coefficients some are 1e-6
        # of their standard error by chance, and a relative error against such a value measures
        # nothing.  beta is held to 1e-6 of max(|beta|, bse / 1000) (a t statistic of 0.001), the
        # intercept and the slopes of the fixed-effects model to 1e-6 of max(|b|, 1e-4) (1e-3 of
        # their typical standard errors), variant_h2 to max(|.|, 1e-9); p-values, bse, prep: plain.
        # Firth fits are only defined to the size of their last step: the iteration stops at a step
        # norm of 1e-4 (model.py:480-483), the last steps are ~1e-7, and whether one of them is halved
        # is again decided by a comparison of two penalised likelihoods that agree to rounding noise
        # (see firth_noise above) -- measured: coefficients of such a fit differ from the oracle's by
        # half a last step, 8e-8.  They are held to 1e-6 of max(|b|, 0.1), bse to 1e-6 relative.
        firth = (cols['flags'][keep][big] & np.uint32(0x1000)) != 0
        if name == 'beta':
            floor = np.maximum(1e-3 * np.nan_to_num(ref['res'][keep, 4][big], nan=0.0, posinf=0.0),
                               np.where(firth, 0.1, 0.0))
        elif name.startswith('b_') or (name == 'extra' and what.startswith('fixed')):
            floor = np.where(firth, 0.1, 1e-4)
        else:
            floor = 1e-9 if name == 'extra' else 0.0
        err = np.abs(a[big] - b[big]) / np.maximum(np.abs(b[big]), floor)
        worst[name] = float(err.max()) if err.size else 0.0
        assert worst[name] < RTOL, (what, name, worst[name], np.where(keep)[0][np.where(big)[0][int(np.argmax(err))]])
        assert np.all(np.abs(a[fin & ~big]) < 1e-280), (what, name)
    return worst


LMM_COLS = [(1, 'prep'), (2, 'pvalue'), (3, 'beta'), (4, 'bse'), (5, 'extra')]


def _run_lmm_tasks(m, h2, n, tasks, ys, continuous, af, planted, thresholds, seed=SEED):
    eng = m.engine(h2)
    cols = {k: [] for k in ('carriers', 'af', 'flags', 'prep', 'pvalue', 'beta', 'bse', 'extra')}
    tested = 0
    for first, count in tasks:
        eng.synth_device(seed, int(first), int(count), af[0], af[1], planted, ys)
        eng.run_lmm(continuous=continuous, **thresholds)
        r = eng.fetch()
        tested += r.counts['tested']
        assert r.counts['loaded'] == count
        for k in cols:
            cols[k].append(getattr(r, k))
    return {k: np.concatenate(v) for k, v in cols.items()}, tested


def _lmm_case(tmp_path, tag, n, X, y, K, tasks, continuous, h2_forced=None, af=(0.02, 0.98),
              planted=1000, min_af=0.01, max_af=0.99, filter_pvalue=1.0, lrt_pvalue=1.0,
              precision=PRECISION, min_tail=None):
    from pyseer_b200 import lmm as plmm
    m = plmm.KinshipLMM(X, y.reshape(-1, 1), K.copy(), precision=precision)
    h2 = float(m.findH2()['h2'])
    S, U = m.getSU()
    out = {}
    ys = np.where(y > np.median(y), 1, -1).astype(np.int8)
    thresholds = dict(min_af=min_af, max_af=max_af, max_missing=0.05, filter_pvalue=filter_pvalue,
                      lrt_pvalue=lrt_pvalue)
    for h in ([h2] if h2_forced is None else h2_forced):
        cols, tested = _run_lmm_tasks(m, h, n, tasks, ys, continuous, af, planted, thresholds)
        ref = _oracle(tmp_path, '%s_%g' % (tag, h), model='lmm', n=n, task_kind='synth', tasks=tasks,
                      X=X, y=y, U=U, S=S, h2=h, continuous=continuous, af_lo=af[0], af_hi=af[1],
                      planted=planted, separated=0, seed=SEED, min_af=min_af, max_af=max_af,
                      filter_pvalue=filter_pvalue, lrt_pvalue=lrt_pvalue)
        assert tested == int(ref['tested'][0])
        worst = _compare(cols, ref, LMM_COLS, 'lmm %s h2=%g' % (tag, h))
        if min_tail is not None:
            assert np.nanmin(cols['pvalue']) < min_tail, np.nanmin(cols['pvalue'])
        out[h] = (worst, tested, float(np.nanmin(cols['pvalue'])))
        print('lmm %s h2=%g precision=%d: %d tested, worst rel err %s' % (tag, h, precision, tested, worst))
    m.close()
    return h2, out


def test_config3_lmm_n5000_continuous(tmp_path):
    """BASELINE configs[3]: LMM, continuous phenotype, N=5000, D=1; ids sampled over the 50M k-mers."""
    from oracle import cpu_arm
    n = 5000
    X, y, K = cpu_arm.lmm_problem(n)
    tasks = _tasks(SAMPLE, 50000000)
    h2, out = _lmm_case(tmp_path, 'c3', n, X, y, K, tasks, True, min_tail=1e-50)
    assert 0.05 < h2 < 0.95                          # interior h2: the 1/Sd weighting is exercised
    worst, tested, pmin = out[h2]
    assert tested >= 0.95 * tasks[:, 1].sum()
    print('configs[3] N=5000: %d tested, worst rel err %s, min p %.3g' % (tested, worst, pmin))


def test_config1_lmm_n1000_binary(tmp_path):
    """BASELINE configs[1]: LMM, binary phenotype (chi-square pre-filter), N=1000."""
    from oracle import cpu_arm
    n = 1000
    X, y, K = cpu_arm.lmm_problem(n)
    yb = (y > np.median(y)).astype(float)
    tasks = _tasks(SAMPLE, 1000000)
    h2, out = _lmm_case(tmp_path, 'c1', n, X, yb, K, tasks, False, min_tail=1e-20)
    # and with gating filters, so that pre-filtered / lrt-filtered rows are in the comparison
    _lmm_case(tmp_path, 'c1f', n, X, yb, K, tasks[:4], False, af=(0.0, 1.0), filter_pvalue=0.5,
              lrt_pvalue=0.3)


@pytest.mark.parametrize('case', ['clonal', 'cov4', 'af_edges'])
def test_lmm_n5000_adversarial(tmp_path, case):
    """Inputs chosen against the triangular int8 form a = x'M''x (cancellation between ~N^2/4 signed
    terms quantised per 32-column tile): near-low-rank kinship at h2 -> 0.99, four covariates, and
    allele frequencies on the AF-filter thresholds."""
    from oracle import cpu_arm
    n = 5000
    tasks = _tasks(2 * CHUNK, 50000000)
    if case == 'clonal':
        X, y, K = cpu_arm.lmm_problem(n, seed=SEED + 1, clonal=40)
        _lmm_case(tmp_path, case, n, X, y, K, tasks, True, h2_forced=[0.0, 0.5, 0.99])
    elif case == 'cov4':
        X, y, K = cpu_arm.lmm_problem(n, seed=SEED + 2, n_cov=3)
        y = y + X[:, 0] * 0.5
        _lmm_case(tmp_path, case, n, X, y, K, tasks, True)
    else:
        X, y, K = cpu_arm.lmm_problem(n, seed=SEED + 3)
        _lmm_case(tmp_path, 'aflo', n, X, y, K, tasks[:1], True, af=(0.006, 0.014), planted=0)
        _lmm_case(tmp_path, 'afhi', n, X, y, K, tasks[:1], True, af=(0.986, 0.994), planted=0)


def test_config2_fixed_n2000_logit_firth(tmp_path):
    """BASELINE configs[2]: fixed effects, logistic + Firth, N=2000, 10 MDS covariates (p = 12)."""
    from oracle import cpu_arm, fixed_oracle as fo
    from pyseer_b200 import model as pm, _lib
    n, dims = 2000, 10
    mds, y = cpu_arm.fixed_problem(n, dims)
    none = np.empty((0, 0))
    onull = fo.fit_null(y, mds, none, False)
    ofirth = fo.fit_null(y, mds, none, False, True)
    gnull = pm.fit_null(y, mds, none, False)
    assert abs(gnull.llf / onull.llf - 1) < 1e-9
    assert abs(pm.fit_null(y, mds, none, False, True) / ofirth - 1) < 1e-8
    tasks = _tasks(SAMPLE, 10000000)
    ys = np.where(y > 0.5, 1, -1).astype(np.int8)
    model = pm.FixedModel(y, mds, none, False, onull.llf, float(ofirth))
    eng = model.engine
    cols = {k: [] for k in ('carriers', 'af', 'flags', 'prep', 'pvalue', 'beta', 'bse', 'extra', 'betas')}
    tested = n_firth = 0
    for first, count in tasks:
        eng.synth_device(SEED, int(first), int(count), 0.02, 0.98, 1000, ys, 100)
        eng.run_fixed(min_af=0.01, max_af=0.99, max_missing=0.05, filter_pvalue=1.0, lrt_pvalue=1.0,
                      continuous=False)
        r = eng.fetch()
        tested += r.counts['tested']
        n_firth += eng.last_stats()['firth_fits']
        for k in cols:
            cols[k].append(getattr(r, k))
    cols = {k: np.concatenate(v) for k, v in cols.items()}
    for j in range(dims):
        cols['b_%d' % j] = cols['betas'][:, j]
    ref = _oracle(tmp_path, 'c2', model='fixed', n=n, task_kind='synth', tasks=tasks, y=y, m=mds,
                  null_llf=onull.llf, null_firth=float(ofirth), continuous=False, af_lo=0.02,
                  af_hi=0.98, planted=1000, separated=100, seed=SEED, min_af=0.01, max_af=0.99,
                  filter_pvalue=1.0, lrt_pvalue=1.0)
    assert tested == int(ref['tested'][0])
    names = [(1, 'prep'), (2, 'pvalue'), (3, 'beta'), (4, 'bse'), (5, 'extra')] + \
            [(6 + j, 'b_%d' % j) for j in range(dims)]
    used = int(((cols['flags'] & np.uint32(_lib.F_FIRTH_USED)) != 0).sum())
    worst = _compare(cols, ref, names, 'fixed configs[2]', firth_noise=max(3, used // 100))
    assert n_firth > 0 and used == n_firth, (n_firth, used)     # configs[2] is "logistic + Firth"
    assert ((ref['flags'] & 0x0004) != 0).sum() > 0
    print('configs[2] N=2000 p=12: %d tested, %d Firth fits, worst rel err %s' % (tested, used, worst))
    model.close()


def test_config4_burden_n10000(tmp_path):
    """BASELINE configs[4]: VCF burden test, N=10000 samples, LMM; every region the union of 1-20
    rare record rows formed on the device."""
    from oracle import cpu_arm
    from pyseer_b200 import lmm as plmm
    n = 10000
    n_regions = int(os.environ.get('PSB_TEST_BURDEN_REGIONS', str(min(SAMPLE, 100000))))
    n_regions = n_regions // CHUNK * CHUNK
    X, y, K = cpu_arm.lmm_problem(n)
    m = plmm.KinshipLMM(X, y.reshape(-1, 1), K.copy(), precision=PRECISION)
    del K
    h2 = float(m.findH2()['h2'])
    S, U = m.getSU()
    offs, mem = cpu_arm.burden_regions(n_regions)
    eng = m.engine(h2)
    n_rec = int(offs[-1])
    eng.synth_device(SEED + 5, 0, n_rec, 0.001, 0.02, 0, None)
    rec_ptr, _, _, rec_w = eng.submitted_device()
    eng.submit_burden_device(rec_ptr, n_rec, rec_w, offs, mem)
    eng.run_lmm(min_af=0.01, max_af=0.99, max_missing=0.05, filter_pvalue=1.0, lrt_pvalue=1.0,
                continuous=True)
    r = eng.fetch()
    cols = {k: getattr(r, k) for k in ('carriers', 'af', 'flags', 'prep', 'pvalue', 'beta', 'bse', 'extra')}
    tasks = np.array([[r0, CHUNK] for r0 in range(0, n_regions, CHUNK)], dtype=np.int64)
    ref = _oracle(tmp_path, 'c4', model='lmm', n=n, task_kind='burden', tasks=tasks, X=X, y=y, U=U, S=S,
                  h2=h2, continuous=True, af_lo=0.001, af_hi=0.02, planted=0, separated=0,
                  seed=SEED + 5, min_af=0.01, max_af=0.99, filter_pvalue=1.0, lrt_pvalue=1.0,
                  offs=offs, mem=mem, rec_first=0)
    assert r.counts['tested'] == int(ref['tested'][0]) and r.counts['loaded'] == n_regions
    worst = _compare(cols, ref, LMM_COLS, 'lmm burden configs[4]')
    print('configs[4] N=10000: %d regions, %d tested, worst rel err %s' % (n_regions, r.counts['tested'], worst))
    m.close()
