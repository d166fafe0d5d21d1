"""The oracle's restatement of the LMM path against the UNMODIFIED reference modules copied to
oracle/_ref/ (oracle/build_ref.py; present wherever /root/reference was mounted at build time),
and the worker harness both share (oracle/cpu_arm.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def _small_state(n=120, n_cov=0, clonal=0):
    from oracle import cpu_arm
    X, y, K = cpu_arm.lmm_problem(n, seed=11, clonal=clonal, n_cov=n_cov)
    U, S, h2 = cpu_arm.lmm_spectral(X, y, K)
    return dict(X=X, y=y, U=U, S=S, h2=h2), K


needs_ref = pytest.mark.skipif(
    not os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'pyseer', 'fastlmm', 'lmm_cov.py')),
    reason='oracle/_ref not built (needs /root/reference at build time)')


@needs_ref
@pytest.mark.parametrize('n_cov,clonal,continuous', [(0, 0, True), (2, 0, False), (0, 6, True)])
def test_port_equals_reference_fit_lmm(n_cov, clonal, continuous):
    """fit_lmm of oracle/lmm_oracle.py == pyseer.lmm.fit_lmm on the same blocks: notes, filters and
    every statistic."""
    from oracle import cpu_arm
    n = 120
    st, _ = _small_state(n, n_cov, clonal)
    if not continuous:
        st['y'] = (st['y'] > np.median(st['y'])).astype(float)
    tasks = [('synth', 0, 400), ('synth', 10 ** 6, 300)]
    outs = {}
    for prefer in (True, False):
        arm = cpu_arm.CpuArm('lmm', n, st, 1, continuous, af=(0.0, 1.0), planted=50, min_af=0.02,
                             max_af=0.98, filter_pvalue=0.7, lrt_pvalue=0.6, collect=True,
                             prefer_reference=prefer)
        assert arm.kind == ('reference' if prefer else 'port')
        outs[prefer], _ = arm.run(tasks)
        arm.close()
    for a, b in zip(outs[True], outs[False]):
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
        assert np.array_equal(np.isnan(a[2]), np.isnan(b[2]))
        assert np.allclose(a[2], b[2], rtol=1e-9, atol=0, equal_nan=True)
    assert sum(o[0] for o in outs[True]) > 100


@needs_ref
def test_reference_findh2_equals_port():
    """initialise_lmm's spectral state and h2: oracle port vs the reference's LMM class."""
    from oracle import cpu_arm, ref_loader
    rcov = ref_loader.load('fastlmm.lmm_cov')
    st, K = _small_state(90)
    m = rcov.LMM(X=st['X'], Y=st['y'].reshape(-1, 1), G=None, K=K.copy(), inplace=True)
    res = m.findH2()
    assert abs(res['h2'] - st['h2']) < 1e-6
    assert np.allclose(m.S, st['S'], rtol=1e-9, atol=1e-12)


@needs_ref
def test_reference_parser_leg_runs():
    from oracle import cpu_arm
    st, _ = _small_state(64)
    r = cpu_arm.reference_parser_leg(st, 64, 150)
    assert r['kmers'] == 150 and 0 < r['tested'] <= 150 and r['parse_s'] > 0


def test_sample_cli_two_workers(tmp_path):
    """`python -m oracle.cpu_arm sample`: the oracle answers the parity tests at the BASELINE sizes
    compare with, computed by forked workers in a process of its own."""
    from oracle import cpu_arm, fixed_oracle as fo
    n = 150
    m, y = cpu_arm.fixed_problem(n, 3)
    none = np.empty((0, 0))
    null = fo.fit_null(y, m, none, False)
    firth = fo.fit_null(y, m, none, False, True)
    state = tmp_path / 'state.npz'
    np.savez(state, model='fixed', n=n, task_kind='synth', tasks=np.array([[0, 40], [1000, 30]]),
             y=y, m=m, null_llf=null.llf, null_firth=firth, continuous=False, af_lo=0.0, af_hi=1.0,
             planted=7, separated=10, seed=5, min_af=0.02, max_af=0.98, filter_pvalue=1.0,
             lrt_pvalue=1.0)
    out = tmp_path / 'out.npz'
    subprocess.check_call([sys.executable, '-m', 'oracle.cpu_arm', 'sample', str(state), str(out),
                           '--cores', '2'], cwd=ROOT)
    with np.load(out) as d:
        assert d['res'].shape == (70, 6 + 3) and d['flags'].shape == (70,)
        assert int(d['tested'][0]) == int(((d['flags'] & 0x0200) == 0).sum())
        assert (d['flags'] & 0x0004).any()                    # separated rows -> bad-chisq
