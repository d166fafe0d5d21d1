"""psb_eigh (cuSOLVER syevd on the device) against numpy.linalg.eigh, the routine it replaces in
LMM.setSU_fromK (fastlmm/lmm_cov.py:88-103), and through KinshipLMM: same h2 and statistics."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _kernel(n, seed=3):
    rng = np.random.RandomState(seed)
    G = (rng.uniform(size=(n, 2 * n)) < rng.uniform(0.05, 0.95, 2 * n)).astype(float)
    K = G.dot(G.T)
    return K * (n / np.diag(K).sum()), G, rng


def test_eigh_matches_numpy():
    from pyseer_b200.engine import Engine
    n = 300
    K, _, rng = _kernel(n)
    K.flat[::n + 1] += 1.0
    # make the upper triangle garbage: like numpy, only the lower triangle may be read
    A = np.tril(K) + np.triu(rng.normal(size=(n, n)), 1)
    with Engine(0) as eng:
        w, V = eng.eigh(A)
    w2, _ = np.linalg.eigh(K)
    assert np.all(np.diff(w) >= 0)
    assert np.max(np.abs(w - w2) / np.abs(w2)) < 1e-12
    assert np.abs(V.T.dot(V) - np.eye(n)).max() < 1e-12
    assert np.abs(K.dot(V) - V * w).max() < 1e-11 * np.abs(w).max()


def test_lmm_statistics_do_not_depend_on_the_eigensolver(monkeypatch):
    from pyseer_b200 import lmm as plmm
    from pyseer_b200.engine import synth_host, unpack_rows
    n = 300
    K, G, rng = _kernel(n)
    g = G.dot(rng.normal(size=2 * n))
    y = (g - g.mean()) / g.std() * np.sqrt(0.5) + np.sqrt(0.5) * rng.normal(size=n)
    snps = unpack_rows(synth_host(11, 0, 256, n), n).T.astype(float)
    out = {}
    for mode in ('numpy', 'device'):
        monkeypatch.setenv('PYSEER_B200_EIGH', mode)
        m = plmm.KinshipLMM(np.ones((n, 1)), y.reshape(-1, 1), K.copy(), precision=0)
        h2 = m.findH2()['h2']
        out[mode] = (h2, plmm.fit_lmm_block(m, 0.5, snps))
        m.close()
    assert abs(out['numpy'][0] - out['device'][0]) < 1e-6
    for key in ('p_values', 'beta', 'bse', 'frac_h2'):
        assert np.allclose(out['numpy'][1][key], out['device'][1][key], rtol=1e-9, atol=0), key


def test_nll_terms_match_numpy():
    """psb_lmm_nll_terms: the two O(N) sums of LMM.nLLeval (lmm_cov.py:597-684) for a grid of h2"""
    from pyseer_b200.engine import Engine
    rng = np.random.RandomState(5)
    for J in (7, 999, 4999):
        S = np.sort(rng.gamma(0.7, 2.0, size=J))
        S[:3] = [0.0, 1e-12, 1e-7]                    # null directions of a clonal kinship
        uy2 = rng.normal(size=J) ** 2
        h2 = np.concatenate([[0.0, 0.99999], rng.uniform(size=9)])
        with Engine(0) as eng:
            yky, ld = eng.nll_terms(S, uy2, h2)
        Sd = h2[:, None] * S[None, :] + (1.0 - h2[:, None])
        assert np.allclose(yky, (uy2[None, :] / Sd).sum(1), rtol=1e-13, atol=0)
        assert np.allclose(ld, np.log(Sd).sum(1), rtol=1e-12, atol=1e-10)
