"""Size-independent properties of the hot path at the BASELINE sample sizes (no oracle run at
these sizes: the properties follow from the algebra of the reference).

LMM (lmm_cov.py:165-194, 799-815): the intercept is regressed out, so M 1 = 0 and therefore
a(1 - x) = a(x), b(1 - x) = -b(x): complementing a variant flips the sign of beta and leaves
bse, variant_h2 and the p-value unchanged.  Results do not depend on the position of a variant
in the batch nor on the batch split.  Fixed effects (model.py:274-344): complementing the variant
is the reparametrisation k -> 1 - k: kbeta -> -kbeta, intercept -> intercept + kbeta, same bse,
same LRT p-value."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _kinship_problem(n, seed=5):
    rng = np.random.RandomState(seed)
    G = (rng.uniform(size=(n, n // 2)) < rng.uniform(0.05, 0.95, n // 2)).astype(np.float32)
    K = (G @ G.T).astype(float) + np.eye(n)
    g = G.astype(float) @ rng.normal(size=n // 2)
    y = (g - g.mean()) / g.std() * np.sqrt(0.5) + np.sqrt(0.5) * rng.normal(size=n)
    return K * (n / np.diag(K).sum()), y


@pytest.mark.parametrize('n,precision', [(5000, 5), (1000, 5), (1000, 0)])
def test_lmm_complement_and_order_invariance(n, precision):
    from pyseer_b200 import lmm as plmm
    from pyseer_b200.engine import synth_host, words_per_row
    K, y = _kinship_problem(n)
    m = plmm.KinshipLMM(np.ones((n, 1)), y.reshape(-1, 1), K, precision=precision)
    h2 = m.findH2()['h2']
    nv = 3000
    ys = np.where(y > np.median(y), 1, -1).astype(np.int8)
    bits = synth_host(99, 0, nv, n, 0.05, 0.95, planted_every=100, y_sign=ys)
    # complement within the N valid samples (padding bits stay zero)
    mask = np.zeros(words_per_row(n) * 32, dtype=bool)
    mask[:n] = True
    maskw = np.packbits(mask, bitorder='little').view('<u4')
    comp = (~bits) & maskw
    kw = dict(min_af=0.01, max_af=0.99, max_missing=0.05)
    r = plmm.run_lmm_bits(m, h2, bits, None, True, 1.0, 1.0, **kw)
    rc = plmm.run_lmm_bits(m, h2, comp, None, True, 1.0, 1.0, **kw)
    ok = np.isfinite(r.pvalue)
    assert ok.sum() > 0.95 * nv and np.array_equal(ok, np.isfinite(rc.pvalue))
    assert np.allclose(rc.beta[ok], -r.beta[ok], rtol=1e-7, atol=0)
    assert np.allclose(rc.bse[ok], r.bse[ok], rtol=1e-7, atol=0)
    big = ok & (r.pvalue > 1e-290)
    assert np.allclose(rc.pvalue[big], r.pvalue[big], rtol=1e-6, atol=0)
    assert r.pvalue[ok].min() < 1e-10                     # planted tail present
    # order / batching: a permutation of the rows permutes the table, a split concatenates it
    perm = np.random.RandomState(1).permutation(nv)
    rp = plmm.run_lmm_bits(m, h2, bits[perm], None, True, 1.0, 1.0, **kw)
    assert np.array_equal(rp.pvalue[ok[perm]], r.pvalue[perm][ok[perm]])
    assert np.array_equal(rp.carriers, r.carriers[perm]) and np.array_equal(rp.flags, r.flags[perm])
    ra = plmm.run_lmm_bits(m, h2, bits[:1111], None, True, 1.0, 1.0, **kw)
    rb = plmm.run_lmm_bits(m, h2, bits[1111:], None, True, 1.0, 1.0, **kw)
    assert np.array_equal(np.r_[ra.beta, rb.beta][ok], r.beta[ok])
    m.close()


def test_fixed_complement_invariance_n2000():
    from pyseer_b200 import model as pm
    from pyseer_b200.engine import synth_host, words_per_row
    n, nv = 2000, 4000
    rng = np.random.RandomState(2)
    mds = rng.uniform(-1, 1, size=(n, 10))
    mds /= np.abs(mds).max(0)
    y = (mds[:, :3].sum(1) + rng.normal(size=n) > 0).astype(float)
    none = np.empty((0, 0))
    null = pm.fit_null(y, mds, none, False)
    firth = pm.fit_null(y, mds, none, False, True)
    model = pm.FixedModel(y, mds, none, False, null.llf, float(firth))
    ys = np.where(y > 0.5, 1, -1).astype(np.int8)
    bits = synth_host(7, 0, nv, n, 0.05, 0.95, planted_every=97, y_sign=ys)
    mask = np.zeros(words_per_row(n) * 32, dtype=bool)
    mask[:n] = True
    comp = (~bits) & np.packbits(mask, bitorder='little').view('<u4')
    r = pm.run_fixed_bits(model, bits, None, 1.0, 1.0, 0.01, 0.99, 0.05)
    rc = pm.run_fixed_bits(model, comp, None, 1.0, 1.0, 0.01, 0.99, 0.05)
    plain = ((r.flags | rc.flags) & 0x107E) == 0          # no Firth / failure notes on either side
    assert plain.sum() > 0.9 * nv
    assert np.allclose(rc.beta[plain], -r.beta[plain], rtol=1e-6, atol=1e-12)
    assert np.allclose(rc.bse[plain], r.bse[plain], rtol=1e-6, atol=0)
    assert np.allclose(rc.extra[plain], (r.extra + r.beta)[plain], rtol=1e-6, atol=1e-9)
    big = plain & (r.pvalue > 1e-290)
    assert np.allclose(rc.pvalue[big], r.pvalue[big], rtol=1e-6, atol=0)
    assert np.allclose(rc.betas[plain], r.betas[plain], rtol=1e-5, atol=1e-9)
    assert r.pvalue[plain].min() < 1e-8
    model.close()
