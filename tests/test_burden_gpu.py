"""GPU parity tests for burden regions (`--vcf --burden`, input.py:395-411): the device union of
VCF record rows (psb_submit_burden) is bit-exact against the host statement of the rule, and the
LMM / fixed-effects results of the fused path equal those of host-unioned rows (bit-identical)
and the oracle (1e-6 relative)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _host_eigh(monkeypatch):
    # these tests only need an engine of the right sample count; keep the set-up off cuSOLVER
    # (loading it costs up to a minute on a cold box)
    monkeypatch.setenv('PYSEER_B200_EIGH', 'numpy')


def _problem(n, seed=5):
    rng = np.random.RandomState(seed)
    G = (rng.uniform(size=(n, 2 * n)) < rng.uniform(0.05, 0.95, 2 * n)).astype(float)
    K = G.dot(G.T)
    g = G.dot(rng.normal(size=2 * n))
    y = (g - g.mean()) / g.std() * np.sqrt(0.5) + np.sqrt(0.5) * rng.normal(size=n)
    return K * (n / np.diag(K).sum()), y


def _records(rng, n_rec, n, with_missing):
    from pyseer_b200.engine import pack_rows
    af = rng.uniform(0.001, 0.05, size=n_rec)
    x = (rng.uniform(size=(n_rec, n)) < af[:, None]).astype(float)
    if with_missing:
        x[rng.uniform(size=x.shape) < 0.02] = np.nan
    bits, miss = pack_rows(x)
    return bits, miss


def _regions(rng, n_reg, n_rec):
    offsets, members = [0], []
    for r in range(n_reg):
        k = 0 if r % 11 == 3 else rng.randint(1, 21)
        members += list(rng.randint(0, n_rec, size=k))
        offsets.append(len(members))
    return np.array(offsets, dtype=np.int64), np.array(members, dtype=np.int32)


@pytest.mark.parametrize('n', [50, 333, 1000, 4999])
@pytest.mark.parametrize('with_missing', [False, True])
def test_device_union_bit_exact(n, with_missing):
    from pyseer_b200 import lmm as plmm
    from oracle.input_oracle import burden_union as burden_union_host
    K, y = _problem(min(n, 333))
    rng = np.random.RandomState(n)
    # the union does not depend on the model: a small LMM context with the right N
    if n > 333:
        K = np.eye(n) + 0.01
        y = rng.normal(size=n)
    m = plmm.KinshipLMM(np.ones((n, 1)), y.reshape(-1, 1), K.copy(), precision=0)
    h2 = m.findH2()['h2']
    eng = m.engine(h2)
    vbits, vmiss = _records(rng, 700, n, with_missing)
    offs, mem = _regions(rng, 500, vbits.shape[0])
    want_b, want_m = burden_union_host(vbits, vmiss, offs, mem)
    eng.submit_burden(vbits, vmiss, offs, mem)
    got_b, got_m = eng.download_rows()
    assert np.array_equal(got_b, want_b)
    if with_missing and vmiss is not None:
        assert got_m is not None and np.array_equal(got_m, want_m)
    else:
        assert got_m is None
    # empty region list and empty member list
    eng.submit_burden(vbits, vmiss, np.array([0], dtype=np.int64), np.array([], dtype=np.int32))
    b0, _ = eng.download_rows()
    assert b0.shape[0] == 0
    m.close()


def test_bad_member_lists_are_rejected():
    from pyseer_b200 import lmm as plmm
    from pyseer_b200._lib import PsbError
    n = 64
    K, y = _problem(n)
    m = plmm.KinshipLMM(np.ones((n, 1)), y.reshape(-1, 1), K.copy(), precision=0)
    eng = m.engine(m.findH2()['h2'])
    rng = np.random.RandomState(0)
    vbits, _ = _records(rng, 10, n, False)
    with pytest.raises(PsbError):
        eng.submit_burden(vbits, None, np.array([0, 2], dtype=np.int64), np.array([1, 10], dtype=np.int32))
    with pytest.raises((PsbError, ValueError)):
        eng.submit_burden(vbits, None, np.array([0, 2, 1], dtype=np.int64), np.array([1, 2], dtype=np.int32))
    with pytest.raises(PsbError):
        eng.submit_burden(vbits, None, np.array([0, 3, 2], dtype=np.int64), np.array([1, 2], dtype=np.int32))
    m.close()


@pytest.mark.parametrize('precision', [0, 5])
def test_burden_lmm_matches_oracle(precision):
    """Regions through the fused device path == host-unioned rows through run_lmm_bits
    (bit-identical) == oracle fit_lmm_block on the unioned 0/1 matrix (1e-6)."""
    from pyseer_b200 import lmm as plmm
    from pyseer_b200.engine import unpack_rows
    from oracle.input_oracle import burden_union as burden_union_host
    from oracle import lmm_oracle
    n = 300
    K, y = _problem(n)
    rng = np.random.RandomState(11)
    vbits, vmiss = _records(rng, 400, n, False)
    offs, mem = _regions(rng, 256, vbits.shape[0])
    m = plmm.KinshipLMM(np.ones((n, 1)), y.reshape(-1, 1), K.copy(), precision=precision)
    olmm, oh2, _ = lmm_oracle.initialise_lmm(y, None, K.copy())
    # both sides are evaluated at the oracle's h2 (the two h2 searches stop within 1e-5 of each
    # other on a flat minimum, which would show in beta at the 1e-5 level)
    assert abs(m.findH2()['h2'] - oh2) < 1e-5
    h2 = oh2
    r = plmm.run_lmm_burden(m, h2, vbits, vmiss, offs, mem, True, 1.0, 1.0, 0.01, 0.99, 0.05)
    ub, um = burden_union_host(vbits, vmiss, offs, mem)
    r2 = plmm.run_lmm_bits(m, h2, ub, um, True, 1.0, 1.0, 0.01, 0.99, 0.05)
    for f in ('carriers', 'missing', 'flags'):
        assert np.array_equal(getattr(r, f), getattr(r2, f)), f
    for f in ('af', 'prep', 'pvalue', 'beta', 'bse', 'extra'):
        assert np.array_equal(getattr(r, f), getattr(r2, f), equal_nan=True), f
    assert r.counts == r2.counts
    snps = unpack_rows(ub, n).T.astype(float)
    ref = lmm_oracle.fit_lmm_block(olmm, oh2, snps)
    tested = np.isfinite(r.pvalue)
    assert tested.sum() > 100
    # carriers / af are exact integers from popcounts
    assert np.array_equal(r.carriers, snps.sum(axis=0).astype(np.int32))
    for key, col in (('p_values', r.pvalue), ('beta', r.beta), ('bse', r.bse)):
        err = np.max(np.abs(col[tested] / np.asarray(ref[key]).reshape(-1)[tested] - 1))
        assert err < 1e-6, (key, err)
    m.close()


def test_burden_fixed_matches_unioned_rows():
    from pyseer_b200 import model as pmodel
    from oracle.input_oracle import burden_union as burden_union_host
    n = 200
    rng = np.random.RandomState(21)
    mds = rng.uniform(-1, 1, size=(n, 3))
    yb = (mds[:, 0] + rng.normal(size=n) > 0).astype(float)
    none = np.empty((0, 0))
    null = pmodel.fit_null(yb, mds, none, False)
    firth = pmodel.fit_null(yb, mds, none, False, True)
    fm = pmodel.FixedModel(yb, mds, none, False, null.llf, float(firth))
    vbits, vmiss = _records(rng, 300, n, True)
    offs, mem = _regions(rng, 128, vbits.shape[0])
    r = pmodel.run_fixed_burden(fm, vbits, vmiss, offs, mem, 1.0, 1.0, 0.01, 0.99, 0.5)
    ub, um = burden_union_host(vbits, vmiss, offs, mem)
    r2 = pmodel.run_fixed_bits(fm, ub, um, 1.0, 1.0, 0.01, 0.99, 0.5)
    for f in ('carriers', 'missing', 'flags'):
        assert np.array_equal(getattr(r, f), getattr(r2, f)), f
    for f in ('af', 'prep', 'pvalue', 'beta', 'bse', 'extra'):
        assert np.array_equal(getattr(r, f), getattr(r2, f), equal_nan=True), f
    fm.close()


def test_device_resident_records():
    """psb_submit_burden_device: record rows already on the device (made by psb_synth_device, whose
    host twin psb_synth_host gives the same rows) -> region rows equal the oracle union."""
    from pyseer_b200 import lmm as plmm
    from pyseer_b200.engine import synth_host
    from oracle.input_oracle import burden_union
    n = 777
    rng = np.random.RandomState(9)
    m = plmm.KinshipLMM(np.ones((n, 1)), rng.normal(size=(n, 1)), np.eye(n) + 0.01, precision=0)
    eng = m.engine(m.findH2()['h2'])
    n_rec = 1500
    eng.synth_device(123, 40, n_rec, 0.001, 0.05, 0, None)
    ptr, mptr, rows, wpr = eng.submitted_device()
    assert rows == n_rec and mptr is None
    offs, mem = _regions(rng, 400, n_rec)
    eng.submit_burden_device(ptr, n_rec, wpr, offs, mem)
    got, gm = eng.download_rows()
    host = synth_host(123, 40, n_rec, n, 0.001, 0.05)
    want, _ = burden_union(host, None, offs, mem)
    assert gm is None and np.array_equal(got, want)
    m.close()
