"""Kinship matrix K = G G' (pyseer/similarity.py:99-116) on the GPU vs NumPy."""
import contextlib
import io
import os

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(params=['tensor', 'popcount'])
def kin_mode(request, monkeypatch):
    """Both contraction paths of psb_kinship_add: tcgen05 int8 (default) and AND + POPCOUNT
    (PSB_KIN_TC=0, read on every call)."""
    if request.param == 'popcount':
        monkeypatch.setenv('PSB_KIN_TC', '0')
    else:
        monkeypatch.delenv('PSB_KIN_TC', raising=False)
    return request.param


@pytest.mark.parametrize('n,nv', [(50, 333), (333, 5000), (1000, 4097), (128, 128), (129, 1)])
def test_kinship_matches_numpy(n, nv, kin_mode):
    from pyseer_b200.engine import Engine, synth_host, unpack_rows
    bits = synth_host(11, 0, nv, n, af_lo=0.0, af_hi=1.0)
    x = unpack_rows(bits, n).astype(np.int64)
    af = x.sum(1) / float(n)
    keep = ~((af < 0.05) | (af > 0.9))
    G = x[keep].T
    ref = G @ G.T
    eng = Engine(0)
    eng.kinship_begin(n)
    half = nv // 3
    eng.kinship_add(bits[:half], None, 0.05, 0.9, 0.05)       # accumulates over batches
    eng.kinship_add(bits[half:], None, 0.05, 0.9, 0.05)
    K = eng.kinship_fetch()
    eng.close()
    assert np.array_equal(K, ref.astype(float))


def test_kinship_tensor_path_chunks_and_splits():
    """More variants than one expansion chunk (65536), a sample count that leaves a ragged last
    tile, fewer tiles than SMs (variant axis split over CTAs, 64-bit atomic adds) -- against BLAS in
    float64 (exact: every entry < 2^53) and against the popcount path bit for bit."""
    from pyseer_b200.engine import Engine, synth_host, unpack_rows
    n, nv = 1300, 70001
    bits = synth_host(5, 0, nv, n, af_lo=0.0, af_hi=1.0)
    x = unpack_rows(bits, n)
    af = x.sum(1) / float(n)
    keep = ~((af < 0.02) | (af > 0.97))
    G = x[keep].T.astype(np.float64)
    ref = G @ G.T
    out = {}
    for mode in ('1', '0'):
        os.environ['PSB_KIN_TC'] = mode
        try:
            eng = Engine(0)
            eng.kinship_begin(n)
            eng.kinship_add(bits, None, 0.02, 0.97, 0.05)
            eng.kinship_add(bits[:1000], None, 0.02, 0.97, 0.05)      # a second, short batch
            out[mode] = eng.kinship_fetch()
            eng.close()
        finally:
            del os.environ['PSB_KIN_TC']
    G2 = G[:, :int(keep[:1000].sum())]
    assert np.array_equal(out['1'], ref + G2 @ G2.T)
    assert np.array_equal(out['1'], out['0'])


def test_kinship_tensor_path_many_tiles():
    """N = 2500 (210 upper-triangular tiles: no split, ragged last tile), both paths bit for bit."""
    from pyseer_b200.engine import Engine, synth_host
    n, nv = 2500, 20000
    bits = synth_host(9, 0, nv, n, af_lo=0.0, af_hi=1.0)
    out = {}
    for mode in ('1', '0'):
        os.environ['PSB_KIN_TC'] = mode
        try:
            eng = Engine(0)
            eng.kinship_begin(n)
            eng.kinship_add(bits, None, 0.01, 0.99, 0.05)
            out[mode] = eng.kinship_fetch()
            eng.close()
        finally:
            del os.environ['PSB_KIN_TC']
    assert out['1'].max() > 0 and np.array_equal(out['1'], out['1'].T)
    assert np.array_equal(out['1'], out['0'])


@pytest.mark.parametrize('text', ['1', '0'])
def test_similarity_tool_on_reference_kmers(tmp_path, monkeypatch, text):
    """The similarity tool on the reference's k-mer fixture, with the text tokenised on the device and
    accumulated from the device rows (psb_kinship_add_submitted) and with the host parser."""
    from pyseer_b200.similarity import main
    monkeypatch.setenv('PYSEER_B200_TEXT', text)
    from pyseer_b200.input import VariantReader, load_phenotypes
    from pyseer_b200.engine import unpack_rows
    p = load_phenotypes(os.path.join(GOLDEN, 'subset.pheno'), None)
    samples = tmp_path / 'samples.txt'
    samples.write_text('\n'.join(p.index) + '\n')
    out = io.StringIO()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(io.StringIO()):
        main([str(samples), '--kmers', os.path.join(GOLDEN, 'kmers.gz')])
    K = pd.read_csv(io.StringIO(out.getvalue()), sep='\t', index_col=0)
    assert list(K.index) == list(p.index) and list(K.columns) == list(p.index)
    rd = VariantReader('kmers', os.path.join(GOLDEN, 'kmers.gz'), p)
    with contextlib.redirect_stderr(io.StringIO()):
        x = np.concatenate([unpack_rows(b.bits, len(p)) for b in rd.batches(100)]).astype(np.int64)
    af = x.sum(1) / float(len(p))
    G = x[(af >= 0.01) & (af <= 0.99)].T
    assert np.array_equal(K.values, (G @ G.T).astype(float))


def test_kinship_with_missing_genotypes(kin_mode):
    """Missing calls count as absent (documented difference, pyseer_b200/similarity.py): the matrix
    equals the reference's G G' with NaN replaced by 0, for variants within --max-missing; the AF
    of the filter counts missing samples as carriers (input.py:439-446)."""
    from pyseer_b200.engine import Engine, pack_rows
    n, nv = 120, 400
    rng = np.random.RandomState(2)
    k = (rng.uniform(size=(nv, n)) < rng.uniform(0.05, 0.9, size=(nv, 1))).astype(float)
    k[rng.uniform(size=(nv, n)) < 0.01] = np.nan
    k[7, :30] = np.nan                                   # 25 % missing: beyond --max-missing
    bits, miss = pack_rows(k)
    nanmask = np.isnan(k)
    af = (np.nansum(k, axis=1) + nanmask.sum(1)) / float(n)
    keep = (af >= 0.01) & (af <= 0.99) & (nanmask.sum(1) / float(n) <= 0.05)
    G = np.nan_to_num(k[keep]).T.astype(np.int64)
    eng = Engine(0)
    eng.kinship_begin(n)
    eng.kinship_add(bits, miss, 0.01, 0.99, 0.05)
    K = eng.kinship_fetch()
    eng.close()
    assert not keep[7] and np.array_equal(K, (G @ G.T).astype(float))
