"""CPU-side checks of the C ABI: the library loads, exports every symbol the header
declares, the host-evaluated special functions match scipy, packing round-trips and the
host synthetic generator is deterministic.  No GPU compute here."""
import os
import re

import numpy as np
import pytest
from scipy import stats

from conftest import ROOT
from pyseer_b200 import _lib
from pyseer_b200.engine import pack_rows, unpack_rows, synth_host, words_per_row


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, 'include', 'pyseer_b200.h')).read()
    declared = set(re.findall(r'\b(psb_[a-z0-9_]+)\s*\(', hdr))
    lib = _lib.load()
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.psb_abi_version() == 2


def test_no_gpu_fails_loudly():
    from pyseer_b200.engine import Engine, device_count
    if device_count() > 0:
        pytest.skip('GPU present')
    with pytest.raises(_lib.PsbError):
        Engine(0)


def test_special_functions():
    lib = _lib.load()
    rng = np.random.RandomState(0)
    for d in [1, 2, 3.7, 48, 998, 4998, 9998]:
        for x in 10 ** rng.uniform(-6, 3.5, 60):
            ref = stats.f.sf(x, 1, d)
            if ref > 1e-290:
                assert abs(lib.psb_host_f_sf_1(x, d) / ref - 1) < 1e-10
            ref = 2 * stats.t.sf(np.sqrt(x), d)
            if ref > 1e-290:
                assert abs(lib.psb_host_t2_sf(np.sqrt(x), d) / ref - 1) < 1e-10
    for x in 10 ** rng.uniform(-8, 3.1, 300):
        ref = stats.chi2.sf(x, 1)
        if ref > 1e-290:
            assert abs(lib.psb_host_chi2_sf1(x) / ref - 1) < 1e-10
    assert lib.psb_host_f_sf_1(0.0, 10) == 1.0
    assert np.isnan(lib.psb_host_f_sf_1(float('nan'), 10))
    assert lib.psb_host_f_sf_1(float('inf'), 10) == 0.0


def test_pack_roundtrip():
    rng = np.random.RandomState(1)
    for n in (1, 31, 32, 33, 50, 100, 129, 1000):
        k = (rng.uniform(size=(7, n)) < 0.4).astype(float)
        bits, miss = pack_rows(k)
        assert miss is None and bits.shape == (7, words_per_row(n)) and bits.shape[1] % 4 == 0
        assert np.array_equal(unpack_rows(bits, n), k.astype(np.uint8))
        k[2, n // 2] = np.nan
        bits, miss = pack_rows(k)
        assert miss is not None and unpack_rows(miss, n)[2, n // 2] == 1
        assert unpack_rows(bits, n)[2, n // 2] == 0


def test_synth_host_deterministic():
    a = synth_host(1, 100, 50, 333)
    b = synth_host(1, 0, 150, 333)
    assert np.array_equal(a, b[100:])          # rows depend on the variant id only
    x = unpack_rows(b, 333)
    af = x.mean(1)
    assert 0.0 < af.min() and af.max() < 1.0 and 0.3 < af.mean() < 0.7
    # padding bits are zero
    full = np.ascontiguousarray(b).view(np.uint8)
    assert np.unpackbits(full, axis=1, bitorder='little')[:, 333:].sum() == 0
    ys = np.where(np.arange(333) % 2 == 0, 1, -1).astype(np.int8)
    c = unpack_rows(synth_host(1, 0, 40, 333, planted_every=10, y_sign=ys), 333)
    r = [abs(np.corrcoef(c[i], ys)[0, 1]) for i in range(40)]
    assert max(r[0], r[10], r[20], r[30]) > 0.15 and np.median(r) < 0.15
