"""The NumPy twin of the synthetic generator (oracle/synth.py) is bit-equal to the library's
psb_synth_host (host code of libpyseer_b200.so: no GPU needed)."""
import numpy as np
import pytest


@pytest.mark.parametrize('n,first,nv,lo,hi,pl,sep', [
    (50, 0, 300, 0.0, 1.0, 7, 0), (333, 12345, 257, 0.02, 0.98, 50, 20),
    (1000, 10 ** 9 + 7, 100, 0.001, 0.02, 0, 0), (2000, 499, 1100, 0.02, 0.98, 1000, 1000),
    (5000, 6250000 * 7, 64, 0.02, 0.98, 1000, 0)])
def test_numpy_twin_equals_library(n, first, nv, lo, hi, pl, sep):
    from oracle import synth
    from pyseer_b200.engine import synth_host
    rng = np.random.RandomState(n)
    ys = np.where(rng.uniform(size=n) > 0.5, 1, -1).astype(np.int8)
    a = synth_host(20261017, first, nv, n, lo, hi, pl, ys, sep)
    b = synth.synth_rows(20261017, first, nv, n, lo, hi, pl, ys, sep, chunk=97)
    assert a.shape == b.shape and np.array_equal(a, b)
    if sep:
        x = synth.unpack_rows(b, n)
        vid = first + np.arange(nv)
        rows = np.where((vid % sep == sep // 2) & ~(vid % pl == 0))[0]
        assert len(rows) > 0
        for r in rows:
            assert x[r][ys < 0].sum() == 0 and x[r].sum() > 0     # carried by positive samples only
