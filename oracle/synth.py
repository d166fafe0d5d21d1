"""Oracle: NumPy twin of the library's synthetic k-mer generator.  TEST INFRASTRUCTURE.

Restates ``pyseer_b200/csrc/psb_synth.cu`` (``psb_synth_host`` / ``psb_synth_device``: splitmix64
keyed by variant id, so a row does not depend on batch boundaries or on the GPU count) without
loading ``libpyseer_b200.so``: the CPU arms of ``bench.py`` (``--impl reference``,
``cpu_baseline``) and the oracle workers of the parity tests build their inputs with this, so
nothing of the product is on their path.  ``tests/test_synth_cpu.py`` holds it bit-equal to
``psb_synth_host``.
"""
import numpy as np

_C1 = np.uint64(0x9E3779B97F4A7C15)
_C2 = np.uint64(0xBF58476D1CE4E5B9)
_C3 = np.uint64(0x94D049BB133111EB)
_C4 = np.uint64(0x632BE59BD9B4E019)


def _mix64(z):
    with np.errstate(over='ignore'):
        z = z + _C1
        z = (z ^ (z >> np.uint64(30))) * _C2
        z = (z ^ (z >> np.uint64(27))) * _C3
        return z ^ (z >> np.uint64(31))


def words_per_row(n_samples):
    w = (n_samples + 31) // 32
    return (w + 3) // 4 * 4


def row_info(seed, vids, af_lo, af_hi, planted_every=0, separated_every=0):
    """(key, af, delta, sep) per variant id (psb_synth_rowinfo)."""
    vids = np.asarray(vids, dtype=np.int64)
    key = _mix64(np.uint64(seed) ^ _mix64(vids.astype(np.uint64)))
    u = (key >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    af = af_lo + (af_hi - af_lo) * u
    delta = np.zeros(vids.shape[0])
    sep = np.zeros(vids.shape[0])
    if planted_every > 0:
        pl = vids % planted_every == 0
        af[pl] = 0.5
        delta[pl] = 0.30 * ((key[pl] >> np.uint64(3)) & np.uint64(0xFFFF)).astype(np.float64) \
            * (1.0 / 65536.0)
    if separated_every > 0:
        sp = vids % separated_every == separated_every // 2
        if planted_every > 0:
            sp &= ~(vids % planted_every == 0)
        sep[sp] = 0.02 + 0.04 * ((key[sp] >> np.uint64(19)) & np.uint64(0xFFFF)).astype(np.float64) \
            * (1.0 / 65536.0)
    return key, af, delta, sep


def synth_rows(seed, first_variant, n_variants, n_samples, af_lo=0.02, af_hi=0.98,
               planted_every=0, y_sign=None, separated_every=0, chunk=2048):
    """Packed rows (n_variants, words_per_row) uint32 of variant ids first_variant .. +n_variants."""
    N = int(n_samples)
    W = words_per_row(N)
    out = np.zeros((n_variants, W * 4), dtype=np.uint8)
    ys = None if y_sign is None else np.asarray(y_sign, dtype=np.float64)
    half = (N + 1) // 2
    j = (np.arange(half, dtype=np.uint64) + np.uint64(1)) * _C4      # C4 * (w*16 + h + 1)
    for lo in range(0, n_variants, chunk):
        hi = min(lo + chunk, n_variants)
        vids = np.arange(first_variant + lo, first_variant + hi, dtype=np.int64)
        key, af, delta, sep = row_info(seed, vids, af_lo, af_hi, planted_every, separated_every)
        with np.errstate(over='ignore'):
            z = _mix64(key[:, None] + j[None, :])
        u32 = np.empty((hi - lo, half * 2), dtype=np.uint64)
        u32[:, 0::2] = z & np.uint64(0xFFFFFFFF)
        u32[:, 1::2] = z >> np.uint64(32)
        u32 = u32[:, :N]
        p = np.repeat(af[:, None], N, axis=1)
        if ys is not None:
            pl = delta != 0.0
            if pl.any():
                p[pl] = af[pl, None] + delta[pl, None] * ys[None, :]
            sp = sep != 0.0
            if sp.any():
                p[sp] = np.where(ys[None, :] > 0, sep[sp, None], 0.0)
        thr = p * 4294967296.0
        t = np.where(thr >= 4294967295.0, 4294967295.0, np.where(thr <= 0.0, 0.0, np.floor(thr)))
        bits = u32 < t.astype(np.uint64)
        by = np.packbits(bits, axis=1, bitorder='little')
        out[lo:hi, :by.shape[1]] = by
    return np.ascontiguousarray(out).view('<u4').reshape(n_variants, W)


def unpack_rows(bits, n_samples):
    """(S, W) uint32 packed rows -> (S, N) uint8."""
    by = np.ascontiguousarray(bits).view(np.uint8).reshape(bits.shape[0], -1)
    return np.unpackbits(by, axis=1, bitorder='little')[:, :n_samples]
