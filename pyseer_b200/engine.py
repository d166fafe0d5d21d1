"""Thin Python driver over the C ABI: one :class:`Engine` per GPU.

This is plumbing only -- packing presence/absence vectors into the bit rows the library
expects, moving result columns back into NumPy arrays, translating flag bits into the
reference's note strings.  All per-variant arithmetic happens in the CUDA library.
"""
import ctypes
from ctypes import c_float, c_int, c_int64, c_int8, c_uint32, c_void_p, byref

import numpy as np

from . import _lib
from ._lib import PsbParams, PsbResults, check


def words_per_row(n_samples):
    """uint32 words per packed row: ceil(N/32) rounded up to 16-byte rows."""
    w = (n_samples + 31) // 32
    return (w + 3) // 4 * 4


def pack_rows(k):
    """Pack presence/absence rows into the library's bit layout.

    k: (S, N) array; non-zero finite entries are carriers, NaN entries are missing
    genotypes (input.py:428-430).  Returns (bits, missing_or_None), uint32 (S, W) arrays
    with bit (i % 32) of word (i // 32) = sample i."""
    k = np.asarray(k)
    if k.ndim == 1:
        k = k.reshape(1, -1)
    S, N = k.shape
    W = words_per_row(N)
    miss = None
    if k.dtype.kind == 'f':
        nanmask = np.isnan(k)
        present = (k == 1) | ((k != 0) & ~nanmask)
        if nanmask.any():
            miss = _pack_bool(nanmask, W)
    else:
        present = k != 0
    return _pack_bool(present, W), miss


def _pack_bool(b, W):
    S, N = b.shape
    by = np.packbits(b, axis=1, bitorder='little')
    out = np.zeros((S, W * 4), dtype=np.uint8)
    out[:, :by.shape[1]] = by
    return np.ascontiguousarray(out).view('<u4').reshape(S, W)


def unpack_rows(bits, n_samples):
    """Inverse of pack_rows for the presence bits -> (S, N) uint8."""
    by = np.ascontiguousarray(bits).view(np.uint8).reshape(bits.shape[0], -1)
    return np.unpackbits(by, axis=1, bitorder='little')[:, :n_samples]


def notes_from_flags(f):
    return set(s for bit, s in _lib.NOTE_BITS if f & bit)


class Results(object):
    """Result table of one run (NumPy columns, submission order)."""
    __slots__ = ['carriers', 'missing', 'af', 'prep', 'pvalue', 'beta', 'bse', 'extra', 'betas',
                 'flags', 'counts', 'lineage']


class Engine(object):
    """One GPU context (``psb_ctx``)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        self._ctx = c_void_p()
        check(self.lib.psb_create(int(device), byref(self._ctx)))
        self.device = device
        self.n_samples = 0
        self.q = 0
        self.model = None
        self._keep = []          # host buffers that must outlive async copies

    # -- lifecycle -------------------------------------------------------------------
    def close(self):
        if self._ctx:
            self.lib.psb_destroy(self._ctx)
            self._ctx = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        check(self.lib.psb_sync(self._ctx))

    # -- model state -----------------------------------------------------------------
    @staticmethod
    def _dptr(a):
        return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

    def lmm_setup(self, X, y, U, S, h2, precision=0):
        """State of lmm_cov.LMM after lmm.initialise_lmm (lmm.py:114-116)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        U = np.ascontiguousarray(U, dtype=np.float64)
        S = np.ascontiguousarray(S, dtype=np.float64)
        n, d = X.shape
        if y.shape[0] != n or U.shape != (n, n - d) or S.shape[0] != n - d:
            raise ValueError('shape mismatch: X %s y %s U %s S %s' %
                             (X.shape, y.shape, U.shape, S.shape))
        check(self.lib.psb_lmm_setup(self._ctx, n, d, self._dptr(X), self._dptr(y),
                                     self._dptr(U), self._dptr(S), float(h2), int(precision)))
        self.n_samples, self.q, self.model = n, 0, 'lmm'
        self.lmm_precision = int(precision)

    def fixed_setup(self, Z, y, continuous, null_llf, null_firth):
        """Shared arguments of fixed_effects_regression (model.py:202-205)."""
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        n, q = Z.shape
        if y.shape[0] != n:
            raise ValueError('shape mismatch')
        check(self.lib.psb_fixed_setup(self._ctx, n, q, self._dptr(Z), self._dptr(y),
                                       int(bool(continuous)), float(null_llf),
                                       float(null_firth)))
        self.n_samples, self.q, self.model = n, q, 'fixed'

    def eigh(self, A):
        """Symmetric eigendecomposition on the device (psb_eigh): ``(w, V)`` as numpy.linalg.eigh.
        Raises PsbError with code ERR_UNSUPPORTED when cuSOLVER cannot be loaded."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        n = A.shape[0]
        if A.ndim != 2 or A.shape[1] != n:
            raise ValueError('A must be square')
        w = np.empty(n, dtype=np.float64)
        V = np.empty((n, n), dtype=np.float64)
        check(self.lib.psb_eigh(self._ctx, n, self._dptr(A), self._dptr(w), self._dptr(V)))
        return w, V

    def spectral(self, K, X, Xdagger):
        """LMM.setSU_fromK on the device (psb_spectral): eigh of regress(regress(K + I)')."""
        K = np.ascontiguousarray(K, dtype=np.float64)
        X = np.ascontiguousarray(X, dtype=np.float64)
        Xd = np.ascontiguousarray(Xdagger, dtype=np.float64)
        n, d = X.shape
        if K.shape != (n, n) or Xd.shape != (d, n):
            raise ValueError('shape mismatch')
        w = np.empty(n, dtype=np.float64)
        V = np.empty((n, n), dtype=np.float64)
        check(self.lib.psb_spectral(self._ctx, n, d, self._dptr(K), self._dptr(X), self._dptr(Xd),
                                    self._dptr(w), self._dptr(V)))
        return w, V

    # -- kinship ------------------------------------------------------------------------
    def nll_terms(self, S, uy2, h2):
        """(yKy, logdetK) of LMM.nLLeval at each h2 (``psb_lmm_nll_terms``)."""
        S = np.ascontiguousarray(S, dtype=np.float64)
        uy2 = np.ascontiguousarray(uy2, dtype=np.float64)
        h2 = np.ascontiguousarray(np.atleast_1d(h2), dtype=np.float64)
        yky = np.empty(h2.shape[0])
        ld = np.empty(h2.shape[0])
        check(self.lib.psb_lmm_nll_terms(self._ctx, S.shape[0], self._dptr(S), self._dptr(uy2), h2.shape[0],
                                         self._dptr(h2), self._dptr(yky), self._dptr(ld)))
        return yky, ld

    def kinship_begin(self, n_samples):
        check(self.lib.psb_kinship_begin(self._ctx, int(n_samples)))
        self._kin_n = int(n_samples)

    def kinship_add(self, bits, missing=None, min_af=0.01, max_af=0.99, max_missing=0.05):
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        mp = None
        if missing is not None:
            missing = np.ascontiguousarray(missing, dtype=np.uint32)
            mp = missing.ctypes.data_as(c_void_p)
        check(self.lib.psb_kinship_add(self._ctx, bits.ctypes.data_as(c_void_p), mp, bits.shape[0],
                                       bits.shape[1], float(min_af), float(max_af),
                                       float(max_missing)))

    def kinship_add_submitted(self, min_af=0.01, max_af=0.99, max_missing=0.05):
        """``kinship_add`` for the rows of the batch last submitted (``submit`` / ``submit_text``)."""
        check(self.lib.psb_kinship_add_submitted(self._ctx, float(min_af), float(max_af), float(max_missing)))

    def kinship_fetch(self):
        K = np.empty((self._kin_n, self._kin_n), dtype=np.float64)
        check(self.lib.psb_kinship_fetch(self._ctx, self._dptr(K)))
        return K

    def lineage_setup(self, lin, cov=None):
        """Design of model.fit_lineage_effect: [1, lin, cov] (model.py:176-180)."""
        lin = np.asarray(lin, dtype=float)
        cols = [np.ones((lin.shape[0], 1)), lin]
        if cov is not None:
            cov = np.asarray(getattr(cov, 'values', cov), dtype=float)
            if cov.ndim == 2 and cov.shape[0] == lin.shape[0] and cov.shape[1] > 0:
                cols.append(cov)
        Z = np.ascontiguousarray(np.concatenate(cols, axis=1))
        check(self.lib.psb_lineage_setup(self._ctx, Z.shape[0], Z.shape[1], self._dptr(Z),
                                         lin.shape[1]))

    def run_lineage(self, lmm_rule=False):
        """Index of the most associated lineage per submitted variant (-1 = None)."""
        check(self.lib.psb_run_lineage(self._ctx, 1 if lmm_rule else 0))
        out = np.empty(getattr(self, 'n_run', self.n_variants), dtype=np.int32)
        check(self.lib.psb_fetch_lineage(self._ctx, out.ctypes.data_as(c_void_p)))
        return out

    def fit_null(self, Z, y, continuous, firth=False, start_zero=False):
        """model.fit_null on the device.  Returns (params, bse, llf, status_flags)."""
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        n, q = Z.shape
        params = np.zeros(q)
        bse = np.zeros(q)
        llf = ctypes.c_double(0.0)
        st = c_uint32(0)
        check(self.lib.psb_fit_null(self._ctx, n, q, self._dptr(Z), self._dptr(y),
                                    int(bool(continuous)), int(bool(firth)) | (2 if start_zero else 0),
                                    self._dptr(params),
                                    self._dptr(bse), byref(llf), byref(st)))
        return params, bse, llf.value, st.value

    # -- variants ----------------------------------------------------------------------
    def submit(self, bits, missing=None):
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        if bits.ndim != 2:
            raise ValueError('bits must be (n_variants, words_per_row)')
        mp = None
        if missing is not None:
            missing = np.ascontiguousarray(missing, dtype=np.uint32)
            if missing.shape != bits.shape:
                raise ValueError('missing must have the shape of bits')
            mp = missing.ctypes.data_as(c_void_p)
        self._keep = [bits, missing]
        check(self.lib.psb_submit(self._ctx, bits.ctypes.data_as(c_void_p), mp,
                                  bits.shape[0], bits.shape[1]))
        self.n_variants = bits.shape[0]

    def text_setup(self, samples):
        """Device lookup table of the sample names (phenotype order) for ``submit_text``."""
        names = (ctypes.c_char_p * len(samples))(*[str(x).encode() for x in samples])
        check(self.lib.psb_text_setup(self._ctx, names, len(samples)))
        self._text_ready = True

    def submit_text(self, text, n_bytes, line_start, line_len, n_lines):
        """``submit`` for k-mer text (``psb_submit_text``): ``text`` (uint8, ideally page-locked) holds
        ``n_lines`` lines located by ``line_start`` (int64) / ``line_len`` (int32), as
        ``TextKmerReader`` produces them; the rows are built on the device."""
        self._keep = [text, line_start, line_len]
        check(self.lib.psb_submit_text(self._ctx, c_void_p(text.ctypes.data), int(n_bytes),
                                       c_void_p(line_start.ctypes.data), c_void_p(line_len.ctypes.data),
                                       int(n_lines)))
        self.n_variants = int(n_lines)

    def text_info(self, n_lines):
        """Per-line flags of the text batch the last run worked on (2: no observation in the selected
        samples, 4: line without '|')."""
        info = np.zeros(int(n_lines), dtype=np.int32)
        check(self.lib.psb_text_info(self._ctx, c_void_p(info.ctypes.data), int(n_lines)))
        return info

    def submit_burden(self, bits, missing, region_offsets, members):
        """Burden regions (input.py:395-411): ``bits`` / ``missing`` hold one packed row per VCF
        record; region r is the union of records ``members[region_offsets[r]:region_offsets[r+1]]``.
        The union is formed on the device and left submitted (one row per region)."""
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        if bits.ndim != 2:
            raise ValueError('bits must be (n_records, words_per_row)')
        mp = None
        if missing is not None:
            missing = np.ascontiguousarray(missing, dtype=np.uint32)
            if missing.shape != bits.shape:
                raise ValueError('missing must have the shape of bits')
            mp = missing.ctypes.data_as(c_void_p)
        offs = np.ascontiguousarray(region_offsets, dtype=np.int64)
        mem = np.ascontiguousarray(members, dtype=np.int32)
        if offs.ndim != 1 or offs.shape[0] < 1 or (offs.shape[0] > 1 and offs[-1] != mem.shape[0]):
            raise ValueError('region_offsets must have n_regions + 1 entries ending at len(members)')
        self._keep = [bits, missing]
        check(self.lib.psb_submit_burden(self._ctx, bits.ctypes.data_as(c_void_p), mp, bits.shape[0],
                                         bits.shape[1], offs.ctypes.data_as(c_void_p),
                                         mem.ctypes.data_as(c_void_p), offs.shape[0] - 1))
        self.n_variants = offs.shape[0] - 1

    def submit_burden_device(self, d_bits_ptr, n_records, wpr, region_offsets, members,
                             d_missing_ptr=None):
        offs = np.ascontiguousarray(region_offsets, dtype=np.int64)
        mem = np.ascontiguousarray(members, dtype=np.int32)
        check(self.lib.psb_submit_burden_device(
            self._ctx, c_void_p(d_bits_ptr), c_void_p(d_missing_ptr) if d_missing_ptr else None,
            int(n_records), int(wpr), offs.ctypes.data_as(c_void_p), mem.ctypes.data_as(c_void_p),
            offs.shape[0] - 1))
        self.n_variants = offs.shape[0] - 1

    def submitted_device(self):
        """(device pointer of the submitted rows, of the missing rows or None, rows, words/row)"""
        b, m = c_void_p(), c_void_p()
        n, w = c_int64(0), ctypes.c_int32(0)
        check(self.lib.psb_submitted_device(self._ctx, byref(b), byref(m), byref(n), byref(w)))
        return b.value, m.value, n.value, w.value

    def download_rows(self):
        """Host copy of the submitted rows: (bits, missing or None)."""
        _, _, n, w = self.submitted_device()
        bits = np.zeros((n, w), dtype=np.uint32)
        miss = np.zeros((n, w), dtype=np.uint32)
        has = ctypes.c_int32(0)
        check(self.lib.psb_download_rows(self._ctx, bits.ctypes.data_as(c_void_p),
                                         miss.ctypes.data_as(c_void_p), byref(has)))
        return bits, (miss if has.value else None)

    def pattern_digests(self):
        """MD5 digests (n x 16 uint8) of the vectors ``input.hash_pattern`` hashes, for every row of
        the batch last submitted / run, computed on the device (``psb_pattern_digests``)."""
        _, _, n, _ = self.submitted_device()
        out = np.zeros((n, 16), dtype=np.uint8)
        check(self.lib.psb_pattern_digests(self._ctx, out.ctypes.data_as(c_void_p)))
        return out

    def submit_device(self, d_bits_ptr, n_variants, wpr, d_missing_ptr=None):
        check(self.lib.psb_submit_device(self._ctx, c_void_p(d_bits_ptr),
                                         c_void_p(d_missing_ptr) if d_missing_ptr else None,
                                         int(n_variants), int(wpr)))
        self.n_variants = int(n_variants)

    def synth_device(self, seed, first_variant, n_variants, af_lo=0.02, af_hi=0.98,
                     planted_every=0, y_sign=None, separated_every=0):
        ys = None
        if y_sign is not None:
            y_sign = np.ascontiguousarray(y_sign, dtype=np.int8)
            ys = y_sign.ctypes.data_as(ctypes.POINTER(c_int8))
        check(self.lib.psb_synth_device(self._ctx, int(seed), int(first_variant), int(n_variants),
                                        self.n_samples, af_lo, af_hi, int(planted_every),
                                        int(separated_every), ys))
        self.n_variants = int(n_variants)

    # -- run ---------------------------------------------------------------------------
    @staticmethod
    def _params(min_af, max_af, max_missing, filter_pvalue, lrt_pvalue, continuous):
        return PsbParams(float(min_af), float(max_af), float(max_missing), float(filter_pvalue),
                         float(lrt_pvalue), int(bool(continuous)), 0)

    def run_lmm(self, min_af=0.01, max_af=0.99, max_missing=0.05, filter_pvalue=1.0,
                lrt_pvalue=1.0, continuous=False):
        p = self._params(min_af, max_af, max_missing, filter_pvalue, lrt_pvalue, continuous)
        check(self.lib.psb_run_lmm(self._ctx, byref(p)))
        self.n_run = self.n_variants

    def run_fixed(self, min_af=0.01, max_af=0.99, max_missing=0.05, filter_pvalue=1.0,
                  lrt_pvalue=1.0, continuous=False):
        p = self._params(min_af, max_af, max_missing, filter_pvalue, lrt_pvalue, continuous)
        check(self.lib.psb_run_fixed(self._ctx, byref(p)))
        self.n_run = self.n_variants

    def fetch(self, columns=None):
        S = getattr(self, 'n_run', self.n_variants)      # rows of the last run, not of a later submit
        r = Results()
        r.carriers = np.empty(S, dtype=np.int32)
        r.missing = np.empty(S, dtype=np.int32)
        r.flags = np.empty(S, dtype=np.uint32)
        for f in ('af', 'prep', 'pvalue', 'beta', 'bse', 'extra'):
            setattr(r, f, np.empty(S, dtype=np.float64))
        nb = self.q - 1 if self.model == 'fixed' and self.q > 1 else 0
        r.betas = np.empty((S, nb), dtype=np.float64)
        out = PsbResults()
        for f in ('carriers', 'missing', 'af', 'prep', 'pvalue', 'beta', 'bse', 'extra', 'flags'):
            if columns is None or f in columns:
                setattr(out, f, getattr(r, f).ctypes.data_as(c_void_p))
        if nb and (columns is None or 'betas' in columns):
            out.betas = r.betas.ctypes.data_as(c_void_p)
        check(self.lib.psb_fetch(self._ctx, byref(out)))
        r.counts = self.counts()
        r.lineage = None
        return r

    def fetch_into(self, pointers):
        """psb_fetch into caller-owned buffers: ``pointers`` maps column name -> raw address
        (host or device, e.g. ``torch.Tensor.data_ptr()``)."""
        out = PsbResults()
        for f, ptr in pointers.items():
            setattr(out, f, c_void_p(int(ptr)))
        check(self.lib.psb_fetch(self._ctx, byref(out)))

    def fetch_begin(self, pointers):
        """Asynchronous ``fetch_into`` (``psb_fetch_begin``): the copies are queued behind the last run
        on the library's fetch stream; the next ``run_*`` may be queued at once.  The buffers must be
        page-locked for the copies to be truly asynchronous, and stay valid until ``fetch_wait``."""
        out = PsbResults()
        for f, ptr in pointers.items():
            setattr(out, f, c_void_p(int(ptr)))
        self._fetch_keep = out
        check(self.lib.psb_fetch_begin(self._ctx, byref(out)))

    def fetch_wait(self):
        """Blocks until the last ``fetch_begin`` has landed; returns that run's counts."""
        c = (c_int64 * 4)()
        check(self.lib.psb_fetch_wait(self._ctx, c))
        return {'loaded': c[0], 'prefiltered': c[1], 'tested': c[2], 'passed': c[3]}

    def download_bits(self, out):
        """Copy the submitted device rows into ``out`` (uint32 (S, W) array, ideally pinned)."""
        assert out.dtype == np.uint32 and out.flags['C_CONTIGUOUS']
        check(self.lib.psb_download_bits(self._ctx, out.ctypes.data_as(c_void_p)))

    def event_record(self, slot):
        check(self.lib.psb_event_record(self._ctx, int(slot)))

    def event_elapsed(self, a, b):
        ms = c_float(0)
        check(self.lib.psb_event_elapsed(self._ctx, int(a), int(b), byref(ms)))
        return ms.value

    def results_device(self):
        out = PsbResults()
        check(self.lib.psb_results_device(self._ctx, byref(out)))
        return out

    def counts(self):
        c = (c_int64 * 4)()
        check(self.lib.psb_counts(self._ctx, c))
        return {'loaded': c[0], 'prefiltered': c[1], 'tested': c[2], 'passed': c[3]}

    def last_stats(self):
        c = (c_int64 * 4)()
        check(self.lib.psb_last_stats(self._ctx, c))
        return {'newton_evaluations': c[0], 'firth_fits': c[1], 'lrt_filtered': c[2],
                'firth_singular_evaluations': c[3]}

    def last_ms(self, which=0):
        ms = c_float(0)
        check(self.lib.psb_last_ms(self._ctx, which, byref(ms)))
        return ms.value

    def measure_peaks(self):
        """Measured int8-tensor and fp64 peak rates under the current clocks (psb_measure_peaks)."""
        v = (ctypes.c_double * 4)()
        check(self.lib.psb_measure_peaks(self._ctx, v))
        return {'int8_tops': v[0], 'fp64_tflops': v[1], 'int8_probe_ms': v[2], 'fp64_probe_ms': v[3]}

    def launch_count(self):
        n = c_int64(0)
        check(self.lib.psb_launch_count(self._ctx, byref(n)))
        return n.value


class PinnedBuffer(object):
    """Page-locked host array (cudaHostAlloc) for psb_submit / psb_fetch staging."""

    def __init__(self, shape, dtype):
        self.lib = _lib.load()
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        self._ptr = c_void_p()
        check(self.lib.psb_host_alloc(nbytes, byref(self._ptr)))
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._ptr:
            self.array = None
            self.lib.psb_host_free(self._ptr)
            self._ptr = c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def device_count():
    n = c_int(0)
    check(_lib.load().psb_device_count(byref(n)))
    return n.value


def synth_host(seed, first_variant, n_variants, n_samples, af_lo=0.02, af_hi=0.98,
               planted_every=0, y_sign=None, separated_every=0):
    """Host twin of Engine.synth_device (same generator, same rows)."""
    lib = _lib.load()
    W = words_per_row(n_samples)
    out = np.zeros((n_variants, W), dtype=np.uint32)
    ys = None
    if y_sign is not None:
        y_sign = np.ascontiguousarray(y_sign, dtype=np.int8)
        ys = y_sign.ctypes.data_as(ctypes.POINTER(c_int8))
    check(lib.psb_synth_host(int(seed), int(first_variant), int(n_variants), int(n_samples),
                             af_lo, af_hi, int(planted_every), int(separated_every), ys,
                             out.ctypes.data_as(ctypes.POINTER(c_uint32)), W))
    return out
