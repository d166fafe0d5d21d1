"""ctypes binding of libpyseer_b200.so (C ABI: include/pyseer_b200.h).

The library is the product; this module fails loudly when it is missing.  It never falls
back to a CPU implementation.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64,
                    c_int8, c_uint32, c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpyseer_b200.so')

PSB_OK = 0
ABI_VERSION = 2
PSB_ERR_CUDA, PSB_ERR_ARG, PSB_ERR_STATE, PSB_ERR_H2, PSB_ERR_NOMEM, PSB_ERR_UNSUPPORTED = \
    -1, -2, -3, -4, -5, -6

F_AF_FILTER = 0x0001
F_PREFILTER_FAILED = 0x0002
F_BAD_CHISQ = 0x0004
F_HIGH_BSE = 0x0008
F_PERFECT_SEP = 0x0010
F_MATRIX_INV = 0x0020
F_FIRTH_FAIL = 0x0040
F_MISSING_DATA = 0x0080
F_LRT_FAILED = 0x0100
F_PREFILTER = 0x0200
F_FILTER = 0x0400
F_TESTED = 0x0800
F_FIRTH_USED = 0x1000
OPT_NO_PREFILTER = 0x1

#: flag bit -> reference note string (model.py / lmm.py)
NOTE_BITS = [
    (F_AF_FILTER, 'af-filter'),
    (F_PREFILTER_FAILED, 'pre-filtering-failed'),
    (F_BAD_CHISQ, 'bad-chisq'),
    (F_HIGH_BSE, 'high-bse'),
    (F_PERFECT_SEP, 'perfectly-separable-data'),
    (F_MATRIX_INV, 'matrix-inversion-error'),
    (F_FIRTH_FAIL, 'firth-fail'),
    (F_MISSING_DATA, 'missing-data-error'),
    (F_LRT_FAILED, 'lrt-filtering-failed'),
]


class PsbParams(Structure):
    _fields_ = [('min_af', c_double), ('max_af', c_double), ('max_missing', c_double),
                ('filter_pvalue', c_double), ('lrt_pvalue', c_double),
                ('continuous', c_int32), ('options', c_int32)]


class PsbResults(Structure):
    _fields_ = [('carriers', c_void_p), ('missing', c_void_p), ('af', c_void_p),
                ('prep', c_void_p), ('pvalue', c_void_p), ('beta', c_void_p),
                ('bse', c_void_p), ('extra', c_void_p), ('betas', c_void_p),
                ('flags', c_void_p)]


class LibraryMissing(ImportError):
    pass


class PsbError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, 'libpyseer_b200 error %d: %s' % (code, msg))
        self.code = code


_lib = None

#: every symbol include/pyseer_b200.h declares
SYMBOLS = ['psb_abi_version', 'psb_last_error', 'psb_device_count', 'psb_create', 'psb_destroy',
           'psb_sync', 'psb_lmm_setup', 'psb_fixed_setup', 'psb_fit_null', 'psb_submit',
           'psb_submit_device', 'psb_run_lmm', 'psb_run_fixed', 'psb_fetch',
           'psb_results_device', 'psb_counts', 'psb_last_ms', 'psb_launch_count',
           'psb_host_alloc', 'psb_host_free', 'psb_download_bits', 'psb_event_record',
           'psb_event_elapsed', 'psb_reader_open', 'psb_reader_next', 'psb_reader_close',
           'psb_lineage_setup', 'psb_run_lineage', 'psb_fetch_lineage', 'psb_last_stats', 'psb_kinship_begin', 'psb_kinship_add', 'psb_kinship_add_submitted', 'psb_kinship_fetch', 'psb_format_matrix',
           'psb_synth_device', 'psb_synth_host', 'psb_host_chi2_sf1', 'psb_host_f_sf_1',
           'psb_host_t2_sf', 'psb_submit_burden', 'psb_submit_burden_device',
           'psb_submitted_device', 'psb_download_rows', 'psb_eigh', 'psb_reader_set_threads', 'psb_pgz_selftest', 'psb_format_rows', 'psb_format_rows_lineage', 'psb_reader_vcf_info', 'psb_hash_patterns', 'psb_pattern_digests',
           'psb_comm_unique_id', 'psb_comm_init_rank', 'psb_comm_init_all', 'psb_comm_destroy',
           'psb_comm_info', 'psb_comm_bcast', 'psb_comm_allreduce', 'psb_comm_barrier',
           'psb_comm_gather_begin', 'psb_comm_gather_wait', 'psb_comm_gather_fetch',
           'psb_comm_gather_bytes', 'psb_measure_peaks', 'psb_reader_at_eof', 'psb_spectral',
           'psb_reader_next_text', 'psb_text_setup', 'psb_submit_text', 'psb_text_info',
           'psb_lmm_nll_terms', 'psb_fetch_begin', 'psb_fetch_wait']


def load():
    """Load the shared library (once) and declare its prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            '%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            'or `make -C pyseer_b200/csrc`.  pyseer_b200 has no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    dp = POINTER(c_double)
    lib.psb_abi_version.restype = c_int
    if lib.psb_abi_version() != ABI_VERSION:
        raise LibraryMissing('%s has ABI version %d, this package needs %d: rebuild it'
                             % (LIB_PATH, lib.psb_abi_version(), ABI_VERSION))
    lib.psb_last_error.restype = c_char_p
    lib.psb_device_count.argtypes = [POINTER(c_int)]
    lib.psb_create.argtypes = [c_int, POINTER(c_void_p)]
    lib.psb_destroy.argtypes = [c_void_p]
    lib.psb_sync.argtypes = [c_void_p]
    lib.psb_lmm_setup.argtypes = [c_void_p, c_int32, c_int32, dp, dp, dp, dp, c_double, c_int32]
    lib.psb_fixed_setup.argtypes = [c_void_p, c_int32, c_int32, dp, dp, c_int32, c_double, c_double]
    lib.psb_fit_null.argtypes = [c_void_p, c_int32, c_int32, dp, dp, c_int32, c_int32, dp, dp, dp,
                                 POINTER(c_uint32)]
    lib.psb_submit.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32]
    lib.psb_submit_device.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32]
    lib.psb_submit_burden.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p,
                                      c_void_p, c_int64]
    lib.psb_submit_burden_device.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                             c_void_p, c_void_p, c_int64]
    lib.psb_submitted_device.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_void_p),
                                         POINTER(c_int64), POINTER(c_int32)]
    lib.psb_download_rows.argtypes = [c_void_p, c_void_p, c_void_p, POINTER(c_int32)]
    lib.psb_eigh.argtypes = [c_void_p, c_int32, dp, dp, dp]
    lib.psb_spectral.argtypes = [c_void_p, c_int32, c_int32, dp, dp, dp, dp, dp]
    lib.psb_reader_set_threads.argtypes = [c_void_p, c_int32]
    lib.psb_reader_next_text.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                                         c_void_p, c_int64, c_void_p, POINTER(c_int64), POINTER(c_int64)]
    lib.psb_text_setup.argtypes = [c_void_p, POINTER(ctypes.c_char_p), c_int32]
    lib.psb_submit_text.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64]
    lib.psb_text_info.argtypes = [c_void_p, c_void_p, c_int64]
    lib.psb_lmm_nll_terms.argtypes = [c_void_p, c_int32, dp, dp, c_int32, dp, dp, dp]
    lib.psb_hash_patterns.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                      POINTER(c_int64)]
    lib.psb_reader_vcf_info.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
    lib.psb_format_rows.argtypes = [c_int32, c_int64, c_void_p, c_void_p, POINTER(PsbResults), c_int32,
                                    c_int32, c_int32, c_int32, c_void_p, c_int64, POINTER(c_int64),
                                    POINTER(c_int64)]
    lib.psb_format_rows_lineage.argtypes = [c_int32, c_int64, c_void_p, c_void_p, POINTER(PsbResults), c_int32,
                                            c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32,
                                            c_void_p, c_int64, POINTER(c_int64), POINTER(c_int64)]
    lib.psb_run_lmm.argtypes = [c_void_p, POINTER(PsbParams)]
    lib.psb_run_fixed.argtypes = [c_void_p, POINTER(PsbParams)]
    lib.psb_fetch.argtypes = [c_void_p, POINTER(PsbResults)]
    lib.psb_fetch_begin.argtypes = [c_void_p, POINTER(PsbResults)]
    lib.psb_fetch_wait.argtypes = [c_void_p, POINTER(c_int64)]
    lib.psb_results_device.argtypes = [c_void_p, POINTER(PsbResults)]
    lib.psb_counts.argtypes = [c_void_p, POINTER(c_int64)]
    lib.psb_last_ms.argtypes = [c_void_p, c_int32, POINTER(c_float)]
    lib.psb_launch_count.argtypes = [c_void_p, POINTER(c_int64)]
    lib.psb_host_alloc.argtypes = [ctypes.c_size_t, POINTER(c_void_p)]
    lib.psb_host_free.argtypes = [c_void_p]
    lib.psb_download_bits.argtypes = [c_void_p, c_void_p]
    lib.psb_event_record.argtypes = [c_void_p, c_int32]
    lib.psb_event_elapsed.argtypes = [c_void_p, c_int32, c_int32, POINTER(c_float)]
    lib.psb_reader_open.argtypes = [c_char_p, c_int32, POINTER(c_char_p), c_int32, POINTER(c_void_p)]
    lib.psb_reader_next.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_int64,
                                    c_void_p, c_void_p, POINTER(c_int64), POINTER(c_int32)]
    lib.psb_reader_close.argtypes = [c_void_p]
    lib.psb_reader_at_eof.argtypes = [c_void_p, POINTER(c_int32)]
    lib.psb_pattern_digests.argtypes = [c_void_p, c_void_p]
    lib.psb_pgz_selftest.argtypes = [c_char_p, c_int32, c_int64, c_void_p, POINTER(c_int64), POINTER(c_int64)]
    lib.psb_lineage_setup.argtypes = [c_void_p, c_int32, c_int32, dp, c_int32]
    lib.psb_run_lineage.argtypes = [c_void_p, c_int32]
    lib.psb_fetch_lineage.argtypes = [c_void_p, c_void_p]
    lib.psb_last_stats.argtypes = [c_void_p, POINTER(c_int64)]
    lib.psb_kinship_begin.argtypes = [c_void_p, c_int32]
    lib.psb_kinship_add.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_double, c_double,
                                    c_double]
    lib.psb_kinship_add_submitted.argtypes = [c_void_p, c_double, c_double, c_double]
    lib.psb_kinship_fetch.argtypes = [c_void_p, dp]
    lib.psb_format_matrix.argtypes = [dp, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_int64, POINTER(c_int64)]
    lib.psb_comm_unique_id.argtypes = [c_void_p]
    lib.psb_comm_init_rank.argtypes = [c_void_p, c_int32, c_int32, c_void_p, POINTER(c_void_p)]
    lib.psb_comm_init_all.argtypes = [POINTER(c_void_p), c_int32, POINTER(c_void_p)]
    lib.psb_comm_destroy.argtypes = [c_void_p]
    lib.psb_comm_info.argtypes = [c_void_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]
    lib.psb_comm_bcast.argtypes = [c_void_p, c_void_p, ctypes.c_size_t, c_int32]
    lib.psb_comm_allreduce.argtypes = [c_void_p, dp, c_int32, c_int32]
    lib.psb_comm_barrier.argtypes = [c_void_p]
    lib.psb_comm_gather_begin.argtypes = [c_void_p, c_int32, c_int64]
    lib.psb_comm_gather_wait.argtypes = [c_void_p]
    lib.psb_comm_gather_fetch.argtypes = [c_void_p, c_int32, POINTER(PsbResults), POINTER(c_int64),
                                          POINTER(c_int64)]
    lib.psb_comm_gather_bytes.argtypes = [c_void_p, POINTER(c_int64)]
    lib.psb_measure_peaks.argtypes = [c_void_p, dp]
    lib.psb_synth_device.argtypes = [c_void_p, c_uint64, c_int64, c_int64, c_int32, c_double,
                                     c_double, c_int32, c_int32, POINTER(c_int8)]
    lib.psb_synth_host.argtypes = [c_uint64, c_int64, c_int64, c_int32, c_double, c_double,
                                   c_int32, c_int32, POINTER(c_int8), POINTER(c_uint32), c_int32]
    for f in ('psb_host_chi2_sf1',):
        getattr(lib, f).restype = c_double
        getattr(lib, f).argtypes = [c_double]
    for f in ('psb_host_f_sf_1', 'psb_host_t2_sf'):
        getattr(lib, f).restype = c_double
        getattr(lib, f).argtypes = [c_double, c_double]
    _lib = lib
    return lib


def check(code):
    if code != PSB_OK:
        raise PsbError(code, load().psb_last_error().decode('utf-8', 'replace'))
