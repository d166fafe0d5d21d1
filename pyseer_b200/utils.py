"""Output formatting with the reference's conventions (pyseer/utils.py:39-105): tab-separated
fields, numbers as ``'%.2E'``, non-finite values left blank."""
from decimal import Decimal

import numpy as np


def _num(x):
    return '%.2E' % Decimal(float(x)) if x is not None and np.isfinite(x) else ''


def format_output(item, lineage_dict=None, model='seer', print_samples=False):
    """One result tuple (``Seer`` or ``LMM``) -> the TSV line pyseer prints."""
    fields = [str(item.kmer), _num(item.af), _num(item.prep), _num(item.pvalue), _num(item.kbeta)]
    if model not in ('enet', 'rf'):
        fields.append(_num(item.bse))
        if model == 'lmm':
            fields.append(_num(item.frac_h2))
        else:
            fields.append(_num(item.intercept))
            betas = item.betas
            # with --no-distances and no covariates there are no further coefficients
            if not np.all(np.equal(betas, None)):
                fields.append('\t'.join(_num(b) for b in betas))
    if lineage_dict is not None:
        ml = item.max_lineage
        fields.append(lineage_dict[ml] if ml is not None and np.isfinite(ml) else 'NA')
    if print_samples:
        fields.append(','.join(item.kstrains))
        fields.append(','.join(item.nkstrains))
    fields.append(','.join(item.notes))
    return '\t'.join(fields)


def format_table(r, names, model='seer', block_size=1, print_filtered=False, threads=1, lineage=None,
                 lineage_names=None):
    """TSV lines of a whole result table (``engine.Results``) through the library's native formatter
    (``psb_format_rows``): what the result loop of ``main()`` prints with ``format_output`` when
    no sample lists are asked for.  ``lineage``: int32 index per row into ``lineage_names`` (negative:
    ``NA``) for the lineage column of ``--lineage`` runs.  Returns ``(text_bytes, prefiltered, tested,
    printed)``."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    n = len(names)
    if hasattr(names, 'blob'):          # input.NameBlob: already in the formatter's layout
        blob, off = names.blob, np.ascontiguousarray(names.off, dtype=np.int64)
    else:
        blob = ('\0'.join(names) + '\0').encode()
        off = np.zeros(n, dtype=np.int64)
    # offsets of the names inside the blob: one past every NUL but the last (names hold no NUL)
    if n > 1 and not hasattr(names, 'blob'):
        nul = np.flatnonzero(np.frombuffer(blob, dtype=np.uint8) == 0)
        if nul.shape[0] == n:
            off[1:] = nul[:-1] + 1
        else:                       # a name with an embedded NUL: the slow, exact way
            lens = np.fromiter((len(s.encode()) + 1 for s in names), dtype=np.int64, count=n)
            np.cumsum(lens[:-1], out=off[1:])
    cols = _lib.PsbResults()
    keep = []
    for f, dt in (('af', np.float64), ('prep', np.float64), ('pvalue', np.float64), ('beta', np.float64),
                  ('bse', np.float64), ('extra', np.float64), ('flags', np.uint32)):
        a = np.ascontiguousarray(getattr(r, f), dtype=dt)
        keep.append(a)
        setattr(cols, f, a.ctypes.data_as(ctypes.c_void_p))
    nb = 0
    if model != 'lmm' and r.betas is not None and r.betas.ndim == 2 and r.betas.shape[1] > 0:
        b = np.ascontiguousarray(r.betas, dtype=np.float64)
        keep.append(b)
        cols.betas = b.ctypes.data_as(ctypes.c_void_p)
        nb = b.shape[1]
    cap = int(len(blob) + n * (32 * (7 + nb) + 256 + (max(len(str(x)) for x in lineage_names) + 1 if lineage is not None and len(lineage_names) else 4)) + 64)
    global _fmt_buf
    if _fmt_buf is None or _fmt_buf.shape[0] < cap:      # kept between calls (one output thread): no 20 MB
        _fmt_buf = np.empty(cap, dtype=np.uint8)         # of zeroed memory per batch
    out = _fmt_buf
    out_len = ctypes.c_int64(0)
    counts = (ctypes.c_int64 * 3)(0, 0, 0)
    if lineage is None:
        _lib.check(lib.psb_format_rows(1 if model == 'lmm' else 0, n, blob, off.ctypes.data_as(ctypes.c_void_p),
                                       ctypes.byref(cols), nb, int(block_size), int(bool(print_filtered)),
                                       int(threads), out.ctypes.data, cap, ctypes.byref(out_len), counts))
    else:
        lin = np.ascontiguousarray(lineage, dtype=np.int32)
        lnames = [str(x) for x in lineage_names]
        lblob = ('\0'.join(lnames) + '\0').encode()
        loff = np.zeros(max(len(lnames), 1), dtype=np.int64)
        if len(lnames) > 1:
            loff[1:len(lnames)] = np.cumsum([len(x.encode()) + 1 for x in lnames[:-1]])
        _lib.check(lib.psb_format_rows_lineage(
            1 if model == 'lmm' else 0, n, blob, off.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cols), nb,
            int(block_size), int(bool(print_filtered)), int(threads), lin.ctypes.data_as(ctypes.c_void_p), lblob,
            loff.ctypes.data_as(ctypes.c_void_p), len(lnames), out.ctypes.data, cap, ctypes.byref(out_len), counts))
    return out[:out_len.value].tobytes(), counts[0], counts[1], counts[2]


_fmt_buf = None
