"""``python -m pyseer_b200.similarity`` -- sample similarity (kinship) matrix from variants.

Same interface and output as pyseer's ``similarity`` tool (pyseer/similarity.py): a list of
sample names, a variant file (``--kmers`` / ``--pres`` / ``--vcf``), AF / missing filters;
writes the N x N matrix ``K = G G'`` as a TSV with sample names.  The product runs on the GPU
as an int8 tensor-core contraction of the bit-expanded packed rows (``psb_kinship_*``); k-mer text is
tokenised on the device as well, so that such a file never exists as host rows.

One deliberate difference: a MISSING genotype counts as absent (0).  The reference keeps NaN in its
variant matrix for variants that pass ``--max-missing`` (input.py:428-430, similarity.py:99-113), so
``np.matmul`` turns the whole row and column of every sample with a missing call into NaN, and the
kinship is unusable for the LMM (``eigh`` of a NaN matrix).  Inputs without missing calls -- k-mer
files, and Rtab / VCF files with complete genotypes -- give the reference's matrix exactly
(tests/test_kinship_gpu.py)."""
import argparse
import sys

import numpy as np
import pandas as pd

from . import __version__
from .engine import Engine
from .input import VariantReader, VcfReader

BLOCK = 65536


def get_options(argv=None):
    ap = argparse.ArgumentParser(prog='similarity', description='Calculate a similarity matrix '
                                 'using variant presence/absence information')
    ap.add_argument('samples', help='List of sample names to use')
    g = ap.add_mutually_exclusive_group(required=True)
    g.add_argument('--kmers', default=None)
    g.add_argument('--vcf', default=None)
    g.add_argument('--pres', default=None)
    ap.add_argument('--min-af', type=float, default=0.01)
    ap.add_argument('--max-af', type=float, default=0.99)
    ap.add_argument('--max-missing', type=float, default=0.05)
    ap.add_argument('--uncompressed', action='store_true', default=False)
    ap.add_argument('--gpu', type=int, default=0)
    ap.add_argument('--version', action='version', version='%(prog)s ' + __version__)
    return ap.parse_args(argv)


TEXT_BLOCK = 12000       # lines per batch when the k-mer text is tokenised on the device


def similarity(p, reader, min_af, max_af, max_missing, device=0):
    """K = G G' over every variant the reader yields (filters as input.load_var_block).  k-mer text
    is tokenised on the device (``psb_submit_text``) and accumulated from there
    (``psb_kinship_add_submitted``): the rows never exist on the host."""
    import os
    eng = Engine(device)
    n_samples = len(p)
    eng.kinship_begin(n_samples)
    n = 0
    text = type(reader) is VariantReader and reader.var_type == 'kmers' and \
        os.environ.get('PYSEER_B200_TEXT', '1') != '0'
    name_bytes = sum(len(x) for x in reader.samples) + 3 * n_samples if text else 0
    if text and name_bytes + 4096 > (1 << 30):
        text = False
    if text:
        from . import pipeline
        # the text parser reads the sample count from a model: the smallest one there is
        eng.fixed_setup(np.ones((n_samples, 1)), np.arange(n_samples, dtype=float) / n_samples, True, 0.0, 0.0)
        eng.text_setup(reader.samples)
        text_bytes = int(min(max(32 << 20, TEXT_BLOCK * (name_bytes // 2 + 512)), 768 << 20))
        text_bytes = max(text_bytes, name_bytes + 4096)
        pool = pipeline.TextPool(3, TEXT_BLOCK, text_bytes)
        batches = pipeline.Prefetch(reader.text_batches(TEXT_BLOCK, pool=pool), depth=1)
        try:
            for batch in batches:
                buf, n_bytes, lstart, llen = batch.text
                eng.submit_text(buf, n_bytes, lstart, llen, batch.n)
                eng.kinship_add_submitted(min_af, max_af, max_missing)      # returns when the device is done
                info = eng.text_info(batch.n)
                for i in np.nonzero(info & 2)[0]:
                    sys.stderr.write('No observations of ' + batch.names[i] + ' in selected samples\n')
                if (info & 4).any():
                    raise ValueError("k-mer line without '|' separator")
                pool.put(batch.token)
                n += batch.n
                sys.stderr.write('Matrix size ' + str(n) + '\n')
        finally:
            batches.cancel()
            pool.close()
    else:
        for batch in reader.batches(BLOCK):
            eng.kinship_add(batch.bits, batch.missing, min_af, max_af, max_missing)
            n += batch.n
            sys.stderr.write('Matrix size ' + str(n) + '\n')
    K = eng.kinship_fetch()
    eng.close()
    return K


def main(argv=None):
    o = get_options(argv)
    with open(o.samples) as fh:
        names = [line.rstrip() for line in fh if line.strip()]
    p = pd.Series(np.zeros(len(names)), index=names)
    sys.stderr.write('Reading in variants\n')
    if o.vcf:
        reader = VcfReader(o.vcf, p)
    else:
        reader = VariantReader('kmers' if o.kmers else 'Rtab', o.kmers or o.pres, p, o.uncompressed)
    sys.stderr.write('Calculating sample similarity\n')
    K = similarity(p, reader, o.min_af, o.max_af, o.max_missing, o.gpu)
    reader.close()
    write_matrix(K, [str(x) for x in p.index], sys.stdout)
    return 0


def write_matrix(K, names, out):
    """``DataFrame(K, index=names, columns=names).to_csv(out, sep='\\t')`` (similarity.py:118-120); integer
    counts and plain names go through the library's writer (``psb_format_matrix``: pandas takes ~12 s for
    25 M entries), anything else through pandas itself."""
    import ctypes
    import os
    from . import _lib
    n = len(names)
    plain = all(x and not any(c in x for c in '\t\n\r",') for x in names)
    if plain and n > 0:
        lib = _lib.load()
        K = np.ascontiguousarray(K, dtype=np.float64)
        blob = ('\0'.join(names) + '\0').encode()
        off = np.zeros(n, dtype=np.int64)
        if n > 1:
            off[1:] = np.flatnonzero(np.frombuffer(blob, dtype=np.uint8) == 0)[:-1] + 1
        cap = int(n) * int(n) * 18 + 2 * len(blob) + n + 64
        buf = np.empty(cap, dtype=np.uint8)
        used = ctypes.c_int64(0)
        rc = lib.psb_format_matrix(K.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n, blob, off.ctypes.data, min(16, os.cpu_count() or 1),
                                   buf.ctypes.data, cap, ctypes.byref(used))
        if rc == _lib.PSB_OK:
            text = buf[:used.value].tobytes()
            if hasattr(out, 'buffer'):
                out.flush()
                out.buffer.write(text)
            else:
                out.write(text.decode())
            return
    pd.DataFrame(K, index=names, columns=names).to_csv(out, sep='\t')


if __name__ == '__main__':
    sys.exit(main())
