"""``python -m pyseer_b200.similarity`` -- sample similarity (kinship) matrix from variants.

Same interface and output as pyseer's ``similarity`` tool (pyseer/similarity.py): a list of
sample names, a variant file (``--kmers`` / ``--pres`` / ``--vcf``), AF / missing filters;
writes the N x N matrix ``K = G G'`` as a TSV with sample names.  The product runs on the GPU
as an int8 tensor-core contraction of the bit-expanded packed rows (``psb_kinship_*``).

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


def similarity(p, reader, min_af, max_af, max_missing, device=0):
    """K = G G' over every variant the reader yields (filters as input.load_var_block)."""
    eng = Engine(device)
    eng.kinship_begin(len(p))
    n = 0
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
    pd.DataFrame(K, index=p.index, columns=p.index).to_csv(sys.stdout, sep='\t')
    return 0


if __name__ == '__main__':
    sys.exit(main())
