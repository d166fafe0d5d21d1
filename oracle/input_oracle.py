"""Oracle: VCF genotype dictionary and burden regions.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Plain-Python restatement of the dictionary that ``/root/reference/pyseer/input.py`` builds per
variant: ``read_vcf_var`` (input.py:457-502) and the burden branch of ``read_variant``
(input.py:395-411), which applies read_vcf_var for every record of every region of one burden
line to the SAME dictionary.  pysam is absent from the image, so records are given as lists of
genotype strings (``'0'``, ``'1'``, ``'.'``, ``'0/1'``, ``'./.'`` ...), one per sample, in the order
pysam would iterate them.  Pinned by the reference's end-to-end burden baselines
(tests/baseline/13.log, 37.log replayed in tests/test_cli_gpu.py) and exercised against the
packed-row rule in tests/test_burden_cpu.py.
"""
import math

import numpy as np


def read_vcf_var(genotypes, d):
    """input.py:484-497: update dictionary ``d`` (sample index -> 1 or NaN) with one record.
    ``genotypes[s]`` is the GT string of sample s, or None when the record has no GT."""
    for sample, gt in enumerate(genotypes):
        haplotypes = [None] if gt is None else gt.replace('|', '/').split('/')
        for haplotype in haplotypes:
            if (haplotype is None or haplotype == '.') and sample not in d:
                d[sample] = float('nan')
            elif haplotype is not None and haplotype != '0' and haplotype != '.':
                d[sample] = 1
                break
            elif sample in d and isinstance(d[sample], float) and math.isnan(d[sample]) \
                    and haplotype != '.':
                del d[sample]
    return d


def burden_region(records, members):
    """input.py:395-407: the dictionary after every member record of a burden line, in order."""
    d = {}
    for m in members:
        read_vcf_var(records[m], d)
    return d


def state_vector(d, n_samples):
    """0 = absent, 1 = carrier, 2 = missing (NaN), in sample order (input.py:439-452)."""
    out = [0] * n_samples
    for s, v in d.items():
        out[s] = 1 if v == 1 else 2
    return out


def burden_union(vbits, vmiss, offsets, members):
    """Union of packed VCF record rows per burden region (input.py:395-411 with the dictionary
    rules of read_vcf_var, input.py:489-497), the NumPy checker of what ``psb_submit_burden``
    computes on the device: carrier if any member record carries, missing if the LAST member
    record is missing and none carries.  Returns (bits, missing or None)."""
    R = len(offsets) - 1
    W = vbits.shape[1]
    bits = np.zeros((R, W), dtype=np.uint32)
    miss = np.zeros((R, W), dtype=np.uint32) if vmiss is not None else None
    for r in range(R):
        mem = members[offsets[r]:offsets[r + 1]]
        if len(mem) == 0:
            continue
        bits[r] = np.bitwise_or.reduce(vbits[mem], axis=0)
        if miss is not None:
            miss[r] = vmiss[mem[-1]] & ~bits[r]
    return bits, miss
