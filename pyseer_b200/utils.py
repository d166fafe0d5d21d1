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
