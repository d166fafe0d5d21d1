"""Oracle: linear mixed model path.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

NumPy restatement of ``/root/reference/pyseer/lmm.py`` and of the slice of the
vendored FaST-LMM code pyseer actually executes (single full-rank kernel,
``dof=None``, ``penalty=0``, ``UW=None``): ``pyseer/fastlmm/lmm_cov.py`` and
``pyseer/fastlmm/mingrid.py``.  Checked against the reference's goldens
(tests/lmm_test.py) and against vectors produced by the unmodified reference
module (tests/golden/lmm_ref_*.npz, made by oracle/gen_golden.py).
"""
from collections import namedtuple

import numpy as np
import scipy.optimize as opt
from scipy import stats

from .fixed_oracle import pre_filtering, fit_lineage_effect

# pyseer/classes.py:3-9
LMM = namedtuple('LMM', ['kmer', 'pattern', 'af', 'prep', 'pvalue', 'kbeta', 'bse',
                         'frac_h2', 'max_lineage', 'kstrains', 'nkstrains', 'notes',
                         'prefilter', 'filter'])


class OracleLMM(object):
    """lmm_cov.py:18-103, 165-218, 427-478, 597-838 restricted to pyseer's use."""

    def __init__(self, X, Y, K):
        self.X = np.asarray(X, dtype=float)
        self.Y = np.asarray(Y, dtype=float).reshape(X.shape[0], -1)
        self.K = None if K is None else np.array(K, dtype=float)
        self.D = self.X.shape[1]
        self.Xdagger = None
        self.U = None
        self.S = None
        self.UY = None

    # Linreg.regress, lmm_cov.py:861-880
    def regress(self, Y):
        if self.Xdagger is None:
            self.Xdagger = np.linalg.pinv(self.X)
        return Y - self.X.dot(self.Xdagger.dot(Y))

    # lmm_cov.py:88-103
    def setSU_fromK(self):
        N = self.K.shape[0]
        self.K.flat[::N + 1] += 1.0
        K_ = self.regress(self.K)
        K_ = self.regress(K_.T)
        S, U = np.linalg.eigh(K_)
        self.U = U[:, self.D:N]
        self.S = S[self.D:N] - 1.0

    def getSU(self):
        if self.U is None or self.S is None:
            self.setSU_fromK()
        return self.S, self.U

    # lmm_cov.py:165-194 (full-rank branch)
    def rotate(self, A):
        S, U = self.getSU()
        A = self.regress(A)
        A_std = A.std(0)
        A[:, A_std <= 1e-10] = 0.0
        return U.T.dot(A)

    def getUY(self):
        if self.UY is None:
            self.UY = self.rotate(self.Y)
        return self.UY

    # lmm_cov.py:597-684 + 686-838
    def nLLeval(self, h2=0.0, snps=None):
        N = self.Y.shape[0] - self.D
        S, U = self.getSU()
        Sd = h2 * S + (1.0 - h2)
        if h2 < 0.0 or h2 >= 1.0:
            return {'nLL': 3e20, 'h2': h2, 'scale': 1.0}
        UY = self.getUY()
        Usnps = None
        if snps is not None:
            assert snps.shape[0] == self.Y.shape[0], "shape missmatch between snps and Y"
            Usnps = self.rotate(np.array(snps, dtype=float))
        YKY = (UY / Sd.reshape(-1, 1) * UY).sum(0)
        logdetK = np.log(Sd).sum()
        res = {'h2': h2, 'scale': 1.0, 'dof': None}
        if Usnps is not None:
            with np.errstate(all='ignore'):
                snpsKsnps = (Usnps / Sd.reshape(-1, 1) * Usnps).sum(0)[:, np.newaxis]
                snpsKY = (Usnps / Sd.reshape(-1, 1)).T.dot(UY)
                beta = snpsKY / snpsKsnps
                if np.isnan(beta.min()):
                    beta[snpsKY == 0] = 0.0
                veb = snpsKY * beta
                r2 = YKY[np.newaxis, :] - veb
                variance_beta = r2 / (N - 1.0) / snpsKsnps
                frac = veb / YKY[np.newaxis, :]
            res.update(beta=beta, variance_beta=variance_beta,
                       variance_explained_beta=veb,
                       fraction_variance_explained_beta=frac)
        else:
            r2 = YKY
        with np.errstate(all='ignore'):
            sigma2 = r2 / N
            res['nLL'] = 0.5 * (logdetK + N * (np.log(2.0 * np.pi * sigma2) + 1))
        return res

    # lmm_cov.py:427-478 (single phenotype branch) + mingrid.py:13-103
    def findH2(self, nGridH2=10, minH2=0.0, maxH2=0.99999):
        resmin = [None]

        def f(x):
            res = self.nLLeval(h2=x)
            if (resmin[0] is None) or (res['nLL'] < resmin[0]['nLL']):
                resmin[0] = res
            return res['nLL'][0]

        minimize1D(f, nGrid=nGridH2, minval=minH2, maxval=maxH2)
        return resmin[0]


def minimize1D(f, nGrid=10, minval=0.0, maxval=0.99999):
    """mingrid.py:13-103."""
    step = (maxval - minval) / nGrid
    evalgrid = np.arange(minval, maxval + step, step)
    resultgrid = np.array([f(x) for x in evalgrid])
    i = resultgrid.argmin()
    minglobal = (evalgrid[i], resultgrid[i])
    if resultgrid[0] < resultgrid[1]:
        ml = opt.fminbound(f, evalgrid[0], evalgrid[1], full_output=True)
        if ml[1] < minglobal[1]:
            minglobal = ml[0:2]
    if resultgrid[-1] < resultgrid[-2]:
        ml = opt.fminbound(f, evalgrid[-2], evalgrid[-1], full_output=True)
        if ml[1] < minglobal[1]:
            minglobal = ml[0:2]
    for i in range(resultgrid.shape[0] - 2):
        if resultgrid[i + 1] < resultgrid[i + 2] and resultgrid[i + 1] < resultgrid[i]:
            ml = opt.brent(f, brack=(evalgrid[i], evalgrid[i + 1], evalgrid[i + 2]),
                           full_output=True)
            if ml[1] < minglobal[1]:
                minglobal = ml[0:2]
    return minglobal


def initialise_lmm(y, covariates, K):
    """lmm.py:93-116 (after sample intersection): normalise K, build the model,
    find h2.  ``covariates`` is an (N, ncov) array or None."""
    y = np.asarray(y, dtype=float).reshape(-1, 1)
    n = y.shape[0]
    if covariates is not None and np.asarray(covariates).size > 0:
        covar = np.c_[np.asarray(covariates, dtype=float), np.ones((n, 1))]
    else:
        covar = np.ones((n, 1))
    K = np.array(K, dtype=float)
    factor = float(n) / np.diag(K).sum()
    if abs(factor - 1.0) > 1e-15:
        K *= factor
    lmm = OracleLMM(covar, y, K)
    res = lmm.findH2()
    return lmm, res['h2'], res


def fit_lmm_block(lmm, h2, variant_block):
    """lmm.py:228-260."""
    res = lmm.nLLeval(h2=h2, snps=variant_block)
    beta = res['beta']          # KeyError when h2 >= 1, as in the reference
    with np.errstate(all='ignore'):
        chi2stats = beta * beta / res['variance_beta']
        out = {
            'p_values': stats.f.sf(chi2stats, 1, lmm.U.shape[0] - (lmm.D + 1))[:, 0],
            'beta': beta[:, 0],
            'bse': np.sqrt(res['variance_beta'][:, 0]),
            'frac_h2': np.sqrt(res['fraction_variance_explained_beta'][:, 0]),
        }
    return out


def fit_lmm(lmm, h2, variants, variant_mat, lineage_effects, lineage_clusters,
            covariates, continuous, filter_pvalue, lrt_pvalue):
    """lmm.py:125-226.  ``variants`` = list of (LMM tuple, y, k)."""
    all_variants = []
    filtered_variants = []
    k = None
    for var_idx, variant in enumerate(variants):
        notes = set()
        var, p, k = variant
        if var.pattern is None or k is None:
            notes.add('af-filter')
            all_variants.append(var._replace(notes=notes, prefilter=True, filter=False))
            variant_mat[:, var_idx] = 0.0
            continue
        prep, bad_chisq = pre_filtering(p, k, continuous)
        if bad_chisq:
            notes.add('bad-chisq')
        if prep >= filter_pvalue or not np.isfinite(prep):
            notes.add('pre-filtering-failed')
            all_variants.append(var._replace(notes=notes, prep=prep, prefilter=True,
                                             filter=False))
            variant_mat[:, var_idx] = 0.0
            continue
        filtered_variants.append(var._replace(prep=prep, notes=notes, prefilter=False))
    variant_mat = variant_mat[:, ~np.all(variant_mat == 0, axis=0)]
    if variant_mat.shape[1] == 0:
        return all_variants
    res = fit_lmm_block(lmm, h2, variant_mat)
    assert len(res['p_values']) == len(filtered_variants)
    for i, tv in enumerate(filtered_variants):
        notes = tv.notes
        pv = res['p_values'][i]
        if pv >= lrt_pvalue or not np.isfinite(pv):
            notes.add('lrt-filtering-failed')
            all_variants.append(tv._replace(notes=notes, pvalue=pv, filter=True))
        else:
            # lmm.py:209-211 -- note the stale ``k`` (last variant of the block)
            max_lineage = fit_lineage_effect(lineage_clusters, covariates, k) \
                if lineage_effects else None
            all_variants.append(tv._replace(pvalue=pv, kbeta=res['beta'][i],
                                            bse=res['bse'][i], frac_h2=res['frac_h2'][i],
                                            notes=notes, filter=False,
                                            max_lineage=max_lineage))
    return all_variants
