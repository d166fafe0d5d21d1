"""LMM front-end with the reference's interface (pyseer/lmm.py).

``initialise_lmm`` does the once-per-run set-up (lmm.py:26-122): kinship normalisation on the host;
for N >= 2048 the projection and O(N^3) eigendecomposition of setSU_fromK (``psb_spectral``,
fastlmm/lmm_cov.py:88-103) and the O(N) sums of every likelihood evaluation of the h2 search
(``psb_lmm_nll_terms``, lmm_cov.py:427-478, 597-684) on the device, the grid / Brent logic of
mingrid.minimize1D around them on the host;
``fit_lmm`` / ``fit_lmm_block`` keep the reference's signatures and error behaviour but
hand every per-variant computation to the GPU engine.
"""
import math
import os
import sys

import numpy as np
import scipy.optimize as opt

from . import classes as var_obj
from . import _lib
from .engine import Engine, pack_rows, notes_from_flags


class KinshipLMM(object):
    """State of ``pyseer.fastlmm.lmm_cov.LMM`` that pyseer uses: covariates X (last column
    ones), phenotype Y, spectral decomposition (U, S) of the projected kernel, and the h2
    search.  Once-per-run host code; the per-variant work lives in the engine."""

    def __init__(self, X, Y, K=None, device=0, precision=None):
        self.X = np.ascontiguousarray(X, dtype=float)
        self.Y = np.asarray(Y, dtype=float).reshape(self.X.shape[0], -1)
        self.K = K
        self.D = self.X.shape[1]
        self.U = None
        self.S = None
        self._UY = None
        self._uy2 = None
        self._nll_grid = None
        self._Xdagger = None
        self.device = device
        self.precision = precision
        self._engine = None
        self._engine_h2 = None

    # mirrors lmm.linreg.D used by lmm.py:253
    @property
    def linreg(self):
        return self

    def _regress(self, A):
        # Linreg.regress, lmm_cov.py:874-880 (pinv-based projection)
        if self._Xdagger is None:
            self._Xdagger = np.linalg.pinv(self.X)
        return A - self.X.dot(self._Xdagger.dot(A))

    def getSU(self):
        # setSU_fromK, lmm_cov.py:88-103
        if self.U is None or self.S is None:
            if self.K is None:
                raise Exception("No Kernel is set. Cannot return U and S.")
            N = self.K.shape[0]
            S = U = None
            mode = os.environ.get('PYSEER_B200_EIGH', 'auto')
            if (mode == 'device' or (mode != 'numpy' and N >= 2048)) and self.D <= 16:
                # projection P (K + I) P and eigh in one device call (psb_spectral)
                if self._Xdagger is None:
                    self._Xdagger = np.linalg.pinv(self.X)
                if self._engine is None:
                    self._engine = Engine(self.device)
                try:
                    S, U = self._engine.spectral(self.K, self.X, self._Xdagger)
                except _lib.PsbError as e:
                    if e.code != _lib.ERR_UNSUPPORTED:
                        raise
                    sys.stderr.write('cuSOLVER not available, eigendecomposition on the host\n')
            if S is None:
                self.K.flat[::N + 1] += 1.0
                K_ = self._regress(self.K)
                K_ = self._regress(K_.T)
                S, U = self._eigh(K_)
            self.U = np.ascontiguousarray(U[:, self.D:N])
            self.S = S[self.D:N] - 1.0
        return self.S, self.U

    def _eigh(self, K_):
        """The O(N^3) step of setSU_fromK: on the device (psb_eigh, cuSOLVER syevd in fp64) for
        N >= 2048 (below that the host routine takes a fraction of a second, less than loading
        cuSOLVER) unless PYSEER_B200_EIGH=numpy / =device; NumPy when cuSOLVER is not installed.
        Either way the per-variant statistics only depend on the eigenspaces, not on the basis
        chosen inside a degenerate one."""
        mode = os.environ.get('PYSEER_B200_EIGH', 'auto')
        if mode == 'device' or (mode != 'numpy' and K_.shape[0] >= 2048):
            if self._engine is None:
                self._engine = Engine(self.device)
            try:
                return self._engine.eigh(K_)
            except _lib.PsbError as e:
                if e.code != _lib.ERR_UNSUPPORTED:
                    raise
                sys.stderr.write('cuSOLVER not available, eigendecomposition on the host\n')
        return np.linalg.eigh(K_)

    def _getUY(self):
        if self._UY is None:
            S, U = self.getSU()
            A = self._regress(self.Y)
            A[:, A.std(0) <= 1e-10] = 0.0          # rotate(), lmm_cov.py:179-181
            self._UY = U.T.dot(A)
        return self._UY

    def _device_h2(self):
        """The O(N) sums of nLLeval on the device (``psb_lmm_nll_terms``) under the same rule as the
        eigendecomposition: N >= 2048 unless PYSEER_B200_EIGH says otherwise; one phenotype column."""
        mode = os.environ.get('PYSEER_B200_EIGH', 'auto')
        return self.Y.shape[1] == 1 and mode != 'numpy' and \
            (mode == 'device' or self.Y.shape[0] >= 2048)

    def _nll_terms(self, h2s):
        """(YKY [n, P], logdetK [n]) of nLLeval at the given h2 values."""
        S, U = self.getSU()
        UY = self._getUY()
        h2s = np.atleast_1d(np.asarray(h2s, dtype=float))
        if self._device_h2():
            if self._engine is None:
                self._engine = Engine(self.device)
            if self._uy2 is None:
                self._uy2 = np.ascontiguousarray(UY[:, 0] ** 2)
            yky, ld = self._engine.nll_terms(S, self._uy2, h2s)
            return yky.reshape(-1, 1), ld
        with np.errstate(all='ignore'):
            Sd = h2s.reshape(-1, 1) * S.reshape(1, -1) + (1.0 - h2s.reshape(-1, 1))
            YKY = np.stack([(UY / sd.reshape(-1, 1) * UY).sum(0) for sd in Sd])
            return YKY, np.log(Sd).sum(1)

    def nLLeval(self, h2=0.0):
        """Null-model negative log-likelihood at h2 (lmm_cov.py:597-684, 726-727, 817-825)."""
        N = self.Y.shape[0] - self.D
        if h2 < 0.0 or h2 >= 1.0:
            return {'nLL': 3e20, 'h2': h2, 'scale': 1.0}
        terms = self._nll_grid.pop(float(h2), None) if self._nll_grid else None
        if terms is None:
            YKY, logdetK = self._nll_terms([h2])
            terms = (YKY[0], logdetK[0])
        YKY, logdetK = terms
        with np.errstate(all='ignore'):
            sigma2 = YKY / N
            nLL = 0.5 * (logdetK + N * (np.log(2.0 * np.pi * sigma2) + 1))
        return {'nLL': nLL, 'h2': h2, 'scale': 1.0, 'dof': None}

    def findH2(self, nGridH2=10, minH2=0.0, maxH2=0.99999):
        """lmm_cov.py:427-478 + mingrid.minimize1D (mingrid.py:13-73): grid, bounded search on
        boundary minima, Brent on interior triplets; returns the best *evaluated* point."""
        resmin = [None]

        def f(x):
            res = self.nLLeval(h2=x)
            if resmin[0] is None or res['nLL'] < resmin[0]['nLL']:
                resmin[0] = res
            return res['nLL'][0]

        step = (maxH2 - minH2) / nGridH2
        grid = np.arange(minH2, maxH2 + step, step)
        # the grid of the search in one device launch (Brent's points follow one by one)
        inside = [float(x) for x in grid if 0.0 <= x < 1.0]
        if inside and self._device_h2():
            YKY, ld = self._nll_terms(inside)
            self._nll_grid = {x: (YKY[i], ld[i]) for i, x in enumerate(inside)}
        vals = np.array([f(x) for x in grid])
        self._nll_grid = None
        if vals[0] < vals[1]:
            opt.fminbound(f, grid[0], grid[1], full_output=True)
        if vals[-1] < vals[-2]:
            opt.fminbound(f, grid[-2], grid[-1], full_output=True)
        for i in range(vals.shape[0] - 2):
            if vals[i + 1] < vals[i + 2] and vals[i + 1] < vals[i]:
                opt.brent(f, brack=(grid[i], grid[i + 1], grid[i + 2]), full_output=True)
        return resmin[0]

    # -- GPU side ----------------------------------------------------------------------
    def engine(self, h2):
        """Engine with the rotated operands for this h2 resident on the GPU."""
        if h2 < 0.0 or h2 >= 1.0:
            # lmm_cov.py:667-670 returns a dict without 'beta' -> KeyError in fit_lmm_block
            raise KeyError('beta')
        if self._engine is None or self._engine_h2 != h2:
            if self._engine is None:
                self._engine = Engine(self.device)
            S, U = self.getSU()
            prec = self.precision
            if prec is None:
                prec = int(os.environ.get('PYSEER_B200_LMM_PRECISION', '46'))
            self._engine.lmm_setup(self.X, self.Y[:, 0], U, S, h2, prec)
            self._engine_h2 = h2
        return self._engine

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None


def initialise_lmm(p, cov, K_in, lmm_cache_in=None, lmm_cache_out=None, lineage_samples=None,
                   device=0, precision=None):
    """lmm.py:26-122 -- same arguments, same return triple ``(p, lmm, h2)``."""
    import pandas as pd

    def _covar(p, cov):
        if len(p.index.intersection(cov.index)) == p.shape[0]:
            return np.c_[cov.loc[p.index].values, np.ones((p.shape[0], 1))]
        elif (cov.shape[0] == 0 and cov.shape[1] == 0) or len(cov.shape) == 0:
            return np.ones((p.shape[0], 1))
        sys.stderr.write("Phenotype and covariate file should have "
                         "matching samples for LMM\n")
        sys.exit(1)

    if lmm_cache_in is not None and os.path.exists(lmm_cache_in):
        covar = _covar(p, cov)
        y = np.reshape(p.values, (-1, 1))
        lmm = KinshipLMM(covar, y, None, device=device, precision=precision)
        with np.load(lmm_cache_in) as data:
            lmm.U = np.ascontiguousarray(data['arr_0'])
            lmm.S = data['arr_1']
            h2 = data['arr_2'][0]
            if lmm.U.shape[0] != len(p):
                sys.stderr.write("Phenotype different length from cache file\n")
                sys.exit(1)
    else:
        K = pd.read_csv(K_in, index_col=0, sep='\t')
        K.index = K.index.astype(str)
        sys.stderr.write("Similarity matrix has dimension " + str(K.shape) + "\n")
        if lineage_samples is not None and set(K.index) != set(lineage_samples):
            sys.stderr.write("Lineage file and similarity matrix contain different sets"
                             " of samples\n")
            sys.exit(1)
        intersecting_samples = p.index.intersection(K.index)
        sys.stderr.write("Analysing " + str(len(intersecting_samples)) + " samples"
                         " found in both phenotype and similarity matrix\n")
        p = p.loc[intersecting_samples]
        y = np.reshape(p.values, (-1, 1))
        K = K.loc[p.index, p.index]
        covar = _covar(p, cov)
        Kv = np.array(K.values, dtype=float)
        with np.errstate(divide='ignore'):
            factor = float(len(p)) / np.diag(Kv).sum()
        if factor == math.inf:
            sys.stderr.write("Invalid similarity matrix. Did you use --calc-C?\n")
            sys.exit(1)
        elif abs(factor - 1.0) > 1e-15:
            Kv *= factor
        lmm = KinshipLMM(covar, y, Kv, device=device, precision=precision)
        result = lmm.findH2()
        h2 = result['h2']
        if lmm_cache_out is not None and not os.path.exists(lmm_cache_out):
            lmm.getSU()
            np.savez(lmm_cache_out, lmm.U, lmm.S, np.array([h2]))
    return (p, lmm, h2)


_NOFILTER = dict(min_af=-1.0, max_af=2.0, max_missing=2.0)


def fit_lmm_block(lmm, h2, variant_block):
    """lmm.py:228-260: ``{'p_values','beta','bse','frac_h2'}`` for an (N, S) 0/1 block.
    No filtering is applied (as in the reference)."""
    eng = lmm.engine(h2)
    variant_block = np.asarray(variant_block)
    assert variant_block.shape[0] == lmm.Y.shape[0], "shape missmatch between snps and Y"
    bits, miss = pack_rows(variant_block.T)
    eng.submit(bits, miss)
    p = eng._params(-1.0, 2.0, 2.0, np.inf, np.inf, False)
    p.options = _lib.OPT_NO_PREFILTER
    _lib.check(eng.lib.psb_run_lmm(eng._ctx, p))
    eng.n_run = eng.n_variants
    r = eng.fetch(('pvalue', 'beta', 'bse', 'extra', 'flags'))
    return {'p_values': r.pvalue, 'beta': r.beta, 'bse': r.bse, 'frac_h2': r.extra}


def run_lmm_bits(lmm, h2, bits, missing, continuous, filter_pvalue, lrt_pvalue,
                 min_af=-1.0, max_af=2.0, max_missing=2.0):
    """Batched entry used by the CLI and the benchmarks: packed rows in, result table out."""
    eng = lmm.engine(h2)
    eng.submit(bits, missing)
    eng.run_lmm(min_af, max_af, max_missing, filter_pvalue, lrt_pvalue, continuous)
    return eng.fetch()


def run_lmm_burden(lmm, h2, vbits, vmiss, region_offsets, members, continuous, filter_pvalue,
                   lrt_pvalue, min_af=-1.0, max_af=2.0, max_missing=2.0):
    """Burden test (``--vcf --burden --lmm``, input.py:395-411 feeding lmm.fit_lmm): one packed row
    per VCF record plus the member lists of the regions; the per-region union is formed on the
    device and every region goes through the LMM path.  Result table, one row per region."""
    eng = lmm.engine(h2)
    eng.submit_burden(vbits, vmiss, region_offsets, members)
    eng.run_lmm(min_af, max_af, max_missing, filter_pvalue, lrt_pvalue, continuous)
    return eng.fetch()


def fit_lmm(lmm, h2, variants, variant_mat, lineage_effects,
            lineage_clusters, covariates, continuous,
            filter_pvalue, lrt_pvalue):
    """lmm.py:125-226 with the same arguments and the same list of LMM tuples back
    (AF/pre-filtered variants first, fitted variants after, as the reference builds it)."""
    all_variants = []
    keep = []
    k = None
    for var_idx, variant in enumerate(variants):
        var, p, k = variant
        if var.pattern is None or k is None:
            all_variants.append(var._replace(notes=set(['af-filter']), prefilter=True,
                                             filter=False))
            variant_mat[:, var_idx] = 0.0
            continue
        keep.append((var_idx, var))
    if not keep:
        return all_variants
    cols = [i for i, _ in keep]
    bits, miss = pack_rows(np.asarray(variant_mat)[:, cols].T)
    r = run_lmm_bits(lmm, h2, bits, miss, continuous, filter_pvalue, lrt_pvalue)
    tested = []
    for j, (var_idx, var) in enumerate(keep):
        f = int(r.flags[j])
        notes = notes_from_flags(f)
        if f & _lib.F_PREFILTER:
            all_variants.append(var._replace(notes=notes, prep=r.prep[j], prefilter=True,
                                             filter=False))
            variant_mat[:, var_idx] = 0.0
            continue
        tested.append((j, var._replace(prep=r.prep[j], notes=notes, prefilter=False)))
    for j, tv in tested:
        f = int(r.flags[j])
        if f & _lib.F_FILTER:
            all_variants.append(tv._replace(pvalue=r.pvalue[j], filter=True))
        else:
            max_lineage = None
            if lineage_effects:
                from .model import fit_lineage_effect
                # lmm.py:209-211 passes the loop variable `k` left over from the first loop
                # (the block's last variant) -- reproduced on purpose
                max_lineage = fit_lineage_effect(lineage_clusters, covariates, k,
                                                 device=lmm.device)
            all_variants.append(tv._replace(pvalue=r.pvalue[j], kbeta=r.beta[j], bse=r.bse[j],
                                            frac_h2=r.extra[j], filter=False,
                                            max_lineage=max_lineage))
    return all_variants
