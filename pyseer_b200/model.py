"""Fixed-effects (SEER) model front-end with the reference's interface (pyseer/model.py).

``fit_null`` and ``fixed_effects_regression`` keep the reference's signatures, return types and
error behaviour; ``run_fixed_bits`` is the batched entry the CLI loop and the benchmarks use
(thousands of packed variants per call).  All arithmetic runs in ``libpyseer_b200.so``.
"""
import sys

import numpy as np

from . import classes as var_obj
from . import _lib
from .engine import Engine, pack_rows, notes_from_flags


class NullFit(object):
    """The slice of a statsmodels results object that pyseer reads from ``fit_null``:
    ``params``, ``bse``, ``llf`` (model.py:118-137, __main__.py:430, :450)."""

    def __init__(self, params, bse, llf):
        self.params = params
        self.bse = bse
        self.llf = llf


def _design(p, m, cov):
    """Null design [1, m, cov] (model.py:97-101)."""
    v = np.ones(p.shape[0]).reshape(-1, 1)
    m = np.asarray(m)
    if m.ndim == 2 and m.shape[1] > 0:
        v = np.concatenate((v, m), axis=1)
    cov = getattr(cov, 'values', cov)
    cov = np.asarray(cov)
    if cov.ndim == 2 and cov.shape[1] > 0:
        v = np.concatenate((v, cov), axis=1)
    return np.ascontiguousarray(v, dtype=float)


_engines = {}


def _engine(device):
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]


def fit_null(p, m, cov, continuous, firth=False, device=0):
    """model.py:73-148: null model ``y ~ [1, m, cov]``.  Returns a :class:`NullFit`
    (``.llf``, ``.params``, ``.bse``), the Firth log-likelihood (float) when ``firth`` is
    set, or ``None`` when the model cannot be fitted (message on stderr as the reference)."""
    p = np.asarray(p, dtype=float).reshape(-1)
    v = _design(p, m, cov)
    if not (np.all(np.isfinite(v)) and np.all(np.isfinite(p))):
        sys.stderr.write('Missing data error for null model\n')
        return None
    params, bse, llf, st = _engine(device).fit_null(v, p, continuous, firth)
    if st & _lib.F_PERFECT_SEP:
        sys.stderr.write('Perfectly separable data error for null model\n')
        return None
    if st & _lib.F_MATRIX_INV:
        if not continuous and not firth:
            # model.py:132-137: "Null fit with default optimiser may fail, Powell optimizer might
            # work" -- once-per-run host set-up, as the reference does it
            return _powell_null(v, p)
        sys.stderr.write('Matrix inversion error for null model\n')
        return None
    if firth:
        if st & _lib.F_FIRTH_FAIL:
            sys.stderr.write('Firth regression did not converge for null model\n')
            return None
        return llf
    return NullFit(params, bse, llf)


def _powell_null(v, p):
    """``Logit.fit(start_params, method='powell')`` of model.py:135-137 -- statsmodels'
    ``_fit_powell``: scipy's ``fmin_powell`` (xtol = ftol = 1e-4, maxiter 35) on ``-loglike / nobs``
    with the perfect-prediction check as callback.  The singular information matrix that sent the
    Newton fit here leaves the result without standard errors (statsmodels warns and carries on)."""
    from scipy import optimize
    n = v.shape[0]
    q = 2.0 * p - 1.0

    def cdf(x):
        with np.errstate(over='ignore'):
            return 1.0 / (1.0 + np.exp(-x))

    def loglike(b):
        with np.errstate(divide='ignore', over='ignore'):
            return np.sum(np.log(cdf(q * v.dot(b))))

    class _Separated(Exception):
        pass

    def check(b):
        if np.allclose(cdf(v.dot(b)) - p, 0):
            raise _Separated()

    start = np.zeros(v.shape[1])
    start[0] = np.log(np.mean(p) / (1 - np.mean(p)))
    try:
        out = optimize.fmin_powell(lambda b: -loglike(b) / n, start, xtol=1e-4, ftol=1e-4, maxiter=35,
                                   maxfun=None, full_output=1, disp=0, callback=check)
    except _Separated:
        sys.stderr.write('Perfectly separable data error for null model\n')
        return None
    params = np.asarray(out[0], dtype=float).reshape(-1)
    return NullFit(params, np.full(params.shape[0], np.nan), float(loglike(params)))


class FixedModel(object):
    """Per-run state shared by every ``fixed_effects_regression`` call: covariate design,
    phenotype, null log-likelihoods -- resident on the GPU."""

    def __init__(self, p, m, cov, continuous, null_res, null_firth, device=0, lineage=None):
        self.p = np.asarray(p, dtype=float).reshape(-1)
        self.Z = _design(self.p, m, cov)
        self.continuous = bool(continuous)
        null_llf = getattr(null_res, 'llf', null_res)
        self.engine = Engine(device)
        self.engine.fixed_setup(self.Z, self.p, self.continuous,
                                float(null_llf) if null_llf is not None and not continuous else 0.0,
                                float(null_firth) if isinstance(null_firth, float) else 0.0)
        if lineage is not None:
            # (lineage design, covariates) of model.fit_lineage_effect, fitted per variant in
            # one batched launch after the regression (model.py:379-380)
            self.engine.lineage_setup(lineage[0], lineage[1])

    def close(self):
        self.engine.close()


def run_fixed_bits(model, bits, missing, filter_pvalue, lrt_pvalue, min_af=-1.0, max_af=2.0,
                   max_missing=2.0, lineage=False):
    """Batched model.fixed_effects_regression over packed rows -> result table
    (``.lineage``: index of the strongest lineage per variant, -1 = None, when asked)."""
    eng = model.engine
    eng.submit(bits, missing)
    eng.run_fixed(min_af, max_af, max_missing, filter_pvalue, lrt_pvalue, model.continuous)
    r = eng.fetch()
    r.lineage = eng.run_lineage(False) if lineage else None
    return r


def run_fixed_burden(model, vbits, vmiss, region_offsets, members, filter_pvalue, lrt_pvalue,
                     min_af=-1.0, max_af=2.0, max_missing=2.0):
    """Burden test with the fixed-effects model: per-region union of VCF record rows on the
    device (input.py:395-411), then model.fixed_effects_regression for every region."""
    eng = model.engine
    eng.submit_burden(vbits, vmiss, region_offsets, members)
    eng.run_fixed(min_af, max_af, max_missing, filter_pvalue, lrt_pvalue, model.continuous)
    return eng.fetch()


_cache = {'key': None, 'model': None}


def _model_for(p, m, c, continuous, null_res, null_firth, device):
    """One resident FixedModel per (p, m, c) identity, as the worker map passes the same
    objects for every variant of a run."""
    key = (id(p), id(m), id(c), bool(continuous), device)
    if _cache['key'] != key:
        if _cache['model'] is not None:
            _cache['model'].close()
        _cache['model'] = FixedModel(p, m, c, continuous, null_res, null_firth, device)
        _cache['key'] = key
        _cache['ref'] = (p, m, c)      # keep the ids alive
    return _cache['model']


def fixed_effects_regression(variant, p, k, m, c, af, pattern, lineage_effects, lin, pret, lrtt,
                             null_res, null_firth, kstrains, nkstrains, continuous, device=0):
    """model.py:202-394 -- same arguments, same :class:`Seer` back (a batch of one)."""
    notes = set()
    nan = np.nan
    if p is None:
        notes.add('af-filter')
        return var_obj.Seer(variant, pattern, af, nan, nan, nan, nan, nan, np.array([]), None,
                            kstrains, nkstrains, notes, True, False)
    model = _model_for(p, m, c, continuous, null_res, null_firth, device)
    bits, miss = pack_rows(np.asarray(k).reshape(1, -1))
    r = run_fixed_bits(model, bits, miss, pret, lrtt)
    return seer_from_row(r, 0, variant, pattern, af, kstrains, nkstrains, lineage_effects, lin, c, k,
                         device)


def seer_from_row(r, j, variant, pattern, af, kstrains, nkstrains, lineage_effects=False, lin=None,
                  c=None, k=None, device=0):
    """Result-table row -> Seer tuple with the reference's conventions."""
    nan = np.nan
    f = int(r.flags[j])
    notes = notes_from_flags(f)
    if f & _lib.F_PREFILTER:
        return var_obj.Seer(variant, pattern, af, r.prep[j], nan, nan, nan, nan, np.array([]), None,
                            kstrains, nkstrains, notes, True, False)
    if f & (_lib.F_FIRTH_FAIL | _lib.F_MISSING_DATA):
        return var_obj.Seer(variant, pattern, af, r.prep[j], nan, nan, nan, nan, np.array([]), None,
                            kstrains, nkstrains, notes, False, True)
    max_lineage = None
    if lineage_effects:
        max_lineage = fit_lineage_effect(lin, c, k, device=device)
    return var_obj.Seer(variant, pattern, af, r.prep[j], r.pvalue[j], r.beta[j], r.bse[j],
                        r.extra[j], np.array(r.betas[j]), max_lineage, kstrains, nkstrains, notes,
                        False, bool(f & _lib.F_FILTER))


def fit_lineage_effect(lin, c, k, device=0):
    """model.py:151-199: Logit ``k ~ [1, lin, c]``, index of the lineage column with the
    largest Wald statistic, or None when the fit fails."""
    lin = np.asarray(lin, dtype=float)
    k = np.asarray(k, dtype=float).reshape(-1)
    c = np.asarray(getattr(c, 'values', c)) if c is not None else np.empty((0, 0))
    cols = [np.ones((lin.shape[0], 1)), lin]
    if c.ndim == 2 and c.shape[0] == lin.shape[0]:
        cols.append(c)
    X = np.ascontiguousarray(np.concatenate(cols, axis=1), dtype=float)
    if not (np.all(np.isfinite(X)) and np.all(np.isfinite(k))):
        return None
    params, bse, llf, st = _engine(device).fit_null(X, k, False, False, start_zero=True)
    if st & (_lib.F_PERFECT_SEP | _lib.F_MATRIX_INV):
        return None
    with np.errstate(all='ignore'):
        wald = np.divide(np.absolute(params), bse)
    return int(np.argmax(wald[1:lin.shape[1] + 1]))
