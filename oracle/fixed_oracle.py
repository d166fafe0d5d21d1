"""Oracle: fixed-effects (SEER) model.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

NumPy restatement of ``/root/reference/pyseer/model.py`` and of the statsmodels
calls it makes (statsmodels is a PyPI dependency that is absent from the image:
``statsmodels>=0.10.0``, requirements.txt:12).  Each function cites the reference
lines it follows.  statsmodels behaviour restated here (published algorithm):

* ``Logit.fit(start_params, method='newton')`` -- ``base/optimizer.py:_fit_newton``:
  ``while it < 35 and any(|new-old| > 1e-8)``: ``H = hessian/n; H[diag] += 1e-10;
  new = old - solve(H, score/n)`` (newton gets the *positive* score and hessian); callback ``_check_perfect_pred`` raises
  ``PerfectSeparationError`` iff ``allclose(cdf(X new) - y, 0)``.
  ``llf = sum(log cdf((2y-1) X b))``; ``bse = sqrt(diag(inv(X' W X)))`` at the
  final parameters.
* ``OLS.fit()`` (pinv): ``b = pinv(X) y``; ``df = n - rank(X)``;
  ``bse = sqrt(diag(pinv pinv') * ssr/df)``; ``p = 2 t.sf(|b/bse|, df)``.
"""
from collections import namedtuple

import math
import numpy as np
from scipy import stats

# pyseer/classes.py:15-22
Seer = namedtuple('Seer', ['kmer', 'pattern', 'af', 'prep', 'pvalue', 'kbeta', 'bse',
                           'intercept', 'betas', 'max_lineage', 'kstrains',
                           'nkstrains', 'notes', 'prefilter', 'filter'])


class PerfectSeparationError(Exception):
    pass


class MissingDataError(Exception):
    pass


# --------------------------------------------------------------------------
# statsmodels Logit pieces (discrete_model.py: Logit.cdf/loglike/score/hessian)
# --------------------------------------------------------------------------
def _cdf(x):
    with np.errstate(over='ignore'):
        return 1.0 / (1.0 + np.exp(-x))


def logit_loglike(beta, X, y):
    q = 2.0 * y - 1.0
    with np.errstate(divide='ignore', over='ignore'):
        return np.sum(np.log(_cdf(q * X.dot(beta))))


def logit_hessian(beta, X):
    L = _cdf(X.dot(beta))
    return -np.dot(L * (1.0 - L) * X.T, X)


def logit_score(beta, X, y):
    return X.T.dot(y - _cdf(X.dot(beta)))


LogitRes = namedtuple('LogitRes', ['params', 'bse', 'llf', 'iterations', 'converged'])


def _check_design(X, y):
    # statsmodels raises MissingDataError at model construction for nan/inf
    if not (np.all(np.isfinite(X)) and np.all(np.isfinite(y))):
        raise MissingDataError('exog contains inf or nans')


def logit_newton(X, y, start, maxiter=35, tol=1e-8, raise_perfect=True):
    """statsmodels ``Logit(y, X).fit(start_params=start, method='newton')``.

    Call sites: model.py:129-131 (null), :188 (lineage), :328-330 (variant)."""
    _check_design(X, y)
    n = X.shape[0]
    new = np.asarray(start, dtype=float).copy()
    old = np.full_like(new, np.inf)
    it = 0
    while it < maxiter and np.any(np.abs(new - old) > tol):
        # statsmodels base/model.py:fit hands _fit_newton score = +score/n and hess =
        # +hessian/n (negative definite; "TODO: why are score and hess positive?"), and
        # base/optimizer.py:_fit_newton adds ridge_factor = 1e-10 to that diagonal: the step
        # is (X'WX/n - 1e-10 I)^-1 score/n.  In separated data the matrix goes indefinite and
        # the iterates run away, which is what trips the perfect-prediction check (pinned by
        # the reference's baseline 18: three separated variants report lineage NA).
        H = logit_hessian(new, X) / n
        H[np.diag_indices(H.shape[0])] += 1e-10
        old = new
        new = old - np.linalg.solve(H, logit_score(old, X, y) / n)
        if raise_perfect and np.allclose(_cdf(X.dot(new)) - y, 0):
            raise PerfectSeparationError()
        it += 1
    Hf = logit_hessian(new, X) / n
    cov = np.linalg.inv(-Hf) / n
    with np.errstate(invalid='ignore'):
        bse = np.sqrt(np.diag(cov))
    return LogitRes(new, bse, logit_loglike(new, X, y), it, it < maxiter)


def logit_powell(X, y, start, maxiter=35, raise_perfect=True):
    """statsmodels ``Logit(y, X).fit(start_params=start, method='powell')`` -- base/optimizer.py:
    _fit_powell: ``scipy.optimize.fmin_powell(f, start, xtol=1e-4, ftol=1e-4, maxiter=maxiter,
    maxfun=None, callback=...)`` on ``f = -loglike / nobs`` (maxiter is DiscreteModel.fit's default
    35); the perfect-prediction callback runs after every Powell iteration.  Afterwards
    base/model.py:fit inverts the hessian through ``eigh`` and, when it is not positive definite,
    warns and leaves the result without bse (NaN here).  Call site: model.py:132-137."""
    from scipy import optimize
    _check_design(X, y)
    n = X.shape[0]

    def f(b):
        return -logit_loglike(b, X, y) / n

    def cb(b):
        if raise_perfect and np.allclose(_cdf(X.dot(b)) - y, 0):
            raise PerfectSeparationError()

    out = optimize.fmin_powell(f, np.asarray(start, dtype=float), xtol=1e-4, ftol=1e-4,
                               maxiter=maxiter, maxfun=None, full_output=1, disp=0, callback=cb)
    params = np.asarray(out[0], dtype=float).reshape(-1)
    bse = np.full(params.shape[0], np.nan)
    H = -logit_hessian(params, X)
    if np.all(np.isfinite(H)):
        ev, evec = np.linalg.eigh(H)
        if ev.min() > 0:
            bse = np.sqrt(np.diag(evec.dot(np.diag(1.0 / ev)).dot(evec.T)))
    return LogitRes(params, bse, logit_loglike(params, X, y), out[3], out[5] == 0)


OLSRes = namedtuple('OLSRes', ['params', 'bse', 'pvalues', 'df_resid', 'llf'])


def ols_fit(X, y):
    """statsmodels ``OLS(y, X).fit()``; call sites model.py:118, :300-312."""
    _check_design(X, y)
    n = X.shape[0]
    pinv = np.linalg.pinv(X)
    params = pinv.dot(y)
    ncp = pinv.dot(pinv.T)
    rank = np.linalg.matrix_rank(X)
    df = n - rank
    resid = y - X.dot(params)
    ssr = resid.dot(resid)
    with np.errstate(divide='ignore', invalid='ignore'):
        bse = np.sqrt(np.diag(ncp) * ssr / df)
        tv = params / bse
    pvalues = 2.0 * stats.t.sf(np.abs(tv), df)
    llf = -n / 2.0 * (np.log(2 * np.pi) + np.log(ssr / n) + 1)
    return OLSRes(params, bse, pvalues, df, llf)


# --------------------------------------------------------------------------
# model.py
# --------------------------------------------------------------------------
def pre_filtering(p, k, continuous):
    """model.py:31-70."""
    bad_chisq = False
    if continuous:
        a = p[k == 1]
        b = p[k == 0]
        n1, n2 = a.shape[0], b.shape[0]
        with np.errstate(all='ignore'):
            if n1 == 0 or n2 == 0:
                return np.nan, bad_chisq
            v1 = a.var(ddof=1) if n1 > 1 else np.nan
            v2 = b.var(ddof=1) if n2 > 1 else np.nan
            vn1, vn2 = v1 / n1, v2 / n2
            df = (vn1 + vn2) ** 2 / (vn1 ** 2 / (n1 - 1) + vn2 ** 2 / (n2 - 1)) \
                if (n1 > 1 and n2 > 1) else np.nan
            if np.isnan(df):
                df = 1.0
            t = (a.mean() - b.mean()) / np.sqrt(vn1 + vn2)
            prep = 2.0 * stats.t.sf(np.abs(t), df)
        return float(prep), bad_chisq
    table = np.array([[np.sum((p == 1) & (k == 1)), np.sum((p == 1) & (k == 0))],
                      [np.sum((p == 0) & (k == 1)), np.sum((p == 0) & (k == 0))]])
    if table[table <= 1].shape[0] > 0 or table[table <= 5].shape[0] > 1:
        bad_chisq = True
    n = table.sum()
    with np.errstate(all='ignore'):
        expected = np.outer(table.sum(1), table.sum(0)) / float(n)
        if n > 0 and np.any(expected == 0):
            # scipy.stats.chi2_contingency raises ValueError here (uncaught in the
            # reference, i.e. pyseer crashes); the restatement reports nan.
            return np.nan, bad_chisq
        chi2 = np.sum((table - expected) ** 2 / expected)
        prep = stats.chi2.sf(chi2, 1)
    return float(prep), bad_chisq


def firth_likelihood(beta, X, y):
    """model.py:397-411."""
    with np.errstate(all='ignore'):
        return -(logit_loglike(beta, X, y) +
                 0.5 * np.log(np.linalg.det(-logit_hessian(beta, X))))


def fit_firth(start_vec, X, y, step_limit=1000, convergence_limit=0.0001):
    """model.py:414-504 (only diag(H) of the hat matrix is formed: :455-462 uses
    nothing else)."""
    betas = [np.asarray(start_vec, dtype=float)]
    i = 0
    for i in range(0, step_limit):
        pi = _cdf(X.dot(betas[i]))
        w = pi * (1 - pi)
        V = np.linalg.pinv(-logit_hessian(betas[i], X))
        # diag( sqrt(W) X V X' sqrt(W) )
        h = w * np.einsum('ij,jk,ik->i', X, V, X)
        U = X.T.dot(y - pi + h * (0.5 - pi))
        new_beta = betas[i] + V.dot(U)
        j = 0
        while firth_likelihood(new_beta, X, y) > firth_likelihood(betas[i], X, y):
            new_beta = betas[i] + 0.5 * (new_beta - betas[i])
            j += 1
            if j > step_limit:
                return None
        betas.append(new_beta)
        if i > 0 and np.linalg.norm(betas[i] - betas[i - 1]) < convergence_limit:
            break
    if np.linalg.norm(betas[i] - betas[i - 1]) >= convergence_limit:
        return None
    fitll = -firth_likelihood(betas[-1], X, y)
    intercept = betas[-1][0]
    if len(betas[-1]) > 1:
        kbeta = betas[-1][1]
        bse = math.sqrt(-logit_hessian(betas[-1], X)[1, 1])
    else:
        kbeta = None
        bse = None
    beta = betas[-1][2:].tolist() if len(betas[-1]) > 2 else None
    return intercept, kbeta, beta, bse, fitll


def null_design(p, m, cov):
    """model.py:97-101."""
    v = np.ones(p.shape[0]).reshape(-1, 1)
    if m.ndim == 2 and m.shape[1] > 0:
        v = np.concatenate((v, m), axis=1)
    cov = np.asarray(cov)
    if cov.ndim == 2 and cov.shape[1] > 0:
        v = np.concatenate((v, cov), axis=1)
    return v


def fit_null(p, m, cov, continuous, firth=False):
    """model.py:73-148.  Returns an OLSRes/LogitRes, a float (firth) or None."""
    v = null_design(p, m, cov)
    try:
        if continuous:
            return ols_fit(v, p)
        start_vec = np.zeros(v.shape[1])
        start_vec[0] = np.log(np.mean(p) / (1 - np.mean(p)))
        if firth:
            firth_res = fit_firth(start_vec, v, p)
            if firth_res is None:
                return None
            return firth_res[4]
        try:
            return logit_newton(v, p, start_vec)
        except np.linalg.LinAlgError:
            # model.py:132-137: "Null fit with default optimiser may fail, Powell optimizer
            # might work"
            return logit_powell(v, p, start_vec)
    except (np.linalg.LinAlgError, PerfectSeparationError, MissingDataError):
        return None


def fit_lineage_effect(lin, c, k):
    """model.py:151-199."""
    c = np.asarray(c)
    if c.ndim == 2 and c.shape[0] == lin.shape[0]:
        X = np.concatenate((np.ones(lin.shape[0]).reshape(-1, 1), lin, c), axis=1)
    else:
        X = np.concatenate((np.ones(lin.shape[0]).reshape(-1, 1), lin), axis=1)
    try:
        res = logit_newton(X, k, np.zeros(X.shape[1]))
        with np.errstate(all='ignore'):
            wald = np.divide(np.absolute(res.params), res.bse)
        return int(np.argmax(wald[1:lin.shape[1] + 1]))
    except (PerfectSeparationError, np.linalg.LinAlgError, MissingDataError):
        return None


def variant_design(p, k, m, c):
    """model.py:274-297."""
    c = np.asarray(c)
    ones = np.ones(p.shape[0]).reshape(-1, 1)
    if m.shape[0] != k.shape[0]:
        if c.ndim == 2 and c.shape[0] == k.shape[0]:
            return np.concatenate((ones, k.reshape(-1, 1), c), axis=1)
        return np.concatenate((ones, k.reshape(-1, 1)), axis=1)
    if c.ndim == 2 and c.shape[0] == m.shape[0]:
        return np.concatenate((ones, k.reshape(-1, 1), m, c), axis=1)
    return np.concatenate((ones, k.reshape(-1, 1), m), axis=1)


def fixed_effects_regression(variant, p, k, m, c, af, pattern, lineage_effects, lin,
                             pret, lrtt, null_res, null_firth, kstrains, nkstrains,
                             continuous):
    """model.py:202-394.  ``null_res`` is the null log-likelihood (binary)."""
    notes = set()
    nan = np.nan
    if p is None:
        notes.add('af-filter')
        return Seer(variant, pattern, af, nan, nan, nan, nan, nan, np.array([]),
                    None, kstrains, nkstrains, notes, True, False)
    prep, bad_chisq = pre_filtering(p, k, continuous)
    if bad_chisq:
        notes.add('bad-chisq')
    if prep > pret or not np.isfinite(prep):
        notes.add('pre-filtering-failed')
        return Seer(variant, pattern, af, prep, nan, nan, nan, nan, np.array([]),
                    None, kstrains, nkstrains, notes, True, False)
    v = variant_design(p, k, m, c)
    try:
        if continuous:
            res = ols_fit(v, p)
            intercept, kbeta, beta = res.params[0], res.params[1], res.params[2:]
            bse = res.bse[1]
            lrt_pvalue = res.pvalues[1]
        else:
            _check_design(v, p)
            start_vec = np.zeros(v.shape[1])
            start_vec[0] = np.log(np.mean(p) / (1 - np.mean(p)))
            if not bad_chisq:
                try:
                    res = logit_newton(v, p, start_vec)
                    # NaN bse (non-PD information) compares False, as in numpy
                    if res.bse[1] > 3:
                        bad_chisq = True
                        notes.add('high-bse')
                    else:
                        lrstat = -2 * (null_res - res.llf)
                        lrt_pvalue = 1
                        if lrstat > 0:
                            lrt_pvalue = stats.chi2.sf(lrstat, 1)
                        intercept, kbeta, beta = res.params[0], res.params[1], res.params[2:]
                        bse = res.bse[1]
                except PerfectSeparationError:
                    bad_chisq = True
                    notes.add('perfectly-separable-data')
                except np.linalg.LinAlgError:
                    bad_chisq = True
                    notes.add('matrix-inversion-error')
            if bad_chisq:
                firth_fit = fit_firth(start_vec, v, p)
                if firth_fit is None:
                    notes.add('firth-fail')
                    return Seer(variant, pattern, af, prep, nan, nan, nan, nan,
                                np.array([]), None, kstrains, nkstrains, notes,
                                False, True)
                intercept, kbeta, beta, bse, fitll = firth_fit
                beta = np.array(beta)
                lrstat = -2 * (null_firth - fitll)
                lrt_pvalue = 1
                if lrstat > 0:
                    lrt_pvalue = stats.chi2.sf(lrstat, 1)
    except MissingDataError:
        notes.add('missing-data-error')
        return Seer(variant, pattern, af, prep, nan, nan, nan, nan, np.array([]),
                    None, kstrains, nkstrains, notes, False, True)
    max_lineage = fit_lineage_effect(lin, c, k) if lineage_effects else None
    if lrt_pvalue > lrtt or not np.isfinite(lrt_pvalue) or not np.isfinite(kbeta):
        notes.add('lrt-filtering-failed')
        return Seer(variant, pattern, af, prep, lrt_pvalue, kbeta, bse, intercept,
                    beta, max_lineage, kstrains, nkstrains, notes, False, True)
    return Seer(variant, pattern, af, prep, lrt_pvalue, kbeta, bse, intercept, beta,
                max_lineage, kstrains, nkstrains, notes, False, False)
