"""Imports the unmodified reference modules copied to ``oracle/_ref/`` by ``oracle/build_ref.py``.
TEST / MEASUREMENT INFRASTRUCTURE.

statsmodels and pysam are not installed in this image.  ``pyseer/model.py`` and ``pyseer/lmm.py``
import statsmodels at module level, ``pyseer/input.py`` imports ``pysam.VariantFile``; the LMM path
(``lmm.fit_lmm`` -> ``model.pre_filtering`` -> ``lmm.fit_lmm_block`` -> ``fastlmm.lmm_cov.LMM.nLLeval``)
and the k-mer / Rtab text parser (``input.read_variant``) never call into either, so name-only
stand-ins are installed when the real packages are absent.  Anything that would actually need
statsmodels (the fixed-effects regressions, lineage effects) raises ``ReferenceUnavailable``.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


class ReferenceUnavailable(RuntimeError):
    pass


def _raise(*a, **k):
    raise ReferenceUnavailable('statsmodels / pysam are not installed: this part of the reference '
                               'cannot run here')


def _install_stubs():
    try:
        import statsmodels  # noqa: F401
        import statsmodels.formula.api  # noqa: F401
    except ImportError:
        sm = types.ModuleType('statsmodels')
        formula = types.ModuleType('statsmodels.formula')
        api = types.ModuleType('statsmodels.formula.api')
        tools = types.ModuleType('statsmodels.tools')
        exc = types.ModuleType('statsmodels.tools.sm_exceptions')
        api.OLS = api.Logit = _raise

        class PerfectSeparationError(Exception):
            pass

        class MissingDataError(Exception):
            pass

        exc.PerfectSeparationError = PerfectSeparationError
        exc.MissingDataError = MissingDataError
        sm.formula, formula.api, sm.tools, tools.sm_exceptions = formula, api, tools, exc
        sm.__stub__ = True
        for name, mod in (('statsmodels', sm), ('statsmodels.formula', formula),
                          ('statsmodels.formula.api', api), ('statsmodels.tools', tools),
                          ('statsmodels.tools.sm_exceptions', exc)):
            sys.modules[name] = mod
    try:
        import pysam  # noqa: F401
    except ImportError:
        ps = types.ModuleType('pysam')
        ps.VariantFile = _raise
        ps.__stub__ = True
        sys.modules['pysam'] = ps


def available():
    return os.path.exists(os.path.join(REF_DIR, 'pyseer', 'fastlmm', 'lmm_cov.py'))


def load(*names):
    """Returns the requested modules of the reference copy, e.g. ``load('lmm', 'fastlmm.lmm_cov')``.
    Raises ReferenceUnavailable when ``oracle/_ref`` has not been built."""
    if not available():
        raise ReferenceUnavailable('%s is empty: run `python oracle/build_ref.py` where /root/reference '
                                   'is mounted' % REF_DIR)
    other = sys.modules.get('pyseer')
    if other is not None and not getattr(other, '__file__', '').startswith(REF_DIR):
        raise ReferenceUnavailable('another `pyseer` package is already imported: %r' % other)
    _install_stubs()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    mods = [importlib.import_module('pyseer.' + n) for n in names]
    for m in mods:
        assert m.__file__.startswith(REF_DIR), m.__file__
    return mods[0] if len(mods) == 1 else mods
