#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference sources of the path, copied from
``/root/reference`` where they lie.  TEST / MEASUREMENT INFRASTRUCTURE.

pyseer is pure Python, so "building" the reference means copying the modules of the path (no edits)
into ``oracle/_ref/pyseer/``: ``lmm.py`` (fit_lmm, fit_lmm_block), ``model.py`` (pre_filtering),
``input.py`` (read_variant, the text parser), ``classes.py``, ``utils.py``, ``cmdscale.py`` and the
vendored FaST-LMM slice ``fastlmm/{lmm_cov,mingrid,util}.py``.  ``oracle/_ref/`` is git-ignored (the
reference's sources never enter the history) but travels to the GPU box with the snapshot, like the
built ``.so``.  Run by ``__graft_entry__.build()`` when ``/root/reference`` exists (this container);
a no-op elsewhere.  ``oracle/ref_loader.py`` imports the copy with stand-ins for the two third-party
packages that are not in the image (statsmodels, pysam): the LMM path never calls into either.
"""
import os
import shutil
import sys

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
FILES = ['pyseer/__init__.py', 'pyseer/classes.py', 'pyseer/utils.py', 'pyseer/cmdscale.py',
         'pyseer/lmm.py', 'pyseer/model.py', 'pyseer/input.py',
         'pyseer/fastlmm/__init__.py', 'pyseer/fastlmm/lmm_cov.py', 'pyseer/fastlmm/mingrid.py',
         'pyseer/fastlmm/util.py', 'pyseer/fastlmm/LICENSE.md', 'pyseer/fastlmm/AUTHORS.txt', 'LICENSE']


def build(quiet=False):
    if not os.path.isdir(os.path.join(REF, 'pyseer')):
        if not quiet:
            sys.stderr.write('oracle/build_ref.py: %s not present, nothing to do\n' % REF)
        return False
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return True


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
