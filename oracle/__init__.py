"""CPU oracle for the pyseer per-variant association path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``pyseer_b200/`` may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker or the timed CPU
baseline -- never as the thing shipped.

Contents
  fixed_oracle.py  NumPy restatement of ``pyseer/model.py`` plus the statsmodels
                   behaviour it calls (statsmodels is not installed in this image;
                   pinned >=0.10.0 by requirements.txt:12).
  lmm_oracle.py    NumPy restatement of ``pyseer/lmm.py`` and the single-kernel
                   full-rank slice of ``pyseer/fastlmm/lmm_cov.py`` / ``mingrid.py``.
  gen_golden.py    Script (run in the build container, where /root/reference is
                   mounted) that imports the reference's own ``fastlmm.lmm_cov``
                   and records golden vectors under ``tests/golden/``.

Parity is PINNED: tests/test_oracle_*.py re-assert every golden constant of the
reference's tests/model_test.py and tests/lmm_test.py against these restatements,
and the LMM restatement is additionally checked against vectors produced by the
reference's unmodified lmm_cov code (tests/golden/lmm_ref_*.npz).
"""
