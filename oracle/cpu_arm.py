"""CPU arm of the path: the reference's per-variant loop on the host cores.
TEST / MEASUREMENT INFRASTRUCTURE -- see oracle/__init__.py.

Nothing here loads ``libpyseer_b200.so``: inputs come from the NumPy twin of the synthetic
generator (oracle/synth.py), the once-per-run state from the oracle's ``initialise_lmm`` /
``fit_null``, the per-variant work from

  kind = "reference"  the UNMODIFIED reference modules copied to ``oracle/_ref/`` by
                      ``oracle/build_ref.py``: ``pyseer.lmm.fit_lmm`` (lmm.py:125-226) ->
                      ``model.pre_filtering`` -> ``fit_lmm_block`` -> ``fastlmm.lmm_cov.LMM.nLLeval``
                      (LMM only: the fixed-effects regressions need statsmodels, which is not
                      installed), and ``pyseer.input.load_var_block`` / ``read_variant`` for the
                      text-parser leg;
  kind = "port"       the NumPy restatements ``oracle/lmm_oracle.py`` / ``oracle/fixed_oracle.py``.

Work is spread over worker processes the way ``pyseer --cpu N`` does it (``Pool.starmap`` over
blocks of ``--block_size`` variants for the LMM, ``__main__.py:539-546``; over single variants for
the fixed effects, ``:777-780``), BLAS pinned to one thread per worker (``__main__.py:16-19``).
The model object reaches the workers by fork instead of being pickled into every task as the
reference does (200 MB per task at N=5000): that favours the CPU arm.

Used by ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg) and, through
``python -m oracle.cpu_arm sample ...``, by the parity tests at the BASELINE sizes, which need the
oracle's answer for >= 1e5 sampled variants in seconds rather than minutes.
"""
import argparse
import math
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import synth                                           # noqa: E402

SEED = 20261017
BLOCK = 3000            # pyseer --block_size default (__main__.py:243-246)

# flag bits of include/pyseer_b200.h (restated: this module must not import the product)
NOTE_BITS = {'af-filter': 0x0001, 'pre-filtering-failed': 0x0002, 'bad-chisq': 0x0004,
             'high-bse': 0x0008, 'perfectly-separable-data': 0x0010,
             'matrix-inversion-error': 0x0020, 'firth-fail': 0x0040, 'missing-data-error': 0x0080,
             'lrt-filtering-failed': 0x0100}
F_PREFILTER, F_FILTER = 0x0200, 0x0400


def cpu_cores(requested=0):
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    c = requested if requested > 0 else avail
    return max(1, min(c, avail, 64))


# ----------------------------------------------------------------------------------------
# synthetic problems: benchdata.py, the definitions both arms of bench.py use
# ----------------------------------------------------------------------------------------
from benchdata import lmm_problem, fixed_problem, fixed_cont_problem, burden_regions      # noqa: E402,F401


def lmm_spectral(X, y, K):
    """Once-per-run set-up of lmm.initialise_lmm with the oracle port: (U, S, h2)."""
    from oracle import lmm_oracle as lo
    m = lo.OracleLMM(X, y.reshape(-1, 1), np.array(K, dtype=float))
    res = m.findH2()
    S, U = m.getSU()
    return np.ascontiguousarray(U), np.ascontiguousarray(S), float(res['h2'])



# ----------------------------------------------------------------------------------------
# workers (state inherited by fork)
# ----------------------------------------------------------------------------------------
_W = {}


def _init_worker():
    try:
        from threadpoolctl import threadpool_limits
        _W['limit'] = threadpool_limits(1)              # __main__.py:16-19
    except Exception:
        pass


def _rows(task):
    """Packed rows of one task: ('synth', first, count) or ('burden', first_region, count)."""
    w = _W
    if task[0] == 'synth':
        return synth.synth_rows(w['seed'], task[1], task[2], w['n'], w['af_lo'], w['af_hi'],
                                w['planted'], w['ys'], w['separated'])
    from oracle import input_oracle as io
    offs, mem = w['offs'], w['mem']
    r0, r1 = task[1], task[1] + task[2]
    m0, m1 = int(offs[r0]), int(offs[r1])
    rec = synth.synth_rows(w['seed'], w['rec_first'] + m0, m1 - m0, w['n'], w['af_lo'], w['af_hi'], 0,
                           None, 0)
    bits, _ = io.burden_union(rec, None, offs[r0:r1 + 1] - m0, (mem[m0:m1] - m0).astype(np.int32))
    return bits


def _flags_of(o):
    f = 0
    for note in o.notes:
        f |= NOTE_BITS[note]
    if o.prefilter:
        f |= F_PREFILTER
    if o.filter:
        f |= F_FILTER
    return f


def _lmm_task(task):
    """One block through fit_lmm (lmm.py:125-226): tuples built as load_var_block does
    (input.py:672-704)."""
    w = _W
    n = w['n']
    x = synth.unpack_rows(_rows(task), n)
    S = x.shape[0]
    nan = float('nan')
    carriers = x.sum(1)
    af = carriers / float(n)
    mat = np.zeros((n, S))
    LMM = w['LMM']
    variants = []
    for s in range(S):
        k = x[s].astype(float)
        if af[s] < w['min_af'] or af[s] > w['max_af']:
            pattern = None
        else:
            pattern = 'p'
            mat[:, s] = k
        variants.append((LMM(s, pattern, af[s], nan, nan, nan, nan, nan, nan, [], [], set(), True,
                             True), w['y'], k))
    out = w['fit_lmm'](w['lmm'], w['h2'], variants, mat, False, [], np.empty((0, 0)),
                       w['continuous'], w['filter_pvalue'], w['lrt_pvalue'])
    tested = sum(1 for o in out if not o.prefilter)
    if not w['collect']:
        return tested
    res = np.full((S, 6), nan)
    flags = np.zeros(S, dtype=np.uint32)
    for o in out:
        res[o.kmer] = (o.af, o.prep, o.pvalue, o.kbeta, o.bse, o.frac_h2)
        flags[o.kmer] = _flags_of(o)
    return tested, carriers.astype(np.int32), res, flags


def _fixed_task(task):
    """Variants of one task through fixed_effects_regression (model.py:202-394), one call per
    variant as the reference's starmap does (__main__.py:777-780)."""
    from oracle import fixed_oracle as fo
    w = _W
    n = w['n']
    x = synth.unpack_rows(_rows(task), n).astype(float)
    S = x.shape[0]
    carriers = x.sum(1)
    af = carriers / float(n)
    none = np.empty((0, 0))
    q = w['m'].shape[1]
    tested = 0
    res = np.full((S, 6 + q), np.nan) if w['collect'] else None
    flags = np.zeros(S, dtype=np.uint32)
    for s in range(S):
        ok = w['min_af'] <= af[s] <= w['max_af']
        o = fo.fixed_effects_regression(s, w['y'] if ok else None, x[s], w['m'], none, af[s], 'p',
                                        False, None, w['filter_pvalue'], w['lrt_pvalue'],
                                        w['null_llf'], w['null_firth'], [], [], w['continuous'])
        tested += not o.prefilter
        if w['collect']:
            res[s, :6] = (o.af, o.prep, o.pvalue, o.kbeta, o.bse, o.intercept)
            b = np.asarray(o.betas, dtype=float).reshape(-1)
            if b.shape[0] == q:
                res[s, 6:] = b
            flags[s] = _flags_of(o)
    if not w['collect']:
        return tested
    return tested, carriers.astype(np.int32), res, flags


class CpuArm(object):
    """`cores` worker processes over a list of tasks; see the module docstring."""

    def __init__(self, model, n, state, cores, continuous, af=(0.02, 0.98), planted=1000,
                 separated=0, seed=SEED, min_af=0.01, max_af=0.99, filter_pvalue=1.0,
                 lrt_pvalue=1.0, collect=False, prefer_reference=True, burden=None):
        import multiprocessing as mp
        self.model = model
        self.cores = cores
        y = np.asarray(state['y'], dtype=float)
        ys = np.where(y > (0.5 if (model == 'fixed' and not continuous) else np.median(y)), 1, -1) \
            .astype(np.int8)
        _W.clear()
        _W.update(n=n, y=y, ys=ys, seed=seed, af_lo=af[0], af_hi=af[1], planted=planted,
                  separated=separated, min_af=min_af, max_af=max_af, filter_pvalue=filter_pvalue,
                  lrt_pvalue=lrt_pvalue, continuous=continuous, collect=collect)
        if burden is not None:
            _W.update(offs=burden['offs'], mem=burden['mem'], rec_first=burden['rec_first'])
        self.kind = 'port'
        if model == 'lmm':
            X = np.asarray(state['X'], dtype=float)
            self.what = 'oracle/lmm_oracle.fit_lmm (NumPy restatement)'
            from oracle import lmm_oracle as lo
            from oracle import ref_loader
            if prefer_reference and ref_loader.available():
                rlmm, rcov, rcls = ref_loader.load('lmm', 'fastlmm.lmm_cov', 'classes')
                lmm = rcov.LMM(X=X, Y=y.reshape(-1, 1), G=None, K=None)     # lmm.py:66-70 (cache branch)
                lmm.U, lmm.S = state['U'], state['S']
                _W.update(lmm=lmm, fit_lmm=rlmm.fit_lmm, LMM=rcls.LMM)
                self.kind = 'reference'
                self.what = ('unmodified pyseer.lmm.fit_lmm -> fastlmm.lmm_cov.LMM.nLLeval '
                             '(oracle/_ref, copied from the reference by oracle/build_ref.py)')
            else:
                lmm = lo.OracleLMM(X, y.reshape(-1, 1), None)
                lmm.U, lmm.S = state['U'], state['S']
                lmm.getUY()
                _W.update(lmm=lmm, fit_lmm=lo.fit_lmm, LMM=lo.LMM)
            _W['h2'] = float(state['h2'])
            self._task = _lmm_task
        else:
            self.what = 'oracle/fixed_oracle.fixed_effects_regression (NumPy restatement of model.py + statsmodels)'
            _W.update(m=np.asarray(state['m'], dtype=float), null_llf=state['null_llf'],
                      null_firth=state['null_firth'])
            self._task = _fixed_task
        self.pool = mp.get_context('fork').Pool(cores, initializer=_init_worker) if cores > 1 else None
        if self.pool is None:
            _init_worker()

    def run(self, tasks):
        """Returns (results per task in task order, seconds)."""
        t0 = time.perf_counter()
        if self.pool is None:
            out = [self._task(t) for t in tasks]
        else:
            out = self.pool.map(self._task, tasks, chunksize=1)
        return out, time.perf_counter() - t0

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None


# ----------------------------------------------------------------------------------------
# text-parser leg: the reference's own load_var_block / read_variant on a k-mer text file
# ----------------------------------------------------------------------------------------
def write_kmer_text(path, bits, n, first=0):
    """pyseer's --kmers format: `<kmer> | s1:1 s2:1 ...` (input.py:330-352)."""
    x = synth.unpack_rows(bits, n)
    names = np.array(['s%d' % i for i in range(n)])
    acgt = 'ACGT'
    with open(path, 'w') as f:
        for s in range(x.shape[0]):
            v, km = first + s, []
            for _ in range(31):
                km.append(acgt[v & 3])
                v >>= 2
            f.write(''.join(km) + ' | ' + ' '.join(t + ':1' for t in names[x[s] != 0]) + '\n')


def reference_parser_leg(state, n, n_kmers, continuous=True, seed=SEED, tmpdir=None):
    """Leg (b) of BASELINE.md 3.1: text -> load_var_block (read_variant, AF filter, hash_pattern,
    block matrix; input.py:638-707) -> fit_lmm, single process, all through the unmodified reference.
    Returns {'kmers', 'tested', 'parse_s', 'fit_s'}."""
    import tempfile
    import pandas as pd
    from oracle import ref_loader
    rinput, rlmm, rcov = ref_loader.load('input', 'lmm', 'fastlmm.lmm_cov')
    y = np.asarray(state['y'], dtype=float)
    ys = np.where(y > np.median(y), 1, -1).astype(np.int8)
    bits = synth.synth_rows(seed, 0, n_kmers, n, 0.02, 0.98, 1000, ys, 0)
    import gzip
    import shutil
    fd, path = tempfile.mkstemp(prefix='psb_kmers_', suffix='.txt', dir=tmpdir)
    os.close(fd)
    try:
        write_kmer_text(path, bits, n)
        with open(path, 'rb') as fi, gzip.open(path + '.gz', 'wb', compresslevel=6) as fo:
            shutil.copyfileobj(fi, fo)         # pyseer's default input: gzipped k-mers
        p = pd.Series(y, index=['s%d' % i for i in range(n)])
        lmm = rcov.LMM(X=np.asarray(state['X'], dtype=float), Y=y.reshape(-1, 1), G=None, K=None)
        lmm.U, lmm.S = state['U'], state['S']
        t0 = time.perf_counter()
        tested = 0
        parse_s = 0.0
        infile, sample_order = rinput.open_variant_file('kmers', path + '.gz', None, [], False)
        if True:                                # __main__.py:476, :527-532
            it = rinput.load_var_block('kmers', p, False, None, infile, set(p.index), sample_order,
                                       0.01, 0.99, 0.05, False, BLOCK)
            while True:
                tp = time.perf_counter()
                variants, mat, eof = next(it)
                parse_s += time.perf_counter() - tp
                if variants is None or len(variants) == 0:
                    break
                out = rlmm.fit_lmm(lmm, float(state['h2']), variants, mat, False, [], np.empty((0, 0)),
                                   continuous, 1.0, 1.0)
                tested += sum(1 for o in out if not o.prefilter)
                if eof:
                    break
        total = time.perf_counter() - t0
        infile.close()
    finally:
        os.unlink(path)
        if os.path.exists(path + '.gz'):
            os.unlink(path + '.gz')
    return {'kmers': n_kmers, 'tested': tested, 'parse_s': parse_s, 'fit_s': total - parse_s,
            'total_s': total}


# ----------------------------------------------------------------------------------------
# sampling CLI for the parity tests
# ----------------------------------------------------------------------------------------
def sample(state_npz, out_npz, cores=0):
    """Oracle answers for the tasks listed in ``state_npz`` (written by the test): columns in task
    order -> ``out_npz``.  Always the NumPy restatement (kind 'port'), so that a test compares the
    CUDA path with the oracle proper; tests/test_ref_cpu.py holds the restatement to oracle/_ref."""
    with np.load(state_npz, allow_pickle=False) as d:
        st = {k: d[k] for k in d.files}
    model = str(st['model'])
    n = int(st['n'])
    tasks = [(str(st['task_kind']), int(a), int(b)) for a, b in st['tasks']]
    burden = None
    if 'offs' in st:
        burden = {'offs': st['offs'], 'mem': st['mem'], 'rec_first': int(st['rec_first'])}
    state = {'y': st['y']}
    if model == 'lmm':
        state.update(X=st['X'], U=st['U'], S=st['S'], h2=float(st['h2']))
    else:
        state.update(m=st['m'], null_llf=float(st['null_llf']), null_firth=float(st['null_firth']))
    arm = CpuArm(model, n, state, cpu_cores(cores), bool(st['continuous']),
                 af=(float(st['af_lo']), float(st['af_hi'])), planted=int(st['planted']),
                 separated=int(st['separated']), seed=int(st['seed']), min_af=float(st['min_af']),
                 max_af=float(st['max_af']), filter_pvalue=float(st['filter_pvalue']),
                 lrt_pvalue=float(st['lrt_pvalue']), collect=True, prefer_reference=False,
                 burden=burden)
    out, secs = arm.run(tasks)
    arm.close()
    np.savez(out_npz, tested=np.array([sum(o[0] for o in out)]),
             carriers=np.concatenate([o[1] for o in out]),
             res=np.concatenate([o[2] for o in out]), flags=np.concatenate([o[3] for o in out]),
             seconds=np.array([secs]), cores=np.array([arm.cores]))
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    sub = ap.add_subparsers(dest='cmd', required=True)
    s = sub.add_parser('sample')
    s.add_argument('state')
    s.add_argument('out')
    s.add_argument('--cores', type=int, default=0)
    a = ap.parse_args(argv)
    return sample(a.state, a.out, a.cores)


if __name__ == '__main__':
    sys.exit(main())
