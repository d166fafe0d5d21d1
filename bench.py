#!/usr/bin/env python
"""bench.py -- k-mers tested/sec of the per-variant association loop on B200.

Workload (BASELINE.json `metric`, configs[3]): LMM, continuous phenotype, N = 5000 samples,
50 M synthetic k-mers sharded by k-mer over 8 GPUs => 6.25 M k-mers per GPU (weak scaling:
`--gpus N` processes N such shards).  One *step* = one pass of the hot path (AF filter,
pre-filter, rotated single-variant LMM test, F-test p-value, lrt filter) over a rank's shard
of packed presence/absence rows.

  value     device-timed, rows already resident in HBM (4 GB per shard, far above the L2)
  e2e       same pass through the C ABI with HOST buffers: psb_submit from pinned host
            memory, psb_run_lmm, psb_fetch of the result table to host -- copies timed
  roofline  the dominant kernel (the rotation / quadratic-form contraction)
  cpu_baseline  the oracle port of the reference's fit_lmm path on the host cores

`--impl reference` times only that CPU path (the oracle port of pyseer's
lmm.fit_lmm -> fit_lmm_block -> fastlmm nLLeval, multiprocessing over blocks of 3000 like
pyseer --cpu N) on the same workload definition.

Launch: `python bench.py --gpus 1 ...` or
`python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`
(torch is used only for the rendezvous, the barrier / max-over-ranks and the NCCL gather of
the result table; the hot path is libpyseer_b200.so through ctypes).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20261017
METRIC = 'kmers_tested_per_sec'
UNIT = 'k-mers/s'
BLOCK = 3000          # pyseer --block_size default (__main__.py:243-246)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--model', default='lmm', choices=['lmm', 'fixed', 'fixed-cont', 'burden'],
                    help="lmm: BASELINE configs[3] (headline); fixed: configs[2], logistic + Firth, "
                         "N=2000, 10 MDS covariates, 10M k-mers")
    ap.add_argument('--samples', type=int, default=0, help='0 = the config default')
    ap.add_argument('--kmers-per-gpu', type=int, default=0, help='0 = the config default')
    ap.add_argument('--precision', type=int,
                    default=int(os.environ.get('PYSEER_B200_LMM_PRECISION', '5')),
                    help='0 = FP64 CUDA-core contraction, 3..8 = exact int8-slice tcgen05 path')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--e2e-chunks', type=int, default=16,
                    help='batches per e2e step (copy of batch i+1 overlaps the kernels of batch i)')
    ap.add_argument('--cpu-cores', type=int, default=0, help='0 = all available (max 64)')
    ap.add_argument('--check', type=int, default=2000,
                    help='variants of the shard re-checked against the oracle after the run')
    return ap.parse_args()


# ----------------------------------------------------------------------------------------
# synthetic problem: kinship from Bernoulli genotypes, heritable phenotype
# ----------------------------------------------------------------------------------------
def make_problem(n):
    """Returns X (N,1 ones), y (continuous, h2 ~ 0.5), normalised kinship K."""
    rng = np.random.RandomState(SEED % (2 ** 31))
    m = 2 * n
    af = rng.uniform(0.05, 0.95, m)
    G = (rng.uniform(size=(n, m)) < af).astype(np.float32)
    K = (G @ G.T).astype(np.float64)
    g = G.astype(np.float64) @ rng.normal(size=m)
    g = (g - g.mean()) / g.std()
    y = math.sqrt(0.5) * g + math.sqrt(0.5) * rng.normal(size=n)
    K *= float(n) / np.diag(K).sum()          # lmm.py:107-112
    return np.ones((n, 1)), y, K


def spectral_state(X, y, K):
    """Once-per-run host set-up of lmm.initialise_lmm: projection, eigh, h2 search."""
    from pyseer_b200.lmm import KinshipLMM
    m = KinshipLMM(X, y.reshape(-1, 1), K, device=0)
    res = m.findH2()
    S, U = m.getSU()
    return np.ascontiguousarray(U), np.ascontiguousarray(S), float(res['h2'])


def make_fixed_problem(n, dims=10):
    """configs[2]: binary phenotype with population structure carried by 10 MDS components
    scaled as input.py:135-136."""
    rng = np.random.RandomState(SEED % (2 ** 31) + 2)
    m = rng.uniform(-1, 1, size=(n, dims))
    m = m / np.abs(m).max(0)
    lin = m[:, :3].sum(1) + rng.normal(size=n)
    y = (lin > np.median(lin)).astype(float)
    return m, y


def _cpu_fixed_block(b):
    from oracle import fixed_oracle as fo
    from pyseer_b200.engine import unpack_rows
    n = _CPU['n']
    x = unpack_rows(_CPU['blocks'][b], n).astype(float)
    af = x.sum(1) / float(n)
    none = np.empty((0, 0))
    tested = 0
    for s in range(x.shape[0]):
        ok = 0.01 <= af[s] <= 0.99
        o = fo.fixed_effects_regression('k', _CPU['y'] if ok else None, x[s], _CPU['m'], none, af[s],
                                        'p', False, None, 1.0, 1.0, _CPU['null_llf'],
                                        _CPU['null_firth'], [], [], _CPU.get('continuous', False))
        tested += not o.prefilter
    return tested


class CpuFixedPath(object):
    """Oracle port of model.fixed_effects_regression, one variant per task as pyseer's
    starmap does (__main__.py:777-780), `cores` workers."""

    def __init__(self, n, m, y, null_llf, null_firth, cores, ys, per_block=150, continuous=False):
        import multiprocessing as mp
        from pyseer_b200.engine import synth_host
        self.cores = cores
        self.per_block = per_block
        _CPU.update(n=n, y=y, m=m, null_llf=null_llf, null_firth=null_firth, continuous=continuous)
        _CPU['blocks'] = [synth_host(SEED, b * per_block, per_block, n, 0.02, 0.98, 1000, ys)
                          for b in range(cores)]
        self.pool = mp.get_context('fork').Pool(cores, initializer=_cpu_init)

    def step(self):
        t0 = time.perf_counter()
        tested = sum(self.pool.map(_cpu_fixed_block, range(self.cores), chunksize=1))
        return tested, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


# ----------------------------------------------------------------------------------------
# CPU path: oracle port of lmm.fit_lmm over blocks of 3000, one worker per core
# ----------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init():
    try:
        from threadpoolctl import threadpool_limits
        _CPU['limit'] = threadpool_limits(1)       # __main__.py:16-19: BLAS pinned to 1 thread
    except Exception:
        pass


def _cpu_block(b):
    from oracle import lmm_oracle as lo
    from pyseer_b200.engine import unpack_rows
    n = _CPU['n']
    bits = _CPU['blocks'][b]
    x = unpack_rows(bits, n)
    mat = np.ascontiguousarray(x.T, dtype=float)
    nan = float('nan')
    y = _CPU['y']
    variants = []
    af = x.sum(1) / float(n)
    for s in range(x.shape[0]):
        ok = 0.01 <= af[s] <= 0.99
        var = lo.LMM('k%d' % s, 'p' if ok else None, af[s], nan, nan, nan, nan, nan, nan, [], [],
                     set(), True, True)
        variants.append((var, y, x[s].astype(float) if ok else None))
    out = lo.fit_lmm(_CPU['lmm'], _CPU['h2'], variants, mat, False, [], np.empty((0, 0)), True,
                     1.0, 1.0)
    return sum(1 for o in out if not o.prefilter)


def cpu_cores(requested):
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    c = requested if requested > 0 else avail
    return max(1, min(c, avail, 64))


class CpuPath(object):
    """The reference's CPU implementation of the path (oracle port) on `cores` workers."""

    def __init__(self, n, X, y, U, S, h2, cores, ys):
        import multiprocessing as mp
        from oracle import lmm_oracle as lo
        from pyseer_b200.engine import synth_host
        olmm = lo.OracleLMM(X, y.reshape(-1, 1), None)
        olmm.U, olmm.S = U, S
        olmm.getUY()
        self.cores = cores
        _CPU.update(n=n, y=y, lmm=olmm, h2=h2)
        # one block of 3000 k-mers per worker per step (same generator and ids as the GPU shard)
        _CPU['blocks'] = [synth_host(SEED, b * BLOCK, BLOCK, n, 0.02, 0.98, 1000, ys)
                          for b in range(cores)]
        self.pool = mp.get_context('fork').Pool(cores, initializer=_cpu_init)

    def step(self):
        t0 = time.perf_counter()
        tested = sum(self.pool.map(_cpu_block, range(self.cores), chunksize=1))
        return tested, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


# ----------------------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.path = tempfile.mktemp(prefix='psb_clocks_', suffix='.csv')
        self.proc = None
        self.dev = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.dev), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(',')]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                      'sw_power_cap'), f[5:9]):
                    if val == 'Active':
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def visible_device(local_rank):
    cvd = os.environ.get('CUDA_VISIBLE_DEVICES')
    if cvd:
        ids = [t for t in cvd.split(',') if t != '']
        if local_rank < len(ids):
            return ids[local_rank]
    return str(local_rank)


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the
    committed ncu --set full capture of the same workload (profiles/ncu_traffic.json), else None."""
    try:
        d = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        return d[key]['traffic_bytes_per_launch']
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


# ----------------------------------------------------------------------------------------
# workloads: what differs between the LMM headline config and the fixed-effects config
# ----------------------------------------------------------------------------------------
class LmmWorkload(object):
    name = 'lmm'
    stats = None
    default_n, default_kpg = 5000, 6250000
    continuous = True

    def __init__(self, a, n, kpg, world):
        self.a, self.n, self.kpg = a, n, kpg
        self.config = {
            'workload': 'LMM continuous phenotype, N=%d samples, %d synthetic k-mers per GPU '
                        '(BASELINE configs[3]: 50M k-mers x 5000 samples sharded by k-mer over 8 '
                        'GPUs), similarity kinship, D=1' % (n, kpg),
            'n_samples': n, 'kmers_per_gpu': kpg, 'kmers_total': kpg * world,
            'af': 'U(0.02,0.98), 0.1% planted causal',
            'filters': 'min_af 0.01 max_af 0.99 filter_pvalue 1 lrt_pvalue 1', 'block_size_cpu': BLOCK,
            'cache': 'inputs (%.2f GB packed rows per GPU) larger than L2'
                     % (kpg * ((n + 127) // 128 * 16) / 1e9)}

    def build_state(self):
        X, y, K = make_problem(self.n)
        U, S, h2 = spectral_state(X, y, K)
        return {'U': U, 'S': S, 'y': y, 'meta': np.array([h2])}

    def state_shapes(self):
        n = self.n
        return {'U': (n, n - 1), 'S': (n - 1,), 'y': (n,), 'meta': (1,)}

    def y_sign(self, st):
        return np.where(st['y'] > np.median(st['y']), 1, -1).astype(np.int8)

    def cpu_path(self, st, cores, ys):
        return CpuPath(self.n, np.ones((self.n, 1)), st['y'], st['U'], st['S'], float(st['meta'][0]),
                       cores, ys), \
            ('blocks of %d k-mers (one block per worker, BLAS pinned to 1 thread as pyseer does), '
             'oracle/lmm_oracle.fit_lmm' % BLOCK)

    def setup_engine(self, eng, st):
        eng.lmm_setup(np.ones((self.n, 1)), st['y'], st['U'], st['S'], float(st['meta'][0]),
                      self.a.precision)

    def run(self, eng):
        eng.run_lmm(min_af=0.01, max_af=0.99, max_missing=0.05, filter_pvalue=1.0, lrt_pvalue=1.0,
                    continuous=True)

    def dtype(self):
        k = self.a.precision
        return ('s8 x%d slices -> s32 (tcgen05) -> f64' % k) if k else 'f64'

    def check(self, st, bits_head, cols):
        from oracle import lmm_oracle as lo
        from pyseer_b200.engine import unpack_rows
        olmm = lo.OracleLMM(np.ones((self.n, 1)), st['y'].reshape(-1, 1), None)
        olmm.U, olmm.S = st['U'], st['S']
        x = unpack_rows(bits_head, self.n)
        ref = lo.fit_lmm_block(olmm, float(st['meta'][0]), np.ascontiguousarray(x.T, dtype=float))
        pv, be = cols['pvalue'], cols['beta']
        ok = np.isfinite(pv) & (ref['p_values'] > 1e-290)
        return {'variants': int(ok.sum()),
                'max_rel_err_pvalue': float(np.max(np.abs(pv[ok] / ref['p_values'][ok] - 1))),
                'max_rel_err_beta': float(np.max(np.abs(be[ok] / ref['beta'][ok] - 1))),
                'min_pvalue': float(np.min(pv[ok]))}

    def roofline(self, tested, k_ms, run_ms, W, pk, pk_kind):
        n, J, k = self.n, self.n - 1, self.a.precision
        achieved = 2.0 * n * J * tested / (k_ms / 1e3) / 1e12     # 2 N (N-D) flop per tested k-mer
        peak = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
        tri = os.environ.get('PSB_TC_TRI', '1') != '0'
        # int8 multiply-adds the kernel issues per k-mer: k slices over N^2/2 (triangular form,
        # K stages of 256 samples from the diagonal down) or N (N-D) entries
        kp = (n + 255) // 256 * 256
        if tri:
            macs = sum(32 * (kp - (32 * jt) // 256 * 256) for jt in range((n + 31) // 32)) * k
        else:
            macs = kp * ((J + 31) // 32 * 32) * k
        return {'bound': 'tensor', 'kernel': 'k_lmm_quadform_tc' if k else 'k_lmm_quadform_fp64',
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                'peak_source': ('%s bf16 dense sustained GEMM (MEASURED_PEAKS.json).  achieved = '
                                'algorithmic fp64-equivalent flops 2 N (N-D) per tested k-mer; the '
                                'kernel executes them as %d exact int8 slices on kind::i8 (nominal '
                                '2x the bf16 rate)%s, so frac ~ %d/%d is the ceiling'
                                % (pk_kind, k, ' over half the index space (triangular form x\'Mx)'
                                   if tri else '', 4 if tri else 2, k)) if k else
                               '%s bf16 dense sustained; FP64 CUDA-core kernel' % pk_kind,
                'algorithmic_flops_per_kmer': 2.0 * n * J,
                'executed_int8_tops': (2.0 * macs * tested / (k_ms / 1e3) / 1e12) if k else None,
                'kernel_ms': k_ms, 'run_ms': run_ms, 'kernel_share_of_step': k_ms / run_ms,
                'hbm_read_frac': (tested * (W * 4 + 56) / (k_ms / 1e3) / 1e9) / pk['hbm_gbs'],
                'algorithmic_bytes': tested * (W * 4 + 24.0),
                'traffic': ncu_traffic('lmm:n=%d:kmers=%d:k=%d' % (n, self.kpg, k))}


class BurdenWorkload(LmmWorkload):
    """BASELINE configs[4]: VCF burden test, 100k regions x 10000 samples, LMM, sharded by region
    over 8 GPUs.  A region is the union of 1-20 rare variant rows (af ~ U(0.001, 0.02), dominant
    encoding, input.py:395-407); the union runs on the device (psb_submit_burden*), then the region
    rows take the LMM path.  `kpg` counts REGIONS per GPU; the metric's unit is regions tested/s."""
    name = 'burden'
    default_n, default_kpg = 10000, 12500

    def __init__(self, a, n, kpg, world):
        LmmWorkload.__init__(self, a, n, kpg, world)
        self.config.update({
            'workload': 'VCF burden test, LMM continuous phenotype, N=%d samples, %d burden regions per GPU, '
                        'each the union of 1-20 rare variant rows (BASELINE configs[4]: 100k regions x '
                        '10000 samples over 8 GPUs); unit = regions' % (n, kpg),
            'af': 'member rows U(0.001,0.02); regions = OR of 1..20 members',
            'cache': 'member rows (%.2f GB) + region rows per GPU'
                     % (kpg * 10.5 * ((n + 127) // 128 * 16) / 1e9)})

    def regions(self, rank):
        rng = np.random.RandomState(SEED % (2 ** 31) + 17 + rank)
        sizes = rng.randint(1, 21, size=self.kpg)
        offs = np.zeros(self.kpg + 1, dtype=np.int64)
        offs[1:] = np.cumsum(sizes)
        return offs, np.arange(int(offs[-1]), dtype=np.int32)     # members of region r are contiguous

    def prepare(self, eng, rank, ys):
        self.offs, self.mem = self.regions(rank)
        self.n_rec = int(self.offs[-1])
        eng.synth_device(SEED + 5, rank * 21 * self.kpg, self.n_rec, 0.001, 0.02, 0, None)
        self.rec_ptr, _, _, self.rec_w = eng.submitted_device()

    def run(self, eng):
        eng.event_record(4)
        eng.submit_burden_device(self.rec_ptr, self.n_rec, self.rec_w, self.offs, self.mem)
        eng.event_record(5)
        LmmWorkload.run(self, eng)

    def check(self, st, rec_head, cols):
        from oracle import input_oracle as io
        nreg = int(np.searchsorted(self.offs, rec_head.shape[0], side='right')) - 1
        nreg = min(nreg, 300)
        bits, _ = io.burden_union(rec_head, None, self.offs[:nreg + 1], self.mem[:int(self.offs[nreg])])
        out = LmmWorkload.check(self, st, bits, {k: v[:nreg] for k, v in cols.items()})
        out['regions'] = nreg
        return out

    def roofline(self, tested, k_ms, run_ms, W, pk, pk_kind):
        r = LmmWorkload.roofline(self, tested, k_ms, run_ms, W, pk, pk_kind)
        r['traffic'] = None
        if getattr(self, 'or_ms', None):
            by = (self.n_rec + self.kpg) * W * 4.0 + self.n_rec * 4.0 + self.kpg * 8.0
            r['burden_or'] = {'bound': 'hbm', 'kernel': 'k_burden_or', 'ms': self.or_ms,
                              'algorithmic_bytes': by, 'achieved': by / (self.or_ms / 1e3) / 1e9,
                              'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                              'frac': by / (self.or_ms / 1e3) / 1e9 / pk['hbm_gbs'],
                              'note': 'CUDA events around psb_submit_burden_device (member lists H2D '
                                      'included); member rows read once, region rows written once'}
        return r


class FixedWorkload(object):
    name = 'fixed'
    stats = None
    default_n, default_kpg = 2000, 10000000
    continuous = False
    DIMS = 10

    def __init__(self, a, n, kpg, world):
        self.a, self.n, self.kpg = a, n, kpg
        self.config = {
            'workload': 'fixed-effects logistic + Firth, N=%d samples, %d MDS covariates, %d '
                        'synthetic k-mers per GPU (BASELINE configs[2])' % (n, self.DIMS, kpg),
            'n_samples': n, 'kmers_per_gpu': kpg, 'kmers_total': kpg * world,
            'af': 'U(0.02,0.98), 0.1% planted causal',
            'filters': 'min_af 0.01 max_af 0.99 filter_pvalue 1 lrt_pvalue 1',
            'cache': 'inputs (%.2f GB packed rows per GPU) larger than L2'
                     % (kpg * ((n + 127) // 128 * 16) / 1e9)}

    def build_state(self):
        from oracle import fixed_oracle as fo
        m, y = make_fixed_problem(self.n, self.DIMS)
        none = np.empty((0, 0))
        null = fo.fit_null(y, m, none, False)
        firth = fo.fit_null(y, m, none, False, True)
        return {'m': m, 'y': y, 'meta': np.array([null.llf, firth])}

    def state_shapes(self):
        return {'m': (self.n, self.DIMS), 'y': (self.n,), 'meta': (2,)}

    def y_sign(self, st):
        return np.where(st['y'] > 0.5, 1, -1).astype(np.int8)

    def cpu_path(self, st, cores, ys):
        return CpuFixedPath(self.n, st['m'], st['y'], float(st['meta'][0]), float(st['meta'][1]),
                            cores, ys), \
            'blocks of 150 k-mers per worker, oracle/fixed_oracle.fixed_effects_regression per variant'

    def setup_engine(self, eng, st):
        Z = np.c_[np.ones(self.n), st['m']]
        eng.fixed_setup(Z, st['y'], False, float(st['meta'][0]), float(st['meta'][1]))

    def run(self, eng):
        eng.run_fixed(min_af=0.01, max_af=0.99, max_missing=0.05, filter_pvalue=1.0, lrt_pvalue=1.0,
                      continuous=False)

    def dtype(self):
        return 'f64'

    def check(self, st, bits_head, cols):
        from oracle import fixed_oracle as fo
        from pyseer_b200.engine import unpack_rows
        x = unpack_rows(bits_head[:300], self.n).astype(float)
        none = np.empty((0, 0))
        errp, errb, nf = 0.0, 0.0, 0
        for s in range(x.shape[0]):
            o = fo.fixed_effects_regression('k', st['y'], x[s], st['m'], none, 0.5, 'p', False, None,
                                            1.0, 1.0, float(st['meta'][0]), float(st['meta'][1]),
                                            [], [], False)
            if o.prefilter or not np.isfinite(o.pvalue):
                continue
            errp = max(errp, abs(cols['pvalue'][s] / o.pvalue - 1))
            errb = max(errb, abs(cols['beta'][s] / o.kbeta - 1))
            nf += 'bad-chisq' in o.notes or 'high-bse' in o.notes
        return {'variants': int(x.shape[0]), 'max_rel_err_pvalue': float(errp),
                'max_rel_err_beta': float(errb), 'firth_fits': int(nf)}

    def roofline(self, tested, k_ms, run_ms, W, pk, pk_kind):
        n, p = self.n, self.DIMS + 2
        per_eval = n * (p * (p + 1) / 2 + 2 * p + 30) * 2.0       # X'WX + score + eta, exp/div
        evals = self.stats['newton_evaluations'] if self.stats else 3.0 * tested
        flops = per_eval * evals                                  # measured evaluation count
        achieved = flops / (k_ms / 1e3) / 1e12
        return {'bound': 'tensor', 'kernel': 'k_fixed_logit(+k_fixed_firth)', 'achieved': achieved,
                'peak': 40.0, 'unit': 'TFLOP/s', 'frac': achieved / 40.0,
                'peak_source': 'FP64 CUDA-core pipe, B200 nominal ~40 TFLOP/s (no measured fp64 peak '
                               'in MEASURED_PEAKS.json); the kernel is FP64-pipe bound, not tensor '
                               'or HBM: flops = measured Newton evaluations x N (p(p+1)/2 + 2p + 30) FMA',
                'newton_evaluations_per_variant': evals / max(tested, 1),
                'kernel_ms': k_ms, 'run_ms': run_ms, 'kernel_share_of_step': k_ms / run_ms,
                'hbm_read_frac': (tested * (W * 4 + 8 * (6 + p)) / (k_ms / 1e3) / 1e9) / pk['hbm_gbs'],
                'traffic': None}


class FixedContWorkload(FixedWorkload):
    """Fixed effects with a continuous phenotype: closed-form OLS t-test per variant."""
    name = 'fixed-cont'
    continuous = True

    def __init__(self, a, n, kpg, world):
        FixedWorkload.__init__(self, a, n, kpg, world)
        self.config['workload'] = self.config['workload'].replace('logistic + Firth', 'OLS (continuous)')

    def build_state(self):
        rng = np.random.RandomState(SEED % (2 ** 31) + 3)
        m, _ = make_fixed_problem(self.n, self.DIMS)
        y = m[:, :3].sum(1) + rng.normal(size=self.n)
        return {'m': m, 'y': y, 'meta': np.array([0.0, 0.0])}

    def y_sign(self, st):
        return np.where(st['y'] > np.median(st['y']), 1, -1).astype(np.int8)

    def cpu_path(self, st, cores, ys):
        cp = CpuFixedPath(self.n, st['m'], st['y'], 0.0, 0.0, cores, ys, per_block=1000, continuous=True)
        return cp, 'blocks of 1000 k-mers per worker, oracle/fixed_oracle.fixed_effects_regression (OLS)'

    def setup_engine(self, eng, st):
        eng.fixed_setup(np.c_[np.ones(self.n), st['m']], st['y'], True, 0.0, 0.0)

    def run(self, eng):
        eng.run_fixed(min_af=0.01, max_af=0.99, max_missing=0.05, filter_pvalue=1.0, lrt_pvalue=1.0,
                      continuous=True)

    def check(self, st, bits_head, cols):
        from oracle import fixed_oracle as fo
        from pyseer_b200.engine import unpack_rows
        x = unpack_rows(bits_head[:500], self.n).astype(float)
        none = np.empty((0, 0))
        errp = errb = 0.0
        for s in range(x.shape[0]):
            o = fo.fixed_effects_regression('k', st['y'], x[s], st['m'], none, 0.5, 'p', False, None,
                                            1.0, 1.0, None, 0.0, [], [], True)
            if o.prefilter or not np.isfinite(o.pvalue) or o.pvalue < 1e-290:
                continue
            errp = max(errp, abs(cols['pvalue'][s] / o.pvalue - 1))
            errb = max(errb, abs(cols['beta'][s] / o.kbeta - 1))
        return {'variants': int(x.shape[0]), 'max_rel_err_pvalue': float(errp),
                'max_rel_err_beta': float(errb)}

    def roofline(self, tested, k_ms, run_ms, W, pk, pk_kind):
        q = self.DIMS + 1
        bytes_alg = tested * (W * 4 + 8.0 * (6 + q))
        achieved = bytes_alg / (run_ms / 1e3) / 1e9
        return {'bound': 'hbm', 'kernel': 'k_bitstats + k_lmm_quadform_tc (linear tile) + k_fixed_ols (whole run)', 'achieved': achieved,
                'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': achieved / pk['hbm_gbs'],
                'peak_source': '%s STREAM-style copy bandwidth (MEASURED_PEAKS.json); algorithmic bytes '
                               '= packed row + result row per tested variant' % pk_kind,
                'kernel_ms': k_ms, 'run_ms': run_ms, 'traffic': None}


# ----------------------------------------------------------------------------------------
def main():
    a = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    wcls = {'lmm': LmmWorkload, 'fixed': FixedWorkload, 'fixed-cont': FixedContWorkload,
            'burden': BurdenWorkload}[a.model]
    n = a.samples or wcls.default_n
    kpg = a.kmers_per_gpu or wcls.default_kpg
    wl = wcls(a, n, kpg, world)
    config = wl.config

    if a.impl == 'reference' and rank != 0:
        return 0

    dist = None
    torch = None
    if world > 1 and a.impl == 'b200':
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    # ---- once-per-run state: rank 0 builds it, the other ranks receive it ---------------
    t_setup = time.time()
    # the CPU arm forks worker processes later: keep this process free of a CUDA context until
    # then (host eigh for the once-per-run set-up); otherwise the set-up uses the device eigh
    forks_later = rank == 0 and (a.impl == 'reference' or (world == 1 and not a.no_cpu_baseline))
    if forks_later:
        os.environ['PYSEER_B200_EIGH'] = 'numpy'
    st = wl.build_state() if rank == 0 else None
    if forks_later:
        os.environ.pop('PYSEER_B200_EIGH', None)
    if dist is not None:
        dev = torch.device('cuda', local_rank)
        out = {}
        for key, shape in wl.state_shapes().items():
            if rank == 0:
                t = torch.from_numpy(np.ascontiguousarray(st[key], dtype=np.float64)).to(dev)
            else:
                t = torch.empty(shape, dtype=torch.float64, device=dev)
            dist.broadcast(t, 0)
            out[key] = t.cpu().numpy()
            del t
        st = out
        torch.cuda.empty_cache()
    ys = wl.y_sign(st)
    t_setup = time.time() - t_setup

    # ---- CPU path (before any CUDA context exists in this process: it forks) ------------
    cpu_line = None
    if rank == 0 and (a.impl == 'reference' or (world == 1 and not a.no_cpu_baseline)):
        cores = cpu_cores(a.cpu_cores)
        cp, what = wl.cpu_path(st, cores, ys)
        if a.impl == 'reference':
            W, Kst = max(a.warmup, 0), max(a.steps, 1)
        else:
            W, Kst = 1, 2
        for _ in range(W):
            cp.step()
        tested = 0
        secs = 0.0
        for _ in range(Kst):
            t, s = cp.step()
            tested += t
            secs += s
        cp.close()
        rate = tested / secs
        cpu_line = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                    'sample': '%d steps x %d %s' % (Kst, cores, what)}
        if a.impl == 'reference':
            line = {'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT,
                    'n_gpus': a.gpus, 'steps': Kst, 'warmup': W,
                    'ms_per_step': 1e3 * secs / Kst, 'higher_is_better': True,
                    'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                    'config': config, 'cpu_baseline': cpu_line,
                    'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                            'd2h_bytes_per_step': 0},
                    'gpu_launches': 0, 'setup_s': t_setup}
            print(json.dumps(line))
            return 0

    # ---- GPU engine -------------------------------------------------------------------
    from pyseer_b200.engine import Engine, PinnedBuffer, words_per_row
    eng = Engine(local_rank)
    wl.setup_engine(eng, st)
    if hasattr(wl, 'prepare'):
        wl.prepare(eng, rank, ys)
    else:
        eng.synth_device(SEED, rank * kpg, kpg, 0.02, 0.98, 1000, ys)
    W = words_per_row(n)

    from pyseer_b200 import sharding
    COLS = tuple((name, b) for name, b, _ in sharding.TABLE_COLUMNS)
    row_bytes = sharding.ROW_BYTES
    if dist is not None:
        table = torch.empty(kpg * row_bytes, dtype=torch.uint8, device=dev)
        ptrs = sharding.table_pointers(table.data_ptr(), kpg)

    def barrier():
        if dist is not None:
            dist.barrier()

    def step():
        wl.run(eng)
        if dist is not None:
            # the one collective of the path: gather the per-variant result table on rank 0
            eng.fetch_into(ptrs)
            sharding.gather_tables(table, [kpg] * world, dst=0)
            torch.cuda.synchronize()

    for _ in range(max(a.warmup, 0)):
        step()
    sampler = ClockSampler(visible_device(local_rank)) if rank == 0 else None
    barrier()
    eng.sync()
    if sampler:
        sampler.start()
    l0 = eng.launch_count()
    eng.event_record(0)
    for _ in range(a.steps):
        step()
    eng.event_record(1)
    eng.sync()
    barrier()
    ms = eng.event_elapsed(0, 1)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    counts = eng.counts()
    tested = counts['tested']
    wl.stats = eng.last_stats() if a.model == 'fixed' else None
    # dominant-kernel time, CUDA events on the library stream around the dominant launch of
    # the last timed step (every step launches the same grid on the same rows)
    k_ms = eng.last_ms(1)
    run_ms = eng.last_ms(0)
    if a.model == 'burden':
        wl.or_ms = eng.event_elapsed(4, 5)
    if dist is not None:
        t = torch.tensor([ms, float(tested)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
        tested_all = float(tsum[1])
    else:
        tested_all = float(tested)
    value = tested_all * a.steps / (ms / 1e3)

    # ---- end to end through the C ABI with host buffers -----------------------------------
    e2e = None
    check = None
    if not a.no_e2e and a.model == 'burden':
        # host record rows + member lists -> psb_submit_burden (H2D, device union) -> LMM -> table
        eng.submit_device(wl.rec_ptr, wl.n_rec, wl.rec_w)
        pin = PinnedBuffer((wl.n_rec, W), np.uint32)
        eng.download_bits(pin.array)
        outs = {name: PinnedBuffer((kpg,), {4: np.int32, 8: np.float64}[b] if name != 'flags'
                                   else np.uint32) for name, b in COLS}
        optr = {name: outs[name].array.ctypes.data for name, _ in COLS}

        def e2e_step():
            eng.submit_burden(pin.array, None, wl.offs, wl.mem)
            LmmWorkload.run(wl, eng)
            eng.fetch_into(optr)

        e2e_step()
        barrier()
        eng.sync()
        eng.event_record(2)
        for _ in range(a.steps):
            e2e_step()
        eng.event_record(3)
        eng.sync()
        barrier()
        ems = eng.event_elapsed(2, 3)
        if dist is not None:
            t = torch.tensor([ems], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t[0])
        e2e = {'value': tested_all * a.steps / (ems / 1e3), 'unit': UNIT,
               'h2d_bytes_per_step': int(wl.n_rec * W * 4 + wl.mem.nbytes + wl.offs.nbytes),
               'd2h_bytes_per_step': int(kpg * row_bytes), 'ms_per_step': ems / a.steps,
               'chunks_per_step': 1}
        if rank == 0 and a.check > 0:
            nrec = int(wl.offs[min(300, kpg)])
            cols = {name: outs[name].array[:300].copy() for name in ('pvalue', 'beta')}
            check = wl.check(st, pin.array[:nrec].copy(), cols)
    elif not a.no_e2e:
        pin = PinnedBuffer((kpg, W), np.uint32)
        eng.download_bits(pin.array)
        outs = {name: PinnedBuffer((kpg,), {4: np.int32, 8: np.float64}[b] if name != 'flags'
                                   else np.uint32) for name, b in COLS}
        optr = {name: outs[name].array.ctypes.data for name, _ in COLS}

        # One step = the whole shard through the public calls a user makes, in `chunks`
        # batches: psb_submit copies batch i+1 (copy stream, second staging slot) while the
        # kernels of batch i run; psb_fetch of batch i then brings its rows of the table back.
        chunks = max(1, a.e2e_chunks)
        bounds = [(kpg * i // chunks, kpg * (i + 1) // chunks) for i in range(chunks)]

        def ptrs_at(lo):
            return {name: outs[name].array[lo:].ctypes.data for name, _ in COLS}

        def e2e_step():
            lo, hi = bounds[0]
            eng.submit(pin.array[lo:hi])
            wl.run(eng)
            for i in range(1, chunks):
                nlo, nhi = bounds[i]
                eng.submit(pin.array[nlo:nhi])
                eng.fetch_into(ptrs_at(lo))
                wl.run(eng)
                lo, hi = nlo, nhi
            eng.fetch_into(ptrs_at(lo))

        e2e_step()
        barrier()
        eng.sync()
        eng.event_record(2)
        for _ in range(a.steps):
            e2e_step()
        eng.event_record(3)
        eng.sync()
        barrier()
        ems = eng.event_elapsed(2, 3)
        if dist is not None:
            t = torch.tensor([ems], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t[0])
        e2e = {'value': tested_all * a.steps / (ems / 1e3), 'unit': UNIT,
               'h2d_bytes_per_step': int(kpg * W * 4), 'd2h_bytes_per_step': int(kpg * row_bytes),
               'ms_per_step': ems / a.steps, 'chunks_per_step': chunks}
        # spot check of the timed output against the oracle (not timed)
        if rank == 0 and a.check > 0:
            cols = {name: outs[name].array[:a.check].copy() for name in ('pvalue', 'beta')}
            check = wl.check(st, pin.array[:a.check].copy(), cols)

    if rank == 0:
        pk, pk_kind = peaks()
        roof = wl.roofline(tested, k_ms, run_ms, W, pk, pk_kind)
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps,
                'warmup': a.warmup, 'ms_per_step': ms / a.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': wl.dtype(),
                'data': 'synthetic', 'config': config, 'clocks': clocks, 'e2e': e2e,
                'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cpu_line,
                'counts': counts, 'check': check, 'setup_s': t_setup}
        if a.model == 'lmm':
            line['h2'] = float(st['meta'][0])
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
