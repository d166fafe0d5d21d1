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
    ap.add_argument('--samples', type=int, default=5000)
    ap.add_argument('--kmers-per-gpu', type=int, default=6250000)
    ap.add_argument('--precision', type=int,
                    default=int(os.environ.get('PYSEER_B200_LMM_PRECISION', '5')),
                    help='0 = FP64 CUDA-core contraction, 3..8 = exact int8-slice tcgen05 path')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
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


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


# ----------------------------------------------------------------------------------------
def main():
    a = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    n = a.samples
    kpg = a.kmers_per_gpu

    config = {'workload': 'LMM continuous phenotype, N=%d samples, %d synthetic k-mers per GPU '
                          '(BASELINE configs[3]: 50M k-mers x 5000 samples sharded by k-mer '
                          'over 8 GPUs), similarity kinship, D=1' % (n, kpg),
              'n_samples': n, 'kmers_per_gpu': kpg, 'kmers_total': kpg * world,
              'af': 'U(0.02,0.98), 0.1% planted causal', 'filters': 'min_af 0.01 max_af 0.99 '
              'filter_pvalue 1 lrt_pvalue 1', 'block_size_cpu': BLOCK,
              'cache': 'inputs (%.2f GB packed rows per GPU) larger than L2'
                       % (kpg * ((n + 127) // 128 * 16) / 1e9)}

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
    if rank == 0:
        X, y, K = make_problem(n)
        U, S, h2 = spectral_state(X, y, K)
        del K
    if dist is not None:
        dev = torch.device('cuda', local_rank)
        meta = torch.zeros(1, dtype=torch.float64, device=dev)
        if rank == 0:
            meta[0] = h2
            tU, tS, ty = (torch.from_numpy(v).to(dev) for v in (U, S, y))
        else:
            tU = torch.empty((n, n - 1), dtype=torch.float64, device=dev)
            tS = torch.empty(n - 1, dtype=torch.float64, device=dev)
            ty = torch.empty(n, dtype=torch.float64, device=dev)
        for t in (meta, tU, tS, ty):
            dist.broadcast(t, 0)
        if rank != 0:
            U, S, y, h2 = tU.cpu().numpy(), tS.cpu().numpy(), ty.cpu().numpy(), float(meta[0])
            X = np.ones((n, 1))
        del tU, tS, ty
        torch.cuda.empty_cache()
    ys = np.where(y > np.median(y), 1, -1).astype(np.int8)
    t_setup = time.time() - t_setup

    # ---- CPU path (before any CUDA context exists in this process: it forks) ------------
    cpu_line = None
    if rank == 0 and (a.impl == 'reference' or (world == 1 and not a.no_cpu_baseline)):
        cores = cpu_cores(a.cpu_cores)
        cp = CpuPath(n, X, y, U, S, h2, cores, ys)
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
                    'sample': '%d steps x %d blocks of %d k-mers (one block per worker, BLAS '
                              'pinned to 1 thread as pyseer does), oracle/lmm_oracle.fit_lmm'
                              % (Kst, cores, BLOCK)}
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
    from pyseer_b200.engine import Engine, PinnedBuffer, words_per_row, unpack_rows
    eng = Engine(local_rank)
    eng.lmm_setup(X, y, U, S, h2, a.precision)
    eng.synth_device(SEED, rank * kpg, kpg, 0.02, 0.98, 1000, ys)
    W = words_per_row(n)
    run_kw = dict(min_af=0.01, max_af=0.99, max_missing=0.05, filter_pvalue=1.0,
                  lrt_pvalue=1.0, continuous=True)

    COLS = (('carriers', 4), ('missing', 4), ('af', 8), ('prep', 8), ('pvalue', 8), ('beta', 8),
            ('bse', 8), ('extra', 8), ('flags', 4))
    row_bytes = sum(b for _, b in COLS)
    gather_buf = None
    if dist is not None:
        table = torch.empty(kpg * row_bytes, dtype=torch.uint8, device=dev)
        ptrs, off = {}, 0
        for name, b in COLS:
            ptrs[name] = table.data_ptr() + off
            off += kpg * b
        if rank == 0:
            gather_buf = [torch.empty_like(table) for _ in range(world)]

    def barrier():
        if dist is not None:
            dist.barrier()

    def step():
        eng.run_lmm(**run_kw)
        if dist is not None:
            # the one collective of the path: gather the per-variant result table on rank 0
            eng.fetch_into(ptrs)
            dist.gather(table, gather_buf, dst=0)
            torch.cuda.synchronize()

    for _ in range(max(a.warmup, 0)):
        step()
    sampler = ClockSampler(visible_device(local_rank)) if rank == 0 else None
    barrier()
    eng.sync()
    if sampler:
        sampler.start()
    l0 = eng.launch_count()
    kern_ms = []
    eng.event_record(0)
    for _ in range(a.steps):
        step()
        kern_ms.append(None)
    eng.event_record(1)
    eng.sync()
    barrier()
    ms = eng.event_elapsed(0, 1)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    counts = eng.counts()
    tested = counts['tested']
    # dominant-kernel time, CUDA events on the library stream around the contraction launch
    # of the last timed step (every step launches the same grid on the same rows)
    k_ms = eng.last_ms(1)
    run_ms = eng.last_ms(0)
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
    if not a.no_e2e:
        pin = PinnedBuffer((kpg, W), np.uint32)
        eng.download_bits(pin.array)
        outs = {name: PinnedBuffer((kpg,), {4: np.int32, 8: np.float64}[b] if name != 'flags'
                                   else np.uint32) for name, b in COLS}
        optr = {name: outs[name].array.ctypes.data for name, _ in COLS}

        def e2e_step():
            eng.submit(pin.array)
            eng.run_lmm(**run_kw)
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
               'h2d_bytes_per_step': int(kpg * W * 4), 'd2h_bytes_per_step': int(kpg * row_bytes),
               'ms_per_step': ems / a.steps}
        pvals_host = outs['pvalue'].array[:a.check].copy()
        beta_host = outs['beta'].array[:a.check].copy()
        bits_head = pin.array[:a.check].copy()

    # ---- spot check of the timed output against the oracle (not timed) ----------------------
    check = None
    if rank == 0 and not a.no_e2e and a.check > 0:
        from oracle import lmm_oracle as lo
        olmm = lo.OracleLMM(X, y.reshape(-1, 1), None)
        olmm.U, olmm.S = U, S
        x = unpack_rows(bits_head, n)
        ref = lo.fit_lmm_block(olmm, h2, np.ascontiguousarray(x.T, dtype=float))
        ok = np.isfinite(pvals_host) & (ref['p_values'] > 1e-290)
        check = {'variants': int(ok.sum()),
                 'max_rel_err_pvalue': float(np.max(np.abs(pvals_host[ok] / ref['p_values'][ok] - 1))),
                 'max_rel_err_beta': float(np.max(np.abs(beta_host[ok] / ref['beta'][ok] - 1))),
                 'min_pvalue': float(np.min(pvals_host[ok]))}

    if rank == 0:
        pk, pk_kind = peaks()
        J = n - 1
        flops_alg = 2.0 * n * J * tested            # per launch: 2 N (N-D) per tested k-mer
        achieved = flops_alg / (k_ms / 1e3) / 1e12
        peak = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
        slices = a.precision
        roof = {'bound': 'tensor', 'kernel': 'k_lmm_quadform_tc' if slices else 'k_lmm_quadform_fp64',
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                'peak_source': '%s bf16 dense sustained (MEASURED_PEAKS.json); the kernel runs '
                               'kind::i8 (nominal 2x bf16) on %d exact slices, so frac <= %.2f'
                               % (pk_kind, slices, 2.0 / slices) if slices else
                               '%s bf16 dense sustained; FP64 CUDA-core kernel' % pk_kind,
                'algorithmic_flops_per_kmer': 2.0 * n * J,
                'executed_int8_tops': achieved * slices if slices else None,
                'kernel_ms': k_ms, 'run_ms': run_ms, 'kernel_share_of_step': k_ms / run_ms,
                'hbm_read_frac': (tested * (W * 4 + 56) / (k_ms / 1e3) / 1e9) / pk['hbm_gbs'],
                'traffic': None}
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps,
                'warmup': a.warmup, 'ms_per_step': ms / a.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None,
                'dtype': ('s8 x%d slices -> s32 (tcgen05) -> f64' % slices) if slices else 'f64',
                'data': 'synthetic', 'config': config, 'clocks': clocks, 'e2e': e2e,
                'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cpu_line,
                'counts': counts, 'h2': h2, 'check': check, 'setup_s': t_setup}
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
