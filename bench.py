#!/usr/bin/env python
"""bench.py -- k-mers tested/sec of the per-variant association loop on B200.

Workload (BASELINE.json `metric`, configs[3]): LMM, continuous phenotype, N = 5000 samples,
50 M synthetic k-mers sharded by k-mer over 8 GPUs => 6.25 M k-mers per GPU (weak scaling:
`--gpus N` processes N such shards).  One *step* = one pass of the hot path (AF filter,
pre-filter, rotated single-variant LMM test, F-test p-value, lrt filter) over a rank's shard
of packed presence/absence rows.

  value        device-timed, rows already resident in HBM (4 GB per shard, far above the L2);
               with N > 1 the NCCL gather of the result table on rank 0 is inside the timed region
  e2e          same pass through the C ABI with HOST buffers: psb_submit from pinned host
               memory, psb_run_lmm, psb_fetch of the result table to host -- copies timed
  roofline     the dominant kernel against the MEASURED peak of the pipe it runs on
               (psb_measure_peaks: int8 tcgen05 MMA rate / fp64 FMA rate, probed right after the
               timed region under the same clocks)
  cpu_baseline the CPU arm (below) on a bounded sample, run in a process of its own
  secondary    (N = 1 default run) the same line for BASELINE configs[1] (LMM, binary phenotype,
               N=1000), configs[2] (fixed effects, logistic + Firth, N=2000, 10 MDS covariates)
               and configs[4] (VCF burden regions, N=10000, LMM)

`--impl reference` is the CPU arm: the reference's own code on the host cores -- for the LMM the
UNMODIFIED pyseer.lmm.fit_lmm -> fastlmm.lmm_cov.LMM.nLLeval from oracle/_ref (copied there from the
reference by oracle/build_ref.py; kind "reference"), else the NumPy restatement (kind "port"; always
for the fixed effects: statsmodels is not installed) -- over blocks of 3000 k-mers on one worker per
core, like pyseer --cpu N.  That arm never loads libpyseer_b200.so: inputs come from the NumPy twin
of the generator (oracle/synth.py), the once-per-run state from the oracle.

Launch: `python bench.py --gpus 1 ...` or
`python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...` (torchrun is only
the launcher: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_PORT are read from the environment; the
rendezvous, barrier, max-over-ranks and the gather go through the library's own NCCL communicator,
pyseer_b200/comm.py -- no torch import).
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
THRESH = dict(min_af=0.01, max_af=0.99, max_missing=0.05, filter_pvalue=1.0, lrt_pvalue=1.0)
MODELS = ('lmm', 'lmm-binary', 'fixed', 'fixed-cont', 'burden')


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--model', default='lmm', choices=MODELS,
                    help='lmm: BASELINE configs[3] (headline); lmm-binary: configs[1] (N=1000, 1M k-mers); '
                         'fixed: configs[2] (logistic + Firth, N=2000, 10 MDS covariates, 10M k-mers); '
                         'burden: configs[4] (N=10000, regions); fixed-cont: OLS')
    ap.add_argument('--samples', type=int, default=0, help='0 = the config default')
    ap.add_argument('--kmers-per-gpu', type=int, default=0, help='0 = the config default')
    ap.add_argument('--precision', type=int,
                    default=int(os.environ.get('PYSEER_B200_LMM_PRECISION', '46')),
                    help='0 = FP64 CUDA-core contraction, 3..7 = exact int8-slice tcgen05 path, 46 = two passes '
                         '(4 slices, 6 again for the far tail)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-secondary', action='store_true')
    ap.add_argument('--e2e-chunks', type=int, default=0,
                    help='batches per e2e step (copy of batch i+1 overlaps the kernels of batch i)')
    ap.add_argument('--cpu-cores', type=int, default=0, help='0 = all available (max 64)')
    ap.add_argument('--check', type=int, default=2000,
                    help='variants of the shard re-checked against the oracle after the run')
    ap.add_argument('--parser-kmers', type=int, default=-1,
                    help='reference arm: k-mers of the text-parser leg (-1: 1500 for the LMM headline, else 0)')
    ap.add_argument('--state-npz', default=None,
                    help='(internal) once-per-run state handed to the CPU arm of a default run')
    return ap.parse_args(argv)


# ----------------------------------------------------------------------------------------
# workload definitions shared by both arms (no product import here)
# ----------------------------------------------------------------------------------------
def workload_spec(model, samples=0, kpg=0):
    d = {'lmm': (5000, 6250000), 'lmm-binary': (1000, 1000000), 'fixed': (2000, 10000000),
         'fixed-cont': (2000, 10000000), 'burden': (10000, 12500)}[model]
    return samples or d[0], kpg or d[1]


def config_of(model, n, kpg, world):
    packed_gb = kpg * ((n + 127) // 128 * 16) / 1e9
    base = {'n_samples': n, 'kmers_per_gpu': kpg, 'kmers_total': kpg * world,
            'filters': 'min_af 0.01 max_af 0.99 filter_pvalue 1 lrt_pvalue 1', 'block_size_cpu': BLOCK,
            'cache': 'inputs (%.2f GB packed rows per GPU) larger than L2' % packed_gb}
    if model == 'lmm':
        base.update(workload='LMM continuous phenotype, N=%d samples, %d synthetic k-mers per GPU '
                             '(BASELINE configs[3]: 50M k-mers x 5000 samples sharded by k-mer over 8 '
                             'GPUs), similarity kinship, D=1' % (n, kpg),
                    af='U(0.02,0.98), 0.1% planted causal')
    elif model == 'lmm-binary':
        base.update(workload='LMM binary phenotype, N=%d samples, %d synthetic k-mers per GPU (BASELINE '
                             'configs[1]: 1M k-mers x 1000 samples, 1 GPU), similarity kinship, D=1' % (n, kpg),
                    af='U(0.02,0.98), 0.1% planted causal')
    elif model == 'burden':
        base.update(workload='VCF burden test, LMM continuous phenotype, N=%d samples, %d burden regions per '
                             'GPU, each the union of 1-20 rare variant rows (BASELINE configs[4]: 100k regions '
                             'x 10000 samples over 8 GPUs); unit = regions' % (n, kpg),
                    af='member rows U(0.001,0.02); regions = OR of 1..20 members',
                    cache='member rows (%.2f GB) + region rows per GPU' % (10.5 * packed_gb))
    elif model == 'fixed':
        base.update(workload='fixed-effects logistic + Firth, N=%d samples, 10 MDS covariates, %d synthetic '
                             'k-mers per GPU (BASELINE configs[2])' % (n, kpg),
                    af='U(0.02,0.98), 0.1% planted causal, 0.1% rare and carried by cases only (-> Firth)')
    else:
        base.update(workload='fixed-effects OLS (continuous), N=%d samples, 10 MDS covariates, %d synthetic '
                             'k-mers per GPU' % (n, kpg), af='U(0.02,0.98), 0.1% planted causal')
    return base


def synth_params(model):
    """(seed, af_lo, af_hi, planted_every, separated_every) of the rows a model runs on."""
    if model == 'burden':
        return SEED + 5, 0.001, 0.02, 0, 0
    if model == 'fixed':
        return SEED, 0.02, 0.98, 1000, 1000
    return SEED, 0.02, 0.98, 1000, 0


import benchdata                                  # noqa: E402 -- problem definitions (plain NumPy)
from benchdata import fixed_cont_problem          # noqa: E402


def y_sign(model, y):
    cut = 0.5 if model in ('fixed', 'lmm-binary') else np.median(y)
    return np.where(y > cut, 1, -1).astype(np.int8)


# ----------------------------------------------------------------------------------------
# CPU arm (--impl reference; also the cpu_baseline leg, as a subprocess of the default run)
# ----------------------------------------------------------------------------------------
def reference_arm(a):
    from oracle import cpu_arm, ref_loader
    model = a.model
    n, kpg = workload_spec(model, a.samples, a.kmers_per_gpu)
    t0 = time.time()
    given = None
    if a.state_npz:
        with np.load(a.state_npz) as d:
            given = {k: d[k] for k in d.files}
    is_lmm = model in ('lmm', 'lmm-binary', 'burden')
    if is_lmm:
        if given is not None:
            state = dict(X=given['X'], y=given['y'], U=given['U'], S=given['S'], h2=float(given['h2']))
        else:
            X, y, K = cpu_arm.lmm_problem(n)
            if model == 'lmm-binary':
                y = (y > np.median(y)).astype(float)
            U, S, h2 = cpu_arm.lmm_spectral(X, y, K)
            del K
            state = dict(X=X, y=y, U=U, S=S, h2=h2)
        continuous = model != 'lmm-binary'
    else:
        from oracle import fixed_oracle as fo
        none = np.empty((0, 0))
        if model == 'fixed':
            m, y = cpu_arm.fixed_problem(n, 10)
            null = fo.fit_null(y, m, none, False)
            state = dict(y=y, m=m, null_llf=null.llf, null_firth=float(fo.fit_null(y, m, none, False, True)))
            continuous = False
        else:
            m, y = fixed_cont_problem(n)
            state = dict(y=y, m=m, null_llf=None, null_firth=0.0)
            continuous = True
    setup_s = time.time() - t0
    cores = cpu_arm.cpu_cores(a.cpu_cores)
    seed, af_lo, af_hi, planted, separated = synth_params(model)
    burden = None
    if model == 'burden':
        offs, mem = cpu_arm.burden_regions(kpg)
        burden = {'offs': offs, 'mem': mem, 'rec_first': 0}
        per_task = min(BLOCK, max(1, kpg // cores))
        tasks = [('burden', b * per_task, per_task) for b in range(cores)]
    else:
        per_task = BLOCK if is_lmm else (1000 if continuous else 150)
        tasks = [('synth', b * per_task, per_task) for b in range(cores)]
    arm = cpu_arm.CpuArm('lmm' if is_lmm else 'fixed', n, state, cores, continuous, af=(af_lo, af_hi),
                         planted=planted, separated=separated, seed=seed, burden=burden, **{
                             k: THRESH[k] for k in ('min_af', 'max_af', 'filter_pvalue', 'lrt_pvalue')})
    W, K = max(a.warmup, 0), max(a.steps, 1)
    for _ in range(W):
        arm.run(tasks)
    tested, secs = 0, 0.0
    for _ in range(K):
        out, s = arm.run(tasks)
        tested += sum(out)
        secs += s
    arm.close()
    rate = tested / secs
    sample = ('%d steps x %d workers x %d %s per worker and step (BLAS pinned to 1 thread as pyseer does); %s'
              % (K, cores, per_task, 'regions' if model == 'burden' else 'k-mers', arm.what))
    cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': arm.kind, 'sample': sample}
    n_parse = a.parser_kmers if a.parser_kmers >= 0 else (1500 if (model == 'lmm' and not a.state_npz) else 0)
    if n_parse > 0 and is_lmm and model != 'burden' and ref_loader.available():
        # leg (b) of BASELINE.md 3.1: gzipped k-mer text -> the reference's own parser -> fit_lmm
        leg = cpu_arm.reference_parser_leg(state, n, n_parse, continuous=continuous)
        cpu['text_leg'] = {'kmers': leg['kmers'], 'tested': leg['tested'], 'cores': 1,
                           'parse_kmers_per_s': leg['kmers'] / leg['parse_s'],
                           'end_to_end_kmers_per_s': leg['tested'] / leg['total_s'],
                           'what': 'gzipped k-mer text -> pyseer.input.load_var_block (read_variant, AF filter, '
                                   'hash_pattern, block matrix) -> pyseer.lmm.fit_lmm, one process'}
    line = {'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': a.gpus,
            'steps': K, 'warmup': W, 'ms_per_step': 1e3 * secs / K, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_of(model, n, kpg, max(a.gpus, 1)), 'cpu_baseline': cpu,
            'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0, 'setup_s': setup_s}
    _emit(line)
    return 0


def cpu_baseline_subprocess(a, model, n, kpg, state):
    """The CPU arm on a bounded sample, in a process of its own (its workers fork; this process holds a
    CUDA context).  The once-per-run state is handed over so that it is not recomputed."""
    fd, path = tempfile.mkstemp(prefix='psb_state_', suffix='.npz')
    os.close(fd)
    try:
        if 'U' in state:
            np.savez(path, X=state['X'], y=state['y'], U=state['U'], S=state['S'], h2=state['h2'])
            extra = ['--state-npz', path]
        else:
            extra = []
        cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--model', model, '--samples',
               str(n), '--kmers-per-gpu', str(kpg), '--steps', '2', '--warmup', '1', '--cpu-cores',
               str(a.cpu_cores), '--parser-kmers', '0'] + extra
        env = dict(os.environ)
        for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
            env.pop(k, None)
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=900)
        for ln in out.stdout.decode().splitlines()[::-1]:
            if ln.startswith('{'):
                return json.loads(ln)['cpu_baseline']
        return {'error': out.stderr.decode()[-400:]}
    except Exception as e:                                   # noqa: BLE001 -- reported, not fatal
        return {'error': repr(e)}
    finally:
        os.unlink(path)


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
                 '--format=csv,noheader,nounits', '-lms', '100'],
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
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(',')]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                    pw.append(float(f[3]))
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
                       reasons=sorted(reasons), samples=len(sm), power_w_max=float(max(pw)))
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


def file_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)), 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, \
        'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------------------
# GPU workloads
# ----------------------------------------------------------------------------------------
class LmmWorkload(object):
    model = 'lmm'
    continuous = True

    def __init__(self, a, n, kpg):
        self.a, self.n, self.kpg = a, n, kpg
        self.stats = None

    def build_state(self, device):
        from pyseer_b200.lmm import KinshipLMM
        X, y, K = benchdata.lmm_problem(self.n)
        if not self.continuous:
            y = (y > np.median(y)).astype(float)
        m = KinshipLMM(X, y.reshape(-1, 1), K, device=device)       # eigh on the device (psb_eigh)
        res = m.findH2()
        S, U = m.getSU()
        m.close()
        return {'X': X, 'y': y, 'U': np.ascontiguousarray(U), 'S': np.ascontiguousarray(S),
                'h2': np.array([float(res['h2'])])}

    def state_shapes(self):
        n = self.n
        return {'X': (n, 1), 'y': (n,), 'U': (n, n - 1), 'S': (n - 1,), 'h2': (1,)}

    def setup_engine(self, eng, st):
        eng.lmm_setup(st['X'], st['y'], st['U'], st['S'], float(st['h2'][0]), self.a.precision)

    def prepare(self, eng, rank, ys):
        seed, lo, hi, planted, sep = synth_params(self.model)
        eng.synth_device(seed, rank * self.kpg, self.kpg, lo, hi, planted, ys, sep)

    def run(self, eng):
        eng.run_lmm(continuous=self.continuous, **THRESH)

    def slices(self):
        """int8 slices of the working precision (precision 46: 4, with 6 again for the far tail)."""
        return 4 if self.a.precision == 46 else self.a.precision

    def dtype(self):
        k = self.a.precision
        if k == 46:
            return 's8 x4 slices (x6 for F > 30) -> s32 (tcgen05) -> f64'
        return ('s8 x%d slices -> s32 (tcgen05) -> f64' % k) if k else 'f64'

    def check(self, st, bits_head, cols):
        from oracle import lmm_oracle as lo, synth
        olmm = lo.OracleLMM(st['X'], st['y'].reshape(-1, 1), None)
        olmm.U, olmm.S = st['U'], st['S']
        x = synth.unpack_rows(bits_head, self.n)
        ref = lo.fit_lmm_block(olmm, float(st['h2'][0]), np.ascontiguousarray(x.T, dtype=float))
        pv, be = cols['pvalue'], cols['beta']
        ok = np.isfinite(pv) & (ref['p_values'] > 1e-290)
        return {'variants': int(ok.sum()),
                'max_rel_err_pvalue': float(np.max(np.abs(pv[ok] / ref['p_values'][ok] - 1))),
                'max_rel_err_beta': float(np.max(np.abs(be[ok] / ref['beta'][ok] - 1))),
                'min_pvalue': float(np.min(pv[ok]))}

    def roofline(self, tested, k_ms, run_ms, W, pk, pk_kind, dev_peaks):
        n, J, k = self.n, self.n - 1, self.slices()
        tri = os.environ.get('PSB_TC_TRI', '1') != '0'
        # int8 multiply-adds the kernel issues per k-mer: k slices over N^2/2 (triangular form,
        # K stages of 256 samples from the diagonal down) or N (N-D) entries
        kp = (n + 255) // 256 * 256
        if tri:
            macs = sum(32 * (kp - (32 * jt) // 256 * 256) for jt in range((n + 31) // 32)) * k
        else:
            macs = kp * ((J + 31) // 32 * 32) * k
        alg = 2.0 * n * J * tested / (k_ms / 1e3) / 1e12
        if not k:
            peak = dev_peaks['fp64_tflops']
            return {'bound': 'tensor', 'kernel': 'k_lmm_quadform_fp64', 'achieved': alg, 'peak': peak,
                    'unit': 'TFLOP/s', 'frac': alg / peak, 'peak_source': 'measured fp64 FMA rate (psb_measure_peaks)',
                    'kernel_ms': k_ms, 'run_ms': run_ms, 'traffic': None}
        executed = 2.0 * macs * tested / (k_ms / 1e3) / 1e12
        peak = dev_peaks['int8_tops']
        return {'bound': 'tensor', 'kernel': 'k_lmm_quadform_tc', 'achieved': executed, 'peak': peak,
                'unit': 'TFLOP/s', 'frac': executed / peak,
                'peak_source': 'MEASURED dense int8 tensor rate of this board under the clocks of this run '
                               '(psb_measure_peaks: tcgen05.mma kind::i8 M128 N256 K32 issued back to back on '
                               'all SMs right after the timed region).  achieved = int8 operations the kernel '
                               'EXECUTES (unit: tera-ops/s, s8 x s8 -> s32): %d exact slices%s'
                               % (k, ' over the lower triangle of M (a = x\'M\'\'x)' if tri else ''),
                'executed_int8_macs_per_kmer': float(macs),
                'algorithmic_flops_per_kmer': 2.0 * n * J,
                'fp64_equivalent_tflops': alg,
                'fp64_equivalent_vs_bf16_sustained': alg / pk.get('bf16_tflops_sustained', pk['bf16_tflops']),
                'kernel_ms': k_ms, 'run_ms': run_ms, 'kernel_share_of_step': k_ms / run_ms,
                'side_kernels_ms': run_ms - k_ms,
                'hbm_read_frac': (tested * (W * 4 + 56) / (k_ms / 1e3) / 1e9) / pk['hbm_gbs'],
                'algorithmic_bytes': tested * (W * 4 + 24.0),
                'traffic': ncu_traffic('lmm:n=%d:kmers=%d:k=%d' % (n, self.kpg, self.a.precision))}


class LmmBinaryWorkload(LmmWorkload):
    model = 'lmm-binary'
    continuous = False


class BurdenWorkload(LmmWorkload):
    """BASELINE configs[4]: a region is the union of 1-20 rare variant rows (af ~ U(0.001, 0.02),
    dominant encoding, input.py:395-407); the union runs on the device (psb_submit_burden*), then the
    region rows take the LMM path.  `kpg` counts REGIONS per GPU; the unit is regions tested/s."""
    model = 'burden'

    def prepare(self, eng, rank, ys):
        self.offs, self.mem = benchdata.burden_regions(self.kpg, rank)
        self.n_rec = int(self.offs[-1])
        seed, lo, hi, _, _ = synth_params(self.model)
        eng.synth_device(seed, rank * 21 * self.kpg, self.n_rec, lo, hi, 0, None)
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

    def roofline(self, tested, k_ms, run_ms, W, pk, pk_kind, dev_peaks):
        r = LmmWorkload.roofline(self, tested, k_ms, run_ms, W, pk, pk_kind, dev_peaks)
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
    model = 'fixed'
    continuous = False
    DIMS = 10

    def __init__(self, a, n, kpg):
        self.a, self.n, self.kpg = a, n, kpg
        self.stats = None

    def build_state(self, device):
        from pyseer_b200 import model as pm
        m, y = benchdata.fixed_problem(self.n, self.DIMS)
        none = np.empty((0, 0))
        null = pm.fit_null(y, m, none, False, device=device)
        firth = pm.fit_null(y, m, none, False, True, device=device)
        return {'m': m, 'y': y, 'meta': np.array([null.llf, firth])}

    def state_shapes(self):
        return {'m': (self.n, self.DIMS), 'y': (self.n,), 'meta': (2,)}

    def setup_engine(self, eng, st):
        Z = np.c_[np.ones(self.n), st['m']]
        eng.fixed_setup(Z, st['y'], self.continuous, float(st['meta'][0]), float(st['meta'][1]))

    def prepare(self, eng, rank, ys):
        seed, lo, hi, planted, sep = synth_params(self.model)
        eng.synth_device(seed, rank * self.kpg, self.kpg, lo, hi, planted, ys, sep)

    def run(self, eng):
        eng.run_fixed(continuous=self.continuous, **THRESH)

    def dtype(self):
        return 'f64'

    def check(self, st, bits_head, cols):
        from oracle import fixed_oracle as fo, synth
        # the head of the shard plus every separated (Firth) row among the first 20000
        x = synth.unpack_rows(bits_head, self.n).astype(float)
        rows = list(range(min(200, x.shape[0]))) + [r for r in range(500, x.shape[0], 1000)][:20]
        none = np.empty((0, 0))
        errp, errb, nf = 0.0, 0.0, 0
        for s in rows:
            o = fo.fixed_effects_regression('k', st['y'], x[s], st['m'], none, 0.5, 'p', False, None,
                                            1.0, 1.0, float(st['meta'][0]), float(st['meta'][1]),
                                            [], [], self.continuous)
            if o.prefilter or not np.isfinite(o.pvalue) or o.pvalue < 1e-290 or 'firth-fail' in o.notes:
                continue
            errp = max(errp, abs(cols['pvalue'][s] / o.pvalue - 1))
            errb = max(errb, abs(cols['beta'][s] - o.kbeta) / max(abs(o.kbeta), 1e-3 * o.bse))
            nf += 'bad-chisq' in o.notes or 'high-bse' in o.notes
        return {'variants': len(rows), 'max_rel_err_pvalue': float(errp),
                'max_rel_err_beta': float(errb), 'firth_fits_checked': int(nf)}

    def roofline(self, tested, k_ms, run_ms, W, pk, pk_kind, dev_peaks):
        n, p = self.n, self.DIMS + 2
        per_eval = n * (p * (p + 1) / 2 + 2 * p + 30) * 2.0       # X'WX + score + eta, exp/div
        evals = self.stats['newton_evaluations'] if self.stats else 3.0 * tested
        achieved = per_eval * evals / (k_ms / 1e3) / 1e12
        peak = dev_peaks['fp64_tflops']
        return {'bound': 'tensor', 'kernel': 'k_fixed_logit(+k_fixed_firth)', 'achieved': achieved,
                'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                'peak_source': 'MEASURED fp64 FMA rate of this board under the clocks of this run '
                               '(psb_measure_peaks).  The kernel is bound by the FP64 pipe, not by tensor cores '
                               'or HBM ("tensor" only names the compute side of the roofline): achieved = '
                               'measured Newton evaluations x N (p(p+1)/2 + 2p + 30) FMA x 2',
                'newton_evaluations_per_variant': evals / max(tested, 1),
                'firth_fits': self.stats['firth_fits'] if self.stats else None,
                'kernel_ms': k_ms, 'run_ms': run_ms, 'kernel_share_of_step': k_ms / run_ms,
                'hbm_read_frac': (tested * (W * 4 + 8 * (6 + p)) / (k_ms / 1e3) / 1e9) / pk['hbm_gbs'],
                'traffic': ncu_traffic('fixed:n=%d:kmers=%d' % (n, self.kpg))}


class FixedContWorkload(FixedWorkload):
    """Fixed effects with a continuous phenotype: closed-form OLS t-test per variant."""
    model = 'fixed-cont'
    continuous = True

    def build_state(self, device):
        m, y = fixed_cont_problem(self.n, self.DIMS)
        return {'m': m, 'y': y, 'meta': np.array([0.0, 0.0])}

    def roofline(self, tested, k_ms, run_ms, W, pk, pk_kind, dev_peaks):
        q = self.DIMS + 1
        bytes_alg = tested * (W * 4 + 8.0 * (6 + q))
        achieved = bytes_alg / (run_ms / 1e3) / 1e9
        return {'bound': 'hbm', 'kernel': 'k_bitstats + k_lmm_quadform_tc (linear tile) + k_fixed_ols (whole run)',
                'achieved': achieved, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': achieved / pk['hbm_gbs'],
                'peak_source': '%s STREAM-style copy bandwidth; algorithmic bytes = packed row + result row '
                               'per tested variant' % pk_kind,
                'kernel_ms': k_ms, 'run_ms': run_ms, 'traffic': None}


WORKLOADS = {'lmm': LmmWorkload, 'lmm-binary': LmmBinaryWorkload, 'fixed': FixedWorkload,
             'fixed-cont': FixedContWorkload, 'burden': BurdenWorkload}


# ----------------------------------------------------------------------------------------
def gpu_line(a, model, n, kpg, rank, local_rank, world, comm, eng, steps, warmup, cpu_baseline=True):
    """One bench line of the GPU arm for `model`.  `eng` is this rank's context, `comm` the library's
    NCCL communicator (None at N = 1).  Returns the line (rank 0) or None."""
    from pyseer_b200.engine import PinnedBuffer, words_per_row
    from pyseer_b200 import sharding
    wl = WORKLOADS[model](a, n, kpg)
    config = config_of(model, n, kpg, world)

    # ---- once-per-run state: rank 0 builds it, the other ranks receive it (ncclBroadcast) ----
    t_setup = time.time()
    st = wl.build_state(local_rank) if rank == 0 else None
    if comm is not None:
        out = {}
        for key, shape in wl.state_shapes().items():
            arr = np.ascontiguousarray(st[key], dtype=np.float64) if rank == 0 else \
                np.empty(shape, dtype=np.float64)
            out[key] = comm.bcast(arr, 0)
        st = out
    ys = y_sign(model, st['y'])
    wl.setup_engine(eng, st)
    wl.prepare(eng, rank, ys)
    t_setup = time.time() - t_setup
    W = words_per_row(n)
    COLS = tuple((name, b) for name, b, _ in sharding.TABLE_COLUMNS)
    row_bytes = sharding.ROW_BYTES

    def barrier():
        if comm is not None:
            comm.barrier()

    def step():
        wl.run(eng)
        if comm is not None:
            # the one collective of the path: NCCL gather of the per-variant result table on rank 0,
            # queued behind the run on the communicator's stream (overlaps the next step's kernels)
            comm.gather_begin(kpg, root=0)

    for _ in range(max(warmup, 0)):
        step()
    if comm is not None:
        comm.gather_wait()
    sampler = ClockSampler(visible_device(local_rank)) if rank == 0 else None
    barrier()
    eng.sync()
    if sampler:
        sampler.start()
    l0 = eng.launch_count()
    eng.event_record(0)
    for _ in range(steps):
        step()
    if comm is not None:
        comm.gather_wait()              # joins the last gather into the compute stream: it is timed
    eng.event_record(1)
    eng.sync()
    barrier()
    ms = eng.event_elapsed(0, 1)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    dev_peaks = eng.measure_peaks() if rank == 0 else None      # same power state as the timed region
    counts = eng.counts()
    tested = counts['tested']
    wl.stats = eng.last_stats() if model == 'fixed' else None
    # dominant-kernel time, CUDA events on the library stream around the dominant launch of
    # the last timed step (every step launches the same grid on the same rows)
    k_ms = eng.last_ms(1)
    run_ms = eng.last_ms(0)
    if model == 'burden':
        wl.or_ms = eng.event_elapsed(4, 5)
    gathered = None
    if comm is not None:
        ms = float(comm.allreduce([ms], 'max')[0])
        tested_all = float(comm.allreduce([float(tested)], 'sum')[0])
        if rank == 0:
            # content check of the gather: rank r's table as received on rank 0
            rows = tested_rows = 0
            for r in range(world):
                _, nr, cnt = comm.gather_fetch(r, pointers={})
                rows += nr
                tested_rows += cnt['tested']
            gathered = {'rows': rows, 'tested': tested_rows, 'bytes_per_rank': comm.gather_bytes(),
                        'nccl_version': comm.info()['nccl_version']}
    else:
        tested_all = float(tested)
    value = tested_all * steps / (ms / 1e3)

    # ---- end to end through the C ABI with host buffers -----------------------------------
    e2e = None
    check = None
    if not a.no_e2e:
        outs = {name: PinnedBuffer((kpg,), {4: np.int32, 8: np.float64}[b] if name != 'flags'
                                   else np.uint32) for name, b in COLS}

        def ptrs_at(lo):
            return {name: outs[name].array[lo:].ctypes.data for name, _ in COLS}

        if model == 'burden':
            # host record rows + member lists -> psb_submit_burden (H2D, device union) -> LMM -> table
            eng.submit_device(wl.rec_ptr, wl.n_rec, wl.rec_w)
            pin = PinnedBuffer((wl.n_rec, W), np.uint32)
            eng.download_bits(pin.array)
            chunks = 1
            h2d = int(wl.n_rec * W * 4 + wl.mem.nbytes + wl.offs.nbytes)

            def e2e_step():
                eng.submit_burden(pin.array, None, wl.offs, wl.mem)
                LmmWorkload.run(wl, eng)
                eng.fetch_into(ptrs_at(0))
        else:
            pin = PinnedBuffer((kpg, W), np.uint32)
            eng.download_bits(pin.array)
            # One step = the whole shard through the public calls a user makes, in `chunks`
            # batches: psb_submit copies batch i+1 (copy stream, second staging slot) while the
            # kernels of batch i run; psb_fetch of batch i then brings its rows of the table back.
            chunks = max(1, a.e2e_chunks or {'lmm': 8, 'lmm-binary': 4, 'fixed': 4, 'fixed-cont': 8}.get(model, 8))
            # a short first batch (its copy is the only one nothing hides) and a short last one (so is
            # its table's way back); the batches between them are large, which keeps the partial last
            # wave of every tensor-kernel launch and the per-launch costs small
            if chunks >= 4:
                edge = max(256, (kpg // (4 * chunks)) // 256 * 256)
                mid = [edge + (kpg - 2 * edge) * i // (chunks - 2) for i in range(chunks - 1)]
                cuts = [0] + mid + [kpg]
            else:
                cuts = [kpg * i // chunks for i in range(chunks + 1)]
            bounds = [(cuts[i], cuts[i + 1]) for i in range(chunks)]
            h2d = int(kpg * W * 4)

            def e2e_step():
                # submit(i+1) on the copy stream, the table of batch i on the fetch stream
                # (psb_fetch_begin) and the kernels of batch i+1 all overlap; every table has landed
                # on the host before the step ends (psb_fetch_wait)
                lo, hi = bounds[0]
                eng.submit(pin.array[lo:hi])
                wl.run(eng)
                for i in range(1, chunks):
                    nlo, nhi = bounds[i]
                    eng.submit(pin.array[nlo:nhi])
                    if i > 1:
                        eng.fetch_wait()
                    eng.fetch_begin(ptrs_at(lo))
                    wl.run(eng)
                    lo, hi = nlo, nhi
                if chunks > 1:
                    eng.fetch_wait()
                eng.fetch_into(ptrs_at(lo))

        e2e_step()
        barrier()
        eng.sync()
        eng.event_record(2)
        for _ in range(steps):
            e2e_step()
        eng.event_record(3)
        eng.sync()
        barrier()
        ems = eng.event_elapsed(2, 3)
        if comm is not None:
            ems = float(comm.allreduce([ems], 'max')[0])
        e2e = {'value': tested_all * steps / (ems / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': int(kpg * row_bytes), 'ms_per_step': ems / steps,
               'chunks_per_step': chunks}
        # spot check of the timed output against the oracle (not timed)
        if rank == 0 and a.check > 0:
            if model == 'burden':
                nrec = int(wl.offs[min(300, kpg)])
                cols = {name: outs[name].array[:300].copy() for name in ('pvalue', 'beta')}
                check = wl.check(st, pin.array[:nrec].copy(), cols)
            else:
                nchk = min(kpg, 20000 if model == 'fixed' else a.check)
                cols = {name: outs[name].array[:nchk].copy() for name in ('pvalue', 'beta')}
                check = wl.check(st, pin.array[:nchk].copy(), cols)
        pin.free()
        for o in outs.values():
            o.free()

    line = None
    if rank == 0:
        pk, pk_kind = file_peaks()
        roof = wl.roofline(tested, k_ms, run_ms, W, pk, pk_kind, dev_peaks)
        cpu_line = None
        if cpu_baseline and not a.no_cpu_baseline:
            cs = dict(st)
            if 'h2' in cs:
                cs['h2'] = float(cs['h2'][0])
            cpu_line = cpu_baseline_subprocess(a, model, n, kpg, cs)
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps,
                'warmup': warmup, 'ms_per_step': ms / steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': wl.dtype(),
                'data': 'synthetic', 'config': config, 'clocks': clocks, 'e2e': e2e,
                'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cpu_line,
                'counts': counts, 'check': check, 'setup_s': t_setup,
                'measured_peaks': dev_peaks, 'gather': gathered}
        if wl.stats:
            line['stats'] = wl.stats
        if 'h2' in st:
            line['h2'] = float(st['h2'][0])
    return line


def _emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = json.dumps(line) + '\n'
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, out.encode())
    else:
        sys.stdout.write(out)
        sys.stdout.flush()


_REAL_STDOUT = None


def _guard_stdout():
    """Libraries below us write to file descriptor 1 (NCCL prints its version banner there when the
    box sets NCCL_DEBUG): everything but the JSON line goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def main(argv=None):
    a = parse_args(argv)
    _guard_stdout()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if a.impl == 'reference':
        return reference_arm(a) if rank == 0 else 0

    from pyseer_b200.engine import Engine
    eng = Engine(local_rank)
    comm = None
    if world > 1:
        from pyseer_b200.comm import Comm
        comm = Comm.from_env(eng)
    n, kpg = workload_spec(a.model, a.samples, a.kmers_per_gpu)
    line = gpu_line(a, a.model, n, kpg, rank, local_rank, world, comm, eng, a.steps, a.warmup)
    if comm is not None:
        comm.close()
    eng.close()
    if rank == 0 and world == 1 and a.model == 'lmm' and not a.no_secondary \
            and not a.samples and not a.kmers_per_gpu:
        # the other BASELINE configs through the same harness (one GPU; configs[4] with all of its
        # 100k regions on this GPU)
        sec = []
        for model, kp in (('lmm-binary', 0), ('fixed', 0), ('burden', 100000)):
            try:
                e2 = Engine(local_rank)
                sn, skpg = workload_spec(model, 0, kp)
                sec.append(gpu_line(a, model, sn, skpg, 0, local_rank, 1, None, e2, max(3, min(a.steps, 5)), 3))
                e2.close()
            except Exception as exc:                           # noqa: BLE001 -- reported in the line
                sec.append({'config': config_of(model, *workload_spec(model, 0, kp), 1), 'error': repr(exc)})
        line['secondary'] = sec
    if rank == 0:
        _emit(line)
    return 0


if __name__ == '__main__':
    sys.exit(main())
