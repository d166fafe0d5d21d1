#!/usr/bin/env python
"""Steady-state rate of `python -m pyseer_b200 --lmm --bits-cache` over a LARGE packed cache (the runs of
scripts/cli_throughput.py last a fraction of a second): a cache of `--kmers` synthetic variants at N
samples is written directly (one chunk of synthetic rows repeated under fresh names), then streamed.
Prints one JSON object; run on the GPU box."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--samples', type=int, default=5000)
    ap.add_argument('--kmers', type=int, default=2000000)
    ap.add_argument('--profile', action='store_true')
    ap.add_argument('--quick', action='store_true', help='default batch only, with and without an lrt filter')
    a = ap.parse_args()
    import benchdata
    from pyseer_b200.engine import synth_host
    from pyseer_b200.input import PackedCacheWriter, VariantBatch
    from pyseer_b200.lmm import KinshipLMM
    n, m = a.samples, a.kmers
    cores = len(os.sched_getaffinity(0))
    d = tempfile.mkdtemp(prefix='psb_cache_')
    X, y, K = benchdata.lmm_problem(n)
    samples = ['s%d' % i for i in range(n)]
    with open(os.path.join(d, 'pheno.tsv'), 'w') as fh:
        fh.write('samples\tpheno\n')
        for i in range(n):
            fh.write('s%d\t%r\n' % (i, float(y[i])))
    src = os.path.join(d, 'kmers.txt')
    with open(src, 'w') as fh:
        fh.write('K | s0:1\n')
    cache = os.path.join(d, 'kmers.bits')
    t0 = time.time()
    chunk = 48000
    rows = synth_host(7, 0, chunk, n)
    w = PackedCacheWriter(cache, 'kmers', src, samples, rows.shape[1])
    for lo in range(0, m, chunk):
        k = min(chunk, m - lo)
        w.add(VariantBatch(['K%09d' % (lo + i) for i in range(k)], rows[:k], None))
    w.close()
    gen_s = time.time() - t0
    lm = KinshipLMM(X, y.reshape(-1, 1), K)
    h2 = float(lm.findH2()['h2'])
    S, U = lm.getSU()
    lm.close()
    np.savez(os.path.join(d, 'lmm.npz'), U, S, np.array([h2]))
    base = [sys.executable] + (['-m', 'cProfile', '-s', 'cumtime'] if a.profile else []) + \
        ['-m', 'pyseer_b200', '--phenotypes', os.path.join(d, 'pheno.tsv'), '--lmm', '--load-lmm',
         os.path.join(d, 'lmm.npz'), '--cpu', str(cores), '--kmers', src, '--uncompressed', '--bits-cache', cache]
    res = {'n_samples': n, 'kmers': m, 'cache_bytes': os.path.getsize(cache), 'write_cache_s': gen_s, 'runs': {}}
    runs = (('default', []), ('lrt_1e-4', ['--lrt-pvalue', '1e-4']), ('default_again', [])) if a.quick else None
    for tag, extra in runs or (('default', []), ('gpu_batch_24000', ['--gpu-batch', '24000']), ('gpu_batch_12000', ['--gpu-batch', '12000']),
                       ('lrt_1e-4', ['--lrt-pvalue', '1e-4']), ('default_again', [])):
        t = time.time()
        env = dict(os.environ, PYSEER_B200_TIMING='1')
        out_path = os.path.join(d, 'out.tsv')
        with open(out_path, 'w') as fo:
            err = subprocess.run(base + extra, stdout=fo, stderr=subprocess.PIPE, cwd=ROOT, env=env,
                                 check=True).stderr.decode()
        stream = [ln for ln in err.splitlines() if ln.startswith('pipeline:')]
        rate = float(stream[-1].split('=')[1].split()[0]) if stream else None
        res['runs'][tag] = {'wall_s': time.time() - t, 'variants_per_s_streaming': rate,
                            'output_bytes': os.path.getsize(out_path)}
        sys.stderr.write('%s: %.2f s, streaming %s variants/s, %d bytes of output\n'
                         % (tag, time.time() - t, rate, os.path.getsize(out_path)))
        if a.profile:
            sys.stderr.write('\n'.join(l for l in open(out_path).read().splitlines()[-60:]) + '\n')
            break
    print(json.dumps(res))
    for f in os.listdir(d):
        os.unlink(os.path.join(d, f))
    os.rmdir(d)


if __name__ == '__main__':
    main()
