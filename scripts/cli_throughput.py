#!/usr/bin/env python
"""CLI-level throughput of `python -m pyseer_b200 --lmm` on real input formats at N = 5000:
gzip text, bgzip text (block-parallel inflate), plain text, and the packed --bits-cache of a second
run.  Prints one JSON object; run on the GPU box (scripts/gpu_r2_*.sh)."""
import argparse
import gzip
import json
import os
import struct
import subprocess
import sys
import tempfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_bgzf(path, src, block=65000, level=1):
    with open(src, 'rb') as fi, open(path, 'wb') as fh:
        while True:
            chunk = fi.read(block)
            co = zlib.compressobj(level, zlib.DEFLATED, -15)
            payload = co.compress(chunk) + co.flush()
            bsize = 12 + 6 + len(payload) + 8
            fh.write(b'\x1f\x8b\x08\x04' + b'\x00' * 4 + b'\x00\xff' + struct.pack('<H', 6) + b'BC' +
                     struct.pack('<HH', 2, bsize - 1) + payload +
                     struct.pack('<II', zlib.crc32(chunk) & 0xffffffff, len(chunk)))
            if not chunk:
                break


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--samples', type=int, default=5000)
    ap.add_argument('--kmers', type=int, default=40000)
    ap.add_argument('--cpu', type=int, default=0)
    a = ap.parse_args()
    import benchdata
    from oracle import synth
    from pyseer_b200.lmm import KinshipLMM
    n, m = a.samples, a.kmers
    cores = a.cpu or len(os.sched_getaffinity(0))
    d = tempfile.mkdtemp(prefix='psb_cli_')
    X, y, K = benchdata.lmm_problem(n)
    lm = KinshipLMM(X, y.reshape(-1, 1), K)
    h2 = float(lm.findH2()['h2'])
    S, U = lm.getSU()
    lm.close()
    np.savez(os.path.join(d, 'lmm.npz'), U, S, np.array([h2]))
    names = np.array(['s%d' % i for i in range(n)])
    with open(os.path.join(d, 'pheno.tsv'), 'w') as fh:
        fh.write('samples\tpheno\n')
        for i in range(n):
            fh.write('s%d\t%r\n' % (i, float(y[i])))
    ys = np.where(y > np.median(y), 1, -1).astype(np.int8)
    txt = os.path.join(d, 'kmers.txt')
    t0 = time.time()
    with open(txt, 'w') as fh:
        for lo in range(0, m, 2000):
            x = synth.unpack_rows(synth.synth_rows(7, lo, min(2000, m - lo), n, 0.02, 0.98, 1000, ys), n)
            tok = np.char.add(names, ':1')
            for s in range(x.shape[0]):
                fh.write('K%08d | ' % (lo + s) + ' '.join(tok[x[s] != 0]) + '\n')
    gen_s = time.time() - t0
    size_txt = os.path.getsize(txt)
    subprocess.check_call('gzip -1 -k -c %s > %s.gz' % (txt, txt), shell=True)
    write_bgzf(txt + '.bgz', txt)
    base = [sys.executable, '-m', 'pyseer_b200', '--phenotypes', os.path.join(d, 'pheno.tsv'), '--lmm',
            '--load-lmm', os.path.join(d, 'lmm.npz'), '--cpu', str(cores)]
    # set-up time of a run (load the LMM cache, psb_lmm_setup): one-line input
    one = os.path.join(d, 'one.txt')
    with open(txt) as fi, open(one, 'w') as fo:
        fo.write(fi.readline())

    def run(extra, tag):
        t = time.time()
        env = dict(os.environ, PYSEER_B200_TIMING='1')
        with open(os.devnull, 'w') as null:
            err = subprocess.run(base + extra, stdout=null, stderr=subprocess.PIPE, cwd=ROOT, env=env,
                                 check=True).stderr.decode()
        stream = [ln for ln in err.splitlines() if ln.startswith('pipeline:')]
        rate = float(stream[-1].split('=')[1].split()[0]) if stream else None
        return time.time() - t, rate

    setup_s = run(['--kmers', one, '--uncompressed'], 'setup')[0]
    cache = os.path.join(d, 'kmers.bits')
    res = {'n_samples': n, 'kmers': m, 'parser_threads': cores, 'text_bytes': size_txt, 'setup_s': setup_s,
           'generate_text_s': gen_s, 'runs': {}}
    for tag, extra in (('gzip_text', ['--kmers', txt + '.gz']),
                       ('bgzip_text', ['--kmers', txt + '.bgz']),
                       ('plain_text', ['--kmers', txt, '--uncompressed']),
                       ('gzip_text_writing_bits_cache', ['--kmers', txt + '.gz', '--bits-cache', cache]),
                       ('bits_cache', ['--kmers', txt + '.gz', '--bits-cache', cache])):
        w, rate = run(extra, tag)
        res['runs'][tag] = {'wall_s': w, 'variants_per_s_whole_run': m / w,
                            'variants_per_s_streaming': rate}
    print(json.dumps(res))
    for f in os.listdir(d):
        os.unlink(os.path.join(d, f))
    os.rmdir(d)


if __name__ == '__main__':
    main()
