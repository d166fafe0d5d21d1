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


def _write_part(job):
    d, i, lo, hi, n, ys = job
    from oracle import synth
    names = np.array(['s%d' % k for k in range(n)])
    tok = np.char.add(names, ':1')
    f = os.path.join(d, 'part%04d.txt' % i)
    with open(f, 'w') as fh:
        for a in range(lo, hi, 2000):
            x = synth.unpack_rows(synth.synth_rows(7, a, min(2000, hi - a), n, 0.02, 0.98, 1000, ys), n)
            for s in range(x.shape[0]):
                fh.write('K%08d | ' % (a + s) + ' '.join(tok[x[s] != 0]) + '\n')
    subprocess.check_call('gzip -1 -k -c %s > %s.gz' % (f, f), shell=True)
    write_bgzf(f + '.bgz', f)
    # this part's share of ONE gzip member (what `gzip` or `pigz` make of the whole text): raw deflate,
    # no final block, flushed to a byte boundary
    co = zlib.compressobj(1, zlib.DEFLATED, -15)
    with open(f, 'rb') as fi, open(f + '.gz1', 'wb') as fo:
        while True:
            blk = fi.read(16 << 20)
            if not blk:
                break
            fo.write(co.compress(blk))
        fo.write(co.flush(zlib.Z_FULL_FLUSH))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--samples', type=int, default=5000)
    ap.add_argument('--kmers', type=int, default=40000)
    ap.add_argument('--cpu', type=int, default=0)
    ap.add_argument('--part', type=int, default=10000, help='k-mers per generator job')
    ap.add_argument('--similarity-only', action='store_true', help='time the similarity tool on the text file and stop')
    ap.add_argument('--sweep', default='', help='comma-separated --gpu-batch values to try on the text formats')
    a = ap.parse_args()
    import benchdata
    from oracle import synth
    from pyseer_b200.lmm import KinshipLMM
    n, m = a.samples, a.kmers
    cores = a.cpu or len(os.sched_getaffinity(0))
    d = tempfile.mkdtemp(prefix='psb_cli_')
    X, y, K = benchdata.lmm_problem(n)
    with open(os.path.join(d, 'pheno.tsv'), 'w') as fh:
        fh.write('samples\tpheno\n')
        for i in range(n):
            fh.write('s%d\t%r\n' % (i, float(y[i])))
    ys = np.where(y > np.median(y), 1, -1).astype(np.int8)
    txt = os.path.join(d, 'kmers.txt')
    t0 = time.time()
    # the text, its gzip and its bgzip image, written in parts on all cores and concatenated (a series
    # of gzip members is a gzip file, a series of BGZF files a BGZF file)
    parts = [(d, i, lo, min(lo + a.part, m), n, ys) for i, lo in enumerate(range(0, m, a.part))]
    import multiprocessing as mp
    with mp.get_context('fork').Pool(min(cores, len(parts))) as pool:
        pool.map(_write_part, parts)
    crc, total = 0, 0
    for ext in ('', '.gz', '.bgz', '.gz1'):
        with open(txt + ext, 'wb') as fo:
            if ext == '.gz1':
                fo.write(b'\x1f\x8b\x08\x00' + b'\x00' * 4 + b'\x00\xff')
            for i in range(len(parts)):
                f = os.path.join(d, 'part%04d.txt%s' % (i, ext))
                with open(f, 'rb') as fi:
                    while True:
                        blk = fi.read(64 << 20)
                        if not blk:
                            break
                        if ext == '':
                            crc = zlib.crc32(blk, crc)
                            total += len(blk)
                        fo.write(blk)
                os.unlink(f)
            if ext == '.gz1':                   # empty final block, CRC-32 and length of the whole text
                fo.write(b'\x03\x00' + struct.pack('<II', crc & 0xffffffff, total & 0xffffffff))
    gen_s = time.time() - t0
    size_txt = os.path.getsize(txt)
    if a.similarity_only:
        # the similarity tool (K = G G') on the same k-mer text: device parser + psb_kinship_add_submitted
        # against the host parser + psb_kinship_add
        with open(os.path.join(d, 'samples.txt'), 'w') as fh:
            fh.write('\n'.join('s%d' % i for i in range(n)) + '\n')
        res = {'n_samples': n, 'kmers': m, 'text_bytes': size_txt, 'runs': {}}
        for tag, src, extra, text in (('plain_text_device_parser', txt, ['--uncompressed'], '1'),
                                      ('gzip_one_member_device_parser', txt + '.gz1', [], '1'),
                                      ('plain_text_host_parser', txt, ['--uncompressed'], '0')):
            t = time.time()
            with open(os.path.join(d, 'K_%s.tsv' % tag), 'w') as fo:
                subprocess.run([sys.executable, '-m', 'pyseer_b200.similarity', os.path.join(d, 'samples.txt'),
                                '--kmers', src] + extra, stdout=fo, stderr=subprocess.DEVNULL, cwd=ROOT,
                               env=dict(os.environ, PYSEER_B200_TEXT=text), check=True)
            res['runs'][tag] = {'wall_s': time.time() - t}
            sys.stderr.write('similarity %s: %.2f s\n' % (tag, time.time() - t))
        ref = open(os.path.join(d, 'K_plain_text_host_parser.tsv')).read()
        res['equal'] = all(open(os.path.join(d, 'K_%s.tsv' % t)).read() == ref for t in res['runs'])
        print(json.dumps(res))
        for f in os.listdir(d):
            os.unlink(os.path.join(d, f))
        os.rmdir(d)
        return
    # (the device is touched only after the generator processes were forked)
    lm = KinshipLMM(X, y.reshape(-1, 1), K)
    h2 = float(lm.findH2()['h2'])
    S, U = lm.getSU()
    lm.close()
    np.savez(os.path.join(d, 'lmm.npz'), U, S, np.array([h2]))
    base = [sys.executable, '-m', 'pyseer_b200', '--phenotypes', os.path.join(d, 'pheno.tsv'), '--lmm',
            '--load-lmm', os.path.join(d, 'lmm.npz'), '--cpu', str(cores)]
    # set-up time of a run (load the LMM cache, psb_lmm_setup): one-line input
    one = os.path.join(d, 'one.txt')
    with open(txt) as fi, open(one, 'w') as fo:
        fo.write(fi.readline())

    def run(extra, tag, text='1'):
        t = time.time()
        env = dict(os.environ, PYSEER_B200_TIMING='1', PYSEER_B200_TEXT='1' if text == 'zlib' else text)
        if text == 'zlib':
            env['PSB_PGZ'] = '0'                # the serial inflater, for comparison
        with open(os.devnull, 'w') as null:
            err = subprocess.run(base + extra, stdout=null, stderr=subprocess.PIPE, cwd=ROOT, env=env,
                                 check=True).stderr.decode()
        stream = [ln for ln in err.splitlines() if ln.startswith('pipeline:')]
        rate = float(stream[-1].split('=')[1].split()[0]) if stream else None
        return time.time() - t, rate

    setup_s = run(['--kmers', one, '--uncompressed'], 'setup')[0]
    # what the box reads from its page cache with plain sequential reads: the ceiling of any text path
    t = time.time()
    with open(txt, 'rb', buffering=0) as fh:
        buf = bytearray(16 << 20)
        while fh.readinto(buf):
            pass
    read_gbs = size_txt / (time.time() - t) / 1e9
    cache = os.path.join(d, 'kmers.bits')
    res = {'n_samples': n, 'kmers': m, 'parser_threads': cores, 'text_bytes': size_txt, 'setup_s': setup_s,
           'generate_text_s': gen_s,
           'page_cache_read_GBps': read_gbs, 'lines_per_s_at_that_rate': read_gbs * 1e9 / (size_txt / m), 'runs': {}}
    # device parser (psb_submit_text, the default) and host parser (PYSEER_B200_TEXT=0) on each format
    for tag, extra, text in (('plain_text_device_parser', ['--kmers', txt, '--uncompressed'], '1'),
                             ('bgzip_text_device_parser', ['--kmers', txt + '.bgz'], '1'),
                             ('gzip_text_device_parser', ['--kmers', txt + '.gz'], '1'),
                             ('gzip_one_member_device_parser', ['--kmers', txt + '.gz1'], '1'),
                             ('gzip_one_member_zlib_device_parser', ['--kmers', txt + '.gz1'], 'zlib'),
                             ('plain_text_device_parser_output_patterns',
                              ['--kmers', txt, '--uncompressed', '--output-patterns', os.path.join(d, 'pat1.txt')], '1'),
                             ('plain_text_host_parser_output_patterns',
                              ['--kmers', txt, '--uncompressed', '--output-patterns', os.path.join(d, 'pat0.txt')], '0'),
                             ('plain_text_host_parser', ['--kmers', txt, '--uncompressed'], '0'),
                             ('bgzip_text_host_parser', ['--kmers', txt + '.bgz'], '0'),
                             ('gzip_text_host_parser', ['--kmers', txt + '.gz'], '0'),
                             ('gzip_one_member_writing_bits_cache', ['--kmers', txt + '.gz1', '--bits-cache', cache], '1'),
                             ('bits_cache', ['--kmers', txt + '.gz1', '--bits-cache', cache], '1')):
        w, rate = run(extra, tag, text)
        res['runs'][tag] = {'wall_s': w, 'variants_per_s_whole_run': m / w,
                            'variants_per_s_streaming': rate}
        sys.stderr.write('%s: %.2f s, streaming %s variants/s\n' % (tag, w, rate))
    if a.sweep:
        res['gpu_batch_sweep'] = {}
        for gb in [int(x) for x in a.sweep.split(',')]:
            for tag, extra in (('plain_text_device_parser', ['--kmers', txt, '--uncompressed']),
                               ('gzip_one_member_device_parser', ['--kmers', txt + '.gz1'])):
                w, rate = run(extra + ['--gpu-batch', str(gb)], tag, '1')
                res['gpu_batch_sweep']['%s@%d' % (tag, gb)] = {'wall_s': w, 'variants_per_s_streaming': rate}
                sys.stderr.write('%s gpu-batch %d: %.2f s, streaming %s variants/s\n' % (tag, gb, w, rate))
    if a.sweep:
        # the inflater alone (psb_pgz_selftest into a 256 MB buffer), phases on stderr
        import ctypes
        from pyseer_b200 import _lib
        lib = _lib.load()
        os.environ['PSB_PGZ_TIMES'] = '1'
        os.environ['PSB_PGZ_SELFTEST_BUF'] = str(256 << 20)
        res['inflate_only'] = {}
        for th in (1, 4, 8, 16):
            ln, st = ctypes.c_int64(), (ctypes.c_int64 * 2)()
            t = time.time()
            rc = lib.psb_pgz_selftest((txt + '.gz1').encode(), th, 0, None, ctypes.byref(ln), st)
            dt = time.time() - t
            res['inflate_only']['threads_%d' % th] = {'rc': rc, 's': dt, 'GB_per_s': ln.value / dt / 1e9, 'chunks': list(st)}
            sys.stderr.write('inflate only, %d threads: %.2f s = %.2f GB/s %s\n' % (th, dt, ln.value / dt / 1e9, list(st)))
    print(json.dumps(res))
    for f in os.listdir(d):
        os.unlink(os.path.join(d, f))
    os.rmdir(d)


if __name__ == '__main__':
    main()
