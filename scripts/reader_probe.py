"""Throughput of the native k-mer reader (psb_reader_*) by parser threads, and of the packed cache."""
import contextlib, io, os, sys, time
import numpy as np, pandas as pd
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyseer_b200.input import VariantReader, open_variants

n, nv = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.RandomState(0)
samples = ['s%d' % i for i in range(n)]
p = pd.Series(np.zeros(n), index=samples)
path = '/tmp/probe_kmers_%d.txt' % n
with open(path, 'w') as f:
    for v in range(nv):
        on = np.nonzero(rng.uniform(size=n) < rng.uniform(0.02, 0.98))[0]
        f.write('K%d | ' % v + ' '.join('s%d:1' % i for i in on) + '\n')
os.system('gzip -kf1 ' + path)
print('file %.1f MB, %d cores' % (os.path.getsize(path) / 1e6, len(os.sched_getaffinity(0))))
for src, unc in ((path, True), (path + '.gz', False)):
    for th in (1, 2, 4, 8, 16):
        t0 = time.time()
        with contextlib.redirect_stderr(io.StringIO()):
            rd = VariantReader('kmers', src, p, uncompressed=unc, threads=th)
            tot = sum(b.n for b in rd.batches(1000))
            rd.close()
        print('%s threads=%d: %.0f variants/s' % ('gz' if not unc else 'txt', th, tot / (time.time() - t0)), flush=True)
cache = path + '.bits'
for tag in ('write', 'read'):
    t0 = time.time()
    with contextlib.redirect_stderr(io.StringIO()):
        rd = open_variants('kmers', path, p, uncompressed=True, cache=cache, threads=8)
        tot = sum(b.n for b in rd.batches(1000))
        rd.close()
    print('cache %s: %.0f variants/s' % (tag, tot / (time.time() - t0)))
