#!/bin/bash
# AddressSanitizer + UBSan over the chunk-parallel gzip reader (host code, no GPU): builds
# csrc/psb_pgz.cu alone with g++, inflates valid files (k-mer text, stored blocks, runs, members) on
# 1 / 4 / 8 threads with 4 KiB .. 1 MiB chunks and 150 damaged / truncated copies.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
W=$(mktemp -d)
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -x c++ -shared -fPIC "$ROOT/pyseer_b200/csrc/psb_pgz.cu" -o "$W/libpgz_asan.so" -lz -lpthread
cat > "$W/run.py" <<'PY'
import ctypes, gzip, os, random, sys
import numpy as np
W = sys.argv[1]
lib = ctypes.CDLL(os.path.join(W, 'libpgz_asan.so'))
lib.psb_pgz_selftest.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
def run(path, th, ch):
    ln = ctypes.c_int64(); st = (ctypes.c_int64 * 2)(); crc = ctypes.c_uint32()
    return lib.psb_pgz_selftest(path.encode(), th, ch, ctypes.cast(ctypes.byref(crc), ctypes.c_void_p), ctypes.byref(ln), st), ln.value
def kmer_text(n_lines, n_samples, seed):
    rng = np.random.RandomState(seed)
    toks = np.array(['sample_%d:1' % i for i in range(n_samples)])
    acgt = np.array(list('ACGT'))
    return ('\n'.join(''.join(rng.choice(acgt, size=31)) + ' | ' + ' '.join(toks[rng.uniform(size=n_samples) < rng.uniform(0.02, 0.98)]) for _ in range(n_lines)) + '\n').encode()
cases = {'kmers': kmer_text(1500, 700, 1), 'random': os.urandom(1 << 20), 'zeros': b'\0' * (3 << 20),
         'mixed': kmer_text(150, 300, 2) + os.urandom(300000) + kmer_text(150, 300, 3) + b'A' * 100000}
bad = 0
for name, data in cases.items():
    for level in (1, 6, 9):
        p = os.path.join(W, '%s_%d.gz' % (name, level))
        open(p, 'wb').write(gzip.compress(data, level))
        for th, ch in ((1, 0), (4, 4096), (8, 30000)):
            rc, n = run(p, th, ch)
            bad += rc != 0 or n != len(data)
p = os.path.join(W, 'multi.gz')
open(p, 'wb').write(gzip.compress(cases['kmers'], 6) + gzip.compress(cases['mixed'], 1) + b'\0' * 64)
for th, ch in ((1, 0), (4, 4096), (8, 30000)):
    rc, n = run(p, th, ch)
    bad += rc != 0 or n != len(cases['kmers']) + len(cases['mixed'])
print('valid inputs: %d wrong' % bad)
random.seed(3)
raw, raw1 = open(os.path.join(W, 'kmers_6.gz'), 'rb').read(), open(os.path.join(W, 'mixed_1.gz'), 'rb').read()
rej = 0
for trial in range(150):
    src = bytearray(raw if trial % 2 else raw1)
    kind = trial % 3
    if kind == 0:
        for _ in range(random.randint(1, 4)):
            src[random.randrange(10, len(src))] ^= 1 << random.randrange(8)
    elif kind == 1:
        a = random.randrange(10, len(src)); b = min(len(src), a + random.randrange(1, 5000))
        src[a:b] = os.urandom(b - a)
    else:
        src = src[:random.randrange(20, len(src))]
    f = os.path.join(W, 'fz.gz')
    open(f, 'wb').write(src)
    rc, n = run(f, random.choice([1, 4, 8]), random.choice([0, 4096, 30000]))
    rej += rc != 0
print('damaged inputs: %d of 150 rejected' % rej)
PY
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python "$W/run.py" "$W" 2>&1 | grep -v "^psb_pgz:"
rm -rf "$W"
