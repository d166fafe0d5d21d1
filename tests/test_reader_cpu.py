"""The native k-mer / Rtab reader (psb_reader_*, host code of the shared library) against a
plain-Python parse of the same files that follows input.read_variant (input.py:377-452)."""
import gzip
import io
import os
import contextlib

import numpy as np
import pandas as pd

from conftest import GOLDEN
from pyseer_b200.engine import unpack_rows
from pyseer_b200.input import VariantReader, hash_pattern, load_phenotypes


def _pheno():
    return load_phenotypes(os.path.join(GOLDEN, 'subset.pheno'), None)


def test_kmers_reader_matches_python_parse():
    p = _pheno()
    samples = list(p.index)
    rd = VariantReader('kmers', os.path.join(GOLDEN, 'kmers.gz'), p)
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        batches = list(rd.batches(64))
    rd.close()
    names = [n for b in batches for n in b.names]
    x = np.concatenate([unpack_rows(b.bits, len(samples)) for b in batches])
    assert [b.n for b in batches] == [64, 64, 64, 8] and all(b.missing is None for b in batches)
    none_seen = []
    with gzip.open(os.path.join(GOLDEN, 'kmers.gz'), 'rt') as fh:
        for i, line in enumerate(fh):
            name = line.split()[0]
            d = {t.split(':')[0] for t in line.rstrip().split('|')[1].split()}
            k = np.array([1 if s in d else 0 for s in samples])
            assert names[i] == name
            assert np.array_equal(x[i], k), name
            if k.sum() == 0:
                none_seen.append(name)
    assert i + 1 == len(names) == 200
    assert err.getvalue().count('No observations of') == len(none_seen) == 6


def test_rtab_reader_and_missing(tmp_path):
    p = _pheno()
    samples = list(p.index)
    rd = VariantReader('Rtab', os.path.join(GOLDEN, 'presence_absence.Rtab.gz'), p)
    with contextlib.redirect_stderr(io.StringIO()):
        batches = list(rd.batches(1000))
    rd.close()
    tab = pd.read_csv(os.path.join(GOLDEN, 'presence_absence.Rtab.gz'), sep='\t', index_col=0)
    tab = tab[samples]
    names = [n for b in batches for n in b.names]
    x = np.concatenate([unpack_rows(b.bits, len(samples)) for b in batches])
    assert names == [str(s) for s in tab.index]
    assert np.array_equal(x, (tab.values == 1).astype(np.uint8))
    # a small file with missing cells, plain text
    f = tmp_path / 'm.Rtab'
    f.write_text('Gene\t' + '\t'.join(samples[:4]) + '\textra\n'
                 'g1\t1\t0\t.\t1\t1\n'
                 'g2\t0\t0\t0\t0\t1\n'
                 'g3\t1\t\t1\t0\t0\n')
    rd = VariantReader('Rtab', str(f), p)
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        (b,) = list(rd.batches(10))
    assert b.names == ['g1', 'g2', 'g3'] and b.missing is not None
    xb, mb = unpack_rows(b.bits, len(samples)), unpack_rows(b.missing, len(samples))
    assert xb[0, :4].tolist() == [1, 0, 0, 1] and mb[0, :4].tolist() == [0, 0, 1, 0]
    assert xb[1].sum() == 0 and 'No observations of g2' in err.getvalue()
    assert xb[2, :4].tolist() == [1, 0, 1, 0] and mb[2, :4].tolist() == [0, 1, 0, 0]
    k = rd.k_vector(b, 0)
    assert k.dtype == np.float64 and np.isnan(k[2]) and k[0] == 1
    ks, nks = rd.sample_lists(b, 0)
    assert set(ks) == {samples[0], samples[2], samples[3]}      # missing counts as a carrier
    assert rd.k_vector(b, 1).dtype == np.int64
    rd.close()


def test_hash_pattern_golden(goldens):
    """input.py:710-723 / tests/input_test.py: MD5 of the 8-byte-per-sample image, base64."""
    assert hash_pattern(np.ones(50, dtype=np.int64)) == b'xxPKpdegG31U5Mx9EHcYXg==\n'
    # tests/input_test.py:840-845: the 'binary' column of subset.pheno
    pb = pd.read_csv(os.path.join(GOLDEN, 'subset.pheno'), index_col=0, sep='\t')['binary']
    assert hash_pattern(pb.values) == goldens['hash_pattern_k'].encode()


def test_packed_cache_roundtrip(tmp_path):
    """--bits-cache: the first pass parses and writes the cache, later passes stream the packed
    rows back (any batch size), and a different sample order invalidates it."""
    from pyseer_b200.input import open_variants, CachedVariantReader, PackedCache
    p = _pheno()
    src = os.path.join(GOLDEN, 'kmers.gz')
    cache = str(tmp_path / 'kmers.bits')
    with contextlib.redirect_stderr(io.StringIO()):
        rd = open_variants('kmers', src, p, cache=cache)
        assert isinstance(rd, VariantReader)
        first = list(rd.batches(64))
        rd.close()
    assert PackedCache.valid(cache, 'kmers', src, [str(s) for s in p.index], first[0].bits.shape[1])
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        rd2 = open_variants('kmers', src, p, cache=cache)
        assert isinstance(rd2, CachedVariantReader)
        second = list(rd2.batches(30))
        rd2.close()
    assert err.getvalue().count('No observations of') == 6
    assert [b.n for b in second] == [30] * 6 + [20]
    assert [n for b in first for n in b.names] == [n for b in second for n in b.names]
    assert np.array_equal(np.concatenate([b.bits for b in first]), np.concatenate([b.bits for b in second]))
    assert all(b.missing is None for b in second)
    assert np.array_equal(rd2.k_vector(second[0], 3), VariantReader.k_vector(rd2, first[0], 3))
    # another sample order: the cache does not apply and is rewritten
    p2 = p.iloc[::-1]
    with contextlib.redirect_stderr(io.StringIO()):
        rd3 = open_variants('kmers', src, p2, cache=cache)
        assert isinstance(rd3, VariantReader)
        third = list(rd3.batches(200))
        rd3.close()
    n = len(p.index)
    assert np.array_equal(unpack_rows(third[0].bits, n)[:, ::-1], unpack_rows(np.concatenate([b.bits for b in first]), n))
    # an interrupted write leaves no usable cache
    cache2 = str(tmp_path / 'partial.bits')
    with contextlib.redirect_stderr(io.StringIO()):
        rd4 = open_variants('kmers', src, p, cache=cache2)
        it = rd4.batches(64)
        next(it)
        it.close()
        rd4.close()
    assert not os.path.exists(cache2)


def test_packed_cache_with_missing(tmp_path):
    from pyseer_b200.input import open_variants, CachedVariantReader
    p = _pheno()
    samples = list(p.index)
    f = tmp_path / 'm.Rtab'
    f.write_text('Gene\t' + '\t'.join(samples[:4]) + '\n' + 'g1\t1\t0\t.\t1\n' + 'g2\t0\t0\t0\t0\n' + 'g3\t1\t1\t1\t.\n')
    cache = str(tmp_path / 'm.bits')
    with contextlib.redirect_stderr(io.StringIO()):
        a = list(open_variants('Rtab', str(f), p, uncompressed=True, cache=cache).batches(2))
        rd = open_variants('Rtab', str(f), p, uncompressed=True, cache=cache)
        assert isinstance(rd, CachedVariantReader)
        b = list(rd.batches(3))
    assert [x for t in a for x in t.names] == b[0].names == ['g1', 'g2', 'g3']
    assert np.array_equal(np.concatenate([t.bits for t in a]), b[0].bits)
    ma = np.concatenate([t.missing if t.missing is not None else np.zeros_like(t.bits) for t in a])
    assert b[0].missing is not None and np.array_equal(ma, b[0].missing)


def test_parser_threads_give_identical_rows(tmp_path):
    """--cpu N: the lines of a batch are parsed by N threads; rows, names, flags do not change."""
    rng = np.random.RandomState(4)
    n, nv = 700, 900
    samples = ['s%d' % i for i in range(n)]
    p = pd.Series(np.zeros(n), index=samples)
    f = tmp_path / 'k.txt'
    with open(f, 'w') as fh:
        for v in range(nv):
            on = np.nonzero(rng.uniform(size=n) < rng.uniform(0.0, 0.9))[0] if v % 50 else []
            extra = ' other:1' if v % 3 == 0 else ''
            fh.write('K%d | %s%s\n' % (v, ' '.join('s%d:1' % i for i in on), extra))
    outs = []
    for threads in (1, 5):
        rd = VariantReader('kmers', str(f), p, uncompressed=True, threads=threads)
        err = io.StringIO()
        with contextlib.redirect_stderr(err):
            bs = list(rd.batches(256))
        rd.close()
        outs.append(([x for b in bs for x in b.names], np.concatenate([b.bits for b in bs]),
                     err.getvalue()))
    assert outs[0][0] == outs[1][0] == ['K%d' % v for v in range(nv)]
    assert np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2] and outs[0][2].count('No observations of') == nv // 50
    # malformed line: the error surfaces from a worker thread as from the serial path
    g = tmp_path / 'bad.txt'
    g.write_text(''.join('K%d | s1:1\n' % v for v in range(40)) + 'K40 s1:1\n')
    from pyseer_b200._lib import PsbError
    import pytest
    for threads in (1, 4):
        rd = VariantReader('kmers', str(g), p, uncompressed=True, threads=threads)
        with pytest.raises(PsbError, match="without '\\|' separator"):
            list(rd.batches(64))
        rd.close()


def test_vcf_parser_threads_give_identical_batches():
    from pyseer_b200.input import VcfReader
    p = _pheno()
    outs = []
    for threads in (1, 6):
        rd = VcfReader(os.path.join(GOLDEN, 'variants50.vcf.gz'), p, threads=threads)
        err = io.StringIO()
        with contextlib.redirect_stderr(err):
            bs = list(rd.batches(250))
        rd.close()
        outs.append(([x for b in bs for x in b.names], np.concatenate([b.bits for b in bs]),
                     [b.missing is None for b in bs], err.getvalue()))
    assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2] and outs[0][3] == outs[1][3]
    assert len(outs[0][0]) == 886


def test_long_names_do_not_shrink_batches(tmp_path):
    """Batches keep their requested size whatever the length of the variant names (unitig names run
    to kilobytes): a native call that stops because its names buffer is full is continued into the
    same batch, a name longer than the whole buffer gets a larger one -- and no line is lost.  The
    LMM path walks blocks of --block_size lines from the start of every batch (lmm.py:158-226), so
    batch boundaries must not depend on the names."""
    p = _pheno()
    samples = list(p.index)
    rng = np.random.RandomState(4)
    path = str(tmp_path / 'long.txt')
    names, rows = [], []
    with open(path, 'w') as fh:
        for i in range(700):
            ln = 70000 if i == 333 else int(rng.randint(150, 260))
            nm = ''.join(rng.choice(list('ACGT'), ln)) + str(i)
            carriers = [s for s in samples if rng.uniform() < 0.4]
            fh.write(nm + ' | ' + ' '.join(s + ':1' for s in carriers) + '\n')
            names.append(nm)
            rows.append(set(carriers))
    rd = VariantReader('kmers', path, p)
    batches = list(rd.batches(300, names_cap=4096))        # ~20 names per native call
    rd.close()
    assert [b.n for b in batches] == [300, 300, 100]
    got = [n for b in batches for n in b.names]
    assert got == names
    x = np.concatenate([unpack_rows(b.bits, len(samples)) for b in batches])
    for i in (0, 299, 300, 333, 334, 699):
        assert set(s for s, on in zip(samples, x[i]) if on) == rows[i]


def test_vcf_short_and_empty_genotype_cells(tmp_path):
    """A sample cell with fewer ':' fields than FORMAT, an empty cell or a trailing '/': pysam reports
    None for such a haplotype, which read_vcf_var treats like '.' (input.py:484-497) -- missing, unless
    another haplotype of the cell is called."""
    from pyseer_b200.input import VcfReader
    p = pd.Series([0.0, 1.0, 0.0, 1.0, 1.0], index=['a', 'b', 'c', 'd', 'e'])
    vcf = str(tmp_path / 't.vcf')
    with open(vcf, 'w') as fh:
        fh.write('##fileformat=VCFv4.2\n')
        fh.write('#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\tc\td\te\n')
        fh.write('chr\t10\t.\tA\tT\t.\tPASS\t.\tDP:GT\t3\t3:1\t3:0\t3:\t3:1/\n')
    rd = VcfReader(vcf, p)
    b = list(rd.batches(10))[0]
    rd.close()
    x = unpack_rows(b.bits, 5)[0]
    m = unpack_rows(b.missing, 5)[0]
    assert list(x) == [0, 1, 0, 0, 1]          # e: '1/' carries through its called haplotype
    assert list(m) == [1, 0, 0, 1, 0]          # a: no GT field at all; d: empty GT


def _write_bgzf(path, data, block=60000):
    """bgzip-style file: independent gzip members with the 'BC' extra field (SAM spec 4.1)."""
    import struct
    import zlib
    with open(path, 'wb') as fh:
        for o in list(range(0, len(data), block)) + [None]:
            chunk = b'' if o is None else data[o:o + block]
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            payload = co.compress(chunk) + co.flush()
            bsize = 12 + 6 + len(payload) + 8
            fh.write(b'\x1f\x8b\x08\x04' + b'\x00' * 4 + b'\x00\xff' + struct.pack('<H', 6) +
                     b'BC' + struct.pack('<HH', 2, bsize - 1) + payload +
                     struct.pack('<II', zlib.crc32(chunk) & 0xffffffff, len(chunk)))


def test_bgzf_input_equals_gzip_input(tmp_path):
    """bgzip'ed k-mer files are inflated block-parallel on the parser threads; rows, names and batch
    boundaries equal those of the plain-gzip file, whatever the thread count (lines straddle blocks)."""
    p = _pheno()
    with gzip.open(os.path.join(GOLDEN, 'kmers.gz'), 'rb') as fh:
        text = fh.read()
    text = text * 8                                    # 1600 lines, many blocks
    bg = str(tmp_path / 'kmers.bgz')
    _write_bgzf(bg, text, block=7001)
    gz = str(tmp_path / 'kmers.gz')
    with gzip.open(gz, 'wb') as fh:
        fh.write(text)
    out = {}
    for tag, path, threads in (('gz', gz, 1), ('bg1', bg, 1), ('bg5', bg, 5)):
        rd = VariantReader('kmers', path, p, threads=threads)
        with contextlib.redirect_stderr(io.StringIO()):
            out[tag] = [(b.names, b.bits.copy()) for b in rd.batches(300)]
        rd.close()
    assert len(out['gz']) == 6
    for tag in ('bg1', 'bg5'):
        assert len(out[tag]) == len(out['gz'])
        for (n1, b1), (n2, b2) in zip(out['gz'], out[tag]):
            assert n1 == n2 and np.array_equal(b1, b2)


def test_truncated_cache_is_rejected(tmp_path):
    """A cache cut off in the middle -- even where the cut leaves zero bytes at the end -- is not
    used: the closing chunk carries the row count and a trailer word."""
    from pyseer_b200.input import PackedCache, open_variants
    p = _pheno()
    src = os.path.join(GOLDEN, 'kmers.gz')
    cache = str(tmp_path / 'k.bits')
    rd = open_variants('kmers', src, p, cache=cache)
    with contextlib.redirect_stderr(io.StringIO()):
        n = sum(b.n for b in rd.batches(64))
    rd.close()
    samples = [str(s) for s in p.index]
    W = rd.W
    assert n == 200 and PackedCache.valid(cache, 'kmers', src, samples, W)
    data = open(cache, 'rb').read()
    cut = str(tmp_path / 'cut.bits')
    with open(cut, 'wb') as fh:
        fh.write(data[:len(data) // 2] + b'\0' * 64)
    assert not PackedCache.valid(cut, 'kmers', src, samples, W)


def _text_lines(rd, size, block=1, text_cap=None):
    """(names, lines) of every batch of the text mode, lines as bytes cut out of the text buffer"""
    out = []
    for b in rd.text_batches(size, block_size=block, text_cap=text_cap):
        text, nb, ls, ll = b.text
        assert b.bits is None and nb <= text.shape[0]
        lines = [bytes(text[ls[i]:ls[i] + ll[i]]) for i in range(b.n)]
        assert all(ls[i] + ll[i] <= nb for i in range(b.n))
        out.append((list(b.names), lines))
    return out


def test_text_mode_cuts_the_same_lines(tmp_path):
    """psb_reader_next_text (the host half of the device parser): same names, same line order and
    batch sizes as the row reader, lines trimmed like it does -- for gzip, bgzip-style and plain
    input, one thread and several, and with a text buffer that forces short batches."""
    p = _pheno()
    src = os.path.join(GOLDEN, 'kmers.gz')
    raw = gzip.open(src, 'rb').read()
    ref_lines = [l.rstrip(b'\r \t') for l in raw.split(b'\n')]
    ref_lines = [l for l in ref_lines if l]
    plain = str(tmp_path / 'k.txt')
    with open(plain, 'wb') as fh:
        fh.write(raw.rstrip(b'\n'))                 # last line without a newline
    crlf = str(tmp_path / 'k_crlf.txt')
    with open(crlf, 'wb') as fh:
        fh.write(raw.replace(b'\n', b' \r\n\r\n'))  # trailing blanks, CRLF and empty lines
    for path in (src, plain, crlf):
        for threads in (1, 4):
            rd = VariantReader('kmers', path, p, threads=threads)
            got = _text_lines(rd, 64)
            rd.close()
            assert [len(n) for n, _ in got] == [64, 64, 64, 8]
            lines = [l for _, ls in got for l in ls]
            names = [x for n, _ in got for x in n]
            assert lines == ref_lines
            assert names == [l.split()[0].decode() for l in ref_lines]
    # a text buffer that holds a fifth of the file: batches shrink to whole blocks of 8, nothing is lost
    rd = VariantReader('kmers', plain, p)
    got = _text_lines(rd, 64, block=8, text_cap=max(len(raw) // 5, 9 * max(len(l) for l in ref_lines)))
    rd.close()
    sizes = [len(n) for n, _ in got]
    assert sum(sizes) == 200 and all(s % 8 == 0 for s in sizes[:-1]) and max(sizes) < 64
    assert [l for _, ls in got for l in ls] == ref_lines


def test_text_mode_threads_on_a_large_file(tmp_path):
    """Reads and newline scans split over threads (slices of >= 4 MB): same lines as one thread."""
    rng = np.random.RandomState(3)
    n = 600
    samples = ['sample_%04d' % i for i in range(n)]
    p = pd.Series(rng.uniform(size=n), index=samples)
    toks = np.array([s + ':1' for s in samples])
    path = str(tmp_path / 'big.txt')
    with open(path, 'w') as fh:
        for v in range(9000):
            fh.write('K%06d | ' % v + ' '.join(toks[rng.uniform(size=n) < rng.uniform(0.05, 0.95)]) + '\n')
    assert os.path.getsize(path) > (24 << 20)
    out = {}
    for threads in (1, 5):
        rd = VariantReader('kmers', path, p, threads=threads)
        out[threads] = _text_lines(rd, 4000, block=100)
        rd.close()
    assert [len(a) for a, _ in out[1]] == [4000, 4000, 1000]
    assert out[1] == out[5]
    with open(path, 'rb') as fh:
        assert [l for _, ls in out[5] for l in ls] == [l.rstrip() for l in fh.read().split(b'\n') if l]


def test_text_mode_keeps_its_text_when_a_buffer_is_too_small(tmp_path):
    """PSB_ERR_NOMEM (here: a names buffer shorter than the first name) loses nothing: the text read so
    far opens the next call, which Python makes with a larger buffer."""
    p = _pheno()
    raw = gzip.open(os.path.join(GOLDEN, 'kmers.gz'), 'rb').read()
    ref = [l.split()[0].decode() for l in raw.split(b'\n') if l.strip()]
    rd = VariantReader('kmers', os.path.join(GOLDEN, 'kmers.gz'), p)
    names = [x for b in rd.text_batches(64, names_cap=8) for x in b.names]
    rd.close()
    assert names == ref


# ---- plain gzip inflated on several threads (csrc/psb_pgz.cu) ----

def _pgz(path, threads, chunk):
    import ctypes
    from pyseer_b200 import _lib
    lib = _lib.load()
    crc, n, st = ctypes.c_uint32(), ctypes.c_int64(), (ctypes.c_int64 * 2)()
    rc = lib.psb_pgz_selftest(path.encode(), threads, chunk, ctypes.cast(ctypes.byref(crc), ctypes.c_void_p),
                              ctypes.byref(n), st)
    return rc, crc.value, n.value, list(st)


def _kmer_text(n_lines, n_samples, seed):
    rng = np.random.RandomState(seed)
    toks = np.array(['sample_%d:1' % i for i in range(n_samples)])
    acgt = np.array(list('ACGT'))
    return ('\n'.join(''.join(rng.choice(acgt, size=31)) + ' | ' +
                      ' '.join(toks[rng.uniform(size=n_samples) < rng.uniform(0.02, 0.98)])
                      for _ in range(n_lines)) + '\n').encode()


def test_parallel_gzip_equals_zlib(tmp_path):
    """The chunk-parallel inflater against zlib's CRC-32 and length: k-mer text, incompressible bytes
    (stored blocks), runs (long matches at distance 1), a mix, the empty and a tiny stream; compression
    levels 1 / 6 / 9; one thread and several; chunks far smaller than a deflate block (most chunks find
    no block start and are decoded by their predecessor) up to one chunk for the whole file."""
    import zlib
    cases = {'kmers': _kmer_text(700, 700, 1), 'random': os.urandom(1 << 20), 'zeros': b'\0' * (3 << 20),
             'mixed': _kmer_text(150, 300, 2) + os.urandom(300000) + _kmer_text(150, 300, 3) + b'A' * 100000,
             'empty': b'', 'tiny': b'hello\n'}
    for name, data in cases.items():
        want = (zlib.crc32(data) & 0xffffffff, len(data))
        for level in (1, 6, 9):
            path = str(tmp_path / ('%s_%d.gz' % (name, level)))
            with open(path, 'wb') as fh:
                fh.write(gzip.compress(data, level))
            for threads, chunk in ((1, 0), (4, 4096), (8, 65536), (5, 20000)):
                rc, crc, n, st = _pgz(path, threads, chunk)
                assert rc == 0 and (crc, n) == want, (name, level, threads, chunk, rc, n)
    # chunks of 64 KiB on the k-mer text: block starts are found and confirmed (work is shared)
    rc, crc, n, st = _pgz(str(tmp_path / 'kmers_6.gz'), 4, 65536)
    assert st[0] >= 4 and st[1] <= st[0] // 2, st


def test_parallel_gzip_falls_back_to_zlib(tmp_path, monkeypatch):
    """A chunk the chain needs but that cannot be decoded here (test hook: a symbol budget of 200 k per
    chunk) hands the stream to zlib at the last confirmed block boundary -- mid-member, at a bit offset,
    with the 32 KiB window -- and the text, CRC-32 and length still come out right."""
    import zlib
    monkeypatch.setenv('PSB_PGZ_MAX_SYMS', '200000')
    for name, data in (('kmers', _kmer_text(700, 700, 1)), ('zeros', b'\0' * (3 << 20)),
                       ('two', None)):
        path = str(tmp_path / (name + '.gz'))
        if data is None:
            a, b = _kmer_text(200, 300, 8), _kmer_text(300, 200, 9)
            data = a + b
            with open(path, 'wb') as fh:
                fh.write(gzip.compress(a, 6) + gzip.compress(b, 9))
        else:
            with open(path, 'wb') as fh:
                fh.write(gzip.compress(data, 6))
        for threads, chunk in ((1, 0), (4, 8192), (8, 65536)):
            rc, crc, n, _ = _pgz(path, threads, chunk)
            assert rc == 0 and (crc, n) == (zlib.crc32(data) & 0xffffffff, len(data)), (name, threads, chunk, rc, n)
    # the switch in the middle of a member: k-mer text decodes within a budget of 1 M symbols per chunk,
    # the run of zeros behind it does not
    monkeypatch.setenv('PSB_PGZ_MAX_SYMS', '1000000')
    data = _kmer_text(700, 700, 1) + b'\0' * (8 << 20) + _kmer_text(100, 100, 2)
    path = str(tmp_path / 'mid.gz')
    with open(path, 'wb') as fh:
        fh.write(gzip.compress(data, 6))
    for threads, chunk in ((2, 30000), (8, 16384)):
        rc, crc, n, _ = _pgz(path, threads, chunk)
        assert rc == 0 and (crc, n) == (zlib.crc32(data) & 0xffffffff, len(data))
    bad = bytearray(gzip.compress(_kmer_text(400, 300, 4), 6))
    bad[len(bad) // 2] ^= 0x10
    with open(str(tmp_path / 'bad.gz'), 'wb') as fh:
        fh.write(bad)
    with contextlib.redirect_stderr(io.StringIO()):
        assert _pgz(str(tmp_path / 'bad.gz'), 4, 8192)[0] < 0


def test_parallel_gzip_members_header_fields_and_corruption(tmp_path):
    import subprocess
    import zlib
    a, b = _kmer_text(300, 400, 5), _kmer_text(200, 300, 6) + os.urandom(50000)
    multi = str(tmp_path / 'multi.gz')
    with open(multi, 'wb') as fh:                  # two members and zero padding, as `cat a.gz b.gz` gives
        fh.write(gzip.compress(a, 6) + gzip.compress(b, 1) + b'\0' * 64)
    for threads, chunk in ((1, 0), (4, 4096), (8, 65536)):
        rc, crc, n, _ = _pgz(multi, threads, chunk)
        assert rc == 0 and (crc, n) == (zlib.crc32(a + b) & 0xffffffff, len(a) + len(b))
    named = tmp_path / 'named.txt'                 # gzip(1) stores the file name in the header
    named.write_bytes(a)
    subprocess.check_call(['gzip', '-k', str(named)])
    rc, crc, n, _ = _pgz(str(named) + '.gz', 3, 30000)
    assert rc == 0 and (crc, n) == (zlib.crc32(a) & 0xffffffff, len(a))
    raw = bytearray(gzip.compress(a, 6))
    for at in (len(raw) // 3, len(raw) // 2, len(raw) - 6):       # damaged data / damaged CRC
        bad = bytearray(raw)
        bad[at] ^= 0x41
        p = str(tmp_path / ('bad%d.gz' % at))
        with open(p, 'wb') as fh:
            fh.write(bad)
        with contextlib.redirect_stderr(io.StringIO()):
            assert _pgz(p, 4, 16384)[0] < 0
    trunc = str(tmp_path / 'trunc.gz')
    with open(trunc, 'wb') as fh:
        fh.write(raw[:len(raw) // 2])
    assert _pgz(trunc, 4, 16384)[0] < 0


def test_reader_on_parallel_gzip(tmp_path, monkeypatch):
    """The row reader and the text reader over a gzip file through the parallel inflater (forced onto
    the small fixture: PSB_PGZ_MIN=0, 8 KiB chunks) give what they give through zlib (PSB_PGZ=0)."""
    p = _pheno()
    with gzip.open(os.path.join(GOLDEN, 'kmers.gz'), 'rb') as fh:
        text = fh.read() * 6
    gz = str(tmp_path / 'k.gz')
    with gzip.open(gz, 'wb') as fh:
        fh.write(text)
    out = {}
    for tag, env in (('zlib', {'PSB_PGZ': '0'}), ('pgz', {'PSB_PGZ_MIN': '0', 'PSB_PGZ_CHUNK': '8192'})):
        for k in ('PSB_PGZ', 'PSB_PGZ_MIN', 'PSB_PGZ_CHUNK'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for threads in (1, 4):
            rd = VariantReader('kmers', gz, p, threads=threads)
            with contextlib.redirect_stderr(io.StringIO()):
                rows = [(b.names, b.bits.copy()) for b in rd.batches(250)]
            rd.close()
            rd = VariantReader('kmers', gz, p, threads=threads)
            lines = _text_lines(rd, 250)
            rd.close()
            out[tag, threads] = (rows, lines)
    ref_rows, ref_lines = out['zlib', 1]
    assert sum(len(n) for n, _ in ref_rows) == 1200
    for key, (rows, lines) in out.items():
        assert lines == ref_lines, key
        assert len(rows) == len(ref_rows), key
        for (n1, b1), (n2, b2) in zip(ref_rows, rows):
            assert n1 == n2 and np.array_equal(b1, b2), key
