"""k-mer text tokenised on the device (psb_submit_text, csrc/psb_text.cu) against the host reader
(psb_reader_next, itself pinned to the reference's read_variant in test_host_goldens.py /
test_reader_cpu.py): bit-identical packed rows and flags, on the reference's k-mer fixture and on
adversarial synthetic lines; the CLI prints the same bytes either way."""
import contextlib
import gzip
import io
import os

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _engine(n):
    """a context with a model set up on n samples (the text parser only needs N)"""
    from pyseer_b200.engine import Engine
    rng = np.random.RandomState(0)
    y = rng.normal(size=n)
    Z = np.ones((n, 1))
    eng = Engine(0)
    eng.fixed_setup(Z, y, True, 0.0, 0.0)
    return eng


def _host_rows(path, p, size):
    from pyseer_b200.input import VariantReader
    rd = VariantReader('kmers', path, p)
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        bs = list(rd.batches(size))
    rd.close()
    return [n for b in bs for n in b.names], np.concatenate([b.bits for b in bs]), err.getvalue()


def _device_rows(path, p, size, threads=1, text_cap=None):
    from pyseer_b200.input import VariantReader
    rd = VariantReader('kmers', path, p, threads=threads)
    eng = _engine(len(p))
    eng.text_setup(rd.samples)
    names, rows, infos = [], [], []
    for b in rd.text_batches(size, text_cap=text_cap):
        text, nb, ls, ll = b.text
        eng.submit_text(text, nb, ls, ll, b.n)
        bits, miss = eng.download_rows()
        assert miss is None and bits.shape == (b.n, rd.W)
        infos.append(eng.text_info(b.n))     # download_rows made the batch the current one
        names += b.names
        rows.append(bits)
    rd.close()
    eng.close()
    return names, np.concatenate(rows), np.concatenate(infos)


def test_fixture_rows_equal_host_reader():
    from pyseer_b200.input import load_phenotypes
    p = load_phenotypes(os.path.join(GOLDEN, 'subset.pheno'), None)
    src = os.path.join(GOLDEN, 'kmers.gz')
    hn, hb, herr = _host_rows(src, p, 64)
    for threads in (1, 3):
        dn, db, info = _device_rows(src, p, 64, threads)
        assert dn == hn and np.array_equal(db, hb)
        none = ['No observations of ' + hn[i] + ' in selected samples' for i in np.nonzero(info & 2)[0]]
        assert none == [l for l in herr.split('\n') if l]
        assert not (info & 4).any()


@pytest.mark.parametrize('n', [50, 333, 5000])
def test_adversarial_lines(tmp_path, n):
    """names that are prefixes of each other, unknown samples, repeated samples, tabs, tokens without
    ':', a long name, empty sample lists, leading blanks, CRLF, no final newline"""
    rng = np.random.RandomState(n)
    samples = ['s%d' % i for i in range(n - 3)] + ['s', 'x' * 70, 'sample:odd'][:3]
    samples[-1] = 'q_1'
    p = pd.Series(rng.uniform(size=n), index=samples)
    lines = []
    for v in range(700):
        af = rng.uniform(0, 1) ** 2
        pick = [samples[i] for i in np.nonzero(rng.uniform(size=n) < af)[0]]
        toks = []
        for s in pick:
            toks.append(s + (':%d' % rng.randint(1, 99) if rng.uniform() < 0.9 else ''))
            if rng.uniform() < 0.02:
                toks.append('unknown%d:1' % rng.randint(1000))
            if rng.uniform() < 0.02:
                toks.append(s + ':7')                       # listed twice
            if rng.uniform() < 0.01:
                toks.append(s + 'x:1')                      # a known name plus a suffix: unknown
        rng.shuffle(toks)
        sep = '\t' if v % 7 == 0 else ' '
        kmer = ''.join(rng.choice(list('ACGT'), size=rng.randint(9, 100)))
        lead = '  ' if v % 11 == 0 else ''
        lines.append(lead + kmer + ' |' + ('' if v % 13 == 0 else ' ') + sep.join(toks) +
                     ('  ' if v % 5 == 0 else ''))
    text = ('\r\n' if n == 333 else '\n').join(lines)           # no newline after the last line
    path = str(tmp_path / 'adv.txt')
    with open(path, 'w') as fh:
        fh.write(text)
    hn, hb, herr = _host_rows(path, p, 256)
    dn, db, info = _device_rows(path, p, 256, threads=2)
    assert dn == hn and np.array_equal(db, hb)
    assert int((info & 2).astype(bool).sum()) == herr.count('No observations of')
    # the same through a text buffer that cuts the batches short
    dn2, db2, _ = _device_rows(path, p, 256, text_cap=len(text) // 4 + 4096)
    assert dn2 == hn and np.array_equal(db2, hb)


def test_line_without_separator(tmp_path):
    """the host half refuses it like the row reader does (PSB_ERR_ARG); the kernel flags it when it is
    handed such a line all the same"""
    from pyseer_b200 import _lib
    from pyseer_b200.input import VariantReader
    p = pd.Series([0.0, 1.0, 1.0], index=['a', 'b', 'c'])
    path = str(tmp_path / 'bad.txt')
    text = b'AAA | a:1 b:1\nCCC a:1\nGGG | c:1\n'
    with open(path, 'wb') as fh:
        fh.write(text)
    rd = VariantReader('kmers', path, p)
    with pytest.raises(_lib.PsbError, match='separator'):
        list(rd.text_batches(8))
    rd.close()
    eng = _engine(3)
    eng.text_setup(['a', 'b', 'c'])
    buf = np.frombuffer(text, dtype=np.uint8).copy()
    ls = np.array([0, 14, 22], dtype=np.int64)
    ll = np.array([13, 7, 9], dtype=np.int32)
    eng.submit_text(buf, buf.shape[0], ls, ll, 3)
    bits, _ = eng.download_rows()
    info = eng.text_info(3)
    eng.close()
    assert list(info & 4) == [0, 4, 0] and list(bits[:, 0]) == [3, 0, 4]


def test_cli_prints_the_same_with_and_without_device_parser(tmp_path, monkeypatch):
    from pyseer_b200.__main__ import main
    G = lambda f: os.path.join(GOLDEN, f)   # noqa: E731
    plain = str(tmp_path / 'kmers.txt')
    with open(plain, 'wb') as fh:
        fh.write(gzip.open(G('kmers.gz'), 'rb').read())
    outs = {}
    for mode in ('1', '0'):
        for src in (G('kmers.gz'), plain):
            for extra in (['--lmm', '--similarity', G('similarity50.tsv')],
                          ['--distances', G('distances50.tsv'), '--max-dimensions', '3', '--print-filtered']):
                monkeypatch.setenv('PYSEER_B200_TEXT', mode)
                out, err = io.StringIO(), io.StringIO()
                with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
                    rc = main(['--kmers', src, '--phenotypes', G('subset.pheno'), '--gpu-batch', '30',
                               '--block_size', '10', '--cpu', '2'] + extra +
                              (['--uncompressed'] if src == plain else []))
                assert rc == 0
                key = (src == plain, extra[0])
                outs.setdefault(key, []).append((out.getvalue(), err.getvalue()))
    for key, (a, b) in outs.items():
        assert a[0] == b[0], key
        assert sorted(a[1].split('\n')) == sorted(b[1].split('\n')), key


@pytest.mark.parametrize('n', [3, 7, 8, 9, 50, 55, 56, 57, 63, 64, 65, 333, 5000])
def test_pattern_digests_equal_host_hashes(n):
    """psb_pattern_digests (one thread per row, the 8 N byte message generated from the packed bits)
    against hashlib on the vector the reference hashes (input.py:710-723: int64 0/1, float64 with NaN
    for rows with a missing genotype) and against the host psb_hash_patterns, for every tail length of
    the last MD5 block."""
    import hashlib
    from binascii import b2a_base64
    from pyseer_b200.engine import pack_rows
    from pyseer_b200.input import hash_patterns
    rng = np.random.RandomState(n)
    nv = 300
    k = (rng.uniform(size=(nv, n)) < rng.uniform(0.05, 0.95, size=(nv, 1))).astype(float)
    k[rng.uniform(size=(nv, n)) < 0.01] = np.nan
    k[::2] = np.nan_to_num(k[::2])                      # every second row without missing genotypes
    bits, miss = pack_rows(k)
    if miss is None:
        miss = np.zeros_like(bits)
    eng = _engine(n)
    for m in (miss, None):
        kk = k if m is not None else np.nan_to_num(k)
        eng.submit(bits, m)
        dig = eng.pattern_digests()
        assert dig.shape == (nv, 16)
        for v in range(nv):
            vec = kk[v] if np.isnan(kk[v]).any() else kk[v].astype(np.int64)
            assert dig[v].tobytes() == hashlib.md5(vec.tobytes()).digest(), (n, v)
        assert b''.join(b2a_base64(d.tobytes()) for d in dig) == hash_patterns(bits, m, n)
    eng.close()


def test_cli_output_patterns_through_device_parser(tmp_path, monkeypatch):
    """--output-patterns with the k-mer text tokenised on the device (hashes from psb_pattern_digests)
    writes the same file as the host path (rows parsed and hashed on the host)."""
    from pyseer_b200.__main__ import main
    G = lambda f: os.path.join(GOLDEN, f)   # noqa: E731
    files = {}
    for mode in ('1', '0'):
        for extra in (['--lmm', '--similarity', G('similarity50.tsv')],
                      ['--distances', G('distances50.tsv'), '--max-dimensions', '3']):
            monkeypatch.setenv('PYSEER_B200_TEXT', mode)
            pat = str(tmp_path / ('patterns_%s_%s.txt' % (mode, extra[0].strip('-'))))
            out, err = io.StringIO(), io.StringIO()
            with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
                rc = main(['--kmers', G('kmers.gz'), '--phenotypes', G('subset.pheno'), '--gpu-batch', '30',
                           '--block_size', '10', '--output-patterns', pat] + extra)
            assert rc == 0
            files.setdefault(extra[0], []).append((open(pat, 'rb').read(), out.getvalue()))
    for key, (a, b) in files.items():
        assert a[0] == b[0] and len(a[0]) > 25 * 100, key
        assert a[1] == b[1], key


def test_bits_cache_written_from_device_rows(tmp_path, monkeypatch):
    """--bits-cache on a first run that tokenises the text on the device: the cache holds the rows the
    device parsed (brought back per batch) -- the same names and rows as the cache the host parser
    writes -- and a run from it prints what the parsing runs printed."""
    from pyseer_b200.__main__ import main
    from pyseer_b200.input import CachedVariantReader, load_phenotypes
    G = lambda f: os.path.join(GOLDEN, f)   # noqa: E731
    p = load_phenotypes(G('subset.pheno'), None)
    outs, caches = {}, {}
    for mode in ('1', '0', 'read'):
        monkeypatch.setenv('PYSEER_B200_TEXT', '1' if mode == 'read' else mode)
        cache = str(tmp_path / ('k%s.bits' % ('1' if mode == 'read' else mode)))
        out, err = io.StringIO(), io.StringIO()
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
            rc = main(['--kmers', G('kmers.gz'), '--phenotypes', G('subset.pheno'), '--lmm', '--similarity',
                       G('similarity50.tsv'), '--gpu-batch', '60', '--block_size', '20', '--bits-cache', cache])
        assert rc == 0
        assert ('Reading packed variants from' in err.getvalue()) == (mode == 'read')
        outs[mode] = out.getvalue()
        caches[mode] = cache
    assert outs['1'] == outs['0'] == outs['read'] and len(outs['1'].split('\n')) > 150
    rows = {}
    for mode in ('1', '0'):
        rd = CachedVariantReader(caches[mode], p)
        with contextlib.redirect_stderr(io.StringIO()):
            bs = list(rd.batches(1000))
        rd.close()
        rows[mode] = ([x for b in bs for x in b.names], np.concatenate([b.bits for b in bs]))
    assert rows['1'][0] == rows['0'][0] and len(rows['1'][0]) == 200
    assert np.array_equal(rows['1'][1], rows['0'][1])
