"""Host-side loaders and the variant reader of the CLI.

Behaviour follows pyseer/input.py (phenotypes :24-59, structure :62-137, covariates :195-247,
k-mer / Rtab lines :301-454) but the variant stream is produced as *batches of packed bit rows*
(what ``psb_submit`` takes) instead of one NumPy vector per variant: sample names are mapped to
bit positions once, and per-variant sample lists are only materialised when asked for
(``--print-samples``).
"""
import binascii
import hashlib
import json
import os
import sys

import numpy as np
import pandas as pd

from .cmdscale import cmdscale
from .engine import words_per_row


def load_phenotypes(infile, column):
    """input.py:24-59: phenotype Series indexed by sample name (last column by default)."""
    p = pd.read_csv(infile, index_col=0, sep='\t')
    if p.shape[1] < 1:
        sys.stderr.write('Phenotype file must contain at least one phenotype column\n')
        sys.exit(1)
    p.index = p.index.astype(str)
    if np.any(p.index.duplicated()):
        sys.stderr.write('Phenotype file contains duplicated sample names\n')
        sys.exit(1)
    p = p[p.columns[-1]] if column is None else p[column]
    p = p.dropna()
    if not pd.api.types.is_numeric_dtype(p.values.dtype):
        sys.stderr.write('Phenotypes must be numeric\n')
        sys.exit(1)
    return p


def load_structure(infile, p, max_dimensions, mds_type='classic', n_cpus=1, seed=None):
    """input.py:62-137: distance matrix -> MDS projection restricted to phenotyped samples,
    every component scaled to max |value| 1."""
    m = pd.read_csv(infile, index_col=0, sep='\t')
    m.index = m.index.astype(str)
    if np.any(m.index.duplicated()):
        sys.stderr.write('Structure file contains duplicated sample names\n')
        sys.exit(1)
    sys.stderr.write('Structure matrix has dimension ' + str(m.shape) + '\n')
    common = p.index.intersection(m.index).intersection(m.columns)
    m = m.loc[common, common]
    if len(common) == 0:
        sys.stderr.write('None of the phenotyped samples were found in population structure matrix\n')
        sys.exit(1)
    if mds_type == 'classic':
        projection, _ = cmdscale(m.values)
    else:
        from sklearn import manifold
        metric = mds_type != 'non-metric'
        if mds_type not in ('metric', 'non-metric'):
            sys.stderr.write('Unsupported mds type chosen. Assuming metric\n')
        try:
            mds = manifold.MDS(n_components=max_dimensions, metric_mds=metric, metric='precomputed',
                               n_jobs=n_cpus, normalized_stress='auto', random_state=seed, n_init=1)
        except TypeError:
            mds = manifold.MDS(n_components=max_dimensions, metric=metric, n_jobs=n_cpus,
                               random_state=seed, dissimilarity='precomputed')
        projection = mds.fit_transform(m.values)
    m = pd.DataFrame(projection, index=m.index)
    for i in range(m.shape[1]):
        m[i] = m[i] / max(abs(m[i]))
    return m


def load_covariates(infile, covariates, p):
    """input.py:195-247: quantitative columns as they are ('Nq'), categorical ones
    dummy-encoded with one level dropped."""
    c = pd.read_csv(infile, index_col=0, header=0, sep='\t')
    c.index = c.index.astype(str)
    if np.any(c.index.duplicated()):
        sys.stderr.write('Covariate file contains duplicated sample names\n')
        sys.exit(1)
    if len(p.index.difference(c.index)) > 0:
        sys.stderr.write('All samples with a phenotype must be present in covariate file\n')
        sys.exit(1)
    c = c.loc[p.index.intersection(c.index)]
    if covariates is None:
        return pd.DataFrame([])
    cov = []
    for col in covariates:
        cnum = int(col.rstrip('q'))
        if cnum == 1 or cnum > c.shape[1] + 1:
            sys.stderr.write('Covariates columns values should be > 1 and less than or equal to '
                             'total number of columns (%d)\n' % (c.shape[1] + 1))
            return None
        series = c.iloc[:, cnum - 2]
        if col[-1] == 'q':
            cov.append(series)
        else:
            categories = set(series)
            categories.pop()
            for i, categ in enumerate(categories):
                cov.append(pd.Series([1 if x == categ else 0 for x in series.values], index=c.index,
                                     name=c.columns[cnum - 2] + '_' + str(i)))
    return pd.concat(cov, axis=1) if len(cov) > 0 else pd.DataFrame([])


def hash_pattern(k):
    """input.py:710-723: base64 of the MD5 of the int64 (float64 when NaN present) byte image."""
    return binascii.b2a_base64(hashlib.md5(np.ascontiguousarray(k).view(np.uint8)).digest())


def hash_patterns(bits, missing, n_samples, flags=None):
    """hash_pattern of every row of a packed batch through the library (``psb_hash_patterns``):
    the concatenated 25-byte entries the result loop writes to ``--output-patterns`` for the
    tested variants (rows whose flags carry F_PREFILTER are skipped when ``flags`` is given)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    bits = np.ascontiguousarray(bits, dtype=np.uint32)
    n, W = bits.shape
    mp = None
    if missing is not None:
        missing = np.ascontiguousarray(missing, dtype=np.uint32)
        mp = missing.ctypes.data_as(ctypes.c_void_p)
    fp = None
    if flags is not None:
        flags = np.ascontiguousarray(flags, dtype=np.uint32)
        fp = flags.ctypes.data_as(ctypes.c_void_p)
    out = ctypes.create_string_buffer(25 * n + 1)
    m = ctypes.c_int64(0)
    _lib.check(lib.psb_hash_patterns(bits.ctypes.data_as(ctypes.c_void_p), mp, n, W, int(n_samples), fp,
                                     ctypes.addressof(out), ctypes.byref(m)))
    return out.raw[:25 * m.value]


class NameBlob(object):
    """The variant names of a batch as the packed cache stores them -- one blob of NUL-terminated names
    plus their offsets -- behaving like the list of strings the result loop expects; the native
    formatter takes blob and offsets as they are (48000 names per batch are otherwise split into
    strings only to be joined again)."""
    __slots__ = ['blob', 'off']

    def __init__(self, blob, off):
        self.blob, self.off = blob, off

    def __len__(self):
        return int(self.off.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        a = int(self.off[i])
        return self.blob[a:self.blob.index(b'\0', a)].decode()

    def __iter__(self):
        return iter(self.blob[:-1].decode().split('\0')) if len(self) else iter(())

    def __eq__(self, other):
        return list(self) == list(other)

    def __ne__(self, other):
        return not self == other


class VariantBatch(object):
    """``n`` variants as packed rows plus what the result loop needs to print them."""
    __slots__ = ['names', 'bits', 'missing', 'n', 'token', 'skipped', 'text', 'info', 'digests', 'report_empty']

    def __init__(self, names, bits, missing):
        self.names = names
        self.bits = bits
        self.missing = missing
        self.n = len(names)
        self.token = None          # buffer-pool token (pipeline.py), released by the consumer
        self.skipped = None        # bool mask: records the reference never hands to a model
        self.text = None           # (uint8 text, n_bytes, line_start, line_len): rows are built on the device
        self.info = None           # ... and their per-line flags (Engine.text_info), set by the runner
        self.digests = None        # ... and the MD5 digests of their patterns (Engine.pattern_digests), on request
        self.report_empty = False  # 'No observations of ...' still to be reported, from the run's carrier counts


class VariantReader(object):
    """Streams a k-mer (``name | s1:1 s2:1 ...``) or Rtab file as VariantBatch objects through
    the library's native reader (``psb_reader_*``: zlib + hash-map lookup straight into packed
    rows).  The sample order of the bit rows is the phenotype index order (``p.index``), as
    the reference builds ``k`` (input.py:450)."""

    def __init__(self, var_type, path, p, uncompressed=False, threads=1):
        import ctypes
        from . import _lib
        if var_type not in ('kmers', 'Rtab'):
            raise ValueError('unsupported variant type %s (VCF input needs pysam, which this '
                             'build does not use)' % var_type)
        self.var_type = var_type
        self.samples = [str(s) for s in p.index]
        self.n_samples = len(self.samples)
        self.W = words_per_row(self.n_samples)
        self._lib = _lib.load()
        names = (ctypes.c_char_p * self.n_samples)(*[s.encode() for s in self.samples])
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.psb_reader_open(str(path).encode(), 0 if var_type == 'kmers' else 1,
                                             names, self.n_samples, ctypes.byref(self._h)))
        if threads and threads > 1:      # pyseer's --cpu: parser threads
            _lib.check(self._lib.psb_reader_set_threads(self._h, int(threads)))

    def close(self):
        if self._h:
            self._lib.psb_reader_close(self._h)
            self._h = None

    def batches(self, size, names_cap=None, pool=None):
        """Batches of exactly ``size`` variants (the last one shorter): block boundaries of the LMM
        path (lmm.py:158-226 works on blocks of --block_size lines) must not depend on how long the
        variant names are.  A call of the native reader that stops because its names buffer is full
        is continued into the same rows; a name longer than the whole buffer gets a larger one.
        ``pool``: optional source of (pinned) row buffers, ``pool.get() -> (bits, missing_or_None,
        token)``; the token comes back in ``VariantBatch.token`` for the consumer to release."""
        import ctypes
        from . import _lib
        cap = int(names_cap or max(1 << 20, 160 * size))
        with_missing = self.var_type == 'Rtab'
        done = False
        while not done:
            token = None
            if pool is not None:
                bits, miss, token = pool.get()
                if not with_missing:
                    miss = None
            else:
                bits = np.empty((size, self.W), dtype=np.uint32)
                miss = np.empty((size, self.W), dtype=np.uint32) if with_missing else None
            off = np.empty(size, dtype=np.int64)
            info = np.empty(size, dtype=np.int32)
            nm = []
            any_missing = False
            filled = 0
            while filled < size:
                names = ctypes.create_string_buffer(cap)
                n = ctypes.c_int64(0)
                anym = ctypes.c_int32(0)
                rc = self._lib.psb_reader_next(
                    self._h, size - filled, bits[filled:].ctypes.data,
                    miss[filled:].ctypes.data if miss is not None else None, self.W,
                    ctypes.addressof(names), cap, off[filled:].ctypes.data, info[filled:].ctypes.data,
                    ctypes.byref(n), ctypes.byref(anym))
                if rc == _lib.PSB_ERR_NOMEM:
                    cap *= 4                     # a single name longer than the buffer: the line was kept
                    continue
                _lib.check(rc)
                n = n.value
                raw = names.raw
                nm.extend(raw[off[filled + i]:raw.index(b'\0', off[filled + i])].decode() for i in range(n))
                any_missing = any_missing or bool(anym.value)
                filled += n
                eof = ctypes.c_int32(0)
                _lib.check(self._lib.psb_reader_at_eof(self._h, ctypes.byref(eof)))
                if eof.value:
                    done = True
                    break
            if filled == 0:
                if pool is not None:
                    pool.put(token)
                return
            for i in np.nonzero(info[:filled] & 2)[0]:
                sys.stderr.write('No observations of ' + nm[i] + ' in selected samples\n')
            b = VariantBatch(nm, bits[:filled], miss[:filled] if (miss is not None and any_missing) else None)
            b.token = token
            yield b

    def text_batches(self, size, block_size=1, pool=None, text_cap=None, names_cap=None):
        """The same batches as TEXT for the device parser (``psb_reader_next_text`` ->
        ``Engine.submit_text``): the host only reads / inflates and cuts lines, ``VariantBatch.text``
        carries ``(text, n_bytes, line_start, line_len)`` and ``bits`` is None.  Batches hold ``size``
        lines; a batch cut short by the text buffer keeps a multiple of ``block_size`` lines.
        ``pool``: ``pool.get() -> (text, line_start, line_len, token)`` page-locked buffers."""
        import ctypes
        from . import _lib
        if self.var_type != 'kmers':
            raise ValueError('text batches exist for k-mer files only')
        ncap = int(names_cap or max(1 << 20, 160 * size))
        while True:
            token = None
            if pool is not None:
                text, lstart, llen, token = pool.get()
            else:
                cap = int(text_cap or max(64 << 20, size * (6 * self.n_samples + 256)))
                text = np.empty(cap, dtype=np.uint8)
                lstart = np.empty(size, dtype=np.int64)
                llen = np.empty(size, dtype=np.int32)
            rows = min(size, lstart.shape[0])
            off = np.empty(rows, dtype=np.int64)
            while True:
                names = ctypes.create_string_buffer(ncap)
                n, nb = ctypes.c_int64(0), ctypes.c_int64(0)
                rc = self._lib.psb_reader_next_text(
                    self._h, rows, int(block_size), text.ctypes.data, text.shape[0], lstart.ctypes.data,
                    llen.ctypes.data, ctypes.addressof(names), ncap, off.ctypes.data, ctypes.byref(n),
                    ctypes.byref(nb))
                if rc == _lib.PSB_ERR_NOMEM and ncap < (1 << 30):      # the text read so far is kept by the reader
                    ncap *= 4
                    continue
                _lib.check(rc)
                break
            n = n.value
            if n == 0:
                if pool is not None:
                    pool.put(token)
                return
            nm = names.raw[:int(off[n - 1])].decode().split('\0')[:n - 1] if n > 1 else []
            raw = names.raw
            nm.append(raw[off[n - 1]:raw.index(b'\0', off[n - 1])].decode())
            b = VariantBatch(nm, None, None)
            b.text = (text, nb.value, lstart, llen)
            b.token = token
            yield b
            eof = ctypes.c_int32(0)
            _lib.check(self._lib.psb_reader_at_eof(self._h, ctypes.byref(eof)))
            if eof.value:
                return

    # -- per-variant detail, only when needed ------------------------------------------
    def sample_lists(self, batch, j):
        """(kstrains, nkstrains) of variant j: sorted names, missing counted as carriers
        (input.py:439-440)."""
        by = batch.bits[j].view(np.uint8)
        present = np.unpackbits(by, bitorder='little')[:self.n_samples].astype(bool)
        if batch.missing is not None:
            present |= np.unpackbits(batch.missing[j].view(np.uint8), bitorder='little')[:self.n_samples].astype(bool)
        ks = sorted(s for s, on in zip(self.samples, present) if on)
        nks = sorted(s for s, on in zip(self.samples, present) if not on)
        return ks, nks

    def k_vector(self, batch, j):
        """The reference's ``k`` array of variant j: int64, or float64 with NaN when the
        variant has missing genotypes (input.py:450)."""
        x = np.unpackbits(batch.bits[j].view(np.uint8), bitorder='little')[:self.n_samples]
        if batch.missing is not None:
            m = np.unpackbits(batch.missing[j].view(np.uint8), bitorder='little')[:self.n_samples]
            if m.any():
                k = x.astype(np.float64)
                k[m.astype(bool)] = np.nan
                return k
        return x.astype(np.int64)


class PackedCache(object):
    """Pre-packed binary cache of a variant file: the packed rows the reader produced,
    stored once so that later runs on the same file and the same sample order skip text parsing
    (the reference parser manages 10^2-10^3 variants/s, input.py:301-454; the native one ~6 k/s at
    N=5000; the cache streams at disk speed).

    Layout: one JSON header line (magic, sample-order digest, source size and mtime, row width),
    then chunks ``int64 n, int64 names_bytes, int64 has_missing | names (NUL separated) | bits
    uint32[n][W] | missing uint32[n][W] if has_missing``; the closing chunk ``0, total rows, TRAILER``
    ends the file -- a cache without it (interrupted run, truncated copy) is not used."""
    MAGIC = 'pyseer_b200-bits-2'
    TRAILER = 0x7073625F62697473            # 'psb_bits': third word of the closing chunk

    @staticmethod
    def _header(var_type, source, samples, W):
        st = os.stat(source)
        digest = hashlib.sha1('\n'.join(samples).encode()).hexdigest()
        return {'magic': PackedCache.MAGIC, 'var_type': var_type, 'n_samples': len(samples), 'W': int(W),
                'samples_sha1': digest, 'source_size': st.st_size, 'source_mtime_ns': st.st_mtime_ns}

    @staticmethod
    def valid(path, var_type, source, samples, W):
        """True when ``path`` is a complete cache of ``source`` for this sample order."""
        try:
            want = PackedCache._header(var_type, source, samples, W)
            with open(path, 'rb') as fh:
                got = json.loads(fh.readline().decode())
                if got != want:
                    return False
                fh.seek(-24, os.SEEK_END)
                tail = np.frombuffer(fh.read(24), dtype='<i8')
                return tail.shape[0] == 3 and tail[0] == 0 and tail[1] >= 0 and \
                    int(tail[2]) == PackedCache.TRAILER
        except (OSError, ValueError):
            return False


class PackedCacheWriter(object):
    def __init__(self, path, var_type, source, samples, W):
        self.fh = open(path, 'wb')
        self.rows = 0
        self.fh.write((json.dumps(PackedCache._header(var_type, source, samples, W), sort_keys=True) + '\n').encode())

    def add(self, batch):
        names = ('\0'.join(batch.names) + '\0').encode()
        has_m = batch.missing is not None
        self.rows += batch.n
        self.fh.write(np.array([batch.n, len(names), int(has_m)], dtype='<i8').tobytes())
        self.fh.write(names)
        self.fh.write(np.ascontiguousarray(batch.bits, dtype='<u4').tobytes())
        if has_m:
            self.fh.write(np.ascontiguousarray(batch.missing, dtype='<u4').tobytes())

    def close(self, complete=True):
        if self.fh:
            if complete:
                self.fh.write(np.array([0, self.rows, PackedCache.TRAILER], dtype='<i8').tobytes())
            self.fh.close()
            self.fh = None


class CachedVariantReader(object):
    """Same interface as VariantReader, over a PackedCache file (no parsing, no native library)."""

    def __init__(self, path, p, threads=1):
        self.samples = [str(s) for s in p.index]
        self.n_samples = len(self.samples)
        self.W = words_per_row(self.n_samples)
        self.threads = max(1, min(int(threads), 8))
        self._tp = None
        self.fh = open(path, 'rb')
        self.header = json.loads(self.fh.readline().decode())
        self._data_pos = self.fh.tell()             # first chunk
        self.var_type = self.header['var_type']

    def close(self):
        if self._tp is not None:
            self._tp.shutdown()
            self._tp = None
        if self.fh:
            self.fh.close()
            self.fh = None

    def _pool(self):
        if self._tp is None:
            from concurrent.futures import ThreadPoolExecutor
            self._tp = ThreadPoolExecutor(self.threads)
        return self._tp

    def batches(self, size, pool=None, defer_empty=False):
        """Batches of exactly ``size`` variants (the last one shorter), the rows read STRAIGHT into the
        batch's buffers with ``pread`` -- page-locked ones when ``pool`` hands them out (``pool.get() ->
        (bits, missing_or_None, token)``), so that ``psb_submit`` is a plain DMA; no intermediate copies.
        ``defer_empty``: rows without any observation are reported by the consumer from the carrier counts
        of the run (``VariantBatch.report_empty``) instead of a pass over the rows here."""
        W, fd = self.W, self.fh.fileno()
        row_bytes = W * 4
        pos = self._data_pos
        # the chunk being consumed: rows, names, file offsets of its bits / missing rows, rows taken
        n = taken = 0
        craw, cstart, cnul, bits_off, miss_off = b'', None, None, 0, None

        def read_slice(mv, off):
            done = 0
            while done < len(mv):
                k = os.preadv(fd, [mv[done:]], off + done)
                if k <= 0:
                    raise IOError('packed cache truncated')
                done += k

        def read_into(arr, off):
            # slices of >= 4 MB on a few threads (preadv releases the GIL): one thread copies ~5 GB/s
            # out of the page cache
            mv = memoryview(arr).cast('B')
            parts = min(self.threads, max(1, len(mv) >> 22))
            if parts <= 1:
                return read_slice(mv, off)
            cut = [len(mv) * i // parts for i in range(parts + 1)]
            list(self._pool().map(lambda i: read_slice(mv[cut[i]:cut[i + 1]], off + cut[i]), range(parts)))

        at_end = False
        while not at_end:
            token = None
            if pool is not None:
                bits, miss, token = pool.get()
            else:
                bits, miss = np.empty((size, W), dtype=np.uint32), None
            pieces, offs, blob_len, filled, any_m = [], [], 0, 0, False
            while filled < size:
                if taken == n:
                    head = np.frombuffer(os.pread(fd, 24, pos), dtype='<i8')
                    if head.shape[0] < 3 or head[0] == 0:
                        at_end = True
                        break
                    n, nb, has_m = int(head[0]), int(head[1]), int(head[2])
                    craw = os.pread(fd, nb, pos + 24)
                    cnul = np.flatnonzero(np.frombuffer(craw, dtype=np.uint8) == 0)[:n].astype(np.int64)
                    if cnul.shape[0] != n:
                        raise IOError('packed cache: names of a chunk are damaged')
                    cstart = np.concatenate(([0], cnul[:-1] + 1))
                    bits_off = pos + 24 + nb
                    miss_off = bits_off + n * row_bytes if has_m else None
                    pos = bits_off + n * row_bytes * (2 if has_m else 1)
                    taken = 0
                k = min(size - filled, n - taken)
                read_into(bits[filled:filled + k], bits_off + taken * row_bytes)
                if miss_off is not None:
                    if miss is None:
                        miss = np.zeros((size, W), dtype=np.uint32)
                    elif not any_m:
                        miss[:filled] = 0
                    read_into(miss[filled:filled + k], miss_off + taken * row_bytes)
                    any_m = True
                elif any_m:
                    miss[filled:filled + k] = 0
                a0, a1 = int(cstart[taken]), int(cnul[taken + k - 1]) + 1
                pieces.append(craw[a0:a1])
                offs.append(cstart[taken:taken + k] - a0 + blob_len)
                blob_len += a1 - a0
                taken += k
                filled += k
            if filled == 0:
                if pool is not None:
                    pool.put(token)
                return
            m = miss[:filled] if any_m and miss[:filled].any() else None
            names = NameBlob(pieces[0] if len(pieces) == 1 else b''.join(pieces),
                             offs[0] if len(offs) == 1 else np.concatenate(offs))
            out = VariantBatch(names, bits[:filled], m)
            out.token = token
            if m is None and defer_empty:
                out.report_empty = True     # no missing genotypes: "no observation" is carriers == 0 in the results
            else:
                empty = ~(out.bits.any(axis=1) | (m.any(axis=1) if m is not None else False))
                for i in np.nonzero(empty)[0]:
                    sys.stderr.write('No observations of ' + out.names[i] + ' in selected samples\n')
            yield out

    sample_lists = VariantReader.sample_lists
    k_vector = VariantReader.k_vector


def open_variants(var_type, path, p, uncompressed=False, cache=None, threads=1):
    """VariantReader, or -- with ``cache`` -- a CachedVariantReader when a valid packed cache of
    ``path`` for this sample order exists, else a VariantReader that writes the cache as it reads
    (``--bits-cache``)."""
    if cache is None:
        return VariantReader(var_type, path, p, uncompressed, threads)
    samples = [str(s) for s in p.index]
    W = words_per_row(len(samples))
    if PackedCache.valid(cache, var_type, path, samples, W):
        sys.stderr.write('Reading packed variants from ' + str(cache) + '\n')
        return CachedVariantReader(cache, p, threads)
    rd = VariantReader(var_type, path, p, uncompressed, threads)
    writer = PackedCacheWriter(cache, var_type, path, samples, W)
    inner = rd.batches

    def batches(size, names_cap=None, pool=None):
        done = False
        try:
            for b in inner(size, names_cap, pool):
                writer.add(b)
                yield b
            done = True
        finally:
            writer.close(complete=done)
            if not done:
                try:
                    os.unlink(cache)
                except OSError:
                    pass

    rd.batches = batches
    # a run that tokenises the text on the device (text_batches) writes the cache itself, from the rows
    # it brings back (pyseer_b200/__main__.py)
    rd.cache_writer = writer
    rd.cache_path = cache
    return rd


class VcfReader(object):
    """VCF (and burden-region) input without pysam: plain or gzip text VCF, dominant encoding, parsed
    by the library's native reader (``psb_reader_*`` with var_type 2).

    Follows input.read_vcf_var (input.py:457-502): a sample carries the variant if any haplotype
    of its GT is a non-reference allele; '.' haplotypes mark the genotype missing unless a
    called haplotype follows; multi-allelic records and records whose FILTER is neither empty
    nor PASS are skipped (they still count as loaded, and end up AF-filtered like the
    reference's ``None`` sentinel).  With ``burden_file`` each row is the OR over every record
    overlapping the region(s) of one line ``name contig:start-end[,contig:start-end...]``
    (input.py:395-411, load_burden :250-266); the packed rows of the whole VCF are held in memory
    for that.  Produces the same VariantBatch objects as VariantReader."""

    def __init__(self, path, p, burden_file=None, reducer=None, threads=1):
        import ctypes
        from . import _lib
        self.samples = [str(s) for s in p.index]
        self.n_samples = len(self.samples)
        self.W = words_per_row(self.n_samples)
        self._lib = _lib.load()
        names = (ctypes.c_char_p * self.n_samples)(*[s.encode() for s in self.samples])
        self._h = ctypes.c_void_p()
        try:
            _lib.check(self._lib.psb_reader_open(str(path).encode(), 2, names, self.n_samples,
                                                 ctypes.byref(self._h)))
        except _lib.PsbError as e:
            if 'no #CHROM header' in str(e):
                raise ValueError('no #CHROM header line found; is this a VCF file?')
            raise
        if threads and threads > 1:
            _lib.check(self._lib.psb_reader_set_threads(self._h, int(threads)))
        self.regions = None
        # burden regions: `reducer(vbits, vmiss, offsets, members) -> (bits, missing)` forms the
        # per-region union of record rows on the device (Engine.submit_burden + download_rows, as
        # the CLI wires it); there is no host implementation in the package
        if burden_file and reducer is None:
            raise ValueError('burden regions need a `reducer` (the device union, '
                             'Engine.submit_burden); pyseer_b200 has no CPU fallback')
        self.reducer = reducer
        if burden_file:
            self.regions = []
            with open(burden_file) as rf:
                for line in rf:
                    name, spec = line.rstrip().split()
                    self.regions.append((name, spec.split(',')))
            # every record is parsed once into a packed row (read_vcf_var on an empty dictionary)
            contig, pos, reflen, skip, bits, miss = [], [], [], [], [], []
            for nm, b, m, info, cg, ps, rl in self._native(8192):
                contig.extend(cg)
                pos.append(ps)
                reflen.append(rl)
                skip.append(info & 12)
                bits.append(b)
                miss.append(m)
            cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dtype=dt)
            self.rec_contig = np.array(contig, dtype=object)
            self.rec_start = cat(pos, np.int64) - 1
            self.rec_end = self.rec_start + cat(reflen, np.int32)
            self.rec_skip = cat(skip, np.int32)
            self.rec_pos = cat(pos, np.int64)
            self.rec_bits = np.concatenate(bits) if bits else np.zeros((0, self.W), dtype=np.uint32)
            self.rec_miss = np.concatenate(miss) if bits else None
            if self.rec_miss is not None and not self.rec_miss.any():
                self.rec_miss = None

    def close(self):
        if self._h:
            self._lib.psb_reader_close(self._h)
            self._h = None

    def _native(self, size):
        """Batches of parsed records: names, carrier rows, missing rows, info flags (1 missing, 2 no
        observation, 4 multi-allelic, 8 filtered), contigs, 1-based positions, REF lengths."""
        import ctypes
        from . import _lib
        cap = max(1 << 20, 512 * size)
        while True:
            bits = np.empty((size, self.W), dtype=np.uint32)
            miss = np.empty((size, self.W), dtype=np.uint32)
            names = ctypes.create_string_buffer(cap)
            off = np.empty(size, dtype=np.int64)
            info = np.empty(size, dtype=np.int32)
            n = ctypes.c_int64(0)
            anym = ctypes.c_int32(0)
            _lib.check(self._lib.psb_reader_next(
                self._h, size, bits.ctypes.data, miss.ctypes.data, self.W, ctypes.addressof(names), cap,
                off.ctypes.data, info.ctypes.data, ctypes.byref(n), ctypes.byref(anym)))
            n = n.value
            if n == 0:
                return
            raw = names.raw
            nm = [raw[off[i]:raw.index(b'\0', off[i])].decode() for i in range(n)]
            cbuf = ctypes.create_string_buffer(cap)
            coff = np.empty(n, dtype=np.int64)
            pos = np.empty(n, dtype=np.int64)
            rl = np.empty(n, dtype=np.int32)
            _lib.check(self._lib.psb_reader_vcf_info(self._h, n, ctypes.addressof(cbuf), cap,
                                                     coff.ctypes.data, pos.ctypes.data, rl.ctypes.data))
            craw = cbuf.raw
            cg = [craw[coff[i]:craw.index(b'\0', coff[i])].decode() for i in range(n)]
            yield nm, bits[:n], miss[:n], info[:n], cg, pos, rl

    def _region_members(self, specs):
        """Record indices a burden line fetches, in fetch order (input.py:398-407), or None when
        a region does not parse (the reference then yields its None sentinel)."""
        import re
        members = []
        for spec in specs:
            mt = re.match(r'^(.+):(\d+)-(\d+)$', spec)
            if not mt:
                sys.stderr.write('Could not parse region %s\n' % str(spec))
                return None
            contig, lo, hi = mt.group(1), int(mt.group(2)) - 1, int(mt.group(3))
            hit = np.nonzero((self.rec_contig == contig) & (self.rec_start < hi) &
                             (self.rec_end > lo))[0]
            for j in hit.tolist():
                if self.rec_skip[j] & 4:
                    sys.stderr.write('Multiple alleles at %s_%s. Skipping\n' %
                                     (self.rec_contig[j], self.rec_pos[j]))
                elif not self.rec_skip[j]:
                    members.append(j)
        return members

    def _burden_batches(self, size):
        for b0 in range(0, len(self.regions), size):
            names, offsets, members = [], [0], []
            for rname, specs in self.regions[b0:b0 + size]:
                mem = self._region_members(specs)
                names.append('NA' if mem is None else rname)
                members.extend(mem or [])
                offsets.append(len(members))
            bits, miss = self.reducer(self.rec_bits, self.rec_miss, np.array(offsets, dtype=np.int64),
                                      np.array(members, dtype=np.int32))
            empty = ~(bits.any(axis=1) | (miss.any(axis=1) if miss is not None else False))
            for j in np.nonzero(empty)[0]:
                if names[j] != 'NA':
                    sys.stderr.write('No observations of ' + names[j] + ' in selected samples\n')
            if miss is not None and not miss.any():
                miss = None
            vb = VariantBatch(names, bits, miss)
            # a burden line whose region does not parse is the reference's None sentinel: never fitted
            vb.skipped = np.array([nm == 'NA' for nm in names], dtype=bool)
            yield vb

    def batches(self, size):
        if self.regions is not None:
            for batch in self._burden_batches(size):
                yield batch
            return
        for nm, bits, miss, info, cg, pos, rl in self._native(size):
            for i in range(len(nm)):
                if info[i] & 4:
                    sys.stderr.write('Multiple alleles at %s_%s. Skipping\n' % (cg[i], pos[i]))
                if info[i] & 12:
                    # skipped record: the reference yields its None sentinel (counted as loaded and
                    # pre-filtered); an empty row takes the same route through the AF filter
                    nm[i] = 'NA'
                elif info[i] & 2:
                    sys.stderr.write('No observations of ' + nm[i] + ' in selected samples\n')
            vb = VariantBatch(nm, np.ascontiguousarray(bits),
                              np.ascontiguousarray(miss) if (info & 1).any() else None)
            # skipped records never reach a model in the reference (k is None: input.py:603-611), whatever
            # --min-af says: the result loop forces them to 'af-filter' from this mask
            vb.skipped = (info & 12) != 0
            yield vb

    sample_lists = VariantReader.sample_lists
    k_vector = VariantReader.k_vector


def load_lineage(infile, p):
    """input.py:139-177: `sample<TAB>cluster` file -> (binary design matrix with one column
    per cluster in sorted label order, list of labels); rows follow the phenotype order."""
    rows = [x.rstrip().split() for x in open(infile) if x.strip()]
    lin = pd.Series([r[1] for r in rows], index=[str(r[0]) for r in rows])
    if np.any(lin.index.duplicated()):
        lin = lin.loc[~lin.index.duplicated()]
    if len(p.index.difference(lin.index)) > 0:
        sys.stderr.write('All samples with a phenotype must be present in lineage file\n')
        sys.exit(1)
    lin = lin.loc[p.index.intersection(lin.index)]
    labels = sorted(set(lin.values))
    design = np.array([[1 if x == lab else 0 for x in lin.values] for lab in labels]).T
    assert np.all(lin.index == p.index)
    return design, labels
