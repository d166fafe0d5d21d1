"""The CLI's streaming pipeline: input.py's k-mer streaming (pyseer/input.py:505-707) as a
pinned-host -> device staging pipeline, and the worker map (pyseer/__main__.py:517-593, 762-827) as
batches in flight on one or several GPUs.

Three stages run concurrently, each in its own thread (the native parser, the CUDA calls and the
native formatter all release the GIL):

  reader     parses batch k+1 of the variant file straight into a page-locked buffer of a small pool
             (``PinnedPool``), so that ``psb_submit`` is a true asynchronous DMA; plain k-mer files go
             one step further (``TextPool``): the reader only reads / inflates the TEXT into the
             page-locked buffer and cuts it into lines, the device tokenises it (``psb_submit_text``);
  GPU        ``BatchRunner``: submit(k+1) on the copy stream and the table of batch k on the fetch
             stream (``psb_fetch_begin``) while the kernels of batch k+1 run; with several GPUs a *super-step* deals one batch to every GPU (contiguous,
             in input order), the runs are queued from one thread per GPU, and the result tables come
             back through the library's NCCL gather on GPU 0 (``comm.gather_begin`` / ``gather_fetch``)
             in rank order = input order, as the reference's ordered ``pool.starmap`` keeps it;
  output     formats batch k-1 (``psb_format_rows``) and writes it, in input order.
"""
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .engine import PinnedBuffer


class PinnedPool(object):
    """``n`` page-locked (bits, missing) buffer pairs of ``rows`` x ``W`` uint32."""

    def __init__(self, n, rows, W, with_missing):
        self._bufs = []
        self._free = queue.Queue()
        for i in range(n):
            b = PinnedBuffer((rows, W), np.uint32)
            m = PinnedBuffer((rows, W), np.uint32) if with_missing else None
            self._bufs.append((b, m))
            self._free.put(i)

    def get(self):
        i = self._free.get()
        b, m = self._bufs[i]
        return b.array, (m.array if m is not None else None), i

    def put(self, token):
        if token is not None:
            self._free.put(token)

    def close(self):
        for b, m in self._bufs:
            b.free()
            if m is not None:
                m.free()
        self._bufs = []


class TextPool(object):
    """``n`` page-locked (text, line_start, line_len) buffer sets for ``VariantReader.text_batches``."""

    def __init__(self, n, rows, text_bytes):
        self._bufs = []
        self._free = queue.Queue()
        for i in range(n):
            self._bufs.append((PinnedBuffer((text_bytes,), np.uint8), PinnedBuffer((rows,), np.int64),
                               PinnedBuffer((rows,), np.int32)))
            self._free.put(i)

    def get(self):
        i = self._free.get()
        t, a, b = self._bufs[i]
        return t.array, a.array, b.array, i

    def put(self, token):
        if token is not None:
            self._free.put(token)

    def close(self):
        for bufs in self._bufs:
            for b in bufs:
                b.free()
        self._bufs = []


class AsyncFetch(object):
    """One page-locked set of result columns for ``Engine.fetch_begin`` / ``fetch_wait``; ``wait``
    returns a ``Results`` holding copies, so the set is free for the next ``begin``."""
    COLS = (('carriers', np.int32), ('missing', np.int32), ('af', np.float64), ('prep', np.float64),
            ('pvalue', np.float64), ('beta', np.float64), ('bse', np.float64), ('extra', np.float64),
            ('flags', np.uint32))

    def __init__(self, eng, rows, n_betas):
        self.eng, self.nb, self.rows = eng, int(n_betas), int(rows)
        self.bufs = {name: PinnedBuffer((rows,), dt) for name, dt in self.COLS}
        self.betas = PinnedBuffer((rows * max(self.nb, 1),), np.float64)
        self.n = 0

    def begin(self, n):
        if n > self.rows:                      # a batch larger than announced: larger buffers
            self.close()
            self.rows = int(n)
            self.bufs = {name: PinnedBuffer((self.rows,), dt) for name, dt in self.COLS}
            self.betas = PinnedBuffer((self.rows * max(self.nb, 1),), np.float64)
        ptrs = {name: self.bufs[name].array.ctypes.data for name, _ in self.COLS}
        if self.nb:
            ptrs['betas'] = self.betas.array.ctypes.data
        self.n = n
        self.eng.fetch_begin(ptrs)

    def wait(self):
        from .engine import Results
        counts = self.eng.fetch_wait()
        r = Results()
        for name, _ in self.COLS:
            setattr(r, name, self.bufs[name].array[:self.n].copy())
        r.betas = self.betas.array[:self.n * self.nb].reshape(self.n, self.nb).copy()
        r.counts = counts
        r.lineage = None
        return r

    def close(self):
        for b in self.bufs.values():
            b.free()
        self.betas.free()
        self.bufs = {}


def submit_batch(eng, b):
    """Rows of a batch to the engine: packed rows from the host, or k-mer text for the device parser."""
    if b.text is not None:
        text, n_bytes, lstart, llen = b.text
        eng.submit_text(text, n_bytes, lstart, llen, b.n)
    else:
        eng.submit(b.bits, b.missing)


class _Stop(object):
    pass


class Prefetch(object):
    """Runs an iterator in a thread, ``depth`` items ahead."""

    def __init__(self, it, depth=2):
        self._q = queue.Queue(maxsize=max(1, depth))
        self._err = None
        self._cancel = False
        self._t = threading.Thread(target=self._run, args=(it,), daemon=True)
        self._t.start()

    def _run(self, it):
        try:
            for item in it:
                if self._cancel:
                    break
                self._q.put(item)
        except BaseException as e:          # noqa: BLE001 -- re-raised in the consumer
            self._err = e
        self._q.put(_Stop)

    def __iter__(self):
        while True:
            item = self._q.get()
            if item is _Stop:
                if self._err is not None:
                    raise self._err
                return
            yield item

    def cancel(self):
        self._cancel = True
        try:
            while True:
                self._q.get_nowait()
        except queue.Empty:
            pass


class BatchRunner(object):
    """Drives batches through one or several engines and yields ``(batch, results)`` in input order.

    ``run(engine)`` queues ``psb_run_lmm`` / ``psb_run_fixed`` for the rows last submitted to that
    engine; ``lineage``: None, or the ``lmm_rule`` flag of ``Engine.run_lineage`` (fixed effects:
    False) whose result is attached to ``results.lineage``."""

    def __init__(self, engines, run, n_betas=0, lineage=None, comm=None, rows_max=0, digests=False, rows=False):
        self.engines = list(engines)
        self._run = run
        self.digests = digests          # --output-patterns on batches parsed on the device
        self.rows = rows                # --bits-cache being written from batches parsed on the device
        self.n_betas = n_betas
        self.lineage = lineage
        self.comm = comm if len(self.engines) > 1 else None
        self.rows_max = rows_max
        self._pool = ThreadPoolExecutor(len(self.engines)) if len(self.engines) > 1 else None

    def close(self):
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None

    def run(self, eng, b=None):
        """Queues the model run for the rows last submitted to ``eng``; a batch whose rows only exist on
        the device gets the MD5 digests of its patterns (``psb_pattern_digests``) while they are there."""
        self._run(eng)
        if self.digests and b is not None and b.bits is None:
            b.digests = eng.pattern_digests()
        if self.rows and b is not None and b.bits is None:
            b.bits, b.missing = eng.download_rows()

    def _fetch(self, eng, b):
        r = eng.fetch()
        if self.lineage is not None:
            r.lineage = eng.run_lineage(self.lineage)
        return r

    def results(self, batches):
        if len(self.engines) == 1:
            for item in self._single(batches):
                yield item
        else:
            for item in self._multi(batches):
                yield item

    def _single(self, batches):
        eng = self.engines[0]
        if self.lineage is not None or self.rows_max <= 0:
            # lineage effects are fitted on the device table of the run they belong to: one run at a time
            prev = None
            for b in batches:
                submit_batch(eng, b)                   # H2D of batch k+1 on the copy stream ...
                if prev is not None:
                    yield prev, self._fetch(eng, prev)  # ... while batch k finishes and comes back
                self.run(eng, b)
                prev = b
            if prev is not None:
                yield prev, self._fetch(eng, prev)
            return
        # Three things in flight: the copy of batch k+1 (copy stream), the table of batch k on its way
        # to the host (psb_fetch_begin, fetch stream) and the kernels of batch k+1, which write the
        # other set of result columns -- the device never waits for the host.
        fetcher = AsyncFetch(eng, self.rows_max, self.n_betas)
        try:
            prev = inflight = None
            for b in batches:
                submit_batch(eng, b)
                if prev is not None:
                    if inflight is not None:
                        yield inflight, fetcher.wait()
                    fetcher.begin(prev.n)
                    inflight = prev
                self.run(eng, b)
                prev = b
            if inflight is not None:
                yield inflight, fetcher.wait()
            if prev is not None:
                fetcher.begin(prev.n)
                yield prev, fetcher.wait()
        finally:
            fetcher.close()

    def _multi(self, batches):
        n = len(self.engines)
        group, prev, prev_gathered = [], None, False

        def retire(grp, gathered):
            if gathered:
                self.comm.gather_wait()
                for g, b in enumerate(grp):
                    r, _, _ = self.comm.gather_fetch(g, n_betas=self.n_betas)
                    yield b, r
            else:
                for g, b in enumerate(grp):
                    yield b, self._fetch(self.engines[g], b)

        def launch(grp):
            for g, b in enumerate(grp):
                submit_batch(self.engines[g], b)

        def run_group(grp):
            list(self._pool.map(self.run, self.engines[:len(grp)], grp))
            # the NCCL gather needs every rank: a short last super-step (fewer batches than GPUs)
            # and runs with lineage effects are fetched from their own GPUs instead
            use_gather = self.comm is not None and len(grp) == n and self.lineage is None
            if use_gather:
                self.comm.gather_begin(max(self.rows_max, max(b.n for b in grp)), root=0)
            return use_gather

        for b in batches:
            group.append(b)
            if len(group) < n:
                continue
            launch(group)
            if prev is not None:
                for item in retire(prev, prev_gathered):
                    yield item
            prev_gathered = run_group(group)
            prev, group = group, []
        if group:
            launch(group)
            if prev is not None:
                for item in retire(prev, prev_gathered):
                    yield item
            prev_gathered = run_group(group)
            prev = group
        if prev is not None:
            for item in retire(prev, prev_gathered):
                yield item
