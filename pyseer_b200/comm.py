"""Multi-GPU plumbing over the library's communicator (``psb_comm_*``, csrc/psb_comm.cu).

K-mers shard by contiguous ranges over the GPUs; the one exchange of the path is the NCCL gather of
the per-variant result table on a root rank, in rank order = input order (the reference keeps input
order through its ordered ``pool.starmap``, __main__.py:541-546, :777-780).  No PyTorch: NCCL is
loaded by the library itself.

Two ways to make a communicator:

``Comm.from_env(engine)``   one process per GPU, launched torchrun-style (``RANK``, ``WORLD_SIZE``,
                            ``LOCAL_RANK``, ``MASTER_PORT`` in the environment).  Rank 0 creates the
                            NCCL id and publishes it through a file (one node: the ranks share /tmp);
                            the name carries the launcher's pid, so consecutive launches on the same
                            port cannot pick up each other's id.
``Comm.local(engines)``     one process driving several GPUs (the CLI's ``--gpus N``).
"""
import ctypes
import os
import tempfile
import time
from ctypes import byref, c_int32, c_int64, c_void_p

import numpy as np

from . import _lib
from ._lib import PsbResults, check
from .engine import Results

ID_BYTES = 128


def shard_range(n_variants, rank, world):
    """Contiguous [first, last) range of variant ids of ``rank``; sizes differ by at most 1."""
    base, rem = divmod(int(n_variants), int(world))
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def _id_path():
    explicit = os.environ.get('PSB_COMM_ID_FILE')
    if explicit:
        return explicit
    tag = '%s_%s_%d' % (os.environ.get('MASTER_PORT', '0'),
                        os.environ.get('TORCHELASTIC_RUN_ID', 'none'), os.getppid())
    return os.path.join(tempfile.gettempdir(), 'psb_nccl_%s.id' % tag)


class Comm(object):
    def __init__(self, handle, engines, world, ranks):
        self.lib = _lib.load()
        self._h = handle
        self.engines = engines
        self.world = world
        self.ranks = ranks              # global ranks of the local engines
        self.rank = ranks[0]

    # -- construction ------------------------------------------------------------------------
    @classmethod
    def from_env(cls, engine, timeout=300.0):
        lib = _lib.load()
        rank = int(os.environ.get('RANK', '0'))
        world = int(os.environ.get('WORLD_SIZE', '1'))
        buf = (ctypes.c_uint8 * ID_BYTES)()
        path = _id_path()
        if rank == 0:
            check(lib.psb_comm_unique_id(buf))
            tmp = '%s.%d.tmp' % (path, os.getpid())
            with open(tmp, 'wb') as f:
                f.write(bytes(buf))
            os.replace(tmp, path)
        else:
            t0 = time.time()
            while True:
                try:
                    with open(path, 'rb') as f:
                        data = f.read()
                    if len(data) == ID_BYTES:
                        break
                except OSError:
                    pass
                if time.time() - t0 > timeout:
                    raise RuntimeError('no NCCL id from rank 0 at %s after %.0f s' % (path, timeout))
                time.sleep(0.02)
            ctypes.memmove(buf, data, ID_BYTES)
        h = c_void_p()
        check(lib.psb_comm_init_rank(engine._ctx, world, rank, buf, byref(h)))
        comm = cls(h, [engine], world, [rank])
        comm.barrier()                  # every rank has read the id: rank 0 may remove the file
        if rank == 0:
            try:
                os.unlink(path)
            except OSError:
                pass
        return comm

    @classmethod
    def local(cls, engines):
        lib = _lib.load()
        arr = (c_void_p * len(engines))(*[e._ctx for e in engines])
        h = c_void_p()
        check(lib.psb_comm_init_all(arr, len(engines), byref(h)))
        return cls(h, list(engines), len(engines), list(range(len(engines))))

    def close(self):
        if self._h:
            self.lib.psb_comm_destroy(self._h)
            self._h = c_void_p()

    def info(self):
        w, n, v = c_int32(0), c_int32(0), c_int32(0)
        check(self.lib.psb_comm_info(self._h, byref(w), byref(n), byref(v)))
        return {'world': w.value, 'n_local': n.value, 'nccl_version': v.value}

    # -- small collectives on host arrays (one engine per process) -----------------------------
    def bcast(self, arr, root=0):
        """In place on a C-contiguous NumPy array."""
        assert arr.flags['C_CONTIGUOUS']
        check(self.lib.psb_comm_bcast(self._h, arr.ctypes.data_as(c_void_p), arr.nbytes, root))
        return arr

    def allreduce(self, values, op='sum'):
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        check(self.lib.psb_comm_allreduce(self._h, v.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                          v.shape[0], {'sum': 0, 'max': 1}[op]))
        return v

    def barrier(self):
        check(self.lib.psb_comm_barrier(self._h))

    # -- the gather of the result table -----------------------------------------------------------
    def gather_begin(self, rows_max, root=0):
        check(self.lib.psb_comm_gather_begin(self._h, int(root), int(rows_max)))

    def gather_wait(self):
        check(self.lib.psb_comm_gather_wait(self._h))

    def gather_bytes(self):
        b = c_int64(0)
        check(self.lib.psb_comm_gather_bytes(self._h, byref(b)))
        return b.value

    def gather_fetch(self, src_rank, n_betas=0, pointers=None):
        """Rank ``src_rank``'s table from the root's receive buffer: a :class:`Results` of NumPy
        columns, or -- with ``pointers`` (column name -> raw address) -- copied there; returns
        (results_or_None, n_rows, counts)."""
        n = c_int64(0)
        cnt = (c_int64 * 4)()
        check(self.lib.psb_comm_gather_fetch(self._h, int(src_rank), None, byref(n), cnt))
        counts = {'loaded': cnt[0], 'prefiltered': cnt[1], 'tested': cnt[2], 'passed': cnt[3]}
        S = n.value
        out = PsbResults()
        r = None
        if pointers is not None:
            for f, ptr in pointers.items():
                setattr(out, f, c_void_p(int(ptr)))
        else:
            r = Results()
            r.carriers = np.empty(S, dtype=np.int32)
            r.missing = np.empty(S, dtype=np.int32)
            r.flags = np.empty(S, dtype=np.uint32)
            for f in ('af', 'prep', 'pvalue', 'beta', 'bse', 'extra'):
                setattr(r, f, np.empty(S, dtype=np.float64))
            r.betas = np.empty((S, n_betas), dtype=np.float64)
            for f in ('carriers', 'missing', 'af', 'prep', 'pvalue', 'beta', 'bse', 'extra', 'flags'):
                setattr(out, f, getattr(r, f).ctypes.data_as(c_void_p))
            if n_betas:
                out.betas = r.betas.ctypes.data_as(c_void_p)
            r.counts = counts
            r.lineage = None
        check(self.lib.psb_comm_gather_fetch(self._h, int(src_rank), byref(out), byref(n), cnt))
        return r, S, counts
