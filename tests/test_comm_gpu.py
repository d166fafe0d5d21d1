"""The library-owned NCCL gather of the result table (psb_comm_*, csrc/psb_comm.cu): the table
gathered from N contexts equals the table of one context over the whole batch, byte for byte and in
input order.  The N = 2 case needs two GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

COLS = ('carriers', 'missing', 'af', 'prep', 'pvalue', 'beta', 'bse', 'extra', 'flags')


def _lmm_state(n):
    import benchdata
    from pyseer_b200.lmm import KinshipLMM
    X, y, K = benchdata.lmm_problem(n, seed=3)
    m = KinshipLMM(X, y.reshape(-1, 1), K)
    h2 = float(m.findH2()['h2'])
    S, U = m.getSU()
    m.close()
    return X, y, U, S, h2


def _gathered_equals_single(n_dev, model):
    from pyseer_b200.comm import Comm, shard_range
    from pyseer_b200.engine import Engine, device_count, synth_host
    if device_count() < n_dev:
        pytest.skip('needs %d GPUs' % n_dev)
    n, nv = 700, 9001                                  # odd count: shards differ by one row
    ys = None
    if model == 'lmm':
        X, y, U, S, h2 = _lmm_state(n)
        setup = lambda e: e.lmm_setup(X, y, U, S, h2, 5)                      # noqa: E731
        run = lambda e: e.run_lmm(0.02, 0.98, 0.05, 0.9, 0.8, True)           # noqa: E731
        nb = 0
    else:
        import benchdata
        from pyseer_b200 import model as pm
        mds, y = benchdata.fixed_problem(n, 4)
        none = np.empty((0, 0))
        null = pm.fit_null(y, mds, none, False)
        firth = pm.fit_null(y, mds, none, False, True)
        Z = np.c_[np.ones(n), mds]
        setup = lambda e: e.fixed_setup(Z, y, False, null.llf, firth)         # noqa: E731
        run = lambda e: e.run_fixed(0.02, 0.98, 0.05, 0.9, 0.8, False)        # noqa: E731
        nb = 4
    ys = np.where(y > np.median(y), 1, -1).astype(np.int8)
    bits = synth_host(11, 0, nv, n, 0.0, 1.0, 50, ys, 20)
    single = Engine(0)
    setup(single)
    single.submit(bits)
    run(single)
    ref = single.fetch()
    engines = [Engine(d) for d in range(n_dev)]
    comm = Comm.local(engines)
    assert comm.info()['world'] == n_dev and comm.info()['nccl_version'] > 20000
    rows_max = 0
    for r, e in enumerate(engines):
        setup(e)
        lo, hi = shard_range(nv, r, n_dev)
        rows_max = max(rows_max, hi - lo)
    for rep in range(2):                                # twice: buffers are reused, streams re-joined
        for r, e in enumerate(engines):
            lo, hi = shard_range(nv, r, n_dev)
            e.submit(bits[lo:hi])
            run(e)
        comm.gather_begin(rows_max, root=0)
        comm.gather_wait()
        got = {c: [] for c in COLS + ('betas',)}
        tested = 0
        for r in range(n_dev):
            t, nr, cnt = comm.gather_fetch(r, n_betas=nb)
            lo, hi = shard_range(nv, r, n_dev)
            assert nr == hi - lo and cnt['loaded'] == nr
            tested += cnt['tested']
            for c in got:
                got[c].append(getattr(t, c))
        assert tested == ref.counts['tested']
        for c in COLS:
            a = np.concatenate(got[c])
            assert a.tobytes() == getattr(ref, c).tobytes(), c
        if nb:
            assert np.concatenate(got['betas']).tobytes() == ref.betas.tobytes()
    comm.close()
    for e in engines + [single]:
        e.close()


@pytest.mark.parametrize('model', ['lmm', 'fixed'])
def test_gather_one_gpu(model):
    """world = 1: the pack / NCCL send-to-self / fetch path on a one-GPU box."""
    _gathered_equals_single(1, model)


@pytest.mark.parametrize('model', ['lmm', 'fixed'])
def test_gather_two_gpus_equals_one(model):
    _gathered_equals_single(2, model)
