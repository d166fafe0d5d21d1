"""Time psb_kinship_add (K = G G', pyseer/similarity.py:99-116) through both contraction paths:
tcgen05 int8 (default) and AND + POPCOUNT on the CUDA cores (PSB_KIN_TC=0).  Host rows in, the
call returns when the device is done: the number includes the host->device copy of the rows.
Usage: python scripts/kinship_throughput.py [N] [variants]  -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyseer_b200.engine import Engine, synth_host  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    nv = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    bits = synth_host(3, 0, nv, n, af_lo=0.02, af_hi=0.98)
    res = {'n_samples': n, 'variants': nv, 'row_bytes': int(bits.nbytes)}
    tables = {}
    for mode, name in (('1', 'tensor'), ('0', 'popcount')):
        os.environ['PSB_KIN_TC'] = mode
        eng = Engine(0)
        eng.kinship_begin(n)
        eng.kinship_add(bits, None, 0.01, 0.99, 0.05)          # allocations, first launches
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            eng.kinship_add(bits, None, 0.01, 0.99, 0.05)
            best = min(best, time.perf_counter() - t0)
        tables[name] = eng.kinship_fetch()
        eng.close()
        res[name + '_ms'] = round(best * 1e3, 2)
        res[name + '_int8_tops' if mode == '1' else name + '_pair_gops'] = round(
            (2.0 if mode == '1' else 1.0) * nv * n * (n + 1) / 2 / best / (1e12 if mode == '1' else 1e9), 1)
    res['equal'] = bool(np.array_equal(tables['tensor'], tables['popcount']))
    print(json.dumps(res))


if __name__ == '__main__':
    main()
