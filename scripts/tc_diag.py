"""Diagnostic: tensor-core LMM contraction vs the FP64 kernel on the same inputs."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyseer_b200 import lmm as plmm
from pyseer_b200.engine import synth_host, unpack_rows


def problem(n, seed=1):
    rng = np.random.RandomState(seed)
    G = (rng.uniform(size=(n, 2 * n)) < rng.uniform(0.05, 0.95, 2 * n)).astype(float)
    K = G.dot(G.T)
    g = G.dot(rng.normal(size=2 * n))
    y = (g - g.mean()) / g.std() * np.sqrt(0.5) + np.sqrt(0.5) * rng.normal(size=n)
    return K * (n / np.diag(K).sum()), y


for n, nv in ((int(a), int(b)) for a, b in (t.split(':') for t in sys.argv[1].split(','))):
    K, y = problem(n)
    bits = synth_host(7, 0, nv, n)
    snps = unpack_rows(bits, n).T.astype(float)
    m0 = plmm.KinshipLMM(np.ones((n, 1)), y.reshape(-1, 1), K.copy(), precision=0)
    h2 = m0.findH2()['h2']
    r0 = plmm.fit_lmm_block(m0, h2, snps)
    for prec in [int(p) for p in sys.argv[2].split(',')]:
        m = plmm.KinshipLMM(np.ones((n, 1)), y.reshape(-1, 1), None, precision=prec)
        m.U, m.S = m0.U, m0.S
        t0 = time.time()
        try:
            r = plmm.fit_lmm_block(m, h2, snps)
        except Exception as e:
            print('n=%d prec=%d FAILED: %s' % (n, prec, e))
            break
        ratio = r0['beta'] / r['beta']          # = a_tc / a_fp64
        perr = np.abs(r['p_values'] / r0['p_values'] - 1)
        print('n=%d nv=%d prec=%d  a ratio: min %.12g max %.12g  max|ratio-1| %.3e  p rel err max %.3e  (%.2fs)'
              % (n, nv, prec, np.nanmin(ratio), np.nanmax(ratio), np.nanmax(np.abs(ratio - 1)),
                 np.nanmax(perr), time.time() - t0))
        if np.nanmax(np.abs(ratio - 1)) > 1e-6:
            print('   first ratios:', ratio[:8])
            print('   ratios @128..136:', ratio[128:136] if nv > 136 else None)
        m.close()
    m0.close()
