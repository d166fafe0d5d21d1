"""Debug: GPU vs oracle coefficients of one variant of the configs[2] parity test."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from oracle import cpu_arm, synth, fixed_oracle as fo
import test_baseline_sizes_gpu as T
from pyseer_b200 import model as pm
n, dims = 2000, 10
mds, y = cpu_arm.fixed_problem(n, dims)
tasks = T._tasks(102000, 10000000)
ys = np.where(y > 0.5, 1, -1).astype(np.int8)
none = np.empty((0, 0))
onull = fo.fit_null(y, mds, none, False)
ofirth = fo.fit_null(y, mds, none, False, True)
idx = int(sys.argv[1]) if len(sys.argv) > 1 else 60949
c, o = divmod(idx, 3000)
first = int(tasks[c][0])
model = pm.FixedModel(y, mds, none, False, onull.llf, float(ofirth))
eng = model.engine
eng.synth_device(T.SEED, first, 3000, 0.02, 0.98, 1000, ys, 100)
eng.run_fixed(min_af=0.01, max_af=0.99, max_missing=0.05, filter_pvalue=1.0, lrt_pvalue=1.0, continuous=False)
r = eng.fetch()
bits = synth.synth_rows(T.SEED, first + o, 1, n, 0.02, 0.98, 1000, ys, 100)
x = synth.unpack_rows(bits, n)[0].astype(float)
ref = fo.fixed_effects_regression('v', y, x, mds, none, x.mean(), 'p', False, None, 1.0, 1.0, onull.llf, ofirth, [], [], False)
g = np.r_[r.extra[o], r.beta[o], r.betas[o]]
rr = np.r_[ref.intercept, ref.kbeta, ref.betas]
print('env', {k: v for k, v in os.environ.items() if k.startswith('PSB_')}, 'stats', eng.last_stats())
print('flags', hex(int(r.flags[o])), 'abs diff', np.abs(g - rr), 'p', r.pvalue[o], ref.pvalue, 'bse', r.bse[o], ref.bse)
# whole chunk: worst abs diff of beta
worst = 0
for s in range(0, 3000, 37):
    b1 = synth.synth_rows(T.SEED, first + s, 1, n, 0.02, 0.98, 1000, ys, 100)
    x1 = synth.unpack_rows(b1, n)[0].astype(float)
    rf = fo.fixed_effects_regression('v', y, x1, mds, none, x1.mean(), 'p', False, None, 1.0, 1.0, onull.llf, ofirth, [], [], False)
    if rf.prefilter or 'firth-fail' in rf.notes:
        continue
    worst = max(worst, np.abs(np.r_[r.extra[s], r.beta[s], r.betas[s]] - np.r_[rf.intercept, rf.kbeta, rf.betas]).max())
print('worst abs coefficient diff over 81 variants of the chunk', worst)
