import sys, os, collections
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_fixed_gpu import _problem, _variants, NONE
from pyseer_b200 import model as pm, _lib
from pyseer_b200.engine import notes_from_flags
from oracle import fixed_oracle as fo
for n, dims, nv in ((100, 0, 120), (333, 3, 200)):
    m, y, rng = _problem(n, dims, 7 + n, True)
    bits, x = _variants(n, nv, y, rng, True)
    mm = m if dims else NONE
    onull = fo.fit_null(y, mm, NONE, False); of = fo.fit_null(y, mm, NONE, False, True)
    model = pm.FixedModel(y, mm, NONE, False, onull.llf, float(of))
    r = pm.run_fixed_bits(model, bits, None, 0.6, 0.5, 0.02, 0.98, 0.05)
    h = collections.Counter()
    for s in range(nv):
        h[(tuple(sorted(notes_from_flags(int(r.flags[s])))), bool(r.flags[s] & _lib.F_FIRTH_USED))] += 1
    for k, v in h.items(): print(n, dims, k, v)
    print('flags of rows 1..5', [hex(int(f)) for f in r.flags[:6]], r.pvalue[:6], r.beta[:6])
    model.close()
