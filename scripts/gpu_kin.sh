#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kinship_gpu.py -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/pytest_kin.log
cat gpurun_out/pytest_kin.log | cut -c1-600
timeout 300 python - <<'PY' 2>&1 | tail -3
import time, numpy as np
from pyseer_b200.engine import Engine, synth_host
n, nv = 5000, 200000
bits = synth_host(3, 0, nv, n)
eng = Engine(0); eng.kinship_begin(n)
t=time.time(); eng.kinship_add(bits); K=eng.kinship_fetch(); dt=time.time()-t
print('kinship N=%d V=%d: %.3f s (incl. H2D + fetch), trace %.0f' % (n, nv, dt, np.trace(K)))
PY
