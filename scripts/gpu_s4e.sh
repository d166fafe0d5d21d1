#!/bin/bash
# session 4, call E: device eigh (cuSOLVER) through the LMM tests, setup timing, ncu of the pair-mode kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lmm_gpu.py tests/test_properties_gpu.py tests/test_burden_gpu.py tests/test_cli_gpu.py -m gpu -q --tb=short 2>&1 | tail -15 | cut -c1-400
timeout 600 python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
from pyseer_b200.engine import Engine
eng = Engine(0)
for n in (1000, 5000, 10000):
    rng = np.random.RandomState(n)
    G = (rng.uniform(size=(n, 2 * n)) < 0.3).astype(np.float32)
    K = (G @ G.T).astype(np.float64); K *= n / np.trace(K); K.flat[::n + 1] += 1.0
    t0 = time.time(); w, V = eng.eigh(K); t1 = time.time()
    if n <= 5000:
        w2, V2 = np.linalg.eigh(K); t2 = time.time()
    else:
        w2, t2 = w, t1
    res = np.abs(K @ V[:, -3:] - V[:, -3:] * w[-3:]).max()
    print('n=%d device eigh %.2fs numpy %.2fs  max|w-w2|/|w| %.2e  residual %.2e  orth %.2e' % (
        n, t1 - t0, t2 - t1, np.max(np.abs(w - w2) / np.abs(w2)), res,
        np.abs(V[:, :50].T @ V[:, :50] - np.eye(50)).max()), flush=True)
PY
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_lmm_quadform_tc -s 1 -c 1 -o gpurun_out/prof_tc_pair python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc_pair.log 2>&1
tail -2 gpurun_out/ncu_tc_pair.log | cut -c1-200
