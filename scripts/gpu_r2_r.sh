#!/bin/bash
# same-box A/B of the skipped zero box; device h2 sums; 2 k=4 runs each
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_eigh_gpu.py -q -x 2>&1 | tail -3
for rep in 1 2; do for skip in 1 0; do
PSB_TC_SKIP=$skip timeout 600 python bench.py --precision 46 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r2r_bench_skip$skip.json 2> gpurun_out/r2r_bench_skip$skip.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2r_bench_skip$skip.json'))
    r=d['roofline']
    print('skip $skip', {k:d[k] for k in ('value','ms_per_step')}, r['frac'], r['kernel_ms'], r['side_kernels_ms'], d['clocks']['sm_mhz'], r['traffic'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2r_bench_skip$skip.err').read()[-1500:])
PY
done; done
