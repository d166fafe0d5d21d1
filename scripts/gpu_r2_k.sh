#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 1500 python -m pytest tests -q -m gpu -x --durations=6 > gpurun_out/r2k_suite.log 2>&1
echo "exit $?" >> gpurun_out/r2k_suite.log; tail -14 gpurun_out/r2k_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for minb in 2; do
  echo "=== fixed bench MINB=$minb"
  PSB_LOGIT_MINB=$minb timeout 600 python bench.py --model fixed --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2k_bench_fixed_minb$minb.json 2> gpurun_out/r2k_bench_fixed_minb$minb.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2k_bench_fixed_minb$minb.json'))
    print({k:d[k] for k in ('value','ms_per_step','stats')}, d['roofline']['frac'], d['roofline']['kernel_ms'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2k_bench_fixed_minb$minb.err').read()[-1500:])
PY
done
