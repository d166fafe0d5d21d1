#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for env in "PSB_LOGIT_FAST=0" "PSB_LOGIT_FAST=0 PSB_LOGIT_WARM=0" "PSB_LOGIT_FAST=1"; do
  echo "=== $env"; env $env timeout 300 python scripts/dbg_v60949.py 60949 2>&1 | tail -4
done
timeout 900 python -m pytest tests/test_fixed_gpu.py tests/test_cli_gpu.py tests/test_baseline_sizes_gpu.py::test_config2_fixed_n2000_logit_firth -x -q --durations=5 > gpurun_out/r2e_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2e_tests.log; tail -30 gpurun_out/r2e_tests.log
for minb in 2 1 3; do
  echo "=== fixed bench MINB=$minb"
  PSB_LOGIT_MINB=$minb timeout 600 python bench.py --model fixed --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_fixed_minb$minb.json 2> gpurun_out/r2e_bench_fixed_minb$minb.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2e_bench_fixed_minb$minb.json'))
print({k:d[k] for k in ('value','ms_per_step','stats','check')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'])
PY
done
echo "=== old kernel"; PSB_LOGIT_FAST=0 timeout 600 python bench.py --model fixed --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['stats'])"
