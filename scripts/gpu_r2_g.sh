#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for env in "PSB_LOGIT_FAST=0" "PSB_LOGIT_FAST=0 PSB_LOGIT_WARM=0" "PSB_LOGIT_FAST=1"; do
  echo "=== $env"; env $env timeout 300 python scripts/dbg_v60949.py 60950 2>&1 | tail -4
done
timeout 900 python -m pytest tests/test_fixed_gpu.py tests/test_cli_gpu.py tests/test_baseline_sizes_gpu.py::test_config2_fixed_n2000_logit_firth -q --durations=5 > gpurun_out/r2g_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2g_tests.log; tail -30 gpurun_out/r2g_tests.log
for minb in 2; do
  echo "=== fixed bench MINB=$minb"
  PSB_LOGIT_MINB=$minb timeout 600 python bench.py --model fixed --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_fixed_minb$minb.json 2> gpurun_out/r2g_bench_fixed_minb$minb.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2g_bench_fixed_minb$minb.json'))
    print({k:d[k] for k in ('value','ms_per_step','stats','check')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2g_bench_fixed_minb$minb.err').read()[-1500:])
PY
done
timeout 900 python scripts/cli_throughput.py --kmers 40000 > gpurun_out/r2g_cli_throughput.json 2> gpurun_out/r2g_cli_throughput.err; echo "cli exit $?"; cat gpurun_out/r2g_cli_throughput.json; tail -3 gpurun_out/r2g_cli_throughput.err
