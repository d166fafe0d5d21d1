#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_fixed_gpu.py tests/test_baseline_sizes_gpu.py::test_config2_fixed_n2000_logit_firth -q -s --durations=5 > gpurun_out/r2h_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2h_tests.log; tail -12 gpurun_out/r2h_tests.log
for minb in 2 3; do
  echo "=== fixed bench MINB=$minb"
  PSB_LOGIT_MINB=$minb timeout 600 python bench.py --model fixed --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2h_bench_fixed_minb$minb.json 2> gpurun_out/r2h_bench_fixed_minb$minb.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2h_bench_fixed_minb$minb.json'))
    print({k:d[k] for k in ('value','ms_per_step','stats')}, d['roofline']['frac'], d['roofline']['kernel_ms'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2h_bench_fixed_minb$minb.err').read()[-1500:])
PY
done
timeout 900 python scripts/cli_throughput.py --kmers 60000 > gpurun_out/r2h_cli_throughput.json 2> gpurun_out/r2h_cli_throughput.err; echo "cli exit $?"; cat gpurun_out/r2h_cli_throughput.json; tail -3 gpurun_out/r2h_cli_throughput.err
