#!/bin/bash
# fast Logit kernel with bulk-copy staging: parity tests + configs[2] bench
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_fixed_gpu.py tests/test_baseline_sizes_gpu.py::test_config2_fixed_n2000_logit_firth -q -x > gpurun_out/r2p_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2p_tests.log; tail -4 gpurun_out/r2p_tests.log | cut -c1-300
timeout 600 python bench.py --model fixed --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_bench_fixed.json 2> gpurun_out/r2p_bench_fixed.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2p_bench_fixed.json'))
    print({k:d[k] for k in ('value','ms_per_step','stats')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['clocks'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2p_bench_fixed.err').read()[-1500:])
PY
