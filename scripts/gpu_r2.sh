#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/tc_diag.py 64:100,1000:1000,1500:5000 5,4 > gpurun_out/tc_diag.log 2>&1
cat gpurun_out/tc_diag.log
for i in 1 2 3; do
timeout 900 python -m pytest tests/test_lmm_gpu.py -m gpu -q --tb=short 2>&1 | tail -4 | cut -c1-300
done > gpurun_out/pytest_lmm_rep.log
cat gpurun_out/pytest_lmm_rep.log
timeout 900 python -m pytest tests/test_properties_gpu.py tests/test_cli_gpu.py -m gpu -q --tb=short 2>&1 | tail -6 | cut -c1-300 > gpurun_out/pytest_prop.log
cat gpurun_out/pytest_prop.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_lmm_int.json 2> gpurun_out/bench_lmm_int.err
cut -c1-220 gpurun_out/bench_lmm_int.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_lmm_int.json; grep -o '"check": {[^}]*}' gpurun_out/bench_lmm_int.json; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_lmm_int.json;  grep -o '"clocks": {[^}]*}' gpurun_out/bench_lmm_int.json; tail -2 gpurun_out/bench_lmm_int.err
