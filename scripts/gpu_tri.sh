#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/tc_diag.py 64:100,1000:1000,1500:5000 5,4,6 > gpurun_out/tc_diag.log 2>&1
cat gpurun_out/tc_diag.log
timeout 900 python -m pytest tests/test_lmm_gpu.py tests/test_properties_gpu.py -m gpu -q --tb=short 2>&1 | tail -12 > gpurun_out/pytest_lmm.log
tail -6 gpurun_out/pytest_lmm.log | cut -c1-400
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_lmm_tri.json 2> gpurun_out/bench_lmm_tri.err
cut -c1-220 gpurun_out/bench_lmm_tri.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_lmm_tri.json; grep -o '"check": {[^}]*}' gpurun_out/bench_lmm_tri.json; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_lmm_tri.json;  grep -o '"clocks": {[^}]*}' gpurun_out/bench_lmm_tri.json; tail -2 gpurun_out/bench_lmm_tri.err
