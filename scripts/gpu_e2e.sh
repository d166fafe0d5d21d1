#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/pytest_all.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_lmm.json 2> gpurun_out/bench_lmm.err
timeout 900 python bench.py --no-cpu-baseline --e2e-chunks 1 --steps 2 --warmup 1 > gpurun_out/bench_lmm_c1.json 2> gpurun_out/bench_lmm_c1.err
tail -5 gpurun_out/pytest_all.log | cut -c1-300; cut -c1-200 gpurun_out/bench_lmm.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_lmm.json gpurun_out/bench_lmm_c1.json; grep -o '"check": {[^}]*}' gpurun_out/bench_lmm.json; tail -2 gpurun_out/bench_lmm.err
