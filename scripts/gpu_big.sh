#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lmm_gpu.py -m gpu -q --tb=short 2>&1 | tail -12 > gpurun_out/pytest_lmm.log
tail -4 gpurun_out/pytest_lmm.log | cut -c1-300
timeout 1500 python bench.py --samples 10000 --kmers-per-gpu 100000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-chunks 1 > gpurun_out/bench_lmm_n10000.json 2> gpurun_out/bench_lmm_n10000.err
cut -c1-200 gpurun_out/bench_lmm_n10000.json; grep -o '"check": {[^}]*}' gpurun_out/bench_lmm_n10000.json; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_lmm_n10000.json; tail -2 gpurun_out/bench_lmm_n10000.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_lmm_q.json 2>/dev/null; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_lmm_q.json
