#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -8 > gpurun_out/pytest_all.log
timeout 900 python bench.py --model fixed --kmers-per-gpu 1000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_fixed_1m.json 2> gpurun_out/bench_fixed_1m.err
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_lmm.json 2> gpurun_out/bench_lmm.err
tail -4 gpurun_out/pytest_all.log; cut -c1-330 gpurun_out/bench_fixed_1m.json; tail -2 gpurun_out/bench_fixed_1m.err;  cut -c1-330 gpurun_out/bench_lmm.json; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_lmm.json
