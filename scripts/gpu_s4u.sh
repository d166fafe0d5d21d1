#!/bin/bash
# session 4, call U: refresh the secondary bench lines with the final kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --model fixed-cont --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ols.json 2> gpurun_out/bench_ols.err
cut -c1-170 gpurun_out/bench_ols.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_ols.json; grep -o '"check": {[^}]*}' gpurun_out/bench_ols.json; tail -2 gpurun_out/bench_ols.err
timeout 600 python bench.py --samples 1000 --kmers-per-gpu 1000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1000.json 2> gpurun_out/bench_n1000.err
cut -c1-170 gpurun_out/bench_n1000.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n1000.json; grep -o '"check": {[^}]*}' gpurun_out/bench_n1000.json; grep -o '"kernel_ms": [0-9.]*, "run_ms": [0-9.]*' gpurun_out/bench_n1000.json; tail -2 gpurun_out/bench_n1000.err
timeout 600 python bench.py --samples 2000 --kmers-per-gpu 4000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_n2000.json 2> gpurun_out/bench_n2000.err
cut -c1-170 gpurun_out/bench_n2000.json; grep -o '"kernel_ms": [0-9.]*, "run_ms": [0-9.]*' gpurun_out/bench_n2000.json; tail -2 gpurun_out/bench_n2000.err
