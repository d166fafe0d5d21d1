#!/bin/bash
# session 4, call Q: burden workload (configs[4] shard), mode-agreement test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lmm_gpu.py -m gpu -q --tb=short -k "modes_agree" 2>&1 | tail -5 | cut -c1-300
timeout 900 python bench.py --model burden --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_burden.json 2> gpurun_out/bench_burden.err
cut -c1-200 gpurun_out/bench_burden.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_burden.json; grep -o '"check": {[^}]*}' gpurun_out/bench_burden.json; grep -o '"burden_or": {[^}]*}' gpurun_out/bench_burden.json; grep -o '"counts": {[^}]*}' gpurun_out/bench_burden.json; grep -o '"kernel_ms": [0-9.]*, "run_ms": [0-9.]*' gpurun_out/bench_burden.json; grep -o '"setup_s": [0-9.]*' gpurun_out/bench_burden.json; tail -3 gpurun_out/bench_burden.err
timeout 900 python bench.py --model burden --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --kmers-per-gpu 100000 > gpurun_out/bench_burden_100k.json 2> gpurun_out/bench_burden_100k.err
cut -c1-200 gpurun_out/bench_burden_100k.json; grep -o '"burden_or": {[^}]*}' gpurun_out/bench_burden_100k.json; grep -o '"kernel_ms": [0-9.]*, "run_ms": [0-9.]*' gpurun_out/bench_burden_100k.json; tail -3 gpurun_out/bench_burden_100k.err
