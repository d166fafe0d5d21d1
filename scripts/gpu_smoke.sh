#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 1500 python bench.py --samples 10000 --kmers-per-gpu 100000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-chunks 1 > gpurun_out/bench_lmm_n10000.json 2> gpurun_out/bench_lmm_n10000.err
cut -c1-300 gpurun_out/bench_lmm_n10000.json; grep -o '"check": {[^}]*}' gpurun_out/bench_lmm_n10000.json; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_lmm_n10000.json; grep -o '"setup_s": [0-9.]*' gpurun_out/bench_lmm_n10000.json; tail -2 gpurun_out/bench_lmm_n10000.err
