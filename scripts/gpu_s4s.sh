#!/bin/bash
# session 4, call S: final full suite + smoke + default bench on the committed tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short --durations=5 2>&1 | tail -12 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
cut -c1-200 gpurun_out/bench_final.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_final.json; grep -o '"check": {[^}]*}' gpurun_out/bench_final.json; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_final.json; grep -o '"executed_int8_tops": [0-9.]*' gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
