#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for B in 2 3 4; do
PSB_LOGIT_MINB=$B timeout 600 python bench.py --model fixed --kmers-per-gpu 1000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_fx_b$B.json 2>/dev/null
echo "MINB=$B $(grep -o '"value": [0-9.]*' gpurun_out/bench_fx_b$B.json | head -1) $(grep -o '"newton_evaluations_per_variant": [0-9.]*' gpurun_out/bench_fx_b$B.json)"
done
