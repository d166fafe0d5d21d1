#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fixed_logit -s 1 -c 1 -o gpurun_out/prof_logit2 python bench.py --model fixed --kmers-per-gpu 200000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_logit.log 2>&1
tail -2 gpurun_out/ncu_logit.log | cut -c1-200
