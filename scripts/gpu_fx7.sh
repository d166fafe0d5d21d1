#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/pytest_all.log
timeout 900 python bench.py --model fixed-cont --kmers-per-gpu 10000000 --steps 3 --warmup 2 > gpurun_out/bench_fixedcont.json 2> gpurun_out/bench_fixedcont.err
tail -5 gpurun_out/pytest_all.log | cut -c1-300; cat gpurun_out/bench_fixedcont.json; tail -2 gpurun_out/bench_fixedcont.err
