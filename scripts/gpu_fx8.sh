#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python bench.py --model fixed > gpurun_out/bench_fixed_10m.json 2> gpurun_out/bench_fixed_10m.err
cat gpurun_out/bench_fixed_10m.json; tail -2 gpurun_out/bench_fixed_10m.err
