#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --model fixed --kmers-per-gpu 1000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_fixed_1m.json 2> gpurun_out/bench_fixed_1m.err
timeout 900 python bench.py --model fixed-cont --kmers-per-gpu 2000000 --steps 2 --warmup 1 > gpurun_out/bench_fixedcont.json 2> gpurun_out/bench_fixedcont.err
cut -c1-260 gpurun_out/bench_fixed_1m.json; tail -2 gpurun_out/bench_fixed_1m.err; cat gpurun_out/bench_fixedcont.json; tail -2 gpurun_out/bench_fixedcont.err
