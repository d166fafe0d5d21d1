#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fixed_gpu.py -m gpu -q --tb=line 2>&1 | tail -15 > gpurun_out/pytest_fx.log
timeout 900 python bench.py --model fixed --kmers-per-gpu 1000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_fixed_1m.json 2> gpurun_out/bench_fixed_1m.err
tail -6 gpurun_out/pytest_fx.log; cat gpurun_out/bench_fixed_1m.json | cut -c1-1800; tail -3 gpurun_out/bench_fixed_1m.err
