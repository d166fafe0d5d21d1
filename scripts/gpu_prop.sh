#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_properties_gpu.py -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/pytest_prop.log
cat gpurun_out/pytest_prop.log | cut -c1-700
timeout 900 python bench.py --samples 1000 --kmers-per-gpu 1000000 > gpurun_out/bench_lmm_n1000.json 2> gpurun_out/bench_lmm_n1000.err
cut -c1-1900 gpurun_out/bench_lmm_n1000.json; tail -2 gpurun_out/bench_lmm_n1000.err
