#!/bin/bash
# first GPU call: fp64 LMM path parity + first bench line + launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt; free -g >> gpurun_out/gpu_info.txt
PSB_TEST_PRECISIONS=0 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_fp64.log
timeout 900 python bench.py --precision 0 --kmers-per-gpu 200000 --steps 2 --warmup 3 > gpurun_out/bench_fp64.json 2> gpurun_out/bench_fp64.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fp64.csv python bench.py --precision 0 --samples 1000 --kmers-per-gpu 100000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/pytest_fp64.log; cat gpurun_out/bench_fp64.json; tail -3 gpurun_out/bench_fp64.err
