#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -15 > gpurun_out/pytest_all.log
timeout 900 python bench.py --model fixed --kmers-per-gpu 1000000 --steps 2 --warmup 1 > gpurun_out/bench_fixed_1m.json 2> gpurun_out/bench_fixed_1m.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fixed_logit -s 1 -c 1 -o gpurun_out/prof_logit python bench.py --model fixed --kmers-per-gpu 100000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_logit.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_fixed.csv python bench.py --model fixed --kmers-per-gpu 200000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fx_launch.log 2>&1
tail -6 gpurun_out/pytest_all.log; cat gpurun_out/bench_fixed_1m.json; tail -3 gpurun_out/bench_fixed_1m.err
