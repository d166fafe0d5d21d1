#!/bin/bash
# session 4, call N: launch list of the two-SM kernel (host eigh so the set-up does not flood the list), eigh tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
PYSEER_B200_EIGH=numpy timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_tc_2sm.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc_launch2.log 2>&1
grep -c quadform gpurun_out/launches_tc_2sm.csv
timeout 600 python -m pytest tests/test_eigh_gpu.py tests/test_burden_gpu.py -m gpu -q --tb=short --durations=3 2>&1 | tail -8 | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_dev_eigh.json 2> gpurun_out/bench_dev_eigh.err
grep -o '"value": [0-9.]*' gpurun_out/bench_dev_eigh.json | head -1; grep -o '"setup_s": [0-9.]*' gpurun_out/bench_dev_eigh.json; grep -o '"executed_int8_tops": [0-9.]*' gpurun_out/bench_dev_eigh.json; tail -2 gpurun_out/bench_dev_eigh.err
