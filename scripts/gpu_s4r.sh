#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
PYSEER_B200_EIGH=numpy timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/launches_vpw8.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_vpw8.log 2>&1
grep "bitstats\|prefilter\|epilogue" gpurun_out/launches_vpw8.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -6
timeout 600 python -m pytest tests/test_lmm_gpu.py tests/test_fixed_gpu.py -m gpu -q --tb=short 2>&1 | tail -4 | cut -c1-300
