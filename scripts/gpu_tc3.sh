#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/tc_diag.py 64:100,1000:1000,1500:5000 5,4,7 > gpurun_out/tc_diag.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_tc.log
timeout 900 python bench.py --precision 5 --no-cpu-baseline > gpurun_out/bench_tc_p5.json 2> gpurun_out/bench_tc_p5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_tc.csv python bench.py --precision 5 --kmers-per-gpu 606208 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc_launch.log 2>&1
cat gpurun_out/tc_diag.log; tail -5 gpurun_out/pytest_tc.log; cat gpurun_out/bench_tc_p5.json; tail -3 gpurun_out/bench_tc_p5.err
grep -E "k_bitstats|quadform|k_prefilter|epilogue" gpurun_out/launches_tc.csv | awk -F'","' '{print $5, $NF}' | cut -c1-40,200- | head -8
