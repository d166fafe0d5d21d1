#!/bin/bash
# session 4, call L: full suite + default bench (with e2e, cpu baseline) + reference arm + ncu evidence for the two-SM kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default2.json 2> gpurun_out/bench_default2.err
cut -c1-200 gpurun_out/bench_default2.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_default2.json; grep -o '"check": {[^}]*}' gpurun_out/bench_default2.json; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_default2.json; grep -o '"clocks": {[^}]*}' gpurun_out/bench_default2.json; grep -o '"setup_s": [0-9.]*' gpurun_out/bench_default2.json; tail -2 gpurun_out/bench_default2.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference2.json 2> gpurun_out/bench_reference2.err
cut -c1-120 gpurun_out/bench_reference2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_tc_2sm.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc_launch2.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_lmm_quadform_tc -s 1 -c 1 -o gpurun_out/prof_tc_2sm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc_full2.log 2>&1
tail -2 gpurun_out/ncu_tc_full2.log | cut -c1-200
timeout 600 python bench.py --model fixed --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_fixed2.json 2> gpurun_out/bench_fixed2.err
cut -c1-160 gpurun_out/bench_fixed2.json; tail -2 gpurun_out/bench_fixed2.err
