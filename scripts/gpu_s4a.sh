#!/bin/bash
# session 4, call A: full GPU suite, smoke, default bench + reference arm, launch list and full
# ncu capture of the triangular tensor kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/pytest_all.log
tail -8 gpurun_out/pytest_all.log | cut -c1-500
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cut -c1-250 gpurun_out/bench_default.json; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cut -c1-400 gpurun_out/bench_reference.json; tail -2 gpurun_out/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_tc_tri.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc_launch.log 2>&1
tail -5 gpurun_out/launches_tc_tri.csv | cut -c1-200
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_lmm_quadform_tc -s 1 -c 1 -o gpurun_out/prof_tc_tri python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc_full.log 2>&1
tail -3 gpurun_out/ncu_tc_full.log | cut -c1-300
ls -la gpurun_out | tail -20
