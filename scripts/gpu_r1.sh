#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/pytest_all.log
tail -3 gpurun_out/pytest_all.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cut -c1-250 gpurun_out/bench_default.json; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/bench_default.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_tc_tri.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc_launch.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_lmm_quadform_tc -s 1 -c 1 -o gpurun_out/prof_tc_tri python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc_full.log 2>&1
tail -3 gpurun_out/ncu_tc_full.log | cut -c1-300
