#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_tc.log
for P in 5 4; do
timeout 900 python bench.py --precision $P > gpurun_out/bench_tc_p$P.json 2> gpurun_out/bench_tc_p$P.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_tc.csv python bench.py --precision 5 --kmers-per-gpu 606208 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lmm_quadform_tc -s 1 -c 1 -o gpurun_out/prof_tc python bench.py --precision 5 --kmers-per-gpu 75776 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc_full.log 2>&1
tail -5 gpurun_out/pytest_tc.log; cat gpurun_out/bench_tc_p5.json; tail -3 gpurun_out/bench_tc_p5.err; cat gpurun_out/bench_tc_p4.json
