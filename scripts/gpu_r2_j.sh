#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_fixed_gpu.py tests/test_lmm_gpu.py tests/test_baseline_sizes_gpu.py::test_config2_fixed_n2000_logit_firth -q -s > gpurun_out/r2j_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2j_tests.log; grep -n "configs\[\|passed\|failed\|^E  " gpurun_out/r2j_tests.log | head
for prec in 5 46; do
PSB_TEST_BASELINE_PRECISION=$prec timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py -q -s -k "config3 or config1 or adversarial or config4" > gpurun_out/r2j_base$prec.log 2>&1; echo "exit $?" >> gpurun_out/r2j_base$prec.log; grep -n "worst rel\|passed\|failed\|^E  " gpurun_out/r2j_base$prec.log | cut -c1-400 | head -20
done
# profiles: launch list of the default LMM step, full capture of the fast Logit kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_lmm_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2j_ncu_lmm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fixed_logit_fast -s 1 -c 1 -o gpurun_out/r02_logit_fast python bench.py --model fixed --kmers-per-gpu 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2j_ncu_logit.log 2>&1
ls -la gpurun_out/*.ncu-rep
