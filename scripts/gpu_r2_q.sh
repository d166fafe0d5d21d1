#!/bin/bash
# skipped zero box in diagonal stages: LMM tests, bench, full ncu capture of the k=4 kernel
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_lmm_gpu.py tests/test_properties_gpu.py tests/test_fixed_gpu.py -q -x > gpurun_out/r2q_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2q_tests.log; tail -4 gpurun_out/r2q_tests.log | cut -c1-300
PSB_TEST_BASELINE_PRECISION=46 timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py -q -s -k "config3 or adversarial" > gpurun_out/r2q_base46.log 2>&1; echo "exit $?" >> gpurun_out/r2q_base46.log; grep -n "worst rel\|passed\|failed\|^E  " gpurun_out/r2q_base46.log | cut -c1-500 | head -20
for prec in 46; do
timeout 600 python bench.py --precision $prec --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r2q_bench_lmm_p$prec.json 2> gpurun_out/r2q_bench_lmm_p$prec.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2q_bench_lmm_p$prec.json'))
    r=d['roofline']
    print('prec $prec', {k:d[k] for k in ('value','ms_per_step')}, r['frac'], r['kernel_ms'], r['side_kernels_ms'], d['check'], d['clocks'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2q_bench_lmm_p$prec.err').read()[-1500:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lmm_quadform_tc -s 1 -c 1 -o gpurun_out/r02_lmm_tc_k4 python bench.py --precision 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2q_ncu_tc.log 2>&1
ls -la gpurun_out/*.ncu-rep
