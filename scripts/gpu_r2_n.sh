#!/bin/bash
# side kernels (vectorised popcount pass, series/Wallis t-distribution tail), expanders that expand
# before they wait: tests, bench at precision 46 / 5 / 4, launch list, full capture of the k=4 kernel
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_lmm_gpu.py tests/test_properties_gpu.py tests/test_burden_gpu.py -q -x > gpurun_out/r2n_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2n_tests.log; tail -5 gpurun_out/r2n_tests.log | cut -c1-300
for prec in 46 5 4; do
timeout 600 python bench.py --precision $prec --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r2n_bench_lmm_p$prec.json 2> gpurun_out/r2n_bench_lmm_p$prec.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2n_bench_lmm_p$prec.json'))
    r=d['roofline']
    print('prec $prec', {k:d[k] for k in ('value','ms_per_step')}, r['frac'], r['kernel_ms'], r['side_kernels_ms'], d['check'], d['clocks'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2n_bench_lmm_p$prec.err').read()[-1500:])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_lmm_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2n_ncu_lmm.log 2>&1
grep -c "k_lmm_quadform_tc" gpurun_out/r02_lmm_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lmm_quadform_tc -s 2 -c 1 -o gpurun_out/r02_lmm_tc_k4 python bench.py --precision 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2n_ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_bitstats_v -s 1 -c 1 -o gpurun_out/r02_bitstats_v python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2n_ncu_bs.log 2>&1
ls -la gpurun_out/*.ncu-rep
