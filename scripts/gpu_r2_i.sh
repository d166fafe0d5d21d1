#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for env in "PSB_LOGIT_FAST=0" "PSB_LOGIT_FAST=1"; do
  echo "=== $env"; env $env timeout 300 python scripts/dbg_v60949.py 52213 2>&1 | tail -4
done
echo "=== LMM precision 46 tests"
PSB_TEST_PRECISIONS=46 timeout 600 python -m pytest tests/test_lmm_gpu.py -q -x > gpurun_out/r2i_lmm46.log 2>&1; echo "exit $?" >> gpurun_out/r2i_lmm46.log; tail -5 gpurun_out/r2i_lmm46.log
PSB_TEST_BASELINE_PRECISION=46 timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py -q -s -k "config3 or config1 or adversarial" > gpurun_out/r2i_base46.log 2>&1; echo "exit $?" >> gpurun_out/r2i_base46.log; grep -n "worst\|passed\|failed\|^E  " gpurun_out/r2i_base46.log | head -20
for prec in 46 5; do
  echo "=== LMM bench precision $prec"
  timeout 600 python bench.py --precision $prec --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2i_bench_lmm_p$prec.json 2> gpurun_out/r2i_bench_lmm_p$prec.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2i_bench_lmm_p$prec.json'))
    print({k:d[k] for k in ('value','ms_per_step','check','clocks')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2i_bench_lmm_p$prec.err').read()[-1500:])
PY
done
