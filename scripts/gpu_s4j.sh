#!/bin/bash
# session 4, call J: two-SM MMA mode (cta_group::2)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/tc_diag.py 64:100,1000:1000,1500:5000 5,4,6,7,3 > gpurun_out/tc_diag_j.log 2>&1
grep -v "^  File\|^    \|first ratios\|ratios @\|^ 1\." gpurun_out/tc_diag_j.log | tail -16
PSB_TC_GLOBAL_BITS=1 timeout 300 python scripts/tc_diag.py 1000:1000 5 2>&1 | tail -1
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA > gpurun_out/bench_2sm_$tag.json 2> gpurun_out/bench_2sm_$tag.err
  echo "$tag: $(grep -o '"value": [0-9.]*' gpurun_out/bench_2sm_$tag.json | head -1) $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_2sm_$tag.json) $(grep -o '"sm_mhz": [0-9.]*' gpurun_out/bench_2sm_$tag.json) $(grep -o '"check": {[^}]*}' gpurun_out/bench_2sm_$tag.json)"
  tail -2 gpurun_out/bench_2sm_$tag.err
}
run two PSB_X=0
run single PSB_TC_PAIR=0
EXTRA="--precision 4" run two_k4 PSB_X=0
EXTRA="--samples 10000 --kmers-per-gpu 400000" run two_n10k PSB_X=0
timeout 900 python -m pytest tests/test_lmm_gpu.py tests/test_properties_gpu.py tests/test_burden_gpu.py -m gpu -q --tb=short 2>&1 | tail -8 | cut -c1-400
