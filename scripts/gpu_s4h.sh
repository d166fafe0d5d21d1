#!/bin/bash
# session 4, call H: 256-sample stages, all precisions, full GPU suite, pair vs no pair
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/tc_diag.py 64:100,1000:1000,1500:5000 5,4,6,7,3 > gpurun_out/tc_diag_h.log 2>&1
grep -v "^  File\|^    " gpurun_out/tc_diag_h.log | tail -16
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12 | cut -c1-400
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA > gpurun_out/bench_k256b_$tag.json 2> gpurun_out/bench_k256b_$tag.err
  echo "$tag: $(grep -o '"value": [0-9.]*' gpurun_out/bench_k256b_$tag.json | head -1) $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_k256b_$tag.json) $(grep -o '"sm_mhz": [0-9.]*' gpurun_out/bench_k256b_$tag.json)"
  tail -2 gpurun_out/bench_k256b_$tag.err
}
run pair PSB_X=0
run nopair PSB_TC_PAIR=0
run pair2 PSB_X=0
run nopair2 PSB_TC_PAIR=0
