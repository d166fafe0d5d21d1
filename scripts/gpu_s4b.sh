#!/bin/bash
# session 4, call B: operand-ring depth experiment on the triangular tensor kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/tc_diag.py 64:100,1000:1000,1500:5000 5,4 > gpurun_out/tc_diag_b.log 2>&1
tail -12 gpurun_out/tc_diag_b.log
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_ring_$tag.json 2> gpurun_out/bench_ring_$tag.err
  echo "$tag: $(grep -o '"value": [0-9.]*' gpurun_out/bench_ring_$tag.json | head -1) $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_ring_$tag.json) $(grep -o '"check": {[^}]*}' gpurun_out/bench_ring_$tag.json) $(grep -o '"sm_mhz": [0-9.]*' gpurun_out/bench_ring_$tag.json)"
  tail -2 gpurun_out/bench_ring_$tag.err
}
run default PSB_X=0
run smem7 PSB_TC_BITS_SMEM=1
run smem6 PSB_TC_BITS_SMEM=1 PSB_TC_STAGES=6
run glob8 PSB_TC_BITS_SMEM=0 PSB_TC_STAGES=8
run glob6 PSB_TC_BITS_SMEM=0 PSB_TC_STAGES=6
timeout 600 python -m pytest tests/test_lmm_gpu.py tests/test_properties_gpu.py -m gpu -q --tb=short 2>&1 | tail -5 | cut -c1-300
