#!/bin/bash
# session 4, call X: e2e chunking and ring depth variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline $EXTRA > gpurun_out/bench_x_$tag.json 2> gpurun_out/bench_x_$tag.err
  echo "$tag: $(grep -o '"value": [0-9.]*' gpurun_out/bench_x_$tag.json | head -1) e2e $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_x_$tag.json) $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_x_$tag.json)"
  tail -1 gpurun_out/bench_x_$tag.err
}
EXTRA="--e2e-chunks 8" run c8 PSB_X=0
EXTRA="--e2e-chunks 4" run c4 PSB_X=0
EXTRA="--e2e-chunks 16" run c16 PSB_X=0
EXTRA="--e2e-chunks 32" run c32 PSB_X=0
