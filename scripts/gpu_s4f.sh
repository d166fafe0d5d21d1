#!/bin/bash
# session 4, call F: knock-out timing experiments on the tensor kernel (results invalid when PSB_TC_DEBUG != 0)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --kmers-per-gpu 2000000 $EXTRA > gpurun_out/bench_ko_$tag.json 2> gpurun_out/bench_ko_$tag.err
  echo "$tag: $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_ko_$tag.json) $(grep -o '"sm_mhz": [0-9.]*' gpurun_out/bench_ko_$tag.json)"
  tail -1 gpurun_out/bench_ko_$tag.err
}
run base PSB_TC_DEBUG=0
run mma1 PSB_TC_DEBUG=1
run noexp PSB_TC_DEBUG=2
run noepi PSB_TC_DEBUG=4
run mma1_noexp PSB_TC_DEBUG=3
run noexp_noepi PSB_TC_DEBUG=6
run all PSB_TC_DEBUG=7
