#!/bin/bash
# session 4, call I: knock-outs at 256-sample stages; N=10000 pair vs no pair
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --kmers-per-gpu 2000000 $EXTRA > gpurun_out/bench_ko2_$tag.json 2> gpurun_out/bench_ko2_$tag.err
  echo "$tag: $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_ko2_$tag.json) $(grep -o '"sm_mhz": [0-9.]*' gpurun_out/bench_ko2_$tag.json)"
  tail -1 gpurun_out/bench_ko2_$tag.err
}
export PSB_TC_PAIR=0
run base PSB_TC_DEBUG=0
run mma1 PSB_TC_DEBUG=1
run noexp PSB_TC_DEBUG=2
run noepi PSB_TC_DEBUG=4
run all PSB_TC_DEBUG=7
EXTRA="--samples 10000 --kmers-per-gpu 400000" 
run n10k_nopair PSB_TC_PAIR=0
run n10k_pair PSB_TC_PAIR=1
