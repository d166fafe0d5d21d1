#!/bin/bash
# session 4, call K: why is the two-SM mode slow? knock-outs + ncu sampling
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --kmers-per-gpu 2000000 $EXTRA > gpurun_out/bench_ko3_$tag.json 2> gpurun_out/bench_ko3_$tag.err
  echo "$tag: $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_ko3_$tag.json) $(grep -o '"sm_mhz": [0-9.]*' gpurun_out/bench_ko3_$tag.json)"
  tail -1 gpurun_out/bench_ko3_$tag.err
}
run single PSB_TC_PAIR=0
run two PSB_TC_PAIR=2
run two_mma1 PSB_TC_PAIR=2 PSB_TC_DEBUG=1
run two_noexp PSB_TC_PAIR=2 PSB_TC_DEBUG=2
run two_noepi PSB_TC_PAIR=2 PSB_TC_DEBUG=4
run two_all PSB_TC_PAIR=2 PSB_TC_DEBUG=7
run two_st3 PSB_TC_PAIR=2 PSB_TC_STAGES=3
PSB_TC_PAIR=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lmm_quadform_tc -s 1 -c 1 -o gpurun_out/prof_tc_two python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --kmers-per-gpu 2000000 > gpurun_out/ncu_tc_two.log 2>&1
tail -2 gpurun_out/ncu_tc_two.log | cut -c1-200
