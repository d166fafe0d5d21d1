#!/bin/bash
# session 4: Welch sums through the tensor pass (deferred pre-filter) -- parity and timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYSEER_B200_EIGH=numpy
timeout 600 python -m pytest tests/test_lmm_gpu.py tests/test_properties_gpu.py tests/test_cli_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -12 | cut -c1-400
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e $EXTRA > gpurun_out/bench_w_$tag.json 2> gpurun_out/bench_w_$tag.err
  echo "$tag: $(grep -o '"value": [0-9.]*' gpurun_out/bench_w_$tag.json | head -1) $(grep -o '"kernel_ms": [0-9.]*, "run_ms": [0-9.]*' gpurun_out/bench_w_$tag.json) $(grep -o '"counts": {[^}]*}' gpurun_out/bench_w_$tag.json)"
  tail -1 gpurun_out/bench_w_$tag.err
}
run tc PSB_X=0
run cuda PSB_WELCH_TC=0
EXTRA="--samples 1000 --kmers-per-gpu 1000000" run tc_n1000 PSB_X=0
EXTRA="--samples 1000 --kmers-per-gpu 1000000" run cuda_n1000 PSB_WELCH_TC=0
