#!/bin/bash
# round 2, call A: parity at the BASELINE sizes, then the whole GPU suite
cd "$GRAFT_REPO_ROOT" || exit 1
nproc > gpurun_out/r2a_nproc.txt; free -g >> gpurun_out/r2a_nproc.txt
timeout 1500 python -m pytest tests/test_baseline_sizes_gpu.py -x -q -s --durations=0 > gpurun_out/r2a_baseline_sizes.log 2>&1
echo "exit $?" >> gpurun_out/r2a_baseline_sizes.log
tail -40 gpurun_out/r2a_baseline_sizes.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_baseline_sizes_gpu.py --durations=10 > gpurun_out/r2a_suite.log 2>&1
echo "exit $?" >> gpurun_out/r2a_suite.log
tail -15 gpurun_out/r2a_suite.log
