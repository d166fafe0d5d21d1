#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_fixed_gpu.py tests/test_baseline_sizes_gpu.py::test_config2_fixed_n2000_logit_firth -x -q -s --durations=5 > gpurun_out/r2c_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2c_tests.log
tail -30 gpurun_out/r2c_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
echo "bench exit $?"; tail -5 gpurun_out/r2c_bench_default.err; head -c 6000 gpurun_out/r2c_bench_default.json
