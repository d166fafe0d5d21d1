#!/bin/bash
# device text parser: parity tests, CLI replays, CLI throughput at N = 5000
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_text_gpu.py tests/test_cli_gpu.py tests/test_comm_gpu.py -q -x > gpurun_out/r2o_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2o_tests.log; tail -15 gpurun_out/r2o_tests.log | cut -c1-300
nproc; free -g | head -2
timeout 1500 python scripts/cli_throughput.py --kmers 160000 > gpurun_out/r2o_cli_throughput.json 2> gpurun_out/r2o_cli_throughput.err
echo "cli exit $?"; tail -12 gpurun_out/r2o_cli_throughput.err | cut -c1-300
