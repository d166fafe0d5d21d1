#!/bin/bash
# new GPU tests (pattern digests, text) then the CLI throughput sweep with the final readers
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_text_gpu.py tests/test_kinship_gpu.py -x -q -m gpu > gpurun_out/r2z_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r2z_tests.log
timeout 900 python scripts/cli_throughput.py --samples 5000 --kmers 160000 --sweep 12000 > gpurun_out/r02_cli_throughput.json 2> gpurun_out/r02_cli_throughput.err
echo "cli exit $?"; tail -24 gpurun_out/r02_cli_throughput.err
