#!/bin/bash
# the whole -m gpu suite on the final code, then the CLI throughput table
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r2z_tests.log
timeout 900 python scripts/cli_throughput.py --samples 5000 --kmers 160000 > gpurun_out/r02_cli_throughput.json 2> gpurun_out/r02_cli_throughput.err
echo "cli exit $?"; tail -14 gpurun_out/r02_cli_throughput.err
