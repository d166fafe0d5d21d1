#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/r2b_suite.log 2>&1
echo "exit $?" >> gpurun_out/r2b_suite.log
tail -40 gpurun_out/r2b_suite.log
