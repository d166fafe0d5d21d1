#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_cli_gpu.py -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/pytest_cli.log
cat gpurun_out/pytest_cli.log | cut -c1-600
