#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/pytest_all.log
cat gpurun_out/pytest_all.log | cut -c1-700
