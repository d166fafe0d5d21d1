#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fixed_gpu.py -m gpu -q --tb=line 2>&1 | tail -30 > gpurun_out/pytest_fx.log
cat gpurun_out/pytest_fx.log
