#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/fx_diag.py > gpurun_out/fx_diag.log 2>&1
cat gpurun_out/fx_diag.log
