#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/tc_diag.py 64:100,200:300,1000:1000,1500:5000 5,4,3,6,7 > gpurun_out/tc_diag.log 2>&1
echo "rc=$?" >> gpurun_out/tc_diag.log
cat gpurun_out/tc_diag.log
