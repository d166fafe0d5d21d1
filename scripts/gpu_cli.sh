#!/bin/bash
# CLI throughput on real input formats (incl. chunk-parallel gzip), then the kinship profile
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 800 python scripts/cli_throughput.py --samples 5000 --kmers 160000 > gpurun_out/r02_cli_throughput.json 2> gpurun_out/r02_cli_throughput.err
echo "cli exit $?"; tail -12 gpurun_out/r02_cli_throughput.err
bash scripts/gpu_kin_profile.sh
