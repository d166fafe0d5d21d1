#!/bin/bash
# CLI-level throughput on one GPU box: real input formats at N=5000 (plain / bgzip / gzip k-mer text
# through the device parser and the host parser, --output-patterns, --bits-cache written and read) and
# the whole CLI over a packed cache of 20 M variants
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python scripts/cli_throughput.py --samples 5000 --kmers 160000 --sweep 12000 > gpurun_out/r02_cli_throughput.json 2> gpurun_out/r02_cli_throughput.err
echo "cli exit $?"; tail -24 gpurun_out/r02_cli_throughput.err
timeout 600 python scripts/cache_throughput.py --kmers 20000000 --quick > gpurun_out/r02_cache_throughput_20m.json 2> gpurun_out/r02_cache_throughput_20m.err
echo "cache exit $?"; tail -4 gpurun_out/r02_cache_throughput_20m.err
