#!/bin/bash
# kinship tensor path: tests, timing of both paths, launch list (one GPU)
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 300 python -m pytest tests/test_kinship_gpu.py -x -q -m gpu > gpurun_out/r2u_kin_tests.log 2>&1; echo "tests exit $?"
tail -15 gpurun_out/r2u_kin_tests.log
timeout 200 python scripts/kinship_throughput.py 5000 200000 > gpurun_out/r2u_kin_5000.json 2> gpurun_out/r2u_kin_5000.err; echo "timing exit $?"
cat gpurun_out/r2u_kin_5000.json; tail -3 gpurun_out/r2u_kin_5000.err
timeout 200 python scripts/kinship_throughput.py 10000 100000 > gpurun_out/r2u_kin_10000.json 2> gpurun_out/r2u_kin_10000.err
cat gpurun_out/r2u_kin_10000.json
timeout 200 python scripts/kinship_throughput.py 1000 200000 > gpurun_out/r2u_kin_1000.json 2> gpurun_out/r2u_kin_1000.err
cat gpurun_out/r2u_kin_1000.json
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2u_kin_launches.csv \
    python scripts/kinship_throughput.py 5000 200000 > gpurun_out/r2u_kin_ncu.log 2>&1
grep -c k_kin gpurun_out/r2u_kin_launches.csv
