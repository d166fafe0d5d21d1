#!/bin/bash
# kinship tensor path: full ncu capture of k_kin_tc, compute-sanitizer memcheck / racecheck over its tests
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 400 ncu --clock-control none --set full --import-source on -k regex:k_kin_tc -s 4 -c 1 -o gpurun_out/r02_kin_tc \
    python scripts/kinship_throughput.py 5000 200000 > gpurun_out/ncu_kin.log 2>&1
ls -la gpurun_out/r02_kin_tc.ncu-rep
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 600 $CS --tool $tool --print-limit 20 python -m pytest "tests/test_kinship_gpu.py::test_kinship_matches_numpy" "tests/test_kinship_gpu.py::test_kinship_with_missing_genotypes" -q -x > gpurun_out/r02_sanitizer_${tool}_kinship.log 2>&1
  echo "exit $?" >> gpurun_out/r02_sanitizer_${tool}_kinship.log
  tail -6 gpurun_out/r02_sanitizer_${tool}_kinship.log
done
