#!/bin/bash
# compute-sanitizer memcheck / racecheck over the smoke test and the tensor-kernel mode-agreement
# test (PSB_TC_PAIR = 0 / 1 / 2 are exercised inside that test); logs -> gpurun_out/ -> profiles/
cd "$GRAFT_REPO_ROOT" || exit 1
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  echo "=== $tool smoke"
  timeout 1200 $CS --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_${tool}_smoke.log 2>&1
  echo "exit $?" >> gpurun_out/r02_sanitizer_${tool}_smoke.log
  tail -4 gpurun_out/r02_sanitizer_${tool}_smoke.log
  echo "=== $tool mode agreement (n=300)"
  timeout 1500 $CS --tool $tool --print-limit 20 python -m pytest "tests/test_lmm_gpu.py::test_tensor_kernel_modes_agree[300-700]" -q -x > gpurun_out/r02_sanitizer_${tool}_modes.log 2>&1
  echo "exit $?" >> gpurun_out/r02_sanitizer_${tool}_modes.log
  tail -4 gpurun_out/r02_sanitizer_${tool}_modes.log
  echo "=== $tool fast logit + comm (one GPU)"
  timeout 1500 $CS --tool $tool --print-limit 20 python -m pytest "tests/test_comm_gpu.py::test_gather_one_gpu" "tests/test_fixed_gpu.py::test_oracle_parity_fixed" -q -x > gpurun_out/r02_sanitizer_${tool}_fixed_comm.log 2>&1
  echo "exit $?" >> gpurun_out/r02_sanitizer_${tool}_fixed_comm.log
  tail -4 gpurun_out/r02_sanitizer_${tool}_fixed_comm.log
  echo "=== $tool kinship tensor path"
  timeout 600 $CS --tool $tool --print-limit 20 python -m pytest "tests/test_kinship_gpu.py::test_kinship_matches_numpy" "tests/test_kinship_gpu.py::test_kinship_with_missing_genotypes" -q -x > gpurun_out/r02_sanitizer_${tool}_kinship.log 2>&1
  echo "exit $?" >> gpurun_out/r02_sanitizer_${tool}_kinship.log
  tail -4 gpurun_out/r02_sanitizer_${tool}_kinship.log
done
