#!/bin/bash
# two GPUs of one box: the multi-GPU tests (gathered table == single-GPU table), the CLI with --gpus 2
# on the device text parser, bench.py under torchrun
cd "$GRAFT_REPO_ROOT" || exit 1
nvidia-smi -L
timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_text_gpu.py tests/test_cli_gpu.py::test_two_gpus_print_the_same_table "tests/test_fixed_gpu.py::test_popcount_pass_variants_agree" -q -x -rs > gpurun_out/r02_two_gpu_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_two_gpu_tests.log; tail -8 gpurun_out/r02_two_gpu_tests.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
echo "bench exit $?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_2gpu.json'))
    print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d.get('gather'))
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02_bench_2gpu.err').read()[-2000:])
PY
