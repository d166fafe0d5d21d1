#!/bin/bash
# 2-GPU box: whole GPU suite (incl. the 2-GPU gather / CLI tests), then bench at N=2 under torchrun
cd "$GRAFT_REPO_ROOT" || exit 1
nvidia-smi -L > gpurun_out/r2d_gpus.txt
timeout 1500 python -m pytest tests -q -m gpu --durations=8 -rs > gpurun_out/r2d_suite_2gpu.log 2>&1
echo "exit $?" >> gpurun_out/r2d_suite_2gpu.log
tail -25 gpurun_out/r2d_suite_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2d_bench_2gpu.json 2> gpurun_out/r2d_bench_2gpu.err
echo "bench exit $?"; tail -5 gpurun_out/r2d_bench_2gpu.err; head -c 3000 gpurun_out/r2d_bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --parser-kmers 300 > gpurun_out/r2d_bench_ref_2gpu.json 2> gpurun_out/r2d_bench_ref_2gpu.err
echo "ref exit $?"; head -c 1500 gpurun_out/r2d_bench_ref_2gpu.json
