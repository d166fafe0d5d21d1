#!/bin/bash
# the driver's scaling run, rehearsed: bench.py under torchrun on all GPUs of the box
cd "$GRAFT_REPO_ROOT" || exit 1
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "bench exit $?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_${N}gpu.json'))
    print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d.get('gather'))
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02_bench_${N}gpu.err').read()[-2500:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 1 --warmup 0 --parser-kmers 0 2>/dev/null | head -c 300; echo
