#!/bin/bash
# ncu evidence for profiles/ (run under gpurun on ONE GPU; numbers printed by runs under ncu are not
# bench values): launch list of the default LMM step, full captures of the dominant kernels.
cd "$GRAFT_REPO_ROOT" || exit 1
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_lmm_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/ncu_lmm_launches.log 2>&1
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_fixed_launches.csv \
    python bench.py --model fixed --kmers-per-gpu 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fixed_launches.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_lmm_quadform_tc -s 1 -c 1 -o gpurun_out/r02_lmm_tc_k4 \
    python bench.py --precision 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/ncu_tc.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_fixed_logit_fast -s 1 -c 1 -o gpurun_out/r02_logit_fast \
    python bench.py --model fixed --kmers-per-gpu 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_logit.log 2>&1
timeout 600 $NCU --set full -k regex:k_bitstats_stream -s 1 -c 1 -o gpurun_out/r02_bitstats_stream \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/ncu_bitstats.log 2>&1
# the kinship tensor kernel (once per run): timing of both contraction paths, launch list, full capture
timeout 300 python scripts/kinship_throughput.py 5000 200000 > gpurun_out/r02_kin_5000.json 2> gpurun_out/r02_kin_5000.err
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_kinship_launches.csv \
    python scripts/kinship_throughput.py 5000 200000 > gpurun_out/ncu_kin_launches.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:k_kin_tc -s 4 -c 1 -o gpurun_out/r02_kin_tc \
    python scripts/kinship_throughput.py 5000 200000 > gpurun_out/ncu_kin.log 2>&1
ls -la gpurun_out/*.ncu-rep
