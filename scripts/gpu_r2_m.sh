#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "bench exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
def show(x):
    print(x['config']['workload'][:60], '| value %.3g e2e %.3g ms %.2f frac %.3f cpu %s' % (x['value'], x['e2e']['value'], x['ms_per_step'], x['roofline']['frac'], (x.get('cpu_baseline') or {}).get('value')), x.get('stats'), x.get('check'))
show(d)
for s in d.get('secondary', []):
    if 'error' in s: print('ERR', s['error'])
    else: show(s)
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref exit $?"; head -c 1800 gpurun_out/r02_bench_reference.json; echo
bash scripts/gpu_sanitizer.sh
