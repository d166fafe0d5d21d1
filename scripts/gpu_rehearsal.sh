#!/bin/bash
# round-end rehearsal (what the driver runs) on one GPU: the whole -m gpu suite, smoke, default bench, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2y_tests.log; tail -5 gpurun_out/r2y_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "bench exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
def show(x):
    print(x['config']['workload'][:60], '| value %.4g e2e %.4g ms %.2f frac %.3f cpu %s' % (x['value'], x['e2e']['value'], x['ms_per_step'], x['roofline']['frac'], (x.get('cpu_baseline') or {}).get('value')), x.get('stats'), x.get('check'), x.get('clocks',{}).get('sm_mhz'))
show(d)
for s in d.get('secondary', []):
    if 'error' in s: print('ERR', s['error'])
    else: show(s)
print('gpu_launches', d.get('gpu_launches'), 'steps', d['steps'], 'warmup', d['warmup'])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref exit $?"; head -c 600 gpurun_out/r02_bench_reference.json; echo
