#!/usr/bin/env python
"""ncu report -> the JSON summaries kept under profiles/:

    python scripts/ncu_summary.py gpurun_out/r02_lmm_tc_k4.ncu-rep profiles/r02_lmm_tc_k4_ncu_full_summary.json "note"

Reads the report with `ncu -i ... --page raw --csv` (first profiled kernel) and keeps the metrics the
DESIGN.md tables quote; with --sass also the executed-instruction mix from the source page."""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.per_second', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg.per_second', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__cluster_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum']


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    rep, out = args[0], args[1]
    note = args[2] if len(args) > 2 else ''
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, check=True).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    res = {'source': rep + ('; ' + note if note else ''), 'metrics': {}}
    for k in KEEP:
        if k in d and d[k][0] != '':
            res['metrics'][k] = {'value': d[k][0], 'unit': d[k][1]}
    if '--sass' in sys.argv:
        src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout.decode()
        rows = list(csv.reader(io.StringIO(src)))
        h = rows[1]
        ix = {x: i for i, x in enumerate(h)}
        ops = collections.Counter()
        tot = 0
        for r in rows[2:]:
            if len(r) != len(h):
                continue
            m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', r[ix['Source']])
            n = int(r[ix['Instructions Executed']] or 0)
            ops[m.group(2) if m else '?'] += n
            tot += n
        res['executed_instruction_mix_pct'] = {k: round(100.0 * v / tot, 2) for k, v in ops.most_common(16)}
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res['metrics'], indent=1)[:1500])


if __name__ == '__main__':
    main()
