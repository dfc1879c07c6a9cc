#!/usr/bin/env python
"""Hot spots of one kernel from an .ncu-rep with source counters: stall samples per SASS instruction (top N) and per
100-instruction bucket, plus the headline counters.  usage: tools/ncu_hot.py report.ncu-rep [topN]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr, units, val = raw[0], raw[1], raw[2]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
for i, h in enumerate(hdr):
    if h in KEYS or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(val[i] or 0) > 0.15):
        print(f'{h:90s} {val[i]:>16s} {units[i]}')
src = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout.splitlines()))
h2 = src[1]; data = src[2:]
iS = h2.index('Warp Stall Sampling (All Samples)'); isrc = h2.index('Source')
tot = sum(int(r[iS]) for r in data)
print('samples', tot, 'instructions', len(data))
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:topn]
for i in sorted(top):
    print(f'{i:5d} {int(data[i][iS]):6d} {100.0 * int(data[i][iS]) / tot:5.1f}%  {data[i][isrc][:100]}')
b = collections.Counter()
for i, r in enumerate(data):
    b[i // 100] += int(r[iS])
print('per 100 instr:', [(k, round(100.0 * v / tot, 1)) for k, v in sorted(b.items())])
