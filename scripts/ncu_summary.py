#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key metrics, stall reasons, hot source lines."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2 + which]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
for k in keys:
    if k in hdr:
        print("%-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
print("-- stalls per issue")
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
        try:
            v = float(r[i])
        except ValueError:
            continue
        if v > 0.1:
            print("   %-40s %.2f" % (h.split('issue_stalled_')[1].split('_per_issue')[0], v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", str(which),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = None
agg = collections.defaultdict(lambda: [0, 0])
tot = [0, 0]
for row in rows:
    if row and row[0] == 'Line No':
        h = row
        continue
    if h is None or len(row) < len(h):
        continue
    try:
        s, n = int(row[h.index('# Samples')]), int(row[h.index('Instructions Executed')])
    except ValueError:
        continue
    agg[row[0]][0] += s
    agg[row[0]][1] += n
    tot[0] += s
    tot[1] += n
print("-- hot source lines (line: samples%, instr%)  ['' = inlined headers]")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print("   %-8s %5.1f%% %5.1f%%" % (k, 100.0 * v[0] / max(1, tot[0]), 100.0 * v[1] / max(1, tot[1])))
