#!/usr/bin/env python
"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the small text summaries committed
under profiles/.

    python tools/ncu_summary.py launches gpurun_out/X_launches.csv  > profiles/X_launches.txt
    python tools/ncu_summary.py full     gpurun_out/X.ncu-rep       > profiles/X_full.txt
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
    'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        name = row['Kernel Name'].split('(')[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print('# ncu --metrics gpu__time_duration.sum --clock-control none : per-launch device time, cold cache, serialised')
    print('# source: %s ; %d launches, %.1f us total' % (path, sum(v[0] for v in agg.values()), tot))
    print('%-72s %8s %14s %7s' % ('kernel', 'launches', 'total_us', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-72s %8d %14.1f %6.1f%%' % (k[:72], v[0], v[1], 100 * v[1] / tot))


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print('# ncu --set full --clock-control none : %s' % path)
    for r in rows[2:]:
        print('kernel: %s' % r[hdr.index('Kernel Name')])
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print('  %-78s %16s %s' % (m, r[i], units[i]))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
