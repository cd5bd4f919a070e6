#!/usr/bin/env python3
"""Summarise an .ncu-rep (or a launch-list csv) into the few numbers DESIGN.md / bench.py quote.
usage: ncu_summary.py report.ncu-rep   |   ncu_summary.py launches.csv"""
import collections, csv, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'launch__waves_per_multiprocessor', 'smsp__warps_eligible.avg.per_cycle_active']


def rep(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('kernel:', vals[hdr.index('Kernel Name')][:90])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'  {w:72s} {vals[i]:>18s} {units[i]}')
        for i, h in enumerate(hdr):
            if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                try:
                    v = float(vals[i])
                except ValueError:
                    continue
                if v > 0.4:
                    print(f'  {h:72s} {v:18.2f}')


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if 'Kernel Name' in r)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except (ValueError, IndexError):
            continue
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        agg.setdefault(r[ki].split('(')[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f'{k:72s} n={len(v):3d} total={sum(v) / 1e3:9.3f} ms avg={sum(v) / len(v) / 1e3:8.3f} ms share={sum(v) / tot * 100:5.1f}%')


if __name__ == '__main__':
    (launches if sys.argv[1].endswith('.csv') else rep)(sys.argv[1])
