#!/usr/bin/env python
"""Text summary of an ncu report for profiles/: python tools/summarize_ncu.py rep.ncu-rep [launch_index] > profiles/x.txt
Also: python tools/summarize_ncu.py --launches launches.csv  (per-kernel totals and shares of a launch list)."""
import csv
import subprocess
import sys
import collections

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = row['Kernel Name'].split('(')[0]
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        ms = v / 1e6 if u.startswith('n') else (v / 1e3 if u.startswith('u') else v)
        tot[k] += ms
        cnt[k] += 1
    T = sum(tot.values())
    print("kernel, launches, total_ms, share   (ncu --metrics gpu__time_duration.sum: cold-cache, serialised -> compare SHARES)")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print("%-60s %6d %12.3f %7.4f" % (k[:60], cnt[k], v, v / T))


def report(path, which):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    r = rows[2 + which]
    print("report:", path, " launch", r[hdr.index('ID')], r[hdr.index('Kernel Name')])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("  %-80s %18s %s" % (k, r[i], units[i]))
    st = []
    for i, h in enumerate(hdr):
        if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
            try:
                st.append((float(r[i]), h))
            except ValueError:
                pass
    print("  top stall reasons (warps per issue-active cycle):")
    for v, h in sorted(st)[::-1][:7]:
        print("    %8.3f %s" % (v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))


if __name__ == '__main__':
    if sys.argv[1] == '--launches':
        launches(sys.argv[2])
    else:
        report(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
