#!/usr/bin/env python
"""Key raw metrics + per-function instruction shares of an ncu report.
usage: python tools/ncu_summary.py report.ncu-rep [path/to/kb_step.cuh for function mapping]"""
import csv, io, os, re, subprocess, sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sass__inst_executed_global_loads',
        'sass__inst_executed_global_stores', 'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'sm__cycles_active.avg', 'sm__cycles_active.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.sum.per_cycle_elapsed']


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for k in KEYS:
        for i, h in enumerate(hdr):
            if h == k:
                print("%-70s %-12s %s" % (k, units[i], vals[i]))
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and 'per_issue_active' in h:
            print("  stall %-30s %s" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), vals[i]))
    srcfile = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gym_kilobots_b200", "csrc", "kb_step.cuh")
    src = open(srcfile).read().split('\n')
    funcs = []
    for i, l in enumerate(src, 1):
        m = re.match(r'\s*__device__ (?:__forceinline__ |__noinline__ )?(?:static )?[\w:<>\*& ]+?\b(\w+)\(', l)
        if m and not l.strip().endswith(';'):
            funcs.append((i, m.group(1)))

    def fn(line):
        name = '?'
        for i, n in funcs:
            if i <= line:
                name = n
            else:
                break
        return name
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    agg = defaultdict(lambda: [0, 0, 0])
    cur = None
    h = None
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == 'File Path':
            cur = os.path.basename(r[1]); h = None; continue
        if r[0] == 'Line No':
            h = {x: i for i, x in enumerate(r)}; continue
        if h is None or cur is None or r[0] == '':
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        key = (cur, fn(line) if cur == 'kb_step.cuh' else '')
        def num(x):
            try:
                return int(float(x))
            except ValueError:
                return 0
        agg[key][0] += num(r[h['Instructions Executed']])
        agg[key][1] += num(r[h['Thread Instructions Executed']])
        agg[key][2] += num(r[h['# Samples']])
    tot = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[2] for v in agg.values()) or 1
    print("--- per function (inlined code is attributed to the function whose source line it came from)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:32]:
        print("%-22s %-24s inst %5.2f%% smp %5.2f%% thr/inst %4.1f" % (k[0], k[1], 100 * v[0] / tot, 100 * v[2] / ts, v[1] / max(v[0], 1)))


if __name__ == "__main__":
    main()
