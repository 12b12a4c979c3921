#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo and --import-source on).

usage: python tools/ncu_lines.py report.ncu-rep [top_n]
Aggregates 'Instructions Executed', thread instructions and stall samples per (file, line) from
`ncu --page source --csv --print-source cuda,sass`.
"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file = None
    hdr = None
    per_line = defaultdict(lambda: [0, 0, 0, ""])   # inst, thread inst, samples, text
    per_file = defaultdict(lambda: [0, 0])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = os.path.basename(r[1])
            hdr = None
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            # two 'Source' columns: first is the CUDA source text
            continue
        if hdr is None or cur_file is None or r[0] in ("Function Name", "Kernel Name", "Address"):
            continue
        if r[0] == "":
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        def num(x):
            try:
                return int(float(x))
            except ValueError:
                return 0
        inst = num(r[hdr["Instructions Executed"]])
        tinst = num(r[hdr["Thread Instructions Executed"]])
        smp = num(r[hdr["# Samples"]])
        k = (cur_file, line)
        per_line[k][0] += inst
        per_line[k][1] += tinst
        per_line[k][2] += smp
        per_line[k][3] = r[1].strip()[:110]
        per_file[cur_file][0] += inst
        per_file[cur_file][1] += smp
    tot = sum(v[0] for v in per_line.values()) or 1
    tots = sum(v[2] for v in per_line.values()) or 1
    print("total warp-inst %d samples %d" % (tot, tots))
    for f, v in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        print("  %-22s inst %6.2f%%  samples %6.2f%%" % (f, 100.0 * v[0] / tot, 100.0 * v[1] / tots))
    print("--- top lines by instructions")
    for (f, l), v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-14s:%5d inst %5.2f%% smp %5.2f%% thr/inst %5.1f  %s" % (
            f, l, 100.0 * v[0] / tot, 100.0 * v[2] / tots, v[1] / max(v[0], 1), v[3]))
    print("--- top lines by stall samples")
    for (f, l), v in sorted(per_line.items(), key=lambda kv: -kv[1][2])[:top // 2]:
        print("%-14s:%5d inst %5.2f%% smp %5.2f%% thr/inst %5.1f  %s" % (
            f, l, 100.0 * v[0] / tot, 100.0 * v[2] / tots, v[1] / max(v[0], 1), v[3]))


if __name__ == "__main__":
    main()
