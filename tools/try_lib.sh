#!/bin/bash
# usage: tools/try_lib.sh <lib.so> [bench args...]
# GPU box only: swaps a variant build of the product library into the box's scratch copy and prints the bench line.
lib=$1; shift
cp "$lib" gym_kilobots_b200/csrc/libkb_b200.so
python bench.py --cpu-envs 8 "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('%.4f ms  value %.3e  e2e %.3e  %s' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['roofline']['launch']))
    elif 'rror' in l: print(l.rstrip())
"
