#!/usr/bin/env python
"""Phase attribution of an `ncu --set full --import-source on` capture of kb_step_kernel<LPE>.

usage: python tools/ncu_phases.py report.ncu-rep libkb_b200.so [LPE]
Every SASS instruction is mapped (nvdisasm -gi inline chains) to the phase of Sim::solve / Sim::worldStep /
the kernel body whose source line is the outermost frame, then executed instructions and stall samples are
summed per phase.  The library must be the build the report was captured from (same instruction count).
"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, lib = sys.argv[1], sys.argv[2]
lpe = sys.argv[3] if len(sys.argv) > 3 else "8"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "gym_kilobots_b200", "csrc", "kb_step.cuh")).read().split("\n")

def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]:
            return i + 1
    raise SystemExit("pattern not found: " + pat)

solve0 = find("__device__ __forceinline__ void solve() {")
kt = {i: find("KB_T(%d);" % i, solve0 if 2 <= i <= 10 else 0) for i in range(12)}
ws0 = find("__device__ __forceinline__ void worldStep() {")
names = {2: "touchlist", 3: "serial DFS/levels", 4: "integrate v + init", 5: "warm start + velocity", 6: "store + integrate x",
         7: "position", 8: "sync transforms + sleep", 9: "synchronizeFixtures", 10: "findNewContacts"}

def phase_of(chain):
    # chain: list of (file, line), innermost first; walk from the outermost frame inwards
    body = None
    for f, l in reversed(chain):
        if f == "kb_b200.cu":
            body = "kernel body: line %d" % l
            continue
        if f == "kb_step.cuh":
            if solve0 <= l <= kt[10]:
                for i in range(2, 11):
                    if l <= kt[i]:
                        return names[i]
            if ws0 <= l <= kt[11] + 1:
                if l <= kt[0]: return "worldStep misc"
                if l <= kt[1] - 2: return "collide"
                if l <= kt[1] + 1: continue   # the solve() call: look deeper
                return "TOI (call site)"
            if body:
                return body
        if f == "kb_toi.cuh": return "TOI"
    return body or ("other:" + (chain[-1][0] if chain else "?"))

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cub], capture_output=True, text=True).stdout.split("\n")
sec = ".text._ZN2kb14kb_step_kernelILi%sEEEvNS_10KernelArgsE:" % lpe
i0 = dis.index(sec)
insts = []  # (addr, chain)
chain = []
pend = []
fre = re.compile(r'File ".*?/([^/"]+)", line (\d+)')
for ln in dis[i0 + 1:]:
    if ln.startswith("//-----") or ln.startswith("\t.section"):
        break
    if "//## File" in ln:
        m = fre.findall(ln)
        pend.append([(f, int(l)) for f, l in m])
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(\S.*);", ln)
    if m:
        if pend:
            # first pending line is innermost "X inlined at Y"; the following lines continue the chain outward
            chain = []
            for p in pend:
                for fl in p:
                    if not chain or chain[-1] != fl:
                        chain.append(fl)
            pend = []
        insts.append((int(m.group(1), 16), chain, m.group(2)))
if rep == "-":   # static mode: SASS instruction count per phase, no report needed
    cnt = defaultdict(int)
    for addr, ch, txt in insts:
        cnt[phase_of(ch)] += 1
    for p_, c_ in sorted(cnt.items(), key=lambda kv: -kv[1]):
        print("%-28s %8d" % (p_, c_))
    print("%-28s %8d" % ("total", len(insts)))
    sys.exit(0)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr and r and r[0].startswith("0x"):
        data.append((int(r[hdr["# Samples"]] or 0), int(r[hdr["Instructions Executed"]] or 0),
                     int(r[hdr["Thread Instructions Executed"]] or 0)))
if len(data) != len(insts):
    print("warning: report has %d instructions, library %d -- different builds?" % (len(data), len(insts)))
agg = defaultdict(lambda: [0, 0, 0, 0])
for (smp, ie, te), (addr, ch, txt) in zip(data, insts):
    p = phase_of(ch)
    a = agg[p]
    a[0] += smp; a[1] += ie; a[2] += te; a[3] += 1
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print("%-28s %8s %8s %9s %8s" % ("phase", "samples%", "inst%", "thr/inst", "SASS"))
for p, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %8.2f %8.2f %9.1f %8d" % (p, 100.0 * a[0] / ts, 100.0 * a[1] / ti, a[2] / max(a[1], 1), a[3]))
if os.environ.get("KB_DUMP_PHASE"):
    want = os.environ["KB_DUMP_PHASE"]
    for (smp, ie, te), (addr, ch, txt) in zip(data, insts):
        if phase_of(ch) == want:
            inner = ch[0] if ch else ("?", 0)
            print("%05x %9d %5.1f %6d  %-14s:%-5d %s" % (addr, ie, te / max(ie, 1), smp, inner[0], inner[1], txt[:70]))
