#!/usr/bin/env python
"""In-kernel phase timers of the large-swarm tier (debug build, -DKB_PROFILE -> build/lib_prof.so):
   here (CPU box):  python tools/prof_swarm.py build
   GPU box:         python tools/prof_swarm.py run [side] [envs] [steps]
Prints cycles per phase of one env-step (thread 0's clock between block-wide marks), mean over envs."""
import ctypes as C
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "gym_kilobots_b200", "csrc")
PROF = os.path.join(ROOT, "build", "lib_prof.so")
NAMES = ["sense/light", "collide", "touching list", "body lists (CSR)", "island DFS (thread 0)", "rows+integrate+init",
         "warm start + velocity", "store + integrate", "position", "transform + sleep", "sync fixtures", "find new contacts",
         "toi", "gather + store", "(of find new contacts: pair hash + grid build)", "(of find new contacts: thread 0 waiting at the scans of the grid pass)"]

if sys.argv[1] == "build":
    os.makedirs(os.path.dirname(PROF), exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
                           "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-DKB_PROFILE", "-shared", "-o", PROF,
                           os.path.join(CSRC, "kb_b200.cu")])
    sys.exit(0)

import numpy as np  # noqa: E402
shutil.copy(PROF, os.path.join(CSRC, "libkb_b200.so"))   # GPU box scratch copy only
from gym_kilobots_b200 import _native as native  # noqa: E402
from gym_kilobots_b200 import scenarios as SC  # noqa: E402

side = int(sys.argv[2]) if len(sys.argv) > 2 else 32
E = int(sys.argv[3]) if len(sys.argv) > 3 else 148
T = int(sys.argv[4]) if len(sys.argv) > 4 else 8
sc = SC.c4_swarm(E, side=side)
nb = native.NativeBatch(sc.scenes, sc.num_envs, sc.env_scene, sc.max_contacts)
nb.reset(sc.body_pose, sc.light_state)
lib = native._lib
lib.kb_get_profile.argtypes = [C.c_void_p, C.c_void_p]
acts = np.zeros((T, E, 2))
cprev = nb.counters().astype(np.float64)
for t in range(T):
    nb.step(acts[t])
    if t not in (0, T // 2, T - 1):
        cprev = nb.counters().astype(np.float64)
    else:
        prof = np.zeros((E, 16), np.uint64)
        lib.kb_get_profile(nb.h, prof.ctypes.data_as(C.c_void_p))
        p = prof.astype(np.float64)
        tot = p[:, :14].sum(1)
        cnow = nb.counters().astype(np.float64)
        c = cnow - cprev
        cprev = cnow
        print("env-step %d: %.2f Mcycles per env-step (mean; max %.2f); touching/substep %.0f, levels/substep %.0f, position sweeps x islands/substep %.1f, islands/substep %.1f, pair tests/substep %.0f" % (
            t, tot.mean() / 1e6, tot.max() / 1e6, c[:, 2].sum() / c[:, 0].sum(), c[:, 3].sum() / c[:, 0].sum(), c[:, 4].sum() / c[:, 0].sum(), c[:, 7].sum() / c[:, 0].sum(), c[:, 6].sum() / c[:, 0].sum()))
        for i, n in enumerate(NAMES):
            if p[:, i].mean() > 0:
                print("   %-28s %9.0f kcycles  %5.1f %%" % (n, p[:, i].mean() / 1e3, 100 * p[:, i].mean() / tot.mean()))
