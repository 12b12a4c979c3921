#!/usr/bin/env python
"""Steady-state rollout timing on one GPU: episodes of EP env-steps with staggered phases (every step the envs with
(env + t) % EP == 0 are rebuilt by a masked kb_reset), split into the reset launch and the step launch.
usage: python tools/steady.py <workload> <envs> <steps> [EP]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from gym_kilobots_b200 import scenarios as SC  # noqa: E402
from gym_kilobots_b200.envs import KilobotsVecEnv  # noqa: E402

name, E, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
EP = int(sys.argv[4]) if len(sys.argv) > 4 else 50
sc = bench.build_scenario(name, E)
env = KilobotsVecEnv(sc, allow_status_flags=True)
env.reset()
dev = env.batch.device
acts = torch.as_tensor(SC.random_actions(sc, E, 64), dtype=torch.float64, device=dev)
pose = torch.as_tensor(sc.body_pose, dtype=torch.float64, device=dev)
light = torch.as_tensor(sc.light_state, dtype=torch.float64, device=dev)
ids = torch.arange(E, device=dev)
masks = [((ids + t) % EP == 0).to(torch.uint8) for t in range(EP)]
for t in range(2 * EP):
    env.batch.reset(pose, light, None, masks[t % EP])
    env.step_device(acts[t % 64])
torch.cuda.synchronize()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(T)]
for t in range(T):
    ev[t][0].record()
    env.batch.reset(pose, light, None, masks[t % EP])
    ev[t][1].record()
    env.step_device(acts[t % 64])
    ev[t][2].record()
torch.cuda.synchronize()
r = np.array([e[0].elapsed_time(e[1]) for e in ev])
s = np.array([e[1].elapsed_time(e[2]) for e in ev])
N = env.batch.N
print("%s E=%d EP=%d: reset %.3f ms, step %.3f ms (min %.3f max %.3f), total %.3f ms -> %.3e kilobot-steps/s; flagged %d" % (
    name, E, EP, r.mean(), s.mean(), s.min(), s.max(), (r + s).mean(),
    E * N / ((r + s).mean() * 1e-3), int((env.batch.get_status() != 0).sum())))
