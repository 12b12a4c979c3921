#!/usr/bin/env python
"""Long-horizon soak on the GPU: steps a full-size batch for many env-steps and reports the per-env status bits
(contact / solver capacity overflow, non-finite poses) and how close the contact counts come to the capacities.
usage: python tools/soak.py <c2|c2p|c3|c5|c1> <envs> <env-steps>"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gym_kilobots_b200 import scenarios as SC  # noqa: E402
from gym_kilobots_b200.envs import KilobotsVecEnv  # noqa: E402

name, E, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
sc = bench.build_scenario(name, E)
env = KilobotsVecEnv(sc)
env.reset()
acts = SC.random_actions(sc, E, 50)
import torch
a = torch.as_tensor(acts, dtype=torch.float64, device="cuda")
flags = np.zeros(E, np.int64)
maxc = 0
maxt = 0
for t in range(T):
    _, _, _, info = env.step_device(a[t % 50])
    if t % 10 == 9 or t == T - 1:
        flags |= info["status"].cpu().numpy()
        pr, ct = env.batch.contacts()
        maxc = max(maxc, int(ct.max()))
        maxt = max(maxt, int(pr[:, :, 2].sum(1).max()))
print("%s: %d envs x %d env-steps: envs with contact-overflow %d, solver-overflow %d, non-finite %d; max persistent "
      "contacts %d (capacity %d), max touching %d (solver capacity 3B+9 = %d)" % (
          name, E, T, int((flags & 1 != 0).sum()), int((flags & 4 != 0).sum()), int((flags & 2 != 0).sum()), maxc,
          env.batch.C, maxt, 3 * env.batch.B + 9))
