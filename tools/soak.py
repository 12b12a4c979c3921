#!/usr/bin/env python
"""Long-horizon soak on the GPU: steps a full-size batch for many env-steps and reports the per-env status bits
(contact / solver capacity overflow, non-finite poses) and how close the contact counts come to the capacities.
usage: python tools/soak.py <c2|c2p|c3|c4|c5|c1> <envs> <env-steps>      (one JSON line)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gym_kilobots_b200 import scenarios as SC  # noqa: E402
from gym_kilobots_b200.envs import KilobotsVecEnv  # noqa: E402

name, E, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
sc = bench.build_scenario(name, E)
env = KilobotsVecEnv(sc, allow_status_flags=True)
env.reset()
acts = SC.random_actions(sc, E, 50)
import torch
a = torch.as_tensor(acts, dtype=torch.float64, device="cuda")
flags = np.zeros(E, np.int64)
maxc = 0
maxt = 0
probe = max(1, T // 20)
for t in range(T):
    _, _, _, info = env.step_device(a[t % 50] if name != "c4" else None)
    if t % probe == probe - 1 or t == T - 1:
        flags |= env.batch.get_status()
        if E * env.batch.C <= 1 << 26:
            pr, ct = env.batch.contacts()
            maxc = max(maxc, int(ct.max()))
            maxt = max(maxt, int(pr[:, :, 2].sum(1).max()))
kmax = 3 * env.batch.B + (32 if env.batch.B > 62 else 9)
print(json.dumps({"workload": name, "envs": E, "env_steps": T, "world_steps": 10 * T,
                  "envs_with_contact_overflow": int((flags & 1 != 0).sum()),
                  "envs_with_solver_overflow": int((flags & 4 != 0).sum()),
                  "envs_with_nonfinite_pose": int((flags & 2 != 0).sum()),
                  "max_persistent_contacts_seen": maxc, "contact_capacity": env.batch.C,
                  "max_touching_contacts_seen": maxt, "solver_capacity": kmax}))
