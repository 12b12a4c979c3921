#!/usr/bin/env python
"""Product-level collective check on R GPUs (torchrun --nproc-per-node R tools/allreduce_check.py):
(1) KilobotsVecEnv.all_reduce_episode_stats over NCCL equals the float64 sum of every rank's per-env statistics;
(2) determinism across rank counts: every rank simulates its slice of the SAME global env ids, and the hashes of the
    per-env observations, all-gathered, equal those of a single-GPU run of all global envs (computed on rank 0).
Prints one JSON line on rank 0; exit code 1 on mismatch."""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gym_kilobots_b200 import _abi as abi  # noqa: E402
from gym_kilobots_b200 import scenarios as SC  # noqa: E402
from gym_kilobots_b200 import scene as S  # noqa: E402
from gym_kilobots_b200.envs import KilobotsVecEnv  # noqa: E402


def run(sc, actions, device):
    task = S.TaskSpec(abi.KB_TASK_SWARM_TO_TARGET, max_episode_steps=4, w_position=1.0, step_penalty=0.01,
                      position_tolerance=0.01)
    env = KilobotsVecEnv(sc, device=device, task=task, targets=np.zeros((sc.num_envs, 3)))
    env.reset()
    for a in actions:
        obs, rew, done, info = env.step(a)
    h = [hashlib.sha1(np.ascontiguousarray(obs["kilobots"][e]).tobytes() + np.ascontiguousarray(obs["objects"][e]).tobytes()
                      + np.ascontiguousarray(obs["light"][e]).tobytes()).hexdigest() for e in range(sc.num_envs)]
    return env, h


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E_total, steps = 4096, 6
    E = E_total // world
    acts_all = SC.random_actions(2, E_total, steps, seed=11)          # [steps, E_total, 2]: indexed by GLOBAL env id
    sc = SC.c5_small(E, env_offset=rank * E)
    env, hashes = run(sc, acts_all[:, rank * E:(rank + 1) * E], local)
    red = env.all_reduce_episode_stats()
    local_sums = env.batch.episode_stats().sum(0)
    gathered = [None] * world
    sums = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, hashes)
        dist.all_gather_object(sums, local_sums)
    else:
        gathered, sums = [hashes], [local_sums]
    ok = True
    out = None
    if rank == 0:
        total = np.sum(np.stack(sums), axis=0)
        ok_stats = red["envs"] == E_total and all(
            np.isclose(red["sum_" + n], total[i], rtol=1e-12, atol=1e-12) for i, n in enumerate(abi.EPISODE_STAT_NAMES))
        env.close()
        ref_env, ref_hashes = run(SC.c5_small(E_total), acts_all, local)
        flat = [h for part in gathered for h in part]
        ok_det = flat == ref_hashes
        ok = ok_stats and ok_det
        out = {"world_size": world, "backend": dist.get_backend() if world > 1 else "none", "envs_total": E_total,
               "env_steps": steps, "all_reduce_matches_sum_of_ranks": bool(ok_stats),
               "per_env_obs_hashes_equal_single_gpu_run": bool(ok_det), "episode_stats": red}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
