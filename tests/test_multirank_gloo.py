"""world_size-2 gloo test of the N>1 path: envs shard across ranks with no data-path collective, per-env
results are independent of the partitioning, and the only collective is the all-reduce of episode
statistics / timings (what bench.py does over NCCL).  The per-rank simulation is the CPU oracle here."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gym_kilobots_b200 import scenarios as SC
    from oracle import kbo
    E = 6
    sc = SC.c2_quad_assembly(E, seed=0, env_offset=rank * E, degenerate=False)
    ob = kbo.OracleBatch(sc.scenes, E, None, sc.max_contacts)
    ob.reset(sc.body_pose, sc.light_state)
    acts = SC.random_actions(sc, world * E, 4, seed=7)[:, rank * E:(rank + 1) * E]
    for a in acts:
        out = ob.step(a)
    stats = torch.tensor([float((out["status"] != 0).sum()), float(E * len(acts)), float(out["kilobots"].sum())],
                         dtype=torch.float64)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ret.put((stats.numpy().copy(), float(t[0])))
    gathered = [torch.zeros(E, 15, 3) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(out["kilobots"]))
    if rank == 0:
        ret.put(torch.cat(gathered).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_process():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    stats, tmax = q.get(timeout=120)
    kb = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from gym_kilobots_b200 import scenarios as SC
    from oracle import kbo
    sc = SC.c2_quad_assembly(12, seed=0, degenerate=False)
    ob = kbo.OracleBatch(sc.scenes, 12, None, sc.max_contacts)
    ob.reset(sc.body_pose, sc.light_state)
    for a in SC.random_actions(sc, 12, 4, seed=7):
        out = ob.step(a)
    assert np.array_equal(kb, out["kilobots"])          # per-env results independent of the sharding
    assert stats[0] == 0 and stats[1] == 12 * 4 and tmax == 2.0
    assert abs(stats[2] - float(out["kilobots"][:6].sum()) - float(out["kilobots"][6:].sum())) < 1e-3


class _FakeBatch:
    """Stands in for NativeBatch on the CPU box: `reduce_episode_stats` returns this rank's KB_REDUCED_STATS sums."""

    def __init__(self, sums):
        self._sums = torch.tensor(sums, dtype=torch.float64)

    def reduce_episode_stats(self):
        return self._sums.clone()


def _stats_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gym_kilobots_b200 import _abi as abi
    from gym_kilobots_b200.envs.vec_env import KilobotsVecEnv
    # the PRODUCT's collective (KilobotsVecEnv.all_reduce_episode_stats) with the device reduction stubbed out:
    # [envs, return, length, position_error, orientation_error, success, done_count, envs_with_status]
    env = object.__new__(KilobotsVecEnv)
    env.batch = _FakeBatch([100.0 + rank, 1.5 * (rank + 1), 40.0, 2.0, 0.5, 3.0 * rank, 7.0, float(rank)])
    out = env.all_reduce_episode_stats()
    assert len(abi.REDUCED_STAT_NAMES) == abi.KB_REDUCED_STATS
    if rank == 0:
        ret.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_product_all_reduce_of_episode_stats_over_gloo():
    """The host side of "NCCL used only to all-reduce episode statistics" on two ranks (gloo on the CPU box; the same
    method runs over NCCL on the GPU box: tests/test_multigpu.py, tools/allreduce_check.py)."""
    world, port = 2, 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_stats_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out["envs"] == 201 and out["envs_with_status"] == 1 and out["episodes_done"] == 14.0
    assert out["sum_return"] == 4.5 and out["sum_success"] == 3.0
    assert abs(out["mean_length"] - 80.0 / 201) < 1e-12 and abs(out["mean_return"] - 4.5 / 201) < 1e-12
