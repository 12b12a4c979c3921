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
