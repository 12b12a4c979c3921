"""Task layer: on-device reward, termination, episode statistics and the flat
Yaml observation.  It is an extension -- the reference's hooks are abstract (kilobots_env.py:123-131) and its
in-tree envs return constants -- so the oracle's version is pinned here against an independent numpy
restatement of the formulas in include/kb_b200.h, and the CUDA path is compared with the oracle bit for bit.
"""
import math

import numpy as np
import pytest

from gym_kilobots_b200 import _abi as abi
from gym_kilobots_b200 import scenarios as SC
from gym_kilobots_b200.scene import TaskSpec


def _targets(sc, seed=3):
    rng = np.random.default_rng(seed)
    t = np.zeros((sc.num_envs, 3))
    t[:, 0] = rng.uniform(-0.8, 0.8, sc.num_envs)
    t[:, 1] = rng.uniform(-0.6, 0.6, sc.num_envs)
    t[:, 2] = rng.uniform(-7.0, 7.0, sc.num_envs)     # beyond +-pi: the angle error must wrap
    return t


def _numpy_error(bodies, M, task, target):
    """(distance m, |angle| rad) per env from the raw body state, float64, the header's formulas."""
    E = bodies.shape[0]
    d = np.zeros(E)
    a = np.zeros(E)
    for e in range(E):
        if task.mode == abi.KB_TASK_OBJECT_TO_TARGET:
            px = np.float64(bodies[e, task.object, 8]) / 25.0
            py = np.float64(bodies[e, task.object, 9]) / 25.0
            a[e] = abs(math.remainder(float(np.float64(bodies[e, task.object, 2])) - target[e, 2], 6.283185307179586))
        else:
            sx = sy = np.float64(0.0)
            for b in range(M, bodies.shape[1]):
                sx = sx + np.float64(bodies[e, b, 8]) / 25.0
                sy = sy + np.float64(bodies[e, b, 9]) / 25.0
            px, py = sx / np.float64(bodies.shape[1] - M), sy / np.float64(bodies.shape[1] - M)
        dx, dy = px - target[e, 0], py - target[e, 1]
        d[e] = np.sqrt(dx * dx + dy * dy)
    return d, a


TASKS = [
    TaskSpec(mode=abi.KB_TASK_OBJECT_TO_TARGET, object=0, w_position=1.0, w_orientation=0.25, step_penalty=0.001,
             success_bonus=2.0, position_tolerance=0.5, orientation_tolerance=1.0, max_episode_steps=7),
    TaskSpec(mode=abi.KB_TASK_SWARM_TO_TARGET, w_position=3.0, step_penalty=0.01, success_bonus=1.0,
             position_tolerance=0.3, max_episode_steps=5),
]


@pytest.mark.parametrize("task", TASKS, ids=["object_to_target", "swarm_to_target"])
def test_oracle_task_layer_matches_numpy_restatement(oracle, task):
    sc = SC.c1_single_env(24, seed=5)
    ob = oracle.OracleBatch(sc.scenes, sc.num_envs, sc.env_scene, sc.max_contacts, threads=4)
    tg = _targets(sc)
    ob.set_task(task, tg)
    flat = ob.bind_flat_observation()
    ob.reset(sc.body_pose, sc.light_state)
    acts = SC.random_actions(sc, sc.num_envs, 12)
    ret = np.zeros(sc.num_envs)
    length = np.zeros(sc.num_envs)
    done_count = np.zeros(sc.num_envs)
    saw_done = saw_success = False
    for t in range(12):
        d0, a0 = _numpy_error(ob.bodies(), ob.M, task, tg)
        out = ob.step(acts[t])
        b = ob.bodies()
        d1, a1 = _numpy_error(b, ob.M, task, tg)
        success = (d1 <= task.position_tolerance) & (a1 <= task.orientation_tolerance)
        r = task.w_position * (d0 - d1)
        r = r + task.w_orientation * (a0 - a1)
        r = r - task.step_penalty
        r = np.where(success, r + task.success_bonus, r)
        length += 1
        done = success | (length >= task.max_episode_steps)
        ret += r
        done_count += done
        assert np.array_equal(out["reward"], r.astype(np.float32)), "reward, step %d" % t
        assert np.array_equal(out["done"].astype(bool), done), "done, step %d" % t
        st = ob.episode_stats()
        assert np.array_equal(st[:, 0], ret) and np.array_equal(st[:, 1], length)
        assert np.array_equal(st[:, 2], d1) and np.array_equal(st[:, 3], a1)
        assert np.array_equal(st[:, 4], success.astype(float)) and np.array_equal(st[:, 5], done_count)
        # flat observation = YamlKilobotsEnv.observation_space order
        N, M, L = ob.N, ob.M, ob.L
        assert np.array_equal(flat[:, :2 * N].reshape(-1, N, 2), out["kilobots"][:, :, :2])
        assert np.array_equal(flat[:, 2 * N:2 * N + L], out["light"].astype(np.float32))
        fo = flat[:, 2 * N + L:].reshape(-1, M, 4)
        assert np.array_equal(fo[:, :, :2], out["objects"][:, :, :2])
        assert np.array_equal(fo[:, :, 2], b[:, :M, 10]) and np.array_equal(fo[:, :, 3], b[:, :M, 11])
        saw_done |= bool(done.any())
        saw_success |= bool(success.any())
        if done.any():   # auto-reset of the finished envs only
            ob.reset(sc.body_pose, sc.light_state, mask=done.astype(np.uint8))
            ret[done] = 0.0
            length[done] = 0.0
    assert saw_done and saw_success, "the case must exercise termination and success"


def test_const_task_is_the_reference_default(oracle):
    """Without a task the hooks return the in-tree constants (yaml_kilobots_env.py:368-375, test envs :89-91)."""
    sc = SC.c1_single_env(4, seed=1)
    for s in sc.scenes:
        s.reward_const = 1.0
    ob = oracle.OracleBatch(sc.scenes, sc.num_envs, sc.env_scene, sc.max_contacts)
    ob.reset(sc.body_pose, sc.light_state)
    out = ob.step(SC.random_actions(sc, sc.num_envs, 1)[0])
    assert np.array_equal(out["reward"], np.ones(4, np.float32)) and not out["done"].any()


@pytest.mark.gpu
@pytest.mark.parametrize("task", TASKS, ids=["object_to_target", "swarm_to_target"])
def test_kernel_task_layer_matches_oracle(oracle, native, task):
    sc = SC.c2_quad_assembly(96, seed=2, degenerate=False)
    ob = oracle.OracleBatch(sc.scenes, sc.num_envs, sc.env_scene, sc.max_contacts, threads=8)
    nb = native.NativeBatch(sc.scenes, sc.num_envs, sc.env_scene, sc.max_contacts)
    tg = _targets(sc)
    if task.mode == abi.KB_TASK_OBJECT_TO_TARGET:
        task = TaskSpec(**{**task.__dict__, "object": 3})   # the movable quad of the assembly scene
        tg[:, :2] = sc.body_pose[:, 3, :2] + 0.02             # near its spawn: some envs succeed, most time out
    ob.set_task(task, tg)
    nb.set_task(task, tg)
    fo, fn = ob.bind_flat_observation(), nb.bind_flat_observation()
    ob.reset(sc.body_pose, sc.light_state)
    nb.reset(sc.body_pose, sc.light_state)
    acts = SC.random_actions(sc, sc.num_envs, 14)
    dones = 0
    for t in range(14):
        oo, on = ob.step(acts[t]), nb.step(acts[t])
        for k in ("kilobots", "objects", "light", "reward", "done", "status"):
            assert np.array_equal(oo[k], on[k]), "step %d: %s differs" % (t, k)
        assert np.array_equal(fo, fn.cpu().numpy()), "step %d: flat observation differs" % t
        assert np.array_equal(ob.episode_stats(), nb.episode_stats()), "step %d: episode statistics differ" % t
        d = oo["done"]
        dones += int(d.sum())
        if d.any():
            ob.reset(sc.body_pose, sc.light_state, mask=d)
            nb.reset(sc.body_pose, sc.light_state, mask=d)
    assert dones > 0
    assert np.array_equal(ob.bodies(), nb.bodies())


@pytest.mark.gpu
def test_vec_env_task_surface(native):
    """KilobotsVecEnv(task=..., flat_observation=True): device and host paths agree, episode stats are exposed."""
    from gym_kilobots_b200.envs import KilobotsVecEnv
    sc = SC.c1_single_env(16, seed=9)
    tg = _targets(sc)
    envs = [KilobotsVecEnv(sc, task=TASKS[0], targets=tg, flat_observation=True) for _ in range(2)]
    for e in envs:
        e.reset()
    acts = SC.random_actions(sc, sc.num_envs, 4)
    for t in range(4):
        od, rd, dd, _ = envs[0].step_device(acts[t])
        oh, rh, dh, _ = envs[1].step(acts[t])
        assert np.array_equal(od["flat"].cpu().numpy(), oh["flat"])
        assert np.array_equal(rd.cpu().numpy(), rh) and np.array_equal(dd.cpu().numpy().astype(bool), dh)
    st = envs[0].episode_stats()
    assert set(st) == set(abi.EPISODE_STAT_NAMES) and np.array_equal(st["length"], np.full(16, 4.0))
