"""Round-2 GPU parity cases (CUDA path through the C-ABI vs. the CPU oracle, bit-exact unless stated):
BASELINE.json sizes for C3 / C5, light constants held per scene, the every-pair contact capacity on stacked
spawns, Body.set_pose on ONE body, status words after reset, and the product's episode-statistics reduction."""
import numpy as np
import pytest

from gym_kilobots_b200 import _abi as abi
from gym_kilobots_b200 import scenarios as SC
from gym_kilobots_b200 import scene as S

from parity_util import assert_same_obs, assert_same_state, run_parity

pytestmark = pytest.mark.gpu


def test_c3_baseline_size_all_shapes(oracle, native):
    """C3 as BASELINE.json states it: 50 kilobots (32 lanes per env) + L-form / triangle / circle object, 258 envs."""
    sc = SC.c3_shapes(258)
    assert sc.scenes[0].num_kilobots == 50 and set(sc.env_scene) == {0, 1, 2}
    run_parity(oracle, native, sc, steps=12, check_every=3)


def test_c5_full_size_vs_oracle_on_a_strided_sample(oracle, native):
    """C5 at BASELINE.json's full size (2^20 envs on one GPU), 6 env-steps; the oracle steps every 256th env
    (4096 envs: the counter-based spawn makes a slice of the batch the batch of the slice) and the observations of
    those envs must agree bit for bit."""
    E, stride, steps = 1 << 20, 256, 6
    sc = SC.c5_small(E)
    sub = SC.Scenario("C5-sample", sc.scenes, None, sc.body_pose[::stride].copy(), sc.light_state[::stride].copy(),
                      sc.max_contacts)
    acts = SC.random_actions(sc, E, steps)
    nb = native.NativeBatch(sc.scenes, E, None, sc.max_contacts)
    ob = oracle.OracleBatch(sub.scenes, sub.num_envs, None, sub.max_contacts, threads=8)
    nb.reset(sc.body_pose, sc.light_state)
    ob.reset(sub.body_pose, sub.light_state)
    for t in range(steps):
        on = nb.step(acts[t])
        oo = ob.step(acts[t][::stride])
        for k in ("kilobots", "objects", "light"):
            assert np.array_equal(on[k][::stride], oo[k]), "C5 full size, step %d: %s differs" % (t, k)
        assert not on["status"].any(), "status flags at full C5 size"
    # size-independent property: the table chain (kilobots_env.py:46-51: left, bottom and right edge; the top is
    # open) keeps every kilobot inside x in [-1, 1], y >= -0.75
    assert np.all(np.abs(on["kilobots"][..., 0]) <= 1.0 + 1e-3) and np.all(on["kilobots"][..., 1] >= -0.75 - 1e-3)
    assert np.isfinite(on["kilobots"]).all() and np.isfinite(on["objects"]).all()
    nb.close()


def test_light_constants_are_per_scene(oracle, native):
    """Envs whose lights differ (radius; order of the components of a shuffled CompositeLight) run with THEIR
    light constants, not scene 0's (round-1 advisor finding)."""
    world = (1.0, 0.5)
    wb = np.array(world) / 2
    bounds = (-wb * 1.1, wb * 1.1)
    act = (np.array([-1, -1]) * .01, np.array([1, 1]) * .01)
    circ = lambda r: S.LightSpec(abi.KB_LIGHT_CIRCULAR, radius=r, bounds=bounds, action_bounds=act)
    mom = lambda r: S.LightSpec(abi.KB_LIGHT_MOMENTUM, radius=r, bounds=bounds, action_bounds=act, max_velocity=.01)
    bodies = [S.quad_body(.1, .1)] + [S.kilobot_body(abi.KB_KILOBOT_SIMPLE_PHOTOTAXIS) for _ in range(9)]
    mk = lambda lights: S.SceneSpec(bodies=bodies, num_objects=1, lights=lights, world_size=world)
    scenes = [mk([circ(.25), mom(.2)]), mk([mom(.2), circ(.25)]), mk([circ(.1), mom(.35)])]
    E = 24
    rng = np.random.default_rng(3)
    pose = np.zeros((E, 10, 3))
    pose[:, 0, :2] = rng.uniform(-.05, .05, size=(E, 2))
    ring = np.array([(np.cos(a), np.sin(a)) for a in np.linspace(0, 2 * np.pi, 9, endpoint=False)]) * .15
    pose[:, 1:, :2] = ring[None] * np.array([1.0, .8]) + rng.uniform(-.01, .01, size=(E, 9, 2))
    pose[:, 1:, 2] = rng.uniform(-np.pi, np.pi, size=(E, 9))
    env_scene = (np.arange(E) % 3).astype(np.int32)
    light = np.zeros((E, 6))
    for e in range(E):
        a, b = rng.uniform(-.2, .2, size=2), rng.uniform(-.2, .2, size=2)
        light[e] = list(a) + [0, 0] + list(b) if env_scene[e] == 1 else list(a) + list(b) + [0, 0]
    sc = SC.Scenario("per-scene-lights", scenes, env_scene, pose, light, max_contacts=96)
    ob, nb = run_parity(oracle, native, sc, steps=25)
    # the three scene kinds must actually behave differently from one another
    k = nb.bodies()[:, 1:, :2]
    assert not np.array_equal(k[0], k[1]) and not np.array_equal(k[0], k[2])


@pytest.mark.parametrize("n", [32, 60])
def test_stacked_spawn_with_every_pair_capacity(oracle, native, n):
    """The reference's spawn -- N(mean, 0.03^2) clipped, no rejection (yaml_kilobots_env.py:346-352) -- stacks
    kilobots on top of each other.  With max_contacts = -1 (what the E = 1 facade passes) no pair is dropped:
    status stays 0 and the CUDA path follows the oracle bit for bit through the explosion."""
    world = (2.0, 1.5)
    sc = S.SceneSpec(bodies=[S.quad_body(.15, .15)] + [S.kilobot_body(abi.KB_KILOBOT_PHOTOTAXIS) for _ in range(n)],
                     num_objects=1, lights=[S.LightSpec(abi.KB_LIGHT_CIRCULAR, radius=.2)], world_size=world)
    E = 6
    rng = np.random.default_rng(n)
    centre = rng.uniform(-.5, .5, size=(E, 2))
    pose = np.zeros((E, n + 1, 3))
    pose[:, 0, :2] = centre + np.array([.3, 0.0])
    pose[:, 1:, :2] = centre[:, None, :] + rng.normal(scale=.03, size=(E, n, 2))
    scen = SC.Scenario("stacked-%d" % n, [sc], None, pose, centre.copy(), max_contacts=-1)
    ob, nb = run_parity(oracle, native, scen, steps=8)
    assert not nb.get_status().any() and not ob.get_status().any()
    P = nb.P
    assert nb.C >= P * (P - 1) // 2


def test_default_capacity_overflow_raises_in_the_vec_env(native):
    """Same stacked spawn through KilobotsVecEnv with the throughput-default capacity: the overflow is an error,
    not a silent status bit (unless allow_status_flags)."""
    from gym_kilobots_b200.envs import KilobotsVecEnv
    n = 40
    sc = S.SceneSpec(bodies=[S.kilobot_body(abi.KB_KILOBOT_PHOTOTAXIS) for _ in range(n)], num_objects=0,
                     lights=[S.LightSpec(abi.KB_LIGHT_CIRCULAR, radius=.2)], world_size=(2.0, 1.5))
    rng = np.random.default_rng(0)
    pose = np.zeros((4, n, 3))
    pose[:, :, :2] = rng.normal(scale=.01, size=(4, n, 2))
    scen = SC.Scenario("overflow", [sc], None, pose, np.zeros((4, 2)), max_contacts=0)
    env = KilobotsVecEnv(scen)
    env.reset()
    with pytest.raises(native.KbStatusError):
        env.check_status()
    with pytest.raises(native.KbStatusError):
        env.step(np.zeros((4, 2)))
    env.close()


def test_set_pose_moves_one_body_only(oracle, native):
    """Body.set_pose (lib/body.py:67-69) is b2Body::SetTransform on ONE body: compound objects with a non-zero
    local centre that were not touched keep their sweep bit for bit (round-1 advisor finding)."""
    sc = SC.pushing_yard(6, light="circular")
    ob, nb = run_parity(oracle, native, sc, steps=5)
    before = nb.bodies().copy()
    b = before
    pose = np.stack([b[..., 8] / 25.0, b[..., 9] / 25.0, b[..., 2]], axis=-1).astype(np.float64)
    mask = np.zeros(pose.shape[:2], np.uint8)
    mask[:, 7] = 1
    pose[:, 7, 0] += 0.01
    ob.set_poses(pose, mask)
    nb.set_poses(pose, mask)
    assert_same_state(ob, nb, "after masked set_poses")
    after = nb.bodies()
    untouched = np.ones(pose.shape[1], bool)
    untouched[7] = False
    assert np.array_equal(after[:, untouched], before[:, untouched])
    acts = SC.random_actions(sc, sc.num_envs, 4, seed=9)
    for t in range(4):
        assert_same_obs(ob.step(acts[t]), nb.step(acts[t]), "after masked set_poses, step %d" % t)


def test_episode_stats_reduction_matches_the_per_env_statistics(native):
    """kb_reduce_episode_stats (the rank-local operand of the NCCL all-reduce) against a float64 numpy sum of
    kb_get_episode_stats; world size 1 here, tools/allreduce_check.py runs the same check on 2 GPUs."""
    from gym_kilobots_b200.envs import KilobotsVecEnv
    sc = SC.c1_single_env(300)
    task = S.TaskSpec(abi.KB_TASK_OBJECT_TO_TARGET, object=0, max_episode_steps=3, w_position=1.0, w_orientation=0.1,
                      step_penalty=0.01, success_bonus=1.0, position_tolerance=0.02, orientation_tolerance=0.2)
    env = KilobotsVecEnv(sc, task=task, targets=np.tile(np.array([[0.2, 0.1, 0.3]]), (300, 1)))
    env.reset()
    acts = SC.random_actions(sc, 300, 5)
    for t in range(5):
        env.step(acts[t])
    st = env.batch.episode_stats()
    red = env.all_reduce_episode_stats()
    assert red["envs"] == 300 and red["envs_with_status"] == 0
    for i, name in enumerate(abi.EPISODE_STAT_NAMES):
        assert np.isclose(red["sum_" + name], st[:, i].sum(), rtol=1e-12, atol=1e-12), name
    assert red["episodes_done"] == st[:, abi.EPISODE_STAT_NAMES.index("done_count")].sum() > 0
    # deterministic: the same launch twice gives the same bits
    a = env.batch.reduce_episode_stats().cpu().numpy()
    b = env.batch.reduce_episode_stats().cpu().numpy()
    assert np.array_equal(a, b)
    env.close()


@pytest.mark.parametrize("chunks", ["1", "3", "8"])
def test_pipelined_host_step_equals_the_device_step(native, chunks, monkeypatch):
    """kb_step_host splits a large batch into chunks whose copies overlap the next chunk's kernel (two side streams);
    the results must not depend on the chunking (ragged last chunk, ragged last block)."""
    from gym_kilobots_b200.envs import KilobotsVecEnv
    monkeypatch.setenv("KB_HOST_CHUNKS", chunks)
    sc = SC.c5_small(1003)
    a = KilobotsVecEnv(sc)
    monkeypatch.setenv("KB_HOST_CHUNKS", "1")
    b = KilobotsVecEnv(sc)
    a.reset()
    b.reset()
    acts = SC.random_actions(sc, sc.num_envs, 6)
    for t in range(6):
        oa, ra, da, ia = a.step(acts[t])
        ob_, rb, db, ib = b.step_device(acts[t])
        for k in ("kilobots", "objects", "light"):
            assert np.array_equal(oa[k], ob_[k].cpu().numpy()), (chunks, t, k)
        assert np.array_equal(ra, rb.cpu().numpy()) and np.array_equal(ia["status"], ib["status"].cpu().numpy())
    assert np.array_equal(a.batch.bodies(), b.batch.bodies())
    a.close()
    b.close()
