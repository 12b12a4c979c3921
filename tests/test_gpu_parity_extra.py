"""More GPU parity cases (CUDA path through the C-ABI vs. the CPU oracle, bit-exact): every object shape with
object-object / object-table contacts, the other controllers and lights, direct control, masked resets,
Body.set_pose, checkpoint resume, every lane-group width, ragged batches, and -- at BASELINE.json's full C2
size -- size-independent properties (partition invariance, run-to-run determinism, the oracle on a sample)."""
import numpy as np
import pytest

from gym_kilobots_b200 import _abi as abi
from gym_kilobots_b200 import scenarios as SC

from parity_util import assert_same_obs, assert_same_state, run_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("light", ["momentum", "composite", "circular"])
def test_pushing_yard_simple_phototaxis(oracle, native, light):
    run_parity(oracle, native, SC.pushing_yard(16, light=light), steps=30)


def test_pushing_yard_linear_light(oracle, native):
    sc = SC.pushing_yard(8, light="linear")
    acts = SC.random_actions(sc, sc.num_envs, 20) * 300.0   # angles in [-3, 3] rad
    run_parity(oracle, native, sc, steps=20, actions=acts)


def test_pushing_yard_phototaxis_no_toi(oracle, native):
    run_parity(oracle, native, SC.pushing_yard(16, kilobot_kind=abi.KB_KILOBOT_PHOTOTAXIS, light="circular",
                                               enable_toi=False), steps=30)


def test_direct_control(oracle, native):
    sc = SC.direct_control(16)
    acts = SC.random_kilobot_actions(sc, sc.num_envs, 30)
    run_parity(oracle, native, sc, steps=30, actions=acts, mode=abi.KB_ACTION_KILOBOTS)


def test_action_none_freezes_the_light(oracle, native):
    sc = SC.c1_single_env(8)
    run_parity(oracle, native, sc, steps=10, actions=[None] * 10, mode=abi.KB_ACTION_NONE)


@pytest.mark.parametrize("lanes", [4, 8, 16, 32])
def test_every_lane_group_width(oracle, native, lanes, monkeypatch):
    """The lane-group width is a launch parameter, not part of the semantics."""
    monkeypatch.setenv("KB_LANES_PER_ENV", str(lanes))
    run_parity(oracle, native, SC.c2_quad_assembly(21, degenerate=False), steps=15)   # 21: ragged last block
    run_parity(oracle, native, SC.pushing_yard(5, light="momentum"), steps=15)


def test_masked_reset_set_pose_and_resume(oracle, native):
    sc = SC.c2_quad_assembly(24, degenerate=False)
    ob, nb = run_parity(oracle, native, sc, steps=6)
    # masked reset of every third env (auto-reset path)
    mask = (np.arange(sc.num_envs) % 3 == 0).astype(np.uint8)
    pose2 = SC.c2_quad_assembly(24, seed=5, degenerate=False)
    ob.reset(pose2.body_pose, pose2.light_state, mask=mask)
    nb.reset(pose2.body_pose, pose2.light_state, mask=mask)
    assert_same_state(ob, nb, "after masked reset")
    acts = SC.random_actions(sc, sc.num_envs, 12, seed=3)
    for t in range(4):
        assert_same_obs(ob.step(acts[t]), nb.step(acts[t]), "after masked reset, step %d" % t)
    # Body.set_pose on every body (lib/body.py:67-69)
    b = nb.bodies()
    pose = np.stack([b[..., 8] / 25.0, b[..., 9] / 25.0, b[..., 2]], axis=-1).astype(np.float64)
    pose[:, 4:, 0] += 0.002
    pose[:, 4:, 2] += 0.1
    ob.set_poses(pose)
    nb.set_poses(pose)
    assert_same_state(ob, nb, "after set_poses")
    for t in range(4, 8):
        assert_same_obs(ob.step(acts[t]), nb.step(acts[t]), "after set_poses, step %d" % t)
    assert_same_state(ob, nb, "after set_poses + steps")
    # checkpoint / bit-exact resume of the CUDA path
    blob = nb.get_state()
    ref = [nb.step(acts[t]) for t in range(8, 12)]
    nb.set_state(blob)
    for t in range(8, 12):
        assert_same_obs(ref[t - 8], nb.step(acts[t]), "resume, step %d" % t)


def test_full_size_c2_properties(oracle, native):
    """BASELINE.json configs[1] at full size: 4096 envs x 15 kilobots."""
    E, steps = 4096, 6
    sc = SC.c2_quad_assembly(E)
    acts = SC.random_actions(sc, E, steps)
    nb = native.NativeBatch(sc.scenes, E, sc.env_scene, sc.max_contacts)
    nb.reset(sc.body_pose, sc.light_state)
    outs = [nb.step(acts[t]) for t in range(steps)]
    assert not outs[-1]["status"].any()
    # run-to-run determinism
    nb2 = native.NativeBatch(sc.scenes, E, sc.env_scene, sc.max_contacts)
    nb2.reset(sc.body_pose, sc.light_state)
    for t in range(steps):
        o2 = nb2.step(acts[t])
    assert_same_obs(outs[-1], o2, "determinism")
    assert np.array_equal(nb.bodies(), nb2.bodies())
    # partition invariance: a slice of the batch is the batch of the slice (what rank sharding relies on)
    lo, n = 1000, 96
    sub = SC.c2_quad_assembly(n, env_offset=lo)
    nb3 = native.NativeBatch(sub.scenes, n, sub.env_scene, sub.max_contacts)
    nb3.reset(sub.body_pose, sub.light_state)
    for t in range(steps):
        o3 = nb3.step(acts[t][lo:lo + n])
    for k in ("kilobots", "objects", "light"):
        assert np.array_equal(o3[k], outs[-1][k][lo:lo + n]), k
    # and the same slice against the oracle
    ob = oracle.OracleBatch(sub.scenes, n, sub.env_scene, sub.max_contacts, threads=8)
    ob.reset(sub.body_pose, sub.light_state)
    for t in range(steps):
        oo = ob.step(acts[t][lo:lo + n])
    assert_same_obs(oo, o3, "slice vs oracle")
    # conservation-style sanity at full size: nothing leaves the table by more than a body radius + slop
    kb = outs[-1]["kilobots"]
    assert np.isfinite(kb).all()
    assert (np.abs(kb[..., 0]) < 1.0 + 0.02).all() and (kb[..., 1] > -0.75 - 0.02).all()


def test_vec_env_host_and_device_paths_agree(oracle, native):
    """KilobotsVecEnv.step (host numpy, packed single D2H copy) == step_device (CUDA tensors) == oracle."""
    import torch
    from gym_kilobots_b200.envs import KilobotsVecEnv
    sc = SC.c2_quad_assembly(37, degenerate=False)
    acts = SC.random_actions(sc, sc.num_envs, 5)
    ea, eb = KilobotsVecEnv(sc), KilobotsVecEnv(sc)
    ob = oracle.OracleBatch(sc.scenes, sc.num_envs, sc.env_scene, sc.max_contacts, threads=8)
    ea.reset(); eb.reset(); ob.reset(sc.body_pose, sc.light_state)
    for t in range(5):
        obs_h, rew_h, done_h, info_h = ea.step(acts[t])
        obs_d, rew_d, done_d, info_d = eb.step_device(torch.as_tensor(acts[t], device="cuda"))
        oo = ob.step(acts[t])
        for k in ("kilobots", "objects", "light"):
            assert np.array_equal(obs_h[k], obs_d[k].cpu().numpy()), (t, k)
            assert np.array_equal(obs_h[k], oo[k]), (t, k)
        assert np.array_equal(rew_h, rew_d.cpu().numpy()) and np.array_equal(rew_h, oo["reward"])
        assert not done_h.any() and not info_h["status"].any()
    h2d, d2h = ea.host_io_bytes()
    assert h2d == 37 * 2 * 8 and d2h == 37 * (15 * 12 + 4 * 12 + 2 * 8 + 4 + 1 + 4)


def _oracle_backed(cls, oracle):
    class Injected(cls):
        def _make_batch(self, spec):
            return oracle.OracleBatch(spec, 1)
    return Injected


def test_single_env_facades_on_the_gpu_match_the_oracle(oracle, native):
    """The E = 1 drop-in classes (gym_kilobots.envs.*) stepping on the B200 vs. the same classes on the oracle."""
    import yaml
    from gym_kilobots_b200.envs import (DirectControlKilobotsEnv, QuadAssemblyKilobotsEnv, YamlKilobotsEnv)
    from gym_kilobots_b200.lib import Quad, SimpleVelocityControlKilobot

    def same(a, b, what):
        for k in ("kilobots", "objects", "light"):
            assert np.array_equal(a[k], b[k]), (what, k)

    # QuadAssemblyKilobotsEnv (kilobots_test_envs.py:23-95)
    eg, eo = QuadAssemblyKilobotsEnv(seed=7), _oracle_backed(QuadAssemblyKilobotsEnv, oracle)(seed=7)
    same(eg.reset(), eo.reset(), "quad assembly reset")
    rng = np.random.default_rng(0)
    for t in range(6):
        a = rng.uniform(-.01, .01, size=2)
        og, rg, dg, ig = eg.step(a)
        oo, ro, do, io = eo.step(a)
        same(og, oo, "quad assembly step %d" % t)
        assert rg == ro == 1.0 and dg is False and ig is None
    assert np.array_equal(eg.kilobots[3].get_pose(), eo.kilobots[3].get_pose())

    # YamlKilobotsEnv (yaml_kilobots_env.py:101-369)
    text = """
!EvalEnv
width: 1.0
height: 1.0
resolution: 600
objects:
  - !ObjectConf {idx: 0, color: null, shape: l_shape, width: .15, height: .15, init: [.1, -.1, .3], symmetry: null}
  - !ObjectConf {idx: 1, color: null, shape: circle, width: .05, height: .05, init: [.2, .2, 0.], symmetry: null}
light: !LightConf {type: momentum, init: [0., 0.], radius: .2}
kilobots: !KilobotsConf {num: 7, mean: [0., 0.], std: .03}
"""
    conf = yaml.load(text, Loader=yaml.Loader)
    import random

    def seeded(cls):
        np.random.seed(3)
        random.seed(3)
        e = cls(configuration=conf)
        return e, e.reset()
    eg, og0 = seeded(YamlKilobotsEnv)
    eo, oo0 = seeded(_oracle_backed(YamlKilobotsEnv, oracle))
    same(og0, oo0, "yaml reset")
    for t in range(4):
        a = rng.uniform(-.01, .01, size=2)
        same(eg.step(a)[0], eo.step(a)[0], "yaml step %d" % t)

    # DirectControlKilobotsEnv (direct_control_kilobots_env.py:8-29)
    class Env(DirectControlKilobotsEnv):
        def _configure_environment(self):
            self._objects = [Quad(world=self.world, width=.1, height=.1, position=(.08, .0))]
            self._kilobots = [SimpleVelocityControlKilobot(self.world, position=(-.04 * i, .0), orientation=.0,
                                                           velocity=[.005, .0]) for i in range(3)]

        def get_reward(self, *a):
            return 0.

    eg, eo = Env(), _oracle_backed(Env, oracle)()
    same(eg.reset(), eo.reset(), "direct reset")
    for t in range(6):
        a = np.stack([rng.uniform(0, .012, size=3), rng.uniform(-1.8, 1.8, size=3)], axis=1)
        same(eg.step(a)[0], eo.step(a)[0], "direct step %d" % t)


def test_vec_env_from_yaml_envs(oracle, native):
    """KilobotsVecEnv.from_envs: reference-style YamlKilobotsEnv instances (random object / light / kilobot
    initialisation per env) vectorised into one CUDA batch == the oracle on the recorded scenario."""
    import yaml
    from gym_kilobots_b200.envs import KilobotsVecEnv, YamlKilobotsEnv
    text = """
!EvalEnv
width: 1.0
height: 1.0
resolution: 600
objects:
  - !ObjectConf {idx: 0, color: null, shape: c_shape, width: .15, height: .15, init: random, symmetry: null}
  - !ObjectConf {idx: 1, color: null, shape: triangle, width: .1, height: .1, init: random, symmetry: null}
light: !LightConf {type: momentum, init: object, radius: .2}
kilobots: !KilobotsConf {num: 12, mean: light, std: .03}
"""
    conf = yaml.load(text, Loader=yaml.Loader)
    np.random.seed(4)
    vec = KilobotsVecEnv.from_envs([YamlKilobotsEnv(configuration=conf) for _ in range(24)])
    sc = vec.scenario
    ob = oracle.OracleBatch(sc.scenes, sc.num_envs, sc.env_scene, sc.max_contacts, threads=8)
    vec.reset()
    ob.reset(sc.body_pose, sc.light_state)
    acts = SC.random_actions(sc, sc.num_envs, 10)
    for t in range(10):
        obs, r, d, info = vec.step(acts[t])
        out = ob.step(acts[t])
        for k in ("kilobots", "objects", "light"):
            assert np.array_equal(obs[k], out[k]), "step %d: %s" % (t, k)
    assert np.array_equal(vec.batch.bodies(), ob.bodies())
    # re-sample the scenes of a few envs the way a reference reset() would, and reset only those
    mask = np.zeros(sc.num_envs, np.uint8)
    mask[::5] = 1
    pose, light = vec.resample(mask)
    vec.reset_done(mask, pose, light)
    ob.reset(pose, light, mask=mask)
    obs, _, _, _ = vec.step(acts[0])
    out = ob.step(acts[0])
    assert np.array_equal(obs["kilobots"], out["kilobots"]) and np.array_equal(vec.batch.bodies(), ob.bodies())
