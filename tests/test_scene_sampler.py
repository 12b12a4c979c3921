"""Scene sampling at reset: YamlKilobotsEnv's random scene initialisation
(yaml_kilobots_env.py:194-198,256-283,299,327-354) as a counter-based function of (seed, global env id, episode),
evaluated inside the reset kernel (csrc/kb_sample.cuh) and, draw for draw, in numpy (gym_kilobots_b200/sampler.py).
The reference draws from numpy's unseeded global generator, so agreement with it is distributional (same supports and
moments as the E = 1 facade, whose `_configure_environment` mirrors the reference line by line); device vs numpy is
exact up to the last ulp of log / sin / cos; the physics after a sampled reset is bit-exact vs the oracle."""
import numpy as np
import pytest
import yaml

from gym_kilobots_b200 import _abi as abi
from gym_kilobots_b200 import scenarios as SC
from gym_kilobots_b200.envs import YamlKilobotsEnv
from gym_kilobots_b200.sampler import SceneSampler

CONF = """
!EvalEnv
width: 1.2
height: 0.8
resolution: 600
objects:
  - !ObjectConf {idx: 0, color: null, shape: square, width: .15, height: .10, init: random, symmetry: null}
  - !ObjectConf {idx: 1, color: null, shape: triangle, width: .1, height: .1, init: [.2, -.1, .5], symmetry: null}
light: !LightConf {type: momentum, init: object, radius: .2}
kilobots: !KilobotsConf {num: 6, mean: light, std: .03}
"""
COMPOSITE = ("light: !LightConf {type: composite, init: random, components: ["
             "!LightConf {type: circular, init: random, radius: .2}, !LightConf {type: momentum, init: [.1, .1], radius: .1}]}")


def _facade_draws(conf, n):
    env = YamlKilobotsEnv(configuration=conf)
    poses, lights = [], []
    for _ in range(n):
        _, pose, light, _ = env._record_scene()
        poses.append(pose)
        lights.append(light)
    return np.stack(poses), np.stack(lights)


def test_sampler_matches_the_facade_in_distribution():
    conf = yaml.load(CONF, Loader=yaml.Loader)
    n = 4000
    np.random.seed(0)
    pf, lf = _facade_draws(conf, n)
    ps, ls, scene = SceneSampler(conf, seed=1).sample_numpy(np.arange(n), 0)
    assert ps.shape == pf.shape == (n, 8, 3) and ls.shape == lf.shape == (n, 4) and not scene.any()
    # supports
    assert np.abs(ps[:, 0, 0]).max() <= 0.7 * 0.6 and np.abs(ps[:, 0, 1]).max() <= 0.7 * 0.4   # U(world) * 0.7
    assert np.abs(ps[:, 0, 2]).max() <= np.pi
    assert np.array_equal(ps[:, 1], np.tile([.2, -.1, .5], (n, 1)))                              # fixed init
    assert np.abs(ps[:, 2:, 0]).max() <= 0.6 - 0.02 + 1e-12 and np.abs(ps[:, 2:, 1]).max() <= 0.4 - 0.02 + 1e-12
    assert np.allclose(np.hypot(ls[:, 2], ls[:, 3]), .01)                                        # momentum light speed
    # the light sits on a circle of radius 1.2 * max(w, h) / 2 around one of the two objects
    d = np.stack([np.hypot(*(ls[:, :2] - ps[:, i, :2]).T) for i in range(2)], 1)
    r = np.array([1.2 * .15 / 2, 1.2 * .1 / 2])
    assert np.all(np.isclose(d, r, atol=1e-12).any(1))
    # moments agree with the facade (z-test at 6 standard errors on every column)
    def close(a, b):
        se = np.sqrt(a.var(0) / len(a) + b.var(0) / len(b)) + 1e-12
        return np.all(np.abs(a.mean(0) - b.mean(0)) <= 6 * se) and np.all(np.abs(a.std(0) - b.std(0)) <= 0.1 * a.std(0) + 1e-9)
    assert close(pf[:, 0], ps[:, 0])                      # random object pose
    assert close(lf, ls)                                  # light position + velocity
    rel_f = (pf[:, 2:, :2] - lf[:, None, :2]).reshape(-1, 2)
    rel_s = (ps[:, 2:, :2] - ls[:, None, :2]).reshape(-1, 2)
    inside = lambda p: (np.abs(p[:, :, 0]) < 0.57) & (np.abs(p[:, :, 1]) < 0.37)   # away from the clip
    mf, ms = inside(pf[:, 2:]).reshape(-1), inside(ps[:, 2:]).reshape(-1)
    assert close(rel_f[mf], rel_s[ms])                    # kilobots ~ N(light, 0.03^2)


def test_draws_are_keyed_by_global_env_id_and_episode():
    conf = yaml.load(CONF, Loader=yaml.Loader)
    whole = SceneSampler(conf, seed=7).sample_numpy(np.arange(64), 3)
    part = SceneSampler(conf, seed=7, env_id_base=32).sample_numpy(np.arange(16), 3)      # "rank 2 of 4"
    assert np.array_equal(whole[0][32:48], part[0]) and np.array_equal(whole[1][32:48], part[1])
    other = SceneSampler(conf, seed=7).sample_numpy(np.arange(64), 4)
    assert not np.array_equal(whole[0][:, 0], other[0][:, 0])                              # a new episode, new draws
    eps = np.arange(64) % 5
    mixed = SceneSampler(conf, seed=7).sample_numpy(np.arange(64), eps)
    assert np.array_equal(mixed[0][eps == 3], whole[0][eps == 3])


def test_shuffled_composite_light_and_random_mean():
    """yaml_kilobots_env.py:299: a composite light with init 'random' shuffles its components -- the ORDER of the
    light types in the state / action vectors changes per episode; every order is a scene template."""
    conf = yaml.load(CONF.replace("light: !LightConf {type: momentum, init: object, radius: .2}", COMPOSITE), Loader=yaml.Loader)
    s = SceneSampler(conf, seed=3)
    specs = s.scene_specs()
    assert len(specs) == 2 and s.shuffle
    assert [l.type for l in specs[0].lights] == [abi.KB_LIGHT_CIRCULAR, abi.KB_LIGHT_MOMENTUM]
    assert [l.type for l in specs[1].lights] == [abi.KB_LIGHT_MOMENTUM, abi.KB_LIGHT_CIRCULAR]
    assert [l.radius for l in specs[1].lights] == [.1, .2]
    pose, light, scene = s.sample_numpy(np.arange(4000), 0)
    assert light.shape == (4000, 6) and 0.45 < scene.mean() < 0.55                   # both orders, evenly
    a, b = scene == 0, scene == 1
    assert np.allclose(light[a, 2:4], [.1, .1]) and np.allclose(np.hypot(light[a, 4], light[a, 5]), .01)   # circ | mom
    assert np.allclose(light[b, 0:2], [.1, .1]) and np.allclose(np.hypot(light[b, 2], light[b, 3]), .01)   # mom | circ
    # every kilobot is drawn around one of the two component lights (:335-338)
    pc = np.where(a[:, None], light[:, 0:2], light[:, 4:6])      # the circular component's position
    pm = np.tile([.1, .1], (4000, 1))
    d0 = np.hypot(*(pose[:, 2:, :2] - pc[:, None]).transpose(2, 0, 1))
    d1 = np.hypot(*(pose[:, 2:, :2] - pm[:, None]).transpose(2, 0, 1))
    assert np.mean(np.minimum(d0, d1) < 0.15) > 0.99 and 0.3 < np.mean(d1 < d0) < 0.7
    conf.kilobots.mean = "random"
    pose2 = SceneSampler(conf, seed=4).sample_numpy(np.arange(2000), 0)[0]
    c = pose2[:, 2:, :2].mean(1)
    assert np.abs(c[:, 0]).max() <= 0.9 * 0.6 + 0.1 and c[:, 0].std() > 0.2


@pytest.mark.gpu
@pytest.mark.parametrize("composite", [False, True])
def test_device_sampler_vs_numpy_and_oracle(oracle, native, composite):
    """(1) what the reset kernel draws equals the numpy evaluation (|diff| <= 1e-12: libm vs CUDA log / sin / cos);
    (2) masked auto-reset re-draws exactly the finished envs, with the next episode's numbers; (3) the oracle, reset
    with the device's draws (and scene switches), then follows the CUDA path bit for bit."""
    from gym_kilobots_b200.envs import KilobotsVecEnv
    import torch
    text = CONF.replace("light: !LightConf {type: momentum, init: object, radius: .2}", COMPOSITE) if composite else CONF
    conf = yaml.load(text, Loader=yaml.Loader)
    E = 96
    vec = KilobotsVecEnv.from_configuration(conf, E, seed=11, env_id_base=1000)
    sm, sc = vec.sampler, vec.scenario
    ob = oracle.OracleBatch(sc.scenes, E, sc.env_scene, sc.max_contacts, threads=8)
    vec.reset()
    pose, light, scene, ep = vec.batch.get_sampled()
    ref = sm.sample_numpy(np.arange(E), 0)
    assert np.abs(pose - ref[0]).max() <= 1e-12 and np.abs(light - ref[1]).max() <= 1e-12
    assert np.array_equal(scene, ref[2]) and np.all(ep == 1)
    if composite:
        assert set(scene) == {0, 1}
    ob.set_env_scene(scene)
    ob.reset(pose, light)
    assert np.array_equal(vec.batch.bodies(), ob.bodies())
    acts = SC.random_actions(sc, E, 6)
    for t in range(3):
        o1 = vec.step(acts[t])[0]
        o2 = ob.step(acts[t])
        for k in ("kilobots", "objects", "light"):
            assert np.array_equal(o1[k], o2[k]), (t, k)
    before = vec.batch.bodies().copy()
    done = torch.zeros(E, dtype=torch.uint8, device="cuda")
    done[::4] = 1
    vec.reset_done(done)
    m = done.cpu().numpy().astype(bool)
    pose2, light2, scene2, ep2 = vec.batch.get_sampled()
    assert np.array_equal(ep2, 1 + m.astype(np.uint32))
    ref2 = sm.sample_numpy(np.arange(E), 1)
    assert np.abs(pose2[m] - ref2[0][m]).max() <= 1e-12 and np.array_equal(scene2[m], ref2[2][m])
    assert np.array_equal(pose2[~m], pose[~m])                                   # the others keep their scene
    after = vec.batch.bodies()
    changed = (before[:, :, 8:10] != after[:, :, 8:10]).any(axis=(1, 2))
    assert np.array_equal(changed, m)
    ob.set_env_scene(scene2)
    ob.reset(pose2, light2, mask=m.astype(np.uint8))
    for t in range(3, 6):
        o1 = vec.step(acts[t])[0]
        o2 = ob.step(acts[t])
        for k in ("kilobots", "objects", "light"):
            assert np.array_equal(o1[k], o2[k]), (t, k)
    assert np.array_equal(vec.batch.bodies(), ob.bodies())
    # rank invariance on the device: a batch that holds only "the second half of the job" draws the same scenes
    half = KilobotsVecEnv.from_configuration(conf, E // 2, seed=11, env_id_base=1000 + E // 2)
    half.reset()
    ph, lh, sh, _ = half.batch.get_sampled()
    assert np.array_equal(ph, pose[E // 2:]) and np.array_equal(lh, light[E // 2:]) and np.array_equal(sh, scene[E // 2:])
    vec.close()
    half.close()


@pytest.mark.gpu
def test_device_sampler_on_the_swarm_tier(oracle, native):
    """80 kilobots and no object select the large-swarm tier; its reset kernel draws the scene the same way."""
    from gym_kilobots_b200.envs import KilobotsVecEnv
    import torch
    text = """
!EvalEnv
width: 1.0
height: 0.8
resolution: 600
objects: []
light: !LightConf {type: circular, init: random, radius: .3}
kilobots: !KilobotsConf {num: 80, mean: [0.1, -0.05], std: 0.08}
"""
    # (mean 'light' with a light at the table's edge stacks dozens of kilobots on the clip line: more touching contacts
    #  than the tier's 3B + 32 solver capacity -- flagged, and a different test)
    conf = yaml.load(text, Loader=yaml.Loader)
    E = 12
    vec = KilobotsVecEnv.from_configuration(conf, E, seed=5, env_id_base=40)
    assert vec.batch.launch_config()["block_threads"] in (128, 256, 512)   # the swarm tier
    ob = oracle.OracleBatch(vec.scenario.scenes, E, vec.scenario.env_scene, vec.scenario.max_contacts, threads=8)
    vec.reset()
    pose, light, scene, ep = vec.batch.get_sampled()
    ref = vec.sampler.sample_numpy(np.arange(E), 0)
    assert np.abs(pose - ref[0]).max() <= 1e-12 and np.abs(light - ref[1]).max() <= 1e-12 and np.all(ep == 1)
    ob.reset(pose, light)
    assert np.array_equal(vec.batch.bodies(), ob.bodies())
    acts = SC.random_actions(vec.scenario, E, 4)
    for t in range(2):
        o1, o2 = vec.step(acts[t])[0], ob.step(acts[t])
        assert np.array_equal(o1["kilobots"], o2["kilobots"]) and np.array_equal(o1["light"], o2["light"])
    done = torch.zeros(E, dtype=torch.uint8, device="cuda")
    done[1::3] = 1
    vec.reset_done(done)
    m = done.cpu().numpy().astype(bool)
    pose2, light2, _, ep2 = vec.batch.get_sampled()
    assert np.array_equal(ep2, 1 + m.astype(np.uint32)) and np.array_equal(pose2[~m], pose[~m])
    ob.reset(pose2, light2, mask=m.astype(np.uint8))
    for t in range(2, 4):
        o1, o2 = vec.step(acts[t])[0], ob.step(acts[t])
        assert np.array_equal(o1["kilobots"], o2["kilobots"])
    assert np.array_equal(vec.batch.bodies(), ob.bodies())
    assert not vec.check_status().any()
    vec.close()
