"""YamlSceneSampler: the vectorised, device-side restatement of YamlKilobotsEnv's random scene
initialisation (yaml_kilobots_env.py:194-198,256-283,327-354).  The reference draws from numpy's unseeded global
generator, so agreement is distributional: same supports and the same moments as the E = 1 facade, whose
`_configure_environment` mirrors the reference line by line (tests/test_host_api.py)."""
import numpy as np
import pytest
import yaml

from gym_kilobots_b200.envs import YamlKilobotsEnv
from gym_kilobots_b200.envs.yaml_sampler import YamlSceneSampler

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
    ps, ls = YamlSceneSampler(conf, n, device="cpu", seed=1).sample()
    ps, ls = ps.numpy(), ls.numpy()
    assert ps.shape == pf.shape == (n, 8, 3) and ls.shape == lf.shape == (n, 4)
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


def test_sampler_composite_and_random_mean():
    text = CONF.replace("light: !LightConf {type: momentum, init: object, radius: .2}",
                        "light: !LightConf {type: composite, init: null, components: ["
                        "{type: circular, init: random, radius: .2}, {type: circular, init: [.1, .1], radius: .1}]}")
    conf = yaml.load(text, Loader=yaml.Loader)
    s = YamlSceneSampler(conf, 2000, device="cpu", seed=3)
    pose, light = s.sample()
    pose, light = pose.numpy(), light.numpy()
    assert light.shape == (2000, 4) and np.array_equal(light[:, 2:], np.tile([.1, .1], (2000, 1)))
    # every kilobot is drawn around one of the two component lights (yaml_kilobots_env.py:335-338)
    d0 = np.hypot(*(pose[:, 2:, :2] - light[:, None, 0:2]).transpose(2, 0, 1))
    d1 = np.hypot(*(pose[:, 2:, :2] - light[:, None, 2:4]).transpose(2, 0, 1))
    near = np.minimum(d0, d1)
    assert np.mean(near < 0.15) > 0.99 and 0.3 < np.mean(d1 < d0) < 0.7
    conf.kilobots.mean = "random"
    pose2, _ = YamlSceneSampler(conf, 2000, device="cpu", seed=4).sample()
    c = pose2.numpy()[:, 2:, :2].mean(1)
    assert np.abs(c[:, 0]).max() <= 0.9 * 0.6 + 0.1 and c[:, 0].std() > 0.2


@pytest.mark.gpu
def test_device_sampler_auto_reset(native):
    from gym_kilobots_b200.envs import KilobotsVecEnv
    conf = yaml.load(CONF, Loader=yaml.Loader)
    np.random.seed(2)
    vec = KilobotsVecEnv.from_envs([YamlKilobotsEnv(configuration=conf) for _ in range(64)])
    vec.use_device_sampler(conf, seed=5)
    vec.reset()
    before = vec.batch.bodies().copy()
    import torch
    done = torch.zeros(64, dtype=torch.uint8, device="cuda")
    done[::4] = 1
    vec.reset_done(done)
    after = vec.batch.bodies()
    changed = (before[:, :, 8:10] != after[:, :, 8:10]).any(axis=(1, 2))
    assert np.array_equal(changed, done.cpu().numpy().astype(bool))       # only the finished envs were re-drawn
    # the object with a fixed init is back at its pose (up to the push of bodies spawned on top of it, resolved by
    # the settle step of reset, kilobots_env.py:157)
    assert np.abs(after[::4, 1, 8:10] - np.float32([.2 * 25, -.1 * 25])).max() < 1.0
    assert (after[::4, 1, 8:10] == np.float32([.2 * 25, -.1 * 25])).all(1).any()
