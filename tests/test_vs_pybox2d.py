"""Oracle vs the REAL reference (gregorgebhardt/gym-kilobots on pybox2d) -- SURVEY.md 8(c) plan item (3).

INERT IN THE BUILD CONTAINER AND ON THE GPU BOX: `import Box2D` fails there (box2d-py is not installable
offline), so every test here skips and the Box2D half of the oracle stays "parity unpinned" (DESIGN.md section 2).
The file exists so that the pin happens automatically wherever box2d-py is installed: it needs `Box2D`, the
reference sources (KB_REFERENCE_DIR, default /root/reference, or an installed `gym_kilobots`) and, if `gym` is
missing, falls back to a three-line stand-in for `gym.Env` / `gym.spaces.Box` (the reference only stores them).

What is compared (BASELINE.json north_star): over a stated short horizon on identical initial states and actions,
body poses within 1e-4 m / 1e-4 rad and bit-identical contact-pair sets; then the Box2D-version switches of
SURVEY Appendix B.9 (damping mode, open/closed table chain) are calibrated by which setting reproduces pybox2d.
"""
import os
import sys
import types

import numpy as np
import pytest

Box2D = pytest.importorskip("Box2D", reason="box2d-py (pybox2d) is not installed: the Box2D half stays unpinned")

HORIZON_ENV_STEPS = 5          # 50 physics sub-steps: the "stated short horizon" of the parity claim
POSE_TOL_M = 1e-4
POSE_TOL_RAD = 1e-4


def _reference():
    try:
        import gym  # noqa: F401
    except Exception:
        gym = types.ModuleType("gym")
        spaces = types.ModuleType("gym.spaces")

        class Box:
            def __init__(self, low, high, dtype=None, shape=None):
                self.low, self.high, self.dtype = np.asarray(low), np.asarray(high), dtype
                self.shape = self.low.shape

            def sample(self):
                return np.random.uniform(self.low, self.high)

        class Env:
            pass

        spaces.Box = Box
        gym.spaces = spaces
        gym.Env = Env
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    ref_dir = os.environ.get("KB_REFERENCE_DIR", "/root/reference")
    if os.path.isdir(os.path.join(ref_dir, "gym_kilobots")) and ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import gym_kilobots.envs.kilobots_env as ref_env
        import gym_kilobots.lib as ref_lib
    except Exception as e:  # pragma: no cover
        pytest.skip("reference sources not importable: %r" % (e,))
    return ref_env, ref_lib


def _make_reference_env(ref_env, ref_lib, object_pose, kilobot_xy, light_xy):
    class Env(ref_env.KilobotsEnv):
        def _configure_environment(self):
            self._light = ref_lib.CircularGradientLight(position=np.array(light_xy, float), radius=.2,
                                                        bounds=np.array(self.world_bounds) * 1.1,
                                                        action_bounds=(np.array([-.01, -.01]), np.array([.01, .01])))
            self._add_object(ref_lib.Quad(width=.15, height=.15, position=object_pose[:2],
                                          orientation=object_pose[2], world=self.world))
            for p in kilobot_xy:
                self._add_kilobot(ref_lib.PhototaxisKilobot(self.world, position=p, light=self._light))

        def get_reward(self, *a):
            return 0.

    return Env()


def _oracle_for(oracle, object_pose, kilobot_xy, light_xy, **scene_kw):
    from gym_kilobots_b200 import _abi as abi, scene as S, scenarios as SC
    bodies = [S.quad_body(.15, .15)] + [S.kilobot_body(abi.KB_KILOBOT_PHOTOTAXIS) for _ in kilobot_xy]
    spec = S.SceneSpec(bodies=bodies, num_objects=1, lights=[SC._circular_light(.2, (2.0, 1.5))], **scene_kw)
    pose = np.zeros((1, len(bodies), 3))
    pose[0, 0] = object_pose
    pose[0, 1:, :2] = kilobot_xy
    ob = oracle.OracleBatch(spec, 1, libm_trig=True)      # libm sinf/cosf: what Box2D itself calls
    ob.reset(pose, np.asarray(light_xy, float)[None])
    return ob


def _scene(seed):
    rng = np.random.default_rng(seed)
    light = rng.uniform(-.5, .5, 2)
    kil = []
    while len(kil) < 10:
        p = light + rng.normal(scale=.05, size=2)
        if all(np.hypot(*(p - q)) > .04 for q in kil):
            kil.append(p)
    return np.array([light[0] + .12, light[1], .3]), np.array(kil), light


def _pair_set_reference(env):
    pairs = set()
    # SWIG hands out a fresh proxy object per access, so bodies are tagged through userData (the table has none)
    for i, b in enumerate(list(env.get_objects()) + list(env.get_kilobots())):
        b._body.userData = i
    for c in env.world.contacts:
        a, b = c.fixtureA.body.userData, c.fixtureB.body.userData
        a, b = (-1 if a is None else int(a)), (-1 if b is None else int(b))
        pairs.add((min(a, b), max(a, b), bool(c.touching)))
    return pairs


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_short_horizon_poses_and_contact_pairs(oracle, seed):
    ref_env, ref_lib = _reference()
    obj, kil, light = _scene(seed)
    env = _make_reference_env(ref_env, ref_lib, obj, kil, light)
    env.reset()
    ob = _oracle_for(oracle, obj, kil, light)
    rng = np.random.default_rng(100 + seed)
    for t in range(HORIZON_ENV_STEPS):
        a = rng.uniform(-.01, .01, 2)
        o_ref, _, _, _ = env.step(a)
        o = ob.step(a[None])
        assert np.abs(o_ref["kilobots"][:, :2] - o["kilobots"][0, :, :2]).max() <= POSE_TOL_M, "env-step %d" % t
        assert np.abs(o_ref["kilobots"][:, 2] - o["kilobots"][0, :, 2]).max() <= POSE_TOL_RAD, "env-step %d" % t
        assert np.abs(o_ref["objects"][:, :2] - o["objects"][0, :, :2]).max() <= POSE_TOL_M, "env-step %d" % t
        assert np.abs(o_ref["objects"][:, 2] - o["objects"][0, :, 2]).max() <= POSE_TOL_RAD, "env-step %d" % t
    # contact-pair sets (body index pairs + touching flag; walls = -1)
    pairs, count = ob.contacts()
    P_wall = ob.P - (1 + len(kil))
    mine = set()
    for k in range(int(count[0])):
        pa, pb, touching, _ = pairs[0, k]
        a = -1 if pa < P_wall else int(pa - P_wall)
        b = -1 if pb < P_wall else int(pb - P_wall)
        mine.add((min(a, b), max(a, b), bool(touching)))
    assert mine == _pair_set_reference(env)


def test_version_switch_calibration(oracle):
    """exactly one (damping_mode, wall_edges) setting should reproduce the installed Box2D build."""
    ref_env, ref_lib = _reference()
    obj, kil, light = _scene(7)
    kil[0] = (0.98, 0.0)                       # one kilobot near the +x wall: the open/closed chain matters there
    env = _make_reference_env(ref_env, ref_lib, obj, kil, light)
    env.reset()
    ref = [env.step(np.zeros(2))[0] for _ in range(3)][-1]
    errs = {}
    for damping_mode in (0, 1):
        for wall_edges in (3, 4):
            ob = _oracle_for(oracle, obj, kil, light, damping_mode=damping_mode, wall_edges=wall_edges)
            for _ in range(3):
                o = ob.step(np.zeros((1, 2)))
            errs[(damping_mode, wall_edges)] = float(np.abs(ref["kilobots"] - o["kilobots"][0]).max())
    best = min(errs, key=errs.get)
    assert errs[best] <= POSE_TOL_M, errs
    assert best == (0, 3), "installed Box2D behaves like damping_mode=%d wall_edges=%d; update the defaults: %r" % (
        best[0], best[1], errs)
