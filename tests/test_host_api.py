"""Host-side logic on CPU: the drop-in gym surface (reference class names / signatures), scene
templates, YAML tags, and that the C-ABI library loads and exports every declared symbol.
The E=1 facade is exercised with the ORACLE injected as backend (tests only; the product factory
`KilobotsEnv._make_batch` builds the CUDA batch and nothing else)."""
import ctypes
import os
import re

import numpy as np
import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "kb_b200.h")).read()
    declared = set(re.findall(r"\b(kb_[a-z_]+)\s*\(", hdr))
    assert {"kb_create", "kb_destroy", "kb_reset", "kb_step", "kb_step_host", "kb_last_error", "kb_get_state",
            "kb_set_state", "kb_get_contacts"} <= declared
    path = os.path.join(ROOT, "gym_kilobots_b200", "csrc", "libkb_b200.so")
    assert os.path.exists(path), "build with python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(path)
    for sym in declared:
        assert hasattr(lib, sym), sym
    from gym_kilobots_b200 import _native
    assert set(_native.exported_symbols()) <= declared


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gym_kilobots_b200 import _native, scenarios as SC
    sc = SC.c1_single_env(1)
    with pytest.raises(_native.NativeLibraryError):
        _native.NativeBatch(sc.scenes, 1)
    from gym_kilobots_b200.envs import QuadAssemblyKilobotsEnv
    env = QuadAssemblyKilobotsEnv(seed=0)
    with pytest.raises(_native.NativeLibraryError):
        env.reset()


def test_kb_create_without_device_reports_no_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gym_kilobots_b200 import _abi as abi, _native, scenarios as SC
    _, fn = _native.load()
    d, keep = SC.c1_single_env(1).scenes[0].to_desc()
    descs = (abi.KbSceneDesc * 1)(d)
    h = ctypes.c_void_p()
    rc = fn["create"](descs, 1, None, 1, 0, 0, ctypes.byref(h))
    assert rc == -2 and b"no CUDA device" in fn["last_error"]()


def _with_oracle(cls, oracle):
    class Injected(cls):
        def _make_batch(self, spec):
            return oracle.OracleBatch(spec, 1)
    return Injected


def test_quad_assembly_facade_matches_batched_oracle(oracle):
    from gym_kilobots_b200.envs import QuadAssemblyKilobotsEnv
    env = _with_oracle(QuadAssemblyKilobotsEnv, oracle)(seed=5)
    assert env.num_kilobots == 0            # reference drops constructor-pass kilobots (kilobots_env.py:67-68)
    obs = env.reset()
    assert obs["kilobots"].shape == (15, 3) and obs["objects"].shape == (4, 3) and obs["light"].shape == (2,)
    assert obs["kilobots"].dtype == np.float64
    assert env.action_space.shape == (2,)
    a = np.array([0.004, -0.02])
    o, r, d, info = env.step(a)
    assert r == 1.0 and d is False and info is None
    # same scene through the batched driver
    spec = env._scene_spec_cache
    ob = oracle.OracleBatch(spec, 1)
    env2 = _with_oracle(QuadAssemblyKilobotsEnv, oracle)(seed=5)
    env2.reset()
    pose = np.stack([b._init_pose for b in list(env2._objects) + list(env2._kilobots)])
    # (reset moved nothing observable in _init_pose)
    ob.reset(pose[None], env2._light._state_vector()[None] * 0 + pose[4, :2])
    out = ob.step(a[None])
    assert np.allclose(o["kilobots"], out["kilobots"][0].astype(np.float64), atol=1e-7)
    k0 = env.kilobots[0]
    assert np.allclose(k0.get_pose(), o["kilobots"][0])
    assert env.objects[0].vertices.shape == (1, 4, 2)
    assert np.allclose(env.get_light().get_state(), o["light"])


def test_yaml_env_tags_spaces_and_shapes(oracle):
    from gym_kilobots_b200.envs import YamlKilobotsEnv, EnvConfiguration
    text = """
!EvalEnv
width: 1.0
height: 1.0
resolution: 600
objects:
  - !ObjectConf {idx: 0, color: null, shape: l_shape, width: .15, height: .15, init: random, symmetry: null}
  - !ObjectConf {idx: 1, color: null, shape: circle, width: .05, height: .05, init: [.2, .2, 0.], symmetry: null}
light: !LightConf {type: momentum, init: object, radius: .2}
kilobots: !KilobotsConf {num: 7, mean: light, std: .03}
"""
    conf = yaml.load(text, Loader=yaml.Loader)
    assert isinstance(conf, EnvConfiguration) and conf.objects[0].object_type == "l_shape"
    env = _with_oracle(YamlKilobotsEnv, oracle)(configuration=conf)
    obs = env.reset()
    assert obs["kilobots"].shape == (7, 3) and obs["light"].shape == (4,)
    assert env.observation_space.shape == (7 * 2 + 4 + 2 * 4,)
    assert env.state_space.shape == (7 * 2 + 4 + 2 * 3,)
    assert env.world_width == 1.0 and env.world_bounds[1][0] == 0.5
    for _ in range(2):
        obs, r, d, info = env.step(env.action_space.sample())
    assert r == .0 and d is False and info == ""
    assert env.objects[1].get_radius() == .05   # sic: yaml 'circle' passes radius=width


def test_direct_control_env(oracle):
    from gym_kilobots_b200.envs import DirectControlKilobotsEnv
    from gym_kilobots_b200.lib import SimpleVelocityControlKilobot, Quad

    class Env(DirectControlKilobotsEnv):
        def _configure_environment(self):
            self._objects = [Quad(world=self.world, width=.1, height=.1, position=(.3, .0))]
            self._kilobots = [SimpleVelocityControlKilobot(self.world, position=(.05 * i, .0), orientation=.0,
                                                           velocity=[.005, .0]) for i in range(3)]

        def get_reward(self, *a):
            return 0.

    env = _with_oracle(Env, oracle)()
    env.reset()
    assert env.action_space.shape == (3, 2)
    x0 = env.get_state()["kilobots"][:, 0].copy()
    obs, r, d, info = env.step(np.array([[.01, 0.], [.01, 0.], [.0, 1.]]))
    assert obs["kilobots"][0, 0] > x0[0] + 0.005 and abs(obs["kilobots"][2, 2] - 1.0 / 1.08 * 1.0) < 0.2
    assert obs["light"].shape == (0,)       # reference would crash without a light (D7)


def test_body_accessors_match_reference_semantics(oracle):
    from gym_kilobots_b200.envs import TriangleTestEnv
    env = _with_oracle(TriangleTestEnv, oracle)()
    obs = env.reset()
    assert np.allclose(obs["objects"][:, :2], [[0, 0], [0, .3], [0, -.3], [.3, 0]], atol=2e-3)
    lform = env.objects[1]
    assert lform.vertices.shape == (2, 4, 2) and lform.width == .15
    p = lform.get_world_point((0.01, 0.02))
    assert np.allclose(lform.get_local_point(p), (0.01, 0.02), atol=1e-6)
    lform.set_pose((0.1, 0.2, 0.3))
    assert np.allclose(lform.get_pose(), (0.1, 0.2, 0.3), atol=1e-6)


def test_scenarios_are_partition_invariant():
    """Per-env draws depend on the global env id only: a rank's slice equals the full batch's slice."""
    from gym_kilobots_b200 import scenarios as SC
    full = SC.c2_quad_assembly(16, seed=3)
    part = SC.c2_quad_assembly(8, seed=3, env_offset=8)
    assert np.array_equal(full.body_pose[8:], part.body_pose)
    assert np.array_equal(full.light_state[8:], part.light_state)
    full3 = SC.c3_shapes(9, seed=1, num_kilobots=8)
    part3 = SC.c3_shapes(3, seed=1, env_offset=6, num_kilobots=8)
    assert np.array_equal(full3.body_pose[6:], part3.body_pose) and np.array_equal(full3.env_scene[6:], part3.env_scene)


_YAML_VEC = """
!EvalEnv
width: 1.0
height: 1.0
resolution: 600
objects:
  - !ObjectConf {idx: 0, color: null, shape: t_shape, width: .15, height: .15, init: random, symmetry: null}
  - !ObjectConf {idx: 1, color: null, shape: square, width: .1, height: .1, init: random, symmetry: null}
light: !LightConf {type: circular, init: object, radius: .2}
kilobots: !KilobotsConf {num: 9, mean: light, std: .03}
"""


def test_from_envs_vectorises_yaml_envs(oracle):
    """scenarios.from_envs / KilobotsVecEnv.from_envs: E YamlKilobotsEnv built from one
    configuration, each with its own random initialisation (yaml_kilobots_env.py:194-198,256-265,327-354),
    become ONE batch whose env e behaves exactly like stepping reference-style env e on its own."""
    from gym_kilobots_b200 import scenarios as SC
    from gym_kilobots_b200.envs import YamlKilobotsEnv
    conf = yaml.load(_YAML_VEC, Loader=yaml.Loader)
    E = 5
    cls = _with_oracle(YamlKilobotsEnv, oracle)
    np.random.seed(11)
    sc = SC.from_envs([cls(configuration=conf) for _ in range(E)])
    assert sc.body_pose.shape == (E, 11, 3) and sc.light_state.shape == (E, 2)
    assert len(sc.scenes) == 1 and np.array_equal(sc.env_scene, np.zeros(E, np.int32))
    assert len(np.unique(sc.body_pose[:, 0, 0])) == E          # every env drew its own object pose
    ob = oracle.OracleBatch(sc.scenes, E, sc.env_scene, sc.max_contacts)
    ob.reset(sc.body_pose, sc.light_state)
    acts = SC.random_actions(sc, E, 3)
    np.random.seed(11)
    singles = [cls(configuration=conf) for _ in range(E)]
    for env in singles:
        env.reset()
    for t in range(3):
        out = ob.step(acts[t])
        for e, env in enumerate(singles):
            o, r, d, info = env.step(acts[t, e])
            assert np.allclose(o["kilobots"], out["kilobots"][e], atol=1e-7)   # float32 observation vs float64 pose
            assert np.allclose(o["objects"], out["objects"][e], atol=1e-7)
            assert np.array_equal(o["light"], out["light"][e])


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """Every struct of include/kb_b200.h as gcc lays it out (sizeof and the offset of every field) against the ctypes
    mirror in gym_kilobots_b200/_abi.py: the boundary is plain C, a drifted field would corrupt scenes silently."""
    import subprocess
    from gym_kilobots_b200 import _abi as abi
    structs = ["KbFixtureDef", "KbBodyDef", "KbLightDef", "KbSceneDesc", "KbDims", "KbTaskDef", "KbLaunchConfig",
               "KbSampleObject", "KbSampleLight", "KbSampleSpec"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "kb_b200.h"', 'int main(void) {']
    for s in structs:
        cls = getattr(abi, s)
        lines.append('printf("%s %%zu", sizeof(%s));' % (s, s))
        for name, _ in cls._fields_:
            lines.append('printf(" %%zu", offsetof(%s, %s));' % (s, name))
        lines.append('printf("\\n");')
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).strip().splitlines()
    assert len(out) == len(structs)
    for line in out:
        parts = line.split()
        cls = getattr(abi, parts[0])
        assert int(parts[1]) == ctypes.sizeof(cls), parts[0]
        for (name, _), off in zip(cls._fields_, parts[2:]):
            assert int(off) == getattr(cls, name).offset, (parts[0], name)
