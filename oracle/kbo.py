"""ctypes driver of the CPU oracle (oracle/libkbo.so).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (no pybox2d available to pin it).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It mirrors the product's NativeBatch API with numpy host arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from gym_kilobots_b200 import _abi as abi

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force=False):
    """Compile the oracle with gcc (seconds)."""
    target = os.path.join(_DIR, "libkbo.so")
    if force or not os.path.exists(target):
        subprocess.check_call(["make", "-C", _DIR, "all"], stdout=subprocess.DEVNULL)
    return target


def load(libm_trig=False):
    name = "libkbo_libm.so" if libm_trig else "libkbo.so"
    if name not in _LIBS:
        path = os.path.join(_DIR, name)
        if not os.path.exists(path):
            build(force=True)
        lib = C.CDLL(path)
        fns = abi.bind(lib, "kbo_")
        lib.kbo_set_threads.restype = C.c_int
        lib.kbo_set_threads.argtypes = [C.c_void_p, C.c_int32]
        lib.kbo_sincosf.restype = None
        lib.kbo_sincosf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        lib.kbo_libm_sincosf.restype = None
        lib.kbo_libm_sincosf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        _LIBS[name] = (lib, fns)
    return _LIBS[name]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleBatch:
    """E independent reference-semantics environments stepped on the CPU."""

    def __init__(self, scenes, num_envs, env_scene=None, max_contacts=0, threads=1, libm_trig=False):
        self.lib, self.fn = load(libm_trig)
        if not isinstance(scenes, (list, tuple)):
            scenes = [scenes]
        self.scenes = list(scenes)
        descs = (abi.KbSceneDesc * len(scenes))()
        self._keep = []
        for i, s in enumerate(scenes):
            d, keep = s.to_desc()
            descs[i] = d
            self._keep.append(keep)
        es = None if env_scene is None else np.ascontiguousarray(env_scene, dtype=np.int32)
        h = C.c_void_p()
        rc = self.fn["create"](descs, len(scenes), _ptr(es), num_envs, max_contacts, 0, C.byref(h))
        if rc != 0:
            raise RuntimeError("kbo_create failed: %s" % self.fn["last_error"]().decode())
        self.h = h
        dims = abi.KbDims()
        self.fn["get_dims"](self.h, C.byref(dims))
        self.E, self.B, self.M, self.N = dims.num_envs, dims.num_bodies, dims.num_objects, dims.num_kilobots
        self.P, self.C, self.L, self.A = dims.num_proxies, dims.max_contacts, dims.light_state_dim, dims.action_dim
        self.set_threads(threads)

    def set_threads(self, n):
        self.lib.kbo_set_threads(self.h, int(n))

    def close(self):
        if self.h:
            self.fn["destroy"](self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, body_pose, light_state=None, kb_velocity=None, mask=None):
        pose = np.ascontiguousarray(body_pose, dtype=np.float64).reshape(self.E, self.B, 3)
        light = None if light_state is None else np.ascontiguousarray(light_state, dtype=np.float64).reshape(self.E, self.L)
        vel = None if kb_velocity is None else np.ascontiguousarray(kb_velocity, dtype=np.float64).reshape(self.E, self.N, 2)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        rc = self.fn["reset"](self.h, _ptr(m), _ptr(pose), _ptr(light), _ptr(vel), None)
        assert rc == 0, self.fn["last_error"]()

    def step(self, action=None, mode=None):
        if action is None:
            mode, act = abi.KB_ACTION_NONE, None
        else:
            if mode is None:
                mode = abi.KB_ACTION_LIGHT
            act = np.ascontiguousarray(action, dtype=np.float64)
        out = {
            "kilobots": np.zeros((self.E, self.N, 3), np.float32),
            "objects": np.zeros((self.E, self.M, 3), np.float32),
            "light": np.zeros((self.E, self.L), np.float64),
            "reward": np.zeros(self.E, np.float32),
            "done": np.zeros(self.E, np.uint8),
            "status": np.zeros(self.E, np.int32),
        }
        rc = self.fn["step"](self.h, _ptr(act), mode, _ptr(out["kilobots"]), _ptr(out["objects"]), _ptr(out["light"]),
                             _ptr(out["reward"]), _ptr(out["done"]), _ptr(out["status"]), None)
        assert rc == 0, self.fn["last_error"]()
        return out

    # ---- introspection -------------------------------------------------------------------
    def bodies(self):
        out = np.zeros((self.E, self.B, abi.KB_BODY_STATE_FLOATS), np.float32)
        self.fn["get_bodies"](self.h, _ptr(out))
        return out

    def set_poses(self, pose, body_mask=None):
        pose = np.ascontiguousarray(pose, dtype=np.float64).reshape(self.E, self.B, 3)
        m = None if body_mask is None else np.ascontiguousarray(body_mask, dtype=np.uint8).reshape(self.E, self.B)
        self.fn["set_poses_masked"](self.h, _ptr(pose), _ptr(m))

    def set_env_scene(self, env_scene):
        es = np.ascontiguousarray(env_scene, dtype=np.int32).reshape(self.E)
        rc = self.fn["set_env_scene"](self.h, _ptr(es))
        assert rc == 0

    def get_status(self):
        out = np.zeros(self.E, np.int32)
        self.fn["get_status"](self.h, _ptr(out))
        return out

    def contacts(self):
        pairs = np.zeros((self.E, self.C, 4), np.int32)
        count = np.zeros(self.E, np.int32)
        self.fn["get_contacts"](self.h, _ptr(pairs), _ptr(count))
        return pairs, count

    def impulses(self):
        out = np.zeros((self.E, self.C, 4), np.float32)
        self.fn["get_impulses"](self.h, _ptr(out))
        return out

    def counters(self):
        out = np.zeros((self.E, abi.KB_NUM_COUNTERS), np.uint64)
        self.fn["get_counters"](self.h, _ptr(out))
        return out

    def proxies(self):
        out = np.zeros((self.E, self.P, 4), np.float32)
        self.fn["get_proxies"](self.h, _ptr(out))
        return out

    def controllers(self):
        ctrl = np.zeros((self.E, self.N, 4), np.float64)
        light = np.zeros((self.E, max(self.L, 1)), np.float64)
        self.fn["get_controllers"](self.h, _ptr(ctrl), _ptr(light))
        return ctrl, light[:, :self.L]

    def mass_data(self):
        out = np.zeros((len(self.scenes), self.B, 4), np.float32)
        self.fn["get_mass_data"](self.h, _ptr(out))
        return out

    # ---- task layer (same extension as the product library, include/kb_b200.h "Task layer")
    def set_task(self, task, targets=None):
        t = task.to_def()
        tg = None if targets is None else np.ascontiguousarray(targets, dtype=np.float64).reshape(self.E, 3)
        rc = self.fn["set_task"](self.h, C.byref(t), _ptr(tg))
        assert rc == 0, "kbo_set_task rejected the task"

    def episode_stats(self):
        out = np.zeros((self.E, abi.KB_EPISODE_STATS), np.float64)
        self.fn["get_episode_stats"](self.h, _ptr(out))
        return out

    def render(self, env_ids=(0,), width=1200, height=900):
        ids = np.ascontiguousarray(env_ids, dtype=np.int32).reshape(-1)
        out = np.zeros((len(ids), height, width, 3), np.uint8)
        self.fn["render"](self.h, _ptr(ids), len(ids), width, height, _ptr(out), None)
        return out

    def bind_flat_observation(self, enable=True):
        if not enable:
            self.obs_flat = None
            self.fn["bind_flat_observation"](self.h, None)
            return None
        self.obs_flat = np.zeros((self.E, self.fn["flat_observation_dim"](self.h)), np.float32)
        self.fn["bind_flat_observation"](self.h, _ptr(self.obs_flat))
        return self.obs_flat


def sincosf(a, libm=False):
    lib, _ = load()
    a = np.ascontiguousarray(a, dtype=np.float32)
    s = np.empty_like(a)
    c = np.empty_like(a)
    (lib.kbo_libm_sincosf if libm else lib.kbo_sincosf)(_ptr(a), _ptr(s), _ptr(c), a.size)
    return s, c


# ---- golden-vector hooks (single pieces of the Python-side restatement) ---------------------------
def _light_defs(specs):
    from gym_kilobots_b200.scene import SceneSpec
    d, keep = SceneSpec(lights=list(specs)).to_desc()
    return d.lights, keep


def eval_light(specs, state, pts):
    lib, _ = load()
    defs, keep = _light_defs(specs)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    state = np.ascontiguousarray(state, dtype=np.float64)
    value = np.zeros(len(pts))
    grad = np.zeros((len(pts), 2))
    lib.kbo_eval_light.restype = None
    lib.kbo_eval_light.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.kbo_eval_light(C.cast(defs, C.c_void_p), len(specs), _ptr(state), _ptr(pts), len(pts), _ptr(value), _ptr(grad))
    return value, grad


def light_step(spec, state, action, substeps):
    lib, _ = load()
    defs, keep = _light_defs([spec])
    state = np.ascontiguousarray(state, dtype=np.float64).copy()
    action = np.ascontiguousarray(action, dtype=np.float64)
    trace = np.zeros((substeps, len(state)))
    lib.kbo_light_step.restype = None
    lib.kbo_light_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.kbo_light_step(C.cast(defs, C.c_void_p), _ptr(state), _ptr(action), substeps, _ptr(trace))
    return state, trace


def controller_trace(kind, pose_b2, feed, velocity=None):
    lib, _ = load()
    feed = np.ascontiguousarray(feed, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((len(feed), 3), np.float32)
    vel = None if velocity is None else np.ascontiguousarray(velocity, dtype=np.float64)
    lib.kbo_controller_trace.restype = None
    lib.kbo_controller_trace.argtypes = [C.c_int32, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p]
    lib.kbo_controller_trace(kind, pose_b2[0], pose_b2[1], pose_b2[2], _ptr(vel), _ptr(feed), len(feed), _ptr(out))
    return out


def sensor_pos(x, y, angle):
    lib, _ = load()
    out = np.zeros(2)
    lib.kbo_sensor_pos.restype = None
    lib.kbo_sensor_pos.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p]
    lib.kbo_sensor_pos(x, y, angle, _ptr(out))
    return out
